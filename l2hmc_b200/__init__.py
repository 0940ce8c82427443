"""l2hmc_b200 -- B200-native (sm_100a) implementation of the batched leapfrog
integrator hot path of saforem2/l2hmc-qcd, behind the reference's own
Dynamics / Lattice / Group / NetworkFactory surface.

Module paths mirror the reference package (`l2hmc.<...>`):

    l2hmc_b200.configs
    l2hmc_b200.group.{su3,u1}.pytorch.group
    l2hmc_b200.lattice.{su3,u1}.pytorch.lattice
    l2hmc_b200.network.pytorch.network
    l2hmc_b200.dynamics.pytorch.dynamics

`l2hmc_b200.ops` holds the tensor-level wrappers of the C ABI (include/l2b.h);
`l2hmc_b200._lib` is the ctypes binding of `libl2b.so`.  There is no CPU
fallback: the CUDA library must be built (`python -m l2hmc_b200._build`).
"""
__version__ = '0.1.0'
