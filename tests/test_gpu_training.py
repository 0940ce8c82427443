"""GPU tier (-m gpu): back-propagation through the CUDA path.  The same scalar
the golden generator differentiated through the REFERENCE's autograd graph
(oracle/make_golden.py: a loss touching x_prop, acc, sumlogdet and the Wilson
loops) is differentiated through our hand-written adjoint kernels; gradients of
every network parameter, of the step sizes and of the input links must agree."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


@pytest.fixture()
def default_dtype():
    old = torch.get_default_dtype()
    yield torch.set_default_dtype
    torch.set_default_dtype(old)


def _build_u1(gu, name, tag):
    from l2hmc_b200.configs import DynamicsConfig, NetworkConfig, ConvolutionConfig, get_input_spec
    from l2hmc_b200.dynamics.pytorch.dynamics import Dynamics
    from l2hmc_b200.lattice.u1.pytorch.lattice import LatticeU1
    from l2hmc_b200.network.pytorch.network import NetworkFactory
    shape, nb, nlf = [int(s) for s in gu['shape']], 3, 2
    pre = f'{name}/'
    cfg = DynamicsConfig(nchains=nb, group='U1', latvolume=shape, nleapfrog=nlf, eps=0.1, eps_hmc=0.1, use_ncp=True,
                         verbose=False, use_split_xnets=True, use_separate_networks=True, merge_directions=True)
    conv = ConvolutionConfig(filters=[4, 8, 8], sizes=[3, 2, 2], pool=[2, 2, 2]) if name == 'conv' else None
    fac = NetworkFactory(input_spec=get_input_spec(cfg),
                         network_config=NetworkConfig(units=[16, 12], activation_fn='leaky_relu', dropout_prob=0.2,
                                                      use_batch_norm=True),
                         conv_config=conv, net_weights=None)
    lat = LatticeU1(nb, shape)
    dyn = Dynamics(potential_fn=lat.action, config=cfg, network_factory=fac)
    sd = {k[len(pre) + 3:]: torch.from_numpy(gu[k]) for k in gu.files if k.startswith(pre + 'sd/')}
    dyn.load_state_dict(sd, strict=False)
    dyn.masks = [dev(m) for m in gu[pre + 'masks']]
    dyn.eval()
    return dyn, lat


@pytest.mark.parametrize('tag,rtol', [('f64', 1e-9), ('f32', 5e-3)])
@pytest.mark.parametrize('name', ['dense', 'conv'])
def test_u1_l2hmc_gradients_match_reference_autograd(golden_dir, default_dtype, tag, rtol, name):
    default_dtype(torch.float64 if tag == 'f64' else torch.float32)
    from l2hmc_b200.dynamics.pytorch.dynamics import State
    gu = np.load(golden_dir / f'u1_{tag}.npz')
    pre = f'{name}/'
    dyn, lat = _build_u1(gu, name, tag)
    x = dev(gu['x']).requires_grad_(True)
    st = State(x, dev(gu[pre + 'v']), torch.tensor(float(gu['beta'])))
    sp, met = dyn.transition_kernel_fb(st)
    xp = sp.x.flatten(1)
    loss = ((met['acc'] * xp.cos().sum(1)).sum() + met['sumlogdet'].sum()
            + (met['acc'] * lat.wilson_loops(sp.x).sin().sum((1, 2))).sum())
    want_loss = float(gu[pre + 'loss'])
    assert abs(float(loss) - want_loss) <= rtol * max(1.0, abs(want_loss)) * 10
    loss.backward()
    checked = 0
    params = dict(dyn.named_parameters())
    for k in gu.files:
        if not k.startswith(pre + 'grad/'):
            continue
        n = k[len(pre) + 5:]
        p = params.get('networks.' + n, params.get(n))
        assert p is not None and p.grad is not None, f'no gradient for {n}'
        want = gu[k]
        got = p.grad.detach().cpu().numpy()
        scale = max(1e-6 if tag == 'f64' else 1e-2, float(np.abs(want).max()))
        assert np.max(np.abs(got - want)) <= rtol * scale, (n, float(np.max(np.abs(got - want))), scale)
        checked += 1
    assert checked >= 100
    gx = x.grad.detach().cpu().numpy()
    assert np.max(np.abs(gx - gu[pre + 'grad_x'])) <= rtol * max(1.0, float(np.abs(gu[pre + 'grad_x']).max()))


def test_u1_adjoint_kernels_vs_finite_differences():
    """each adjoint kernel against central differences of its own forward kernel (fp64)"""
    from l2hmc_b200 import autograd as ag
    torch.manual_seed(0)
    nb, T, X = 2, 6, 5
    xdim = 2 * T * X
    x = ((torch.rand(nb, 2, T, X, dtype=torch.float64, device=DEV) * 2 - 1) * 3).requires_grad_(True)
    v = torch.randn(nb, xdim, dtype=torch.float64, device=DEV, requires_grad=True)
    s, t, q = (0.3 * torch.randn(nb, xdim, dtype=torch.float64, device=DEV, requires_grad=True) for _ in range(3))
    eps = torch.tensor(0.09, dtype=torch.float64, device=DEV, requires_grad=True)
    mask = (torch.rand(xdim, device=DEV) > 0.5).float()
    wts = [torch.randn(nb, xdim, dtype=torch.float64, device=DEV), torch.randn(nb, dtype=torch.float64, device=DEV)]

    def fd_check(fn, inputs, tol=1e-6):
        outs = fn(*inputs)
        outs = outs if isinstance(outs, tuple) else (outs,)
        loss = sum((o.reshape(nb, -1) * w.reshape(nb, -1)[:, :o.reshape(nb, -1).shape[1]]).sum() for o, w in zip(outs, wts))
        grads = torch.autograd.grad(loss, inputs, allow_unused=True)
        h = 1e-6
        for inp, g in zip(inputs, grads):
            flat = inp.detach().reshape(-1)
            for idx in torch.randint(0, flat.numel(), (4,)).tolist():
                def val(delta):
                    z = flat.clone()
                    z[idx] += delta
                    args = [z.reshape(inp.shape) if a is inp else a.detach() for a in inputs]
                    o = fn(*args)
                    o = o if isinstance(o, tuple) else (o,)
                    return float(sum((oo.reshape(nb, -1) * w.reshape(nb, -1)[:, :oo.reshape(nb, -1).shape[1]]).sum()
                                     for oo, w in zip(o, wts)))
                num = (val(h) - val(-h)) / (2 * h)
                assert abs(num - float(g.reshape(-1)[idx])) <= tol * max(1.0, abs(num)), (fn, idx, num, float(g.reshape(-1)[idx]))

    fd_check(lambda x_: ag.U1Force.apply(x_, 2.5, [T, X]).reshape(nb, -1), [x])
    fd_check(lambda x_: ag.U1Action.apply(x_, 2.5, [T, X]).reshape(nb, 1), [x])
    fd_check(lambda x_: ag.U1WilsonLoops.apply(x_, [T, X]).reshape(nb, -1), [x])
    fd_check(lambda v_: ag.U1Kinetic.apply(v_).reshape(nb, 1), [v])
    f = torch.randn(nb, xdim, dtype=torch.float64, device=DEV, requires_grad=True)
    for sign in (+1, -1):
        fd_check(lambda v_, f_, s_, t_, q_, e_: ag.U1VUpdate.apply(v_, f_, s_, t_, q_, e_, sign), [v, f, s, t, q, eps])
        for ncp in (True, False):
            xf = x.detach().reshape(nb, -1).clone().requires_grad_(True)
            fd_check(lambda x_, v_, s_, t_, q_, e_: ag.U1XUpdate.apply(x_, v_, s_, t_, q_, mask, e_, sign, ncp),
                     [xf, v, s, t, q, eps], tol=2e-5)


def test_su3_training_raises_until_adjoints_exist(default_dtype):
    default_dtype(torch.float64)
    from l2hmc_b200.configs import DynamicsConfig
    from l2hmc_b200.dynamics.pytorch.dynamics import Dynamics
    from l2hmc_b200.lattice.su3.pytorch.lattice import LatticeSU3
    cfg = DynamicsConfig(nchains=1, group='SU3', latvolume=[2, 2, 2, 2], nleapfrog=1, eps=0.05, verbose=False,
                         use_split_xnets=False, use_separate_networks=False)
    lat = LatticeSU3(1, [2, 2, 2, 2])
    dyn = Dynamics(potential_fn=lat.action, config=cfg, network_factory=None)
    with pytest.raises(NotImplementedError):
        dyn((lat.random(), torch.tensor(6.0)))
