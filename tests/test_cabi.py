"""CPU tier: libl2b.so builds for sm_100a, loads, and exports exactly the
symbols include/l2b.h declares; argument validation that needs no GPU works."""
import ctypes
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]


@pytest.fixture(scope='module')
def lib():
    from l2hmc_b200 import _build
    _build.build()
    from l2hmc_b200 import _lib
    return _lib


def header_symbols():
    txt = (ROOT / 'include' / 'l2b.h').read_text()
    txt = re.sub(r'/\*.*?\*/', '', txt, flags=re.S)
    return sorted(set(re.findall(r'\b(l2b_[a-z0-9_]+)\s*\(', txt)))


def test_every_declared_symbol_is_exported(lib):
    syms = header_symbols()
    assert len(syms) >= 30
    cdll = ctypes.CDLL(str(lib.LIB_PATH))
    missing = [s for s in syms if not hasattr(cdll, s)]
    assert not missing, f'declared in l2b.h but not exported: {missing}'
    assert sorted(lib.EXPORTS) == syms, 'ctypes signatures and l2b.h disagree'


def test_sm100a_cubin_present(lib):
    import subprocess
    out = subprocess.run(['cuobjdump', '-lelf', str(lib.LIB_PATH)], capture_output=True, text=True).stdout
    assert 'sm_100a' in out


def test_argument_validation_without_gpu(lib):
    assert lib.version() >= 100
    d = lib.dims4([4, 4, 4, 4])
    n = lib.su3_ws_bytes(2, [4, 4, 4, 4])
    assert n >= 2 * 2 * 4 * 256 * 9 * 16
    with pytest.raises(lib.L2BError, match='null'):
        lib.call('l2b_su3_exp', None, 1.0, None, 4, lib.L2B_F64, None)
    with pytest.raises(lib.L2BError, match='L2B_F64 only'):
        lib.call('l2b_su3_exp', ctypes.c_void_p(16), 1.0, ctypes.c_void_p(16), 4, lib.L2B_F32, None)
    with pytest.raises(lib.L2BError, match='non-positive'):
        lib.call('l2b_su3_aos_to_soa', ctypes.c_void_p(16), ctypes.c_void_p(16), 0, d, lib.L2B_F64, None)
    with pytest.raises(lib.L2BError, match='workspace too small'):
        lib.call('l2b_su3_plaq_sums', ctypes.c_void_p(256), ctypes.c_void_p(256), 2, d, lib.L2B_F64,
                 ctypes.c_void_p(256), 128, None)
    with pytest.raises(lib.L2BError, match='nlf'):
        lib.call('l2b_u1_hmc_trajectory', ctypes.c_void_p(16), ctypes.c_void_p(16), 1.0, 0.1, 0, ctypes.c_void_p(16),
                 ctypes.c_void_p(16), ctypes.c_void_p(16), 2, 8, 8, lib.L2B_F32, None)
    with pytest.raises(lib.L2BError, match='shared memory'):
        lib.call('l2b_u1_hmc_trajectory', ctypes.c_void_p(16), ctypes.c_void_p(16), 1.0, 0.1, 2, ctypes.c_void_p(16),
                 ctypes.c_void_p(16), ctypes.c_void_p(16), 2, 512, 512, lib.L2B_F32, None)


def test_cpu_tensors_are_rejected_not_silently_computed(lib):
    import torch
    from l2hmc_b200 import ops
    x = torch.zeros(1, 4, 2, 2, 2, 2, 3, 3, dtype=torch.complex128)
    with pytest.raises(lib.L2BError, match='no CPU fallback'):
        ops.su3_plaq_sums(x)
    with pytest.raises(lib.L2BError, match='no CPU fallback'):
        ops.u1_force(torch.zeros(1, 2, 4, 4), 1.0)
