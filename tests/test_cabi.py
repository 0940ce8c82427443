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


def test_header_is_plain_c(tmp_path):
    """the boundary is a C ABI: include/l2b.h must compile as C99 without any C++ / CUDA / torch type"""
    import subprocess
    src = tmp_path / 'hc.c'
    src.write_text('#include "include/l2b.h"\nint main(void) { return l2b_version() ? 0 : 1; }\n')
    subprocess.check_call(['gcc', '-std=c99', '-Wall', '-Wextra', '-pedantic', '-fsyntax-only', '-I', str(ROOT), str(src)])


def test_new_entry_points_validate_arguments_without_gpu(lib):
    p = ctypes.c_void_p(256)
    d = lib.dims4([4, 4, 4, 4])
    # tcgen05 heads kernel: hidden size outside the supported range, bad sign, missing workspace for logdet
    args = [p, p, p, p, p, p, p, 1.0, p, p, 0.1, None, 1, p, p, None, 4, 1152]
    with pytest.raises(lib.L2BError, match='hidden'):
        lib.call('l2b_su3_heads_vupdate', *args, 260, p, 1 << 20, None)
    with pytest.raises(lib.L2BError, match='sign'):
        lib.call('l2b_su3_heads_vupdate', *(args[:12] + [0] + args[13:]), 64, p, 1 << 20, None)
    with pytest.raises(lib.L2BError, match='workspace'):
        lib.call('l2b_su3_heads_vupdate', *args, 64, None, 0, None)
    assert lib._lib.l2b_vnet_heads_packed_bytes(1152, 64) == 9 * 3 * 64 * 128 * 2
    assert lib._lib.l2b_vnet_heads_packed_bytes(1000, 40) == 8 * 3 * 64 * 128 * 2      # ragged tile, K padded to 64
    # improved-action force: needs at least one output
    with pytest.raises(lib.L2BError, match='null'):
        lib.call('l2b_su3_force_c1', p, 6.0, -0.331, None, None, 1, d, lib.L2B_F64, p, 1 << 30, None)
    with pytest.raises(lib.L2BError, match='workspace'):
        lib.call('l2b_su3_force_c1', p, 6.0, -0.331, p, None, 1, d, lib.L2B_F64, p, 16, None)
    # U(1) fused kernels
    with pytest.raises(lib.L2BError, match='hidden <= 32'):
        lib.call('l2b_u1_heads_update', 0, p, 48, p, p, p, p, p, p, p, p, 1.0, 1.0, 1.0, p, p, None, 0.1, None, 1, 1, p, p, 4,
                 128, lib.L2B_F32, p, 1 << 20, None)
    with pytest.raises(lib.L2BError, match='mask'):
        lib.call('l2b_u1_heads_update', 1, p, 16, p, p, p, p, p, p, p, p, 1.0, 1.0, 1.0, p, p, None, 0.1, None, 1, 1, p, p, 4,
                 128, lib.L2B_F32, p, 1 << 20, None)
    with pytest.raises(lib.L2BError, match='units'):
        lib.call('l2b_u1_input_layer', 0, p, p, None, p, p, p, p, 24, p, 4, 128, lib.L2B_F32, p, 1 << 20, None)
    # adjoints / planar variants
    with pytest.raises(lib.L2BError, match='vec_dtype'):
        lib.call('l2b_su3_project_vec', p, p, 7, 16, lib.L2B_F64, None)
    with pytest.raises(lib.L2BError, match='null'):
        lib.call('l2b_su3_project_bwd', p, None, None, lib.L2B_F64, p, 16, lib.L2B_F64, None)
    with pytest.raises(lib.L2BError, match='null'):
        lib.call('l2b_su3_update_gauge_planar', p, None, 1.0, None, None, 0, p, 2, d, lib.L2B_F64, None)


def test_plain_c_program_links_against_the_library(tmp_path):
    """examples/c_abi_cold_start.c: a C99 consumer with no Python and no torch links against libl2b.so
    (+ the CUDA runtime for its own allocations); without a GPU it reports that and exits 77, with one
    it checks the cold-start known answers"""
    import shutil
    import subprocess
    exe = tmp_path / 'cold'
    cudalib = Path(shutil.which('nvcc') or '/usr/local/cuda/bin/nvcc').resolve().parents[1] / 'lib64'
    subprocess.check_call(['gcc', '-std=c99', '-Wall', '-Wextra', '-I', str(ROOT), str(ROOT / 'examples' / 'c_abi_cold_start.c'),
                           '-o', str(exe), '-L', str(ROOT / 'l2hmc_b200'), '-ll2b', '-L', str(cudalib), '-lcudart',
                           f'-Wl,-rpath,{ROOT / "l2hmc_b200"}', f'-Wl,-rpath,{cudalib}', '-lm'])
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    assert r.returncode in (0, 77), (r.returncode, r.stdout, r.stderr)
    if r.returncode == 77:
        assert 'no CPU fallback' in r.stderr
    else:
        assert '-> ok' in r.stdout
