"""ctypes binding of libl2b.so (include/l2b.h).

The CUDA library is the product; there is NO CPU fallback.  Importing this
module where the library has not been built, or calling into it without a CUDA
device, raises immediately.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_double, c_float, c_int, c_longlong, c_size_t, c_uint64, c_void_p, POINTER
from pathlib import Path

PKG = Path(__file__).resolve().parent
LIB_PATH = PKG / 'libl2b.so'

L2B_F32, L2B_F64, L2B_BF16 = 0, 1, 2


class L2BError(RuntimeError):
    pass


def _hashes(lib: ctypes.CDLL) -> tuple[int, int]:
    out = []
    for name in ('l2b_abi_hash', 'l2b_source_hash'):
        f = getattr(lib, name, None)
        if f is None:
            out.append(-1)            # a binary older than these exports
            continue
        f.argtypes, f.restype = [], ctypes.c_uint32
        out.append(int(f()))
    return out[0], out[1]


def _load() -> ctypes.CDLL:
    """dlopen libl2b.so and make sure it is THE library of this tree: its `l2b_abi_hash()` must equal the
    hash of include/l2b.h (the signatures below are transcribed from that header; a stale binary under new
    positional pointer arguments would corrupt memory silently) and its `l2b_source_hash()` the hash of csrc/.
    A mismatching or missing library is rebuilt when a compiler is there (L2B_AUTOBUILD=1, the default; one
    builder at a time under a file lock), otherwise this raises.  There is no CPU fallback."""
    from . import _build
    auto = os.environ.get('L2B_AUTOBUILD', '1') == '1'
    if not LIB_PATH.exists():
        if auto:
            _build.build(force=True)
        if not LIB_PATH.exists():
            raise L2BError(
                f'{LIB_PATH} is missing: build it with `python -m l2hmc_b200._build` '
                '(or __graft_entry__.build()); there is no CPU fallback')
    lib = ctypes.CDLL(str(LIB_PATH))
    want = (_build.abi_hash(), _build.source_hash()) if _build.HEADER.exists() else None
    if want is None or _hashes(lib) == want:
        return lib
    if auto:
        import _ctypes
        _ctypes.dlclose(lib._handle)
        del lib
        try:
            _build.build(force=True)
        except RuntimeError as e:
            raise L2BError(f'{LIB_PATH} does not match this source tree and could not be rebuilt: {e}') from e
        lib = ctypes.CDLL(str(LIB_PATH))
    got = _hashes(lib)
    if got[0] != want[0]:
        raise L2BError(f'{LIB_PATH} was built against a different include/l2b.h (abi hash {got[0]:#x}, header '
                       f'{want[0]:#x}): rebuild with `python -m l2hmc_b200._build --force`')
    if got[1] != want[1]:
        raise L2BError(f'{LIB_PATH} was built from different sources (hash {got[1]:#x}, tree {want[1]:#x}): '
                       'rebuild with `python -m l2hmc_b200._build --force`')
    return lib


_lib = _load()

_P = c_void_p
_DIMS = POINTER(c_int)
# name -> argtypes (restype is int unless listed in _RES)
_SIGS = {
    'l2b_su3_aos_to_soa': [_P, _P, c_int, _DIMS, c_int, _P],
    'l2b_su3_soa_to_aos': [_P, _P, c_int, _DIMS, c_int, _P],
    'l2b_su3_wilson_loops': [_P, _P, c_int, _DIMS, c_int, _P, c_size_t, _P],
    'l2b_su3_plaq_sums': [_P, _P, c_int, _DIMS, c_int, _P, c_size_t, _P],
    'l2b_su3_force': [_P, c_double, _P, _P, c_int, _DIMS, c_int, _P, c_size_t, _P],
    'l2b_su3_action_grad_c1': [_P, _P, c_double, c_double, _P, _P, c_int, _DIMS, c_int, _P, c_size_t, _P],
    'l2b_su3_force_c1': [_P, c_double, c_double, _P, _P, c_int, _DIMS, c_int, _P, c_size_t, _P],
    'l2b_su3_exp': [_P, c_double, _P, c_size_t, c_int, _P],
    'l2b_su3_update_gauge': [_P, _P, c_double, _P, _P, c_int, _P, c_int, _DIMS, c_int, _P],
    'l2b_su3_project': [_P, _P, _P, c_size_t, c_int, _P],
    'l2b_su3_to_vec': [_P, _P, c_size_t, c_int, _P],
    'l2b_su3_from_vec': [_P, _P, c_size_t, c_int, _P],
    'l2b_su3_tah': [_P, _P, c_size_t, c_int, _P],
    'l2b_su3_kinetic': [_P, _P, c_int, _DIMS, c_int, _P, c_size_t, _P],
    'l2b_su3_check': [_P, _P, _P, c_int, _DIMS, c_int, _P, c_size_t, _P],
    'l2b_su3_rand_momentum': [c_uint64, c_uint64, _P, _P, _P, c_int, _DIMS, c_int, _P, c_size_t, _P],
    'l2b_su3_vupdate': [_P, _P, _P, _P, _P, c_double, _P, c_int, _P, _P, c_int, _DIMS, c_int, _P, c_size_t, _P],
    'l2b_su3_hmc_trajectory': [_P, _P, c_double, c_double, c_int, _P, _P, _P, c_int, _DIMS, c_int, _P, c_size_t, _P],
    'l2b_su3_action_grad': [_P, _P, _P, c_int, _DIMS, c_int, _P, c_size_t, _P],
    'l2b_su3_vupdate_bwd': [_P, _P, _P, _P, _P, c_double, _P, c_int, _P, _P, _P, _P, _P, _P, _P, _P, c_int, _DIMS, c_int, _P, c_size_t, _P],
    'l2b_su3_update_gauge_bwd': [_P, _P, c_double, _P, _P, c_int, _P, _P, _P, _P, _P, c_int, _DIMS, c_int, _P, c_size_t, _P],
    'l2b_su3_to_vec_bwd': [_P, _P, c_size_t, c_int, _P],
    'l2b_su3_wilson_loops_bwd': [_P, _P, _P, c_int, _DIMS, c_int, _P, c_size_t, _P],
    'l2b_su3_force_bwd': [_P, c_double, _P, _P, c_int, _DIMS, c_int, _P, c_size_t, _P],
    'l2b_su3_force_planar': [_P, c_double, _P, c_int, _DIMS, c_int, _P],
    'l2b_su3_project_vec_planar': [_P, _P, c_int, c_int, _DIMS, c_int, _P],
    'l2b_su3_update_gauge_planar': [_P, _P, c_double, _P, _P, c_int, _P, c_int, _DIMS, c_int, _P],
    'l2b_su3_project_vec': [_P, _P, c_int, c_size_t, c_int, _P],
    'l2b_su3_project_vec_planar_lm': [_P, _P, c_int, c_int, _DIMS, c_int, _P],
    'l2b_su3_update_gauge_planar_pair': [_P, _P, c_double, _P, _P, c_int, _P, c_int, _DIMS, c_int, _P],
    'l2b_su3_heads_vupdate_pair': [_P, _P, _P, _P, _P, _P, _P, c_float, _P, _P, c_double, _P, c_int, c_double, _P, c_int,
                                   c_int, _P, _P, c_int, c_int, c_int, _P, c_size_t, _P],
    'l2b_su3_input_pack': [_P, _P, c_int, _P, c_int, c_int, _P],
    'l2b_su3_input_layer': [_P, _P, _P, _P, _P, c_int, _P, c_int, c_int, c_int, c_int, _P, c_size_t, _P],
    'l2b_gemm_bf16': [POINTER(_P), c_longlong, c_int, POINTER(_P), c_longlong, c_int, c_int, c_int, c_int, c_int, c_int, _P,
                      c_int, c_longlong, c_int, _P, c_int, c_int, _P, c_size_t, _P],
    'l2b_split_bf16x3': [_P, c_longlong, c_longlong, c_longlong, _P, c_longlong, _P],
    'l2b_conv_im2col': [_P, c_int, c_int, c_int, c_int, c_int, c_int, POINTER(c_longlong), _P, c_int, c_int, _P],
    'l2b_conv_col2im': [_P, c_int, c_longlong, c_int, c_int, c_int, c_int, c_int, _P, POINTER(c_longlong), c_int, _P],
    'l2b_pool_act': [_P, c_int, c_int, c_int, c_int, c_int, c_int, c_int, _P, _P, _P, _P],
    'l2b_pool_act_bwd': [_P, _P, c_int, _P, _P, c_int, c_int, c_int, c_int, c_int, c_int, _P, _P],
    'l2b_su3_project_bwd': [_P, _P, _P, c_int, _P, c_size_t, c_int, _P],
    'l2b_su3_force_kick_planar': [_P, _P, c_double, c_double, _P, c_int, _DIMS, c_int, _P, c_size_t, _P],
    'l2b_su3_force_kick_drift_planar': [_P, _P, _P, c_double, c_double, c_double, _P, c_int, _DIMS, c_int, _P, c_size_t, _P],
    'l2b_su3_drift_planar': [_P, _P, c_double, c_int, _DIMS, c_int, _P],
    'l2b_set_option': [c_char_p, c_int],
    'l2b_su3_heads_vupdate_bwd': [_P, _P, _P, _P, _P, c_float, c_double, _P, c_int, _P, _P, _P, _P, _P, c_int, _P, _P,
                                  c_int, c_int, _P, c_size_t, _P],
    'l2b_vnet_pack_heads': [_P, _P, _P, c_int, _P, c_int, c_int, _P],
    'l2b_su3_heads_vupdate': [_P, _P, _P, _P, _P, _P, _P, c_float, _P, _P, c_double, _P, c_int, _P, _P, _P, c_int, c_int,
                              c_int, _P, c_size_t, _P],
    'l2b_u1_wilson_loops': [_P, _P, c_int, c_int, c_int, c_int, _P],
    'l2b_u1_wilson_loops4x4': [_P, _P, c_int, c_int, c_int, c_int, _P],
    'l2b_u1_observables': [_P, c_double, _P, c_int, c_int, c_int, c_int, _P],
    'l2b_u1_force': [_P, c_double, _P, c_int, c_int, c_int, c_int, _P],
    'l2b_u1_hmc_trajectory': [_P, _P, c_double, c_double, c_int, _P, _P, _P, c_int, c_int, c_int, c_int, _P],
    'l2b_u1_vupdate': [_P, _P, _P, _P, _P, c_double, _P, c_int, _P, _P, c_int, c_int, c_int, _P],
    'l2b_u1_xupdate': [_P, _P, _P, _P, _P, _P, c_double, _P, c_int, c_int, _P, _P, c_int, c_int, c_int, _P],
    'l2b_u1_kinetic': [_P, _P, c_int, c_int, c_int, _P],
    'l2b_u1_compat_proj': [_P, _P, c_size_t, c_int, _P],
    'l2b_u1_wilson_loops_bwd': [_P, _P, c_int, c_int, c_int, c_int, _P],
    'l2b_u1_force_bwd': [_P, c_double, _P, _P, c_int, c_int, c_int, c_int, _P],
    'l2b_u1_vupdate_bwd': [_P, _P, _P, _P, _P, c_double, _P, c_int, _P, _P, _P, _P, _P, _P, _P, _P, c_int, c_int, c_int, _P],
    'l2b_u1_xupdate_bwd': [_P, _P, _P, _P, _P, _P, c_double, _P, c_int, c_int, _P, _P, _P, _P, _P, _P, _P, _P, c_int, c_int, c_int, _P],
    'l2b_u1_heads_update': [c_int, _P, c_int, _P, _P, _P, _P, _P, _P, _P, _P, c_double, c_double, c_double, _P, _P, _P,
                            c_double, _P, c_int, c_int, _P, _P, c_int, c_int, c_int, _P, c_size_t, _P],
    'l2b_u1_input_layer': [c_int, _P, _P, _P, _P, _P, _P, _P, c_int, _P, c_int, c_int, c_int, _P, c_size_t, _P],
    'l2b_rowscale': [_P, _P, _P, c_int, c_int, c_int, _P],
    'l2b_accept_mix': [POINTER(_P), POINTER(_P), POINTER(_P), POINTER(c_size_t), c_int, _P, c_int, _P],
}
_RES = {
    'l2b_last_error': ([], c_char_p),
    'l2b_version': ([], c_int),
    'l2b_abi_hash': ([], ctypes.c_uint32),
    'l2b_source_hash': ([], ctypes.c_uint32),
    'l2b_launch_count': ([], c_uint64),
    'l2b_su3_ws_bytes': ([c_int, _DIMS, c_int], c_size_t),
    'l2b_u1_ws_bytes': ([c_int, c_int, c_int, c_int], c_size_t),
    'l2b_vnet_heads_packed_bytes': ([c_int, c_int], c_size_t),
    'l2b_vnet_heads_ws_bytes': ([c_int, c_int], c_size_t),
    'l2b_su3_input_packed_bytes': ([c_int, c_int], c_size_t),
    'l2b_su3_input_ws_bytes': ([c_int, c_int], c_size_t),
    'l2b_gemm_bf16_splits': ([c_int, c_int, c_int, c_int, c_int], c_int),
    'l2b_gemm_bf16_ws_bytes': ([c_int, c_int, c_int], c_size_t),
    'l2b_u1_heads_ws_bytes': ([c_int, c_int], c_size_t),
    'l2b_u1_input_ws_bytes': ([c_int, c_int], c_size_t),
}

EXPORTS = sorted(list(_SIGS) + list(_RES))

for _name, _args in _SIGS.items():
    _f = getattr(_lib, _name)
    _f.argtypes = _args
    _f.restype = c_int
for _name, (_args, _res) in _RES.items():
    _f = getattr(_lib, _name)
    _f.argtypes = _args
    _f.restype = _res


def last_error() -> str:
    return (_lib.l2b_last_error() or b'').decode()


def version() -> int:
    return int(_lib.l2b_version())


def set_option(key: str, value: int) -> None:
    call('l2b_set_option', key.encode(), int(value))


def launch_count() -> int:
    return int(_lib.l2b_launch_count())


def dims4(shape) -> ctypes.Array:
    assert len(shape) == 4
    return (c_int * 4)(*[int(s) for s in shape])


def su3_ws_bytes(nb: int, shape) -> int:
    n = int(_lib.l2b_su3_ws_bytes(int(nb), dims4(shape), L2B_F64))
    if n == 0:
        raise L2BError(f'l2b_su3_ws_bytes: {last_error()}')
    return n


def call(name: str, *args) -> None:
    """Invoke an int-returning entry point; raise L2BError with the library's
    message on a non-zero return (the reference's error style is Python
    exceptions/asserts, dynamics.py:1268, configs.py:482)."""
    rc = getattr(_lib, name)(*args)
    if rc != 0:
        raise L2BError(f'{name} failed (rc={rc}): {last_error()}')
