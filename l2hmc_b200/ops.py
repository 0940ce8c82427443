"""Tensor-level entry points over the C ABI (include/l2b.h).

Every function here takes/returns CUDA `torch.Tensor`s, hands raw device
pointers + the current CUDA stream to libl2b through ctypes and allocates the
outputs/workspace the library asks for.  torch is used for device memory and
streams only; no arithmetic of the hot path happens in torch here, and there is
no CPU path: a CPU tensor raises.
"""
from __future__ import annotations

import ctypes
from ctypes import c_size_t, c_void_p
from typing import Optional, Sequence

import torch

from . import _lib
from ._lib import L2B_BF16, L2B_F32, L2B_F64, L2BError, call, dims4

Tensor = torch.Tensor

_WS: dict = {}


try:                                                     # raw handle of the current stream without building a Stream object:
    _raw_stream = torch._C._cuda_getCurrentRawStream    # ~0.3 us instead of ~10 us, and `_stream` runs once per kernel
    _cur_device = torch._C._cuda_getDevice
except AttributeError:                                   # pragma: no cover  (older / newer torch without the private hooks)
    _raw_stream = _cur_device = None


def _stream() -> c_void_p:
    if _raw_stream is not None:
        return c_void_p(_raw_stream(_cur_device()))
    return c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t: Optional[Tensor]) -> c_void_p:
    return c_void_p(0 if t is None else t.data_ptr())


def _need_cuda(*ts: Optional[Tensor]) -> None:
    for t in ts:
        if t is not None and not t.is_cuda:
            raise L2BError(
                'l2hmc_b200 runs on CUDA tensors only (got a CPU tensor); '
                'there is no CPU fallback')


def _workspace(nbytes: int, device: torch.device) -> Tensor:
    """Grow-only scratch buffer per device.  Calls are stream-ordered on the
    current stream, so one buffer per device is safe for a single-stream caller
    (the reference is single-threaded on the default stream)."""
    key = (device.index, torch.cuda.current_stream(device).cuda_stream)
    ws = _WS.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.empty(max(nbytes, 256), dtype=torch.uint8, device=device)
        _WS[key] = ws
    return ws


def _eps_args(eps, dtype: torch.dtype, mult: float = 1.0):
    """Step size for the L2HMC update kernels: a Python number goes by value; a 0-dim CUDA tensor
    (the trainable `xeps`/`veps` after sigmoid(log(.))) stays on the device and the kernel reads
    eps = mult * (*ptr), so nothing synchronises and the call can be captured in a CUDA graph.
    -> (by-value double, device pointer, keep-alive)"""
    if isinstance(eps, Tensor):
        e = eps.detach()
        if e.is_cuda:
            e = e.to(dtype).reshape(1).contiguous()
            return float(mult), c_void_p(e.data_ptr()), e
        return float(mult) * float(e), c_void_p(0), None
    return float(mult) * float(eps), c_void_p(0), None


def free_workspaces() -> None:
    _WS.clear()


# ---------------------------------------------------------------------------
# SU(3)
# ---------------------------------------------------------------------------
def _su3_field(x: Tensor, shape: Optional[Sequence[int]] = None) -> tuple[Tensor, int, list[int]]:
    """-> (contiguous complex128 [nb,4,T,X,Y,Z,3,3], nb, [T,X,Y,Z])"""
    _need_cuda(x)
    if x.dtype != torch.complex128:
        raise L2BError(f'SU(3) fields must be complex128 (got {x.dtype})')
    if x.dim() != 8:
        if shape is None:
            raise L2BError(f'SU(3) field of shape {tuple(x.shape)} needs an explicit lattice shape')
        x = x.reshape(x.shape[0], 4, *shape, 3, 3)
    if tuple(x.shape[-2:]) != (3, 3) or x.shape[1] != 4:
        raise L2BError(f'bad SU(3) field shape {tuple(x.shape)}')
    return x.contiguous(), int(x.shape[0]), [int(s) for s in x.shape[2:6]]


def _su3_ws(nb: int, dims: Sequence[int], device) -> tuple[Tensor, int]:
    n = _lib.su3_ws_bytes(nb, dims)
    return _workspace(n, device), n


def su3_wilson_loops(x: Tensor) -> Tensor:
    x, nb, dims = _su3_field(x)
    out = torch.empty((6, nb, *dims), dtype=torch.complex128, device=x.device)
    ws, n = _su3_ws(nb, dims, x.device)
    call('l2b_su3_wilson_loops', _ptr(x), _ptr(out), nb, dims4(dims), L2B_F64, _ptr(ws), n, _stream())
    return out


def su3_plaq_sums(x: Tensor) -> Tensor:
    """[nb, 2] = (sum Re tr P, sum Im tr P)"""
    x, nb, dims = _su3_field(x)
    out = torch.empty((nb, 2), dtype=torch.float64, device=x.device)
    ws, n = _su3_ws(nb, dims, x.device)
    call('l2b_su3_plaq_sums', _ptr(x), _ptr(out), nb, dims4(dims), L2B_F64, _ptr(ws), n, _stream())
    return out


def su3_force(x: Tensor, beta: float, want_plaq_sum: bool = False):
    x, nb, dims = _su3_field(x)
    f = torch.empty_like(x)
    ps = torch.empty(nb, dtype=torch.float64, device=x.device) if want_plaq_sum else None
    ws, n = _su3_ws(nb, dims, x.device)
    call('l2b_su3_force', _ptr(x), float(beta), _ptr(f), _ptr(ps), nb, dims4(dims), L2B_F64, _ptr(ws), n, _stream())
    return (f, ps) if want_plaq_sum else f


def su3_force_c1(x: Tensor, beta: float, c1: float, want_force: bool = True, want_sums: bool = False):
    """improved-action force and / or loop sums [nb, 2] = (sum Re tr P, sum Re tr R)
    (lattice/su3/pytorch/lattice.py:96-112,252-269,299-308 with c1 != 0)"""
    x, nb, dims = _su3_field(x)
    f = torch.empty_like(x) if want_force else None
    sums = torch.empty(nb, 2, dtype=torch.float64, device=x.device) if want_sums else None
    ws, n = _su3_ws(nb, dims, x.device)
    call('l2b_su3_force_c1', _ptr(x), float(beta), float(c1), _ptr(f), _ptr(sums), nb, dims4(dims), L2B_F64, _ptr(ws), n,
         _stream())
    if want_force and want_sums:
        return f, sums
    return f if want_force else sums


def _mats(x: Tensor) -> tuple[Tensor, int]:
    _need_cuda(x)
    if x.dtype != torch.complex128 or tuple(x.shape[-2:]) != (3, 3):
        raise L2BError(f'expected [..., 3, 3] complex128 (got {tuple(x.shape)} {x.dtype})')
    x = x.contiguous()
    return x, x.numel() // 9


def su3_exp(p: Tensor, scale: float = 1.0) -> Tensor:
    p, n = _mats(p)
    out = torch.empty_like(p)
    call('l2b_su3_exp', _ptr(p), float(scale), _ptr(out), n, L2B_F64, _stream())
    return out


def su3_tah(x: Tensor) -> Tensor:
    x, n = _mats(x)
    out = torch.empty_like(x)
    call('l2b_su3_tah', _ptr(x), _ptr(out), n, L2B_F64, _stream())
    return out


def su3_project(x: Tensor, want_matrix: bool = True, want_vec: bool = False):
    x, n = _mats(x)
    m = torch.empty_like(x) if want_matrix else None
    v = torch.empty((*x.shape[:-2], 8), dtype=torch.float64, device=x.device) if want_vec else None
    call('l2b_su3_project', _ptr(x), _ptr(m), _ptr(v), n, L2B_F64, _stream())
    if want_matrix and want_vec:
        return m, v
    return m if want_matrix else v


_NET_DT = {torch.float64: L2B_F64, torch.float32: L2B_F32, torch.bfloat16: L2B_BF16}


def _net_dt(dtype: torch.dtype) -> int:
    if dtype not in _NET_DT:
        raise L2BError(f'net-side buffers must be float64 / float32 / bfloat16 (got {dtype})')
    return _NET_DT[dtype]


def su3_project_vec(x: Tensor, dtype: torch.dtype = torch.float64) -> Tensor:
    """group_to_vec = su3_to_vec(projectSU(x)) written in the nets' element type"""
    x, n = _mats(x)
    v = torch.empty((*x.shape[:-2], 8), dtype=dtype, device=x.device)
    call('l2b_su3_project_vec', _ptr(x), _ptr(v), _net_dt(dtype), n, L2B_F64, _stream())
    return v


def su3_project_bwd(x: Tensor, gmat: Optional[Tensor] = None, gvec: Optional[Tensor] = None) -> Tensor:
    """adjoint of projectSU (gmat) and/or of group_to_vec (gvec), closed form"""
    x, n = _mats(x)
    vdt = L2B_F64
    if gmat is not None:
        gmat, ng = _mats(gmat.to(torch.complex128))
        if ng != n:
            raise L2BError('gmat and x differ in size')
    if gvec is not None:
        _need_cuda(gvec)
        vdt = _net_dt(gvec.dtype)
        gvec = gvec.contiguous()
        if gvec.numel() != 8 * n:
            raise L2BError(f'gvec must have {8 * n} entries (got {gvec.numel()})')
    gx = torch.empty_like(x)
    call('l2b_su3_project_bwd', _ptr(x), _ptr(gmat), _ptr(gvec), vdt, _ptr(gx), n, L2B_F64, _stream())
    return gx


def su3_to_vec(x: Tensor) -> Tensor:
    x, n = _mats(x)
    v = torch.empty((*x.shape[:-2], 8), dtype=torch.float64, device=x.device)
    call('l2b_su3_to_vec', _ptr(x), _ptr(v), n, L2B_F64, _stream())
    return v


def su3_from_vec(v: Tensor) -> Tensor:
    _need_cuda(v)
    if v.dtype != torch.float64 or v.shape[-1] != 8:
        raise L2BError(f'expected [..., 8] float64 (got {tuple(v.shape)} {v.dtype})')
    v = v.contiguous()
    x = torch.empty((*v.shape[:-1], 3, 3), dtype=torch.complex128, device=v.device)
    call('l2b_su3_from_vec', _ptr(v), _ptr(x), v.numel() // 8, L2B_F64, _stream())
    return x


def su3_update_gauge(x: Tensor, p: Tensor, eps=1.0, mask: Optional[Tensor] = None,
                     mask_complement: bool = False, eps_mult: float = 1.0) -> Tensor:
    """x' = m*x + exp(eps_mult * eps * p) ((1-m)*x); eps: number or 0-dim CUDA tensor"""
    x, nb, dims = _su3_field(x)
    p, nbp, _ = _su3_field(p, dims)
    if p.shape != x.shape:
        raise L2BError(f'x {tuple(x.shape)} and p {tuple(p.shape)} differ')
    if mask is not None:
        _need_cuda(mask)
        mask = mask.to(torch.float32).contiguous()
        if mask.numel() != x[0].numel():
            raise L2BError(f'mask has {mask.numel()} entries, expected {x[0].numel()}')
    out = torch.empty_like(x)
    ev, ep, _keep = _eps_args(eps, torch.float64, eps_mult)
    call('l2b_su3_update_gauge', _ptr(x), _ptr(p), ev, ep, _ptr(mask), int(mask_complement), _ptr(out), nb,
         dims4(dims), L2B_F64, _stream())
    return out


def su3_kinetic(p: Tensor) -> Tensor:
    p, nb, dims = _su3_field(p)
    ke = torch.empty(nb, dtype=torch.float64, device=p.device)
    ws, n = _su3_ws(nb, dims, p.device)
    call('l2b_su3_kinetic', _ptr(p), _ptr(ke), nb, dims4(dims), L2B_F64, _ptr(ws), n, _stream())
    return ke


def su3_check(x: Tensor) -> tuple[Tensor, Tensor]:
    x, nb, dims = _su3_field(x)
    avg = torch.empty(nb, dtype=torch.float64, device=x.device)
    mx = torch.empty_like(avg)
    ws, n = _su3_ws(nb, dims, x.device)
    call('l2b_su3_check', _ptr(x), _ptr(avg), _ptr(mx), nb, dims4(dims), L2B_F64, _ptr(ws), n, _stream())
    return avg, mx


def su3_rand_momentum(nb: int, dims: Sequence[int], seed: int, offset: int, device, want_ke: bool = False,
                      offset_dev: Optional[Tensor] = None):
    """offset_dev: optional int64 CUDA scalar added to `offset` and bumped by one after the draw
    (CUDA-graph replays then draw fresh momenta)"""
    device = torch.device(device)
    if device.type != 'cuda':
        raise L2BError('su3_rand_momentum needs a CUDA device')
    p = torch.empty((nb, 4, *dims, 3, 3), dtype=torch.complex128, device=device)
    ke = torch.empty(nb, dtype=torch.float64, device=device) if want_ke else None
    ws, n = _su3_ws(nb, dims, device)
    with torch.cuda.device(device):
        call('l2b_su3_rand_momentum', int(seed) & (2**64 - 1), int(offset), _ptr(offset_dev), _ptr(p), _ptr(ke), nb,
             dims4(dims),
             L2B_F64, _ptr(ws), n, _stream())
    return (p, ke) if want_ke else p


def su3_vupdate(v: Tensor, force: Tensor, s: Optional[Tensor], t: Optional[Tensor], q: Optional[Tensor],
                eps, sign: int) -> tuple[Tensor, Tensor]:
    v, nb, dims = _su3_field(v)
    force, _, _ = _su3_field(force, dims)
    xdim = v[0].numel()

    def real(a):
        if a is None:
            return None
        _need_cuda(a)
        a = a.to(torch.float64).contiguous()
        if a.numel() != nb * xdim:
            raise L2BError(f's/t/q must have {nb * xdim} entries (got {a.numel()})')
        return a
    s, t, q = real(s), real(t), real(q)
    out = torch.empty_like(v)
    logdet = torch.empty(nb, dtype=torch.float64, device=v.device)
    ws, n = _su3_ws(nb, dims, v.device)
    ev, ep, _keep = _eps_args(eps, torch.float64)
    call('l2b_su3_vupdate', _ptr(v), _ptr(force), _ptr(s), _ptr(t), _ptr(q), ev, ep, int(sign), _ptr(out),
         _ptr(logdet), nb, dims4(dims), L2B_F64, _ptr(ws), n, _stream())
    return out, logdet


def su3_hmc_trajectory(x: Tensor, v: Tensor, beta: float, eps: float, nlf: int):
    """-> (x_prop, v_prop, energies[nb, 4] = (KE0, S0, KE1, S1))"""
    x, nb, dims = _su3_field(x)
    v, _, _ = _su3_field(v, dims)
    if v.shape != x.shape:
        raise L2BError(f'x {tuple(x.shape)} and v {tuple(v.shape)} differ')
    xo = torch.empty_like(x)
    vo = torch.empty_like(v)
    en = torch.empty((nb, 4), dtype=torch.float64, device=x.device)
    ws, n = _su3_ws(nb, dims, x.device)
    call('l2b_su3_hmc_trajectory', _ptr(x), _ptr(v), float(beta), float(eps), int(nlf), _ptr(xo), _ptr(vo),
         _ptr(en), nb, dims4(dims), L2B_F64, _ptr(ws), n, _stream())
    return xo, vo, en


def su3_aos_to_soa(x: Tensor) -> Tensor:
    x, nb, dims = _su3_field(x)
    out = torch.empty_like(x)
    call('l2b_su3_aos_to_soa', _ptr(x), _ptr(out), nb, dims4(dims), L2B_F64, _stream())
    return out


def su3_soa_to_aos(x: Tensor) -> Tensor:
    x, nb, dims = _su3_field(x)
    out = torch.empty_like(x)
    call('l2b_su3_soa_to_aos', _ptr(x), _ptr(out), nb, dims4(dims), L2B_F64, _stream())
    return out


# ---- planar-state variants (L2HMC inference sweep without layout conversions) ----
def su3_force_planar(x_planar: Tensor, beta: float) -> Tensor:
    x, nb, dims = _su3_field(x_planar)
    f = torch.empty_like(x)
    call('l2b_su3_force_planar', _ptr(x), float(beta), _ptr(f), nb, dims4(dims), L2B_F64, _stream())
    return f


def su3_project_vec_planar(x_planar: Tensor, dtype: torch.dtype = torch.float64) -> Tensor:
    """vec8 [nb, 4, T, X, Y, Z, 8] (same order as su3_project_vec) from a planar field"""
    x, nb, dims = _su3_field(x_planar)
    v = torch.empty((nb, 4, *dims, 8), dtype=dtype, device=x.device)
    call('l2b_su3_project_vec_planar', _ptr(x), _ptr(v), _net_dt(dtype), nb, dims4(dims), L2B_F64, _stream())
    return v


def su3_update_gauge_planar_pair(x_planar: Tensor, p_planar: Tensor, eps, mask_planar: Tensor,
                                 first_complement: bool = False, eps_mult: float = 1.0) -> Tensor:
    """both masked link updates of one leapfrog layer in one pass (mask, then its complement; swapped when
    `first_complement`): include/l2b.h, l2b_su3_update_gauge_planar_pair"""
    x, nb, dims = _su3_field(x_planar)
    p, _, _ = _su3_field(p_planar, dims)
    _need_cuda(mask_planar)
    mask_planar = mask_planar.to(torch.float32).contiguous()
    if mask_planar.numel() != x[0].numel():
        raise L2BError(f'mask has {mask_planar.numel()} entries, expected {x[0].numel()}')
    out = torch.empty_like(x)
    ev, ep, _keep = _eps_args(eps, torch.float64, eps_mult)
    call('l2b_su3_update_gauge_planar_pair', _ptr(x), _ptr(p), ev, ep, _ptr(mask_planar), int(first_complement),
         _ptr(out), nb, dims4(dims), L2B_F64, _stream())
    return out


def su3_update_gauge_planar(x_planar: Tensor, p_planar: Tensor, eps=1.0, mask_planar: Optional[Tensor] = None,
                            mask_complement: bool = False, eps_mult: float = 1.0) -> Tensor:
    x, nb, dims = _su3_field(x_planar)
    p, _, _ = _su3_field(p_planar, dims)
    if mask_planar is not None:
        _need_cuda(mask_planar)
        mask_planar = mask_planar.to(torch.float32).contiguous()
        if mask_planar.numel() != x[0].numel():
            raise L2BError(f'mask has {mask_planar.numel()} entries, expected {x[0].numel()}')
    out = torch.empty_like(x)
    ev, ep, _keep = _eps_args(eps, torch.float64, eps_mult)
    call('l2b_su3_update_gauge_planar', _ptr(x), _ptr(p), ev, ep, _ptr(mask_planar), int(mask_complement), _ptr(out),
         nb, dims4(dims), L2B_F64, _stream())
    return out


# ---------------------------------------------------------------------------
# U(1)
# ---------------------------------------------------------------------------
def _dt(t: Tensor) -> int:
    if t.dtype == torch.float32:
        return L2B_F32
    if t.dtype == torch.float64:
        return L2B_F64
    raise L2BError(f'U(1) fields must be float32/float64 (got {t.dtype})')


def _u1_field(x: Tensor, shape: Optional[Sequence[int]] = None) -> tuple[Tensor, int, int, int]:
    _need_cuda(x)
    if x.dim() != 4:
        if shape is None:
            raise L2BError(f'U(1) field of shape {tuple(x.shape)} needs an explicit lattice shape')
        x = x.reshape(x.shape[0], 2, *shape)
    if x.shape[1] != 2:
        raise L2BError(f'bad U(1) field shape {tuple(x.shape)}')
    _dt(x)
    return x.contiguous(), int(x.shape[0]), int(x.shape[2]), int(x.shape[3])


def u1_wilson_loops(x: Tensor, shape=None) -> Tensor:
    x, nb, T, X = _u1_field(x, shape)
    w = torch.empty((nb, T, X), dtype=x.dtype, device=x.device)
    call('l2b_u1_wilson_loops', _ptr(x), _ptr(w), nb, T, X, _dt(x), _stream())
    return w


def u1_wilson_loops4x4(x: Tensor, shape=None) -> Tensor:
    """[nb, T, X] angles of the reference's 4x4 loops (before its trailing `.T`)"""
    x, nb, T, X = _u1_field(x, shape)
    w = torch.empty((nb, T, X), dtype=x.dtype, device=x.device)
    call('l2b_u1_wilson_loops4x4', _ptr(x), _ptr(w), nb, T, X, _dt(x), _stream())
    return w


def u1_observables(x: Tensor, beta: float, shape=None) -> Tensor:
    """[nb, 4] = (action, plaq, sinQ, intQ)"""
    x, nb, T, X = _u1_field(x, shape)
    obs = torch.empty((nb, 4), dtype=x.dtype, device=x.device)
    call('l2b_u1_observables', _ptr(x), float(beta), _ptr(obs), nb, T, X, _dt(x), _stream())
    return obs


def u1_force(x: Tensor, beta: float, shape=None) -> Tensor:
    x, nb, T, X = _u1_field(x, shape)
    f = torch.empty_like(x)
    call('l2b_u1_force', _ptr(x), float(beta), _ptr(f), nb, T, X, _dt(x), _stream())
    return f


# shared memory one block may opt into on sm_100a: the whole-trajectory U(1) kernel needs 5 T X elements of it
# (csrc/l2b_u1.cu hmc_smem_bytes) and returns L2B_ERR_UNSUPPORTED above; `Dynamics` takes the per-step path there
U1_TRAJECTORY_SMEM_LIMIT = 227 * 1024


def u1_hmc_trajectory(x: Tensor, v: Tensor, beta: float, eps: float, nlf: int, shape=None):
    x, nb, T, X = _u1_field(x, shape)
    v, _, _, _ = _u1_field(v, (T, X))
    if v.dtype != x.dtype:
        raise L2BError('x and v dtypes differ')
    xo, vo = torch.empty_like(x), torch.empty_like(v)
    en = torch.empty((nb, 4), dtype=x.dtype, device=x.device)
    call('l2b_u1_hmc_trajectory', _ptr(x), _ptr(v), float(beta), float(eps), int(nlf), _ptr(xo), _ptr(vo), _ptr(en),
         nb, T, X, _dt(x), _stream())
    return xo, vo, en


def _rows(a: Optional[Tensor], nb: int, like: Tensor) -> Optional[Tensor]:
    if a is None:
        return None
    _need_cuda(a)
    return a.to(like.dtype).reshape(nb, -1).contiguous()


def u1_vupdate(v: Tensor, force: Tensor, s, t, q, eps, sign: int) -> tuple[Tensor, Tensor]:
    _need_cuda(v, force)
    nb = v.shape[0]
    v2 = v.reshape(nb, -1).contiguous()
    f2 = force.to(v.dtype).reshape(nb, -1).contiguous()
    xdim = v2.shape[1]
    s, t, q = _rows(s, nb, v2), _rows(t, nb, v2), _rows(q, nb, v2)
    out = torch.empty_like(v2)
    logdet = torch.empty(nb, dtype=v.dtype, device=v.device)
    ev, ep, _keep = _eps_args(eps, v2.dtype)
    call('l2b_u1_vupdate', _ptr(v2), _ptr(f2), _ptr(s), _ptr(t), _ptr(q), ev, ep, int(sign), _ptr(out),
         _ptr(logdet), nb, xdim, _dt(v2), _stream())
    return out.reshape(v.shape), logdet


def u1_xupdate(x: Tensor, v: Tensor, s, t, q, mask: Tensor, eps, sign: int, use_ncp: bool):
    _need_cuda(x, v, mask)
    nb = x.shape[0]
    x2 = x.reshape(nb, -1).contiguous()
    v2 = v.to(x.dtype).reshape(nb, -1).contiguous()
    xdim = x2.shape[1]
    s, t, q = _rows(s, nb, x2), _rows(t, nb, x2), _rows(q, nb, x2)
    mask = mask.to(torch.float32).reshape(-1).contiguous()
    if mask.numel() != xdim:
        raise L2BError(f'mask has {mask.numel()} entries, expected {xdim}')
    out = torch.empty_like(x2)
    logdet = torch.empty(nb, dtype=x.dtype, device=x.device)
    ev, ep, _keep = _eps_args(eps, x2.dtype)
    call('l2b_u1_xupdate', _ptr(x2), _ptr(v2), _ptr(s), _ptr(t), _ptr(q), _ptr(mask), ev, ep, int(sign),
         int(bool(use_ncp)), _ptr(out), _ptr(logdet), nb, xdim, _dt(x2), _stream())
    return out.reshape(x.shape), logdet


def u1_kinetic(v: Tensor) -> Tensor:
    _need_cuda(v)
    nb = v.shape[0]
    v2 = v.reshape(nb, -1).contiguous()
    ke = torch.empty(nb, dtype=v.dtype, device=v.device)
    call('l2b_u1_kinetic', _ptr(v2), _ptr(ke), nb, v2.shape[1], _dt(v2), _stream())
    return ke


def u1_compat_proj(x: Tensor) -> Tensor:
    _need_cuda(x)
    x = x.contiguous()
    out = torch.empty_like(x)
    call('l2b_u1_compat_proj', _ptr(x), _ptr(out), x.numel(), _dt(x), _stream())
    return out


# ---------------------------------------------------------------------------
# accept / reject
# ---------------------------------------------------------------------------
def accept_mix(accept: Tensor, pairs: Sequence[tuple[Tensor, Tensor]]) -> list[Tensor]:
    """out_k[b] = accept[b] ? prop_k[b] : init_k[b] for every (init_k, prop_k).
    accept: [nb] float32 0/1 (dynamics.py:1081-1087).  Outputs are [nb, -1]."""
    _need_cuda(accept)
    accept = accept.to(torch.float32).contiguous()
    nb = accept.numel()
    n = len(pairs)
    inits, props, outs = [], [], []
    for a, b in pairs:
        _need_cuda(a, b)
        a = a.reshape(nb, -1).contiguous()
        b = b.to(a.dtype).reshape(nb, -1).contiguous()
        if a.shape != b.shape:
            raise L2BError(f'init {tuple(a.shape)} and proposed {tuple(b.shape)} differ')
        inits.append(a)
        props.append(b)
        outs.append(torch.empty_like(a))
    arr = c_void_p * n
    pi = arr(*[a.data_ptr() for a in inits])
    pp = arr(*[b.data_ptr() for b in props])
    po = arr(*[o.data_ptr() for o in outs])
    rb = (c_size_t * n)(*[a.shape[1] * a.element_size() for a in inits])
    call('l2b_accept_mix', ctypes.cast(pi, ctypes.POINTER(c_void_p)), ctypes.cast(pp, ctypes.POINTER(c_void_p)),
         ctypes.cast(po, ctypes.POINTER(c_void_p)), rb, n, _ptr(accept), nb, _stream())
    return outs


# ---------------------------------------------------------------------------
# adjoints (training path)
# ---------------------------------------------------------------------------
def u1_wilson_loops_bwd(gw: Tensor) -> Tensor:
    _need_cuda(gw)
    gw = gw.contiguous()
    nb, T, X = gw.shape
    gx = torch.empty((nb, 2, T, X), dtype=gw.dtype, device=gw.device)
    call('l2b_u1_wilson_loops_bwd', _ptr(gw), _ptr(gx), nb, T, X, _dt(gw), _stream())
    return gx


def u1_force_bwd(x: Tensor, beta: float, gforce: Tensor, shape=None) -> Tensor:
    x, nb, T, X = _u1_field(x, shape)
    gforce = gforce.to(x.dtype).reshape(x.shape).contiguous()
    gx = torch.empty_like(x)
    call('l2b_u1_force_bwd', _ptr(x), float(beta), _ptr(gforce), _ptr(gx), nb, T, X, _dt(x), _stream())
    return gx


def u1_vupdate_bwd(v, force, s, t, q, eps, sign: int, gv_out, glogdet):
    nb = v.shape[0]
    v2 = v.reshape(nb, -1).contiguous()
    f2 = force.to(v.dtype).reshape(nb, -1).contiguous()
    xdim = v2.shape[1]
    s, t, q = _rows(s, nb, v2), _rows(t, nb, v2), _rows(q, nb, v2)
    go = gv_out.to(v.dtype).reshape(nb, -1).contiguous()
    gl = None if glogdet is None else glogdet.to(v.dtype).contiguous()
    gv, gf = torch.empty_like(v2), torch.empty_like(v2)
    gs = torch.empty_like(v2) if s is not None else None
    gt = torch.empty_like(v2) if t is not None else None
    gq = torch.empty_like(v2) if q is not None else None
    geps = torch.empty(nb, dtype=v.dtype, device=v.device)
    ev, ep, _keep = _eps_args(eps, v2.dtype)
    call('l2b_u1_vupdate_bwd', _ptr(v2), _ptr(f2), _ptr(s), _ptr(t), _ptr(q), ev, ep, int(sign), _ptr(go),
         _ptr(gl), _ptr(gv), _ptr(gf), _ptr(gs), _ptr(gt), _ptr(gq), _ptr(geps), nb, xdim, _dt(v2), _stream())
    return gv, gf, gs, gt, gq, geps


def u1_xupdate_bwd(x, v, s, t, q, mask, eps, sign: int, use_ncp: bool, gx_out, glogdet):
    nb = x.shape[0]
    x2 = x.reshape(nb, -1).contiguous()
    v2 = v.to(x.dtype).reshape(nb, -1).contiguous()
    xdim = x2.shape[1]
    s, t, q = _rows(s, nb, x2), _rows(t, nb, x2), _rows(q, nb, x2)
    mask = mask.to(torch.float32).reshape(-1).contiguous()
    go = gx_out.to(x.dtype).reshape(nb, -1).contiguous()
    gl = None if glogdet is None else glogdet.to(x.dtype).contiguous()
    gx, gv = torch.empty_like(x2), torch.empty_like(x2)
    gs = torch.empty_like(x2) if s is not None else None
    gt = torch.empty_like(x2) if t is not None else None
    gq = torch.empty_like(x2) if q is not None else None
    geps = torch.empty(nb, dtype=x.dtype, device=x.device)
    ev, ep, _keep = _eps_args(eps, x2.dtype)
    call('l2b_u1_xupdate_bwd', _ptr(x2), _ptr(v2), _ptr(s), _ptr(t), _ptr(q), _ptr(mask), ev, ep, int(sign),
         int(bool(use_ncp)), _ptr(go), _ptr(gl), _ptr(gx), _ptr(gv), _ptr(gs), _ptr(gt), _ptr(gq), _ptr(geps), nb, xdim,
         _dt(x2), _stream())
    return gx, gv, gs, gt, gq, geps


def rowscale(a: Tensor, scale: Tensor) -> Tensor:
    """out[b, ...] = scale[b] * a[b, ...]   (real fields)"""
    _need_cuda(a, scale)
    nb = a.shape[0]
    a2 = a.reshape(nb, -1).contiguous()
    sc = scale.to(a.dtype).contiguous()
    out = torch.empty_like(a2)
    call('l2b_rowscale', _ptr(a2), _ptr(sc), _ptr(out), nb, a2.shape[1], _dt(a2), _stream())
    return out.reshape(a.shape)


def su3_action_grad(x: Tensor, coef: Tensor) -> Tensor:
    """gx = coef[b] * A^+ (A = staple sum)"""
    x, nb, dims = _su3_field(x)
    coef = coef.to(torch.float64).contiguous()
    gx = torch.empty_like(x)
    ws, n = _su3_ws(nb, dims, x.device)
    call('l2b_su3_action_grad', _ptr(x), _ptr(coef), _ptr(gx), nb, dims4(dims), L2B_F64, _ptr(ws), n, _stream())
    return gx


def su3_action_grad_c1(x: Tensor, c1: float, coef: Optional[Tensor] = None, scale: float = 0.0,
                       gforce: Optional[Tensor] = None) -> Tensor:
    """improved action (c1 != 0): gx = coef[b] Aimp^+, or with `gforce` the force adjoint
    TAH(gforce)^+ (scale Aimp^+); Aimp = (1 - 8 c1) A + c1 R"""
    x, nb, dims = _su3_field(x)
    if coef is not None:
        coef = coef.to(torch.float64).contiguous()
    if gforce is not None:
        gforce, _, _ = _su3_field(gforce.to(torch.complex128), dims)
    gx = torch.empty_like(x)
    ws, n = _su3_ws(nb, dims, x.device)
    call('l2b_su3_action_grad_c1', _ptr(x), _ptr(coef), float(scale), float(c1), _ptr(gforce), _ptr(gx), nb, dims4(dims),
         L2B_F64, _ptr(ws), n, _stream())
    return gx


def su3_force_bwd(x: Tensor, beta: float, gforce: Tensor) -> Tensor:
    """gx = TAH(gforce)^+ dsdx with dsdx = -(beta/3) A^+ (one stencil kernel)"""
    x, nb, dims = _su3_field(x)
    gforce, _, _ = _su3_field(gforce.to(torch.complex128), dims)
    gx = torch.empty_like(x)
    ws, n = _su3_ws(nb, dims, x.device)
    call('l2b_su3_force_bwd', _ptr(x), float(beta), _ptr(gforce), _ptr(gx), nb, dims4(dims), L2B_F64, _ptr(ws), n,
         _stream())
    return gx


def su3_wilson_loops_bwd(x: Tensor, gw: Tensor) -> Tensor:
    """adjoint of su3_wilson_loops: gw [6, nb, T, X, Y, Z] complex -> gx like x"""
    x, nb, dims = _su3_field(x)
    _need_cuda(gw)
    gw = gw.to(torch.complex128).contiguous()
    if tuple(gw.shape) != (6, nb, *dims):
        raise L2BError(f'gw must be {(6, nb, *dims)} (got {tuple(gw.shape)})')
    gx = torch.empty_like(x)
    ws, n = _su3_ws(nb, dims, x.device)
    call('l2b_su3_wilson_loops_bwd', _ptr(x), _ptr(gw), _ptr(gx), nb, dims4(dims), L2B_F64, _ptr(ws), n, _stream())
    return gx


def su3_vupdate_bwd(v, force, s, t, q, eps, sign: int, gv_out, glogdet):
    v, nb, dims = _su3_field(v)
    force, _, _ = _su3_field(force, dims)
    gv_out, _, _ = _su3_field(gv_out.to(torch.complex128), dims)
    xdim = v[0].numel()

    def real(a):
        return None if a is None else a.to(torch.float64).reshape(nb, xdim).contiguous()
    s, t, q = real(s), real(t), real(q)
    gl = None if glogdet is None else glogdet.to(torch.float64).contiguous()
    gv, gf = torch.empty_like(v), torch.empty_like(v)
    mk = lambda ref: None if ref is None else torch.empty((nb, xdim), dtype=torch.float64, device=v.device)  # noqa: E731
    gs, gt, gq = mk(s), mk(t), mk(q)
    geps = torch.empty(nb, dtype=torch.float64, device=v.device)
    ws, n = _su3_ws(nb, dims, v.device)
    ev, ep, _keep = _eps_args(eps, torch.float64)
    call('l2b_su3_vupdate_bwd', _ptr(v), _ptr(force), _ptr(s), _ptr(t), _ptr(q), ev, ep, int(sign), _ptr(gv_out),
         _ptr(gl), _ptr(gv), _ptr(gf), _ptr(gs), _ptr(gt), _ptr(gq), _ptr(geps), nb, dims4(dims), L2B_F64, _ptr(ws), n,
         _stream())
    return gv, gf, gs, gt, gq, geps


def su3_update_gauge_bwd(x, p, eps, mask, mask_complement: bool, gx_out, eps_mult: float = 1.0):
    x, nb, dims = _su3_field(x)
    p, _, _ = _su3_field(p, dims)
    gx_out, _, _ = _su3_field(gx_out.to(torch.complex128), dims)
    if mask is not None:
        mask = mask.to(torch.float32).contiguous()
    gx, gp = torch.empty_like(x), torch.empty_like(x)
    geps = torch.empty(nb, dtype=torch.float64, device=x.device)
    bad = torch.zeros(1, dtype=torch.int32, device=x.device)
    ws, n = _su3_ws(nb, dims, x.device)
    ev, ep, _keep = _eps_args(eps, torch.float64, eps_mult)
    call('l2b_su3_update_gauge_bwd', _ptr(x), _ptr(p), ev, ep, _ptr(mask), int(mask_complement), _ptr(gx_out),
         _ptr(gx), _ptr(gp), _ptr(geps), _ptr(bad), nb, dims4(dims), L2B_F64, _ptr(ws), n, _stream())
    return gx, gp, geps, bad


def su3_to_vec_bwd(gvec: Tensor) -> Tensor:
    _need_cuda(gvec)
    gvec = gvec.to(torch.float64).contiguous()
    gx = torch.empty((*gvec.shape[:-1], 3, 3), dtype=torch.complex128, device=gvec.device)
    call('l2b_su3_to_vec_bwd', _ptr(gvec), _ptr(gx), gvec.numel() // 8, L2B_F64, _stream())
    return gx


# ---------------------------------------------------------------------------
# vnet input layer on the tensor cores (split-K tcgen05 GEMM fed by link-major vec8 images)
# ---------------------------------------------------------------------------
INPUT_ACTIVATIONS = {'identity': 0, 'tanh': 1, 'relu': 2, 'swish': 3, 'leaky_relu': 4, 'elu': 5}


class InputPack:
    """bf16 K-major image of (W_x, W_v) of the vnet InputLayer + both biases (include/l2b.h, l2b_su3_input_pack)"""
    __slots__ = ('packed', 'bias_x', 'bias_v', 'nlinks', 'hidden', 'activation')

    def __init__(self, packed, bias_x, bias_v, nlinks, hidden, activation):
        self.packed, self.bias_x, self.bias_v = packed, bias_x, bias_v
        self.nlinks, self.hidden, self.activation = int(nlinks), int(hidden), int(activation)


def input_layer_supported(nlinks: int, hidden: int, nb: int) -> bool:
    return nlinks % 8 == 0 and 0 < hidden <= 256 and 0 < nb <= 256


def input_nb_pad(nb: int) -> int:
    return (int(nb) + 15) // 16 * 16


def su3_input_pack(w_x: Tensor, w_v: Tensor, b_x: Tensor, b_v: Tensor, activation: str) -> InputPack:
    """w_x, w_v: [hidden, 8 * nlinks] as nn.Linear stores them (x part, force part)"""
    _need_cuda(w_x, w_v)
    hidden, K = int(w_x.shape[0]), int(w_x.shape[1])
    if w_v.shape != w_x.shape or w_v.dtype != w_x.dtype or K % 64 != 0:
        raise L2BError('input weights must share shape and dtype, with in_features a multiple of 64')
    if activation not in INPUT_ACTIVATIONS:
        raise L2BError(f'activation {activation!r} not supported by the fused input layer')
    nlinks = K // 8
    nbytes = int(_lib._lib.l2b_su3_input_packed_bytes(nlinks, hidden))
    if nbytes == 0:
        raise L2BError(f'fused input layer does not support nlinks={nlinks}, hidden={hidden}')
    packed = torch.empty(nbytes, dtype=torch.uint8, device=w_x.device)
    call('l2b_su3_input_pack', _ptr(w_x.detach().contiguous()), _ptr(w_v.detach().contiguous()), _net_dt(w_x.dtype),
         _ptr(packed), nlinks, hidden, _stream())
    f32 = lambda a: a.detach().to(torch.float32).reshape(-1).contiguous()  # noqa: E731
    return InputPack(packed, f32(b_x), f32(b_v), nlinks, hidden, INPUT_ACTIVATIONS[activation])


def su3_project_vec_planar_lm(x_planar: Tensor, out: Optional[Tensor] = None) -> Tensor:
    """su3_to_vec(projectSU(x)) of a planar field as the link-major bf16 operand image of `su3_input_layer`:
    [4 V links, nb_pad, 8].  `out`: a buffer from a previous call (its pad rows are already zero)."""
    x, nb, dims = _su3_field(x_planar)
    nbp = input_nb_pad(nb)
    V = int(dims[0] * dims[1] * dims[2] * dims[3])
    if out is None:
        out = (torch.zeros if nbp != nb else torch.empty)((4 * V, nbp, 8), dtype=torch.bfloat16, device=x.device)
    call('l2b_su3_project_vec_planar_lm', _ptr(x), _ptr(out), nb, nbp, dims4(dims), L2B_F64, _stream())
    return out


def su3_input_layer(act_x: Tensor, act_f: Tensor, pack: InputPack, nb: int) -> Tensor:
    """z [nb, hidden] bf16 = act(W_x vec_x + b_x + W_v vec_f + b_v) from two link-major images"""
    _need_cuda(act_x, act_f)
    nbp = int(act_x.shape[1])
    if act_x.shape != act_f.shape or act_x.dtype != torch.bfloat16 or act_f.dtype != torch.bfloat16:
        raise L2BError('act_x / act_f must be bf16 images of the same shape')
    if int(act_x.shape[0]) != pack.nlinks or nbp < nb:
        raise L2BError(f'activation image {tuple(act_x.shape)} does not match nlinks={pack.nlinks}, nb={nb}')
    z = torch.empty((nb, pack.hidden), dtype=torch.bfloat16, device=act_x.device)
    nws = int(_lib._lib.l2b_su3_input_ws_bytes(nbp, pack.hidden))
    ws = _workspace(nws, act_x.device)
    call('l2b_su3_input_layer', _ptr(act_x), _ptr(act_f), _ptr(pack.packed), _ptr(pack.bias_x), _ptr(pack.bias_v),
         pack.activation, _ptr(z), int(nb), nbp, pack.nlinks, pack.hidden, _ptr(ws), nws, _stream())
    return z


# ---------------------------------------------------------------------------
# dense layers: the general bf16 tensor-core GEMM (include/l2b.h, l2b_gemm_bf16)
# ---------------------------------------------------------------------------
_ACT_CODES = {None: 0, 'identity': 0, 'tanh': 1, 'relu': 2, 'swish': 3, 'leaky_relu': 4, 'elu': 5}


def _pad2(t: Tensor, rows: int, cols: int) -> Tensor:
    """zero-pad a 2-D matrix to [rows, cols]; a no-op (no copy) for the production shapes"""
    r, c = int(t.shape[0]), int(t.shape[1])
    if r == rows and c == cols:
        return t
    return torch.nn.functional.pad(t, (0, cols - c, 0, rows - r))


def _gemm_operand(t: Tensor) -> Tensor:
    if t.stride(1) != 1 or t.stride(0) % 8 != 0 or t.data_ptr() % 16 != 0 or t.stride(0) < t.shape[1]:
        t = t.contiguous()
    return t


def _r8(n: int) -> int:
    return (n + 7) // 8 * 8


def _gemm_fast(a_list, b_list, a_kmajor, b_kmajor, out, out_dtype, bias, act, accumulate, splits, seg_inner):
    """the common case of `gemm_bf16` -- aligned contiguous CUDA bf16 operands whose extents are multiples of 8, a
    fresh or aligned output -- with the least Python per launch (the training steps of small nets are host-bound);
    returns None when anything is unusual and the general path must look at it"""
    bf16 = torch.bfloat16
    a0, b0 = a_list[0], b_list[0]
    if a0.dim() != 2 or b0.dim() != 2:
        return None
    sa, sb = a0.shape, b0.shape
    lda, ldb = a0.stride(0), b0.stride(0)
    M, K = (sa[0], sa[1]) if a_kmajor else (sa[1], sa[0])
    N, Kb = (sb[0], sb[1]) if b_kmajor else (sb[1], sb[0])
    if K != Kb or N % 8 or (K % 8 if (a_kmajor or b_kmajor) else 0) or (0 if a_kmajor else M % 8):
        return None
    ptrs_a, ptrs_b = [], []
    seen: dict = {}
    for lst, ptrs, shp, ld in ((a_list, ptrs_a, sa, lda), (b_list, ptrs_b, sb, ldb)):
        for t in lst:
            p = seen.get(id(t))
            if p is None:
                if (t.dtype is not bf16 or not t.is_cuda or t.shape != shp or t.stride(1) != 1 or t.stride(0) != ld
                        or ld % 8 or ld < shp[1]):
                    return None
                p = t.data_ptr()
                if p % 16:
                    return None
                seen[id(t)] = p
            ptrs.append(p)
    if out is not None:
        if (out.dim() != 2 or out.shape[0] != M or out.shape[1] != N or out.stride(1) != 1 or out.stride(0) % 8
                or out.data_ptr() % 16 or out.dtype not in (bf16, torch.float32) or not out.is_cuda):
            return None
        dst = out
    else:
        if accumulate:
            return None
        dst = torch.empty((M, N), dtype=out_dtype, device=a0.device)
    if bias is not None:
        if bias.dtype is not torch.float32 or not bias.is_cuda or bias.numel() != N or not bias.is_contiguous():
            return None
    n = len(a_list)
    if splits <= 0:
        splits = int(_lib._lib.l2b_gemm_bf16_splits(M, N, K, n, int(b_kmajor)))
    ws, nws = None, 0
    if splits > 1:
        nws = int(_lib._lib.l2b_gemm_bf16_ws_bytes(M, N, splits))
        ws = _workspace(nws, a0.device)
    call('l2b_gemm_bf16', (c_void_p * n)(*ptrs_a), lda, int(a_kmajor), (c_void_p * n)(*ptrs_b), ldb, int(b_kmajor), n,
         int(seg_inner), M, N, K, _ptr(dst), L2B_BF16 if dst.dtype is bf16 else L2B_F32, dst.stride(0), int(accumulate),
         _ptr(bias), _ACT_CODES[act], max(1, splits), _ptr(ws), nws, _stream())
    return dst


def gemm_bf16(a: Sequence[Tensor] | Tensor, b: Sequence[Tensor] | Tensor, a_kmajor: bool, b_kmajor: bool, *,
              out: Optional[Tensor] = None, out_dtype: torch.dtype = torch.bfloat16, bias: Optional[Tensor] = None,
              act: Optional[str] = None, accumulate: bool = False, splits: int = 0, seg_inner: bool = False) -> Tensor:
    """D[m, n] = act(sum_seg sum_k A_seg(m, k) B_seg(n, k) + bias[n]) on the tensor cores (bf16 in, fp32 accumulate).

    a / b: one matrix or a list of up to 32 same-shaped matrices (segments summed in one launch).  K-major
    operands are stored [MN, K], MN-major ones [K, MN]; `out` may be a (row-strided) view, bf16 or fp32.
    Extents that are not multiples of 8 (16-byte units) are zero-padded here (a copy); the production shapes
    never are."""
    a_list = [a] if isinstance(a, Tensor) else list(a)
    b_list = [b] if isinstance(b, Tensor) else list(b)
    if len(a_list) != len(b_list) or not 1 <= len(a_list) <= 32:
        raise L2BError('gemm_bf16: 1..32 (A, B) segment pairs')
    fast = _gemm_fast(a_list, b_list, a_kmajor, b_kmajor, out, out_dtype, bias, act, accumulate, splits, seg_inner)
    if fast is not None:
        return fast
    # the segments of an fp32 product repeat the same few planes: every distinct tensor is checked / padded once
    ua = {id(t): t for t in a_list}
    ub = {id(t): t for t in b_list}
    _need_cuda(*ua.values(), *ub.values(), out, bias)
    a0, b0 = a_list[0], b_list[0]
    for t in (*ua.values(), *ub.values()):
        if t.dtype != torch.bfloat16 or t.dim() != 2:
            raise L2BError(f'gemm operands must be 2-D bfloat16 matrices (got {t.dtype}, {tuple(t.shape)})')
    if any(t.shape != a0.shape for t in ua.values()) or any(t.shape != b0.shape for t in ub.values()):
        raise L2BError('gemm_bf16: segments must share shapes')
    M, K = (int(d) for d in (a0.shape if a_kmajor else a0.shape[::-1]))
    N, Kb = (int(d) for d in (b0.shape if b_kmajor else b0.shape[::-1]))
    if Kb != K:
        # one side may carry the zero padding of a previous call (a sliced fp32 result re-split to a multiple of 8)
        if _r8(Kb) != _r8(K):
            raise L2BError(f'gemm_bf16: contraction lengths differ ({K} vs {Kb})')
        K = max(K, Kb)
    Kp = _r8(K) if (a_kmajor or b_kmajor) else K
    Mp = M if a_kmajor else _r8(M)
    Np = _r8(N)
    pa = {k: _gemm_operand(_pad2(t, M, Kp) if a_kmajor else _pad2(t, Kp, Mp)) for k, t in ua.items()}
    pb = {k: _gemm_operand(_pad2(t, Np, Kp) if b_kmajor else _pad2(t, Kp, Np)) for k, t in ub.items()}
    if len({int(t.stride(0)) for t in pa.values()}) != 1:
        pa = {k: t.contiguous() for k, t in pa.items()}
    if len({int(t.stride(0)) for t in pb.values()}) != 1:
        pb = {k: t.contiguous() for k, t in pb.items()}
    a_list = [pa[id(t)] for t in a_list]
    b_list = [pb[id(t)] for t in b_list]
    dev = a_list[0].device
    direct = (out is not None and Mp == M and Np == N and out.dim() == 2 and tuple(out.shape) == (M, N)
              and out.stride(1) == 1 and out.stride(0) % 8 == 0 and out.data_ptr() % 16 == 0
              and out.dtype in (torch.bfloat16, torch.float32))
    if direct:
        dst = out
    else:
        if accumulate:
            raise L2BError('gemm_bf16: accumulate needs an aligned bf16 / fp32 `out` of an unpadded shape')
        dst = torch.empty((Mp, Np), dtype=out_dtype, device=dev)
    if bias is not None:
        bias = bias.detach().to(torch.float32).reshape(-1).contiguous()
        if Np != N:
            bias = torch.nn.functional.pad(bias, (0, Np - N))
    if splits <= 0:
        splits = int(_lib._lib.l2b_gemm_bf16_splits(Mp, Np, Kp, len(a_list), int(b_kmajor)))
    splits = max(1, splits)
    nws = int(_lib._lib.l2b_gemm_bf16_ws_bytes(Mp, Np, splits))
    ws = _workspace(nws, dev) if nws else None
    ap = (c_void_p * len(a_list))(*[t.data_ptr() for t in a_list])
    bp = (c_void_p * len(b_list))(*[t.data_ptr() for t in b_list])
    call('l2b_gemm_bf16', ap, int(a_list[0].stride(0)), int(a_kmajor), bp, int(b_list[0].stride(0)), int(b_kmajor),
         len(a_list), int(seg_inner), Mp, Np, Kp, _ptr(dst), L2B_BF16 if dst.dtype == torch.bfloat16 else L2B_F32,
         int(dst.stride(0)), int(accumulate), _ptr(bias), _ACT_CODES[act], splits, _ptr(ws), nws, _stream())
    if direct:
        return out
    res = dst if (Mp == M and Np == N) else dst[:M, :N].contiguous()
    if out is not None:
        out.copy_(res)
        return out
    return res


def split_bf16x3(x: Tensor) -> Tensor:
    """x [R, C] float32 -> [3, R, C8] bfloat16 with x = x1 + x2 + x3 (C8 = C rounded up to a multiple of 8, zero
    padded): the operands of the fp32-accurate GEMM `gemm_f32`"""
    _need_cuda(x)
    if x.dtype != torch.float32 or x.dim() != 2:
        raise L2BError(f'split_bf16x3 expects a 2-D float32 matrix (got {x.dtype}, {tuple(x.shape)})')
    if x.stride(1) != 1:
        x = x.contiguous()
    R, C = int(x.shape[0]), int(x.shape[1])
    out = torch.empty((3, R, _r8(C)), dtype=torch.bfloat16, device=x.device)
    call('l2b_split_bf16x3', _ptr(x), R, C, int(x.stride(0)), _ptr(out), int(out.shape[2]), _stream())
    return out


# (a_i, b_j) pairs (0-based planes) with i + j <= planes - 1, smallest contributions first
_X3_PAIRS = ((2, 0), (1, 1), (0, 2), (1, 0), (0, 1), (0, 0))
_X2_PAIRS = ((1, 0), (0, 1), (0, 0))


def gemm_f32(a3: Sequence[Tensor] | Tensor, b3: Sequence[Tensor] | Tensor, a_kmajor: bool, b_kmajor: bool, *,
             out: Optional[Tensor] = None, bias: Optional[Tensor] = None, act: Optional[str] = None,
             accumulate: bool = False) -> Tensor:
    """fp32-accurate D = act(sum_u A_u B_u^T + bias) on the bf16 tensor cores: every operand comes as its bf16x3 split
    (`split_bf16x3`, [3, R, C8]); a3 / b3 may be lists (the updates whose dW share a weight matrix).  Six products per
    pair, at most 30 segments per launch (the rest accumulates in further launches)."""
    a_list = [a3] if isinstance(a3, Tensor) else list(a3)
    b_list = [b3] if isinstance(b3, Tensor) else list(b3)
    if len(a_list) > 5 and (act is not None or bias is not None):
        raise L2BError('gemm_f32: bias / activation need all segments in one launch (at most 5 operand pairs)')
    # two planes per operand (bf16x2: 16 mantissa bits, three products) is the TF32-class variant used for the
    # convolutions on request; three planes (six products) is fp32-accurate
    pairs = _X3_PAIRS if min(int(a_list[0].shape[0]), int(b_list[0].shape[0])) >= 3 else _X2_PAIRS
    res = out
    a_planes = [t.unbind(0) for t in a_list]                 # one view per plane, shared by the segments that use it
    b_planes = [t.unbind(0) for t in b_list]
    for i in range(0, len(a_list), 5):
        segs_a, segs_b = [], []
        for a_, b_ in zip(a_planes[i:i + 5], b_planes[i:i + 5]):
            for ia, ib in pairs:
                segs_a.append(a_[ia])
                segs_b.append(b_[ib])
        last = i + 5 >= len(a_list)
        res = gemm_bf16(segs_a, segs_b, a_kmajor, b_kmajor, out=res, out_dtype=torch.float32,
                        bias=bias if last else None, act=act if last else None, accumulate=accumulate or i > 0,
                        seg_inner=True)
    return res


def linear_fwd(x: Tensor, w: Tensor, bias: Optional[Tensor] = None, act: Optional[str] = None,
               out_dtype: torch.dtype = torch.bfloat16) -> Tensor:
    """act(x W^T + b): x [nb, in], W [out, in] (nn.Linear layout), both contracted over their contiguous axis"""
    return gemm_bf16(x, w, True, True, bias=bias, act=act, out_dtype=out_dtype)


def linear_dx(gy: Sequence[Tensor] | Tensor, w: Sequence[Tensor] | Tensor, out_dtype: torch.dtype = torch.bfloat16) -> Tensor:
    """dX = sum_seg dY_seg W_seg: dY [nb, out] K-major, W [out, in] contracted over its ROWS (MN-major) -- no W^T copy"""
    return gemm_bf16(gy, w, True, False, out_dtype=out_dtype)


def linear_dw(gy: Tensor, x: Tensor, out: Optional[Tensor] = None, out_dtype: torch.dtype = torch.float32,
              accumulate: bool = False) -> Tensor:
    """dW = dY^T X: dY [r, out], X [r, in], both contracted over their rows (MN-major) -- no transposed copies"""
    return gemm_bf16(gy, x, False, False, out=out, out_dtype=out_dtype, accumulate=accumulate)


# ---------------------------------------------------------------------------
# U(1) xnet convolution stack: periodic-padding gather / its adjoint, pooling + activation (csrc/l2b_conv.cu)
# ---------------------------------------------------------------------------
def _strides4(t: Tensor, nchw: bool):
    """element strides (batch, channel, row, column) of a 4-D activation stored NCHW or NHWC"""
    sb, s1, s2, s3 = (int(v) for v in t.stride())
    return (ctypes.c_longlong * 4)(*((sb, s1, s2, s3) if nchw else (sb, s3, s1, s2)))


def conv_im2col(x: Tensor, n: int, nchw: bool, planes: int, tap_major: bool = False) -> Tensor:
    """col[planes, nb OH OW, K8] bf16 of a periodic-padding convolution block (see include/l2b.h); x float32 / bf16,
    [nb, C, H, W] (nchw) or [nb, H, W, C]; columns k = (ci, kh, kw), or (kh, kw, ci) with `tap_major`"""
    _need_cuda(x)
    if x.dtype not in (torch.float32, torch.bfloat16) or x.dim() != 4:
        raise L2BError(f'conv_im2col expects a 4-D float32 / bfloat16 activation (got {x.dtype}, {tuple(x.shape)})')
    nb = int(x.shape[0])
    C, H, W = (int(x.shape[1]), int(x.shape[2]), int(x.shape[3])) if nchw else (int(x.shape[3]), int(x.shape[1]), int(x.shape[2]))
    OH, OW, K8 = H + n - 1, W + n - 1, _r8(C * n * n)
    col = torch.empty((planes, nb * OH * OW, K8), dtype=torch.bfloat16, device=x.device)
    call('l2b_conv_im2col', _ptr(x), _net_dt(x.dtype), nb, C, H, W, int(n), _strides4(x, nchw), _ptr(col), int(planes),
         int(bool(tap_major)), _stream())
    return col


def conv_col2im(dcol: Tensor, like: Tensor, n: int, nchw: bool, tap_major: bool = False) -> Tensor:
    """adjoint of conv_im2col: dcol [nb OH OW, ld] float32 / bf16 -> float32 gradient shaped and laid out like `like`"""
    _need_cuda(dcol)
    if dcol.dtype not in (torch.float32, torch.bfloat16) or dcol.dim() != 2 or dcol.stride(1) != 1:
        raise L2BError('conv_col2im expects a row-major 2-D float32 / bfloat16 matrix')
    nb = int(like.shape[0])
    C, H, W = (int(like.shape[1]), int(like.shape[2]), int(like.shape[3])) if nchw else (int(like.shape[3]), int(like.shape[1]), int(like.shape[2]))
    din = torch.empty(like.shape, dtype=torch.float32, device=dcol.device)
    call('l2b_conv_col2im', _ptr(dcol), _net_dt(dcol.dtype), int(dcol.stride(0)), nb, C, H, W, int(n), _ptr(din),
         _strides4(din, nchw), int(bool(tap_major)), _stream())
    return din


def pool_act(x: Tensor, pool: int, act: Optional[str]):
    """MaxPool2d(pool) + activation on an NHWC activation; returns (y, idx, pre) (pre only for swish)"""
    _need_cuda(x)
    x = x.contiguous()
    nb, H, W, C = (int(v) for v in x.shape)
    y = torch.empty((nb, H // pool, W // pool, C), dtype=x.dtype, device=x.device)
    idx = torch.empty(y.shape, dtype=torch.uint8, device=x.device)
    pre = torch.empty(y.shape, dtype=torch.float32, device=x.device) if act == 'swish' else None
    call('l2b_pool_act', _ptr(x), _net_dt(x.dtype), nb, H, W, C, int(pool), _ACT_CODES[act], _ptr(y), _ptr(idx), _ptr(pre),
         _stream())
    return y, idx, pre


def pool_act_bwd(gy: Tensor, y: Tensor, pre: Optional[Tensor], idx: Tensor, in_shape, pool: int, act: Optional[str]) -> Tensor:
    nb, H, W, C = (int(v) for v in in_shape)
    gy = gy.to(torch.float32).contiguous()
    gx = torch.empty((nb, H, W, C), dtype=torch.float32, device=gy.device)
    call('l2b_pool_act_bwd', _ptr(gy), _ptr(y), _net_dt(y.dtype), _ptr(pre), _ptr(idx), nb, H, W, C, int(pool),
         _ACT_CODES[act], _ptr(gx), _stream())
    return gx


# ---------------------------------------------------------------------------
# vnet output heads on the tensor cores, fused with the momentum update
# ---------------------------------------------------------------------------
class HeadsPack:
    """bf16 UMMA tile image of (W_s, W_t, W_q) plus the per-column epilogue constants
    (include/l2b.h, l2b_vnet_pack_heads)"""
    __slots__ = ('packed', 'bias', 'scale_s', 'scale_q', 'scale_t', 'xdim', 'hidden')

    def __init__(self, packed, bias, scale_s, scale_q, scale_t, xdim, hidden):
        self.packed, self.bias, self.scale_s, self.scale_q = packed, bias, scale_s, scale_q
        self.scale_t, self.xdim, self.hidden = float(scale_t), int(xdim), int(hidden)


def heads_supported(hidden: int) -> bool:
    return hidden % 8 == 0 and 0 < hidden <= 256


def vnet_pack_heads(w_s: Tensor, w_t: Tensor, w_q: Tensor, b_s: Tensor, b_t: Tensor, b_q: Tensor,
                    coeff_s: Tensor, coeff_q: Tensor, nw_s: float = 1.0, nw_t: float = 1.0,
                    nw_q: float = 1.0) -> HeadsPack:
    """weights [xdim, hidden] as nn.Linear stores them; coeff_* are ScaledTanh.coeff [1, xdim]"""
    _need_cuda(w_s, w_t, w_q)
    xdim, hidden = int(w_s.shape[0]), int(w_s.shape[1])
    if w_t.shape != w_s.shape or w_q.shape != w_s.shape or w_t.dtype != w_s.dtype or w_q.dtype != w_s.dtype:
        raise L2BError('head weights must share shape and dtype')
    ws_, wt_, wq_ = (w.detach().contiguous() for w in (w_s, w_t, w_q))
    nbytes = int(_lib._lib.l2b_vnet_heads_packed_bytes(xdim, hidden))
    packed = torch.empty(nbytes, dtype=torch.uint8, device=w_s.device)
    call('l2b_vnet_pack_heads', _ptr(ws_), _ptr(wt_), _ptr(wq_), _net_dt(w_s.dtype), _ptr(packed), xdim, hidden,
         _stream())
    f32 = lambda a: a.detach().to(torch.float32).reshape(-1).contiguous()  # noqa: E731
    bias = torch.stack([f32(b_s), f32(b_t), f32(b_q)]).contiguous()
    scale_s = (float(nw_s) * f32(coeff_s).exp()).contiguous()
    scale_q = (float(nw_q) * f32(coeff_q).exp()).contiguous()
    return HeadsPack(packed, bias, scale_s, scale_q, nw_t, xdim, hidden)


def su3_heads_vupdate(z: Tensor, pack: HeadsPack, v: Tensor, force: Tensor, eps, sign: int,
                      want_stq: bool = False):
    """(v', logdet[, stq]) with s, t, q = heads(z) never materialised (unless want_stq:
    f32 [3, nb, xdim])"""
    _need_cuda(z, v, force)
    if z.dtype != torch.bfloat16:
        z = z.to(torch.bfloat16)
    z = z.contiguous()
    nb = int(z.shape[0])
    if int(z.shape[1]) != pack.hidden:
        raise L2BError(f'z has {z.shape[1]} features, the packed heads expect {pack.hidden}')
    if v.dtype != torch.complex128 or force.dtype != torch.complex128:
        raise L2BError('v and force must be complex128')
    v, force = v.contiguous(), force.contiguous()
    if v.numel() != nb * pack.xdim or force.numel() != nb * pack.xdim:
        raise L2BError(f'v / force must have {nb * pack.xdim} complex entries')
    out = torch.empty_like(v)
    logdet = torch.empty(nb, dtype=torch.float64, device=v.device)
    stq = torch.empty((3, nb, pack.xdim), dtype=torch.float32, device=v.device) if want_stq else None
    nws = int(_lib._lib.l2b_vnet_heads_ws_bytes(nb, pack.xdim))
    ws = _workspace(nws, v.device)
    ev, ep, _keep = _eps_args(eps, torch.float64)
    call('l2b_su3_heads_vupdate', _ptr(z), _ptr(pack.packed), _ptr(pack.bias[0]), _ptr(pack.bias[1]),
         _ptr(pack.bias[2]), _ptr(pack.scale_s), _ptr(pack.scale_q), pack.scale_t, _ptr(v), _ptr(force), ev, ep,
         int(sign), _ptr(out), _ptr(logdet), _ptr(stq), nb, pack.xdim, pack.hidden, _ptr(ws), nws, _stream())
    return (out, logdet, stq) if want_stq else (out, logdet)


def su3_heads_vupdate_pair(z: Tensor, pack: HeadsPack, v: Tensor, force: Tensor, eps1, sign1: int, eps2, sign2: int,
                           negate_between: bool = False):
    """(v'', logdet_1 + logdet_2): two consecutive momentum updates that share (s, t, q) and the force -- no link
    update in between, one vnet for all layers -- in one pass (include/l2b.h, l2b_su3_heads_vupdate_pair)"""
    _need_cuda(z, v, force)
    if z.dtype != torch.bfloat16:
        z = z.to(torch.bfloat16)
    z = z.contiguous()
    nb = int(z.shape[0])
    if int(z.shape[1]) != pack.hidden:
        raise L2BError(f'z has {z.shape[1]} features, the packed heads expect {pack.hidden}')
    if v.dtype != torch.complex128 or force.dtype != torch.complex128:
        raise L2BError('v and force must be complex128')
    v, force = v.contiguous(), force.contiguous()
    if v.numel() != nb * pack.xdim or force.numel() != nb * pack.xdim:
        raise L2BError(f'v / force must have {nb * pack.xdim} complex entries')
    out = torch.empty_like(v)
    logdet = torch.empty(nb, dtype=torch.float64, device=v.device)
    nws = int(_lib._lib.l2b_vnet_heads_ws_bytes(nb, pack.xdim))
    ws = _workspace(nws, v.device)
    e1, p1, _k1 = _eps_args(eps1, torch.float64)
    e2, p2, _k2 = _eps_args(eps2, torch.float64)
    call('l2b_su3_heads_vupdate_pair', _ptr(z), _ptr(pack.packed), _ptr(pack.bias[0]), _ptr(pack.bias[1]),
         _ptr(pack.bias[2]), _ptr(pack.scale_s), _ptr(pack.scale_q), pack.scale_t, _ptr(v), _ptr(force), e1, p1,
         int(sign1), e2, p2, int(sign2), int(bool(negate_between)), _ptr(out), _ptr(logdet), nb, pack.xdim, pack.hidden,
         _ptr(ws), nws, _stream())
    return out, logdet


# ---------------------------------------------------------------------------
# U(1): output heads fused with the update (CUDA-core kernel, hidden <= 32)
# ---------------------------------------------------------------------------
def u1_heads_supported(hidden: int) -> bool:
    return 0 < hidden <= 32


def u1_heads_update(mode: int, z: Tensor, head_params, nw, a: Tensor, b: Tensor, eps, sign: int,
                    mask: Optional[Tensor] = None, use_ncp: bool = True):
    """mode 0: (v', logdet) = vupdate(a = v, b = force); mode 1: (x', logdet) = xupdate(a = x, b = v, mask).
    head_params = (W_s, b_s, c_s, W_t, b_t, W_q, b_q, c_q) as `LeapfrogLayer.head_params()`, nw = (s, t, q)."""
    ws_, bs, cs, wt, bt, wq, bq, cq = (p.detach().contiguous() for p in head_params)
    _need_cuda(z, a, b, ws_)
    dt = a.dtype
    nb = int(a.shape[0])
    a2, b2 = a.reshape(nb, -1).contiguous(), b.reshape(nb, -1).to(dt).contiguous()
    xdim, hidden = int(a2.shape[1]), int(ws_.shape[1])
    if any(p.dtype != dt for p in (ws_, bs, cs, wt, bt, wq, bq, cq)) or tuple(ws_.shape) != (xdim, hidden):
        raise L2BError('head parameters must have the field dtype and shape [xdim, hidden]')
    z = z.detach().to(dt).contiguous()
    if tuple(z.shape) != (nb, hidden):
        raise L2BError(f'z must be [{nb}, {hidden}] (got {tuple(z.shape)})')
    if mode == 1:
        if mask is None:
            raise L2BError('the x-update needs a mask')
        mask = mask.to(torch.float32).reshape(-1).contiguous()
    out = torch.empty_like(a2)
    logdet = torch.empty(nb, dtype=dt, device=a.device)
    nws = int(_lib._lib.l2b_u1_heads_ws_bytes(nb, xdim))
    ws = _workspace(nws, a.device)
    ev, ep, _keep = _eps_args(eps, dt)
    call('l2b_u1_heads_update', int(mode), _ptr(z), hidden, _ptr(ws_), _ptr(wt), _ptr(wq), _ptr(bs), _ptr(bt), _ptr(bq),
         _ptr(cs.reshape(-1)), _ptr(cq.reshape(-1)), float(nw[0]), float(nw[1]), float(nw[2]), _ptr(a2), _ptr(b2),
         _ptr(mask), ev, ep, int(sign), int(bool(use_ncp)), _ptr(out), _ptr(logdet), nb, xdim, _dt(a2), _ptr(ws), nws,
         _stream())
    return out, logdet


def u1_input_supported(units0: int) -> bool:
    return 0 < units0 <= 16


def u1_input_layer(mode: int, x: Tensor, v: Tensor, w_x: Tensor, b_x: Tensor, w_v: Tensor, b_v: Tensor,
                   mask: Optional[Tensor] = None) -> Tensor:
    """pre-activation of InputLayer without conv stack: mode 1 (xnet) W_x . cat(cos(mask x), sin(mask x)) + b_x +
    W_v . v + b_v, mode 0 (vnet) W_x . x + b_x + W_v . v + b_v; x and v are read once.  -> [nb, units]"""
    _need_cuda(x, v, w_x, w_v)
    dt = x.dtype
    nb = int(x.shape[0])
    x2, v2 = x.reshape(nb, -1).contiguous(), v.reshape(nb, -1).to(dt).contiguous()
    xdim, units = int(x2.shape[1]), int(w_x.shape[0])
    w_x, b_x, w_v, b_v = (p.detach().contiguous() for p in (w_x, b_x, w_v, b_v))
    if any(p.dtype != dt for p in (w_x, b_x, w_v, b_v)):
        raise L2BError('input-layer parameters must have the field dtype')
    if tuple(w_x.shape) != (units, (2 if mode == 1 else 1) * xdim) or tuple(w_v.shape) != (units, xdim):
        raise L2BError(f'unexpected input-layer weight shapes {tuple(w_x.shape)}, {tuple(w_v.shape)}')
    if mode == 1:
        if mask is None:
            raise L2BError('the xnet input needs the mask')
        mask = mask.to(torch.float32).reshape(-1).contiguous()
    pre = torch.empty((nb, units), dtype=dt, device=x.device)
    nws = int(_lib._lib.l2b_u1_input_ws_bytes(nb, xdim))
    ws = _workspace(nws, x.device)
    call('l2b_u1_input_layer', int(mode), _ptr(x2), _ptr(v2), _ptr(mask), _ptr(w_x), _ptr(b_x), _ptr(w_v), _ptr(b_v),
         units, _ptr(pre), nb, xdim, _dt(x2), _ptr(ws), nws, _stream())
    return pre


def su3_heads_vupdate_bwd(v: Tensor, force: Tensor, stq: Tensor, pack: HeadsPack, eps, sign: int, gv_out: Tensor,
                          glogdet: Optional[Tensor], gpre_dtype: torch.dtype, want_gforce: bool = True):
    """element-wise adjoint of su3_heads_vupdate -> (gv, gforce, gpre[3, nb, xdim], colsum[5, xdim], geps[nb]);
    colsum rows: the three heads' bias gradients, then sum_b gs*s and sum_b gq*q (the ScaledTanh.coeff gradients)"""
    _need_cuda(v, force, stq, gv_out)
    nb, xdim = int(stq.shape[1]), int(stq.shape[2])
    v2, f2 = v.contiguous(), force.contiguous()
    go = gv_out.to(torch.complex128).contiguous()
    if v2.numel() != nb * xdim or f2.numel() != nb * xdim or go.numel() != nb * xdim:
        raise L2BError('v, force and gv_out must have nb * xdim complex entries')
    stq = stq.to(torch.float32).contiguous()
    gl = None if glogdet is None else glogdet.to(torch.float64).contiguous()
    gv = torch.empty_like(v2)
    gf = torch.empty_like(v2) if want_gforce else None
    gpre = torch.empty((3, nb, xdim), dtype=gpre_dtype, device=v.device)
    colsum = torch.empty((5, xdim), dtype=torch.float32, device=v.device)
    geps = torch.empty(nb, dtype=torch.float64, device=v.device)
    nws = nb * ((xdim + 255) // 256) * 8
    ws = _workspace(nws, v.device)
    if gpre_dtype not in (torch.float32, torch.bfloat16):
        raise L2BError(f'gpre dtype must be float32 or bfloat16 (got {gpre_dtype})')
    ev, ep, _keep = _eps_args(eps, torch.float64)
    call('l2b_su3_heads_vupdate_bwd', _ptr(v2), _ptr(f2), _ptr(stq), _ptr(pack.scale_s), _ptr(pack.scale_q), pack.scale_t,
         ev, ep, int(sign), _ptr(go), _ptr(gl), _ptr(gv), _ptr(gf), _ptr(gpre), _net_dt(gpre_dtype), _ptr(colsum),
         _ptr(geps), nb, xdim, _ptr(ws), nws, _stream())
    return gv, gf, gpre, colsum, geps
