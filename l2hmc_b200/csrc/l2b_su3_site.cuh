// l2b_su3_site.cuh -- lattice geometry, the internal planar (SoA) link layout
// and the per-link / per-site stencil bodies of the SU(3) kernels.
//
// Boundary layout (what callers hand in; reference configs.py:501-507):
//     x[b, mu, t, x, y, z, i, j]  complex, interleaved re/im  ("AoS", 144 B/link in c128)
// Internal layout used by every stencil kernel ("SoA"):
//     U[b][mu][e][site]  one complex<T> per entry, e = 3*i + j, site = ((t*X + x)*Y + y)*Z + z
// so a warp that owns 32 consecutive sites reads each matrix entry with ONE fully
// coalesced 128-bit load per lane, and a neighbour in +-z is the same row shifted by
// 16 B.  The conversion happens once per trajectory, not per leapfrog step.
//
// The bodies are __host__ __device__ so tests/hostemu can run them on the CPU.
#pragma once
#include "l2b_su3_math.cuh"

namespace l2b {

#if defined(__CUDACC__)
template <typename T> struct cplx_of;
template <> struct cplx_of<double> { using type = double2; };
template <> struct cplx_of<float> { using type = float2; };
#else
template <typename T> struct cplx_host { T x, y; };
template <typename T> struct cplx_of { using type = cplx_host<T>; };
#endif

struct Lat {
  int L[4];        // T, X, Y, Z
  int stride[4];   // site stride of each direction
  int V;
};

L2B_HD Lat make_lat(int T, int X, int Y, int Z) {
  Lat l;
  l.L[0] = T; l.L[1] = X; l.L[2] = Y; l.L[3] = Z;
  l.stride[3] = 1; l.stride[2] = Z; l.stride[1] = Y * Z; l.stride[0] = X * Y * Z;
  l.V = T * X * Y * Z;
  return l;
}

L2B_HD void site_coords(const Lat& l, int site, int c[4]) {
  c[3] = site % l.L[3]; site /= l.L[3];
  c[2] = site % l.L[2]; site /= l.L[2];
  c[1] = site % l.L[1]; site /= l.L[1];
  c[0] = site;
}

// site index of n + mu_hat / n - mu_hat (periodic)
L2B_HD int site_fwd(const Lat& l, int site, const int c[4], int mu) {
  return (c[mu] == l.L[mu] - 1) ? site - (l.L[mu] - 1) * l.stride[mu] : site + l.stride[mu];
}
L2B_HD int site_bwd(const Lat& l, int site, const int c[4], int mu) {
  return (c[mu] == 0) ? site + (l.L[mu] - 1) * l.stride[mu] : site - l.stride[mu];
}

#if defined(__CUDA_ARCH__)
#define L2B_LDG(p) __ldg(p)
#else
#define L2B_LDG(p) (*(p))
#endif

// plane pointer of (chain b, direction mu): 9 rows of V complex numbers
template <typename C>
L2B_HD const C* soa_plane(const C* f, const Lat& l, int b, int mu) {
  return f + ((size_t)b * 4 + mu) * 9 * (size_t)l.V;
}
template <typename C>
L2B_HD C* soa_plane(C* f, const Lat& l, int b, int mu) {
  return f + ((size_t)b * 4 + mu) * 9 * (size_t)l.V;
}

template <typename T, typename C>
L2B_HD void soa_load(Mat3<T>& m, const C* plane, int V, int site) {
  L2B_UNROLL
  for (int e = 0; e < 9; ++e) {
    const C v = L2B_LDG(plane + (size_t)e * V + site);
    m.re[e] = v.x; m.im[e] = v.y;
  }
}

template <typename T, typename C>
L2B_HD void soa_store(C* plane, int V, int site, const Mat3<T>& m) {
  L2B_UNROLL
  for (int e = 0; e < 9; ++e) {
    C v; v.x = m.re[e]; v.y = m.im[e];
    plane[(size_t)e * V + site] = v;
  }
}

// ---------------------------------------------------------------------------
// G = U_mu(n) * A_mu(n),   A = sum of the six staples around link (mu, n):
//   A_mu(n) = sum_{nu != mu} [ U_nu(n+mu) U_mu(n+nu)^+ U_nu(n)^+
//                            + U_nu(n+mu-nu)^+ U_mu(n-nu)^+ U_nu(n-nu) ]
// Then  force = (beta/3) TAH(G)  (== projectTAH(dS/dU U^+), reference
// lattice/su3/pytorch/lattice.py:299-308, which gets it by autograd) and
// sum_{mu,n} Re tr G = 4 * sum_plaquettes Re tr P  (every plaquette is seen from
// each of its four links), which gives the Wilson action for free.
// ---------------------------------------------------------------------------
#if defined(__CUDA_ARCH__)
#define L2B_PREFETCH_L1(p) asm volatile("prefetch.global.L1 [%0];" ::"l"(p))
#define L2B_PREFETCH_L2(p) asm volatile("prefetch.global.L2 [%0];" ::"l"(p))
#else
#define L2B_PREFETCH_L1(p) ((void)(p))
#define L2B_PREFETCH_L2(p) ((void)(p))
#endif

template <int LEVEL, typename C>
L2B_HD void soa_prefetch(const C* plane, int V, int site) {
  L2B_UNROLL
  for (int e = 0; e < 9; ++e) {
    if (LEVEL == 1) L2B_PREFETCH_L1(plane + (size_t)e * V + site);
    else L2B_PREFETCH_L2(plane + (size_t)e * V + site);
  }
}

L2B_HD int sel4(int a0, int a1, int a2, int a3, int i) {
  // register-only replacement for a[i] with a runtime i (no local-memory array)
  return (i == 0) ? a0 : (i == 1) ? a1 : (i == 2) ? a2 : a3;
}

// PF = 1: while the staples of direction k are being multiplied, prefetch (to L1)
// the six matrices of direction k+1 -- latency hiding that costs no registers.
// WITH_LINK = false returns the staple sum A itself (the action's adjoint is
// dS/dU = -(beta/3) A^+), true returns G = U A.
template <typename T, typename C, int PF = 0, bool WITH_LINK = true>
L2B_HD void link_times_staples(Mat3<T>& g, const C* U, const Lat& l, int b, int mu, int site) {
  const int V = l.V;
  // coordinates, and for every direction d: site offset of a forward / backward hop
  int r = site;
  const int c3 = r % l.L[3]; r /= l.L[3];
  const int c2 = r % l.L[2]; r /= l.L[2];
  const int c1 = r % l.L[1]; r /= l.L[1];
  const int c0 = r;
  const int f0 = (c0 == l.L[0] - 1) ? -(l.L[0] - 1) * l.stride[0] : l.stride[0];
  const int f1 = (c1 == l.L[1] - 1) ? -(l.L[1] - 1) * l.stride[1] : l.stride[1];
  const int f2 = (c2 == l.L[2] - 1) ? -(l.L[2] - 1) * l.stride[2] : l.stride[2];
  const int f3 = (c3 == l.L[3] - 1) ? -(l.L[3] - 1) : 1;
  const int b0 = (c0 == 0) ? (l.L[0] - 1) * l.stride[0] : -l.stride[0];
  const int b1 = (c1 == 0) ? (l.L[1] - 1) * l.stride[1] : -l.stride[1];
  const int b2 = (c2 == 0) ? (l.L[2] - 1) * l.stride[2] : -l.stride[2];
  const int b3 = (c3 == 0) ? (l.L[3] - 1) : -1;
  const size_t plane_sz = (size_t)9 * V;
  const C* chain = U + (size_t)b * 4 * plane_sz;
  const C* pmu = chain + (size_t)mu * plane_sz;
  const int n_pmu = site + sel4(f0, f1, f2, f3, mu);
  Mat3<T> a, x, y, m;
  mat_zero(a);
  L2B_UNROLL
  for (int k = 1; k < 4; ++k) {
    const int nu = (mu + k) & 3;
    const C* pnu = chain + (size_t)nu * plane_sz;
    const int fnu = sel4(f0, f1, f2, f3, nu);
    const int bnu = sel4(b0, b1, b2, b3, nu);
    const int n_pnu = site + fnu;
    const int n_mnu = site + bnu;
    const int n_pmu_mnu = n_pmu + bnu;           // the mu hop does not change the nu coordinate
    if (PF == 1 && k < 3) {
      const int nu2 = (mu + k + 1) & 3;
      const C* pnu2 = chain + (size_t)nu2 * plane_sz;
      const int f2n = sel4(f0, f1, f2, f3, nu2), b2n = sel4(b0, b1, b2, b3, nu2);
      soa_prefetch<1>(pnu2, V, n_pmu);
      soa_prefetch<1>(pmu, V, site + f2n);
      soa_prefetch<1>(pnu2, V, site);
      soa_prefetch<1>(pnu2, V, n_pmu + b2n);
      soa_prefetch<1>(pmu, V, site + b2n);
      soa_prefetch<1>(pnu2, V, site + b2n);
    }
    // forward staple
    soa_load(x, pnu, V, n_pmu);                  // U_nu(n+mu)
    soa_load(y, pmu, V, n_pnu);                  // U_mu(n+nu)
    mat_mul<false, true, false>(m, x, y);        // U_nu(n+mu) U_mu(n+nu)^+
    soa_load(x, pnu, V, site);                   // U_nu(n)
    mat_mul<false, true, true>(a, m, x);         // a += m U_nu(n)^+
    // backward staple
    soa_load(x, pnu, V, n_pmu_mnu);              // U_nu(n+mu-nu)
    soa_load(y, pmu, V, n_mnu);                  // U_mu(n-nu)
    mat_mul<true, true, false>(m, x, y);         // U_nu(n+mu-nu)^+ U_mu(n-nu)^+
    soa_load(x, pnu, V, n_mnu);                  // U_nu(n-nu)
    mat_mul<false, false, true>(a, m, x);        // a += m U_nu(n-nu)
  }
  if (WITH_LINK) {
    soa_load(x, pmu, V, site);
    mat_mul<false, false, false>(g, x, a);
  } else {
    g = a;
  }
}

// Same staple sum and link product as link_times_staples (identical operation order, so identical
// bits), with a caller-supplied hook run right before staple direction HOOK_AT (1..3; 0 = before
// the first).  The force kernel's "early momentum" variants use it to put the 9 momentum loads in
// flight while the remaining staples are multiplied.
template <typename T, typename C, int HOOK_AT, typename Hook>
L2B_HD void link_times_staples_hook(Mat3<T>& g, const C* U, const Lat& l, int b, int mu, int site, Hook&& hook) {
  const int V = l.V;
  int r = site;
  const int c3 = r % l.L[3]; r /= l.L[3];
  const int c2 = r % l.L[2]; r /= l.L[2];
  const int c1 = r % l.L[1]; r /= l.L[1];
  const int c0 = r;
  const int f0 = (c0 == l.L[0] - 1) ? -(l.L[0] - 1) * l.stride[0] : l.stride[0];
  const int f1 = (c1 == l.L[1] - 1) ? -(l.L[1] - 1) * l.stride[1] : l.stride[1];
  const int f2 = (c2 == l.L[2] - 1) ? -(l.L[2] - 1) * l.stride[2] : l.stride[2];
  const int f3 = (c3 == l.L[3] - 1) ? -(l.L[3] - 1) : 1;
  const int b0 = (c0 == 0) ? (l.L[0] - 1) * l.stride[0] : -l.stride[0];
  const int b1 = (c1 == 0) ? (l.L[1] - 1) * l.stride[1] : -l.stride[1];
  const int b2 = (c2 == 0) ? (l.L[2] - 1) * l.stride[2] : -l.stride[2];
  const int b3 = (c3 == 0) ? (l.L[3] - 1) : -1;
  const size_t plane_sz = (size_t)9 * V;
  const C* chain = U + (size_t)b * 4 * plane_sz;
  const C* pmu = chain + (size_t)mu * plane_sz;
  const int n_pmu = site + sel4(f0, f1, f2, f3, mu);
  Mat3<T> a, x, y, m;
  mat_zero(a);
  if (HOOK_AT == 0) hook();
  L2B_UNROLL
  for (int k = 1; k < 4; ++k) {
    if (k == HOOK_AT) hook();
    const int nu = (mu + k) & 3;
    const C* pnu = chain + (size_t)nu * plane_sz;
    const int fnu = sel4(f0, f1, f2, f3, nu);
    const int bnu = sel4(b0, b1, b2, b3, nu);
    const int n_pnu = site + fnu;
    const int n_mnu = site + bnu;
    const int n_pmu_mnu = n_pmu + bnu;
    soa_load(x, pnu, V, n_pmu);
    soa_load(y, pmu, V, n_pnu);
    mat_mul<false, true, false>(m, x, y);
    soa_load(x, pnu, V, site);
    mat_mul<false, true, true>(a, m, x);
    soa_load(x, pnu, V, n_pmu_mnu);
    soa_load(y, pmu, V, n_mnu);
    mat_mul<true, true, false>(m, x, y);
    soa_load(x, pnu, V, n_mnu);
    mat_mul<false, false, true>(a, m, x);
  }
  soa_load(x, pmu, V, site);
  mat_mul<false, false, false>(g, x, a);
}

// ---------------------------------------------------------------------------
// Low-register form of the same staple sum: the second operand of every 3x3 product is streamed from
// memory ROW BY ROW (3 loads at a time) instead of being held as a whole matrix, so the peak is
// a + X + m + one row (120 registers) instead of a + X + Y + m (144).  Every output element sees the
// same fused multiply-adds in the same order as mat_mul, so the result is bit-identical to
// link_times_staples (tests/hostemu).  HOOK_AT / hook as in link_times_staples_hook.
// ---------------------------------------------------------------------------
template <typename T, typename C>
L2B_HD void soa_load_row(T rr[3], T ri[3], const C* plane, int V, int site, int row) {
  L2B_UNROLL
  for (int j = 0; j < 3; ++j) {
    const C v = L2B_LDG(plane + (size_t)(3 * row + j) * V + site);
    rr[j] = v.x; ri[j] = v.y;
  }
}

// m(:, k) = op(X)(: , :) . conj(Y(k, :))   for the streamed row k of Y;   ADJ_X: op = conjugate transpose
template <bool ADJ_X, typename T>
L2B_HD void col_from_row_adj(Mat3<T>& m, const Mat3<T>& x, const T yr[3], const T yi[3], int k) {
  L2B_UNROLL
  for (int i = 0; i < 3; ++i) {
    T sr = T(0), si = T(0);
    L2B_UNROLL
    for (int l = 0; l < 3; ++l) {
      const int ex = ADJ_X ? (3 * l + i) : (3 * i + l);
      const T ar = x.re[ex], ai = ADJ_X ? -x.im[ex] : x.im[ex];
      const T br = yr[l], bi = -yi[l];
      sr = fma(ar, br, sr);
      sr = fma(-ai, bi, sr);
      si = fma(ar, bi, si);
      si = fma(ai, br, si);
    }
    m.re[3 * i + k] = sr;
    m.im[3 * i + k] = si;
  }
}

template <typename T, typename C, int HOOK_AT, typename Hook>
L2B_HD void link_times_staples_lowreg(Mat3<T>& g, const C* U, const Lat& l, int b, int mu, int site, Hook&& hook) {
  const int V = l.V;
  int r = site;
  const int c3 = r % l.L[3]; r /= l.L[3];
  const int c2 = r % l.L[2]; r /= l.L[2];
  const int c1 = r % l.L[1]; r /= l.L[1];
  const int c0 = r;
  const int f0 = (c0 == l.L[0] - 1) ? -(l.L[0] - 1) * l.stride[0] : l.stride[0];
  const int f1 = (c1 == l.L[1] - 1) ? -(l.L[1] - 1) * l.stride[1] : l.stride[1];
  const int f2 = (c2 == l.L[2] - 1) ? -(l.L[2] - 1) * l.stride[2] : l.stride[2];
  const int f3 = (c3 == l.L[3] - 1) ? -(l.L[3] - 1) : 1;
  const int b0 = (c0 == 0) ? (l.L[0] - 1) * l.stride[0] : -l.stride[0];
  const int b1 = (c1 == 0) ? (l.L[1] - 1) * l.stride[1] : -l.stride[1];
  const int b2 = (c2 == 0) ? (l.L[2] - 1) * l.stride[2] : -l.stride[2];
  const int b3 = (c3 == 0) ? (l.L[3] - 1) : -1;
  const size_t plane_sz = (size_t)9 * V;
  const C* chain = U + (size_t)b * 4 * plane_sz;
  const C* pmu = chain + (size_t)mu * plane_sz;
  const int n_pmu = site + sel4(f0, f1, f2, f3, mu);
  Mat3<T> a, x, m;
  T rr[3], ri[3];
  mat_zero(a);
  if (HOOK_AT == 0) hook();
  L2B_UNROLL
  for (int k = 1; k < 4; ++k) {
    if (k == HOOK_AT) hook();
    const int nu = (mu + k) & 3;
    const C* pnu = chain + (size_t)nu * plane_sz;
    const int fnu = sel4(f0, f1, f2, f3, nu);
    const int bnu = sel4(b0, b1, b2, b3, nu);
    const int n_pnu = site + fnu;
    const int n_mnu = site + bnu;
    const int n_pmu_mnu = n_pmu + bnu;
    // forward staple: m = U_nu(n+mu) U_mu(n+nu)^+ ;  a += m U_nu(n)^+
    soa_load(x, pnu, V, n_pmu);
    L2B_UNROLL
    for (int j = 0; j < 3; ++j) {
      soa_load_row(rr, ri, pmu, V, n_pnu, j);
      col_from_row_adj<false>(m, x, rr, ri, j);
    }
    L2B_UNROLL
    for (int j = 0; j < 3; ++j) {                    // a(:, j) += m . conj(Z(j, :))
      soa_load_row(rr, ri, pnu, V, site, j);
      L2B_UNROLL
      for (int i = 0; i < 3; ++i) {
        T sr = a.re[3 * i + j], si = a.im[3 * i + j];
        L2B_UNROLL
        for (int q = 0; q < 3; ++q) {
          const T ar = m.re[3 * i + q], ai = m.im[3 * i + q];
          const T br = rr[q], bi = -ri[q];
          sr = fma(ar, br, sr);
          sr = fma(-ai, bi, sr);
          si = fma(ar, bi, si);
          si = fma(ai, br, si);
        }
        a.re[3 * i + j] = sr; a.im[3 * i + j] = si;
      }
    }
    // backward staple: m = U_nu(n+mu-nu)^+ U_mu(n-nu)^+ ;  a += m U_nu(n-nu)
    soa_load(x, pnu, V, n_pmu_mnu);
    L2B_UNROLL
    for (int j = 0; j < 3; ++j) {
      soa_load_row(rr, ri, pmu, V, n_mnu, j);
      col_from_row_adj<true>(m, x, rr, ri, j);
    }
    L2B_UNROLL
    for (int q = 0; q < 3; ++q) {                    // a += m(:, q) Z(q, :), q ascending: mat_mul's order per element
      soa_load_row(rr, ri, pnu, V, n_mnu, q);
      L2B_UNROLL
      for (int i = 0; i < 3; ++i) {
        const T ar = m.re[3 * i + q], ai = m.im[3 * i + q];
        L2B_UNROLL
        for (int j = 0; j < 3; ++j) {
          T sr = a.re[3 * i + j], si = a.im[3 * i + j];
          sr = fma(ar, rr[j], sr);
          sr = fma(-ai, ri[j], sr);
          si = fma(ar, ri[j], si);
          si = fma(ai, rr[j], si);
          a.re[3 * i + j] = sr; a.im[3 * i + j] = si;
        }
      }
    }
  }
  soa_load(x, pmu, V, site);
  mat_mul<false, false, false>(g, x, a);
}

// ---------------------------------------------------------------------------
// Plaquette traces at one site, the reference's six planes in ITS order
// (u = 1..3, v < u):  tr[ U_u(n) U_v(n+u) (U_v(n) U_u(n+v))^+ ]
// (lattice/su3/pytorch/lattice.py:157-199).  tr_re/tr_im[p], p = 0..5.
// ---------------------------------------------------------------------------
template <typename T, typename C>
L2B_HD void site_plaquette_traces(T tr_re[6], T tr_im[6], const C* U, const Lat& l, int b, int site) {
  int c[4];
  site_coords(l, site, c);
  const int V = l.V;
  int p = 0;
  L2B_UNROLL
  for (int u = 1; u < 4; ++u) {
    L2B_UNROLL
    for (int v = 0; v < 3; ++v) {
      if (v < u) {
        const C* pu = soa_plane(U, l, b, u);
        const C* pv = soa_plane(U, l, b, v);
        Mat3<T> a, bb, yuv, yvu;
        soa_load(a, pu, V, site);
        soa_load(bb, pv, V, site_fwd(l, site, c, u));
        mat_mul<false, false, false>(yuv, a, bb);
        soa_load(a, pv, V, site);
        soa_load(bb, pu, V, site_fwd(l, site, c, v));
        mat_mul<false, false, false>(yvu, a, bb);
        trace_mul_adj(yuv, yvu, tr_re[p], tr_im[p]);
        ++p;
      }
    }
  }
}
// ---------------------------------------------------------------------------
// Adjoint of the per-site Wilson loops w_p(n) = tr[U_u(n) U_v(n+u) U_u(n+v)^+ U_v(n)^+]
// (p = plane (u, v), u > v, reference order) for a cotangent gw[p][b][n] (torch
// convention dL/dRe w + i dL/dIm w).  Link (mu, n) sits in six plaquettes, two per other
// direction nu, and with S_f / S_b the forward / backward staples of link_times_staples
//   G_mu(n) = sum_nu [ c_f S_f^+ + c_b S_b^+ ],
//   mu > nu: c_f = gw_p(n),       c_b = conj(gw_p(n - nu))
//   mu < nu: c_f = conj(gw_p(n)), c_b = gw_p(n - nu),        p = plane(max, min).
// (a real, site-independent gw gives back coef * A^+, the action's adjoint.)
// ---------------------------------------------------------------------------
template <typename T, typename C>
L2B_HD void wloops_adjoint_link(Mat3<T>& g, const C* U, const C* gw, const Lat& l, int nb, int b, int mu, int site) {
  const int V = l.V;
  int r = site;
  const int c3 = r % l.L[3]; r /= l.L[3];
  const int c2 = r % l.L[2]; r /= l.L[2];
  const int c1 = r % l.L[1]; r /= l.L[1];
  const int c0 = r;
  const int f0 = (c0 == l.L[0] - 1) ? -(l.L[0] - 1) * l.stride[0] : l.stride[0];
  const int f1 = (c1 == l.L[1] - 1) ? -(l.L[1] - 1) * l.stride[1] : l.stride[1];
  const int f2 = (c2 == l.L[2] - 1) ? -(l.L[2] - 1) * l.stride[2] : l.stride[2];
  const int f3 = (c3 == l.L[3] - 1) ? -(l.L[3] - 1) : 1;
  const int b0 = (c0 == 0) ? (l.L[0] - 1) * l.stride[0] : -l.stride[0];
  const int b1 = (c1 == 0) ? (l.L[1] - 1) * l.stride[1] : -l.stride[1];
  const int b2 = (c2 == 0) ? (l.L[2] - 1) * l.stride[2] : -l.stride[2];
  const int b3 = (c3 == 0) ? (l.L[3] - 1) : -1;
  const size_t plane_sz = (size_t)9 * V;
  const C* chain = U + (size_t)b * 4 * plane_sz;
  const C* pmu = chain + (size_t)mu * plane_sz;
  const int n_pmu = site + sel4(f0, f1, f2, f3, mu);
  Mat3<T> a, x, y, m, st;
  mat_zero(a);                       // a = sum conj(c) S, so that G = a^+
  L2B_UNROLL
  for (int k = 1; k < 4; ++k) {
    const int nu = (mu + k) & 3;
    const C* pnu = chain + (size_t)nu * plane_sz;
    const int fnu = sel4(f0, f1, f2, f3, nu);
    const int bnu = sel4(b0, b1, b2, b3, nu);
    const int n_pnu = site + fnu;
    const int n_mnu = site + bnu;
    const int n_pmu_mnu = n_pmu + bnu;
    const int hi = mu > nu ? mu : nu, lo = mu > nu ? nu : mu;
    const int p = hi * (hi - 1) / 2 + lo;
    const C* gwp = gw + ((size_t)p * nb + b) * V;
    const C wf = L2B_LDG(gwp + site), wb = L2B_LDG(gwp + n_mnu);
    // conj(c_f), conj(c_b)
    const T cfr = wf.x, cfi = (mu > nu) ? -wf.y : wf.y;
    const T cbr = wb.x, cbi = (mu > nu) ? wb.y : -wb.y;
    soa_load(x, pnu, V, n_pmu);
    soa_load(y, pmu, V, n_pnu);
    mat_mul<false, true, false>(m, x, y);
    soa_load(x, pnu, V, site);
    mat_mul<false, true, false>(st, m, x);          // S_f
    L2B_UNROLL
    for (int e = 0; e < 9; ++e) {
      a.re[e] += cfr * st.re[e] - cfi * st.im[e];
      a.im[e] += cfr * st.im[e] + cfi * st.re[e];
    }
    soa_load(x, pnu, V, n_pmu_mnu);
    soa_load(y, pmu, V, n_mnu);
    mat_mul<true, true, false>(m, x, y);
    soa_load(x, pnu, V, n_mnu);
    mat_mul<false, false, false>(st, m, x);         // S_b
    L2B_UNROLL
    for (int e = 0; e < 9; ++e) {
      a.re[e] += cbr * st.re[e] - cbi * st.im[e];
      a.im[e] += cbr * st.im[e] + cbi * st.re[e];
    }
  }
  L2B_UNROLL
  for (int i = 0; i < 3; ++i) {
    L2B_UNROLL
    for (int j = 0; j < 3; ++j) { g.re[3 * i + j] = a.re[3 * j + i]; g.im[3 * i + j] = -a.im[3 * j + i]; }
  }
}

// ---------------------------------------------------------------------------
// Rectangle (c1 != 0) staples.  The reference's improved action adds the traces of the 2x1 and 1x2
// rectangles of every plane, S = -(beta/3) [ (1 - 8 c1) sum Re tr P + c1 sum Re tr R ]
// (lattice/su3/pytorch/lattice.py:96-112,180-196,252-269), and gets the force from autograd.  Here,
// as for the plaquette: force = (beta/3) TAH( U_mu(n) [ (1 - 8 c1) A_mu(n) + c1 R_mu(n) ] ) with
// R_mu(n) the sum of the 18 rectangle staples of link (mu, n): for each nu != mu and each side
// s = +-nu, the paths of five links from n+mu back to n
//   A  (1x2, mu short):                +s +s -mu -s -s
//   B  (2x1, link first on its side):  +mu +s -mu -mu -s
//   C  (2x1, link second on its side): +s -mu -mu -s +mu
// (a forward hop in d from p multiplies by U_d(p), a backward hop by U_d(p - d)^+).
// sum_{mu,n} Re tr(U_mu(n) R_mu(n)) = 6 * sum_rectangles Re tr R (six links per rectangle).
// Extents 1 and 2 wrap onto themselves; the displacement table handles it.
// ---------------------------------------------------------------------------
L2B_HD int pmod(int a, int n) {
  a %= n;
  return a < 0 ? a + n : a;
}
L2B_HD int sel5(int a0, int a1, int a2, int a3, int a4, int i) {
  return (i == 0) ? a0 : (i == 1) ? a1 : (i == 2) ? a2 : (i == 3) ? a3 : a4;
}

template <typename T, typename C>
L2B_HD void rect_staples(Mat3<T>& a, const C* U, const Lat& l, int b, int mu, int site) {
  const int V = l.V;
  int r = site;
  const int c3 = r % l.L[3]; r /= l.L[3];
  const int c2 = r % l.L[2]; r /= l.L[2];
  const int c1 = r % l.L[1]; r /= l.L[1];
  const int c0 = r;
  const size_t plane_sz = (size_t)9 * V;
  const C* chain = U + (size_t)b * 4 * plane_sz;
  const C* pmu = chain + (size_t)mu * plane_sz;
  const int cm = sel4(c0, c1, c2, c3, mu), Lm = sel4(l.L[0], l.L[1], l.L[2], l.L[3], mu);
  const int sm = sel4(l.stride[0], l.stride[1], l.stride[2], l.stride[3], mu);
  // site offset of a displacement of -1, 0, 1, 2 along mu (periodic)
  const int m_m1 = (pmod(cm - 1, Lm) - cm) * sm, m_p1 = (pmod(cm + 1, Lm) - cm) * sm,
            m_p2 = (pmod(cm + 2, Lm) - cm) * sm;
  // hop tables: direction (1 = mu, 0 = nu) and sign (the nu signs are multiplied by the side s)
  const int hop_mu[3][5] = {{0, 0, 1, 0, 0}, {1, 0, 1, 1, 0}, {0, 1, 1, 0, 1}};
  const int hop_sg[3][5] = {{1, 1, -1, -1, -1}, {1, 1, -1, -1, -1}, {1, -1, -1, -1, 1}};
  mat_zero(a);
  L2B_UNROLL
  for (int k = 1; k < 4; ++k) {
    const int nu = (mu + k) & 3;
    const C* pnu = chain + (size_t)nu * plane_sz;
    const int cn = sel4(c0, c1, c2, c3, nu), Ln = sel4(l.L[0], l.L[1], l.L[2], l.L[3], nu);
    const int sn = sel4(l.stride[0], l.stride[1], l.stride[2], l.stride[3], nu);
    const int n_m2 = (pmod(cn - 2, Ln) - cn) * sn, n_m1 = (pmod(cn - 1, Ln) - cn) * sn,
              n_p1 = (pmod(cn + 1, Ln) - cn) * sn, n_p2 = (pmod(cn + 2, Ln) - cn) * sn;
    L2B_UNROLL
    for (int side = 0; side < 2; ++side) {
      const int s = side == 0 ? 1 : -1;
      L2B_UNROLL
      for (int shape = 0; shape < 3; ++shape) {
        int i = 1, j = 0;          // displacement from n along mu / nu: the path starts at n + mu
        Mat3<T> t, x, w;
        L2B_UNROLL
        for (int h = 0; h < 5; ++h) {
          const bool along_mu = hop_mu[shape][h] != 0;
          const int sg = along_mu ? hop_sg[shape][h] : hop_sg[shape][h] * s;
          if (sg < 0) { if (along_mu) --i; else --j; }      // backward hop: the link starts one site back
          const int at = site + sel5(0, m_m1, 0, m_p1, m_p2, i + 2) + sel5(n_m2, n_m1, 0, n_p1, n_p2, j + 2);
          soa_load(x, along_mu ? pmu : pnu, V, at);
          if (sg > 0) { if (along_mu) ++i; else ++j; }
          if (h == 0) {
            if (sg < 0) {
              L2B_UNROLL
              for (int e = 0; e < 9; ++e) { t.re[e] = x.re[3 * (e % 3) + e / 3]; t.im[e] = -x.im[3 * (e % 3) + e / 3]; }
            } else {
              t = x;
            }
          } else if (h < 4) {
            if (sg < 0) mat_mul<false, true, false>(w, t, x); else mat_mul<false, false, false>(w, t, x);
            t = w;
          } else {
            if (sg < 0) mat_mul<false, true, true>(a, t, x); else mat_mul<false, false, true>(a, t, x);
          }
        }
      }
    }
  }
}

// G = U_mu(n) [ (1 - 8 c1) A_mu(n) + c1 R_mu(n) ];  retr_p = Re tr(U A) (4x the plaquette sum when
// summed over links), retr_r = Re tr(U R) (6x the rectangle sum).
template <typename T, typename C>
L2B_HD void link_times_improved_staples(Mat3<T>& g, T& retr_p, T& retr_r, const C* U, const Lat& l, int b, int mu,
                                        int site, T c1) {
  Mat3<T> ap, ar, x, gp, gr;
  link_times_staples<T, C, 0, false>(ap, U, l, b, mu, site);
  rect_staples<T, C>(ar, U, l, b, mu, site);
  soa_load(x, soa_plane(U, l, b, mu), l.V, site);
  mat_mul<false, false, false>(gp, x, ap);
  mat_mul<false, false, false>(gr, x, ar);
  retr_p = re_trace(gp);
  retr_r = re_trace(gr);
  const T cp = T(1) - T(8) * c1;
  L2B_UNROLL
  for (int e = 0; e < 9; ++e) {
    g.re[e] = cp * gp.re[e] + c1 * gr.re[e];
    g.im[e] = cp * gp.im[e] + c1 * gr.im[e];
  }
}

// Adjoint of the improved action / force at one link (the c1 != 0 counterpart of k_action_grad):
//   gf == nullptr:  g = c * Aimp^+,                      Aimp = (1 - 8 c1) A + c1 R
//   gf != nullptr:  g = TAH(gf)^+ (c * Aimp^+)           (force adjoint at fixed dsdx)
template <typename T, typename C>
L2B_HD void improved_action_adjoint_link(Mat3<T>& g, const C* U, const Mat3<T>* gf, const Lat& l, int b, int mu,
                                         int site, T c, T c1) {
  Mat3<T> ap, ar, ah;
  link_times_staples<T, C, 0, false>(ap, U, l, b, mu, site);
  rect_staples<T, C>(ar, U, l, b, mu, site);
  const T cp = c * (T(1) - T(8) * c1), cr = c * c1;
  L2B_UNROLL
  for (int i = 0; i < 3; ++i) {
    L2B_UNROLL
    for (int j = 0; j < 3; ++j) {
      ah.re[3 * i + j] = cp * ap.re[3 * j + i] + cr * ar.re[3 * j + i];
      ah.im[3 * i + j] = -(cp * ap.im[3 * j + i] + cr * ar.im[3 * j + i]);
    }
  }
  if (gf != nullptr) {
    Mat3<T> th;
    project_tah(th, *gf);
    mat_mul<true, false, false>(g, th, ah);
  } else {
    g = ah;
  }
}

}  // namespace l2b
