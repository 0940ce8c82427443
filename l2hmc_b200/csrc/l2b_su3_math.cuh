// l2b_su3_math.cuh -- per-link 3x3 complex arithmetic for the SU(3) hot path.
//
// Everything here is a register-resident, fully unrolled __host__ __device__
// inline so that (a) the CUDA kernels keep whole matrices in registers and the
// complex multiply-adds compile to back-to-back DFMA, and (b) the very same
// arithmetic can be compiled for the host by tests/hostemu to be diffed against
// the numpy oracle without a GPU.
//
// Reference semantics followed (paths relative to /root/reference/src/l2hmc):
//   mul / adjoint / trace      group/su3/pytorch/group.py:55-75
//   projectTAH                 group/su3/pytorch/group.py:92-103
//   exp / update_gauge         group/su3/pytorch/group.py:45-50,88-90 (torch.matrix_exp)
//   projectSU / eigs3x3 / ...  group/su3/pytorch/utils.py:227-346
//   su3_to_vec / vec_to_su3    group/su3/pytorch/utils.py:394-445
//   randTAH3 (layout of the 8 normals)  group/su3/pytorch/utils.py:171-195
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define L2B_HD __host__ __device__ __forceinline__
#define L2B_UNROLL _Pragma("unroll")
#else
#define L2B_HD inline
#define L2B_UNROLL
#endif

namespace l2b {

// 3x3 complex matrix, row-major element e = 3*i + j, split re/im so that every
// scalar lands in its own register.
template <typename T>
struct Mat3 {
  T re[9];
  T im[9];
};

template <typename T>
L2B_HD void mat_zero(Mat3<T>& a) {
  L2B_UNROLL
  for (int e = 0; e < 9; ++e) { a.re[e] = T(0); a.im[e] = T(0); }
}

template <typename T>
L2B_HD void mat_identity(Mat3<T>& a) {
  mat_zero(a);
  a.re[0] = a.re[4] = a.re[8] = T(1);
}

// C (+)= op(A) * op(B),  op = identity or conjugate transpose.
template <bool ADJ_A, bool ADJ_B, bool ACC, typename T>
L2B_HD void mat_mul(Mat3<T>& c, const Mat3<T>& a, const Mat3<T>& b) {
  L2B_UNROLL
  for (int i = 0; i < 3; ++i) {
    L2B_UNROLL
    for (int j = 0; j < 3; ++j) {
      T sr = ACC ? c.re[3 * i + j] : T(0);
      T si = ACC ? c.im[3 * i + j] : T(0);
      L2B_UNROLL
      for (int k = 0; k < 3; ++k) {
        const int ea = ADJ_A ? (3 * k + i) : (3 * i + k);
        const int eb = ADJ_B ? (3 * j + k) : (3 * k + j);
        const T ar = a.re[ea];
        const T ai = ADJ_A ? -a.im[ea] : a.im[ea];
        const T br = b.re[eb];
        const T bi = ADJ_B ? -b.im[eb] : b.im[eb];
        sr = fma(ar, br, sr);
        sr = fma(-ai, bi, sr);
        si = fma(ar, bi, si);
        si = fma(ai, br, si);
      }
      c.re[3 * i + j] = sr;
      c.im[3 * i + j] = si;
    }
  }
}

// tr(A * B^+) = sum_ij A_ij conj(B_ij)
template <typename T>
L2B_HD void trace_mul_adj(const Mat3<T>& a, const Mat3<T>& b, T& tr_re, T& tr_im) {
  T sr = T(0), si = T(0);
  L2B_UNROLL
  for (int e = 0; e < 9; ++e) {
    sr = fma(a.re[e], b.re[e], sr);
    sr = fma(a.im[e], b.im[e], sr);
    si = fma(a.im[e], b.re[e], si);
    si = fma(-a.re[e], b.im[e], si);
  }
  tr_re = sr;
  tr_im = si;
}

template <typename T>
L2B_HD T re_trace(const Mat3<T>& a) { return a.re[0] + a.re[4] + a.re[8]; }

template <typename T>
L2B_HD T norm2(const Mat3<T>& a) {
  T s = T(0);
  L2B_UNROLL
  for (int e = 0; e < 9; ++e) { s = fma(a.re[e], a.re[e], s); s = fma(a.im[e], a.im[e], s); }
  return s;
}

// R = (X - X^+)/2 - tr(.)/3       (group.py:92-103)
template <typename T>
L2B_HD void project_tah(Mat3<T>& r, const Mat3<T>& x) {
  L2B_UNROLL
  for (int i = 0; i < 3; ++i) {
    L2B_UNROLL
    for (int j = 0; j < 3; ++j) {
      r.re[3 * i + j] = T(0.5) * (x.re[3 * i + j] - x.re[3 * j + i]);
      r.im[3 * i + j] = T(0.5) * (x.im[3 * i + j] + x.im[3 * j + i]);
    }
  }
  // the diagonal of (X - X^+)/2 is purely imaginary
  const T d = (r.im[0] + r.im[4] + r.im[8]) / T(3);
  r.im[0] -= d; r.im[4] -= d; r.im[8] -= d;
}

template <typename T>
L2B_HD void cmul_(T ar, T ai, T br, T bi, T& cr, T& ci) {
  cr = ar * br - ai * bi;
  ci = ar * bi + ai * br;
}

template <typename T>
L2B_HD void det3(const Mat3<T>& a, T& dr, T& di) {
  // cofactor expansion along the first row, complex
#define cm cmul_<T>
  T m0r, m0i, m1r, m1i, t0r, t0i, t1r, t1i;
  // c0 = a11 a22 - a12 a21
  cm(a.re[4], a.im[4], a.re[8], a.im[8], t0r, t0i);
  cm(a.re[5], a.im[5], a.re[7], a.im[7], t1r, t1i);
  m0r = t0r - t1r; m0i = t0i - t1i;
  cm(a.re[0], a.im[0], m0r, m0i, m1r, m1i);
  dr = m1r; di = m1i;
  // c1 = a10 a22 - a12 a20
  cm(a.re[3], a.im[3], a.re[8], a.im[8], t0r, t0i);
  cm(a.re[5], a.im[5], a.re[6], a.im[6], t1r, t1i);
  m0r = t0r - t1r; m0i = t0i - t1i;
  cm(a.re[1], a.im[1], m0r, m0i, m1r, m1i);
  dr -= m1r; di -= m1i;
  // c2 = a10 a21 - a11 a20
  cm(a.re[3], a.im[3], a.re[7], a.im[7], t0r, t0i);
  cm(a.re[4], a.im[4], a.re[6], a.im[6], t1r, t1i);
  m0r = t0r - t1r; m0i = t0i - t1i;
  cm(a.re[2], a.im[2], m0r, m0i, m1r, m1i);
  dr += m1r; di += m1i;
#undef cm
}

// ---------------------------------------------------------------------------
// exp(A) for an arbitrary complex 3x3 A (torch.matrix_exp in the reference).
// Cayley-Hamilton: A^3 = t A^2 - c A + d, so A^n = a_n + b_n A + c_n A^2 with
//   (a, b, c)_{n+1} = (d c_n, a_n - c c_n, b_n + t c_n),
// i.e. the Taylor series is summed on three complex scalars instead of on
// matrices (2 matrix products in total instead of ~18).  |c_n| <= C(n,2) rho^(n-2),
// so with ||A||_F <= 1 after scaling by 2^-s, 20 terms truncate below 1e-16;
// the scaling is undone by s squarings.
// ---------------------------------------------------------------------------
template <typename T>
L2B_HD void mat_exp(Mat3<T>& out, const Mat3<T>& ain) {
  Mat3<T> a = ain;
  const T n2 = norm2(a);
  int s = 0;
  if (n2 > T(1)) {
    // s = ceil(log2(||A||_F)) = ceil(log2(n2) / 2)
    int ex;
    const T fr = frexp(n2, &ex);         // n2 = fr * 2^ex, fr in [0.5, 1)
    (void)fr;
    s = (ex + 1) / 2;                    // 2^(2s) >= n2
    if (s > 60) s = 60;
    const T sc = ldexp(T(1), -s);
    L2B_UNROLL
    for (int e = 0; e < 9; ++e) { a.re[e] *= sc; a.im[e] *= sc; }
  }
  Mat3<T> a2;
  mat_mul<false, false, false>(a2, a, a);
  // characteristic polynomial coefficients
  const T tr = a.re[0] + a.re[4] + a.re[8], ti = a.im[0] + a.im[4] + a.im[8];
  const T t2r = a2.re[0] + a2.re[4] + a2.re[8], t2i = a2.im[0] + a2.im[4] + a2.im[8];
  // c = (t^2 - tr A^2) / 2
  const T cr = T(0.5) * (tr * tr - ti * ti - t2r), ci = T(0.5) * (T(2) * tr * ti - t2i);
  T dr, di;
  det3(a, dr, di);
  // running (a_n, b_n, c_n) and their 1/n!-weighted sums
  T anr = T(0), ani = T(0), bnr = T(0), bni = T(0), cnr = T(1), cni = T(0);  // n = 2
  T sar = T(1), sai = T(0), sbr = T(1), sbi = T(0), scr = T(0.5), sci = T(0);
  T inv_fact = T(0.5);
  // terms needed for a truncation error below 1e-16 at this norm (rho = ||A||_F after scaling):
  // the omitted tail is ~ rho^n / (2 (n-2)!).  Leapfrog arguments eps*P have rho ~ 0.1 ... 0.3.
  const T n2s = norm2(a);
  const int nterms = (n2s <= T(0.01)) ? 10 : (n2s <= T(0.09)) ? 13 : (n2s <= T(0.25)) ? 16 : 20;
  L2B_UNROLL
  for (int n = 3; n <= 20; ++n) {
    if (n > nterms) break;
    // (a, b, c)_{n} from (a, b, c)_{n-1}
    const T nar = dr * cnr - di * cni, nai = dr * cni + di * cnr;
    const T nbr = anr - (cr * cnr - ci * cni), nbi = ani - (cr * cni + ci * cnr);
    const T ncr = bnr + (tr * cnr - ti * cni), nci = bni + (tr * cni + ti * cnr);
    anr = nar; ani = nai; bnr = nbr; bni = nbi; cnr = ncr; cni = nci;
    inv_fact = inv_fact / T(n);
    sar = fma(inv_fact, anr, sar); sai = fma(inv_fact, ani, sai);
    sbr = fma(inv_fact, bnr, sbr); sbi = fma(inv_fact, bni, sbi);
    scr = fma(inv_fact, cnr, scr); sci = fma(inv_fact, cni, sci);
  }
  // out = sa + sb A + sc A^2
  L2B_UNROLL
  for (int e = 0; e < 9; ++e) {
    T r = sbr * a.re[e] - sbi * a.im[e];
    T i = sbr * a.im[e] + sbi * a.re[e];
    r = fma(scr, a2.re[e], r); r = fma(-sci, a2.im[e], r);
    i = fma(scr, a2.im[e], i); i = fma(sci, a2.re[e], i);
    out.re[e] = r; out.im[e] = i;
  }
  out.re[0] += sar; out.im[0] += sai;
  out.re[4] += sar; out.im[4] += sai;
  out.re[8] += sar; out.im[8] += sai;
  for (int k = 0; k < s; ++k) {
    Mat3<T> tmp;
    mat_mul<false, false, false>(tmp, out, out);
    out = tmp;
  }
}

// ---------------------------------------------------------------------------
// exp(A) for A in su(3) (traceless anti-Hermitian): the argument of every HMC drift,
// U' = exp(eps P) U  (group.py:45-50 via dynamics.py:900-913), where P is a momentum
// (randTAH3) kicked by projectTAH forces and therefore stays in the algebra.
// Same Cayley-Hamilton series as mat_exp, specialised:  t = tr A = 0,  c = -tr(A^2)/2 =
// ||A||_F^2 / 2 is REAL,  d = det A is purely IMAGINARY (det A = -conj(det A)), so the
// coefficients of A^n = a_n + b_n A + c_n A^2 alternate between purely real and purely
// imaginary:    n even: (a, b, c) = (re, i im, re),    n odd: (i im, re, i im)
// and one Taylor term costs 2 multiply-adds + 3 accumulations instead of 16 + a division
// (1/n! is a table).  A^2 = -A A^+ is Hermitian: 6 of its 9 entries are computed.
// No scaling / squaring: the caller (`mat_exp_alg`) sends ||A||_F > 1 to mat_exp.
// Differences to mat_exp are at rounding level (the neglected trace of a kicked
// momentum is ~1e-17); the trajectory goldens are reproduced to 1e-15.
// ---------------------------------------------------------------------------
template <typename T>
L2B_HD void mat_exp_tah(Mat3<T>& out, const Mat3<T>& a, T n2) {
  // A^2 = -A A^+ : diagonal real, (j, i) = conj(i, j)
  Mat3<T> a2;
  L2B_UNROLL
  for (int i = 0; i < 3; ++i) {
    L2B_UNROLL
    for (int j = i; j < 3; ++j) {
      T sr = T(0), si = T(0);
      L2B_UNROLL
      for (int k = 0; k < 3; ++k) {
        const T ar = a.re[3 * i + k], ai = a.im[3 * i + k], br = a.re[3 * j + k], bi = a.im[3 * j + k];
        sr = fma(-ar, br, sr);
        sr = fma(-ai, bi, sr);
        if (j != i) { si = fma(ar, bi, si); si = fma(-ai, br, si); }
      }
      a2.re[3 * i + j] = sr; a2.im[3 * i + j] = si;
      if (j != i) { a2.re[3 * j + i] = sr; a2.im[3 * j + i] = -si; }
    }
  }
  const T c = T(-0.5) * (a2.re[0] + a2.re[4] + a2.re[8]);     // = ||A||_F^2 / 2
  T dr, dl;
  det3(a, dr, dl);                                            // d = i dl (Re d = 0 in exact arithmetic)
  (void)dr;
  // 1/n!, n = 0 .. 20
  const T inv_fact[21] = {T(1), T(1), T(0.5), T(1.0 / 6), T(1.0 / 24), T(1.0 / 120), T(1.0 / 720), T(1.0 / 5040),
                          T(1.0 / 40320), T(1.0 / 362880), T(1.0 / 3628800), T(1.0 / 39916800), T(1.0 / 479001600),
                          T(1.0 / 6227020800.0), T(1.0 / 87178291200.0), T(1.0 / 1307674368000.0),
                          T(1.0 / 20922789888000.0), T(1.0 / 355687428096000.0), T(1.0 / 6402373705728000.0),
                          T(1.0 / 121645100408832000.0), T(1.0 / 2432902008176640000.0)};
  // running coefficients: pa, pb, pc hold the single non-zero component of a_n, b_n, c_n
  T pa = T(0), pb = T(0), pc = T(1);                          // n = 2: (0, 0, 1), all "real" slots
  T sar = T(1), sai = T(0), sbr = T(1), sbi = T(0), scr = T(0.5), sci = T(0);
  const int nterms = (n2 <= T(0.01)) ? 10 : (n2 <= T(0.09)) ? 13 : (n2 <= T(0.25)) ? 16 : 20;
  L2B_UNROLL
  for (int n = 3; n <= 20; ++n) {
    if (n > nterms) break;
    T na, nb;
    if (n & 1) {            // from even n-1 (a re, b im, c re)  ->  a = i dl c,  b = a - c c (re),  c = i b
      na = dl * pc;
      nb = fma(-c, pc, pa);
    } else {                // from odd n-1 (a im, b re, c im)   ->  a = -dl c (re),  b = i (a - c c),  c = b (re)
      na = -dl * pc;
      nb = fma(-c, pc, pa);
    }
    pc = pb; pa = na; pb = nb;
    const T w = inv_fact[n];
    if (n & 1) { sai = fma(w, pa, sai); sbr = fma(w, pb, sbr); sci = fma(w, pc, sci); }
    else       { sar = fma(w, pa, sar); sbi = fma(w, pb, sbi); scr = fma(w, pc, scr); }
  }
  // out = sa + sb A + sc A^2
  L2B_UNROLL
  for (int e = 0; e < 9; ++e) {
    T r = sbr * a.re[e] - sbi * a.im[e];
    T i = sbr * a.im[e] + sbi * a.re[e];
    r = fma(scr, a2.re[e], r); r = fma(-sci, a2.im[e], r);
    i = fma(scr, a2.im[e], i); i = fma(sci, a2.re[e], i);
    out.re[e] = r; out.im[e] = i;
  }
  out.re[0] += sar; out.im[0] += sai;
  out.re[4] += sar; out.im[4] += sai;
  out.re[8] += sar; out.im[8] += sai;
}

// exp(A) for an argument that is EXPECTED to lie in su(3) (HMC drifts): the specialised series when A is
// anti-Hermitian and traceless to 1e-14 relative and ||A||_F <= 1, the general mat_exp otherwise -- so any
// input gets the reference's torch.matrix_exp semantics, and the expected ones get it ~2.5x cheaper.
template <typename T>
L2B_HD void mat_exp_alg(Mat3<T>& out, const Mat3<T>& a) {
  const T n2 = norm2(a);
  // || A + A^+ ||_F^2 and |tr A|^2 (the diagonal of A + A^+ is 2 Re a_ii; tr of an anti-Hermitian A is i Im tr)
  T dev = T(0);
  L2B_UNROLL
  for (int i = 0; i < 3; ++i) {
    dev = fma(a.re[4 * i], a.re[4 * i], dev);
    L2B_UNROLL
    for (int j = i + 1; j < 3; ++j) {
      const T er = a.re[3 * i + j] + a.re[3 * j + i], ei = a.im[3 * i + j] - a.im[3 * j + i];
      dev = fma(er, er, dev);
      dev = fma(ei, ei, dev);
    }
  }
  const T ti = a.im[0] + a.im[4] + a.im[8];
  dev = fma(ti, ti, dev);
  if (dev <= T(1e-28) * n2 && n2 <= T(1)) mat_exp_tah(out, a, n2);
  else mat_exp(out, a);
}

// ---------------------------------------------------------------------------
// projectSU  (utils.py:227-346), clamps included
// ---------------------------------------------------------------------------
template <typename T>
L2B_HD void eigs3x3(T tr, T p2, T det, T& e0, T& e1, T& e2) {
  const T third = T(1) / T(3);
  const T tr3 = third * tr;
  const T p23 = third * p2;
  const T tr32 = tr3 * tr3;
  const T q = fabs(T(0.5) * (p23 - tr32));
  const T r = T(0.25) * tr3 * (T(5) * tr32 - p2) - T(0.5) * det;
  const T sq = sqrt(q);
  const T sq3 = q * sq;
  T isq3 = T(1) / sq3;
  isq3 = fmin(T(3e38), fmax(T(-3e38), isq3));
  T rsq3 = r * isq3;
  rsq3 = fmin(T(1), fmax(T(-1), rsq3));
  const T lim = T(1) - T(1e-12);
  rsq3 = fmin(lim, fmax(-lim, rsq3));
  const T t = third * acos(rsq3);
  const T st = sin(t), ct = cos(t);
  const T sqc = sq * ct;
  const T sqs = T(1.7320508075688772935) * sq * st;
  const T ll = tr3 + sqc;
  e0 = tr3 - T(2) * sqc;
  e1 = ll + sqs;
  e2 = ll - sqs;
}

template <typename T>
L2B_HD void rsqrt_phm3_coeffs(T tr, T p2, T det, T& c0, T& c1, T& c2) {
  T e0, e1, e2;
  eigs3x3(tr, p2, det, e0, e1, e2);
  const T se0 = sqrt(fabs(e0)), se1 = sqrt(fabs(e1)), se2 = sqrt(fabs(e2));
  const T u = se0 + se1 + se2;
  const T w = se0 * se1 * se2;
  const T d = w * (se0 + se1) * (se0 + se2) * (se1 + se2);
  const T di = T(1) / d;
  c0 = di * (w * u * u + e0 * se0 * (e1 + e2) + e1 * se1 * (e0 + e2) + e2 * se2 * (e0 + e1));
  c1 = -(tr * u + w) * di;
  c2 = u * di;
}

// NEAR: the caller rounds the result to bf16 (vnet inputs of bf16 nets).  A link that is already special unitary
// to 1e-10 -- every link the integrator produces; ||X^+X - 1||_F^2 + |det X - 1|^2 < 1e-20, the checkSU summand of
// utils.py:376-391 -- is returned as it is: projectSU(X) = X up to that deviation, five orders below bf16's
// resolution, for a quarter of the FP64 work (one product and a determinant instead of four products and the
// trigonometric eigenvalue solve).  Anything else (the force, arbitrary user input) takes the full path.
template <typename T, bool NEAR = false>
L2B_HD void project_su(Mat3<T>& y, const Mat3<T>& x) {
  Mat3<T> t, t2, r, m;
  mat_mul<true, false, false>(t, x, x);        // X^+ X
  if (NEAR) {
    T dev = T(0);
    L2B_UNROLL
    for (int e = 0; e < 9; ++e) {
      const T d = t.re[e] - ((e == 0 || e == 4 || e == 8) ? T(1) : T(0));
      dev = fma(d, d, fma(t.im[e], t.im[e], dev));
    }
    if (dev < T(1e-20)) {
      T xr, xi;
      det3(x, xr, xi);
      xr -= T(1);
      if (fma(xr, xr, xi * xi) < T(1e-20)) {
        y = x;
        return;
      }
    }
  }
  mat_mul<false, false, false>(t2, t, t);
  const T tr = re_trace(t);
  const T p2 = re_trace(t2);
  T dr, di;
  det3(t, dr, di);
  T c0, c1, c2;
  rsqrt_phm3_coeffs(tr, p2, dr, c0, c1, c2);
  L2B_UNROLL
  for (int e = 0; e < 9; ++e) {
    r.re[e] = fma(c1, t.re[e], c2 * t2.re[e]);
    r.im[e] = fma(c1, t.im[e], c2 * t2.im[e]);
  }
  r.re[0] += c0; r.re[4] += c0; r.re[8] += c0;
  mat_mul<false, false, false>(m, x, r);       // projectU
  det3(m, dr, di);
  const T p = -atan2(di, dr) / T(3);
  const T cp = cos(p), sp = sin(p);
  L2B_UNROLL
  for (int e = 0; e < 9; ++e) {
    y.re[e] = m.re[e] * cp - m.im[e] * sp;
    y.im[e] = m.re[e] * sp + m.im[e] * cp;
  }
}

// checkSU summand: ||X^+X - 1||_F^2 + |det X - 1|^2   (utils.py:376-391)
template <typename T>
L2B_HD T check_su_dev(const Mat3<T>& x) {
  Mat3<T> t;
  mat_mul<true, false, false>(t, x, x);
  t.re[0] -= T(1); t.re[4] -= T(1); t.re[8] -= T(1);
  T dr, di;
  det3(x, dr, di);
  dr -= T(1);
  return norm2(t) + dr * dr + di * di;
}

// su3_to_vec (utils.py:394-420)
template <typename T>
L2B_HD void su3_to_vec(T v[8], const Mat3<T>& x) {
  v[0] = T(-2) * x.im[1];
  v[1] = T(-2) * x.re[1];
  v[2] = x.im[4] - x.im[0];
  v[3] = T(-2) * x.im[2];
  v[4] = T(-2) * x.re[2];
  v[5] = T(-2) * x.im[5];
  v[6] = T(-2) * x.re[5];
  v[7] = T(0.57735026918962576451) * (T(2) * x.im[8] - x.im[4] - x.im[0]);
}

// vec_to_su3 (utils.py:423-445)
template <typename T>
L2B_HD void vec_to_su3(Mat3<T>& m, const T v[8]) {
  const T c = T(-0.5);
  const T x01r = c * v[1], x01i = c * v[0];
  const T x02r = c * v[4], x02i = c * v[3];
  const T x12r = c * v[6], x12i = c * v[5];
  const T x2i = T(0.57735026918962576451) * v[7];
  const T x0i = c * (x2i + v[2]);
  const T x1i = c * (x2i - v[2]);
  m.re[0] = T(0); m.im[0] = x0i;
  m.re[4] = T(0); m.im[4] = x1i;
  m.re[8] = T(0); m.im[8] = x2i;
  m.re[1] = x01r; m.im[1] = x01i;
  m.re[2] = x02r; m.im[2] = x02i;
  m.re[5] = x12r; m.im[5] = x12i;
  m.re[3] = -x01r; m.im[3] = x01i;   // -conj(x01)
  m.re[6] = -x02r; m.im[6] = x02i;
  m.re[7] = -x12r; m.im[7] = x12i;
}

// randTAH3 with the eight N(0,1) draws explicit, in the reference's draw order
// (r3, r8, r01, r02, r12, i01, i02, i12)   (utils.py:171-195)
template <typename T>
L2B_HD void tah_from_normals(Mat3<T>& m, const T n[8]) {
  const T s2 = T(0.70710678118654752440);
  const T s3 = T(0.57735026918962576451);
  const T r3 = s2 * n[0];
  const T r8 = s2 * s3 * n[1];
  m.re[0] = T(0); m.im[0] = r8 + r3;
  m.re[4] = T(0); m.im[4] = r8 - r3;
  m.re[8] = T(0); m.im[8] = T(-2) * r8;
  m.re[1] = s2 * n[2]; m.im[1] = s2 * n[5];
  m.re[3] = -m.re[1];  m.im[3] = m.im[1];
  m.re[2] = s2 * n[3]; m.im[2] = s2 * n[6];
  m.re[6] = -m.re[2];  m.im[6] = m.im[2];
  m.re[5] = s2 * n[4]; m.im[5] = s2 * n[7];
  m.re[7] = -m.re[5];  m.im[7] = m.im[5];
}

// ---------------------------------------------------------------------------
// adjoints used by the training path (torch convention for a real loss L:
// G_Z = dL/dRe Z + i dL/dIm Z, so for Y = A B:  G_A = G_Y B^+,  G_B = A^+ G_Y)
// ---------------------------------------------------------------------------

// Adjoint of E = exp(A):  G_A = sum_{n>=1} 1/n! sum_{j+k=n-1} B^j G_E B^k,  B = A^+.
// With B^m = alpha_0(m) + alpha_1(m) B + alpha_2(m) B^2 (Cayley-Hamilton, as in mat_exp) the double sum collapses to
//   G_A = sum_{p,q=0..2} w_pq B^p G B^q,   w = sum_n W(n) / n!,   W(n) = sum_{j+k=n-1} alpha(j) alpha(k)^T,
// and W obeys a SCALAR recurrence: alpha(m+1) = M alpha(m) with the companion matrix M of the characteristic
// polynomial (B^3 = t B^2 - c B + d), hence  W(n+1) = alpha(n) e_0^T + W(n) M^T  (W is symmetric: six complex
// numbers per term, ~70 flops, against one 3x3 product + three matrix axpys = 255 flops per term of the matrix
// recurrence D_n = B D_{n-1} + G B^{n-1} this replaces: k_update_gauge_bwd 260 -> 178 us at 8^4 x 32 chains).
// The matrix work is done one COLUMN of G_A at a time (Horner in B from the left):
//   u_q = G (B^q e_c),   s_p = sum_q w_pq u_q,   G_A e_c = s_0 + B (s_1 + B s_2),
// five matrix-vector products per column; only B, G and a few 3-vectors are live (the matrix form held six
// matrices).  The argument is passed as B = eps * pb with pb = P^+ unscaled, so that the caller can form
// d/d eps = Re sum conj(G_A) P from pb and never needs P itself; the powers of eps are folded into w_pq.
// `sink(c, r, i)` receives column c of G_A (r[k], i[k]: row k).
// No scaling/squaring: at most 34 terms (3^34/34! ~ 6e-23 at ||A||_F = 3), fewer at the small norms of a
// leapfrog step (11 at ||A||_F <= 0.1); L2HMC arguments are eps*v with eps < 1, ||v||_F ~ 2.8.  `ok` is cleared
// when the norm is outside that range.
template <typename T, typename Sink>
L2B_HD void mat_exp_adjoint_cols(const Mat3<T>& pb, T eps, const Mat3<T>& ge, bool& ok, Sink&& sink) {
  const T n2a = eps * eps * norm2(pb);
  ok = n2a <= T(9);
  // invariants of B = eps pb:  t = tr B,  tr B^2 = sum_ij B_ij B_ji,  c = (t^2 - tr B^2) / 2,  d = det B
  const T tr = eps * (pb.re[0] + pb.re[4] + pb.re[8]), ti = eps * (pb.im[0] + pb.im[4] + pb.im[8]);
  T t2r = T(0), t2i = T(0);
  L2B_UNROLL
  for (int i = 0; i < 3; ++i) {
    L2B_UNROLL
    for (int j = 0; j < 3; ++j) {
      t2r = fma(pb.re[3 * i + j], pb.re[3 * j + i], t2r); t2r = fma(-pb.im[3 * i + j], pb.im[3 * j + i], t2r);
      t2i = fma(pb.re[3 * i + j], pb.im[3 * j + i], t2i); t2i = fma(pb.im[3 * i + j], pb.re[3 * j + i], t2i);
    }
  }
  const T e2 = eps * eps;
  t2r *= e2; t2i *= e2;
  const T cr = T(0.5) * (tr * tr - ti * ti - t2r), ci = T(0.5) * (T(2) * tr * ti - t2i);
  T dr, di;
  det3(pb, dr, di);
  dr *= e2 * eps; di *= e2 * eps;
  // alpha(n-1) and the upper triangle of W(n), at n = 1: alpha(0) = e_0, W(1) = e_0 e_0^T; s = sum_n W(n) / n!
  T a0r = T(1), a0i = T(0), a1r = T(0), a1i = T(0), a2r = T(0), a2i = T(0);
  T w00r = T(1), w00i = T(0), w01r = T(0), w01i = T(0), w02r = T(0), w02i = T(0);
  T w11r = T(0), w11i = T(0), w12r = T(0), w12i = T(0), w22r = T(0), w22i = T(0);
  T s00r = T(1), s00i = T(0), s01r = T(0), s01i = T(0), s02r = T(0), s02i = T(0);
  T s11r = T(0), s11i = T(0), s12r = T(0), s12i = T(0), s22r = T(0), s22i = T(0);
  T inv_fact = T(1);
  // terms needed at this norm: the n-th term is bounded by rho^(n-1) / (n-1)!  (rho = ||A||_F <= 3)
  const int nterms = (n2a <= T(0.01)) ? 11 : (n2a <= T(0.09)) ? 14 : (n2a <= T(0.25)) ? 17
                     : (n2a <= T(1)) ? 21 : (n2a <= T(4)) ? 28 : 34;
  for (int n = 2; n <= nterms; ++n) {
    // alpha(n-2) -> alpha(n-1) = M alpha(n-2):  (d a2,  a0 - c a2,  a1 + t a2)
    const T n0r = dr * a2r - di * a2i, n0i = dr * a2i + di * a2r;
    const T n1r = a0r - (cr * a2r - ci * a2i), n1i = a0i - (cr * a2i + ci * a2r);
    const T n2r = a1r + (tr * a2r - ti * a2i), n2i = a1i + (tr * a2i + ti * a2r);
    a0r = n0r; a0i = n0i; a1r = n1r; a1i = n1i; a2r = n2r; a2i = n2i;
    // W(n) = alpha(n-1) e_0^T + W(n-1) M^T, rows p <= columns q:
    //   (p,0) = alpha_p + d W(p,2),  (p,1) = W(p,0) - c W(p,2),  (p,2) = W(p,1) + t W(p,2)
    const T v00r = a0r + (dr * w02r - di * w02i), v00i = a0i + (dr * w02i + di * w02r);
    const T v01r = w00r - (cr * w02r - ci * w02i), v01i = w00i - (cr * w02i + ci * w02r);
    const T v02r = w01r + (tr * w02r - ti * w02i), v02i = w01i + (tr * w02i + ti * w02r);
    const T v11r = w01r - (cr * w12r - ci * w12i), v11i = w01i - (cr * w12i + ci * w12r);
    const T v12r = w11r + (tr * w12r - ti * w12i), v12i = w11i + (tr * w12i + ti * w12r);
    const T v22r = w12r + (tr * w22r - ti * w22i), v22i = w12i + (tr * w22i + ti * w22r);
    w00r = v00r; w00i = v00i; w01r = v01r; w01i = v01i; w02r = v02r; w02i = v02i;
    w11r = v11r; w11i = v11i; w12r = v12r; w12i = v12i; w22r = v22r; w22i = v22i;
    inv_fact /= T(n);
    s00r = fma(inv_fact, w00r, s00r); s00i = fma(inv_fact, w00i, s00i);
    s01r = fma(inv_fact, w01r, s01r); s01i = fma(inv_fact, w01i, s01i);
    s02r = fma(inv_fact, w02r, s02r); s02i = fma(inv_fact, w02i, s02i);
    s11r = fma(inv_fact, w11r, s11r); s11i = fma(inv_fact, w11i, s11i);
    s12r = fma(inv_fact, w12r, s12r); s12i = fma(inv_fact, w12i, s12i);
    s22r = fma(inv_fact, w22r, s22r); s22i = fma(inv_fact, w22i, s22i);
  }
  // B^p G B^q = eps^(p+q) pb^p G pb^q
  s01r *= eps; s01i *= eps;
  s02r *= e2; s02i *= e2; s11r *= e2; s11i *= e2;
  s12r *= e2 * eps; s12i *= e2 * eps;
  s22r *= e2 * e2; s22i *= e2 * e2;
  // y = M x for a 3-vector (complex)
  auto mv = [](const Mat3<T>& m, const T xr[3], const T xi[3], T yr[3], T yi[3]) {
    L2B_UNROLL
    for (int i = 0; i < 3; ++i) {
      T r = T(0), im = T(0);
      L2B_UNROLL
      for (int k = 0; k < 3; ++k) {
        r = fma(m.re[3 * i + k], xr[k], r); r = fma(-m.im[3 * i + k], xi[k], r);
        im = fma(m.re[3 * i + k], xi[k], im); im = fma(m.im[3 * i + k], xr[k], im);
      }
      yr[i] = r; yi[i] = im;
    }
  };
  L2B_UNROLL
  for (int c = 0; c < 3; ++c) {
    T u0r[3], u0i[3], u1r[3], u1i[3], u2r[3], u2i[3], br[3], bi[3], vr[3], vi[3], tr_[3], ti_[3];
    L2B_UNROLL
    for (int k = 0; k < 3; ++k) {
      u0r[k] = ge.re[3 * k + c]; u0i[k] = ge.im[3 * k + c];
      br[k] = pb.re[3 * k + c]; bi[k] = pb.im[3 * k + c];
    }
    mv(ge, br, bi, u1r, u1i);             // G pb e_c
    mv(pb, br, bi, vr, vi);               // pb^2 e_c
    mv(ge, vr, vi, u2r, u2i);             // G pb^2 e_c
    auto comb = [&](T x0r, T x0i, T x1r, T x1i, T x2r, T x2i, T or_[3], T oi_[3], bool add) {
      L2B_UNROLL
      for (int k = 0; k < 3; ++k) {
        T r = add ? or_[k] : T(0), im = add ? oi_[k] : T(0);
        r = fma(x0r, u0r[k], r); r = fma(-x0i, u0i[k], r); im = fma(x0r, u0i[k], im); im = fma(x0i, u0r[k], im);
        r = fma(x1r, u1r[k], r); r = fma(-x1i, u1i[k], r); im = fma(x1r, u1i[k], im); im = fma(x1i, u1r[k], im);
        r = fma(x2r, u2r[k], r); r = fma(-x2i, u2i[k], r); im = fma(x2r, u2i[k], im); im = fma(x2i, u2r[k], im);
        or_[k] = r; oi_[k] = im;
      }
    };
    comb(s02r, s02i, s12r, s12i, s22r, s22i, vr, vi, false);      // s_2
    mv(pb, vr, vi, tr_, ti_);                                     // pb s_2
    comb(s01r, s01i, s11r, s11i, s12r, s12i, tr_, ti_, true);     // + s_1
    mv(pb, tr_, ti_, vr, vi);                                     // pb (s_1 + pb s_2)
    comb(s00r, s00i, s01r, s01i, s02r, s02i, vr, vi, true);       // + s_0
    sink(c, vr, vi);
  }
}

// matrix-argument form: G_A for E = exp(A)
template <typename T>
L2B_HD void mat_exp_adjoint(Mat3<T>& ga, const Mat3<T>& a, const Mat3<T>& ge, bool& ok) {
  Mat3<T> b;
  L2B_UNROLL
  for (int i = 0; i < 3; ++i) {
    L2B_UNROLL
    for (int j = 0; j < 3; ++j) { b.re[3 * i + j] = a.re[3 * j + i]; b.im[3 * i + j] = -a.im[3 * j + i]; }
  }
  mat_exp_adjoint_cols(b, T(1), ge, ok, [&](int c, const T r[3], const T i[3]) {
    L2B_UNROLL
    for (int k = 0; k < 3; ++k) { ga.re[3 * k + c] = r[k]; ga.im[3 * k + c] = i[k]; }
  });
}

// adjoint of su3_to_vec (a real-linear map of the matrix entries)
template <typename T>
L2B_HD void su3_to_vec_adjoint(Mat3<T>& g, const T v[8]) {
  const T s3 = T(0.57735026918962576451);
  mat_zero(g);
  g.re[1] = T(-2) * v[1]; g.im[1] = T(-2) * v[0];
  g.re[2] = T(-2) * v[4]; g.im[2] = T(-2) * v[3];
  g.re[5] = T(-2) * v[6]; g.im[5] = T(-2) * v[5];
  g.im[0] = -v[2] - s3 * v[7];
  g.im[4] = v[2] - s3 * v[7];
  g.im[8] = T(2) * s3 * v[7];
}
// inverse of a complex 3x3 matrix by cofactors (used on the Hermitian positive
// definite I1 H^2 + I3 of the polar-factor adjoint below)
template <typename T>
L2B_HD void mat_inv3(Mat3<T>& inv, const Mat3<T>& a) {
  T dr, di;
  det3(a, dr, di);
  const T dn = T(1) / (dr * dr + di * di);
  const T ir = dr * dn, ii = -di * dn;          // 1 / det
  L2B_UNROLL
  for (int i = 0; i < 3; ++i) {
    L2B_UNROLL
    for (int j = 0; j < 3; ++j) {
      // cofactor C_ji (transpose): rows != j, cols != i, cyclic order keeps the sign
      const int r0 = (j + 1) % 3, r1 = (j + 2) % 3, c0 = (i + 1) % 3, c1 = (i + 2) % 3;
      T pr, pi, qr, qi;
      cmul_<T>(a.re[3 * r0 + c0], a.im[3 * r0 + c0], a.re[3 * r1 + c1], a.im[3 * r1 + c1], pr, pi);
      cmul_<T>(a.re[3 * r0 + c1], a.im[3 * r0 + c1], a.re[3 * r1 + c0], a.im[3 * r1 + c0], qr, qi);
      const T mr = pr - qr, mi = pi - qi;
      inv.re[3 * i + j] = mr * ir - mi * ii;
      inv.im[3 * i + j] = mr * ii + mi * ir;
    }
  }
}

// Adjoint of Y = projectSU(X)  (utils.py:227-346 differentiated by the reference's
// autograd; here analytically).  With X = M H the polar decomposition (M unitary,
// H = (X^+X)^{1/2}) and Y = M e^{-i phi/3}, phi = arg det M:
//   G_M = G_Y conj(c) + i kappa M,  c = e^{-i phi/3},  kappa = Im(c tr(G_Y^+ M)) / 3
//   dM  = M Omega,  H Omega + Omega H = M^+ dX - dX^+ M   (Omega anti-Hermitian)
//   =>  G_X = M Z,  H Z + Z H = Cm := M^+ G_M - G_M^+ M.
// The 3x3 Sylvester equation is solved in closed form: for eigenvalues x, y of H,
//   1/(x+y) = [x^2 - xy + y^2 + I1 (y - x) + I2] / p(y),  p(H) = 2 (I1 H^2 + I3),
// (I1, I2, I3 the invariants of H), hence
//   Z = [Cm H^2 - H Cm H + H^2 Cm + I1 (Cm H - H Cm) + I2 Cm] (2 (I1 H^2 + I3))^{-1}.
// Exact wherever the forward's clamps are inactive (H positive definite).
template <typename T>
L2B_HD void project_su_adjoint(Mat3<T>& gx, const Mat3<T>& x, const Mat3<T>& gy) {
  Mat3<T> t, t2, r, m;
  mat_mul<true, false, false>(t, x, x);
  mat_mul<false, false, false>(t2, t, t);
  T dr, di;
  det3(t, dr, di);
  T c0, c1, c2;
  rsqrt_phm3_coeffs(re_trace(t), re_trace(t2), dr, c0, c1, c2);
  L2B_UNROLL
  for (int e = 0; e < 9; ++e) {
    r.re[e] = fma(c1, t.re[e], c2 * t2.re[e]);
    r.im[e] = fma(c1, t.im[e], c2 * t2.im[e]);
  }
  r.re[0] += c0; r.re[4] += c0; r.re[8] += c0;
  mat_mul<false, false, false>(m, x, r);
  det3(m, dr, di);
  const T p = -atan2(di, dr) / T(3);
  const T cp = cos(p), sp = sin(p);             // c = cp + i sp
  // tr(G_Y^+ M) = sum conj(gy) m  = conj(tr(M G_Y^+))
  T ur, ui;
  trace_mul_adj(m, gy, ur, ui);
  const T kappa = (cp * ui + sp * ur) / T(3);
  Mat3<T> gm;
  L2B_UNROLL
  for (int e = 0; e < 9; ++e) {
    gm.re[e] = gy.re[e] * cp + gy.im[e] * sp - kappa * m.im[e];
    gm.im[e] = gy.im[e] * cp - gy.re[e] * sp + kappa * m.re[e];
  }
  Mat3<T> g, cm, h, h2, ch, hc, num, tmp;
  mat_mul<true, false, false>(g, m, gm);        // M^+ G_M
  L2B_UNROLL
  for (int i = 0; i < 3; ++i) {
    L2B_UNROLL
    for (int j = 0; j < 3; ++j) {
      cm.re[3 * i + j] = g.re[3 * i + j] - g.re[3 * j + i];
      cm.im[3 * i + j] = g.im[3 * i + j] + g.im[3 * j + i];
    }
  }
  mat_mul<true, false, false>(h, x, m);         // H = X^+ M
  mat_mul<false, false, false>(h2, h, h);
  const T i1 = re_trace(h);
  const T i2 = T(0.5) * (i1 * i1 - re_trace(h2));
  det3(h, dr, di);
  const T i3 = dr;
  mat_mul<false, false, false>(ch, cm, h);
  mat_mul<false, false, false>(hc, h, cm);
  L2B_UNROLL
  for (int e = 0; e < 9; ++e) {
    num.re[e] = fma(i1, ch.re[e] - hc.re[e], i2 * cm.re[e]);
    num.im[e] = fma(i1, ch.im[e] - hc.im[e], i2 * cm.im[e]);
  }
  mat_mul<false, false, true>(num, ch, h);      // + Cm H^2
  mat_mul<false, false, true>(num, h, hc);      // + H^2 Cm
  mat_mul<false, false, false>(tmp, h, ch);     // H Cm H
  L2B_UNROLL
  for (int e = 0; e < 9; ++e) {
    num.re[e] -= tmp.re[e]; num.im[e] -= tmp.im[e];
    tmp.re[e] = T(2) * i1 * h2.re[e]; tmp.im[e] = T(2) * i1 * h2.im[e];
  }
  tmp.re[0] += T(2) * i3; tmp.re[4] += T(2) * i3; tmp.re[8] += T(2) * i3;
  mat_inv3(h2, tmp);                            // h2 <- (2 (I1 H^2 + I3))^{-1}
  mat_mul<false, false, false>(tmp, num, h2);   // Z
  mat_mul<false, false, false>(gx, m, tmp);
}

}  // namespace l2b
