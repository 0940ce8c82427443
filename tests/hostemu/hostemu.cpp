// tests/hostemu/hostemu.cpp -- TEST INFRASTRUCTURE ONLY (never loaded by the package).
//
// Compiles the __host__ __device__ arithmetic bodies of the CUDA kernels
// (l2hmc_b200/csrc/l2b_su3_math.cuh, l2b_su3_site.cuh) for the CPU with g++ and
// drives them with plain loops that mirror the kernels' thread mapping, so the
// index math and the per-link arithmetic can be diffed against the numpy oracle
// in the `-m "not gpu"` tier.  It shares no code with the oracle.
#include <cstddef>
#include <cstdint>
#include <vector>

#include "../../l2hmc_b200/csrc/l2b_su3_site.cuh"

using namespace l2b;
using T = double;
using C = cplx_host<double>;

static void aos_get(Mat3<T>& m, const C* aos, size_t i) {
  for (int e = 0; e < 9; ++e) { m.re[e] = aos[i * 9 + e].x; m.im[e] = aos[i * 9 + e].y; }
}
static void aos_put(C* aos, size_t i, const Mat3<T>& m) {
  for (int e = 0; e < 9; ++e) { aos[i * 9 + e].x = m.re[e]; aos[i * 9 + e].y = m.im[e]; }
}
static void to_soa(std::vector<C>& soa, const C* aos, int nb, const Lat& l) {
  soa.resize((size_t)nb * 4 * 9 * l.V);
  for (int p = 0; p < nb * 4; ++p)
    for (int s = 0; s < l.V; ++s)
      for (int e = 0; e < 9; ++e) soa[((size_t)p * 9 + e) * l.V + s] = aos[((size_t)p * l.V + s) * 9 + e];
}
static void to_aos(C* aos, const std::vector<C>& soa, int nb, const Lat& l) {
  for (int p = 0; p < nb * 4; ++p)
    for (int s = 0; s < l.V; ++s)
      for (int e = 0; e < 9; ++e) aos[((size_t)p * l.V + s) * 9 + e] = soa[((size_t)p * 9 + e) * l.V + s];
}

extern "C" {

// mode: 0 exp(scale*x), 1 TAH, 2 projectSU, 3 identity;  vec8 optional
void emu_unary(int mode, const C* x, double scale, C* out, double* vec8, size_t n) {
  for (size_t i = 0; i < n; ++i) {
    Mat3<T> m, r;
    aos_get(m, x, i);
    if (mode == 0) { for (int e = 0; e < 9; ++e) { m.re[e] *= scale; m.im[e] *= scale; } mat_exp(r, m); }
    else if (mode == 4) { for (int e = 0; e < 9; ++e) { m.re[e] *= scale; m.im[e] *= scale; } mat_exp_alg(r, m); }
    else if (mode == 1) project_tah(r, m);
    else if (mode == 2) project_su(r, m);
    else r = m;
    if (out) aos_put(out, i, r);
    if (vec8) su3_to_vec(vec8 + i * 8, r);
  }
}
// adjoint of su3_to_vec(projectSU(x)) (gvec8 given) or of projectSU(x) (gmat given)
void emu_project_bwd(const C* x, const C* gmat, const double* gvec8, C* gx, size_t n) {
  for (size_t i = 0; i < n; ++i) {
    Mat3<T> m, g, r;
    aos_get(m, x, i);
    if (gvec8) su3_to_vec_adjoint(g, gvec8 + i * 8); else mat_zero(g);
    if (gmat) { Mat3<T> a; aos_get(a, gmat, i); for (int e = 0; e < 9; ++e) { g.re[e] += a.re[e]; g.im[e] += a.im[e]; } }
    project_su_adjoint(r, m, g);
    aos_put(gx, i, r);
  }
}
// adjoint of the per-site Wilson loops (k_wloops_bwd): x, gx AoS; gw [6][nb][V]
void emu_wloops_bwd(const C* x, const C* gw, C* gx, int nb, const int* dims) {
  const Lat l = make_lat(dims[0], dims[1], dims[2], dims[3]);
  std::vector<C> soa, out((size_t)nb * 4 * 9 * l.V);
  to_soa(soa, x, nb, l);
  for (int b = 0; b < nb; ++b)
    for (int mu = 0; mu < 4; ++mu)
      for (int s = 0; s < l.V; ++s) {
        Mat3<T> g;
        wloops_adjoint_link<T, C>(g, soa.data(), gw, l, nb, b, mu, s);
        soa_store(soa_plane(out.data(), l, b, mu), l.V, s, g);
      }
  to_aos(gx, out, nb, l);
}
void emu_from_vec(const double* vec8, C* out, size_t n) {
  for (size_t i = 0; i < n; ++i) { Mat3<T> m; vec_to_su3(m, vec8 + i * 8); aos_put(out, i, m); }
}
void emu_tah_from_normals(const double* n8, C* out, size_t n) {
  for (size_t i = 0; i < n; ++i) { Mat3<T> m; tah_from_normals(m, n8 + i * 8); aos_put(out, i, m); }
}
void emu_check(const C* x, double* d, size_t n) {
  for (size_t i = 0; i < n; ++i) { Mat3<T> m; aos_get(m, x, i); d[i] = check_su_dev(m); }
}
// force = (beta/3) TAH(U A); retr[b] = sum_links Re tr(U A)
void emu_force(const C* x, double beta, C* force, double* retr, int nb, const int* dims) {
  const Lat l = make_lat(dims[0], dims[1], dims[2], dims[3]);
  std::vector<C> U;
  to_soa(U, x, nb, l);
  for (int b = 0; b < nb; ++b) {
    double acc = 0.0;
    for (int mu = 0; mu < 4; ++mu)
      for (int s = 0; s < l.V; ++s) {
        Mat3<T> g, f;
        link_times_staples<T, C>(g, U.data(), l, b, mu, s);
        acc += re_trace(g);
        project_tah(f, g);
        for (int e = 0; e < 9; ++e) { f.re[e] *= beta / 3.0; f.im[e] *= beta / 3.0; }
        aos_put(force, ((size_t)b * 4 + mu) * l.V + s, f);
      }
    if (retr) retr[b] = acc;
  }
}
// improved (c1 != 0) action: force = (beta/3) TAH(U [(1 - 8 c1) A + c1 R]); sums[b, 2] = (sum Re tr P, sum Re tr R)
void emu_force_c1(const C* x, double beta, double c1, C* force, double* sums, int nb, const int* dims) {
  const Lat l = make_lat(dims[0], dims[1], dims[2], dims[3]);
  std::vector<C> U;
  to_soa(U, x, nb, l);
  for (int b = 0; b < nb; ++b) {
    double sp = 0.0, sr = 0.0;
    for (int mu = 0; mu < 4; ++mu)
      for (int s = 0; s < l.V; ++s) {
        Mat3<T> g, f;
        T rp, rr;
        link_times_improved_staples<T, C>(g, rp, rr, U.data(), l, b, mu, s, c1);
        sp += rp;
        sr += rr;
        project_tah(f, g);
        for (int e = 0; e < 9; ++e) { f.re[e] *= beta / 3.0; f.im[e] *= beta / 3.0; }
        aos_put(force, ((size_t)b * 4 + mu) * l.V + s, f);
      }
    sums[2 * b] = sp / 4.0;
    sums[2 * b + 1] = sr / 6.0;
  }
}
// improved-action adjoints: gx = coef[b] Aimp^+ (gf == NULL) or TAH(gf)^+ (scale Aimp^+)
void emu_action_grad_c1(const C* x, const double* coef, double scale, double c1, const C* gf, C* gx, int nb,
                        const int* dims) {
  const Lat l = make_lat(dims[0], dims[1], dims[2], dims[3]);
  std::vector<C> U;
  to_soa(U, x, nb, l);
  for (int b = 0; b < nb; ++b)
    for (int mu = 0; mu < 4; ++mu)
      for (int s = 0; s < l.V; ++s) {
        const size_t li = ((size_t)b * 4 + mu) * l.V + s;
        Mat3<T> g, f;
        if (gf) aos_get(f, gf, li);
        improved_action_adjoint_link<T, C>(g, U.data(), gf ? &f : nullptr, l, b, mu, s, gf ? scale : coef[b], c1);
        aos_put(gx, li, g);
      }
}
// the force through link_times_staples_hook (what the default kick kernels k_force_ep call): must be
// bit-identical to link_times_staples; `hooks[b]` counts the hook invocations (one per link)
void emu_force_hook(const C* x, double beta, C* force, long long* hooks, int hook_at, int nb, const int* dims) {
  const Lat l = make_lat(dims[0], dims[1], dims[2], dims[3]);
  std::vector<C> U;
  to_soa(U, x, nb, l);
  for (int b = 0; b < nb; ++b) {
    long long n = 0;
    for (int mu = 0; mu < 4; ++mu)
      for (int s = 0; s < l.V; ++s) {
        Mat3<T> g, f;
        auto hook = [&]() { ++n; };
        if (hook_at == 0) link_times_staples_hook<T, C, 0>(g, U.data(), l, b, mu, s, hook);
        else if (hook_at == 2) link_times_staples_hook<T, C, 2>(g, U.data(), l, b, mu, s, hook);
        else if (hook_at == 3) link_times_staples_hook<T, C, 3>(g, U.data(), l, b, mu, s, hook);
        else if (hook_at == 13) link_times_staples_lowreg<T, C, 3>(g, U.data(), l, b, mu, s, hook);
        else link_times_staples_lowreg<T, C, 2>(g, U.data(), l, b, mu, s, hook);   // 12
        project_tah(f, g);
        for (int e = 0; e < 9; ++e) { f.re[e] *= beta / 3.0; f.im[e] *= beta / 3.0; }
        aos_put(force, ((size_t)b * 4 + mu) * l.V + s, f);
      }
    hooks[b] = n;
  }
}
// wloops[6, nb, V] complex
void emu_wloops(const C* x, C* wl, int nb, const int* dims) {
  const Lat l = make_lat(dims[0], dims[1], dims[2], dims[3]);
  std::vector<C> U;
  to_soa(U, x, nb, l);
  for (int b = 0; b < nb; ++b)
    for (int s = 0; s < l.V; ++s) {
      T tr[6], ti[6];
      site_plaquette_traces<T, C>(tr, ti, U.data(), l, b, s);
      for (int p = 0; p < 6; ++p) { wl[((size_t)p * nb + b) * l.V + s].x = tr[p]; wl[((size_t)p * nb + b) * l.V + s].y = ti[p]; }
    }
}
// the kernel sequence of l2b_su3_hmc_trajectory; energies[nb,4] = (KE0, S0, KE1, S1)
void emu_hmc(const C* x, const C* v, double beta, double eps, int nlf, C* xo, C* vo, double* en, int nb,
             const int* dims) {
  const Lat l = make_lat(dims[0], dims[1], dims[2], dims[3]);
  std::vector<C> U, P;
  to_soa(U, x, nb, l);
  to_soa(P, v, nb, l);
  const double b3 = beta / 3.0, shift = 0.0;
  auto norm2_chain = [&](int b) {
    double a = 0.0;   // per link (|P|^2 - 8), like the kernels
    for (int mu = 0; mu < 4; ++mu)
      for (int s = 0; s < l.V; ++s) {
        Mat3<T> m;
        soa_load(m, soa_plane(P.data(), l, b, mu), l.V, s);
        a += norm2(m) - 8.0;
      }
    return a;
  };
  auto force_kick = [&](double coef, double* retr) {
    for (int b = 0; b < nb; ++b) {
      double acc = 0.0;
      for (int mu = 0; mu < 4; ++mu)
        for (int s = 0; s < l.V; ++s) {
          Mat3<T> g, f;
          link_times_staples<T, C>(g, U.data(), l, b, mu, s);
          acc += re_trace(g);
          project_tah(f, g);
          C* pp = soa_plane(P.data(), l, b, mu) + s;
          for (int e = 0; e < 9; ++e) { pp[(size_t)e * l.V].x -= coef * f.re[e]; pp[(size_t)e * l.V].y -= coef * f.im[e]; }
        }
      if (retr) retr[b] = acc;
    }
  };
  std::vector<double> retr(nb);
  for (int b = 0; b < nb; ++b) en[b * 4 + 0] = 0.5 * norm2_chain(b) + shift;
  force_kick(0.5 * eps * b3, retr.data());
  for (int b = 0; b < nb; ++b) en[b * 4 + 1] = -b3 * 0.25 * retr[b];
  for (int k = 1; k <= nlf; ++k) {
    for (int p = 0; p < nb * 4; ++p)
      for (int s = 0; s < l.V; ++s) {
        Mat3<T> pm, ex, u, r;
        soa_load(pm, P.data() + (size_t)p * 9 * l.V, l.V, s);
        for (int e = 0; e < 9; ++e) { pm.re[e] *= eps; pm.im[e] *= eps; }
        soa_load(u, U.data() + (size_t)p * 9 * l.V, l.V, s);
        mat_exp_alg(ex, pm);
        mat_mul<false, false, false>(r, ex, u);
        soa_store(U.data() + (size_t)p * 9 * l.V, l.V, s, r);
      }
    force_kick((k == nlf ? 0.5 : 1.0) * eps * b3, retr.data());
  }
  for (int b = 0; b < nb; ++b) { en[b * 4 + 2] = 0.5 * norm2_chain(b) + shift; en[b * 4 + 3] = -b3 * 0.25 * retr[b]; }
  to_aos(xo, U, nb, l);
  to_aos(vo, P, nb, l);
}

}  // extern "C"

// ---- adjoints --------------------------------------------------------------
extern "C" {
// ga = adjoint of exp at a applied to ge (n matrices)
void emu_exp_adjoint(const C* a, const C* ge, C* ga, size_t n) {
  for (size_t i = 0; i < n; ++i) {
    Mat3<T> A, G, R;
    aos_get(A, a, i); aos_get(G, ge, i);
    bool ok;
    mat_exp_adjoint(R, A, G, ok);
    aos_put(ga, i, R);
  }
}
void emu_to_vec_adjoint(const double* gv, C* gx, size_t n) {
  for (size_t i = 0; i < n; ++i) { Mat3<T> g; su3_to_vec_adjoint(g, gv + i * 8); aos_put(gx, i, g); }
}
// staple sums A_mu(n) in the boundary layout
void emu_staples(const C* x, C* out, int nb, const int* dims) {
  const Lat l = make_lat(dims[0], dims[1], dims[2], dims[3]);
  std::vector<C> U;
  to_soa(U, x, nb, l);
  for (int b = 0; b < nb; ++b)
    for (int mu = 0; mu < 4; ++mu)
      for (int s = 0; s < l.V; ++s) {
        Mat3<T> a;
        link_times_staples<T, C, 0, false>(a, U.data(), l, b, mu, s);
        aos_put(out, ((size_t)b * 4 + mu) * l.V + s, a);
      }
}
}
