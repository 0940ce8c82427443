// l2b_u1.cu -- sm_100a kernels + C ABI for the 2-D U(1) leapfrog hot path.
//
// x[b, mu, t, x] real angles (mu = 0, 1), plaquette angle
//   w(t,x) = x0(t,x) + x1(t+1,x) - x0(t,x+1) - x1(t,x)
// (reference lattice/u1/pytorch/lattice.py:154-159; dims=1 is T, dims=2 is X).
//
// The HMC trajectory kernel keeps one chain's links, momenta and sin(w) in
// shared memory for the WHOLE trajectory (one thread block per chain), so HBM
// sees each field once on the way in and once on the way out regardless of the
// number of leapfrog steps.
#include <math.h>

#include "l2b_common.cuh"

namespace l2b {
namespace {

constexpr double kPi = 3.14159265358979323846;
constexpr double kTwoPi = 6.28318530717958647692;

template <typename T> struct Num;
template <> struct Num<float> {
  static __device__ __forceinline__ float sin_(float a) { return sinf(a); }
  static __device__ __forceinline__ float cos_(float a) { return cosf(a); }
  static __device__ __forceinline__ float tan_(float a) { return tanf(a); }
  static __device__ __forceinline__ float atan_(float a) { return atanf(a); }
  static __device__ __forceinline__ float exp_(float a) { return expf(a); }
  static __device__ __forceinline__ float log_(float a) { return logf(a); }
  static __device__ __forceinline__ float floor_(float a) { return floorf(a); }
  static __device__ __forceinline__ float fmod_(float a, float b) { return fmodf(a, b); }
};
template <> struct Num<double> {
  static __device__ __forceinline__ double sin_(double a) { return sin(a); }
  static __device__ __forceinline__ double cos_(double a) { return cos(a); }
  static __device__ __forceinline__ double tan_(double a) { return tan(a); }
  static __device__ __forceinline__ double atan_(double a) { return atan(a); }
  static __device__ __forceinline__ double exp_(double a) { return exp(a); }
  static __device__ __forceinline__ double log_(double a) { return log(a); }
  static __device__ __forceinline__ double floor_(double a) { return floor(a); }
  static __device__ __forceinline__ double fmod_(double a, double b) { return fmod(a, b); }
};

// ((x + pi) mod 2 pi) - pi with Python-style modulo (group/u1/pytorch/group.py:130-131)
template <typename T>
__device__ __forceinline__ T wrap_pi(T x) {
  const T pi = (T)kPi, tp = (T)kTwoPi;
  T r = Num<T>::fmod_(x + pi, tp);
  if (r < T(0)) r += tp;
  return r - pi;
}
// x - 2 pi floor((x + pi) / 2 pi)   (lattice/u1/pytorch/lattice.py:45-47)
template <typename T>
__device__ __forceinline__ T project_angle(T x) {
  const T pi = (T)kPi, tp = (T)kTwoPi;
  return x - tp * Num<T>::floor_((x + pi) / tp);
}

template <typename T>
__device__ __forceinline__ T plaq_angle(const T* __restrict__ x0, const T* __restrict__ x1, int t, int xx, int Tt, int X) {
  const int tp = (t + 1 == Tt) ? 0 : t + 1;
  const int xp = (xx + 1 == X) ? 0 : xx + 1;
  return x0[t * X + xx] + x1[tp * X + xx] - x0[t * X + xp] - x1[t * X + xx];
}

// ---------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) k_u1_wloops(const T* __restrict__ x, T* __restrict__ w, int Tt, int X) {
  const int N = Tt * X;
  const T* x0 = x + (size_t)blockIdx.y * 2 * N;
  const T* x1 = x0 + N;
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i < N) w[(size_t)blockIdx.y * N + i] = plaq_angle(x0, x1, i / X, i % X, Tt, X);
}

// the reference's "4x4 Wilson loop" (lattice/u1/pytorch/lattice.py:161-186), term by term and in its summation order
// (five x0 links along x at row t, three x1 links up the far side, three x0 links back along row t + 4, five x1 links
// down the near side -- the roll arguments as upstream has them):
//   w(t, x) = x0(t,x) + x0(t,x+1) + x0(t,x+2) + x0(t,x+3) + x0(t,x+4) + x1(t+1,x+4) + x1(t+2,x+4) + x1(t+3,x+4)
//             - x0(t+4,x+3) - x0(t+4,x+2) - x0(t+4,x+1) - x1(t+4,x) - x1(t+3,x) - x1(t+2,x) - x1(t+1,x) - x1(t,x)
template <typename T>
__global__ void __launch_bounds__(256) k_u1_wloops4x4(const T* __restrict__ x, T* __restrict__ w, int Tt, int X) {
  const int N = Tt * X;
  const T* x0 = x + (size_t)blockIdx.y * 2 * N;
  const T* x1 = x0 + N;
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= N) return;
  const int t = i / X, xx = i % X;
  int tr[5], xc[5];
#pragma unroll
  for (int k = 0; k < 5; ++k) { tr[k] = ((t + k) % Tt) * X; xc[k] = (xx + k) % X; }
  T s = x0[tr[0] + xc[0]];
  s += x0[tr[0] + xc[1]]; s += x0[tr[0] + xc[2]]; s += x0[tr[0] + xc[3]]; s += x0[tr[0] + xc[4]];
  s += x1[tr[1] + xc[4]]; s += x1[tr[2] + xc[4]]; s += x1[tr[3] + xc[4]];
  s -= x0[tr[4] + xc[3]]; s -= x0[tr[4] + xc[2]]; s -= x0[tr[4] + xc[1]];
  s -= x1[tr[4] + xc[0]]; s -= x1[tr[3] + xc[0]]; s -= x1[tr[2] + xc[0]]; s -= x1[tr[1] + xc[0]]; s -= x1[tr[0] + xc[0]];
  w[(size_t)blockIdx.y * N + i] = s;
}

// obs[b] = (action, plaq, sinQ, intQ); one block per chain, fixed-order reduction
template <typename T>
__global__ void __launch_bounds__(256) k_u1_obs(const T* __restrict__ x, T beta, T* __restrict__ obs, int Tt, int X) {
  __shared__ double red[8];
  const int N = Tt * X;
  const T* x0 = x + (size_t)blockIdx.x * 2 * N;
  const T* x1 = x0 + N;
  double sc = 0.0, ss = 0.0, sq = 0.0;
  for (int i = threadIdx.x; i < N; i += 256) {
    const T w = plaq_angle(x0, x1, i / X, i % X, Tt, X);
    sc += (double)(T(1) - Num<T>::cos_(w));
    ss += (double)Num<T>::sin_(w);
    sq += (double)project_angle(w);
  }
  sc = block_sum<256>(sc, red, threadIdx.x);
  ss = block_sum<256>(ss, red, threadIdx.x);
  sq = block_sum<256>(sq, red, threadIdx.x);
  if (threadIdx.x == 0) {
    T* o = obs + (size_t)blockIdx.x * 4;
    o[0] = (T)((double)beta * sc);
    o[1] = (T)(1.0 - sc / N);
    o[2] = (T)(ss / kTwoPi);
    o[3] = (T)(sq / kTwoPi);
  }
}

// F0 = beta (sin w(t,x) - sin w(t,x-1)),  F1 = beta (-sin w(t,x) + sin w(t-1,x))
template <typename T>
__global__ void __launch_bounds__(256) k_u1_force(const T* __restrict__ x, T beta, T* __restrict__ f, int Tt, int X) {
  const int N = Tt * X;
  const T* x0 = x + (size_t)blockIdx.y * 2 * N;
  const T* x1 = x0 + N;
  T* f0 = f + (size_t)blockIdx.y * 2 * N;
  T* f1 = f0 + N;
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= N) return;
  const int t = i / X, xx = i % X;
  const int tm = (t == 0) ? Tt - 1 : t - 1;
  const int xm = (xx == 0) ? X - 1 : xx - 1;
  const T s = Num<T>::sin_(plaq_angle(x0, x1, t, xx, Tt, X));
  const T sxm = Num<T>::sin_(plaq_angle(x0, x1, t, xm, Tt, X));
  const T stm = Num<T>::sin_(plaq_angle(x0, x1, tm, xx, Tt, X));
  f0[i] = beta * (s - sxm);
  f1[i] = beta * (-s + stm);
}

// ---------------------------------------------------------------------------
// whole-trajectory HMC, one block per chain, state resident in shared memory
// ---------------------------------------------------------------------------
template <typename T, int NT>
__global__ void __launch_bounds__(NT) k_u1_hmc(const T* __restrict__ x, const T* __restrict__ v, T beta, T eps, int nlf,
                                               T* __restrict__ xo, T* __restrict__ vo, T* __restrict__ energies,
                                               int Tt, int X) {
  extern __shared__ __align__(16) unsigned char smraw[];
  __shared__ double red[NT / 32];
  const int N = Tt * X;
  T* sx = reinterpret_cast<T*>(smraw);   // [2N] links
  T* sv = sx + 2 * N;                    // [2N] momenta
  T* sw = sv + 2 * N;                    // [N]  sin(w)
  const size_t row = (size_t)blockIdx.x * 2 * N;
  const int tid = threadIdx.x;
  double ke = 0.0;
  for (int i = tid; i < 2 * N; i += NT) {
    sx[i] = x[row + i];
    const T p = v[row + i];
    sv[i] = p;
    ke += (double)p * (double)p;
  }
  __syncthreads();
  ke = block_sum<NT>(ke, red, tid);
  if (tid == 0) energies[(size_t)blockIdx.x * 4 + 0] = (T)(0.5 * ke);

  for (int k = 0; k <= nlf; ++k) {
    // sin(w) at the current links; the action rides along on the first and last pass
    const bool want_s = (k == 0) || (k == nlf);
    double sc = 0.0;
    for (int i = tid; i < N; i += NT) {
      const T w = plaq_angle(sx, sx + N, i / X, i % X, Tt, X);
      sw[i] = Num<T>::sin_(w);
      if (want_s) sc += (double)(T(1) - Num<T>::cos_(w));
    }
    __syncthreads();
    if (want_s) {
      sc = block_sum<NT>(sc, red, tid);
      if (tid == 0) energies[(size_t)blockIdx.x * 4 + (k == 0 ? 1 : 3)] = (T)((double)beta * sc);
    }
    // kick (half at both ends, merged full kicks in between), then drift
    const T c = ((k == 0 || k == nlf) ? T(0.5) : T(1)) * eps;
    for (int i = tid; i < N; i += NT) {
      const int t = i / X, xx = i % X;
      const int tm = (t == 0) ? Tt - 1 : t - 1;
      const int xm = (xx == 0) ? X - 1 : xx - 1;
      const T s = sw[i];
      const T f0 = beta * (s - sw[t * X + xm]);
      const T f1 = beta * (-s + sw[tm * X + xx]);
      const T v0 = sv[i] - c * f0;
      const T v1 = sv[N + i] - c * f1;
      sv[i] = v0;
      sv[N + i] = v1;
      if (k < nlf) {
        sx[i] += eps * v0;
        sx[N + i] += eps * v1;
      }
    }
    __syncthreads();
  }
  ke = 0.0;
  for (int i = tid; i < 2 * N; i += NT) {
    xo[row + i] = sx[i];
    const T p = sv[i];
    vo[row + i] = p;
    ke += (double)p * (double)p;
  }
  ke = block_sum<NT>(ke, red, tid);
  if (tid == 0) energies[(size_t)blockIdx.x * 4 + 2] = (T)(0.5 * ke);
}

// ---------------------------------------------------------------------------
// L2HMC element-wise updates on real fields; one block per chain
// ---------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) k_u1_vupdate(const T* __restrict__ v, const T* __restrict__ f,
                                                    const T* __restrict__ s, const T* __restrict__ t,
                                                    const T* __restrict__ q, T eps_in, const T* __restrict__ eps_dev, int sign, T* __restrict__ out,
                                                    T* __restrict__ logdet, int xdim) {
  const T eps = eps_dev ? eps_in * eps_dev[0] : eps_in;   // device-resident step size (CUDA graphs)
  __shared__ double red[8];
  const size_t row = (size_t)blockIdx.x * xdim;
  double ld = 0.0;
  const T half_eps = T(0.5) * eps;
  for (int i = threadIdx.x; i < xdim; i += 256) {
    const T si = s ? s[row + i] : T(0), ti = t ? t[row + i] : T(0), qi = q ? q[row + i] : T(0);
    const T logjac = (T)sign * eps * si / T(2);
    ld += (double)logjac;
    const T es = Num<T>::exp_(logjac);
    const T eq = Num<T>::exp_(eps * qi);
    const T fn = f[row + i] * eq + ti;
    out[row + i] = (sign > 0) ? (es * v[row + i] - half_eps * fn) : (es * (v[row + i] + half_eps * fn));
  }
  if (logdet != nullptr) {
    ld = block_sum<256>(ld, red, threadIdx.x);
    if (threadIdx.x == 0) logdet[blockIdx.x] = (T)ld;
  }
}

template <typename T>
__global__ void __launch_bounds__(256) k_u1_xupdate(const T* __restrict__ x, const T* __restrict__ v,
                                                    const T* __restrict__ s, const T* __restrict__ t,
                                                    const T* __restrict__ q, const float* __restrict__ mask, T eps_in, const T* __restrict__ eps_dev,
                                                    int sign, int use_ncp, T* __restrict__ out,
                                                    T* __restrict__ logdet, int xdim) {
  const T eps = eps_dev ? eps_in * eps_dev[0] : eps_in;   // device-resident step size (CUDA graphs)
  __shared__ double red[8];
  const size_t row = (size_t)blockIdx.x * xdim;
  double ld = 0.0;
  for (int i = threadIdx.x; i < xdim; i += 256) {
    const T m = (T)mask[i], mb = T(1) - m;
    const T xi = x[row + i], vi = v[row + i];
    const T si = ((T)sign * eps) * (s ? s[row + i] : T(0));
    const T qi = eps * (q ? q[row + i] : T(0));
    const T ti = t ? t[row + i] : T(0);
    const T es = Num<T>::exp_(si), eq = Num<T>::exp_(qi);
    const T tr = eps * (vi * eq + ti);
    T xn, lj;
    if (use_ncp) {
      const T hx = xi / T(2);
      const T x1 = T(2) * Num<T>::atan_(Num<T>::tan_(hx) * es);
      xn = (sign > 0) ? (x1 + tr) : (x1 - es * tr);
      const T ct = Num<T>::cos_(hx), st = es * Num<T>::sin_(hx);
      lj = Num<T>::log_(es / (ct * ct + st * st));
    } else {
      xn = (sign > 0) ? (xi * es + tr) : (es * (xi - tr));
      lj = si;
    }
    ld += (double)(mb * lj);
    out[row + i] = wrap_pi(m * xi + mb * xn);
  }
  if (logdet != nullptr) {
    ld = block_sum<256>(ld, red, threadIdx.x);
    if (threadIdx.x == 0) logdet[blockIdx.x] = (T)ld;
  }
}

// ---------------------------------------------------------------------------
// The three output heads of a U(1) LeapfrogLayer (network.py:536-548) fused with the update that
// consumes them (dynamics.py:1266-1297 v-update / :1398-1467 x-update): s, t, q ([nb, xdim] each,
// three times the size of the field) never reach HBM.  The hidden width of the U(1) nets is small
// (conf/network/default.yaml: 16), far below a tensor-core K step worth filling, so this is a
// CUDA-core kernel: each thread owns one output column, keeps its three weight rows (3 x HP
// registers) and walks over a tile of chains whose hidden vectors z sit in shared memory (broadcast
// reads); x / v / F accesses are coalesced across the columns of a warp.
//   MODE 0: v' = vupdate(v, F; s, t, q)     MODE 1: x' = xupdate(x, v; s, t, q; mask)
// ---------------------------------------------------------------------------
template <typename T> __device__ __forceinline__ T tanh_acc(T a) { return tanh(a); }
template <> __device__ __forceinline__ float tanh_acc<float>(float a) { return tanhf(a); }
template <typename T> __device__ __forceinline__ T tanh_(T a);
template <> __device__ __forceinline__ float tanh_<float>(float a) {   // 1 - 2 / (1 + e^{2a}): abs error ~2e-7
  return 1.0f - __fdividef(2.0f, 1.0f + __expf(2.0f * a));
}
template <> __device__ __forceinline__ double tanh_<double>(double a) { return tanh(a); }

// element-wise math of the fused kernel.  float: SFU intrinsics (absolute errors ~1e-7 on the arguments
// that occur here: half-angles in [-pi/2, pi/2], exponents of O(eps)), well inside the 1e-5 fp32 budget;
// double: the accurate library functions.
template <typename T> struct FNum : Num<T> {
  static __device__ __forceinline__ void sincos_(T a, T& s, T& c) { s = Num<T>::sin_(a); c = Num<T>::cos_(a); }
  static __device__ __forceinline__ T div_(T a, T b) { return a / b; }
};
template <> struct FNum<float> : Num<float> {
  static __device__ __forceinline__ float exp_(float a) { return __expf(a); }
  static __device__ __forceinline__ float log_(float a) { return __logf(a); }
  static __device__ __forceinline__ void sincos_(float a, float& s, float& c) { __sincosf(a, &s, &c); }
  static __device__ __forceinline__ float div_(float a, float b) { return __fdividef(a, b); }
};

constexpr int kHeadsChains = 64;      // most chains per block tile (amortises the per-thread weight-row loads); small
                                      // batches take shorter tiles (`cpb`) so that the grid still fills the GPU:
                                      // 128 chains x 512 columns were 4 blocks of 64 chains (62 us), now 64 of 4

template <typename T, int HP, int MODE>
__global__ void __launch_bounds__(256, 3) k_u1_heads_update(
    const T* __restrict__ z, int H, const T* __restrict__ Ws, const T* __restrict__ Wt, const T* __restrict__ Wq,
    const T* __restrict__ bs, const T* __restrict__ bt, const T* __restrict__ bq, const T* __restrict__ cs,
    const T* __restrict__ cq, T nws, T nwt, T nwq, const T* __restrict__ a, const T* __restrict__ bfield,
    const float* __restrict__ mask, T eps_in, const T* __restrict__ eps_dev, int sign, int use_ncp,
    T* __restrict__ out, double* __restrict__ part, int nb, int xdim, int cpb) {
  __shared__ T zt[kHeadsChains][HP];
  __shared__ double ldw[8][kHeadsChains];
  const T eps = eps_dev ? eps_in * eps_dev[0] : eps_in;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int j = blockIdx.x * 256 + tid;
  const int b0 = blockIdx.y * cpb;
  const int nbt = min(cpb, nb - b0);
  for (int idx = tid; idx < cpb * HP; idx += 256) {
    const int c = idx / HP, k = idx % HP;
    zt[c][k] = (c < nbt && k < H) ? z[(size_t)(b0 + c) * H + k] : T(0);
  }
  const bool ok = j < xdim;
  T ws[HP], wt[HP], wq[HP];
  if (sizeof(T) == 4 && (H & 3) == 0 && ok) {            // rows are 16-byte aligned: 128-bit loads
    const float4* r0 = reinterpret_cast<const float4*>(Ws + (size_t)j * H);
    const float4* r1 = reinterpret_cast<const float4*>(Wt + (size_t)j * H);
    const float4* r2 = reinterpret_cast<const float4*>(Wq + (size_t)j * H);
#pragma unroll
    for (int k4 = 0; k4 < HP / 4; ++k4) {
      const bool kk = 4 * k4 < H;
      const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
      const float4 u0 = kk ? __ldg(r0 + k4) : z4, u1 = kk ? __ldg(r1 + k4) : z4, u2 = kk ? __ldg(r2 + k4) : z4;
      ws[4 * k4] = (T)u0.x; ws[4 * k4 + 1] = (T)u0.y; ws[4 * k4 + 2] = (T)u0.z; ws[4 * k4 + 3] = (T)u0.w;
      wt[4 * k4] = (T)u1.x; wt[4 * k4 + 1] = (T)u1.y; wt[4 * k4 + 2] = (T)u1.z; wt[4 * k4 + 3] = (T)u1.w;
      wq[4 * k4] = (T)u2.x; wq[4 * k4 + 1] = (T)u2.y; wq[4 * k4 + 2] = (T)u2.z; wq[4 * k4 + 3] = (T)u2.w;
    }
  } else {
#pragma unroll
    for (int k = 0; k < HP; ++k) {
      const bool kk = ok && k < H;
      ws[k] = kk ? Ws[(size_t)j * H + k] : T(0);
      wt[k] = kk ? Wt[(size_t)j * H + k] : T(0);
      wq[k] = kk ? Wq[(size_t)j * H + k] : T(0);
    }
  }
  const T b_s = ok ? bs[j] : T(0), b_t = ok ? bt[j] : T(0), b_q = ok ? bq[j] : T(0);
  const T a_s = ok ? nws * Num<T>::exp_(cs[j]) : T(0), a_q = ok ? nwq * Num<T>::exp_(cq[j]) : T(0);
  const T m = (MODE == 1 && ok) ? (T)mask[j] : T(0), mb = T(1) - m;
  const T sg = (T)sign;
  __syncthreads();
  constexpr int G = 4;                                   // chains per group: their field loads are in flight together
  for (int c0 = 0; c0 < nbt; c0 += G) {
    T av[G], bv[G];
#pragma unroll
    for (int g = 0; g < G; ++g) {
      const bool live = ok && (c0 + g < nbt);
      const size_t at = (size_t)(b0 + c0 + g) * xdim + j;
      av[g] = live ? a[at] : T(0);
      bv[g] = live ? bfield[at] : T(0);
    }
#pragma unroll
    for (int g = 0; g < G; ++g) {
      const int c = c0 + g;
      if (c >= nbt) break;                               // block-uniform
      T sp = b_s, tp = b_t, qp = b_q;
#pragma unroll
      for (int k = 0; k < HP; ++k) {
        const T zk = zt[c][k];
        sp = fma(ws[k], zk, sp);
        tp = fma(wt[k], zk, tp);
        qp = fma(wq[k], zk, qp);
      }
      // s enters the log-Jacobian linearly and is summed over the lattice: accurate tanh; q only scales v / F
      const T sv = a_s * tanh_acc<T>(sp), tv = nwt * tp, qv = a_q * tanh_<T>(qp);
      const size_t at = (size_t)(b0 + c) * xdim + j;
      T lj = T(0);
      if (ok) {
        if (MODE == 0) {                                 // k_u1_vupdate
          const T logjac = sg * eps * sv / T(2);
          const T es = FNum<T>::exp_(logjac), eq = FNum<T>::exp_(eps * qv);
          const T fn = bv[g] * eq + tv;
          const T he = T(0.5) * eps;
          out[at] = (sign > 0) ? (es * av[g] - he * fn) : (es * (av[g] + he * fn));
          lj = logjac;
        } else {                                         // k_u1_xupdate
          const T xi = av[g], vi = bv[g];
          const T si = (sg * eps) * sv, qi = eps * qv;
          // everything the log-Jacobian depends on (e^{s}, sin, cos) uses the accurate functions: those
          // terms are summed over the lattice and a biased 1e-7 per element becomes 1e-2 per chain
          // (profiles/check_u1_fused_accuracy.py); e^{q} only scales v and may come from the SFU
          const T es = Num<T>::exp_(si), eq = FNum<T>::exp_(qi);
          const T tr = eps * (vi * eq + tv);
          T xn, l1;
          if (use_ncp) {
            const T hx = xi / T(2);
            const T sh = Num<T>::sin_(hx), ch = Num<T>::cos_(hx);   // tan = sin / cos: one tanf saved
            const T x1 = T(2) * Num<T>::atan_((sh / ch) * es);
            xn = (sign > 0) ? (x1 + tr) : (x1 - es * tr);
            const T st = es * sh;
            l1 = si - Num<T>::log_(ch * ch + st * st);   // = log(es / (..)) without the exp/log round trip; accurate log:
                                                         // these terms are SUMMED over the lattice, a biased 3e-7 shows
          } else {
            xn = (sign > 0) ? (xi * es + tr) : (es * (xi - tr));
            l1 = si;
          }
          lj = mb * l1;
          out[at] = wrap_pi(m * xi + mb * xn);
        }
      }
      double r = (double)lj;                             // fixed shuffle tree, then a fixed order over the 8 warps
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) r += __shfl_down_sync(0xffffffffu, r, o);
      if (lane == 0) ldw[warp][c] = r;
    }
  }
  __syncthreads();
  if (part != nullptr && tid < nbt) {
    double x = 0.0;
#pragma unroll
    for (int w = 0; w < 8; ++w) x += ldw[w][tid];
    part[(size_t)(b0 + tid) * gridDim.x + blockIdx.x] = x;
  }
}

// ---------------------------------------------------------------------------
// Input layer of a U(1) LeapfrogLayer without conv stack (network.py:349-451), both Linears in ONE pass
// over the two fields:   pre[b, u] = sum_j f1(x_bj) W1[u, j] + f2(x_bj) W2[u, j] + v_bj Wv[u, j]
//   MODE 1 (xnet): f1 = cos(m_j x), f2 = sin(m_j x)   (group_to_vec of the masked links, dynamics.py:1169-1178;
//                  W1 = W_x[:, :xdim], W2 = W_x[:, xdim:])
//   MODE 0 (vnet): f1 = x, no f2                      (raw x and force, dynamics.py:1157-1159)
// cuBLAS runs these as two [nb x K] x [K x 16] GEMMs at 0.75 TB/s plus a cat(cos, sin) pass; here x and v are
// read once.  Thread = column j with its (2 or 3) x UP weights in registers, loop over a tile of chains; the
// per-chain sums over the block's 256 columns use a reduce-scatter butterfly (16 + 16 shuffles/adds for UP = 16
// instead of 16 x 5), per-warp partials in shared memory, fixed order -> deterministic.
// part[colblk][b][UP]; k_u1_input_finish adds the column blocks and the two biases.
// ---------------------------------------------------------------------------
template <typename T> struct InTile { static constexpr int CB = sizeof(T) == 8 ? 32 : 64; };   // 32 KB of smem partials

template <typename T, int UP, int MODE>
__global__ void __launch_bounds__(256, 3) k_u1_input_layer(const T* __restrict__ x, const T* __restrict__ v,
                                                           const float* __restrict__ mask,
                                                           const T* __restrict__ Wx, const T* __restrict__ Wv, int U,
                                                           T* __restrict__ part, int nb, int xdim) {
  static_assert(UP == 16, "butterfly below is written for 16 outputs");
  constexpr int CB = InTile<T>::CB;
  __shared__ T wpart[8][CB][UP];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int j = blockIdx.x * 256 + tid;
  const int b0 = blockIdx.y * CB;
  const int nbt = min(CB, nb - b0);
  const bool ok = j < xdim;
  const int ldx = (MODE == 1) ? 2 * xdim : xdim;
  T w1[UP], w2[UP], wv[UP];
#pragma unroll
  for (int u = 0; u < UP; ++u) {
    const bool uu = ok && u < U;
    w1[u] = uu ? Wx[(size_t)u * ldx + j] : T(0);
    w2[u] = (uu && MODE == 1) ? Wx[(size_t)u * ldx + xdim + j] : T(0);
    wv[u] = uu ? Wv[(size_t)u * xdim + j] : T(0);
  }
  const T m = (MODE == 1 && ok) ? (T)mask[j] : T(1);
  constexpr int G = 4;
  for (int c0 = 0; c0 < nbt; c0 += G) {
    T xv[G], vv[G];
#pragma unroll
    for (int g = 0; g < G; ++g) {
      const bool live = ok && (c0 + g < nbt);
      const size_t at = (size_t)(b0 + c0 + g) * xdim + j;
      xv[g] = live ? x[at] : T(0);
      vv[g] = live ? v[at] : T(0);
    }
#pragma unroll
    for (int g = 0; g < G; ++g) {
      const int c = c0 + g;
      if (c >= nbt) break;
      T f1, f2 = T(0);
      if (MODE == 1) FNum<T>::sincos_(m * xv[g], f2, f1);      // f1 = cos, f2 = sin
      else f1 = xv[g];
      if (!ok) { f1 = T(0); f2 = T(0); }                        // cos(0) = 1 must not leak from padded columns
      T acc[UP];
#pragma unroll
      for (int u = 0; u < UP; ++u) {
        T a = f1 * w1[u];
        if (MODE == 1) a = fma(f2, w2[u], a);
        acc[u] = fma(vv[g], wv[u], a);
      }
      // reduce-scatter over the 32 lanes: after the 4 halving steps lane l holds the partial sum of output
      // (l >> 1) over its 16-lane... pairs; one more exchange finishes it
#define L2B_RS_STEP(O, Hh)                                                              \
      {                                                                                   \
        const bool up = (lane & (O)) != 0;                                                \
        _Pragma("unroll") for (int i = 0; i < (Hh); ++i) {                                \
          const T send = up ? acc[i] : acc[i + (Hh)];   /* the half the partner keeps */    \
          const T keep = up ? acc[i + (Hh)] : acc[i];                                       \
          acc[i] = keep + __shfl_xor_sync(0xffffffffu, send, (O));                         \
        }                                                                                 \
      }
      L2B_RS_STEP(16, 8)
      L2B_RS_STEP(8, 4)
      L2B_RS_STEP(4, 2)
      L2B_RS_STEP(2, 1)
#undef L2B_RS_STEP
      acc[0] += __shfl_xor_sync(0xffffffffu, acc[0], 1);
      if ((lane & 1) == 0) {
        // which output this lane ended up with: bit k of u is set iff the lane kept the upper half at step k
        const int u = ((lane & 16) ? 8 : 0) | ((lane & 8) ? 4 : 0) | ((lane & 4) ? 2 : 0) | ((lane & 2) ? 1 : 0);
        wpart[warp][c][u] = acc[0];
      }
    }
  }
  __syncthreads();
  for (int idx = tid; idx < nbt * UP; idx += 256) {
    const int c = idx / UP, u = idx % UP;
    T s_ = T(0);
#pragma unroll
    for (int w = 0; w < 8; ++w) s_ += wpart[w][c][u];
    part[((size_t)blockIdx.x * nb + (b0 + c)) * UP + u] = s_;
  }
}

template <typename T, int UP>
__global__ void __launch_bounds__(256) k_u1_input_finish(const T* __restrict__ part, int ncolblk, const T* __restrict__ bx,
                                                         const T* __restrict__ bv, int U, T* __restrict__ out, int nb) {
  const int idx = blockIdx.x * 256 + threadIdx.x;
  if (idx >= nb * U) return;
  const int b = idx / U, u = idx % U;
  T s_ = bx[u] + bv[u];
  for (int k = 0; k < ncolblk; ++k) s_ += part[((size_t)k * nb + b) * UP + u];
  out[idx] = s_;
}

template <typename T>
__global__ void __launch_bounds__(256) k_u1_sum_rows(const double* __restrict__ part, int n, T* __restrict__ out) {
  __shared__ double red[8];
  const double* row = part + (size_t)blockIdx.x * n;
  double s = 0.0;
  for (int k = threadIdx.x; k < n; k += 256) s += row[k];
  s = block_sum<256>(s, red, threadIdx.x);
  if (threadIdx.x == 0) out[blockIdx.x] = (T)s;
}

template <typename T>
__global__ void __launch_bounds__(256) k_u1_kinetic(const T* __restrict__ v, T* __restrict__ ke, int xdim) {
  __shared__ double red[8];
  const size_t row = (size_t)blockIdx.x * xdim;
  double acc = 0.0;
  for (int i = threadIdx.x; i < xdim; i += 256) { const double p = (double)v[row + i]; acc += p * p; }
  acc = block_sum<256>(acc, red, threadIdx.x);
  if (threadIdx.x == 0) ke[blockIdx.x] = (T)(0.5 * acc);
}

template <typename T>
__global__ void __launch_bounds__(256) k_u1_wrap(const T* __restrict__ x, T* __restrict__ out, size_t n) {
  const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
  if (i < n) out[i] = wrap_pi(x[i]);
}

// ---------------------------------------------------------------------------
// backward kernels of the U(1) L2HMC path (training).  The reference gets these
// from autograd (force with create_graph=True, lattice/u1/pytorch/lattice.py:102-117;
// updates dynamics.py:1266-1297,1398-1467); here they are written out.
// ---------------------------------------------------------------------------
// adjoint of the plaquette-angle map: gx0 = gw(t,x) - gw(t,x-1), gx1 = gw(t-1,x) - gw(t,x)
template <typename T>
__global__ void __launch_bounds__(256) k_u1_wloops_bwd(const T* __restrict__ gw, T* __restrict__ gx, int Tt, int X) {
  const int N = Tt * X;
  const T* g = gw + (size_t)blockIdx.y * N;
  T* g0 = gx + (size_t)blockIdx.y * 2 * N;
  T* g1 = g0 + N;
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= N) return;
  const int t = i / X, xx = i % X;
  const int tm = (t == 0) ? Tt - 1 : t - 1;
  const int xm = (xx == 0) ? X - 1 : xx - 1;
  const T c = g[i];
  g0[i] = c - g[t * X + xm];
  g1[i] = g[tm * X + xx] - c;
}

// d<gF, F(x)>/dx with F = beta (sin w - sin w(x-1), -sin w + sin w(t-1)):
//   gw(a) = beta cos w(a) [gF0(a) - gF0(a+x) - gF1(a) + gF1(a+t)], then the adjoint above.
template <typename T>
__device__ __forceinline__ T u1_hvp_gw(const T* x0, const T* x1, const T* f0, const T* f1, T beta, int t, int xx, int Tt,
                                       int X) {
  const int tp = (t + 1 == Tt) ? 0 : t + 1;
  const int xp = (xx + 1 == X) ? 0 : xx + 1;
  const T w = x0[t * X + xx] + x1[tp * X + xx] - x0[t * X + xp] - x1[t * X + xx];
  const T c = f0[t * X + xx] - f0[t * X + xp] - f1[t * X + xx] + f1[tp * X + xx];
  return beta * Num<T>::cos_(w) * c;
}
template <typename T>
__global__ void __launch_bounds__(256) k_u1_force_bwd(const T* __restrict__ x, T beta, const T* __restrict__ gf,
                                                      T* __restrict__ gx, int Tt, int X) {
  const int N = Tt * X;
  const T* x0 = x + (size_t)blockIdx.y * 2 * N;
  const T* x1 = x0 + N;
  const T* f0 = gf + (size_t)blockIdx.y * 2 * N;
  const T* f1 = f0 + N;
  T* g0 = gx + (size_t)blockIdx.y * 2 * N;
  T* g1 = g0 + N;
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= N) return;
  const int t = i / X, xx = i % X;
  const int tm = (t == 0) ? Tt - 1 : t - 1;
  const int xm = (xx == 0) ? X - 1 : xx - 1;
  const T c = u1_hvp_gw(x0, x1, f0, f1, beta, t, xx, Tt, X);
  g0[i] = c - u1_hvp_gw(x0, x1, f0, f1, beta, t, xm, Tt, X);
  g1[i] = u1_hvp_gw(x0, x1, f0, f1, beta, tm, xx, Tt, X) - c;
}

// adjoint of k_u1_vupdate.  geps[b] = d/d eps of chain b (summed over chains by the caller).
template <typename T>
__global__ void __launch_bounds__(256) k_u1_vupdate_bwd(const T* __restrict__ v, const T* __restrict__ f,
                                                        const T* __restrict__ s, const T* __restrict__ t,
                                                        const T* __restrict__ q, T eps_in, const T* __restrict__ eps_dev, int sign,
                                                        const T* __restrict__ gout, const T* __restrict__ glogdet,
                                                        T* __restrict__ gv, T* __restrict__ gf, T* __restrict__ gs,
                                                        T* __restrict__ gt, T* __restrict__ gq, T* __restrict__ geps,
                                                        int xdim) {
  const T eps = eps_dev ? eps_in * eps_dev[0] : eps_in;   // device-resident step size (CUDA graphs)
  __shared__ double red[8];
  const size_t row = (size_t)blockIdx.x * xdim;
  const T gl = glogdet ? glogdet[blockIdx.x] : T(0);
  const T sg = (T)sign, he = T(0.5) * eps;
  double ge = 0.0;
  for (int i = threadIdx.x; i < xdim; i += 256) {
    const T si = s ? s[row + i] : T(0), ti = t ? t[row + i] : T(0), qi = q ? q[row + i] : T(0);
    const T vi = v[row + i], fi = f[row + i], go = gout[row + i];
    const T lj = sg * eps * si / T(2);
    const T es = Num<T>::exp_(lj), eq = Num<T>::exp_(eps * qi);
    const T fn = fi * eq + ti;
    T g_es, g_fn;
    if (sign > 0) {        // v' = es v - he fn
      g_es = go * vi;
      g_fn = -he * go;
      ge += (double)(-T(0.5) * fn * go);
    } else {               // v' = es (v + he fn)
      g_es = go * (vi + he * fn);
      g_fn = go * es * he;
      ge += (double)(T(0.5) * fn * go * es);
    }
    gv[row + i] = go * es;
    const T g_lj = g_es * es + gl;
    ge += (double)(g_lj * sg * si / T(2));
    const T g_eq = g_fn * fi;
    ge += (double)(g_eq * eq * qi);
    gf[row + i] = g_fn * eq;
    if (gs) gs[row + i] = g_lj * sg * eps / T(2);
    if (gt) gt[row + i] = g_fn;
    if (gq) gq[row + i] = g_eq * eq * eps;
  }
  ge = block_sum<256>(ge, red, threadIdx.x);
  if (threadIdx.x == 0) geps[blockIdx.x] = (T)ge;
}

// adjoint of k_u1_xupdate (the wrap to [-pi, pi) has unit derivative)
template <typename T>
__global__ void __launch_bounds__(256) k_u1_xupdate_bwd(const T* __restrict__ x, const T* __restrict__ v,
                                                        const T* __restrict__ s, const T* __restrict__ t,
                                                        const T* __restrict__ q, const float* __restrict__ mask, T eps_in, const T* __restrict__ eps_dev,
                                                        int sign, int use_ncp, const T* __restrict__ gout,
                                                        const T* __restrict__ glogdet, T* __restrict__ gx,
                                                        T* __restrict__ gv, T* __restrict__ gs, T* __restrict__ gt,
                                                        T* __restrict__ gq, T* __restrict__ geps, int xdim) {
  const T eps = eps_dev ? eps_in * eps_dev[0] : eps_in;   // device-resident step size (CUDA graphs)
  __shared__ double red[8];
  const size_t row = (size_t)blockIdx.x * xdim;
  const T gl = glogdet ? glogdet[blockIdx.x] : T(0);
  const T sg = (T)sign;
  double ge = 0.0;
  for (int i = threadIdx.x; i < xdim; i += 256) {
    const T m = (T)mask[i], mb = T(1) - m;
    const T xi = x[row + i], vi = v[row + i];
    const T s0 = s ? s[row + i] : T(0), q0 = q ? q[row + i] : T(0), ti = t ? t[row + i] : T(0);
    const T sp = sg * eps * s0, qp = eps * q0;
    const T es = Num<T>::exp_(sp), eq = Num<T>::exp_(qp);
    const T u = vi * eq + ti;            // tr = eps * u
    const T tr = eps * u;
    const T go = gout[row + i];
    const T g = go * mb;                 // gradient reaching the updated branch xn
    const T glj = gl * mb;               // gradient reaching the per-element log-Jacobian
    T g_x, g_sp, g_tr;
    if (use_ncp) {
      const T hx = xi / T(2);
      const T sn = Num<T>::sin_(hx), cs = Num<T>::cos_(hx);
      const T den = cs * cs + es * es * sn * sn;
      const T y = Num<T>::tan_(hx) * es;
      const T dx1_dsp = T(2) * y / (T(1) + y * y);
      const T dlj_dsp = T(1) - T(2) * es * es * sn * sn / den;
      const T dlj_dx = -(sn * cs * (es * es - T(1))) / den;
      g_x = g * (es / den) + glj * dlj_dx;
      if (sign > 0) {                    // xn = x1 + tr
        g_sp = g * dx1_dsp + glj * dlj_dsp;
        g_tr = g;
      } else {                           // xn = x1 - es tr
        g_sp = g * (dx1_dsp - es * tr) + glj * dlj_dsp;
        g_tr = -g * es;
      }
    } else {
      if (sign > 0) {                    // xn = x es + tr, lj = sp
        g_x = g * es;
        g_sp = g * xi * es + glj;
        g_tr = g;
      } else {                           // xn = es (x - tr), lj = sp
        g_x = g * es;
        g_sp = g * es * (xi - tr) + glj;
        g_tr = -g * es;
      }
    }
    gx[row + i] = go * m + g_x;
    const T g_u = g_tr * eps;            // tr = eps u
    ge += (double)(g_tr * u);
    gv[row + i] = g_u * eq;
    const T g_qp = g_u * vi * eq;
    if (gt) gt[row + i] = g_u;
    if (gq) gq[row + i] = g_qp * eps;
    ge += (double)(g_qp * q0);
    if (gs) gs[row + i] = g_sp * sg * eps;
    ge += (double)(g_sp * sg * s0);
  }
  ge = block_sum<256>(ge, red, threadIdx.x);
  if (threadIdx.x == 0) geps[blockIdx.x] = (T)ge;
}

// out[b, :] = scale[b] * in[b, :]   (chain-wise scaling: action / kinetic-energy adjoints)
template <typename T>
__global__ void __launch_bounds__(256) k_rowscale(const T* __restrict__ in, const T* __restrict__ scale,
                                                  T* __restrict__ out, int xdim) {
  const size_t row = (size_t)blockIdx.y * xdim;
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i < xdim) out[row + i] = scale[blockIdx.y] * in[row + i];
}

int check_u1(int nb, int Tt, int X, int dtype) {
  L2B_REQUIRE(nb > 0 && Tt > 0 && X > 0, L2B_ERR_INVALID, "non-positive size: nb=%d T=%d X=%d", nb, Tt, X);
  L2B_REQUIRE(dtype == L2B_F32 || dtype == L2B_F64, L2B_ERR_UNSUPPORTED, "unknown dtype %d", dtype);
  L2B_REQUIRE((long long)Tt * X < (1ll << 28), L2B_ERR_UNSUPPORTED, "lattice too large");
  return L2B_OK;
}

size_t hmc_smem_bytes(int Tt, int X, int dtype) {
  return (size_t)5 * Tt * X * (dtype == L2B_F64 ? 8 : 4);
}

template <typename T, int NT>
int launch_hmc(const void* x, const void* v, double beta, double eps, int nlf, void* xo, void* vo, void* en, int nb,
               int Tt, int X, size_t smem, cudaStream_t st) {
  L2B_CUDA(cudaFuncSetAttribute(k_u1_hmc<T, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k_u1_hmc<T, NT><<<nb, NT, smem, st>>>((const T*)x, (const T*)v, (T)beta, (T)eps, nlf, (T*)xo, (T*)vo, (T*)en, Tt, X);
  L2B_LAUNCHED("k_u1_hmc");
  return L2B_OK;
}

template <typename T>
int dispatch_hmc(const void* x, const void* v, double beta, double eps, int nlf, void* xo, void* vo, void* en,
                 int nb, int Tt, int X, size_t smem, cudaStream_t st) {
  const int N = Tt * X;
  if (N >= 2048) return launch_hmc<T, 512>(x, v, beta, eps, nlf, xo, vo, en, nb, Tt, X, smem, st);
  if (N >= 512) return launch_hmc<T, 256>(x, v, beta, eps, nlf, xo, vo, en, nb, Tt, X, smem, st);
  return launch_hmc<T, 128>(x, v, beta, eps, nlf, xo, vo, en, nb, Tt, X, smem, st);
}

#define L2B_DISPATCH_T(dtype, CALL_F32, CALL_F64) \
  do {                                            \
    if ((dtype) == L2B_F32) { CALL_F32; }         \
    else { CALL_F64; }                            \
  } while (0)

}  // namespace
}  // namespace l2b

using namespace l2b;

namespace l2b {
namespace {
template <typename T, int MODE>
int launch_u1_heads(int hp, dim3 grid, cudaStream_t st, const void* z, int H, const void* const w[3],
                    const void* const b[3], const void* cs, const void* cq, const double nw[3], const void* a,
                    const void* bf, const float* mask, double eps, const void* eps_dev, int sign, int use_ncp,
                    void* out, double* part, int nb, int xdim, int cpb) {
#define L2B_U1H(HP)                                                                                              \
  k_u1_heads_update<T, HP, MODE><<<grid, 256, 0, st>>>(                                                           \
      (const T*)z, H, (const T*)w[0], (const T*)w[1], (const T*)w[2], (const T*)b[0], (const T*)b[1],             \
      (const T*)b[2], (const T*)cs, (const T*)cq, (T)nw[0], (T)nw[1], (T)nw[2], (const T*)a, (const T*)bf, mask,   \
      (T)eps, (const T*)eps_dev, sign, use_ncp, (T*)out, part, nb, xdim, cpb)
  if (hp == 8) L2B_U1H(8);
  else if (hp == 16) L2B_U1H(16);
  else L2B_U1H(32);
#undef L2B_U1H
  L2B_LAUNCHED("k_u1_heads_update");
  return L2B_OK;
}
}  // namespace
}  // namespace l2b

extern "C" {

size_t l2b_u1_ws_bytes(int nb, int T, int X, int dtype) {
  (void)nb; (void)T; (void)X; (void)dtype;
  return 0;  // every U(1) reduction is one block per chain: no scratch needed
}

int l2b_u1_wilson_loops(const void* x, void* w, int nb, int T, int X, int dtype, void* stream) {
  if (int rc = check_u1(nb, T, X, dtype)) return rc;
  L2B_REQUIRE(x && w, L2B_ERR_INVALID, "null pointer");
  L2B_REQUIRE(nb <= 65535, L2B_ERR_UNSUPPORTED, "nb exceeds grid.y limit");
  const dim3 grid((T * X + 255) / 256, nb);
  cudaStream_t st = (cudaStream_t)stream;
  L2B_DISPATCH_T(dtype, (k_u1_wloops<float><<<grid, 256, 0, st>>>((const float*)x, (float*)w, T, X)),
                 (k_u1_wloops<double><<<grid, 256, 0, st>>>((const double*)x, (double*)w, T, X)));
  L2B_LAUNCHED("k_u1_wloops");
  return L2B_OK;
}

int l2b_u1_wilson_loops4x4(const void* x, void* w, int nb, int T, int X, int dtype, void* stream) {
  if (int rc = check_u1(nb, T, X, dtype)) return rc;
  L2B_REQUIRE(x && w, L2B_ERR_INVALID, "null pointer");
  L2B_REQUIRE(nb <= 65535, L2B_ERR_UNSUPPORTED, "nb exceeds grid.y limit");
  const dim3 grid((T * X + 255) / 256, nb);
  cudaStream_t st = (cudaStream_t)stream;
  L2B_DISPATCH_T(dtype, (k_u1_wloops4x4<float><<<grid, 256, 0, st>>>((const float*)x, (float*)w, T, X)),
                 (k_u1_wloops4x4<double><<<grid, 256, 0, st>>>((const double*)x, (double*)w, T, X)));
  L2B_LAUNCHED("k_u1_wloops4x4");
  return L2B_OK;
}

int l2b_u1_observables(const void* x, double beta, void* obs, int nb, int T, int X, int dtype, void* stream) {
  if (int rc = check_u1(nb, T, X, dtype)) return rc;
  L2B_REQUIRE(x && obs, L2B_ERR_INVALID, "null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  L2B_DISPATCH_T(dtype, (k_u1_obs<float><<<nb, 256, 0, st>>>((const float*)x, (float)beta, (float*)obs, T, X)),
                 (k_u1_obs<double><<<nb, 256, 0, st>>>((const double*)x, beta, (double*)obs, T, X)));
  L2B_LAUNCHED("k_u1_obs");
  return L2B_OK;
}

int l2b_u1_force(const void* x, double beta, void* force, int nb, int T, int X, int dtype, void* stream) {
  if (int rc = check_u1(nb, T, X, dtype)) return rc;
  L2B_REQUIRE(x && force, L2B_ERR_INVALID, "null pointer");
  L2B_REQUIRE(nb <= 65535, L2B_ERR_UNSUPPORTED, "nb exceeds grid.y limit");
  const dim3 grid((T * X + 255) / 256, nb);
  cudaStream_t st = (cudaStream_t)stream;
  L2B_DISPATCH_T(dtype, (k_u1_force<float><<<grid, 256, 0, st>>>((const float*)x, (float)beta, (float*)force, T, X)),
                 (k_u1_force<double><<<grid, 256, 0, st>>>((const double*)x, beta, (double*)force, T, X)));
  L2B_LAUNCHED("k_u1_force");
  return L2B_OK;
}

int l2b_u1_hmc_trajectory(const void* x, const void* v, double beta, double eps, int nlf, void* x_prop,
                          void* v_prop, void* energies, int nb, int T, int X, int dtype, void* stream) {
  if (int rc = check_u1(nb, T, X, dtype)) return rc;
  L2B_REQUIRE(x && v && x_prop && v_prop && energies, L2B_ERR_INVALID, "null pointer");
  L2B_REQUIRE(nlf >= 1, L2B_ERR_INVALID, "nlf must be >= 1 (got %d)", nlf);
  const size_t smem = hmc_smem_bytes(T, X, dtype);
  L2B_REQUIRE(smem <= 227 * 1024, L2B_ERR_UNSUPPORTED,
              "U(1) %dx%d chain state (%zu B) does not fit one SM's shared memory", T, X, smem);
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == L2B_F32) return dispatch_hmc<float>(x, v, beta, eps, nlf, x_prop, v_prop, energies, nb, T, X, smem, st);
  return dispatch_hmc<double>(x, v, beta, eps, nlf, x_prop, v_prop, energies, nb, T, X, smem, st);
}

int l2b_u1_vupdate(const void* v, const void* force, const void* s, const void* t, const void* q, double eps, const void* eps_dev,
                   int sign, void* v_out, void* logdet, int nb, int xdim, int dtype, void* stream) {
  L2B_REQUIRE(nb > 0 && xdim > 0, L2B_ERR_INVALID, "non-positive size");
  L2B_REQUIRE(dtype == L2B_F32 || dtype == L2B_F64, L2B_ERR_UNSUPPORTED, "unknown dtype %d", dtype);
  L2B_REQUIRE(v && force && v_out, L2B_ERR_INVALID, "null pointer");
  L2B_REQUIRE(sign == 1 || sign == -1, L2B_ERR_INVALID, "sign must be +1 or -1");
  cudaStream_t st = (cudaStream_t)stream;
  L2B_DISPATCH_T(dtype,
                 (k_u1_vupdate<float><<<nb, 256, 0, st>>>((const float*)v, (const float*)force, (const float*)s,
                                                          (const float*)t, (const float*)q, (float)eps, (const float*)eps_dev, sign,
                                                          (float*)v_out, (float*)logdet, xdim)),
                 (k_u1_vupdate<double><<<nb, 256, 0, st>>>((const double*)v, (const double*)force, (const double*)s,
                                                           (const double*)t, (const double*)q, eps, (const double*)eps_dev, sign,
                                                           (double*)v_out, (double*)logdet, xdim)));
  L2B_LAUNCHED("k_u1_vupdate");
  return L2B_OK;
}

int l2b_u1_xupdate(const void* x, const void* v, const void* s, const void* t, const void* q, const float* mask,
                   double eps, const void* eps_dev, int sign, int use_ncp, void* x_out, void* logdet, int nb, int xdim, int dtype,
                   void* stream) {
  L2B_REQUIRE(nb > 0 && xdim > 0, L2B_ERR_INVALID, "non-positive size");
  L2B_REQUIRE(dtype == L2B_F32 || dtype == L2B_F64, L2B_ERR_UNSUPPORTED, "unknown dtype %d", dtype);
  L2B_REQUIRE(x && v && mask && x_out, L2B_ERR_INVALID, "null pointer");
  L2B_REQUIRE(sign == 1 || sign == -1, L2B_ERR_INVALID, "sign must be +1 or -1");
  cudaStream_t st = (cudaStream_t)stream;
  L2B_DISPATCH_T(dtype,
                 (k_u1_xupdate<float><<<nb, 256, 0, st>>>((const float*)x, (const float*)v, (const float*)s,
                                                          (const float*)t, (const float*)q, mask, (float)eps, (const float*)eps_dev, sign,
                                                          use_ncp, (float*)x_out, (float*)logdet, xdim)),
                 (k_u1_xupdate<double><<<nb, 256, 0, st>>>((const double*)x, (const double*)v, (const double*)s,
                                                           (const double*)t, (const double*)q, mask, eps, (const double*)eps_dev, sign,
                                                           use_ncp, (double*)x_out, (double*)logdet, xdim)));
  L2B_LAUNCHED("k_u1_xupdate");
  return L2B_OK;
}

int l2b_u1_kinetic(const void* v, void* ke, int nb, int xdim, int dtype, void* stream) {
  L2B_REQUIRE(nb > 0 && xdim > 0, L2B_ERR_INVALID, "non-positive size");
  L2B_REQUIRE(dtype == L2B_F32 || dtype == L2B_F64, L2B_ERR_UNSUPPORTED, "unknown dtype %d", dtype);
  L2B_REQUIRE(v && ke, L2B_ERR_INVALID, "null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  L2B_DISPATCH_T(dtype, (k_u1_kinetic<float><<<nb, 256, 0, st>>>((const float*)v, (float*)ke, xdim)),
                 (k_u1_kinetic<double><<<nb, 256, 0, st>>>((const double*)v, (double*)ke, xdim)));
  L2B_LAUNCHED("k_u1_kinetic");
  return L2B_OK;
}

int l2b_u1_compat_proj(const void* x, void* out, size_t n, int dtype, void* stream) {
  L2B_REQUIRE(dtype == L2B_F32 || dtype == L2B_F64, L2B_ERR_UNSUPPORTED, "unknown dtype %d", dtype);
  L2B_REQUIRE(x && out, L2B_ERR_INVALID, "null pointer");
  if (n == 0) return L2B_OK;
  const unsigned nblk = (unsigned)((n + 255) / 256);
  cudaStream_t st = (cudaStream_t)stream;
  L2B_DISPATCH_T(dtype, (k_u1_wrap<float><<<nblk, 256, 0, st>>>((const float*)x, (float*)out, n)),
                 (k_u1_wrap<double><<<nblk, 256, 0, st>>>((const double*)x, (double*)out, n)));
  L2B_LAUNCHED("k_u1_wrap");
  return L2B_OK;
}

int l2b_u1_wilson_loops_bwd(const void* gw, void* gx, int nb, int T, int X, int dtype, void* stream) {
  if (int rc = check_u1(nb, T, X, dtype)) return rc;
  L2B_REQUIRE(gw && gx, L2B_ERR_INVALID, "null pointer");
  L2B_REQUIRE(nb <= 65535, L2B_ERR_UNSUPPORTED, "nb exceeds grid.y limit");
  const dim3 grid((T * X + 255) / 256, nb);
  cudaStream_t st = (cudaStream_t)stream;
  L2B_DISPATCH_T(dtype, (k_u1_wloops_bwd<float><<<grid, 256, 0, st>>>((const float*)gw, (float*)gx, T, X)),
                 (k_u1_wloops_bwd<double><<<grid, 256, 0, st>>>((const double*)gw, (double*)gx, T, X)));
  L2B_LAUNCHED("k_u1_wloops_bwd");
  return L2B_OK;
}

int l2b_u1_force_bwd(const void* x, double beta, const void* gforce, void* gx, int nb, int T, int X, int dtype,
                     void* stream) {
  if (int rc = check_u1(nb, T, X, dtype)) return rc;
  L2B_REQUIRE(x && gforce && gx, L2B_ERR_INVALID, "null pointer");
  L2B_REQUIRE(nb <= 65535, L2B_ERR_UNSUPPORTED, "nb exceeds grid.y limit");
  const dim3 grid((T * X + 255) / 256, nb);
  cudaStream_t st = (cudaStream_t)stream;
  L2B_DISPATCH_T(dtype,
                 (k_u1_force_bwd<float><<<grid, 256, 0, st>>>((const float*)x, (float)beta, (const float*)gforce,
                                                             (float*)gx, T, X)),
                 (k_u1_force_bwd<double><<<grid, 256, 0, st>>>((const double*)x, beta, (const double*)gforce,
                                                              (double*)gx, T, X)));
  L2B_LAUNCHED("k_u1_force_bwd");
  return L2B_OK;
}

int l2b_u1_vupdate_bwd(const void* v, const void* force, const void* s, const void* t, const void* q, double eps, const void* eps_dev,
                       int sign, const void* gv_out, const void* glogdet, void* gv, void* gforce, void* gs, void* gt,
                       void* gq, void* geps, int nb, int xdim, int dtype, void* stream) {
  L2B_REQUIRE(nb > 0 && xdim > 0, L2B_ERR_INVALID, "non-positive size");
  L2B_REQUIRE(dtype == L2B_F32 || dtype == L2B_F64, L2B_ERR_UNSUPPORTED, "unknown dtype %d", dtype);
  L2B_REQUIRE(v && force && gv_out && gv && gforce && geps, L2B_ERR_INVALID, "null pointer");
  L2B_REQUIRE(sign == 1 || sign == -1, L2B_ERR_INVALID, "sign must be +1 or -1");
  cudaStream_t st = (cudaStream_t)stream;
  L2B_DISPATCH_T(dtype,
                 (k_u1_vupdate_bwd<float><<<nb, 256, 0, st>>>(
                     (const float*)v, (const float*)force, (const float*)s, (const float*)t, (const float*)q, (float)eps, (const float*)eps_dev,
                     sign, (const float*)gv_out, (const float*)glogdet, (float*)gv, (float*)gforce, (float*)gs,
                     (float*)gt, (float*)gq, (float*)geps, xdim)),
                 (k_u1_vupdate_bwd<double><<<nb, 256, 0, st>>>(
                     (const double*)v, (const double*)force, (const double*)s, (const double*)t, (const double*)q, eps, (const double*)eps_dev,
                     sign, (const double*)gv_out, (const double*)glogdet, (double*)gv, (double*)gforce, (double*)gs,
                     (double*)gt, (double*)gq, (double*)geps, xdim)));
  L2B_LAUNCHED("k_u1_vupdate_bwd");
  return L2B_OK;
}

int l2b_u1_xupdate_bwd(const void* x, const void* v, const void* s, const void* t, const void* q, const float* mask,
                       double eps, const void* eps_dev, int sign, int use_ncp, const void* gx_out, const void* glogdet, void* gx, void* gv,
                       void* gs, void* gt, void* gq, void* geps, int nb, int xdim, int dtype, void* stream) {
  L2B_REQUIRE(nb > 0 && xdim > 0, L2B_ERR_INVALID, "non-positive size");
  L2B_REQUIRE(dtype == L2B_F32 || dtype == L2B_F64, L2B_ERR_UNSUPPORTED, "unknown dtype %d", dtype);
  L2B_REQUIRE(x && v && mask && gx_out && gx && gv && geps, L2B_ERR_INVALID, "null pointer");
  L2B_REQUIRE(sign == 1 || sign == -1, L2B_ERR_INVALID, "sign must be +1 or -1");
  cudaStream_t st = (cudaStream_t)stream;
  L2B_DISPATCH_T(dtype,
                 (k_u1_xupdate_bwd<float><<<nb, 256, 0, st>>>(
                     (const float*)x, (const float*)v, (const float*)s, (const float*)t, (const float*)q, mask,
                     (float)eps, (const float*)eps_dev, sign, use_ncp, (const float*)gx_out, (const float*)glogdet, (float*)gx, (float*)gv,
                     (float*)gs, (float*)gt, (float*)gq, (float*)geps, xdim)),
                 (k_u1_xupdate_bwd<double><<<nb, 256, 0, st>>>(
                     (const double*)x, (const double*)v, (const double*)s, (const double*)t, (const double*)q, mask, eps, (const double*)eps_dev,
                     sign, use_ncp, (const double*)gx_out, (const double*)glogdet, (double*)gx, (double*)gv,
                     (double*)gs, (double*)gt, (double*)gq, (double*)geps, xdim)));
  L2B_LAUNCHED("k_u1_xupdate_bwd");
  return L2B_OK;
}

int l2b_rowscale(const void* in, const void* scale, void* out, int nb, int xdim, int dtype, void* stream) {
  L2B_REQUIRE(nb > 0 && xdim > 0, L2B_ERR_INVALID, "non-positive size");
  L2B_REQUIRE(dtype == L2B_F32 || dtype == L2B_F64, L2B_ERR_UNSUPPORTED, "unknown dtype %d", dtype);
  L2B_REQUIRE(in && scale && out, L2B_ERR_INVALID, "null pointer");
  L2B_REQUIRE(nb <= 65535, L2B_ERR_UNSUPPORTED, "nb exceeds grid.y limit");
  const dim3 grid((xdim + 255) / 256, nb);
  cudaStream_t st = (cudaStream_t)stream;
  L2B_DISPATCH_T(dtype, (k_rowscale<float><<<grid, 256, 0, st>>>((const float*)in, (const float*)scale, (float*)out, xdim)),
                 (k_rowscale<double><<<grid, 256, 0, st>>>((const double*)in, (const double*)scale, (double*)out, xdim)));
  L2B_LAUNCHED("k_rowscale");
  return L2B_OK;
}

size_t l2b_u1_heads_ws_bytes(int nb, int xdim) {
  if (nb <= 0 || xdim <= 0) return 0;
  return align_up((size_t)nb * ((xdim + 255) / 256) * sizeof(double), 256);
}

int l2b_u1_heads_update(int mode, const void* z, int hidden, const void* w_s, const void* w_t, const void* w_q,
                        const void* b_s, const void* b_t, const void* b_q, const void* coeff_s, const void* coeff_q,
                        double nw_s, double nw_t, double nw_q, const void* a, const void* b, const float* mask,
                        double eps, const void* eps_dev, int sign, int use_ncp, void* out, void* logdet, int nb,
                        int xdim, int dtype, void* ws, size_t ws_bytes, void* stream) {
  L2B_REQUIRE(nb > 0 && xdim > 0, L2B_ERR_INVALID, "non-positive size");
  L2B_REQUIRE(dtype == L2B_F32 || dtype == L2B_F64, L2B_ERR_UNSUPPORTED, "unknown dtype %d", dtype);
  L2B_REQUIRE(mode == 0 || mode == 1, L2B_ERR_INVALID, "mode must be 0 (v-update) or 1 (x-update)");
  L2B_REQUIRE(hidden > 0 && hidden <= 32, L2B_ERR_UNSUPPORTED, "fused U(1) heads need hidden <= 32 (got %d)", hidden);
  L2B_REQUIRE(z && w_s && w_t && w_q && b_s && b_t && b_q && coeff_s && coeff_q && a && b && out, L2B_ERR_INVALID,
              "null pointer");
  L2B_REQUIRE(mode == 0 || mask != nullptr, L2B_ERR_INVALID, "the x-update needs a mask");
  L2B_REQUIRE(sign == 1 || sign == -1, L2B_ERR_INVALID, "sign must be +1 or -1");
  const int nblk = (xdim + 255) / 256;
  double* part = nullptr;
  if (logdet) {
    L2B_REQUIRE(ws && ws_bytes >= l2b_u1_heads_ws_bytes(nb, xdim), L2B_ERR_WORKSPACE, "workspace too small");
    part = (double*)ws;
  }
  // chains per block: 64, halved while the grid is short of two blocks per SM (results do not depend on it: the
  // per-chain sums run over the columns of a block, in a fixed order)
  int cpb = kHeadsChains;
  while (cpb > 4 && (long long)nblk * ((nb + cpb - 1) / cpb) < 2 * 148) cpb >>= 1;
  const int nyb = (nb + cpb - 1) / cpb;
  L2B_REQUIRE(nyb <= 65535, L2B_ERR_UNSUPPORTED, "too many chains for one launch");
  const dim3 grid(nblk, nyb);
  const int hp = hidden <= 8 ? 8 : (hidden <= 16 ? 16 : 32);
  const void* const w[3] = {w_s, w_t, w_q};
  const void* const bb[3] = {b_s, b_t, b_q};
  const double nw[3] = {nw_s, nw_t, nw_q};
  cudaStream_t st = (cudaStream_t)stream;
  int rc;
  if (dtype == L2B_F32)
    rc = mode == 0 ? launch_u1_heads<float, 0>(hp, grid, st, z, hidden, w, bb, coeff_s, coeff_q, nw, a, b, mask, eps, eps_dev, sign, use_ncp, out, part, nb, xdim, cpb)
                   : launch_u1_heads<float, 1>(hp, grid, st, z, hidden, w, bb, coeff_s, coeff_q, nw, a, b, mask, eps, eps_dev, sign, use_ncp, out, part, nb, xdim, cpb);
  else
    rc = mode == 0 ? launch_u1_heads<double, 0>(hp, grid, st, z, hidden, w, bb, coeff_s, coeff_q, nw, a, b, mask, eps, eps_dev, sign, use_ncp, out, part, nb, xdim, cpb)
                   : launch_u1_heads<double, 1>(hp, grid, st, z, hidden, w, bb, coeff_s, coeff_q, nw, a, b, mask, eps, eps_dev, sign, use_ncp, out, part, nb, xdim, cpb);
  if (rc != L2B_OK) return rc;
  if (logdet) {
    if (dtype == L2B_F32) k_u1_sum_rows<float><<<nb, 256, 0, st>>>(part, nblk, (float*)logdet);
    else k_u1_sum_rows<double><<<nb, 256, 0, st>>>(part, nblk, (double*)logdet);
    L2B_LAUNCHED("k_u1_sum_rows");
  }
  return L2B_OK;
}

size_t l2b_u1_input_ws_bytes(int nb, int xdim) {
  if (nb <= 0 || xdim <= 0) return 0;
  return align_up((size_t)((xdim + 255) / 256) * nb * 16 * sizeof(double), 256);
}

int l2b_u1_input_layer(int mode, const void* x, const void* v, const float* mask, const void* w_x, const void* b_x,
                       const void* w_v, const void* b_v, int units, void* pre, int nb, int xdim, int dtype, void* ws,
                       size_t ws_bytes, void* stream) {
  L2B_REQUIRE(nb > 0 && xdim > 0, L2B_ERR_INVALID, "non-positive size");
  L2B_REQUIRE(dtype == L2B_F32 || dtype == L2B_F64, L2B_ERR_UNSUPPORTED, "unknown dtype %d", dtype);
  L2B_REQUIRE(mode == 0 || mode == 1, L2B_ERR_INVALID, "mode must be 0 (vnet) or 1 (xnet)");
  L2B_REQUIRE(units > 0 && units <= 16, L2B_ERR_UNSUPPORTED, "fused U(1) input layer needs units[0] <= 16 (got %d)", units);
  L2B_REQUIRE(x && v && w_x && b_x && w_v && b_v && pre, L2B_ERR_INVALID, "null pointer");
  L2B_REQUIRE(mode == 0 || mask != nullptr, L2B_ERR_INVALID, "the xnet input needs the mask");
  L2B_REQUIRE(ws && ws_bytes >= l2b_u1_input_ws_bytes(nb, xdim), L2B_ERR_WORKSPACE, "workspace too small");
  const int ncolblk = (xdim + 255) / 256;
  const int cb = dtype == L2B_F32 ? 64 : 32;
  const int nyb = (nb + cb - 1) / cb;
  L2B_REQUIRE(nyb <= 65535, L2B_ERR_UNSUPPORTED, "too many chains for one launch");
  const dim3 grid(ncolblk, nyb);
  cudaStream_t st = (cudaStream_t)stream;
#define L2B_U1IN(T, MODE)                                                                                           \
  k_u1_input_layer<T, 16, MODE><<<grid, 256, 0, st>>>((const T*)x, (const T*)v, mask, (const T*)w_x, (const T*)w_v,  \
                                                       units, (T*)ws, nb, xdim)
  if (dtype == L2B_F32) { if (mode == 1) L2B_U1IN(float, 1); else L2B_U1IN(float, 0); }
  else { if (mode == 1) L2B_U1IN(double, 1); else L2B_U1IN(double, 0); }
#undef L2B_U1IN
  L2B_LAUNCHED("k_u1_input_layer");
  const int nfin = (nb * units + 255) / 256;
  if (dtype == L2B_F32)
    k_u1_input_finish<float, 16><<<nfin, 256, 0, st>>>((const float*)ws, ncolblk, (const float*)b_x, (const float*)b_v, units, (float*)pre, nb);
  else
    k_u1_input_finish<double, 16><<<nfin, 256, 0, st>>>((const double*)ws, ncolblk, (const double*)b_x, (const double*)b_v, units, (double*)pre, nb);
  L2B_LAUNCHED("k_u1_input_finish");
  return L2B_OK;
}

}  // extern "C"