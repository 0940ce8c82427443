// l2b_vnet.cu -- tensor-core side of the L2HMC momentum update (sm_100a, tcgen05).
//
// k_heads_vupdate fuses the three output heads of the vnet `LeapfrogLayer`
// (reference network/pytorch/network.py:536-548: scale / transl / transf) with
// the momentum-update epilogue that consumes them (dynamics.py:1266-1297):
//
//   s = a_s tanh(W_s z + b_s),  t = a_t (W_t z + b_t),  q = a_q tanh(W_q z + b_q)      [nb, xdim] each
//   fwd: v' = e^{eps s/2} v - eps/2 (F e^{eps q} + t)     logdet = sum  eps s/2
//   bwd: v' = e^{-eps s/2} (v + eps/2 (F e^{eps q} + t))  logdet = sum -eps s/2
//
// so that s, t, q (3 x nb x xdim numbers) never exist in HBM: the only field-sized
// traffic is v (r), F (r), v' (w) plus the bf16 weights.
//
// GEMM shape: the xdim outputs are the UMMA M axis (128 per CTA), the chains are the
// UMMA N axis (64 per CTA), K = hidden units.  One CTA = one (128-column, 64-chain) tile:
//   * the z tile (B operand, all of K) is written to shared memory in the canonical
//     no-swizzle K-major core-matrix layout by the CTA's threads;
//   * the weight tile (A operand) streams in with 1-D TMA bulk copies
//     (cp.async.bulk ... mbarrier::complete_tx) from a PRE-PACKED bf16 image whose 16 KB
//     stages are byte-for-byte the shared-memory layout (packed once per optimizer step
//     by k_pack_heads, which is also the fp32 -> bf16 cast);
//   * one elected thread issues tcgen05.mma (cta_group::1, kind::f16, M128 N64 K16) into three
//     fp32 accumulators in TMEM (3 x 64 columns), tcgen05.commit recycles the stages;
//   * the epilogue reads TMEM with tcgen05.ld (lane = xdim column, column = chain), so the
//     v / F / v' accesses of a warp are 512 contiguous bytes per chain.
// (Measured negative: a cp.async.bulk.prefetch.L2 of the tile's v / F rows at kernel start made the
// kernel 14 % slower -- the epilogue's own 16 loads in flight per thread already cover the latency.)
// Two CTAs of 8 warps are resident per SM (~107 KB smem, 256 TMEM columns each): one CTA's
// HBM-bound epilogue overlaps the other's weight streaming and MMAs.
#include <cuda_bf16.h>

#include "l2b_common.cuh"
#include "l2b_tc.cuh"

namespace l2b {
namespace {

constexpr int BM = 128;                       // xdim columns per tile (UMMA M)
constexpr int BN = 64;                        // chains per tile (UMMA N)
constexpr int KC = 64;                        // K per pipeline stage
constexpr int NST = 4;                        // weight stages in flight
constexpr int A_STAGE_BYTES = BM * KC * 2;    // 16 KB
constexpr int KMAX = 256;                     // largest (padded) hidden size with z resident
constexpr int TMEM_COLS = 256;                // 3 x BN = 192 accumulator columns, power of two
constexpr int NTH = 256;                       // 8 warps: 1 TMA/MMA thread, all 8 warps in the epilogue
constexpr int CH = 8;                          // chains per epilogue chunk

struct Smem {
  alignas(128) unsigned char a[NST][A_STAGE_BYTES];     // weight ring
  alignas(128) unsigned char b[BN * KMAX * 2];          // z tile, [kcore][chain][8]
  alignas(8) unsigned long long full[NST];
  unsigned long long empty[NST];
  unsigned long long accum;
  float lj[NTH / 32][CH * 33];                          // per-warp log-Jacobian staging (padded rows)
  float ld[NTH / 32][32];                               // logdet partials: [warp][chain of its half]
  uint32_t tmem_base;
};

// instruction descriptor, kind::f16: D = f32, A = B = bf16, both K-major, M = 128, N = BN
constexpr uint32_t kIdesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);

// ---------------------------------------------------------------------------
// weight packing: W_h[xdim, H] (nn.Linear layout, f64 / f32 / bf16) -> bf16 image
//   packed[tile][head][kcore][row 0..127][8],  kcore < KP/8,  zero padded rows / columns
// one thread per 16-byte chunk
// ---------------------------------------------------------------------------
template <typename W>
__device__ __forceinline__ float w_load(const W* p) { return (float)*p; }
template <>
__device__ __forceinline__ float w_load<__nv_bfloat16>(const __nv_bfloat16* p) { return __bfloat162float(*p); }

template <typename W>
__global__ void __launch_bounds__(256) k_pack_heads(const W* __restrict__ w0, const W* __restrict__ w1,
                                                    const W* __restrict__ w2, uint4* __restrict__ packed, int xdim,
                                                    int H, int KP, size_t nchunks) {
  const size_t id = (size_t)blockIdx.x * 256 + threadIdx.x;
  if (id >= nchunks) return;
  const int r = (int)(id % BM);
  size_t rest = id / BM;
  const int kcore = (int)(rest % (KP / 8));
  rest /= (KP / 8);
  const int head = (int)(rest % 3);
  const size_t tile = rest / 3;
  const size_t row = tile * BM + r;
  const W* w = head == 0 ? w0 : (head == 1 ? w1 : w2);
  __align__(16) __nv_bfloat16 h[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const int k = kcore * 8 + e;
    h[e] = __float2bfloat16((row < (size_t)xdim && k < H) ? w_load(w + row * H + k) : 0.0f);
  }
  packed[id] = *reinterpret_cast<const uint4*>(h);
}

// ---------------------------------------------------------------------------
// the fused kernel
// ---------------------------------------------------------------------------
struct HeadsArgs {
  const __nv_bfloat16* z;       // [nb, H]
  const unsigned char* packed;  // k_pack_heads image
  const float* bias[3];         // [xdim] each
  const float* scale_s;         // [xdim]  nw.s * exp(coeff_s)
  const float* scale_q;         // [xdim]  nw.q * exp(coeff_q)
  float scale_t;                // nw.t
  const double2* v;             // [nb, xdim] complex128
  const double2* f;
  double2* out;
  double* part;                 // [nb, ntiles] logdet partials (may be null)
  float* stq;                   // optional [3, nb, xdim] fp32 dump of (s, t, q)
  double eps;                   // multiplies *eps_dev when that is given
  const double* eps_dev;        // device-resident step size (CUDA graphs) or null
  int sign, nb, xdim, H, KP, ntiles;
  // pair mode (sign2 != 0): a SECOND momentum update with the same (s, t, q, F) applied to the result of the
  // first, v'' = upd(+-upd(v; eps, sign); eps2, sign2) -- two consecutive v-updates of an L2HMC sweep between which
  // the links do not move see the same network outputs (dynamics.py:1187-1228: second half of leapfrog layer i,
  // first half of layer i+1; `negate` = the v -> -v of the turn-around, dynamics.py:1002).  One pass over
  // v, F and the weights instead of two; logdet is the sum of both.
  double eps2;
  const double* eps2_dev;
  int sign2, negate;
};

// one momentum update of one element (dynamics.py:1266-1297); returns the log-Jacobian term sign * eps * s / 2
__device__ __forceinline__ float vupd(double2& v, const double2 f, float s, float t, float q, float epsf, double he,
                                      bool fwd) {
  const float logjac = (fwd ? 0.5f : -0.5f) * epsf * s;
  const double es = (double)exp_fast(logjac), eq = (double)exp_fast(epsf * q);
  const double fr = fma(f.x, eq, (double)t), fi = f.y * eq;
  if (fwd) { v.x = fma(es, v.x, -he * fr); v.y = fma(es, v.y, -he * fi); }
  else { v.x = es * fma(he, fr, v.x); v.y = es * fma(he, fi, v.y); }
  return logjac;
}

// FULL: interior tile (all 64 chains and 128 columns valid, no s/t/q dump): no per-element predicates
template <bool FWD, bool FULL>
__device__ __forceinline__ void epilogue(const HeadsArgs& a, Smem& sm, uint32_t tmem, int tile, int chain0, int warp,
                                         int lane) {
  const int quarter = warp & 3, half = warp >> 2;
  const size_t j = (size_t)tile * BM + quarter * 32 + lane;
  const bool col_ok = FULL || j < (size_t)a.xdim;
  const float bs = col_ok ? __ldg(a.bias[0] + j) : 0.f, bt = col_ok ? __ldg(a.bias[1] + j) : 0.f,
              bq = col_ok ? __ldg(a.bias[2] + j) : 0.f;
  const float as = col_ok ? __ldg(a.scale_s + j) : 0.f, aq = col_ok ? __ldg(a.scale_q + j) : 0.f, at = a.scale_t;
  const double epsd = a.eps_dev ? a.eps * a.eps_dev[0] : a.eps;
  const float epsf = (float)epsd;
  const double he = 0.5 * epsd;
  const double epsd2 = a.eps2_dev ? a.eps2 * a.eps2_dev[0] : a.eps2;
  const float epsf2 = (float)epsd2;
  const double he2 = 0.5 * epsd2;
  const uint32_t trow = tmem + ((uint32_t)(quarter * 32) << 16);
  float* ljw = sm.lj[warp];
  const size_t xd = (size_t)a.xdim;
#pragma unroll 1
  for (int ck = 0; ck < 4; ++ck) {
    const int c0 = half * 32 + ck * CH;
    if (!FULL && chain0 + c0 >= a.nb) break;            // warp-uniform: nothing left for this warp
    uint32_t rs[CH], rt[CH], rq[CH];
    tmem_ld8(trow + 0 * BN + c0, rs);
    tmem_ld8(trow + 1 * BN + c0, rt);
    tmem_ld8(trow + 2 * BN + c0, rq);
    const size_t base = (size_t)(chain0 + c0) * xd + j;
    const double2* pv = a.v + base;
    const double2* pf = a.f + base;
    double2* po = a.out + base;
    double2 vv[CH], ff[CH];
#pragma unroll
    for (int c = 0; c < CH; ++c) {
      const bool ok = FULL || (col_ok && chain0 + c0 + c < a.nb);
      vv[c] = ok ? __ldg(pv + c * xd) : make_double2(0., 0.);
      ff[c] = ok ? __ldg(pf + c * xd) : make_double2(0., 0.);
    }
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int c = 0; c < CH; ++c) {
      const bool ok = FULL || (col_ok && chain0 + c0 + c < a.nb);
      const float s = as * tanhf(__uint_as_float(rs[c]) + bs);   // summed into logdet over xdim elements: accurate tanh
      const float t = at * (__uint_as_float(rt[c]) + bt);
      const float q = aq * tanh_fast(__uint_as_float(rq[c]) + bq);
      double2 o = vv[c];
      float logjac = vupd(o, ff[c], s, t, q, epsf, he, FWD);      // sign * eps * s / 2
      if (a.sign2 != 0) {
        if (a.negate) { o.x = -o.x; o.y = -o.y; }
        logjac += vupd(o, ff[c], s, t, q, epsf2, he2, a.sign2 > 0);
      }
      ljw[c * 33 + lane] = ok ? logjac : 0.f;
      if (ok) {
        po[c * xd] = o;
        if (!FULL && a.stq != nullptr) {
          const size_t plane = (size_t)a.nb * xd, at_ = base + c * xd;
          a.stq[at_] = s; a.stq[plane + at_] = t; a.stq[2 * plane + at_] = q;
        }
      }
    }
    __syncwarp();
    {                                                   // fixed-order sum over the warp's 32 columns:
      const int c = lane & 7, part = lane >> 3;         // lane (c, part) adds columns 8 part .. +7 of chain c
      float x = 0.f;
#pragma unroll
      for (int k = 0; k < 8; ++k) x += ljw[c * 33 + part * 8 + k];
      x += __shfl_xor_sync(0xffffffffu, x, 8);
      x += __shfl_xor_sync(0xffffffffu, x, 16);
      if (lane < CH) sm.ld[warp][ck * CH + lane] = x;
    }
    __syncwarp();
  }
}

// Interior tiles: software-pipelined epilogue.  4 chains per chunk, the v / F loads of chunk k+1 are in
// flight while chunk k is computed, and chunk 0's loads are issued BEFORE waiting for the accumulators, so
// the epilogue's first DRAM latency hides under the weight streaming + MMAs of the same CTA.
template <bool FWD>
__device__ __forceinline__ void epilogue_full(const HeadsArgs& a, Smem& sm, uint32_t tmem, int tile, int chain0,
                                              int warp, int lane) {
  constexpr int C4 = 4, NCK = 32 / C4;
  const int quarter = warp & 3, half = warp >> 2;
  const size_t j = (size_t)tile * BM + quarter * 32 + lane;
  const size_t xd = (size_t)a.xdim;
  const size_t base0 = (size_t)(chain0 + half * 32) * xd + j;
  const double2* __restrict__ pv = a.v + base0;
  const double2* __restrict__ pf = a.f + base0;
  double2* __restrict__ po = a.out + base0;
  double2 vv[2][C4], ff[2][C4];
#pragma unroll
  for (int c = 0; c < C4; ++c) { vv[0][c] = __ldg(pv + c * xd); ff[0][c] = __ldg(pf + c * xd); }
  const float bs = __ldg(a.bias[0] + j), bt = __ldg(a.bias[1] + j), bq = __ldg(a.bias[2] + j);
  const float as = __ldg(a.scale_s + j), aq = __ldg(a.scale_q + j), at = a.scale_t;
  const double epsd = a.eps_dev ? a.eps * a.eps_dev[0] : a.eps;
  const float epsf = (float)epsd;
  const double he = 0.5 * epsd;
  const double epsd2 = a.eps2_dev ? a.eps2 * a.eps2_dev[0] : a.eps2;
  const float epsf2 = (float)epsd2;
  const double he2 = 0.5 * epsd2;
  const uint32_t trow = tmem + ((uint32_t)(quarter * 32) << 16) + half * 32;
  float* ljw = sm.lj[warp];
  mbar_wait<64>(smem_u32(&sm.accum), 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
  for (int ck = 0; ck < NCK; ++ck) {
    const int cur = ck & 1, nxt = cur ^ 1;
    if (ck + 1 < NCK) {
#pragma unroll
      for (int c = 0; c < C4; ++c) {
        vv[nxt][c] = __ldg(pv + ((ck + 1) * C4 + c) * xd);
        ff[nxt][c] = __ldg(pf + ((ck + 1) * C4 + c) * xd);
      }
    }
    uint32_t rs[C4], rt[C4], rq[C4];
    tmem_ld4(trow + 0 * BN + ck * C4, rs);
    tmem_ld4(trow + 1 * BN + ck * C4, rt);
    tmem_ld4(trow + 2 * BN + ck * C4, rq);
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int c = 0; c < C4; ++c) {
      const float s = as * tanhf(__uint_as_float(rs[c]) + bs);   // summed into logdet over xdim elements: accurate tanh
      const float t = at * (__uint_as_float(rt[c]) + bt);
      const float q = aq * tanh_fast(__uint_as_float(rq[c]) + bq);
      double2 o = vv[cur][c];
      float logjac = vupd(o, ff[cur][c], s, t, q, epsf, he, FWD);
      if (a.sign2 != 0) {
        if (a.negate) { o.x = -o.x; o.y = -o.y; }
        logjac += vupd(o, ff[cur][c], s, t, q, epsf2, he2, a.sign2 > 0);
      }
      ljw[c * 33 + lane] = logjac;
      po[(ck * C4 + c) * xd] = o;
    }
    __syncwarp();
    {                                                   // fixed-order sum over the warp's 32 columns
      const int c = lane & 3, part = lane >> 2;         // lane (c, part) adds columns 4 part .. +3 of chain c
      float x = 0.f;
#pragma unroll
      for (int k = 0; k < 4; ++k) x += ljw[c * 33 + part * 4 + k];
      x += __shfl_xor_sync(0xffffffffu, x, 4);
      x += __shfl_xor_sync(0xffffffffu, x, 8);
      x += __shfl_xor_sync(0xffffffffu, x, 16);
      if (lane < C4) sm.ld[warp][ck * C4 + lane] = x;
    }
    __syncwarp();
  }
}

__global__ void __launch_bounds__(NTH, 2) k_heads_vupdate(const HeadsArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  Smem& sm = *reinterpret_cast<Smem*>(smem_raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nct = (a.nb + BN - 1) / BN;
  const int tile = blockIdx.x / nct;          // chain tiles of one column tile are adjacent:
  const int chain0 = (blockIdx.x % nct) * BN;  // they run together and share the weights in L2
  const int NKC = a.KP / KC;
  const int total = 3 * NKC;

  sm.ld[warp][lane] = 0.f;
  if (tid == 0) {
    for (int s = 0; s < NST; ++s) { mbar_init(smem_u32(&sm.full[s]), 1); mbar_init(smem_u32(&sm.empty[s]), 1); }
    mbar_init(smem_u32(&sm.accum), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    // first ring fill goes out NOW: the weights' DRAM latency overlaps the TMEM allocation and the z tile load
    const unsigned char* wsrc0 = a.packed + (size_t)tile * total * A_STAGE_BYTES;
    for (int s = 0; s < NST && s < total; ++s) {
      mbar_expect_tx(smem_u32(&sm.full[s]), A_STAGE_BYTES);
      bulk_g2s(smem_u32(sm.a) + s * A_STAGE_BYTES, wsrc0 + (size_t)s * A_STAGE_BYTES, A_STAGE_BYTES,
               smem_u32(&sm.full[s]));
    }
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sm.tmem_base)),
                 "n"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // z tile -> canonical K-major image [kcore][chain][8 bf16]; chains >= nb and k >= H are zero
  {
    const int nchunk = BN * (a.KP / 8);
    uint4* bimg = reinterpret_cast<uint4*>(sm.b);
    for (int idx = tid; idx < nchunk; idx += NTH) {
      const int c = idx % BN, i = idx / BN;
      const int b = chain0 + c;
      uint4 val = make_uint4(0, 0, 0, 0);
      if (b < a.nb && i * 8 < a.H) val = __ldg(reinterpret_cast<const uint4*>(a.z + (size_t)b * a.H + i * 8));
      bimg[idx] = val;
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the tensor core
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = sm.tmem_base;

  if (tid == 0) {
    // ---- TMA producer + MMA issuer (one thread) ------------------------------------------
    const unsigned char* wsrc = a.packed + (size_t)tile * total * A_STAGE_BYTES;
    const uint32_t a_base = smem_u32(sm.a), b_base = smem_u32(sm.b);
    const uint32_t a_lbo = BM * 16, b_lbo = BN * 16, sbo = 128;   // K step of each image; 8-row groups are adjacent
    for (int s = 0; s < total; ++s) {
      const int slot = s % NST;
      const uint32_t ph = (uint32_t)(s / NST) & 1u;
      mbar_wait(smem_u32(&sm.full[slot]), ph);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const int head = s / NKC, kc = s % NKC;
#pragma unroll
      for (int kk = 0; kk < KC / 16; ++kk) {
        const uint64_t ad = umma_desc(a_base + slot * A_STAGE_BYTES + (2 * kk) * (BM * 16), a_lbo, sbo);
        const uint64_t bd = umma_desc(b_base + (2 * (kc * (KC / 16) + kk)) * (BN * 16), b_lbo, sbo);
        umma_f16(tmem + head * BN, ad, bd, kIdesc, (kc | kk) != 0 ? 1u : 0u);
      }
      umma_commit(smem_u32(&sm.empty[slot]));         // arrives when these MMAs have read the stage
      if (s + NST < total) {
        mbar_wait(smem_u32(&sm.empty[slot]), ph);
        mbar_expect_tx(smem_u32(&sm.full[slot]), A_STAGE_BYTES);
        bulk_g2s(a_base + slot * A_STAGE_BYTES, wsrc + (size_t)(s + NST) * A_STAGE_BYTES, A_STAGE_BYTES,
                 smem_u32(&sm.full[slot]));
      }
    }
    umma_commit(smem_u32(&sm.accum));                  // all three accumulators complete
  }
  __syncwarp();

  // ---- epilogue: TMEM lane = xdim column, TMEM column = chain -----------------------------
  // 8 warps: warp w reads TMEM lanes 32 (w % 4) .. +31 (the hardware's lane quarter of a warp)
  // and owns the chains [32 (w / 4), +32) of the tile, 8 chains at a time.
  {
    const bool full = (chain0 + BN <= a.nb) && ((size_t)(tile + 1) * BM <= (size_t)a.xdim) && a.stq == nullptr;
    if (full) {                                         // (waits for the accumulators inside, after its first loads)
      if (a.sign > 0) epilogue_full<true>(a, sm, tmem, tile, chain0, warp, lane);
      else epilogue_full<false>(a, sm, tmem, tile, chain0, warp, lane);
    } else {
      mbar_wait<128>(smem_u32(&sm.accum), 0);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (a.sign > 0) epilogue<true, false>(a, sm, tmem, tile, chain0, warp, lane);
      else epilogue<false, false>(a, sm, tmem, tile, chain0, warp, lane);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (a.part != nullptr && tid < BN && chain0 + tid < a.nb) {
    const int h = tid >> 5, c = tid & 31;
    const double x = (double)sm.ld[4 * h + 0][c] + (double)sm.ld[4 * h + 1][c] + (double)sm.ld[4 * h + 2][c] +
                     (double)sm.ld[4 * h + 3][c];
    a.part[(size_t)(chain0 + tid) * a.ntiles + tile] = x;
  }
  if (warp == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(TMEM_COLS) : "memory");
  }
}

// ---------------------------------------------------------------------------
// Adjoint of the fused heads + momentum update, element-wise part.  Given (s, t, q) as the forward kernel
// dumped them and the cotangents of (v', logdet), one pass produces gv, gF, the per-chain d/d eps partials,
// the cotangents of the three heads' PRE-activations (ready to be the A operand of the dz / dW GEMMs, in
// the GEMM's dtype) and the products gs*s, gq*q whose column sums are the ScaledTanh.coeff gradients:
//   s = a_s tanh(pre_s):  d/dpre_s = gs a_s (1 - (s/a_s)^2),  d/dc_s = gs s      (a_s = nw.s e^{c_s}; same for q)
//   t = nw.t pre_t:       d/dpre_t = gt nw.t
// Replaces k_vupdate_bwd + three fp32->fp64 conversions + ~20 element-wise torch kernels per v-update.
// ---------------------------------------------------------------------------
template <typename GP> __device__ __forceinline__ GP to_gp(float x);
template <> __device__ __forceinline__ float to_gp<float>(float x) { return x; }
template <> __device__ __forceinline__ __nv_bfloat16 to_gp<__nv_bfloat16>(float x) { return __float2bfloat16(x); }

template <typename GP>
__global__ void __launch_bounds__(256) k_heads_vupdate_bwd(const double2* __restrict__ v, const double2* __restrict__ f,
                                                           const float* __restrict__ stq,
                                                           const float* __restrict__ scale_s,
                                                           const float* __restrict__ scale_q, float scale_t,
                                                           double eps_in, const double* __restrict__ eps_dev, int sign,
                                                           const double2* __restrict__ gout,
                                                           const double* __restrict__ glogdet,
                                                           double2* __restrict__ gv, double2* __restrict__ gf,
                                                           GP* __restrict__ gpre, float* __restrict__ colsum,
                                                           double* __restrict__ part, int nb, int xdim) {
  // One thread owns a column j and walks the chains: the five sums over the chains that the parameter gradients need
  // -- the three heads' bias gradients sum_b gpre_h[b][j] and sum_b gs s, sum_b gq q for the ScaledTanh coefficients --
  // stay in registers and are written once (colsum [5][xdim]); the first version wrote gs s / gq q as [nb, xdim]
  // arrays and left five reductions per update to torch.  Chains in a fixed order: deterministic.
  __shared__ double red[8];
  const double eps = eps_dev ? eps_in * eps_dev[0] : eps_in;
  const int j = blockIdx.x * 256 + threadIdx.x;
  const bool live = j < xdim;
  const size_t plane = (size_t)nb * xdim;
  const double as = live ? (double)scale_s[j] : 0.0, aq = live ? (double)scale_q[j] : 0.0;
  const double sg = (double)sign, he = 0.5 * eps;
  float cs0 = 0.f, cs1 = 0.f, cs2 = 0.f, cs3 = 0.f, cs4 = 0.f;
  for (int b = 0; b < nb; ++b) {
    double ge = 0.0;
    if (live) {
      const size_t at = (size_t)b * xdim + j;
      const double sv = (double)stq[at], tv = (double)stq[plane + at], qv = (double)stq[2 * plane + at];
      const double2 V = v[at], F = f[at], G = gout[at];
      const double gl = glogdet ? glogdet[b] : 0.0;
      const double lj = sg * eps * sv / 2.0;
      const double es = exp(lj), eq = exp(eps * qv);
      const double fr = fma(F.x, eq, tv), fi = F.y * eq;
      double g_es, g_fr, g_fi;
      if (sign > 0) {
        g_es = G.x * V.x + G.y * V.y;
        g_fr = -he * G.x; g_fi = -he * G.y;
        ge += -0.5 * (fr * G.x + fi * G.y);
      } else {
        g_es = G.x * (V.x + he * fr) + G.y * (V.y + he * fi);
        g_fr = es * he * G.x; g_fi = es * he * G.y;
        ge += 0.5 * es * (fr * G.x + fi * G.y);
      }
      gv[at] = make_double2(es * G.x, es * G.y);
      if (gf != nullptr) gf[at] = make_double2(g_fr * eq, g_fi * eq);
      const double g_lj = g_es * es + gl;
      ge += g_lj * sg * sv / 2.0;
      const double g_eq = g_fr * F.x + g_fi * F.y;
      ge += g_eq * eq * qv;
      const double gs = g_lj * sg * eps / 2.0, gt = g_fr, gq = g_eq * eq * eps;
      const double ths = as != 0.0 ? sv / as : 0.0, thq = aq != 0.0 ? qv / aq : 0.0;
      const float p0 = (float)(gs * as * (1.0 - ths * ths)), p1 = (float)(gt * (double)scale_t),
                  p2 = (float)(gq * aq * (1.0 - thq * thq));
      gpre[at] = to_gp<GP>(p0);
      gpre[plane + at] = to_gp<GP>(p1);
      gpre[2 * plane + at] = to_gp<GP>(p2);
      cs0 += p0; cs1 += p1; cs2 += p2;
      cs3 += (float)(gs * sv);
      cs4 += (float)(gq * qv);
    }
    ge = block_sum<256>(ge, red, threadIdx.x);
    if (threadIdx.x == 0) part[(size_t)b * gridDim.x + blockIdx.x] = ge;
  }
  if (live) {
    const size_t xd = (size_t)xdim;
    colsum[j] = cs0; colsum[xd + j] = cs1; colsum[2 * xd + j] = cs2; colsum[3 * xd + j] = cs3; colsum[4 * xd + j] = cs4;
  }
}

// fixed-order sum of the per-tile partials: out[b] = sum_tile part[b][tile]
__global__ void __launch_bounds__(256) k_sum_rows(const double* __restrict__ part, int n, double* __restrict__ out) {
  __shared__ double red[8];
  const double* row = part + (size_t)blockIdx.x * n;
  double s = 0.0;
  for (int k = threadIdx.x; k < n; k += 256) s += row[k];
  s = block_sum<256>(s, red, threadIdx.x);
  if (threadIdx.x == 0) out[blockIdx.x] = s;
}

// ---------------------------------------------------------------------------
// SU(3) vnet INPUT layer on the tensor cores (reference network/pytorch/network.py:349-451 `InputLayer`, its two
// Linears at :415-422, called from dynamics.py:1142-1160):
//     z[b, h] = act( sum_k ax[b, k] Wx[h, k] + bx[h]  +  sum_k af[b, k] Wv[h, k] + bv[h] ),   k < 8 * 4V
// ax = su3_to_vec(projectSU(x)), af = su3_to_vec(projectSU(F)) -- 8 reals per link.  A GEMM with tiny M (H <= 256) and
// N (chains <= 256) and an enormous K (8^4: 131 072, twice): pure operand streaming, 268 MB for 34 GF at nb = H = 256.
//
// k_su3_input_gemm: split-K over one CTA per SM.  The concatenated K axis [x part | F part] is cut into chunks of 64
// (= 8 links); a CTA owns a contiguous run of chunks and accumulates D[h, b] (UMMA M = 128 rows of h per accumulator,
// N = the padded chain count) in TMEM over all of them.  Both operands are stored in HBM as the canonical no-swizzle
// K-major core-matrix image of their chunks, so a pipeline stage is two contiguous 1-D bulk copies (UBLKCP):
//   * weights: packed once per weight version by k_pack_input, [part][chunk][kcore 0..7][row h < HP][8 bf16];
//   * activations: written in exactly this form by the producer itself, k_project_vec_planar_lm -- one link IS one
//     K core (8 reals), so the image is simply "link-major": [link][chain < NBP][8 bf16].  vec8 never exists in the
//     [chain][link][8] layout a library GEMM would need.
// One thread issues the copies and the tcgen05.mma's (cta_group::1, kind::f16, M128 x N(NBP) x K16); 4 warps drain
// the accumulators with tcgen05.ld into fp32 partials [cta][h][b].
// k_su3_input_reduce: fixed-order sum of the partials over the CTAs (deterministic), both biases, the activation,
// cast -> z[b][h] (bf16, the B operand of k_heads_vupdate).
// ---------------------------------------------------------------------------
constexpr int IL_KC = 64;             // K per chunk = 8 links
constexpr int IL_NST = 3;             // pipeline stages
constexpr int IL_NTH = 128;           // 4 warps: TMEM lane quarters 0..3

struct InputArgs {
  const unsigned char* wpk;           // k_pack_input image
  const unsigned char* act[2];        // link-major activation images of the x part and the F part
  float* part;                        // [ncta][HP][NBP] fp32
  int nch;                            // chunks per part (= nlinks / 8)
  int HP;                             // padded rows of the weight image: 128 or 256
  int NBP;                            // padded chains: multiple of 16, <= 256
  uint32_t tmem_cols;                 // power of two >= (HP / 128) * NBP, >= 32
};

__global__ void __launch_bounds__(IL_NTH, 1) k_su3_input_gemm(const InputArgs a) {
  extern __shared__ __align__(128) unsigned char il_smem[];
  __shared__ __align__(8) unsigned long long full[IL_NST], empty[IL_NST], accum;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t w_bytes = (uint32_t)a.HP * 128u, act_bytes = (uint32_t)a.NBP * 128u, stage_bytes = w_bytes + act_bytes;
  const long long total = 2ll * a.nch;
  const long long c0 = (long long)blockIdx.x * total / gridDim.x, c1 = (long long)(blockIdx.x + 1) * total / gridDim.x;
  const int n = (int)(c1 - c0);
  const int n_mt = a.HP / 128;

  auto issue = [&](int s) {            // stage s of this CTA's run -> ring slot s % IL_NST
    const long long c = c0 + s;
    const int part = c >= a.nch ? 1 : 0;
    const long long ci = c - (long long)part * a.nch;
    const uint32_t slot = (uint32_t)(s % IL_NST);
    const uint32_t dst = smem_u32(il_smem) + slot * stage_bytes;
    mbar_expect_tx(smem_u32(&full[slot]), stage_bytes);
    bulk_g2s(dst, a.wpk + ((size_t)part * a.nch + (size_t)ci) * w_bytes, w_bytes, smem_u32(&full[slot]));
    bulk_g2s(dst + w_bytes, a.act[part] + (size_t)ci * act_bytes, act_bytes, smem_u32(&full[slot]));
  };

  if (tid == 0) {
    for (int s = 0; s < IL_NST; ++s) { mbar_init(smem_u32(&full[s]), 1); mbar_init(smem_u32(&empty[s]), 1); }
    mbar_init(smem_u32(&accum), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    for (int s = 0; s < IL_NST && s < n; ++s) issue(s);      // the first ring fill overlaps the TMEM allocation
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)),
                 "r"(a.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base_s;

  if (tid == 0) {
    // instruction descriptor, kind::f16: D = f32, A = B = bf16, both K-major, M = 128, N = NBP
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(a.NBP >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t w_lbo = (uint32_t)a.HP * 16u, b_lbo = (uint32_t)a.NBP * 16u, sbo = 128u;
    for (int s = 0; s < n; ++s) {
      const uint32_t slot = (uint32_t)(s % IL_NST), ph = (uint32_t)(s / IL_NST) & 1u;
      mbar_wait(smem_u32(&full[slot]), ph);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t w_base = smem_u32(il_smem) + slot * stage_bytes, b_base = w_base + w_bytes;
      for (int mt = 0; mt < n_mt; ++mt) {
#pragma unroll
        for (int kk = 0; kk < IL_KC / 16; ++kk) {
          const uint64_t ad = umma_desc(w_base + (2 * kk) * w_lbo + mt * (128 * 16), w_lbo, sbo);
          const uint64_t bd = umma_desc(b_base + (2 * kk) * b_lbo, b_lbo, sbo);
          umma_f16(tmem + mt * a.NBP, ad, bd, idesc, (s | kk) != 0 ? 1u : 0u);
        }
      }
      umma_commit(smem_u32(&empty[slot]));
      if (s + IL_NST < n) {
        mbar_wait(smem_u32(&empty[slot]), ph);
        issue(s + IL_NST);
      }
    }
    umma_commit(smem_u32(&accum));
  }
  __syncwarp();
  // ---- drain: TMEM lane = h (row of the weight tile), column = chain ----------------------------------------
  mbar_wait<128>(smem_u32(&accum), 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  for (int mt = 0; mt < n_mt; ++mt) {
    float* row = a.part + ((size_t)blockIdx.x * a.HP + mt * 128 + warp * 32 + lane) * a.NBP;
    const uint32_t trow = tmem + ((uint32_t)(warp * 32) << 16) + mt * a.NBP;
    for (int c = 0; c < a.NBP; c += 8) {
      uint32_t r[8];
      tmem_ld8(trow + c, r);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      *reinterpret_cast<uint4*>(row + c) = make_uint4(r[0], r[1], r[2], r[3]);
      *reinterpret_cast<uint4*>(row + c + 4) = make_uint4(r[4], r[5], r[6], r[7]);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(a.tmem_cols) : "memory");
  }
}

// z[b][h] = act( sum_cta part[cta][h][b] + bx[h] + bv[h] ).  Block = one row h x 32 chains x 8 "z lanes": lane zl sums the
// partials of CTAs zl, zl + 8, ... (coalesced 128-byte reads, 8 x shorter dependent chains, 2048 blocks instead of 64),
// then the eight lane sums are added in lane order through shared memory: a fixed summation order, deterministic.
// (The first version walked all partials of four outputs per thread on 64 blocks: 87 us for 39 MB.)
__global__ void __launch_bounds__(256) k_su3_input_reduce(const float* __restrict__ part, int ncta, int HP, int NBP,
                                                          const float* __restrict__ bx, const float* __restrict__ bv,
                                                          int act, int H, int nb, __nv_bfloat16* __restrict__ z) {
  __shared__ float red[8][32];
  const int tx = threadIdx.x & 31, zl = threadIdx.x >> 5;
  const int b = blockIdx.x * 32 + tx, h = blockIdx.y;
  float s = 0.f;
  if (b < NBP) {
    const float* p = part + (size_t)h * NBP + b;
    const size_t plane = (size_t)HP * NBP;
#pragma unroll 4
    for (int c = zl; c < ncta; c += 8) s += __ldg(p + (size_t)c * plane);
  }
  red[zl][tx] = s;
  __syncthreads();
  if (zl != 0 || b >= nb) return;
#pragma unroll
  for (int k = 1; k < 8; ++k) s += red[k][tx];
  z[(size_t)b * H + h] = __float2bfloat16(il_act(s + bx[h] + bv[h], act));
}

// W_x, W_v [H, K] (nn.Linear layout; f64 / f32 / bf16), K = 8 * nlinks -> [part][chunk][kcore][row < HP][8] bf16
template <typename W>
__global__ void __launch_bounds__(256) k_pack_input(const W* __restrict__ wx, const W* __restrict__ wv,
                                                    uint4* __restrict__ packed, int H, int HP, int nch, size_t n16) {
  const size_t id = (size_t)blockIdx.x * 256 + threadIdx.x;
  if (id >= n16) return;
  const int r = (int)(id % HP);
  size_t rest = id / HP;
  const int kcore = (int)(rest % 8);
  rest /= 8;
  const size_t chunk = rest % nch;
  const int part = (int)(rest / nch);
  const W* w = part == 0 ? wx : wv;
  const size_t K = (size_t)nch * IL_KC;
  __align__(16) __nv_bfloat16 h[8];
#pragma unroll
  for (int e = 0; e < 8; ++e)
    h[e] = __float2bfloat16(r < H ? w_load(w + (size_t)r * K + chunk * IL_KC + kcore * 8 + e) : 0.0f);
  packed[id] = *reinterpret_cast<const uint4*>(h);
}

}  // namespace
}  // namespace l2b

using namespace l2b;

extern "C" {

size_t l2b_su3_input_packed_bytes(int nlinks, int hidden) {
  if (nlinks <= 0 || nlinks % 8 != 0 || hidden <= 0 || hidden > 256) return 0;
  const size_t HP = hidden <= 128 ? 128 : 256;
  return (size_t)2 * (nlinks / 8) * 8 * HP * 16;
}

size_t l2b_su3_input_ws_bytes(int nb_pad, int hidden) {
  if (nb_pad <= 0 || hidden <= 0 || hidden > 256) return 0;
  int dev = 0, nsm = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess)
    return 0;
  const size_t HP = hidden <= 128 ? 128 : 256;
  return align_up((size_t)nsm * HP * nb_pad * sizeof(float), 256);
}

int l2b_su3_input_pack(const void* w_x, const void* w_v, int w_dtype, void* packed, int nlinks, int hidden,
                       void* stream) {
  L2B_REQUIRE(w_x && w_v && packed, L2B_ERR_INVALID, "null pointer");
  L2B_REQUIRE(nlinks > 0 && nlinks % 8 == 0, L2B_ERR_UNSUPPORTED, "nlinks must be a positive multiple of 8 (got %d)", nlinks);
  L2B_REQUIRE(hidden > 0 && hidden <= 256, L2B_ERR_UNSUPPORTED, "hidden must be in [1, 256] (got %d)", hidden);
  L2B_REQUIRE(((uintptr_t)packed & 15) == 0, L2B_ERR_INVALID, "packed image must be 16-byte aligned");
  const int HP = hidden <= 128 ? 128 : 256, nch = nlinks / 8;
  const size_t n16 = l2b_su3_input_packed_bytes(nlinks, hidden) / 16;
  const unsigned nblk = (unsigned)((n16 + 255) / 256);
  cudaStream_t st = (cudaStream_t)stream;
  if (w_dtype == L2B_F32)
    k_pack_input<float><<<nblk, 256, 0, st>>>((const float*)w_x, (const float*)w_v, (uint4*)packed, hidden, HP, nch, n16);
  else if (w_dtype == L2B_F64)
    k_pack_input<double><<<nblk, 256, 0, st>>>((const double*)w_x, (const double*)w_v, (uint4*)packed, hidden, HP, nch, n16);
  else if (w_dtype == L2B_BF16)
    k_pack_input<__nv_bfloat16><<<nblk, 256, 0, st>>>((const __nv_bfloat16*)w_x, (const __nv_bfloat16*)w_v, (uint4*)packed, hidden, HP, nch, n16);
  else
    L2B_REQUIRE(false, L2B_ERR_UNSUPPORTED, "w_dtype must be L2B_F32, L2B_F64 or L2B_BF16");
  L2B_LAUNCHED("k_pack_input");
  return L2B_OK;
}

int l2b_su3_input_layer(const void* act_x, const void* act_f, const void* packed, const float* bias_x,
                        const float* bias_v, int activation, void* z_bf16, int nb, int nb_pad, int nlinks, int hidden,
                        void* ws, size_t ws_bytes, void* stream) {
  L2B_REQUIRE(act_x && act_f && packed && bias_x && bias_v && z_bf16, L2B_ERR_INVALID, "null pointer");
  L2B_REQUIRE(nb > 0 && nb_pad >= nb && nb_pad % 16 == 0 && nb_pad <= 256, L2B_ERR_UNSUPPORTED,
              "nb_pad must be a multiple of 16 in [nb, 256] (nb=%d nb_pad=%d)", nb, nb_pad);
  L2B_REQUIRE(nlinks > 0 && nlinks % 8 == 0, L2B_ERR_UNSUPPORTED, "nlinks must be a positive multiple of 8 (got %d)", nlinks);
  L2B_REQUIRE(hidden > 0 && hidden <= 256, L2B_ERR_UNSUPPORTED, "hidden must be in [1, 256] (got %d)", hidden);
  L2B_REQUIRE(activation >= 0 && activation <= 5, L2B_ERR_INVALID, "activation code must be in [0, 5]");
  L2B_REQUIRE((((uintptr_t)act_x | (uintptr_t)act_f | (uintptr_t)packed | (uintptr_t)ws) & 15) == 0, L2B_ERR_INVALID,
              "act_x, act_f, packed, ws must be 16-byte aligned");
  L2B_REQUIRE(ws != nullptr && ws_bytes >= l2b_su3_input_ws_bytes(nb_pad, hidden), L2B_ERR_WORKSPACE,
              "workspace too small for the split-K partials");
  int dev = 0, nsm = 0;
  L2B_CUDA(cudaGetDevice(&dev));
  L2B_CUDA(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev));
  InputArgs a;
  a.wpk = (const unsigned char*)packed;
  a.act[0] = (const unsigned char*)act_x;
  a.act[1] = (const unsigned char*)act_f;
  a.part = (float*)ws;
  a.nch = nlinks / 8;
  a.HP = hidden <= 128 ? 128 : 256;
  a.NBP = nb_pad;
  uint32_t cols = 32;
  while (cols < (uint32_t)((a.HP / 128) * a.NBP)) cols <<= 1;
  a.tmem_cols = cols;
  const long long total = 2ll * a.nch;
  const int ncta = (int)(total < nsm ? total : nsm);
  const size_t smem = (size_t)IL_NST * ((size_t)a.HP + a.NBP) * 128;
  cudaStream_t st = (cudaStream_t)stream;
  L2B_CUDA(cudaFuncSetAttribute((const void*)k_su3_input_gemm, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k_su3_input_gemm<<<ncta, IL_NTH, smem, st>>>(a);
  L2B_LAUNCHED("k_su3_input_gemm");
  const dim3 rgrid((nb_pad + 31) / 32, hidden);
  k_su3_input_reduce<<<rgrid, 256, 0, st>>>(a.part, ncta, a.HP, a.NBP, bias_x, bias_v, activation, hidden, nb,
                                             (__nv_bfloat16*)z_bf16);
  L2B_LAUNCHED("k_su3_input_reduce");
  return L2B_OK;
}

size_t l2b_vnet_heads_packed_bytes(int xdim, int hidden) {
  if (xdim <= 0 || hidden <= 0) return 0;
  const size_t ntiles = ((size_t)xdim + BM - 1) / BM;
  const size_t KP = ((size_t)hidden + KC - 1) / KC * KC;
  return ntiles * 3 * KP * BM * 2;
}

size_t l2b_vnet_heads_ws_bytes(int nb, int xdim) {
  if (xdim <= 0 || nb <= 0) return 0;
  const size_t ntiles = ((size_t)xdim + BM - 1) / BM;
  return align_up((size_t)nb * ntiles * sizeof(double), 256);
}

int l2b_vnet_pack_heads(const void* w_s, const void* w_t, const void* w_q, int w_dtype, void* packed, int xdim,
                        int hidden, void* stream) {
  L2B_REQUIRE(w_s && w_t && w_q && packed, L2B_ERR_INVALID, "null pointer");
  L2B_REQUIRE(xdim > 0 && hidden > 0, L2B_ERR_INVALID, "xdim and hidden must be positive");
  L2B_REQUIRE(((uintptr_t)packed & 15) == 0, L2B_ERR_INVALID, "packed image must be 16-byte aligned");
  const int KP = (hidden + KC - 1) / KC * KC;
  const size_t nchunks = l2b_vnet_heads_packed_bytes(xdim, hidden) / 16;
  const unsigned nblk = (unsigned)((nchunks + 255) / 256);
  cudaStream_t st = (cudaStream_t)stream;
  if (w_dtype == L2B_F32)
    k_pack_heads<float><<<nblk, 256, 0, st>>>((const float*)w_s, (const float*)w_t, (const float*)w_q, (uint4*)packed, xdim, hidden, KP, nchunks);
  else if (w_dtype == L2B_F64)
    k_pack_heads<double><<<nblk, 256, 0, st>>>((const double*)w_s, (const double*)w_t, (const double*)w_q, (uint4*)packed, xdim, hidden, KP, nchunks);
  else if (w_dtype == L2B_BF16)
    k_pack_heads<__nv_bfloat16><<<nblk, 256, 0, st>>>((const __nv_bfloat16*)w_s, (const __nv_bfloat16*)w_t, (const __nv_bfloat16*)w_q, (uint4*)packed, xdim, hidden, KP, nchunks);
  else
    L2B_REQUIRE(false, L2B_ERR_UNSUPPORTED, "w_dtype must be L2B_F32, L2B_F64 or L2B_BF16");
  L2B_LAUNCHED("k_pack_heads");
  return L2B_OK;
}

static int heads_vupdate_impl(const void* z, const void* packed, const float* bias_s, const float* bias_t,
                              const float* bias_q, const float* scale_s, const float* scale_q, float scale_t,
                              const void* v, const void* force, double eps, const double* eps_dev, int sign,
                              double eps2, const double* eps2_dev, int sign2, int negate, void* v_out, double* logdet,
                              float* stq_or_null, int nb, int xdim, int hidden, void* ws, size_t ws_bytes,
                              void* stream) {
  L2B_REQUIRE(z && packed && bias_s && bias_t && bias_q && scale_s && scale_q && v && force && v_out,
              L2B_ERR_INVALID, "null pointer");
  L2B_REQUIRE(nb > 0 && xdim > 0, L2B_ERR_INVALID, "nb and xdim must be positive");
  L2B_REQUIRE(sign == 1 || sign == -1, L2B_ERR_INVALID, "sign must be +1 or -1");
  L2B_REQUIRE(sign2 == 0 || sign2 == 1 || sign2 == -1, L2B_ERR_INVALID, "sign2 must be 0, +1 or -1");
  L2B_REQUIRE(sign2 == 0 || stq_or_null == nullptr, L2B_ERR_UNSUPPORTED, "the paired update has no (s, t, q) dump");
  L2B_REQUIRE(hidden > 0 && hidden % 8 == 0 && hidden <= KMAX, L2B_ERR_UNSUPPORTED,
              "fused heads kernel needs hidden %% 8 == 0 and hidden <= %d (got %d)", KMAX, hidden);
  L2B_REQUIRE((((uintptr_t)z | (uintptr_t)packed | (uintptr_t)v | (uintptr_t)force | (uintptr_t)v_out) & 15) == 0,
              L2B_ERR_INVALID, "z, packed, v, force, v_out must be 16-byte aligned");
  HeadsArgs a;
  a.z = (const __nv_bfloat16*)z;
  a.packed = (const unsigned char*)packed;
  a.bias[0] = bias_s; a.bias[1] = bias_t; a.bias[2] = bias_q;
  a.scale_s = scale_s; a.scale_q = scale_q; a.scale_t = scale_t;
  a.v = (const double2*)v; a.f = (const double2*)force; a.out = (double2*)v_out;
  a.stq = stq_or_null;
  a.eps = eps; a.eps_dev = eps_dev; a.sign = sign; a.nb = nb; a.xdim = xdim; a.H = hidden;
  a.eps2 = eps2; a.eps2_dev = eps2_dev; a.sign2 = sign2; a.negate = negate;
  a.KP = (hidden + KC - 1) / KC * KC;
  a.ntiles = (xdim + BM - 1) / BM;
  a.part = nullptr;
  if (logdet) {
    L2B_REQUIRE(ws != nullptr && ws_bytes >= l2b_vnet_heads_ws_bytes(nb, xdim), L2B_ERR_WORKSPACE,
                "workspace too small for the logdet partials");
    a.part = (double*)ws;
  }
  const int nct = (nb + BN - 1) / BN;
  const size_t nblk = (size_t)a.ntiles * nct;
  L2B_REQUIRE(nblk < (1ull << 31), L2B_ERR_UNSUPPORTED, "grid too large");
  cudaStream_t st = (cudaStream_t)stream;
  L2B_CUDA(cudaFuncSetAttribute((const void*)k_heads_vupdate, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)sizeof(Smem)));
  k_heads_vupdate<<<(unsigned)nblk, NTH, sizeof(Smem), st>>>(a);
  L2B_LAUNCHED("k_heads_vupdate");
  if (logdet) {
    k_sum_rows<<<nb, 256, 0, st>>>(a.part, a.ntiles, logdet);
    L2B_LAUNCHED("k_sum_rows");
  }
  return L2B_OK;
}

int l2b_su3_heads_vupdate(const void* z, const void* packed, const float* bias_s, const float* bias_t,
                          const float* bias_q, const float* scale_s, const float* scale_q, float scale_t,
                          const void* v, const void* force, double eps, const double* eps_dev, int sign, void* v_out,
                          double* logdet,
                          float* stq_or_null, int nb, int xdim, int hidden, void* ws, size_t ws_bytes, void* stream) {
  return heads_vupdate_impl(z, packed, bias_s, bias_t, bias_q, scale_s, scale_q, scale_t, v, force, eps, eps_dev, sign,
                            0.0, nullptr, 0, 0, v_out, logdet, stq_or_null, nb, xdim, hidden, ws, ws_bytes, stream);
}

int l2b_su3_heads_vupdate_pair(const void* z, const void* packed, const float* bias_s, const float* bias_t,
                               const float* bias_q, const float* scale_s, const float* scale_q, float scale_t,
                               const void* v, const void* force, double eps1, const double* eps1_dev, int sign1,
                               double eps2, const double* eps2_dev, int sign2, int negate_between, void* v_out,
                               double* logdet, int nb, int xdim, int hidden, void* ws, size_t ws_bytes, void* stream) {
  L2B_REQUIRE(sign2 == 1 || sign2 == -1, L2B_ERR_INVALID, "sign2 must be +1 or -1");
  return heads_vupdate_impl(z, packed, bias_s, bias_t, bias_q, scale_s, scale_q, scale_t, v, force, eps1, eps1_dev,
                            sign1, eps2, eps2_dev, sign2, negate_between ? 1 : 0, v_out, logdet, nullptr, nb, xdim,
                            hidden, ws, ws_bytes, stream);
}

int l2b_su3_heads_vupdate_bwd(const void* v, const void* force, const float* stq, const float* scale_s,
                              const float* scale_q, float scale_t, double eps, const double* eps_dev, int sign,
                              const void* gv_out, const double* glogdet, void* gv, void* gforce_or_null, void* gpre,
                              int gpre_dtype, float* colsum, double* geps, int nb, int xdim, void* ws,
                              size_t ws_bytes, void* stream) {
  L2B_REQUIRE(v && force && stq && scale_s && scale_q && gv_out && gv && gpre && colsum && geps, L2B_ERR_INVALID,
              "null pointer");
  L2B_REQUIRE(nb > 0 && xdim > 0 && nb <= 65535, L2B_ERR_INVALID, "nb must be in [1, 65535], xdim positive");
  L2B_REQUIRE(sign == 1 || sign == -1, L2B_ERR_INVALID, "sign must be +1 or -1");
  const int nblk = (xdim + 255) / 256;
  L2B_REQUIRE(ws && ws_bytes >= (size_t)nb * nblk * sizeof(double), L2B_ERR_WORKSPACE, "workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  const dim3 grid(nblk, 1);
  double* part = (double*)ws;
#define L2B_HVB(GP)                                                                                               \
  k_heads_vupdate_bwd<GP><<<grid, 256, 0, st>>>((const double2*)v, (const double2*)force, stq, scale_s, scale_q,    \
                                                scale_t, eps, eps_dev, sign, (const double2*)gv_out, glogdet,      \
                                                (double2*)gv, (double2*)gforce_or_null, (GP*)gpre, colsum, part,  \
                                                nb, xdim)
  if (gpre_dtype == L2B_F32) L2B_HVB(float);
  else if (gpre_dtype == L2B_BF16) L2B_HVB(__nv_bfloat16);
  else L2B_REQUIRE(false, L2B_ERR_UNSUPPORTED, "gpre_dtype must be L2B_F32 or L2B_BF16");
#undef L2B_HVB
  L2B_LAUNCHED("k_heads_vupdate_bwd");
  k_sum_rows<<<nb, 256, 0, st>>>(part, nblk, geps);
  L2B_LAUNCHED("k_sum_rows");
  return L2B_OK;
}

}  // extern "C"