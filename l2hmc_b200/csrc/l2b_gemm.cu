// l2b_gemm.cu -- the dense layers of the L2HMC networks as ONE hand-written bf16 GEMM on the 5th-generation tensor
// cores (sm_100a: TMA -> shared memory -> tcgen05.mma -> TMEM), for every Linear the fused kernels of l2b_vnet.cu do
// not already cover: the hidden Linears (reference network/pytorch/network.py:489-493, 538-541), the input Linears
// under autograd (:415-422), and the three GEMMs of every Linear's backward pass (dX = dY W, dW = dY^T X) that the
// reference gets from ATen autograd / cuBLAS.
//
//     D[m][n] = sum_seg sum_k A_seg(m, k) B_seg(n, k)          (fp32 accumulation in TMEM)
//
// Each operand is an ordinary ROW-MAJOR bf16 matrix and may be contracted over either of its axes:
//     K-major  : stored [MN][K]  (the contraction index is contiguous)          -- x and W of a Linear's forward
//     MN-major : stored [K][MN]  (the contraction index is the row)             -- W in dX = dY W, dY and X in dW
// so no transposed copy of an activation, a weight or a cotangent is ever made.  Up to 32 (A, B) pairs of the same
// shape are summed in one launch ("segments": the two input Linears, the three heads of dz, the 2 N_LF x 2 momentum
// updates whose dW contributions share one weight matrix).
//
// Data path.  Every segment of every operand gets a rank-2 TMA descriptor (cuTensorMapEncodeTiled, 128-byte swizzle,
// zero fill outside the matrix -- ragged M, N, K need no special case) passed as a __grid_constant__ parameter.  One
// producer thread issues cp.async.bulk.tensor.2d per stage: a K-major tile is one box [rows x 64 k] (128-byte rows),
// an MN-major tile is one box [64 k x 64 mn] per 64 columns; both land in exactly the swizzled layouts the UMMA
// shared-memory descriptors name (K-major: SBO = 1024; MN-major: LBO = 8192 between 64-column blocks, SBO = 1024
// between 8-row groups; the instruction descriptor's a_major / b_major bits select the orientation).  One MMA thread
// waits on the stage's mbarrier, issues four tcgen05.mma (M128 x N(BN) x K16, kind::f16) and recycles the stage with
// tcgen05.commit.  Four epilogue warps drain TMEM (tcgen05.ld 32x32b: lane = row m, columns = n): bias, activation,
// optional accumulate, bf16 / fp32 store -- or fp32 split-K partials that k_gemm_reduce finishes in a fixed order
// (deterministic).  Tiles with a short K loop run two CTAs per SM so that one CTA's epilogue overlaps the other's
// loads.  (First version, measured: 16-byte cp.async.cg loaders into the no-swizzle layout stall on the LSU at
// ~14 GB/s per SM -- 2 TB/s in all, `profiles/r2j_gemm_*`; the TMA path replaced it.)
#include <cuda.h>
#include <cuda_bf16.h>

#include "l2b_common.cuh"
#include "l2b_tc.cuh"

namespace l2b {
namespace {

constexpr int G_BM = 128;            // UMMA M
constexpr int G_BK = 64;             // K per pipeline stage: one 128-byte swizzle span
constexpr int G_NST = 4;             // ring slots at most (GemmArgs::nst of them are used)
constexpr int G_NTH = 192;           // 4 epilogue warps + TMA producer warp + MMA warp
constexpr int G_MAXSEG = 32;         // (A, B) pairs summed in one launch
constexpr uint32_t G_BOX = 8192;     // one MN-major box: 64 k-rows x 128 bytes

template <int MAXSEG>
struct GemmMaps {
  CUtensorMap a[MAXSEG];
  CUtensorMap b[MAXSEG];
};

struct GemmArgs {
  int a_kmajor, b_kmajor;            // 1: stored [MN][K]; 0: stored [K][MN]
  int nseg, M, N, K;
  int BN;                            // N tile: multiple of 16 (64 when B is MN-major), <= 256
  int n_mt, n_nt;
  void* out;                         // [M][ldo] bf16 or fp32 (splits == 1)
  long long ldo;
  int out_f32, accumulate, act;
  const float* bias;                 // [N] or null
  float* part;                       // [splits][M][N] fp32 (splits > 1)
  uint32_t tmem_cols;
  int nst;                           // pipeline stages in use: 2 (two CTAs per SM) .. G_NST
  int seg_inner;                     // chunk order: 1 = segments innermost (all segments of K chunk 0, then chunk 1, ...)
  int kpack;                         // 2: both operands MN-major with K <= 32: two segments share one 64-row stage
  int trans;                         // the kernel computes D^T (operands swapped by the host): out[n][m] <- D'[m][n]
};

// the reference's activations (codes as il_act) on the SFU: ex2.approx / rcp.approx, absolute error ~2e-7 -- the
// results are rounded to bf16 (bf16 nets); fp32 outputs use the libm-accurate functions
template <int ACT, bool PRECISE>
__device__ __forceinline__ float act_fast(float x) {
  if (PRECISE) return il_act(x, ACT);          // fp32 output (fp32 nets): libm-accurate tanhf / expf / expm1f
  if (ACT == 1) return tanh_fast(x);
  if (ACT == 2) return fmaxf(x, 0.f);
  if (ACT == 3) return x * rcp_approx(1.f + exp_fast(-x));
  if (ACT == 4) return x > 0.f ? x : 0.01f * x;
  if (ACT == 5) return x > 0.f ? x : exp_fast(x) - 1.f;
  return x;
}

__device__ __forceinline__ void tma2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
      ::"r"(dst), "l"((uint64_t)map), "r"(c0), "r"(c1), "r"(bar)
      : "memory");
}
// shared-memory matrix descriptor, 128-byte swizzle (layout type 2 in bits 61-63), descriptor version 1 (bit 46)
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) |
         (1ull << 46) | (2ull << 61);
}

template <int MAXSEG>
__global__ void __launch_bounds__(G_NTH, 2) k_gemm_bf16(const __grid_constant__ GemmMaps<MAXSEG> maps, const GemmArgs g) {
  extern __shared__ unsigned char gsm_raw[];
  __shared__ __align__(8) unsigned long long full[G_NST], empty[G_NST], accum;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int mt = blockIdx.x % g.n_mt, nt = blockIdx.x / g.n_mt;      // M tiles of one N tile are adjacent (share B in L2)
  const int m0 = mt * G_BM, n0 = nt * g.BN;
  const int nkc = (g.K + G_BK - 1) / G_BK;
  const long long total = g.kpack == 2 ? (g.nseg + 1) / 2 : (long long)g.nseg * nkc;
  const long long c0 = (long long)blockIdx.y * total / gridDim.y, c1 = (long long)(blockIdx.y + 1) * total / gridDim.y;
  const int n = (int)(c1 - c0);
  const uint32_t a_bytes = G_BM * G_BK * 2, b_bytes = (uint32_t)g.BN * G_BK * 2, stage_bytes = a_bytes + b_bytes;
  const uint32_t smem0 = (smem_u32(gsm_raw) + 1023u) & ~1023u;       // swizzled tiles: 1024-byte aligned

  if (tid == 0) {
    for (int s = 0; s < G_NST; ++s) { mbar_init(smem_u32(&full[s]), 1); mbar_init(smem_u32(&empty[s]), 1); }
    mbar_init(smem_u32(&accum), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 5) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)),
                 "r"(g.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base_s;

  if (tid == 128) {
    // ---- TMA producer ---------------------------------------------------------------------------------------------
    for (int s = 0; s < n; ++s) {
      const uint32_t slot = (uint32_t)(s % g.nst);
      if (s >= g.nst) mbar_wait(smem_u32(&empty[slot]), (uint32_t)(s / g.nst - 1) & 1u);
      const long long c = c0 + s;
      const uint32_t a_dst = smem0 + slot * stage_bytes, b_dst = a_dst + a_bytes, bar = smem_u32(&full[slot]);
      mbar_expect_tx(bar, stage_bytes);
      if (g.kpack == 2) {
        // rows 0..31 of the stage from segment 2c, rows 32..63 from segment 2c + 1 (boxes of 32 k-rows; a missing
        // last partner is requested outside the matrix: zero fill)
        for (int h = 0; h < 2; ++h) {
          const int sg = (int)(2 * c + h);
          const bool ok = sg < g.nseg;
          const int kk0 = ok ? 0 : g.K;
          const uint32_t off = (uint32_t)h * 4096u;               // 32 rows x 128 bytes into each 64-row block
          tma2d(a_dst + off, &maps.a[ok ? sg : 0], m0, kk0, bar);
          tma2d(a_dst + G_BOX + off, &maps.a[ok ? sg : 0], m0 + 64, kk0, bar);
          for (int j = 0; j < g.BN / 64; ++j) tma2d(b_dst + j * G_BOX + off, &maps.b[ok ? sg : 0], n0 + 64 * j, kk0, bar);
        }
        continue;
      }
      // segments outermost (streams each operand once, front to back) or innermost (the bf16x3 products of an fp32
      // GEMM re-use the same K chunk of a1, a2, a3 / b1, b2, b3 in consecutive stages: L2 hits)
      const int seg = g.seg_inner ? (int)(c % g.nseg) : (int)(c / nkc);
      const int k0 = (g.seg_inner ? (int)(c / g.nseg) : (int)(c - (long long)seg * nkc)) * G_BK;
      if (g.a_kmajor) tma2d(a_dst, &maps.a[seg], k0, m0, bar);
      else {
        tma2d(a_dst, &maps.a[seg], m0, k0, bar);
        tma2d(a_dst + G_BOX, &maps.a[seg], m0 + 64, k0, bar);
      }
      if (g.b_kmajor) tma2d(b_dst, &maps.b[seg], k0, n0, bar);
      else
        for (int j = 0; j < g.BN / 64; ++j) tma2d(b_dst + j * G_BOX, &maps.b[seg], n0 + 64 * j, k0, bar);
    }
  } else if (tid == 160) {
    // ---- MMA issuer -----------------------------------------------------------------------------------------------
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((g.a_kmajor ? 0u : 1u) << 15) |
                           ((g.b_kmajor ? 0u : 1u) << 16) | ((uint32_t)(g.BN >> 3) << 17) | ((uint32_t)(G_BM >> 4) << 24);
    // K-major, 128-byte swizzle: 8-row groups 1024 bytes apart (SBO); a K16 step moves 32 bytes inside the span.
    // MN-major: 64-column blocks G_BOX apart (LBO), 8-k-row groups 1024 bytes apart (SBO); a K16 step = 16 rows.
    const uint32_t a_lbo = g.a_kmajor ? 16u : G_BOX, b_lbo = g.b_kmajor ? 16u : G_BOX;
    const uint32_t a_step = g.a_kmajor ? 32u : 2048u, b_step = g.b_kmajor ? 32u : 2048u;
    for (int s = 0; s < n; ++s) {
      const uint32_t slot = (uint32_t)(s % g.nst), ph = (uint32_t)(s / g.nst) & 1u;
      mbar_wait(smem_u32(&full[slot]), ph);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t a_base = smem0 + slot * stage_bytes, b_base = a_base + a_bytes;
#pragma unroll
      for (int kk = 0; kk < G_BK / 16; ++kk) {
        const uint64_t ad = umma_desc_sw128(a_base + kk * a_step, a_lbo, 1024u);
        const uint64_t bd = umma_desc_sw128(b_base + kk * b_step, b_lbo, 1024u);
        umma_f16(tmem, ad, bd, idesc, (s | kk) != 0 ? 1u : 0u);
      }
      umma_commit(smem_u32(&empty[slot]));
    }
    umma_commit(smem_u32(&accum));
  }
  __syncwarp();

  // ---- epilogue: TMEM lane = row m of the tile, columns = n ------------------------------------------------------
  // One warp per lane quarter and nothing else to hide its latencies behind: the loop body is kept short (16 columns
  // per TMEM round trip, the bias / activation / accumulate work only when asked for).
  if (warp < 4 && m0 + warp * 32 < g.M) {
    if (n > 0) {
      mbar_wait<64>(smem_u32(&accum), 0);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
    const int m = m0 + warp * 32 + lane;
    const bool row_ok = m < g.M;
    const uint32_t trow = tmem + ((uint32_t)(warp * 32) << 16);
    const bool split = gridDim.y > 1;
    const bool plain = g.bias == nullptr && g.act == 0 && !g.accumulate;
    const float bias_t = (g.trans && g.bias != nullptr && row_ok) ? __ldg(g.bias + m) : 0.f;
    for (int c = 0; c < g.BN; c += 16) {
      const int nn = n0 + c;
      if (nn >= g.N) break;                                     // N % 8 == 0: an 8-column group is all in or all out
      const bool hi_ok = nn + 8 < g.N;
      uint32_t r[16];
      if (n > 0) {
        tmem_ld8(trow + c, r);
        tmem_ld8(trow + c + 8, r + 8);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      } else {
#pragma unroll
        for (int i = 0; i < 16; ++i) r[i] = 0u;
      }
      if (!row_ok) continue;
      if (split) {                                              // fp32 partials [z][M][N] of the kernel's own D
        float* p = g.part + ((size_t)blockIdx.y * g.M + m) * g.N + nn;
        *reinterpret_cast<uint4*>(p) = make_uint4(r[0], r[1], r[2], r[3]);
        *reinterpret_cast<uint4*>(p + 4) = make_uint4(r[4], r[5], r[6], r[7]);
        if (hi_ok) {
          *reinterpret_cast<uint4*>(p + 8) = make_uint4(r[8], r[9], r[10], r[11]);
          *reinterpret_cast<uint4*>(p + 12) = make_uint4(r[12], r[13], r[14], r[15]);
        }
        continue;
      }
      const int ncol = hi_ok ? 16 : 8;
      if (!plain) {
        float bv[16];
        if (g.bias != nullptr && !g.trans) {                     // all bias loads in flight before any of them is used
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float4 b4 = (q * 4 < ncol) ? __ldg(reinterpret_cast<const float4*>(g.bias + nn) + q)
                                             : make_float4(0.f, 0.f, 0.f, 0.f);
            bv[4 * q] = b4.x; bv[4 * q + 1] = b4.y; bv[4 * q + 2] = b4.z; bv[4 * q + 3] = b4.w;
          }
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i) bv[i] = bias_t;
        }
        // the activation is chosen once per 16 columns, not per element (a per-element switch compiles to an
        // indirect branch each: measured 3.4 k cycles per 16 columns against 0.4 k for the plain path)
#define L2B_ACT_LOOP(CODE)                                                                                      \
  if (g.out_f32) {                                                                                              \
    _Pragma("unroll") for (int i = 0; i < 16; ++i)                                                              \
        r[i] = __float_as_uint(act_fast<CODE, true>(__uint_as_float(r[i]) + bv[i]));                            \
  } else {                                                                                                      \
    _Pragma("unroll") for (int i = 0; i < 16; ++i)                                                              \
        r[i] = __float_as_uint(act_fast<CODE, false>(__uint_as_float(r[i]) + bv[i]));                           \
  }
        if (g.act == 1) { L2B_ACT_LOOP(1) }
        else if (g.act == 2) { L2B_ACT_LOOP(2) }
        else if (g.act == 3) { L2B_ACT_LOOP(3) }
        else if (g.act == 4) { L2B_ACT_LOOP(4) }
        else if (g.act == 5) { L2B_ACT_LOOP(5) }
        else { L2B_ACT_LOOP(0) }
#undef L2B_ACT_LOOP
      }
      if (g.trans) {                                            // out[nn + i][m]: lanes write consecutive addresses
        if (g.out_f32) {
          float* p = reinterpret_cast<float*>(g.out) + (size_t)nn * g.ldo + m;
#pragma unroll
          for (int i = 0; i < 16; ++i)
            if (i < ncol) {
              float x = __uint_as_float(r[i]);
              if (g.accumulate) x += p[(size_t)i * g.ldo];
              p[(size_t)i * g.ldo] = x;
            }
        } else {
          __nv_bfloat16* p = reinterpret_cast<__nv_bfloat16*>(g.out) + (size_t)nn * g.ldo + m;
#pragma unroll
          for (int i = 0; i < 16; ++i)
            if (i < ncol) {
              float x = __uint_as_float(r[i]);
              if (g.accumulate) x += __bfloat162float(p[(size_t)i * g.ldo]);
              p[(size_t)i * g.ldo] = __float2bfloat16(x);
            }
        }
        continue;
      }
      if (g.out_f32) {
        float* p = reinterpret_cast<float*>(g.out) + (size_t)m * g.ldo + nn;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          if (q * 4 >= ncol) break;
          float4 o = make_float4(__uint_as_float(r[4 * q]), __uint_as_float(r[4 * q + 1]), __uint_as_float(r[4 * q + 2]),
                                 __uint_as_float(r[4 * q + 3]));
          if (g.accumulate) {
            const float4 old = *reinterpret_cast<const float4*>(p + 4 * q);
            o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w;
          }
          *reinterpret_cast<float4*>(p + 4 * q) = o;
        }
      } else {
        __nv_bfloat16* p = reinterpret_cast<__nv_bfloat16*>(g.out) + (size_t)m * g.ldo + nn;
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          if (q * 8 >= ncol) break;
          __align__(16) __nv_bfloat16 h[8];
          float v[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[8 * q + i]);
          if (g.accumulate) {
            *reinterpret_cast<uint4*>(h) = *reinterpret_cast<const uint4*>(p + 8 * q);
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] += __bfloat162float(h[i]);
          }
#pragma unroll
          for (int i = 0; i < 8; ++i) h[i] = __float2bfloat16(v[i]);
          *reinterpret_cast<uint4*>(p + 8 * q) = *reinterpret_cast<const uint4*>(h);
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 5) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(g.tmem_cols) : "memory");
  }
}

// split-K tail: out[m][n] = act( sum_z part[z][m][n] + bias[n] ) (+ out).  A block owns 32 groups of four consecutive
// outputs; its 8 "z lanes" per group each sum every 8th partial, then the 8 lane sums are added in lane order through
// shared memory: a fixed summation order (deterministic) with 8 x shorter dependent-load chains than a serial loop.
__global__ void __launch_bounds__(256) k_gemm_reduce(const float* __restrict__ part, int splits, long long M, long long N,
                                                     const float* __restrict__ bias, int act, void* out, long long ldo,
                                                     int out_f32, int accumulate, int trans) {
  __shared__ float4 red[8][32];
  const int q = threadIdx.x & 31, zl = threadIdx.x >> 5;
  const long long id = ((long long)blockIdx.x * 32 + q) * 4;
  const bool ok = id < M * N;
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  if (ok) {
    const float* p0 = part + id;
    const size_t plane = (size_t)M * N;
#pragma unroll 4
    for (int z = zl; z < splits; z += 8) {
      const float4 p = __ldg(reinterpret_cast<const float4*>(p0 + (size_t)z * plane));
      s.x += p.x; s.y += p.y; s.z += p.z; s.w += p.w;
    }
  }
  red[zl][q] = s;
  __syncthreads();
  if (zl != 0 || !ok) return;
#pragma unroll
  for (int z = 1; z < 8; ++z) {
    const float4 p = red[z][q];
    s.x += p.x; s.y += p.y; s.z += p.z; s.w += p.w;
  }
  const long long m = id / N, nn = id % N;                       // N % 8 == 0: the 4 elements share a row
  float v[4] = {s.x, s.y, s.z, s.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    if (bias) v[i] += __ldg(bias + (trans ? m : nn + i));
    v[i] = il_act(v[i], act);
  }
  if (trans) {                                                   // the partials are D^T: out[nn + i][m]
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const size_t at = (size_t)(nn + i) * ldo + m;
      if (out_f32) {
        float* p = reinterpret_cast<float*>(out) + at;
        *p = accumulate ? *p + v[i] : v[i];
      } else {
        __nv_bfloat16* p = reinterpret_cast<__nv_bfloat16*>(out) + at;
        *p = __float2bfloat16(accumulate ? __bfloat162float(*p) + v[i] : v[i]);
      }
    }
    return;
  }
  if (out_f32) {
    float* p = reinterpret_cast<float*>(out) + m * ldo + nn;
    if (accumulate) { v[0] += p[0]; v[1] += p[1]; v[2] += p[2]; v[3] += p[3]; }
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  } else {
    __nv_bfloat16* p = reinterpret_cast<__nv_bfloat16*>(out) + m * ldo + nn;
    if (accumulate) {
#pragma unroll
      for (int i = 0; i < 4; ++i) v[i] += __bfloat162float(p[i]);
    }
    __align__(8) __nv_bfloat16 h[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __float2bfloat16(v[i]);
    *reinterpret_cast<uint2*>(p) = *reinterpret_cast<const uint2*>(h);
  }
}

// x = x1 + x2 + x3 with three bf16 terms (8 + 8 + 8 mantissa bits): the operands of an fp32-accurate GEMM on the
// bf16 tensor cores (six products a_i b_j, i + j <= 4, fp32 accumulation).  out[t][r][c], c < out_ld, zero padded.
__global__ void __launch_bounds__(256) k_split_bf16x3(const float* __restrict__ x, long long rows, long long cols,
                                                      long long ld, __nv_bfloat16* __restrict__ out, long long out_ld) {
  const long long id = (long long)blockIdx.x * 256 + threadIdx.x;          // one thread per 8 output columns
  const long long per_row = out_ld / 8;
  if (id >= rows * per_row) return;
  const long long r = id / per_row, c0 = (id % per_row) * 8;
  __align__(16) __nv_bfloat16 h[3][8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float v = (c0 + i < cols) ? __ldg(x + r * ld + c0 + i) : 0.f;
    const __nv_bfloat16 a = __float2bfloat16(v);
    const float r1 = v - __bfloat162float(a);
    const __nv_bfloat16 b = __float2bfloat16(r1);
    const float r2 = r1 - __bfloat162float(b);
    h[0][i] = a; h[1][i] = b; h[2][i] = __float2bfloat16(r2);
  }
  const size_t plane = (size_t)rows * out_ld;
#pragma unroll
  for (int t = 0; t < 3; ++t)
    *reinterpret_cast<uint4*>(out + t * plane + r * out_ld + c0) = *reinterpret_cast<const uint4*>(h[t]);
}

int pick_bn(int N, int b_kmajor, int M = 0, long long chunks = 1 << 30, int nsm = 0) {
  const int q = b_kmajor ? 16 : 64;
  int bn = (N + q - 1) / q * q;
  if (bn > 256) bn = 256;
  // a short K loop on a handful of tiles (a hidden Linear): the epilogue's serial TMEM round trips are the critical
  // path, so narrower tiles on more SMs finish sooner
  if (M > 0 && chunks < 16) {
    const long long mt = (M + G_BM - 1) / G_BM;
    while (bn >= 2 * q && bn % (2 * q) == 0 && bn / 2 >= 32 && mt * ((N + bn - 1) / bn) * 4 <= nsm) bn /= 2;
  }
  return bn;
}

using EncodeFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                              const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                              CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// rank-2 descriptor of a row-major bf16 matrix [rows][cols] with leading dimension ld, box [box_rows][64 columns]
// (128-byte rows, 128-byte swizzle), zeros outside the matrix
int make_map(CUtensorMap* m, const void* base, long long rows, long long cols, long long ld, int box_rows) {
  static EncodeFn encode = nullptr;
  if (encode == nullptr) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qr;
    L2B_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr));
    L2B_REQUIRE(fn != nullptr && qr == cudaDriverEntryPointSuccess, L2B_ERR_CUDA, "cuTensorMapEncodeTiled not available");
    encode = (EncodeFn)fn;
  }
  const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  const cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult rc = encode(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  L2B_REQUIRE(rc == CUDA_SUCCESS, L2B_ERR_CUDA, "cuTensorMapEncodeTiled failed (CUresult %d; rows=%lld cols=%lld ld=%lld)",
              (int)rc, rows, cols, ld);
  return L2B_OK;
}

template <int MAXSEG>
int launch_gemm(const void* const* a_ptrs, long long lda, const void* const* b_ptrs, long long ldb, const GemmArgs& g,
                dim3 grid, size_t smem, cudaStream_t st) {
  GemmMaps<MAXSEG> maps;
  for (int s = 0; s < g.nseg; ++s) {
    const int kbox = g.kpack == 2 ? 32 : 64;
    int rc = g.a_kmajor ? make_map(&maps.a[s], a_ptrs[s], g.M, g.K, lda, G_BM) : make_map(&maps.a[s], a_ptrs[s], g.K, g.M, lda, kbox);
    if (rc != L2B_OK) return rc;
    rc = g.b_kmajor ? make_map(&maps.b[s], b_ptrs[s], g.N, g.K, ldb, g.BN) : make_map(&maps.b[s], b_ptrs[s], g.K, g.N, ldb, kbox);
    if (rc != L2B_OK) return rc;
  }
  for (int s = g.nseg; s < MAXSEG; ++s) { maps.a[s] = maps.a[0]; maps.b[s] = maps.b[0]; }
  static bool attr_set = false;
  if (!attr_set) {
    L2B_CUDA(cudaFuncSetAttribute((const void*)k_gemm_bf16<MAXSEG>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  G_NST * (G_BM + 256) * G_BK * 2 + 1024));
    attr_set = true;
  }
  k_gemm_bf16<MAXSEG><<<grid, G_NTH, smem, st>>>(maps, g);
  L2B_LAUNCHED("k_gemm_bf16");
  return L2B_OK;
}

}  // namespace
}  // namespace l2b

using namespace l2b;

extern "C" {

int l2b_gemm_bf16_splits(int M, int N, int K, int nseg, int b_kmajor) {
  if (M <= 0 || N <= 0 || K <= 0 || nseg <= 0) return 0;
  int dev = 0, nsm = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess)
    return 0;
  const int bn = pick_bn(N, b_kmajor);
  const long long tiles = (long long)((M + G_BM - 1) / G_BM) * ((N + bn - 1) / bn);
  const long long chunks = (long long)nseg * ((K + G_BK - 1) / G_BK);
  if (tiles * 2 > nsm || chunks < 16) return 1;                  // enough tiles, or too little K to be worth a second pass
  long long s = nsm / tiles;
  if (s > chunks / 4) s = chunks / 4;                            // at least four stages per CTA
  return (int)(s < 1 ? 1 : s);
}

size_t l2b_gemm_bf16_ws_bytes(int M, int N, int splits) {
  if (M <= 0 || N <= 0 || splits <= 1) return 0;
  return align_up((size_t)splits * M * N * sizeof(float), 256);
}

int l2b_split_bf16x3(const float* x, long long rows, long long cols, long long ld, void* out, long long out_ld,
                     void* stream) {
  L2B_REQUIRE(x && out, L2B_ERR_INVALID, "null pointer");
  L2B_REQUIRE(rows > 0 && cols > 0 && ld >= cols, L2B_ERR_INVALID, "rows, cols must be positive and ld >= cols");
  L2B_REQUIRE(out_ld % 8 == 0 && out_ld >= cols && ((uintptr_t)out & 15) == 0, L2B_ERR_INVALID,
              "out_ld must be a multiple of 8 >= cols and out 16-byte aligned");
  const long long n = rows * (out_ld / 8);
  k_split_bf16x3<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(x, rows, cols, ld, (__nv_bfloat16*)out,
                                                                                out_ld);
  L2B_LAUNCHED("k_split_bf16x3");
  return L2B_OK;
}

int l2b_gemm_bf16(const void* const* a_ptrs, long long lda, int a_kmajor, const void* const* b_ptrs, long long ldb,
                  int b_kmajor, int nseg, int seg_inner, int M, int N, int K, void* out, int out_dtype, long long ldo,
                  int accumulate, const float* bias, int activation, int splits, void* ws, size_t ws_bytes,
                  void* stream) {
  L2B_REQUIRE(a_ptrs && b_ptrs && out, L2B_ERR_INVALID, "null pointer");
  L2B_REQUIRE(nseg >= 1 && nseg <= G_MAXSEG, L2B_ERR_INVALID, "nseg must be in [1, %d] (got %d)", G_MAXSEG, nseg);
  L2B_REQUIRE(M > 0 && N > 0 && K > 0, L2B_ERR_INVALID, "M, N, K must be positive");
  L2B_REQUIRE(out_dtype == L2B_BF16 || out_dtype == L2B_F32, L2B_ERR_UNSUPPORTED, "out_dtype must be L2B_BF16 or L2B_F32");
  L2B_REQUIRE(activation >= 0 && activation <= 5, L2B_ERR_INVALID, "activation code must be in [0, 5]");
  L2B_REQUIRE(N % 8 == 0 && ldo % 8 == 0 && ldo >= N, L2B_ERR_UNSUPPORTED,
              "N and ldo must be multiples of 8 with ldo >= N (N=%d ldo=%lld)", N, ldo);
  L2B_REQUIRE(lda % 8 == 0 && ldb % 8 == 0, L2B_ERR_UNSUPPORTED, "lda, ldb must be multiples of 8 (16-byte units)");
  L2B_REQUIRE(a_kmajor ? (K % 8 == 0 && lda >= K) : (M % 8 == 0 && lda >= M), L2B_ERR_UNSUPPORTED,
              "A: the stored row length (K if K-major, M if MN-major) must be a multiple of 8 and <= lda");
  L2B_REQUIRE(b_kmajor ? (K % 8 == 0 && ldb >= K) : ldb >= N, L2B_ERR_UNSUPPORTED,
              "B: the stored row length (K if K-major, N if MN-major) must be a multiple of 8 and <= ldb");
  L2B_REQUIRE(((uintptr_t)out & 15) == 0, L2B_ERR_INVALID, "out must be 16-byte aligned");
  L2B_REQUIRE(splits >= 1, L2B_ERR_INVALID, "splits must be >= 1");
  for (int s = 0; s < nseg; ++s) {
    L2B_REQUIRE(a_ptrs[s] && b_ptrs[s], L2B_ERR_INVALID, "null operand pointer");
    L2B_REQUIRE((((uintptr_t)a_ptrs[s] | (uintptr_t)b_ptrs[s]) & 15) == 0, L2B_ERR_INVALID,
                "operands must be 16-byte aligned");
  }
  GemmArgs g;
  // dW = sum_u dY_u^T X_u with 32 chains per update: two updates fill one 64-row stage
  g.kpack = (!a_kmajor && !b_kmajor && K <= 32 && nseg >= 2 && !seg_inner) ? 2 : 1;
  g.seg_inner = seg_inner ? 1 : 0;
  const long long chunks = g.kpack == 2 ? (nseg + 1) / 2 : (long long)nseg * ((K + G_BK - 1) / G_BK);
  if (splits > chunks) splits = (int)chunks;
  // A skinny M with a short K loop (dX = dY W of the input layer: 32 chains x 131 072 columns) would leave three of
  // the four TMEM lane quarters -- and epilogue warps -- idle and pad the A tile to 128 rows: compute D^T instead,
  // the long axis on the 128 TMEM lanes, the chains as UMMA N; the epilogue writes out[n][m].
  g.trans = (M <= 64 && M % 8 == 0 && N > M && splits == 1) ? 1 : 0;
  if (g.trans) {
    const void* const* tp = a_ptrs; a_ptrs = b_ptrs; b_ptrs = tp;
    const long long tl = lda; lda = ldb; ldb = tl;
    const int tk = a_kmajor; a_kmajor = b_kmajor; b_kmajor = tk;
    const int tm = M; M = N; N = tm;
  }
  g.a_kmajor = a_kmajor ? 1 : 0;
  g.b_kmajor = b_kmajor ? 1 : 0;
  g.nseg = nseg; g.M = M; g.N = N; g.K = K;
  int nsm = 0, dev = 0;
  L2B_CUDA(cudaGetDevice(&dev));
  L2B_CUDA(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev));
  g.BN = pick_bn(N, g.b_kmajor, M, chunks, nsm);
  g.n_mt = (M + G_BM - 1) / G_BM;
  g.n_nt = (N + g.BN - 1) / g.BN;
  g.out = out; g.ldo = ldo; g.out_f32 = out_dtype == L2B_F32; g.accumulate = accumulate ? 1 : 0; g.act = activation;
  g.bias = bias;
  g.part = nullptr;
  if (splits > 1) {
    L2B_REQUIRE(ws != nullptr && ((uintptr_t)ws & 15) == 0 && ws_bytes >= l2b_gemm_bf16_ws_bytes(M, N, splits),
                L2B_ERR_WORKSPACE, "workspace too small for the split-K partials");
    g.part = (float*)ws;
  }
  uint32_t cols = 32;
  while (cols < (uint32_t)g.BN) cols <<= 1;
  g.tmem_cols = cols;
  const long long tiles = (long long)g.n_mt * g.n_nt;
  L2B_REQUIRE(tiles <= 0x7fffffffLL && splits <= 65535, L2B_ERR_UNSUPPORTED, "grid too large");
  // Several waves of tiles with a short K loop each: two CTAs per SM (two stages each), so that one CTA's prologue /
  // epilogue overlaps the other's loads; otherwise one CTA per SM with a deeper ring.
  const size_t stage = (size_t)(G_BM + g.BN) * G_BK * 2;
  g.nst = G_NST;
  if (tiles * splits > nsm) g.nst = stage > 36 * 1024 ? 2 : 3;
  const long long per_cta = (chunks + splits - 1) / splits;
  if (g.nst > per_cta) g.nst = per_cta < 2 ? 2 : (int)per_cta;
  const size_t smem = (size_t)g.nst * stage + 1024;              // + alignment slack for the 1024-byte swizzle atoms
  cudaStream_t st = (cudaStream_t)stream;
  const dim3 grid((unsigned)tiles, (unsigned)splits);
  const int rc = nseg <= 4 ? launch_gemm<4>(a_ptrs, lda, b_ptrs, ldb, g, grid, smem, st)
                           : launch_gemm<G_MAXSEG>(a_ptrs, lda, b_ptrs, ldb, g, grid, smem, st);
  if (rc != L2B_OK) return rc;
  if (splits > 1) {
    const long long n4 = ((long long)M * N + 3) / 4;
    k_gemm_reduce<<<(unsigned)((n4 + 31) / 32), 256, 0, st>>>(g.part, splits, M, N, bias, activation, out, ldo,
                                                              g.out_f32, g.accumulate, g.trans);
    L2B_LAUNCHED("k_gemm_reduce");
  }
  return L2B_OK;
}

}  // extern "C"
