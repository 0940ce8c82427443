// l2b_gemm.cu -- the dense layers of the L2HMC networks as ONE hand-written bf16 GEMM on the 5th-generation tensor
// cores (sm_100a, tcgen05.mma + TMEM), for every Linear the fused kernels of l2b_vnet.cu do not already cover:
// the hidden Linears (reference network/pytorch/network.py:489-493, 538-541), the input Linears under autograd
// (:415-422), and the three GEMMs of every Linear's backward pass (dX = dY W, dW = dY^T X) that the reference gets
// from ATen autograd / cuBLAS.
//
//     D[m][n] = sum_seg sum_k A_seg(m, k) B_seg(n, k)          (fp32 accumulation in TMEM)
//
// Each operand is an ordinary ROW-MAJOR bf16 matrix and may be contracted over either of its axes:
//     K-major  : stored [MN][K]  (the contraction index is contiguous)          -- x and W of a Linear's forward
//     MN-major : stored [K][MN]  (the contraction index is the row)             -- W in dX = dY W, dY and X in dW
// so no transposed copy of an activation, a weight or a cotangent is ever made.  Up to three (A, B) pairs of the same
// shape are summed in one launch ("segments": the two input Linears, the three heads of dz).
//
// Data path.  Four loader warps move 16-byte units (8 bf16 along the stored row) with cp.async.cg straight into the
// canonical NO-SWIZZLE UMMA layout: unit (row r, column block c) of an R-row tile lands at (c R + r) 16 B.  That one
// rule produces the K-major core matrices (R = tile rows, c = K block) and the MN-major ones (R = 64 K-rows, c = MN
// block) alike; a lane quad-of-8 mapping keeps the global reads in full 64-byte runs and the shared-memory writes in
// conflict-free 128-byte runs.  Units outside the matrix are zero-filled (src-size 0), so ragged M, N, K need no
// special case.  Completion is signalled per stage by cp.async.mbarrier.arrive.noinc; one thread of a fifth warp
// waits, crosses the proxy fence and issues four tcgen05.mma (M128 x N(BN) x K16, kind::f16) per stage, recycling
// stages with tcgen05.commit.  The loader warps then drain TMEM (tcgen05.ld 32x32b: lane = row m, columns = n):
// bias, activation, optional accumulate, bf16 / fp32 store -- or fp32 split-K partials that k_gemm_reduce finishes
// in a fixed order (deterministic).
#include <cuda_bf16.h>

#include "l2b_common.cuh"
#include "l2b_tc.cuh"

namespace l2b {
namespace {

constexpr int G_BM = 128;            // UMMA M
constexpr int G_BK = 64;             // K per pipeline stage
constexpr int G_NST = 4;             // stages
constexpr int G_LOADERS = 128;       // 4 loader / epilogue warps
constexpr int G_NTH = 160;           // + 1 MMA warp

struct GemmOperand {
  const __nv_bfloat16* ptr[3];
  long long ld;                      // elements between stored rows
  int kmajor;                        // 1: stored [MN][K]; 0: stored [K][MN]
};

struct GemmArgs {
  GemmOperand a, b;
  int nseg, M, N, K;
  int BN;                            // N tile: multiple of 16 (32 when B is MN-major), <= 256
  int n_mt, n_nt;
  void* out;                         // [M][ldo] bf16 or fp32 (splits == 1)
  long long ldo;
  int out_f32, accumulate, act;
  const float* bias;                 // [N] or null
  float* part;                       // [splits][M][N] fp32 (splits > 1)
  uint32_t tmem_cols;
};

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_arrive_noinc(uint32_t bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
}

// one operand tile of one stage: `R` stored rows x `C8` 16-byte units per row
__device__ __forceinline__ void load_tile(const __nv_bfloat16* base, long long ld, long long row0, long long col0,
                                          long long rows_total, long long cols_total, int R, int C8, uint32_t dst,
                                          int t) {
  const int units = R * C8;
  const int cq = C8 >> 2;
  for (int u = t; u < units; u += G_LOADERS) {
    const int rl = u & 7, cl = (u >> 3) & 3, rest = u >> 5;
    const int cg = rest % cq, rg = rest / cq;
    const int r = rg * 8 + rl, c8 = cg * 4 + cl;
    const long long gr = row0 + r, gc = col0 + (long long)c8 * 8;
    const bool ok = gr < rows_total && gc < cols_total;
    const __nv_bfloat16* src = ok ? base + gr * ld + gc : base;
    cp_async16(dst + (uint32_t)(c8 * R + r) * 16u, src, ok ? 16u : 0u);
  }
}

__global__ void __launch_bounds__(G_NTH, 1) k_gemm_bf16(const GemmArgs g) {
  extern __shared__ __align__(128) unsigned char gsm[];
  __shared__ __align__(8) unsigned long long full[G_NST], empty[G_NST], accum;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int mt = blockIdx.x % g.n_mt, nt = blockIdx.x / g.n_mt;      // M tiles of one N tile are adjacent (share B in L2)
  const long long m0 = (long long)mt * G_BM, n0 = (long long)nt * g.BN;
  const int nkc = (g.K + G_BK - 1) / G_BK;
  const long long total = (long long)g.nseg * nkc;
  const long long c0 = (long long)blockIdx.y * total / gridDim.y, c1 = (long long)(blockIdx.y + 1) * total / gridDim.y;
  const int n = (int)(c1 - c0);
  const uint32_t a_bytes = G_BM * G_BK * 2, b_bytes = (uint32_t)g.BN * G_BK * 2, stage_bytes = a_bytes + b_bytes;
  const uint32_t smem0 = smem_u32(gsm);

  if (tid == 0) {
    for (int s = 0; s < G_NST; ++s) { mbar_init(smem_u32(&full[s]), G_LOADERS); mbar_init(smem_u32(&empty[s]), 1); }
    mbar_init(smem_u32(&accum), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 4) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)),
                 "r"(g.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base_s;

  if (warp < 4) {
    // ---- loaders ------------------------------------------------------------------------------------------------
    for (int s = 0; s < n; ++s) {
      const uint32_t slot = (uint32_t)(s % G_NST);
      if (s >= G_NST) mbar_wait<32>(smem_u32(&empty[slot]), (uint32_t)(s / G_NST - 1) & 1u);
      const long long c = c0 + s;
      const int seg = (int)(c / nkc);
      const long long k0 = (c - (long long)seg * nkc) * G_BK;
      const uint32_t a_dst = smem0 + slot * stage_bytes, b_dst = a_dst + a_bytes;
      if (g.a.kmajor) load_tile(g.a.ptr[seg], g.a.ld, m0, k0, g.M, g.K, G_BM, G_BK / 8, a_dst, tid);
      else load_tile(g.a.ptr[seg], g.a.ld, k0, m0, g.K, g.M, G_BK, G_BM / 8, a_dst, tid);
      if (g.b.kmajor) load_tile(g.b.ptr[seg], g.b.ld, n0, k0, g.N, g.K, g.BN, G_BK / 8, b_dst, tid);
      else load_tile(g.b.ptr[seg], g.b.ld, k0, n0, g.K, g.N, G_BK, g.BN / 8, b_dst, tid);
      cp_async_arrive_noinc(smem_u32(&full[slot]));
    }
  } else if (tid == G_LOADERS) {
    // ---- MMA issuer -----------------------------------------------------------------------------------------------
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((g.a.kmajor ? 0u : 1u) << 15) |
                           ((g.b.kmajor ? 0u : 1u) << 16) | ((uint32_t)(g.BN >> 3) << 17) | ((uint32_t)(G_BM >> 4) << 24);
    // K-major: LBO = bytes between K blocks (R 16), SBO = 128 (8-row groups adjacent); one K16 step = 2 K blocks
    // MN-major: LBO = 128 (8-K-row groups adjacent), SBO = bytes between MN blocks (64 16); one K16 step = 16 rows
    const uint32_t a_lbo = g.a.kmajor ? G_BM * 16u : 128u, a_sbo = g.a.kmajor ? 128u : G_BK * 16u;
    const uint32_t b_lbo = g.b.kmajor ? (uint32_t)g.BN * 16u : 128u, b_sbo = g.b.kmajor ? 128u : G_BK * 16u;
    const uint32_t a_step = g.a.kmajor ? 2u * G_BM * 16u : 256u, b_step = g.b.kmajor ? 2u * (uint32_t)g.BN * 16u : 256u;
    for (int s = 0; s < n; ++s) {
      const uint32_t slot = (uint32_t)(s % G_NST), ph = (uint32_t)(s / G_NST) & 1u;
      mbar_wait(smem_u32(&full[slot]), ph);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // cp.async (generic proxy) writes -> tensor core
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t a_base = smem0 + slot * stage_bytes, b_base = a_base + a_bytes;
#pragma unroll
      for (int kk = 0; kk < G_BK / 16; ++kk) {
        const uint64_t ad = umma_desc(a_base + kk * a_step, a_lbo, a_sbo);
        const uint64_t bd = umma_desc(b_base + kk * b_step, b_lbo, b_sbo);
        umma_f16(tmem, ad, bd, idesc, (s | kk) != 0 ? 1u : 0u);
      }
      umma_commit(smem_u32(&empty[slot]));
    }
    umma_commit(smem_u32(&accum));
  }
  __syncwarp();

  // ---- epilogue: TMEM lane = row m of the tile, columns = n ------------------------------------------------------
  if (warp < 4) {
    if (n > 0) {
      mbar_wait<64>(smem_u32(&accum), 0);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
    const long long m = m0 + warp * 32 + lane;
    const uint32_t trow = tmem + ((uint32_t)(warp * 32) << 16);
    const bool split = gridDim.y > 1;
    for (int c = 0; c < g.BN; c += 8) {
      uint32_t r[8];
      if (n > 0) {
        tmem_ld8(trow + c, r);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) r[i] = 0u;
      }
      const long long nn = n0 + c;
      if (m >= g.M || nn >= g.N) continue;                      // N % 8 == 0: an 8-column group is all in or all out
      if (split) {
        float* p = g.part + ((size_t)blockIdx.y * g.M + m) * g.N + nn;
        *reinterpret_cast<uint4*>(p) = make_uint4(r[0], r[1], r[2], r[3]);
        *reinterpret_cast<uint4*>(p + 4) = make_uint4(r[4], r[5], r[6], r[7]);
        continue;
      }
      float v[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float x = __uint_as_float(r[i]);
        if (g.bias) x += __ldg(g.bias + nn + i);
        v[i] = il_act(x, g.act);
      }
      if (g.out_f32) {
        float* p = reinterpret_cast<float*>(g.out) + m * g.ldo + nn;
        if (g.accumulate) {
          const float4 o0 = *reinterpret_cast<const float4*>(p), o1 = *reinterpret_cast<const float4*>(p + 4);
          v[0] += o0.x; v[1] += o0.y; v[2] += o0.z; v[3] += o0.w; v[4] += o1.x; v[5] += o1.y; v[6] += o1.z; v[7] += o1.w;
        }
        *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
        *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
      } else {
        __nv_bfloat16* p = reinterpret_cast<__nv_bfloat16*>(g.out) + m * g.ldo + nn;
        __align__(16) __nv_bfloat16 h[8];
        if (g.accumulate) {
          *reinterpret_cast<uint4*>(h) = *reinterpret_cast<const uint4*>(p);
#pragma unroll
          for (int i = 0; i < 8; ++i) v[i] += __bfloat162float(h[i]);
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) h[i] = __float2bfloat16(v[i]);
        *reinterpret_cast<uint4*>(p) = *reinterpret_cast<const uint4*>(h);
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 4) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(g.tmem_cols) : "memory");
  }
}

// split-K tail: out[m][n] = act( sum_z part[z][m][n] + bias[n] ) (+ out), fixed summation order
__global__ void __launch_bounds__(256) k_gemm_reduce(const float* __restrict__ part, int splits, long long M, long long N,
                                                     const float* __restrict__ bias, int act, void* out, long long ldo,
                                                     int out_f32, int accumulate) {
  const long long id = ((long long)blockIdx.x * 256 + threadIdx.x) * 4;
  if (id >= M * N) return;
  const long long m = id / N, nn = id % N;                       // N % 8 == 0: the 4 elements share a row
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int z = 0; z < splits; ++z) {
    const float4 p = *reinterpret_cast<const float4*>(part + (size_t)z * M * N + id);
    s.x += p.x; s.y += p.y; s.z += p.z; s.w += p.w;
  }
  float v[4] = {s.x, s.y, s.z, s.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    if (bias) v[i] += __ldg(bias + nn + i);
    v[i] = il_act(v[i], act);
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

int pick_bn(int N, int b_kmajor) {
  const int q = b_kmajor ? 16 : 32;
  int bn = (N + q - 1) / q * q;
  return bn > 256 ? 256 : bn;
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

int l2b_gemm_bf16(const void* const* a_ptrs, long long lda, int a_kmajor, const void* const* b_ptrs, long long ldb,
                  int b_kmajor, int nseg, int M, int N, int K, void* out, int out_dtype, long long ldo, int accumulate,
                  const float* bias, int activation, int splits, void* ws, size_t ws_bytes, void* stream) {
  L2B_REQUIRE(a_ptrs && b_ptrs && out, L2B_ERR_INVALID, "null pointer");
  L2B_REQUIRE(nseg >= 1 && nseg <= 3, L2B_ERR_INVALID, "nseg must be 1, 2 or 3 (got %d)", nseg);
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
  GemmArgs g;
  for (int s = 0; s < 3; ++s) {
    g.a.ptr[s] = (const __nv_bfloat16*)a_ptrs[s < nseg ? s : 0];
    g.b.ptr[s] = (const __nv_bfloat16*)b_ptrs[s < nseg ? s : 0];
    L2B_REQUIRE(g.a.ptr[s] && g.b.ptr[s], L2B_ERR_INVALID, "null operand pointer");
    L2B_REQUIRE((((uintptr_t)g.a.ptr[s] | (uintptr_t)g.b.ptr[s]) & 15) == 0, L2B_ERR_INVALID,
                "operands must be 16-byte aligned");
  }
  g.a.ld = lda; g.a.kmajor = a_kmajor ? 1 : 0;
  g.b.ld = ldb; g.b.kmajor = b_kmajor ? 1 : 0;
  g.nseg = nseg; g.M = M; g.N = N; g.K = K;
  g.BN = pick_bn(N, g.b.kmajor);
  g.n_mt = (M + G_BM - 1) / G_BM;
  g.n_nt = (N + g.BN - 1) / g.BN;
  const long long chunks = (long long)nseg * ((K + G_BK - 1) / G_BK);
  if (splits > chunks) splits = (int)chunks;
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
  const size_t smem = (size_t)G_NST * (G_BM + g.BN) * G_BK * 2;
  cudaStream_t st = (cudaStream_t)stream;
  L2B_CUDA(cudaFuncSetAttribute((const void*)k_gemm_bf16, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k_gemm_bf16<<<dim3((unsigned)tiles, (unsigned)splits), G_NTH, smem, st>>>(g);
  L2B_LAUNCHED("k_gemm_bf16");
  if (splits > 1) {
    const long long n4 = ((long long)M * N + 3) / 4;
    k_gemm_reduce<<<(unsigned)((n4 + 255) / 256), 256, 0, st>>>(g.part, splits, M, N, bias, activation, out, ldo,
                                                                g.out_f32, g.accumulate);
    L2B_LAUNCHED("k_gemm_reduce");
  }
  return L2B_OK;
}

}  // extern "C"
