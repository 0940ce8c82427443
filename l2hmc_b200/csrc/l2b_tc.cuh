// l2b_tc.cuh -- tcgen05 / TMEM / mbarrier / bulk-copy primitives (inline PTX, sm_100a) shared by the tensor-core
// translation units (l2b_vnet.cu: fused heads + input layer; l2b_gemm.cu: the general bf16 GEMM).
#pragma once
#include <cuda_bf16.h>
#include <stdint.h>

namespace l2b {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// bounded spin: a protocol bug traps (CUDA error) instead of hanging the GPU
template <int SLEEP_NS = 0>
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  for (uint32_t spin = 0; !ok; ++spin) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    if (!ok) {
      if (SLEEP_NS > 0) __nanosleep(SLEEP_NS);          // waiting epilogue warps stay off the issue ports
      if (spin > (1u << 24)) __trap();
    }
  }
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
// shared-memory matrix descriptor, no swizzle, K-major canonical layout:
// core matrix = 8 rows x 16 B contiguous; LBO = byte step between core matrices along K,
// SBO = byte step between 8-row groups along M/N; bits 46-47 = 1 (sm_100 descriptor version)
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) |
         (1ull << 46);
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t r[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld4(uint32_t taddr, uint32_t r[4]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(taddr));
}
// tanh(x) = 1 - 2 / (1 + e^{2x}) on the SFU (ex2.approx + rcp.approx): absolute error ~2e-7, which is
// what the epilogue needs (s, q enter through eps*s/2 and eps*q); saturates correctly for large |x|
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float tanh_fast(float x) {
  const float e = ex2_approx(2.885390081777927f * x);   // e^{2x}; inf -> rcp = 0 -> 1, 0 -> -1
  return fmaf(-2.0f, rcp_approx(1.0f + e), 1.0f);
}
__device__ __forceinline__ float exp_fast(float x) { return ex2_approx(1.4426950408889634f * x); }

// activations of the reference (network.py:40-46): 0 identity, 1 tanh, 2 relu, 3 swish (SiLU), 4 leaky_relu(0.01), 5 elu
__device__ __forceinline__ float il_act(float x, int act) {
  switch (act) {
    case 1: return tanhf(x);
    case 2: return fmaxf(x, 0.f);
    case 3: return x / (1.f + __expf(-x));
    case 4: return x > 0.f ? x : 0.01f * x;
    case 5: return x > 0.f ? x : expm1f(x);
    default: return x;
  }
}

}  // namespace l2b
