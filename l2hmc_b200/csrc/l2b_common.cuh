// l2b_common.cuh -- error plumbing, launch accounting and block-level helpers
// shared by the SU(3) and U(1) translation units of libl2b.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/l2b.h"

namespace l2b {

// thread-local message + process-wide launch counter (defined in l2b_capi.cu)
void set_error(const char* fmt, ...);
void count_launch(int n = 1);

#define L2B_REQUIRE(cond, code, ...)  \
  do {                                \
    if (!(cond)) {                    \
      ::l2b::set_error(__VA_ARGS__);  \
      return (code);                  \
    }                                 \
  } while (0)

// call right after a <<<>>> launch
#define L2B_LAUNCHED(name)                                                           \
  do {                                                                               \
    ::l2b::count_launch();                                                           \
    cudaError_t e__ = cudaGetLastError();                                            \
    if (e__ != cudaSuccess) {                                                        \
      ::l2b::set_error("launch of %s failed: %s", name, cudaGetErrorString(e__));    \
      return L2B_ERR_CUDA;                                                           \
    }                                                                                \
  } while (0)

#define L2B_CUDA(call)                                                               \
  do {                                                                               \
    cudaError_t e__ = (call);                                                        \
    if (e__ != cudaSuccess) {                                                        \
      ::l2b::set_error("%s failed: %s", #call, cudaGetErrorString(e__));             \
      return L2B_ERR_CUDA;                                                           \
    }                                                                                \
  } while (0)

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

#if defined(__CUDACC__)
// Sum over the thread block; result valid in linear thread 0.  `NT` = threads per
// block (multiple of 32, <= 1024); sm must hold NT/32 values.  Fixed shuffle tree
// => bit-reproducible for a fixed launch geometry.
template <int NT, typename T>
__device__ __forceinline__ T block_sum(T v, T* sm, int tid) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  const int lane = tid & 31, warp = tid >> 5;
  if (lane == 0) sm[warp] = v;
  __syncthreads();
  if (warp == 0) {
    v = (lane < NT / 32) ? sm[lane] : T(0);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  }
  __syncthreads();
  return v;
}

template <int NT, typename T>
__device__ __forceinline__ T block_max(T v, T* sm, int tid) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = max(v, __shfl_down_sync(0xffffffffu, v, o));
  const int lane = tid & 31, warp = tid >> 5;
  if (lane == 0) sm[warp] = v;
  __syncthreads();
  if (warp == 0) {
    v = (lane < NT / 32) ? sm[lane] : sm[0];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = max(v, __shfl_down_sync(0xffffffffu, v, o));
  }
  __syncthreads();
  return v;
}
#endif

}  // namespace l2b
