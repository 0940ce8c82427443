// l2b_capi.cu -- error reporting, launch accounting and the group-agnostic
// Metropolis accept/reject mix of libl2b.
#include <atomic>
#include <stdarg.h>
#include <string.h>

#include "l2b_common.cuh"

namespace l2b {

static thread_local char g_err[512] = "";
static std::atomic<uint64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

void count_launch(int n) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }

namespace {

// out[b, :] = accept[b] ? prop[b, :] : init[b, :]; one row = `row_vecs` elements of V.
// The reference forms ma*prop + mr*init with ma in {0, 1} (dynamics.py:639-649), which
// selects bit-exactly for finite inputs.
template <typename V>
__global__ void __launch_bounds__(256) k_accept_mix(const V* __restrict__ init, const V* __restrict__ prop,
                                                    V* __restrict__ out, size_t row_vecs,
                                                    const float* __restrict__ accept) {
  const int b = blockIdx.y;
  const bool acc = __ldg(accept + b) != 0.0f;
  const V* src = (acc ? prop : init) + (size_t)b * row_vecs;
  V* dst = out + (size_t)b * row_vecs;
  for (size_t k = (size_t)blockIdx.x * 256 + threadIdx.x; k < row_vecs; k += (size_t)gridDim.x * 256)
    dst[k] = __ldg(src + k);
}

template <typename V>
int launch_mix(const void* init, const void* prop, void* out, size_t row_bytes, const float* accept, int nb,
               cudaStream_t st) {
  const size_t row_vecs = row_bytes / sizeof(V);
  size_t nblk = (row_vecs + 256 * 4 - 1) / (256 * 4);
  if (nblk < 1) nblk = 1;
  if (nblk > 4096) nblk = 4096;
  k_accept_mix<V><<<dim3((unsigned)nblk, nb), 256, 0, st>>>((const V*)init, (const V*)prop, (V*)out, row_vecs, accept);
  L2B_LAUNCHED("k_accept_mix");
  return L2B_OK;
}

}  // namespace
}  // namespace l2b

using namespace l2b;

extern "C" {

const char* l2b_last_error(void) { return g_err; }

int l2b_version(void) { return 200; }

#ifndef L2B_ABI_HASH
#define L2B_ABI_HASH 0u
#endif
uint32_t l2b_abi_hash(void) { return (uint32_t)L2B_ABI_HASH; }
#ifndef L2B_SRC_HASH
#define L2B_SRC_HASH 0u
#endif
uint32_t l2b_source_hash(void) { return (uint32_t)L2B_SRC_HASH; }

uint64_t l2b_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

int l2b_accept_mix(const void* const* host_init, const void* const* host_prop, void* const* host_out,
                   const size_t* host_row_bytes, int nfields, const float* accept, int nb, void* stream) {
  L2B_REQUIRE(host_init && host_prop && host_out && host_row_bytes && accept, L2B_ERR_INVALID, "null pointer");
  L2B_REQUIRE(nb > 0 && nfields > 0, L2B_ERR_INVALID, "nb and nfields must be positive");
  L2B_REQUIRE(nb <= 65535, L2B_ERR_UNSUPPORTED, "nb=%d exceeds grid.y limit", nb);
  cudaStream_t st = (cudaStream_t)stream;
  for (int k = 0; k < nfields; ++k) {
    const void* a = host_init[k];
    const void* b = host_prop[k];
    void* o = host_out[k];
    const size_t rb = host_row_bytes[k];
    L2B_REQUIRE(a && b && o, L2B_ERR_INVALID, "null field pointer (field %d)", k);
    if (rb == 0) continue;
    const uintptr_t bits = (uintptr_t)a | (uintptr_t)b | (uintptr_t)o | (uintptr_t)rb;
    int rc;
    if ((bits & 15) == 0) rc = launch_mix<uint4>(a, b, o, rb, accept, nb, st);
    else if ((bits & 3) == 0) rc = launch_mix<uint32_t>(a, b, o, rb, accept, nb, st);
    else rc = launch_mix<unsigned char>(a, b, o, rb, accept, nb, st);
    if (rc != L2B_OK) return rc;
  }
  return L2B_OK;
}

}  // extern "C"
