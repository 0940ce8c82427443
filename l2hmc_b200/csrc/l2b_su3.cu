// l2b_su3.cu -- sm_100a kernels + C ABI for the 4-D SU(3) leapfrog hot path.
//
// Kernel inventory (DESIGN.md has the roofline of each):
//   k_aos_to_soa / k_soa_to_aos   boundary layout <-> planar layout (+ |P|^2 partials)
//   k_force                       staples -> (beta/3) TAH(U A), fused momentum kick,
//                                 fused Wilson-action and kinetic-energy partial sums
//   k_drift                       U <- exp(eps P) U  (Cayley-Hamilton exponential)
//   k_plaq                        plaquette traces / per-chain sums
//   k_update_gauge, k_project, k_vupdate, ... link-local ops on the boundary layout,
//                                 staged through shared memory so that global traffic is
//                                 contiguous 128-bit accesses while each thread keeps a
//                                 whole 3x3 matrix in registers
//   k_reduce_affine               deterministic second reduction stage
//
// Reference functions each entry point replaces are listed in include/l2b.h.
#include <stdlib.h>
#include <string.h>

#include <cuda.h>
#include <cuda_bf16.h>

#include "l2b_common.cuh"
#include "l2b_su3_site.cuh"

namespace l2b {
namespace {

using T = double;
using C = double2;
constexpr int NTL = 128;       // threads (= links) per block of the link-local kernels
constexpr int PF_AHEAD_SITES = 16384;  // ~ (148 SMs x 3 blocks x 32 sites) + margin

// ---------------------------------------------------------------------------
// shared-memory staging of the interleaved (AoS) layout
// ---------------------------------------------------------------------------
template <int NT, typename E>
__device__ __forceinline__ void stage_in(E* sm, const E* __restrict__ g, size_t first_item, int nitems, int width) {
  const E* src = g + first_item * width;
  const int n = nitems * width;
  for (int k = threadIdx.x; k < n; k += NT) sm[k] = __ldg(src + k);
}
template <int NT, typename E>
__device__ __forceinline__ void stage_out(E* __restrict__ g, const E* sm, size_t first_item, int nitems, int width) {
  E* dst = g + first_item * width;
  const int n = nitems * width;
  for (int k = threadIdx.x; k < n; k += NT) dst[k] = sm[k];
}
__device__ __forceinline__ void sm_get(Mat3<T>& m, const C* sm, int i) {
#pragma unroll
  for (int e = 0; e < 9; ++e) { const C v = sm[i * 9 + e]; m.re[e] = v.x; m.im[e] = v.y; }
}
__device__ __forceinline__ void sm_put(C* sm, int i, const Mat3<T>& m) {
#pragma unroll
  for (int e = 0; e < 9; ++e) sm[i * 9 + e] = make_double2(m.re[e], m.im[e]);
}
// load this thread's matrix out of a contiguous run of `n` AoS matrices
template <int NT>
__device__ __forceinline__ void block_load_mat(Mat3<T>& m, C* sm, const C* g, size_t first, int n) {
  stage_in<NT>(sm, g, first, n, 9);
  __syncthreads();
  if ((int)threadIdx.x < n) sm_get(m, sm, threadIdx.x);
  __syncthreads();
}
template <int NT>
__device__ __forceinline__ void block_store_mat(C* g, C* sm, const Mat3<T>& m, size_t first, int n) {
  if ((int)threadIdx.x < n) sm_put(sm, threadIdx.x, m);
  __syncthreads();
  stage_out<NT>(g, sm, first, n, 9);
  __syncthreads();
}
template <int NT>
__device__ __forceinline__ void block_load_real9(T r[9], T* sm, const T* g, size_t first, int n) {
  if (g == nullptr) {
#pragma unroll
    for (int e = 0; e < 9; ++e) r[e] = T(0);
    return;                                  // uniform across the block
  }
  stage_in<NT>(sm, g, first, n, 9);
  __syncthreads();
  if ((int)threadIdx.x < n) {
#pragma unroll
    for (int e = 0; e < 9; ++e) r[e] = sm[threadIdx.x * 9 + e];
  }
  __syncthreads();
}

// ---------------------------------------------------------------------------
// layout conversion
// ---------------------------------------------------------------------------
template <bool WITH_NORM2>
__global__ void __launch_bounds__(NTL) k_aos_to_soa(const C* __restrict__ aos, C* __restrict__ soa, int V,
                                                    double* __restrict__ part) {
  __shared__ C sm[NTL * 9];
  __shared__ double red[NTL / 32];
  const int plane = blockIdx.y;            // b * 4 + mu
  const int site0 = blockIdx.x * NTL;
  const int n = min(NTL, V - site0);
  stage_in<NTL>(sm, aos, (size_t)plane * V + site0, n, 9);
  __syncthreads();
  double acc = 0.0;
  if ((int)threadIdx.x < n) {
    C* out = soa + (size_t)plane * 9 * V + site0 + threadIdx.x;
#pragma unroll
    for (int e = 0; e < 9; ++e) {
      const C v = sm[threadIdx.x * 9 + e];
      out[(size_t)e * V] = v;
      if (WITH_NORM2) { acc = fma(v.x, v.x, acc); acc = fma(v.y, v.y, acc); }
    }
  }
  if (WITH_NORM2) {
    // per link (|P|_F^2 - 8), as the reference sums it (group.py:125-126)
    if ((int)threadIdx.x < n) acc -= 8.0;
    acc = block_sum<NTL>(acc, red, threadIdx.x);
    if (threadIdx.x == 0) part[(size_t)blockIdx.y * gridDim.x + blockIdx.x] = acc;
  }
}

__global__ void __launch_bounds__(NTL) k_soa_to_aos(const C* __restrict__ soa, C* __restrict__ aos, int V) {
  __shared__ C sm[NTL * 9];
  const int plane = blockIdx.y;
  const int site0 = blockIdx.x * NTL;
  const int n = min(NTL, V - site0);
  if ((int)threadIdx.x < n) {
    const C* in = soa + (size_t)plane * 9 * V + site0 + threadIdx.x;
#pragma unroll
    for (int e = 0; e < 9; ++e) sm[threadIdx.x * 9 + e] = __ldg(in + (size_t)e * V);
  }
  __syncthreads();
  stage_out<NTL>(aos, sm, (size_t)plane * V + site0, n, 9);
}

// ---------------------------------------------------------------------------
// force (+ kick) on the planar layout.  Block = FORCE_TS consecutive sites x 4
// directions: the four warps that share a site neighbourhood run together so the
// 18 neighbour matrices of a link are mostly served by L1.  (TS = 16 packs two
// directions into one warp.)
// ---------------------------------------------------------------------------
// Brick tiling: the TS sites of a block form a compact 4-D brick ((1,2,2,8) for 32 sites,
// (2,2,2,8) for 64) instead of TS consecutive sites (two or four z-rows).  Each z-run of 8
// sites is still one aligned 128-byte segment per matrix entry, so coalescing is unchanged,
// but the brick's surface is smaller: the distinct neighbour matrices a block has to pull
// from L2 drop from 6.4 per link to 5.5 (32 sites) / 4.3 (64 sites).
template <int TS>
__device__ __forceinline__ int brick_site(const Lat& lat, int blk, int tx) {
  constexpr int BT = (TS == 64) ? 2 : 1;
  const int nz = lat.L[3] >> 3, ny = lat.L[2] >> 1, nx = lat.L[1] >> 1;
  int r = blk;
  const int iz = r % nz; r /= nz;
  const int iy = r % ny; r /= ny;
  const int ix = r % nx; r /= nx;
  const int it = r;
  const int z = (iz << 3) + (tx & 7);
  const int y = (iy << 1) + ((tx >> 3) & 1);
  const int x = (ix << 1) + ((tx >> 4) & 1);
  const int t = it * BT + ((TS == 64) ? (tx >> 5) : 0);
  return ((t * lat.L[1] + x) * lat.L[2] + y) * lat.L[3] + z;
}

template <int TS, int MINB, bool KICK, int PF, bool DRIFT, bool BRICK = false>
__global__ void __launch_bounds__(TS * 4, MINB) k_force(const C* __restrict__ U, C* __restrict__ P, Lat lat, double coef,
                                                  double* __restrict__ part, C* __restrict__ Uout,
                                                  double eps_drift) {
  __shared__ double red[TS * 4 / 32];
  const int b = blockIdx.y;
  const int mu = threadIdx.y;
  const int site = BRICK ? brick_site<TS>(lat, blockIdx.x, threadIdx.x) : blockIdx.x * TS + threadIdx.x;
  const int tid = threadIdx.y * TS + threadIdx.x;
  double retr = 0.0, p2 = 0.0;
  if (site < lat.V) {
    Mat3<T> g, f;
    if (PF == 2) {
      // pull this direction's own links / momenta of a block far enough ahead into L2,
      // so that by the time that block runs its first-touch loads hit L2 instead of HBM
      const int ahead = site + PF_AHEAD_SITES;
      if (ahead < lat.V) {
        soa_prefetch<2>(soa_plane(U, lat, b, mu), lat.V, ahead);
        if (KICK) soa_prefetch<2>(soa_plane((const C*)P, lat, b, mu), lat.V, ahead);
      }
    }
    link_times_staples<T, C, (PF == 1 ? 1 : 0)>(g, U, lat, b, mu, site);
    retr = re_trace(g);
    project_tah(f, g);
    C* pp = soa_plane(P, lat, b, mu) + site;
    const size_t V = lat.V;
#pragma unroll
    for (int e = 0; e < 9; ++e) {
      C v;
      if (KICK) {
        v = __ldcs(pp + e * V);          // momenta are touched once per launch: keep them out of
        v.x = fma(-coef, f.re[e], v.x);  // the way of the link matrices in L1/L2 (evict-first)
        v.y = fma(-coef, f.im[e], v.y);
        p2 = fma(v.x, v.x, p2);
        p2 = fma(v.y, v.y, p2);
      } else {
        v.x = coef * f.re[e];
        v.y = coef * f.im[e];
      }
      __stcs(pp + e * V, v);
      if (DRIFT) { f.re[e] = eps_drift * v.x; f.im[e] = eps_drift * v.y; }   // f <- eps * P_new
    }
    if (KICK) p2 -= 8.0;   // per link (|P|_F^2 - 8), group.py:125-126
    if (DRIFT) {
      // fused drift into the OTHER link buffer: U'(mu,n) = exp(eps P_new) U(mu,n).  Neighbours of
      // other threads keep reading the old links from `U`, so no grid-wide sync is needed and a
      // leapfrog step costs 4 field transfers (r U, r P, w P, w U') instead of 6.
      Mat3<T> ex, w, un;
      soa_load(w, soa_plane(U, lat, b, mu), lat.V, site);
      mat_exp_alg(ex, f);
      mat_mul<false, false, false>(un, ex, w);
      soa_store(soa_plane(Uout, lat, b, mu), lat.V, site, un);
    }
  }
  if (part != nullptr) {
    retr = block_sum<TS * 4>(retr, red, tid);
    p2 = block_sum<TS * 4>(p2, red, tid);
    if (tid == 0) {
      double* o = part + ((size_t)b * gridDim.x + blockIdx.x) * 2;
      o[0] = retr;
      o[1] = p2;
    }
  }
}


// "early momentum" variants of the kick kernels.  In k_force the 9 momentum loads are issued right before
// the kick, after ~2000 cycles of staple arithmetic, and their DRAM latency is fully exposed: 25 % of the
// kernel's stall samples (profiles/r1d_force_stall_profile.md).  Here they are issued HOOK_AT directions
// earlier into 18 more live doubles (the register allocator decides what to spill under the same cap).
// Same arithmetic in the same order: results are bit-identical to k_force.
// PIO / UAOS fold the trajectory's layout conversions into its first and last launches (each saves one
// full pass over a field):
//   PIO = 1  the momenta come from the caller's boundary-layout tensor `Paos` ([nb,4,V,3,3], 144 B per link:
//            every thread reads its own 9 contiguous entries -- a warp reads 4.6 KB contiguous, all sectors
//            fully used) and are written planar; the kernel also sums |P_in|^2 - 8 per link (KE of the
//            initial state, group.py:125-126) into a third partial (partials have stride 3);
//   PIO = 2  the momenta are read planar and written to `Paos` only (final half kick -> v_prop);
//   UAOS     the drifted links are written planar AND to the boundary-layout tensor `Uaos` (last drift -> x_prop).
// The arithmetic and its order are those of PIO = 0: same bits.
template <int TS, bool DRIFT, int HOOK_AT, bool LOWREG, int PIO, bool UAOS>
__device__ __forceinline__ void force_ep_body(const C* __restrict__ U, C* __restrict__ P, const Lat& lat, double coef,
                                              double* __restrict__ part, C* __restrict__ Uout, double eps_drift,
                                              C* __restrict__ Paos, C* __restrict__ Uaos) {
  __shared__ double red[TS * 4 / 32];
  const int b = blockIdx.y;
  const int mu = threadIdx.y;
  const int site = blockIdx.x * TS + threadIdx.x;
  const int tid = threadIdx.y * TS + threadIdx.x;
  double retr = 0.0, p2 = 0.0, p2in = 0.0;
  if (site < lat.V) {
    Mat3<T> g, f;
    const int ahead = site + PF_AHEAD_SITES;
    if (ahead < lat.V) {
      soa_prefetch<2>(soa_plane(U, lat, b, mu), lat.V, ahead);
      if (PIO != 1) soa_prefetch<2>(soa_plane((const C*)P, lat, b, mu), lat.V, ahead);
    }
    C* pp = soa_plane(P, lat, b, mu) + site;
    const size_t V = lat.V;
    const size_t aos_link = (((size_t)b * 4 + mu) * V + site) * 9;
    C pv[9];
    auto load_p = [&]() {
#pragma unroll
      for (int e = 0; e < 9; ++e) pv[e] = (PIO == 1) ? __ldcs(Paos + aos_link + e) : __ldcs(pp + e * V);
    };
    if (LOWREG) link_times_staples_lowreg<T, C, HOOK_AT>(g, U, lat, b, mu, site, load_p);   // row-streamed operands
    else link_times_staples_hook<T, C, HOOK_AT>(g, U, lat, b, mu, site, load_p);
    retr = re_trace(g);
    project_tah(f, g);
#pragma unroll
    for (int e = 0; e < 9; ++e) {
      C v = pv[e];
      if (PIO == 1) { p2in = fma(v.x, v.x, p2in); p2in = fma(v.y, v.y, p2in); }
      v.x = fma(-coef, f.re[e], v.x);
      v.y = fma(-coef, f.im[e], v.y);
      p2 = fma(v.x, v.x, p2);
      p2 = fma(v.y, v.y, p2);
      if (PIO == 2) __stcs(Paos + aos_link + e, v);
      else __stcs(pp + e * V, v);
      if (DRIFT) { f.re[e] = eps_drift * v.x; f.im[e] = eps_drift * v.y; }
    }
    p2 -= 8.0;
    if (PIO == 1) p2in -= 8.0;
    if (DRIFT) {
      Mat3<T> ex, w, un;
      soa_load(w, soa_plane(U, lat, b, mu), lat.V, site);
      mat_exp_alg(ex, f);
      mat_mul<false, false, false>(un, ex, w);
      soa_store(soa_plane(Uout, lat, b, mu), lat.V, site, un);
      if (UAOS) {
#pragma unroll
        for (int e = 0; e < 9; ++e) {
          C v;
          v.x = un.re[e]; v.y = un.im[e];
          __stcs(Uaos + aos_link + e, v);
        }
      }
    }
  }
  if (part != nullptr) {
    retr = block_sum<TS * 4>(retr, red, tid);
    p2 = block_sum<TS * 4>(p2, red, tid);
    if (PIO == 1) p2in = block_sum<TS * 4>(p2in, red, tid);
    if (tid == 0) {
      double* o = part + ((size_t)b * gridDim.x + blockIdx.x) * (PIO == 1 ? 3 : 2);
      o[0] = retr;
      o[1] = p2;
      if (PIO == 1) o[2] = p2in;
    }
  }
}

template <int TS, int MINB, bool DRIFT, int HOOK_AT, bool LOWREG = false>
__global__ void __launch_bounds__(TS * 4, MINB) k_force_ep(const C* __restrict__ U, C* __restrict__ P, Lat lat,
                                                           double coef, double* __restrict__ part,
                                                           C* __restrict__ Uout, double eps_drift) {
  force_ep_body<TS, DRIFT, HOOK_AT, LOWREG, 0, false>(U, P, lat, coef, part, Uout, eps_drift, nullptr, nullptr);
}

// first / last launches of a trajectory: momenta in or out of the boundary layout, links out to it
template <int TS, int MINB, bool DRIFT, int HOOK_AT, int PIO, bool UAOS>
__global__ void __launch_bounds__(TS * 4, MINB) k_force_epx(const C* __restrict__ U, C* __restrict__ P, Lat lat,
                                                            double coef, double* __restrict__ part,
                                                            C* __restrict__ Uout, double eps_drift,
                                                            C* __restrict__ Paos, C* __restrict__ Uaos) {
  force_ep_body<TS, DRIFT, HOOK_AT, false, PIO, UAOS>(U, P, lat, coef, part, Uout, eps_drift, Paos, Uaos);
}

// ---------------------------------------------------------------------------
// k_force_tma: the fused step with the block's OWN momenta and links -- the two first-touch, DRAM-latency
// loads of a link's update, (9 + 9) LDG.128 per thread that k_force_ep keeps in registers from the last
// staple direction on -- brought in by the TMA instead: one elected thread issues eight
// `cp.async.bulk.tensor.2d` (SASS UTMALDG) for the [9 entries x 32 sites] boxes of P_mu and U_mu, mu = 0..3,
// at kernel entry; they land in shared memory ([field][mu][e][site], 36 KB) behind an mbarrier while all
// four warps multiply staples, and are read back with conflict-free LDS.128 right before the kick.
// The planar field is a rank-2 tensor for the TMA: inner dimension = the 2 V doubles of one (chain, mu, e)
// row, outer = the nb * 36 rows.  Same arithmetic in the same order as k_force_ep: same bits.
// Variant 32 (`su3_force_variant`); needs V % 32 == 0, otherwise variant 22 runs.
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
      ::"r"(smem_addr(dst)), "l"((uint64_t)map), "r"(c0), "r"(c1), "r"(smem_addr(bar))
      : "memory");
}

template <int MINB, bool DRIFT>
__global__ void __launch_bounds__(128, MINB) k_force_tma(const __grid_constant__ CUtensorMap mapP,
                                                         const __grid_constant__ CUtensorMap mapU,
                                                         const C* __restrict__ U, C* __restrict__ P, Lat lat,
                                                         double coef, double* __restrict__ part,
                                                         C* __restrict__ Uout, double eps_drift) {
  constexpr int TS = 32;
  extern __shared__ __align__(128) unsigned char tma_smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ double red[4];
  C* tileP = reinterpret_cast<C*>(tma_smem);            // [mu][e][site]
  C* tileU = tileP + 4 * 9 * TS;
  const int b = blockIdx.y;
  const int mu = threadIdx.y;
  const int site0 = blockIdx.x * TS;
  const int site = site0 + threadIdx.x;
  const int tid = threadIdx.y * TS + threadIdx.x;
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_addr(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (tid == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(&bar)),
                 "r"((uint32_t)(2 * 4 * 9 * TS * sizeof(C)))
                 : "memory");
#pragma unroll
    for (int m = 0; m < 4; ++m) {
      tma_load_2d(tileP + m * 9 * TS, &mapP, 2 * site0, (b * 4 + m) * 9, &bar);
      tma_load_2d(tileU + m * 9 * TS, &mapU, 2 * site0, (b * 4 + m) * 9, &bar);
    }
  }
  double retr = 0.0, p2 = 0.0;
  {
    Mat3<T> a, g, f, w;
    const int ahead = site + PF_AHEAD_SITES;
    if (ahead < lat.V) soa_prefetch<2>(soa_plane(U, lat, b, mu), lat.V, ahead);   // neighbours' L2 hits live on this
    link_times_staples<T, C, 0, false>(a, U, lat, b, mu, site);                   // A = sum of the six staples
    // the tiles have had the whole staple phase to land
    {
      uint32_t ok = 0;
      for (uint32_t spin = 0; !ok; ++spin) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(smem_addr(&bar))
            : "memory");
        if (!ok && spin > (1u << 24)) __trap();
      }
    }
    const C* tu = tileU + (mu * 9) * TS + threadIdx.x;
    const C* tp = tileP + (mu * 9) * TS + threadIdx.x;
#pragma unroll
    for (int e = 0; e < 9; ++e) { const C v = tu[e * TS]; w.re[e] = v.x; w.im[e] = v.y; }
    mat_mul<false, false, false>(g, w, a);
    retr = re_trace(g);
    project_tah(f, g);
    C* pp = soa_plane(P, lat, b, mu) + site;
    const size_t V = lat.V;
#pragma unroll
    for (int e = 0; e < 9; ++e) {
      C v = tp[e * TS];
      v.x = fma(-coef, f.re[e], v.x);
      v.y = fma(-coef, f.im[e], v.y);
      p2 = fma(v.x, v.x, p2);
      p2 = fma(v.y, v.y, p2);
      __stcs(pp + e * V, v);
      if (DRIFT) { f.re[e] = eps_drift * v.x; f.im[e] = eps_drift * v.y; }
    }
    p2 -= 8.0;
    if (DRIFT) {
      Mat3<T> ex, un;
      mat_exp_alg(ex, f);
      mat_mul<false, false, false>(un, ex, w);
      soa_store(soa_plane(Uout, lat, b, mu), lat.V, site, un);
    }
  }
  if (part != nullptr) {
    retr = block_sum<128>(retr, red, tid);
    p2 = block_sum<128>(p2, red, tid);
    if (tid == 0) {
      double* o = part + ((size_t)b * gridDim.x + blockIdx.x) * 2;
      o[0] = retr;
      o[1] = p2;
    }
  }
}

// ---------------------------------------------------------------------------
// k_force_async: same arithmetic as k_force, but the 18 neighbour matrices of a
// link travel global -> shared memory with cp.async (LDGSTS) through a per-thread
// ring of R slots, issued R-1 operands ahead of their use.  In-flight loads then
// live in the otherwise idle shared memory instead of in registers, which is what
// caps k_force (its stalls are ~65 % long-scoreboard at 12 warps/SM).  Every
// thread copies exactly the 16-byte words it will read itself, so cp.async
// wait_group is the only synchronisation needed; slot layout [slot][e][thread]
// keeps both the asynchronous writes and the LDS.128 reads conflict free.
// ---------------------------------------------------------------------------
__device__ __forceinline__ void cp_async16(void* smem, const void* gptr) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16;\n" ::"r"(sa), "l"(gptr) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

template <int R, int MINB, bool KICK>
__global__ void __launch_bounds__(128, MINB) k_force_async(const C* __restrict__ U, C* __restrict__ P, Lat lat,
                                                           double coef, double* __restrict__ part, C*, double) {
  extern __shared__ __align__(16) unsigned char smraw[];
  __shared__ double red[4];
  C* ring = reinterpret_cast<C*>(smraw);          // [R][9][128]
  const int b = blockIdx.y;
  const int mu = threadIdx.y;
  const int tid = threadIdx.y * 32 + threadIdx.x;
  int site = blockIdx.x * 32 + threadIdx.x;
  const bool live = site < lat.V;
  if (!live) site = lat.V - 1;                    // keep the pipeline uniform; results discarded
  const int V = lat.V;
  int r = site;
  const int c3 = r % lat.L[3]; r /= lat.L[3];
  const int c2 = r % lat.L[2]; r /= lat.L[2];
  const int c1 = r % lat.L[1]; r /= lat.L[1];
  const int c0 = r;
  const int f0 = (c0 == lat.L[0] - 1) ? -(lat.L[0] - 1) * lat.stride[0] : lat.stride[0];
  const int f1 = (c1 == lat.L[1] - 1) ? -(lat.L[1] - 1) * lat.stride[1] : lat.stride[1];
  const int f2 = (c2 == lat.L[2] - 1) ? -(lat.L[2] - 1) * lat.stride[2] : lat.stride[2];
  const int f3 = (c3 == lat.L[3] - 1) ? -(lat.L[3] - 1) : 1;
  const int b0 = (c0 == 0) ? (lat.L[0] - 1) * lat.stride[0] : -lat.stride[0];
  const int b1 = (c1 == 0) ? (lat.L[1] - 1) * lat.stride[1] : -lat.stride[1];
  const int b2 = (c2 == 0) ? (lat.L[2] - 1) * lat.stride[2] : -lat.stride[2];
  const int b3 = (c3 == 0) ? (lat.L[3] - 1) : -1;
  const size_t plane_sz = (size_t)9 * V;
  const C* chain = U + (size_t)b * 4 * plane_sz;
  const C* pmu = chain + (size_t)mu * plane_sz;
  const int n_pmu = site + sel4(f0, f1, f2, f3, mu);

  // operand j = 6*(k-1) + w, k = 1..3 (nu = mu + k), w: 0 X1=U_nu(n+mu) 1 Y1=U_mu(n+nu) 2 Z1=U_nu(n)
  //                                                     3 X2=U_nu(n+mu-nu) 4 Y2=U_mu(n-nu) 5 Z2=U_nu(n-nu); j = 18: U_mu(n)
  auto issue = [&](int j) {
    const C* src;
    if (j >= 18) {
      src = pmu + site;
    } else {
      const int k = j / 6 + 1, w = j % 6;
      const int nu = (mu + k) & 3;
      const C* pnu = chain + (size_t)nu * plane_sz;
      const int fnu = sel4(f0, f1, f2, f3, nu), bnu = sel4(b0, b1, b2, b3, nu);
      src = (w == 0) ? pnu + n_pmu : (w == 1) ? pmu + site + fnu : (w == 2) ? pnu + site
          : (w == 3) ? pnu + n_pmu + bnu : (w == 4) ? pmu + site + bnu : pnu + site + bnu;
    }
    C* dst = ring + ((size_t)(j % R) * 9) * 128 + tid;
#pragma unroll
    for (int e = 0; e < 9; ++e) cp_async16(dst + e * 128, src + (size_t)e * V);
  };
  auto take = [&](Mat3<T>& m, int j) {
    const C* src = ring + ((size_t)(j % R) * 9) * 128 + tid;
#pragma unroll
    for (int e = 0; e < 9; ++e) { const C v = src[e * 128]; m.re[e] = v.x; m.im[e] = v.y; }
  };

#pragma unroll
  for (int j = 0; j < R - 1; ++j) { issue(j); cp_async_commit(); }
  Mat3<T> a, x, y, m;
  mat_zero(a);
#pragma unroll
  for (int j = 0; j < 19; ++j) {
    if (j + R - 1 < 19) issue(j + R - 1);
    cp_async_commit();                           // one group per step (possibly empty) keeps the count uniform
    cp_async_wait<R - 1>();                      // operand j has landed
    const int w = j % 6;
    if (j == 18) {
      take(x, j);
    } else if (w == 0 || w == 3) {
      take(x, j);
    } else if (w == 1) {
      take(y, j);
      mat_mul<false, true, false>(m, x, y);      // U_nu(n+mu) U_mu(n+nu)^+
    } else if (w == 4) {
      take(y, j);
      mat_mul<true, true, false>(m, x, y);       // U_nu(n+mu-nu)^+ U_mu(n-nu)^+
    } else if (w == 2) {
      take(x, j);
      mat_mul<false, true, true>(a, m, x);       // += m U_nu(n)^+
    } else {
      take(x, j);
      mat_mul<false, false, true>(a, m, x);      // += m U_nu(n-nu)
    }
  }
  Mat3<T> g, f;
  mat_mul<false, false, false>(g, x, a);
  double retr = 0.0, p2 = 0.0;
  if (live) {
    retr = re_trace(g);
    project_tah(f, g);
    C* pp = soa_plane(P, lat, b, mu) + site;
    const size_t Vs = lat.V;
#pragma unroll
    for (int e = 0; e < 9; ++e) {
      C v;
      if (KICK) {
        v = pp[e * Vs];
        v.x = fma(-coef, f.re[e], v.x);
        v.y = fma(-coef, f.im[e], v.y);
        p2 = fma(v.x, v.x, p2);
        p2 = fma(v.y, v.y, p2);
      } else {
        v.x = coef * f.re[e];
        v.y = coef * f.im[e];
      }
      pp[e * Vs] = v;
    }
    if (KICK) p2 -= 8.0;
  }
  if (part != nullptr) {
    retr = block_sum<128>(retr, red, tid);
    p2 = block_sum<128>(p2, red, tid);
    if (tid == 0) {
      double* o = part + ((size_t)b * gridDim.x + blockIdx.x) * 2;
      o[0] = retr;
      o[1] = p2;
    }
  }
}

// ---------------------------------------------------------------------------
// adjoint kernels of the SU(3) L2HMC path (training)
// ---------------------------------------------------------------------------
// gx(mu,n) = coef[b] * A_mu(n)^+  : adjoint of the Wilson action, dS/dU = -(beta/3) A^+
// (the staple sum of the force kernel without the final link product), planar output
// With GF != nullptr (adjoint of the force at fixed dsdx):
//   gx(mu,n) = TAH(GF(mu,n))^+ * (scale * A_mu(n)^+),  scale = -(beta/3)
template <int TS>
__global__ void __launch_bounds__(TS * 4, 3) k_action_grad(const C* __restrict__ U, C* __restrict__ G, Lat lat,
                                                           const double* __restrict__ coef, double scale,
                                                           const C* __restrict__ GF) {
  const int b = blockIdx.y;
  const int mu = threadIdx.y;
  const int site = blockIdx.x * TS + threadIdx.x;
  if (site >= lat.V) return;
  Mat3<T> a, ah;
  link_times_staples<T, C, 0, false>(a, U, lat, b, mu, site);
  const double c = coef ? coef[b] : scale;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
#pragma unroll
    for (int j = 0; j < 3; ++j) { ah.re[3 * i + j] = c * a.re[3 * j + i]; ah.im[3 * i + j] = -c * a.im[3 * j + i]; }
  }
  if (GF != nullptr) {
    Mat3<T> gf, th;
    soa_load(gf, soa_plane(GF, lat, b, mu), lat.V, site);
    project_tah(th, gf);
    mat_mul<true, false, false>(a, th, ah);
    ah = a;
  }
  soa_store(soa_plane(G, lat, b, mu), lat.V, site, ah);
}

// Improved action (c1 != 0): force = coef * TAH(U [(1 - 8 c1) A + c1 R]) with the 18
// rectangle staples per link (rect_staples, l2b_su3_site.cuh); per-block partials (sum Re tr(U A), sum Re tr(U R)).
// 91 more matrix loads and 72 more products per link than the plaquette force: FP64-bound, not tuned
// (the rectangle term is off in every shipped config of the reference).
template <int TS>
__global__ void __launch_bounds__(TS * 4, 2) k_force_c1(const C* __restrict__ U, C* __restrict__ F, Lat lat, double coef,
                                                        double c1, double* __restrict__ part) {
  __shared__ double red[TS * 4 / 32];
  const int b = blockIdx.y;
  const int mu = threadIdx.y;
  const int site = blockIdx.x * TS + threadIdx.x;
  const int tid = threadIdx.y * TS + threadIdx.x;
  double rp = 0.0, rr = 0.0;
  if (site < lat.V) {
    Mat3<T> g, f;
    link_times_improved_staples<T, C>(g, rp, rr, U, lat, b, mu, site, c1);
    if (F != nullptr) {
      project_tah(f, g);
#pragma unroll
      for (int e = 0; e < 9; ++e) { f.re[e] *= coef; f.im[e] *= coef; }
      soa_store(soa_plane(F, lat, b, mu), lat.V, site, f);
    }
  }
  if (part != nullptr) {
    rp = block_sum<TS * 4>(rp, red, tid);
    rr = block_sum<TS * 4>(rr, red, tid);
    if (tid == 0) {
      double* o = part + ((size_t)b * gridDim.x + blockIdx.x) * 2;
      o[0] = rp;
      o[1] = rr;
    }
  }
}

// adjoint of the improved action (GF == nullptr: gx = coef[b] Aimp^+) and of the improved force at fixed dsdx
// (GF != nullptr: gx = TAH(GF)^+ (scale Aimp^+)); Aimp = (1 - 8 c1) A + c1 R.  c1 != 0 counterpart of k_action_grad.
template <int TS>
__global__ void __launch_bounds__(TS * 4, 2) k_action_grad_c1(const C* __restrict__ U, C* __restrict__ G, Lat lat,
                                                              const double* __restrict__ coef, double scale, double c1,
                                                              const C* __restrict__ GF) {
  const int b = blockIdx.y;
  const int mu = threadIdx.y;
  const int site = blockIdx.x * TS + threadIdx.x;
  if (site >= lat.V) return;
  Mat3<T> g, gf;
  if (GF != nullptr) soa_load(gf, soa_plane(GF, lat, b, mu), lat.V, site);
  improved_action_adjoint_link<T, C>(g, U, GF != nullptr ? &gf : nullptr, lat, b, mu, site, coef ? coef[b] : scale, c1);
  soa_store(soa_plane(G, lat, b, mu), lat.V, site, g);
}

// adjoint of k_plaq<WRITE_LOOPS>: gx(mu,n) from the cotangent of the per-site loops
template <int TS>
__global__ void __launch_bounds__(TS * 4, 3) k_wloops_bwd(const C* __restrict__ U, const C* __restrict__ gw,
                                                          C* __restrict__ G, Lat lat, int nb) {
  const int b = blockIdx.y;
  const int mu = threadIdx.y;
  const int site = blockIdx.x * TS + threadIdx.x;
  if (site >= lat.V) return;
  Mat3<T> g;
  wloops_adjoint_link<T, C>(g, U, gw, lat, nb, b, mu, site);
  soa_store(soa_plane(G, lat, b, mu), lat.V, site, g);
}

// adjoint of k_vupdate (force is a constant of the graph)
__global__ void __launch_bounds__(NTL) k_vupdate_bwd(const C* __restrict__ v, const C* __restrict__ f,
                                                     const T* __restrict__ s, const T* __restrict__ t,
                                                     const T* __restrict__ q, double eps_in, const double* __restrict__ eps_dev, int sign,
                                                     const C* __restrict__ gout, const double* __restrict__ glogdet,
                                                     C* __restrict__ gv, C* __restrict__ gf, T* __restrict__ gs,
                                                     T* __restrict__ gt, T* __restrict__ gq, double* __restrict__ part,
                                                     size_t links_per_chain) {
  const double eps = eps_dev ? eps_in * eps_dev[0] : eps_in;   // device-resident step size (CUDA graphs)
  __shared__ C sm[NTL * 9];
  __shared__ double red[NTL / 32];
  const size_t row0 = (size_t)blockIdx.y * links_per_chain;
  const size_t l0 = (size_t)blockIdx.x * NTL;
  const int n = (int)min((size_t)NTL, links_per_chain - l0);
  Mat3<T> Vm, Fm, Gm, R;
  T sv[9], tv[9], qv[9];
  block_load_mat<NTL>(Vm, sm, v, row0 + l0, n);
  block_load_mat<NTL>(Fm, sm, f, row0 + l0, n);
  block_load_mat<NTL>(Gm, sm, gout, row0 + l0, n);
  T* smr = reinterpret_cast<T*>(sm);
  block_load_real9<NTL>(sv, smr, s, row0 + l0, n);
  block_load_real9<NTL>(tv, smr, t, row0 + l0, n);
  block_load_real9<NTL>(qv, smr, q, row0 + l0, n);
  const double gl = glogdet ? glogdet[blockIdx.y] : 0.0;
  const T sg = (T)sign, he = T(0.5) * eps;
  double ge = 0.0;
  const bool live = (int)threadIdx.x < n;
  if (live) {
#pragma unroll
    for (int e = 0; e < 9; ++e) {
      const T lj = sg * eps * sv[e] / T(2);
      const T es = exp(lj), eq = exp(eps * qv[e]);
      const T fr = fma(Fm.re[e], eq, tv[e]), fi = Fm.im[e] * eq;
      const T a = Gm.re[e], b = Gm.im[e];
      T g_es, g_fr, g_fi;
      if (sign > 0) {
        g_es = a * Vm.re[e] + b * Vm.im[e];
        g_fr = -he * a; g_fi = -he * b;
        ge += -T(0.5) * (fr * a + fi * b);
      } else {
        g_es = a * (Vm.re[e] + he * fr) + b * (Vm.im[e] + he * fi);
        g_fr = es * he * a; g_fi = es * he * b;
        ge += T(0.5) * es * (fr * a + fi * b);
      }
      R.re[e] = es * a; R.im[e] = es * b;
      Gm.re[e] = g_fr * eq; Gm.im[e] = g_fi * eq;   // d/dF (the reference's force keeps its `@ x^+` factor attached)
      const T g_lj = g_es * es + gl;
      ge += g_lj * sg * sv[e] / T(2);
      const T g_eq = g_fr * Fm.re[e] + g_fi * Fm.im[e];
      ge += g_eq * eq * qv[e];
      sv[e] = g_lj * sg * eps / T(2);     // reuse the registers for the outputs
      tv[e] = g_fr;
      qv[e] = g_eq * eq * eps;
    }
  }
  block_store_mat<NTL>(gv, sm, R, row0 + l0, n);
  if (gf != nullptr) block_store_mat<NTL>(gf, sm, Gm, row0 + l0, n);
  // real outputs: stage 9 values per link through shared memory for contiguous stores
  T* outs[3] = {gs, gt, gq};
  for (int k = 0; k < 3; ++k) {
    if (outs[k] == nullptr) continue;
    if (live) {
#pragma unroll
      for (int e = 0; e < 9; ++e) smr[threadIdx.x * 9 + e] = (k == 0) ? sv[e] : (k == 1) ? tv[e] : qv[e];
    }
    __syncthreads();
    stage_out<NTL>(outs[k], smr, row0 + l0, n, 9);
    __syncthreads();
  }
  ge = block_sum<NTL>(ge, red, threadIdx.x);
  if (threadIdx.x == 0) part[(size_t)blockIdx.y * gridDim.x + blockIdx.x] = ge;
}

// adjoint of k_update_gauge:  R = m*X + E ((1-m)*X),  E = exp(eps P)
//   G_X = m*G + (1-m)*(E^+ G),  G_E = G ((1-m)*X)^+,  G_P = eps * expadj(eps P; G_E),
//   g_eps = Re sum conj(G_A) P
// Two phases, so that the matrix exponential and the exponential's adjoint never hold their registers at the same
// time: G_X leaves first; the adjoint then runs column by column on (P^+, G_E) and puts each finished column of
// G_P straight into the block's staging buffer.
template <int MINB>
__global__ void __launch_bounds__(NTL, MINB) k_update_gauge_bwd(const C* __restrict__ x, const C* __restrict__ p, double eps_in, const double* __restrict__ eps_dev,
                                                          const float* __restrict__ mask, int mask_complement,
                                                          const C* __restrict__ gout, C* __restrict__ gx,
                                                          C* __restrict__ gp, double* __restrict__ part,
                                                          int* __restrict__ bad, size_t links_per_chain) {
  const double eps = eps_dev ? eps_in * eps_dev[0] : eps_in;   // device-resident step size (CUDA graphs)
  __shared__ C sm[NTL * 9];
  __shared__ double red[NTL / 32];
  const size_t row0 = (size_t)blockIdx.y * links_per_chain;
  const size_t l0 = (size_t)blockIdx.x * NTL;
  const int n = (int)min((size_t)NTL, links_per_chain - l0);
  const bool live = (int)threadIdx.x < n;
  Mat3<T> Pb, GE;       // P^+ and G_E: all the second phase needs
  {
    Mat3<T> X, Pm, G, E, Xb, GX;
    block_load_mat<NTL>(Pm, sm, p, row0 + l0, n);
    block_load_mat<NTL>(X, sm, x, row0 + l0, n);
    block_load_mat<NTL>(G, sm, gout, row0 + l0, n);
    if (live) {
      Mat3<T> A;
#pragma unroll
      for (int e = 0; e < 9; ++e) { A.re[e] = eps * Pm.re[e]; A.im[e] = eps * Pm.im[e]; }
      mat_exp(E, A);
      T md[9];
#pragma unroll
      for (int e = 0; e < 9; ++e) {
        float m = (mask == nullptr) ? 0.0f : __ldg(mask + (l0 + threadIdx.x) * 9 + e);
        if (mask != nullptr && mask_complement) m = 1.0f - m;
        md[e] = (T)m;
        const T mb = (T)(1.0f - m);
        Xb.re[e] = mb * X.re[e]; Xb.im[e] = mb * X.im[e];
      }
      mat_mul<true, false, false>(GX, E, G);          // E^+ G
      mat_mul<false, true, false>(GE, G, Xb);         // G Xb^+
#pragma unroll
      for (int e = 0; e < 9; ++e) {
        const T mb = T(1) - md[e];
        GX.re[e] = md[e] * G.re[e] + mb * GX.re[e];
        GX.im[e] = md[e] * G.im[e] + mb * GX.im[e];
      }
#pragma unroll
      for (int i = 0; i < 3; ++i) {
#pragma unroll
        for (int j = 0; j < 3; ++j) { Pb.re[3 * i + j] = Pm.re[3 * j + i]; Pb.im[3 * i + j] = -Pm.im[3 * j + i]; }
      }
    }
    block_store_mat<NTL>(gx, sm, GX, row0 + l0, n);   // ends with a barrier: `sm` is free again
  }
  double ge = 0.0;
  if (live) {
    bool ok;
    C* mine = sm + threadIdx.x * 9;
    mat_exp_adjoint_cols(Pb, (T)eps, GE, ok, [&](int c, const T r[3], const T i[3]) {
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        // P(k, c) = conj(Pb(c, k))
        ge = fma(r[k], Pb.re[3 * c + k], ge);
        ge = fma(-i[k], Pb.im[3 * c + k], ge);
        mine[3 * k + c] = make_double2(eps * r[k], eps * i[k]);
      }
    });
    if (!ok) atomicExch(bad, 1);
  }
  __syncthreads();
  stage_out<NTL>(gp, sm, row0 + l0, n, 9);
  ge = block_sum<NTL>(ge, red, threadIdx.x);
  if (threadIdx.x == 0) part[(size_t)blockIdx.y * gridDim.x + blockIdx.x] = ge;
}

__global__ void __launch_bounds__(NTL) k_to_vec_bwd(const T* __restrict__ gvec8, C* __restrict__ gx, size_t nmat) {
  __shared__ C sm[NTL * 9];
  const size_t first = (size_t)blockIdx.x * NTL;
  const int n = (int)min((size_t)NTL, nmat - first);
  Mat3<T> m;
  if ((int)threadIdx.x < n) {
    T v[8];
    const double2* in = reinterpret_cast<const double2*>(gvec8 + (first + threadIdx.x) * 8);
#pragma unroll
    for (int k = 0; k < 4; ++k) { const double2 w = __ldg(in + k); v[2 * k] = w.x; v[2 * k + 1] = w.y; }
    su3_to_vec_adjoint(m, v);
  }
  block_store_mat<NTL>(gx, sm, m, first, n);
}

// 8 reals per link in the element type of the nets (vnet input packing): f64 / f32 / bf16
template <typename V> struct Vec8IO;
template <> struct Vec8IO<double> {
  static __device__ __forceinline__ void store(double* o, const T v[8]) {
    double2* d = reinterpret_cast<double2*>(o);
#pragma unroll
    for (int k = 0; k < 4; ++k) d[k] = make_double2(v[2 * k], v[2 * k + 1]);
  }
  static __device__ __forceinline__ void load(T v[8], const double* i) {
    const double2* d = reinterpret_cast<const double2*>(i);
#pragma unroll
    for (int k = 0; k < 4; ++k) { const double2 w = __ldg(d + k); v[2 * k] = w.x; v[2 * k + 1] = w.y; }
  }
};
template <> struct Vec8IO<float> {
  static __device__ __forceinline__ void store(float* o, const T v[8]) {
    float4* d = reinterpret_cast<float4*>(o);
    d[0] = make_float4((float)v[0], (float)v[1], (float)v[2], (float)v[3]);
    d[1] = make_float4((float)v[4], (float)v[5], (float)v[6], (float)v[7]);
  }
  static __device__ __forceinline__ void load(T v[8], const float* i) {
    const float4* d = reinterpret_cast<const float4*>(i);
    const float4 a = __ldg(d), b = __ldg(d + 1);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  }
};
template <> struct Vec8IO<__nv_bfloat16> {
  static __device__ __forceinline__ void store(__nv_bfloat16* o, const T v[8]) {
    __align__(16) __nv_bfloat162 h[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) h[k] = __floats2bfloat162_rn((float)v[2 * k], (float)v[2 * k + 1]);
    *reinterpret_cast<uint4*>(o) = *reinterpret_cast<const uint4*>(h);
  }
  static __device__ __forceinline__ void load(T v[8], const __nv_bfloat16* i) {
    __align__(16) __nv_bfloat162 h[4];
    *reinterpret_cast<uint4*>(h) = __ldg(reinterpret_cast<const uint4*>(i));
#pragma unroll
    for (int k = 0; k < 4; ++k) { const float2 f = __bfloat1622float2(h[k]); v[2 * k] = f.x; v[2 * k + 1] = f.y; }
  }
};

template <typename V> constexpr bool kVecIsBf16 = false;
template <> constexpr bool kVecIsBf16<__nv_bfloat16> = true;

// vec8 = su3_to_vec(projectSU(x)) written straight in the nets' element type
template <typename V>
__global__ void __launch_bounds__(NTL) k_project_vec(const C* __restrict__ x, V* __restrict__ vec8, size_t nmat) {
  __shared__ C sm[NTL * 9];
  const size_t first = (size_t)blockIdx.x * NTL;
  const int n = (int)min((size_t)NTL, nmat - first);
  Mat3<T> m, r;
  block_load_mat<NTL>(m, sm, x, first, n);
  if ((int)threadIdx.x < n) {
    project_su<T, kVecIsBf16<V>>(r, m);               // bf16 output: links already in SU(3) pass through (see project_su)
    T v[8];
    su3_to_vec(v, r);
    Vec8IO<V>::store(vec8 + (first + threadIdx.x) * 8, v);
  }
}

// adjoint of y = projectSU(x) (cotangent gmat) and/or vec8 = su3_to_vec(projectSU(x))
// (cotangent gvec8): closed form, project_su_adjoint in l2b_su3_math.cuh
template <typename V>
__global__ void __launch_bounds__(NTL) k_project_bwd(const C* __restrict__ x, const C* __restrict__ gmat,
                                                     const V* __restrict__ gvec8, C* __restrict__ gx, size_t nmat) {
  __shared__ C sm[NTL * 9];
  const size_t first = (size_t)blockIdx.x * NTL;
  const int n = (int)min((size_t)NTL, nmat - first);
  Mat3<T> m, g, r;
  block_load_mat<NTL>(m, sm, x, first, n);
  if (gmat != nullptr) block_load_mat<NTL>(g, sm, gmat, first, n);
  else mat_zero(g);
  if ((int)threadIdx.x < n) {
    if (gvec8 != nullptr) {
      T v[8];
      Vec8IO<V>::load(v, gvec8 + (first + threadIdx.x) * 8);
      Mat3<T> gv;
      su3_to_vec_adjoint(gv, v);
#pragma unroll
      for (int e = 0; e < 9; ++e) { g.re[e] += gv.re[e]; g.im[e] += gv.im[e]; }
    }
    project_su_adjoint(r, m, g);
  }
  block_store_mat<NTL>(gx, sm, r, first, n);
}

// ---------------------------------------------------------------------------
// planar (internal layout) variants of the link-local L2HMC kernels: with the state kept planar
// through a whole L2HMC sweep no layout conversion surrounds the stencil kernels.  One thread per
// link, grid (ceil(V / NTL), nb * 4); every access is a coalesced 128-bit load / store.
// ---------------------------------------------------------------------------
template <typename V>
__global__ void __launch_bounds__(NTL) k_project_vec_planar(const C* __restrict__ x, V* __restrict__ vec8, int Vs) {
  const int site = blockIdx.x * NTL + threadIdx.x;
  if (site >= Vs) return;
  const size_t plane = blockIdx.y;                        // b * 4 + mu
  Mat3<T> m, r;
  soa_load(m, x + plane * 9 * (size_t)Vs, Vs, site);
  project_su<T, kVecIsBf16<V>>(r, m);
  T v[8];
  su3_to_vec(v, r);
  Vec8IO<V>::store(vec8 + (plane * Vs + site) * 8, v);    // [b][mu][site][8]: the order the vnet input expects
}

// Both masked link updates of one leapfrog layer in one pass (dynamics.py:1195-1198 / 1217-1220: `_update_x_fwd(m)`
// then `_update_x_fwd(1 - m)`, no momentum update in between, same step size): E = exp(eps p) is formed once,
//   x1 = m (.) x + E ((1 - m) (.) x),     x2 = (1 - m) (.) x1 + E (m (.) x1)
// (mask_complement = 1 swaps the roles: the backward layer applies 1 - m first).  Same arithmetic per update as
// k_update_gauge_planar; one read of x and p and one write instead of two of each, one exponential instead of two.
__global__ void __launch_bounds__(NTL) k_update_gauge_planar_pair(const C* __restrict__ x, const C* __restrict__ p,
                                                                  double eps_in, const double* __restrict__ eps_dev,
                                                                  const float* __restrict__ mask, int mask_complement,
                                                                  C* __restrict__ out, int Vs) {
  const double eps = eps_dev ? eps_in * eps_dev[0] : eps_in;
  const int site = blockIdx.x * NTL + threadIdx.x;
  if (site >= Vs) return;
  const size_t plane = blockIdx.y;
  const int mu = (int)(plane & 3);
  Mat3<T> X, Pm, E, R, Xb;
  soa_load(Pm, p + plane * 9 * (size_t)Vs, Vs, site);
  soa_load(X, x + plane * 9 * (size_t)Vs, Vs, site);
#pragma unroll
  for (int e = 0; e < 9; ++e) { Pm.re[e] *= eps; Pm.im[e] *= eps; }
  mat_exp(E, Pm);
  float mk[9];
  const float* mp = mask + (size_t)mu * 9 * Vs + site;
#pragma unroll
  for (int e = 0; e < 9; ++e) {
    const float m = __ldg(mp + (size_t)e * Vs);
    mk[e] = mask_complement ? 1.0f - m : m;
  }
#pragma unroll
  for (int pass = 0; pass < 2; ++pass) {
#pragma unroll
    for (int e = 0; e < 9; ++e) {
      const float m = pass == 0 ? mk[e] : 1.0f - mk[e];
      const T md = (T)m, mb = (T)(1.0f - m);
      R.re[e] = md * X.re[e]; R.im[e] = md * X.im[e];
      Xb.re[e] = mb * X.re[e]; Xb.im[e] = mb * X.im[e];
    }
    mat_mul<false, false, true>(R, E, Xb);
    X = R;
  }
  soa_store(out + plane * 9 * (size_t)Vs, Vs, site, X);
}

// same, written as the K-major operand image of the tensor-core input layer (k_su3_input_gemm, l2b_vnet.cu):
// one link is one K core of 8 reals, so the image is link-major, vec8[link][chain < nbp][8] bf16 (rows >= nb stay
// as the caller zeroed them).  A warp's 32 sites store 32 separate 16-byte pieces (stride nbp * 16 B); the 144-byte
// reads stay coalesced, which is the side that matters.
__global__ void __launch_bounds__(NTL) k_project_vec_planar_lm(const C* __restrict__ x, __nv_bfloat16* __restrict__ vec8,
                                                               int Vs, int nbp) {
  const int site = blockIdx.x * NTL + threadIdx.x;
  if (site >= Vs) return;
  const size_t plane = blockIdx.y;                        // b * 4 + mu
  const int b = (int)(plane >> 2), mu = (int)(plane & 3);
  Mat3<T> m, r;
  soa_load(m, x + plane * 9 * (size_t)Vs, Vs, site);
  project_su<T, true>(r, m);
  T v[8];
  su3_to_vec(v, r);
  Vec8IO<__nv_bfloat16>::store(vec8 + (((size_t)mu * Vs + site) * nbp + b) * 8, v);
}

// x' = m*x + exp(eps p) ((1-m)*x), planar x, p, x'; mask planar [4][9][V] float (nullptr: x' = exp(eps p) x)
__global__ void __launch_bounds__(NTL) k_update_gauge_planar(const C* __restrict__ x, const C* __restrict__ p,
                                                             double eps_in, const double* __restrict__ eps_dev,
                                                             const float* __restrict__ mask, int mask_complement,
                                                             C* __restrict__ out, int Vs) {
  const double eps = eps_dev ? eps_in * eps_dev[0] : eps_in;
  const int site = blockIdx.x * NTL + threadIdx.x;
  if (site >= Vs) return;
  const size_t plane = blockIdx.y;
  const int mu = (int)(plane & 3);
  Mat3<T> X, Pm, E, R;
  soa_load(Pm, p + plane * 9 * (size_t)Vs, Vs, site);
  soa_load(X, x + plane * 9 * (size_t)Vs, Vs, site);
#pragma unroll
  for (int e = 0; e < 9; ++e) { Pm.re[e] *= eps; Pm.im[e] *= eps; }
  mat_exp(E, Pm);
  if (mask == nullptr) {
    mat_mul<false, false, false>(R, E, X);
  } else {
    Mat3<T> Xb;
    const float* mk = mask + (size_t)mu * 9 * Vs + site;
#pragma unroll
    for (int e = 0; e < 9; ++e) {
      float m = __ldg(mk + (size_t)e * Vs);
      if (mask_complement) m = 1.0f - m;
      const T md = (T)m, mb = (T)(1.0f - m);
      R.re[e] = md * X.re[e]; R.im[e] = md * X.im[e];
      Xb.re[e] = mb * X.re[e]; Xb.im[e] = mb * X.im[e];
    }
    mat_mul<false, false, true>(R, E, Xb);
  }
  soa_store(out + plane * 9 * (size_t)Vs, Vs, site, R);
}

// U <- exp(eps P) U on the planar layout
__global__ void __launch_bounds__(128, 4) k_drift(C* __restrict__ U, const C* __restrict__ P, int V, double eps) {
  const int plane = blockIdx.y;
  const int site = blockIdx.x * 128 + threadIdx.x;
  if (site >= V) return;
  const C* pp = P + (size_t)plane * 9 * V;
  C* up = U + (size_t)plane * 9 * V;
  Mat3<T> p, ex, u, r;
  soa_load(p, pp, V, site);
#pragma unroll
  for (int e = 0; e < 9; ++e) { p.re[e] *= eps; p.im[e] *= eps; }
#pragma unroll
  for (int e = 0; e < 9; ++e) { const C v = up[(size_t)e * V + site]; u.re[e] = v.x; u.im[e] = v.y; }
  mat_exp_alg(ex, p);
  mat_mul<false, false, false>(r, ex, u);
  soa_store(up, V, site, r);
}

// plaquette traces: per-chain partial sums and (optionally) the full wloops tensor
template <bool WRITE_LOOPS>
__global__ void __launch_bounds__(128) k_plaq(const C* __restrict__ U, Lat lat, double* __restrict__ part,
                                              C* __restrict__ wloops, int nb) {
  __shared__ double red[4];
  const int b = blockIdx.y;
  const int site = blockIdx.x * 128 + threadIdx.x;
  double sr = 0.0, si = 0.0;
  if (site < lat.V) {
    T tr[6], ti[6];
    site_plaquette_traces<T, C>(tr, ti, U, lat, b, site);
#pragma unroll
    for (int p = 0; p < 6; ++p) {
      sr += tr[p];
      si += ti[p];
      if (WRITE_LOOPS) wloops[((size_t)p * nb + b) * lat.V + site] = make_double2(tr[p], ti[p]);
    }
  }
  if (part != nullptr) {
    sr = block_sum<128>(sr, red, threadIdx.x);
    si = block_sum<128>(si, red, threadIdx.x);
    if (threadIdx.x == 0) {
      double* o = part + ((size_t)b * gridDim.x + blockIdx.x) * 2;
      o[0] = sr;
      o[1] = si;
    }
  }
}

// out[b*out_stride + out_off] = scale * sum_j part[(b*nblk + j)*ncomp + comp] + shift
__global__ void __launch_bounds__(256) k_reduce_affine(const double* __restrict__ part, int nblk, int ncomp, int comp,
                                                       double scale, double shift, double* __restrict__ out,
                                                       int out_stride, int out_off) {
  __shared__ double red[8];
  const int b = blockIdx.x;
  double acc = 0.0;
  for (int j = threadIdx.x; j < nblk; j += 256) acc += part[((size_t)b * nblk + j) * ncomp + comp];
  acc = block_sum<256>(acc, red, threadIdx.x);
  if (threadIdx.x == 0) out[(size_t)b * out_stride + out_off] = fma(scale, acc, shift);
}

// ---------------------------------------------------------------------------
// link-local kernels on the boundary (AoS) layout.  grid.x covers `n_per_row`
// matrices of one row (a chain, or the whole array when rows == 1), grid.y = rows.
// ---------------------------------------------------------------------------
enum UnaryOp { OP_EXP = 0, OP_TAH = 1, OP_PROJECT = 2, OP_TOVEC = 3 };

template <int OP>
__global__ void __launch_bounds__(NTL) k_unary(const C* __restrict__ x, C* __restrict__ out, T* __restrict__ vec8,
                                               size_t nmat, double scale) {
  __shared__ C sm[NTL * 9];
  const size_t first = (size_t)blockIdx.x * NTL;
  const int n = (int)min((size_t)NTL, nmat - first);
  Mat3<T> m, r;
  block_load_mat<NTL>(m, sm, x, first, n);
  const bool live = (int)threadIdx.x < n;
  if (live) {
    if (OP == OP_EXP) {
#pragma unroll
      for (int e = 0; e < 9; ++e) { m.re[e] *= scale; m.im[e] *= scale; }
      mat_exp(r, m);
    } else if (OP == OP_TAH) {
      project_tah(r, m);
    } else if (OP == OP_PROJECT) {
      project_su(r, m);
    } else {
      r = m;
    }
    if (vec8 != nullptr) {
      T v[8];
      su3_to_vec(v, r);
      double2* o = reinterpret_cast<double2*>(vec8 + (first + threadIdx.x) * 8);
#pragma unroll
      for (int k = 0; k < 4; ++k) o[k] = make_double2(v[2 * k], v[2 * k + 1]);
    }
  }
  if (out != nullptr) block_store_mat<NTL>(out, sm, r, first, n);
}

__global__ void __launch_bounds__(NTL) k_from_vec(const T* __restrict__ vec8, C* __restrict__ out, size_t nmat) {
  __shared__ C sm[NTL * 9];
  const size_t first = (size_t)blockIdx.x * NTL;
  const int n = (int)min((size_t)NTL, nmat - first);
  Mat3<T> m;
  if ((int)threadIdx.x < n) {
    T v[8];
    const double2* in = reinterpret_cast<const double2*>(vec8 + (first + threadIdx.x) * 8);
#pragma unroll
    for (int k = 0; k < 4; ++k) { const double2 w = __ldg(in + k); v[2 * k] = w.x; v[2 * k + 1] = w.y; }
    vec_to_su3(m, v);
  }
  block_store_mat<NTL>(out, sm, m, first, n);
}

// x_out = m*x + exp(eps p) ((1-m)*x)   (mask == nullptr: x_out = exp(eps p) x)
__global__ void __launch_bounds__(NTL) k_update_gauge(const C* __restrict__ x, const C* __restrict__ p, double eps_in, const double* __restrict__ eps_dev,
                                                      const float* __restrict__ mask, int mask_complement,
                                                      C* __restrict__ out, size_t links_per_chain) {
  const double eps = eps_dev ? eps_in * eps_dev[0] : eps_in;   // device-resident step size (CUDA graphs)
  __shared__ C sm[NTL * 9];
  const size_t row0 = (size_t)blockIdx.y * links_per_chain;
  const size_t l0 = (size_t)blockIdx.x * NTL;               // link index inside the chain
  const int n = (int)min((size_t)NTL, links_per_chain - l0);
  Mat3<T> X, Pm, E, R;
  block_load_mat<NTL>(Pm, sm, p, row0 + l0, n);
  block_load_mat<NTL>(X, sm, x, row0 + l0, n);
  if ((int)threadIdx.x < n) {
#pragma unroll
    for (int e = 0; e < 9; ++e) { Pm.re[e] *= eps; Pm.im[e] *= eps; }
    mat_exp(E, Pm);
    if (mask == nullptr) {
      mat_mul<false, false, false>(R, E, X);
    } else {
      Mat3<T> Xm, Xb;
      const float* mk = mask + (l0 + threadIdx.x) * 9;
#pragma unroll
      for (int e = 0; e < 9; ++e) {
        float m = __ldg(mk + e);
        if (mask_complement) m = 1.0f - m;
        const T md = (T)m, mb = (T)(1.0f - m);
        Xm.re[e] = md * X.re[e]; Xm.im[e] = md * X.im[e];
        Xb.re[e] = mb * X.re[e]; Xb.im[e] = mb * X.im[e];
      }
      R = Xm;
      mat_mul<false, false, true>(R, E, Xb);
    }
  }
  block_store_mat<NTL>(out, sm, R, row0 + l0, n);
}

// L2HMC momentum update epilogue (see l2b.h)
__global__ void __launch_bounds__(NTL) k_vupdate(const C* __restrict__ v, const C* __restrict__ f,
                                                 const T* __restrict__ s, const T* __restrict__ t,
                                                 const T* __restrict__ q, double eps_in, const double* __restrict__ eps_dev, int sign, C* __restrict__ out,
                                                 double* __restrict__ part, size_t links_per_chain) {
  const double eps = eps_dev ? eps_in * eps_dev[0] : eps_in;   // device-resident step size (CUDA graphs)
  __shared__ C sm[NTL * 9];
  __shared__ double red[NTL / 32];
  const size_t row0 = (size_t)blockIdx.y * links_per_chain;
  const size_t l0 = (size_t)blockIdx.x * NTL;
  const int n = (int)min((size_t)NTL, links_per_chain - l0);
  Mat3<T> Vm, Fm, R;
  T sv[9], tv[9], qv[9];
  block_load_mat<NTL>(Vm, sm, v, row0 + l0, n);
  block_load_mat<NTL>(Fm, sm, f, row0 + l0, n);
  T* smr = reinterpret_cast<T*>(sm);
  block_load_real9<NTL>(sv, smr, s, row0 + l0, n);
  block_load_real9<NTL>(tv, smr, t, row0 + l0, n);
  block_load_real9<NTL>(qv, smr, q, row0 + l0, n);
  double ld = 0.0;
  if ((int)threadIdx.x < n) {
    const T sg = (T)sign;
    const T he = T(0.5) * eps;
#pragma unroll
    for (int e = 0; e < 9; ++e) {
      const T logjac = sg * eps * sv[e] / T(2);
      ld += logjac;
      const T es = exp(logjac);
      const T eq = exp(eps * qv[e]);
      const T fr = fma(Fm.re[e], eq, tv[e]);     // t is real: it only shifts the real part
      const T fi = Fm.im[e] * eq;
      if (sign > 0) {
        R.re[e] = es * Vm.re[e] - he * fr;
        R.im[e] = es * Vm.im[e] - he * fi;
      } else {
        R.re[e] = es * (Vm.re[e] + he * fr);
        R.im[e] = es * (Vm.im[e] + he * fi);
      }
    }
  }
  block_store_mat<NTL>(out, sm, R, row0 + l0, n);
  if (part != nullptr) {
    ld = block_sum<NTL>(ld, red, threadIdx.x);
    if (threadIdx.x == 0) part[(size_t)blockIdx.y * gridDim.x + blockIdx.x] = ld;
  }
}

// sum |p|^2 over a chain row of `row_len` complex numbers (layout agnostic)
__global__ void __launch_bounds__(256) k_norm2_rows(const C* __restrict__ p, size_t row_len, double* __restrict__ part) {
  __shared__ double red[8];
  const C* row = p + (size_t)blockIdx.y * row_len;
  double acc = 0.0;
  for (size_t k = (size_t)blockIdx.x * 256 + threadIdx.x; k < row_len; k += (size_t)gridDim.x * 256) {
    const C v = __ldg(row + k);
    acc += fma(v.x, v.x, fma(v.y, v.y, -8.0 / 9.0));   // 9 entries per link share the -8
  }
  acc = block_sum<256>(acc, red, threadIdx.x);
  if (threadIdx.x == 0) part[(size_t)blockIdx.y * gridDim.x + blockIdx.x] = acc;
}

// checkSU partials: (sum d, max d) per block
__global__ void __launch_bounds__(NTL) k_check(const C* __restrict__ x, double* __restrict__ part,
                                               size_t links_per_chain) {
  __shared__ C sm[NTL * 9];
  __shared__ double red[NTL / 32];
  const size_t row0 = (size_t)blockIdx.y * links_per_chain;
  const size_t l0 = (size_t)blockIdx.x * NTL;
  const int n = (int)min((size_t)NTL, links_per_chain - l0);
  Mat3<T> X;
  block_load_mat<NTL>(X, sm, x, row0 + l0, n);
  double d = 0.0;
  if ((int)threadIdx.x < n) d = check_su_dev(X);
  const double dsum = block_sum<NTL>(d, red, threadIdx.x);
  const double dmax = block_max<NTL>(d, red, threadIdx.x);
  if (threadIdx.x == 0) {
    double* o = part + ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * 2;
    o[0] = dsum;
    o[1] = dmax;
  }
}

__global__ void __launch_bounds__(256) k_check_final(const double* __restrict__ part, int nblk, double nlinks,
                                                     double* __restrict__ avg, double* __restrict__ mx) {
  __shared__ double red[8];
  const int b = blockIdx.x;
  double s = 0.0, m = 0.0;
  for (int j = threadIdx.x; j < nblk; j += 256) {
    s += part[((size_t)b * nblk + j) * 2];
    m = max(m, part[((size_t)b * nblk + j) * 2 + 1]);
  }
  s = block_sum<256>(s, red, threadIdx.x);
  m = block_max<256>(m, red, threadIdx.x);
  if (threadIdx.x == 0) {
    const double c = 2.0 * (3 * 3 + 1);
    avg[b] = sqrt(s / nlinks / c);
    mx[b] = sqrt(m / c);
  }
}

// ---------------------------------------------------------------------------
// Philox4x32-10 Gaussian momenta
// ---------------------------------------------------------------------------
__device__ __forceinline__ void philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
    const uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
}
// two 32-bit words -> uniform double in (0, 1)
__device__ __forceinline__ double u01(uint32_t a, uint32_t b) {
  const unsigned long long z = ((unsigned long long)a << 32) | b;
  return ((double)(z >> 11) + 0.5) * (1.0 / 9007199254740992.0);
}

__global__ void k_u64_add(unsigned long long* p, unsigned long long inc) { *p += inc; }

__global__ void __launch_bounds__(NTL) k_rand_momentum(uint64_t seed, uint64_t offset_in,
                                                       const unsigned long long* __restrict__ offset_dev,
                                                       C* __restrict__ p, double* __restrict__ part,
                                                       size_t links_per_chain) {
  // Philox stream offset: by value, plus a device-resident call counter when given (so that a
  // CUDA-graph replay draws fresh momenta: the counter is bumped by k_u64_add after this kernel)
  const uint64_t offset = offset_in + (offset_dev ? (uint64_t)offset_dev[0] : 0ull);
  __shared__ C sm[NTL * 9];
  __shared__ double red[NTL / 32];
  const size_t row0 = (size_t)blockIdx.y * links_per_chain;
  const size_t l0 = (size_t)blockIdx.x * NTL;
  const int n = (int)min((size_t)NTL, links_per_chain - l0);
  Mat3<T> m;
  double acc = 0.0;
  if ((int)threadIdx.x < n) {
    const uint64_t link = row0 + l0 + threadIdx.x;
    T nrm[8];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      uint32_t c[4] = {(uint32_t)link, (uint32_t)(link >> 32), (uint32_t)(offset * 4 + k), (uint32_t)((offset * 4 + k) >> 32)};
      philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
      const double u1 = u01(c[0], c[1]), u2 = u01(c[2], c[3]);
      const double r = sqrt(-2.0 * log(u1));
      double sn, cs;
      sincospi(2.0 * u2, &sn, &cs);
      nrm[2 * k] = r * cs;
      nrm[2 * k + 1] = r * sn;
    }
    tah_from_normals(m, nrm);
    acc = norm2(m) - 8.0;
  }
  block_store_mat<NTL>(p, sm, m, row0 + l0, n);
  if (part != nullptr) {
    acc = block_sum<NTL>(acc, red, threadIdx.x);
    if (threadIdx.x == 0) part[(size_t)blockIdx.y * gridDim.x + blockIdx.x] = acc;
  }
}

// ---------------------------------------------------------------------------
// host-side helpers
// ---------------------------------------------------------------------------
// launch geometries of k_force: (sites per block, min resident blocks per SM ->
// register cap).  Selected with l2b_set_option("su3_force_variant", i); the
// default is the one measured fastest on B200 (profiles/).
using ForceFn = void (*)(const C*, C*, Lat, double, double*, C*, double);
using ForceFnX = void (*)(const C*, C*, Lat, double, double*, C*, double, C*, C*);
using ForceFnTma = void (*)(const CUtensorMap, const CUtensorMap, const C*, C*, Lat, double, double*, C*, double);
struct ForceVariant {
  int ts;
  ForceFn kick, nokick;
  int smem;   // dynamic shared memory (cp.async operand ring), 0 for the register-only kernel
  ForceFn kick_drift;   // kick + fused drift into the second link buffer (nullptr: not available)
  int brick_fallback;   // >= 0: brick-tiled variant; use this linear variant when the lattice does not tile
  // trajectory ends with the layout conversions folded in (nullptr: separate conversion kernels):
  // [0] first step (momenta from the boundary layout), [1] last step (links also to the boundary layout),
  // [2] first == last (N_LF = 1), [3] final half kick (momenta to the boundary layout)
  ForceFnX ends[4];
  // [0] kick, [1] kick + drift with the block's own P / U tiles staged by the TMA (nullptr: none)
  ForceFnTma tma[2];
};
#define L2B_ENDS(HOOK)                                                                                   \
  {k_force_epx<32, 3, true, HOOK, 1, false>, k_force_epx<32, 3, true, HOOK, 0, true>,                      \
   k_force_epx<32, 3, true, HOOK, 1, true>, k_force_epx<32, 3, false, HOOK, 2, false>}
#define L2B_FV(TS, MINB) L2B_FVP(TS, MINB, 0)
#define L2B_FVP(TS, MINB, PF) \
  {TS, k_force<TS, MINB, true, PF, false>, k_force<TS, MINB, false, PF, false>, 0, k_force<TS, MINB, true, PF, true>, -1}
#define L2B_FVA(R, MINB) {32, k_force_async<R, MINB, true>, k_force_async<R, MINB, false>, R * 9 * 128 * 16, nullptr, -1}
#define L2B_FVB(TS, MINB, PF, FALLBACK)                                                              \
  {TS, k_force<TS, MINB, true, PF, false, true>, k_force<TS, MINB, false, PF, false, true>, 0,       \
   k_force<TS, MINB, true, PF, true, true>, FALLBACK}
const ForceVariant kForceVariants[] = {
    L2B_FV(32, 1),   // 0: 128 threads, uncapped registers
    L2B_FV(32, 3),   // 1: <= 168 registers, 12 warps / SM
    L2B_FV(32, 4),   // 2: <= 128 registers, 16 warps / SM
    L2B_FV(64, 2),   // 3: 256 threads, <= 128 registers
    L2B_FV(32, 5),   // 4: <= 96 registers (spills), 20 warps / SM
    L2B_FV(64, 1),   // 5: 256 threads, uncapped
    L2B_FV(16, 8),   // 6: 64 threads (2 directions per warp), <= 128 registers
    L2B_FVP(32, 3, 1),  // 7: variant 1 + L1 prefetch of the next direction's operands
    L2B_FVP(32, 3, 2),  // 8: variant 1 + L2 look-ahead prefetch of own links/momenta
    L2B_FVP(32, 4, 1),  // 9: variant 2 + L1 prefetch
    L2B_FVP(32, 4, 2),  // 10: variant 2 + L2 look-ahead
    L2B_FVA(4, 3),      // 11: cp.async ring of 4 operands, 12 warps / SM (216 KB smem / SM)
    L2B_FVA(3, 4),      // 12: ring of 3, 16 warps / SM
    L2B_FVA(6, 2),      // 13: ring of 6, 8 warps / SM
    L2B_FVA(3, 3),      // 14: ring of 3, 12 warps / SM
    L2B_FVA(2, 4),      // 15: ring of 2, 16 warps / SM
    L2B_FVB(32, 3, 0, 1),   // 16: brick (1,2,2,8), <= 168 registers
    L2B_FVB(32, 4, 0, 2),   // 17: brick (1,2,2,8), <= 128 registers
    L2B_FVB(64, 1, 0, 1),   // 18: brick (2,2,2,8), 256 threads, uncapped registers
    L2B_FVB(64, 2, 0, 2),   // 19: brick (2,2,2,8), 256 threads, <= 128 registers
    L2B_FVB(32, 3, 2, 8),   // 20: brick (1,2,2,8) + L2 look-ahead
    L2B_FVB(64, 2, 2, 10),  // 21: brick (2,2,2,8) + L2 look-ahead
    // early-momentum variants of variant 8 (kick kernels only; no-kick falls back to variant 8's)
    {32, k_force_ep<32, 3, false, 3>, k_force<32, 3, false, 2, false>, 0, k_force_ep<32, 3, true, 3>, -1, L2B_ENDS(3)},   // 22: before the last direction
    {32, k_force_ep<32, 3, false, 2>, k_force<32, 3, false, 2, false>, 0, k_force_ep<32, 3, true, 2>, -1, L2B_ENDS(2)},   // 23: before the 2nd direction
    {32, k_force_ep<32, 3, false, 0>, k_force<32, 3, false, 2, false>, 0, k_force_ep<32, 3, true, 0>, -1},   // 24: before the staples
    {32, k_force_ep<32, 2, false, 0>, k_force<32, 2, false, 2, false>, 0, k_force_ep<32, 2, true, 0>, -1},   // 25: as 24, uncapped registers (8 warps / SM)
    {32, k_force_ep<32, 4, false, 3>, k_force<32, 4, false, 2, false>, 0, k_force_ep<32, 4, true, 3>, -1},   // 26: as 22, <= 128 registers (16 warps / SM)
    {32, k_force_ep<32, 2, false, 3>, k_force<32, 2, false, 2, false>, 0, k_force_ep<32, 2, true, 3>, -1},   // 27: as 22, uncapped registers (8 warps / SM)
    {32, k_force_ep<32, 2, false, 2>, k_force<32, 2, false, 2, false>, 0, k_force_ep<32, 2, true, 2>, -1},   // 28: as 23, uncapped registers
    // row-streaming (low-register) staple products, bit-identical (tests/hostemu)
    {32, k_force_ep<32, 3, false, 3, true>, k_force<32, 3, false, 2, false>, 0, k_force_ep<32, 3, true, 3, true>, -1},   // 29: hook before the last direction
    {32, k_force_ep<32, 3, false, 2, true>, k_force<32, 3, false, 2, false>, 0, k_force_ep<32, 3, true, 2, true>, -1},   // 30: before the 2nd direction
    {32, k_force_ep<32, 4, false, 3, true>, k_force<32, 4, false, 2, false>, 0, k_force_ep<32, 4, true, 3, true>, -1},   // 31: as 29, <= 128 registers
    // 32: variant 22 with the block's own momenta / links staged through shared memory by the TMA (k_force_tma)
    {32, k_force_ep<32, 3, false, 3>, k_force<32, 3, false, 2, false>, 0, k_force_ep<32, 3, true, 3>, -1, L2B_ENDS(3),
     {k_force_tma<3, false>, k_force_tma<3, true>}},
    // 33: as 32, <= 128 registers (16 warps / SM)
    {32, k_force_ep<32, 3, false, 3>, k_force<32, 3, false, 2, false>, 0, k_force_ep<32, 3, true, 3>, -1, L2B_ENDS(3),
     {k_force_tma<4, false>, k_force_tma<4, true>}},
};
constexpr int kNumForceVariants = (int)(sizeof(kForceVariants) / sizeof(kForceVariants[0]));
int g_fuse_conversions = 1;   // fold the momentum / output layout conversions into the trajectory's end launches
int g_force_variant = 22;   // r1d: momentum loads issued before the last staple direction (-10 % on the 16^4 trajectory)
int g_fuse_drift = 1;
int g_gauge_bwd_minb = 3;
int g_force_carveout = -1;   // -1: driver default; 0..100: preferred shared-memory carve-out in percent

struct Geo {
  int force_variant;
  Lat lat;
  int nb;
  size_t links_per_chain;   // 4 V
  size_t field_elems;       // nb * 4 * V * 9 complex numbers
  int nblk_force;           // blocks per chain of k_force
  int nblk_link;            // blocks per chain of link-local kernels
  int nblk_conv;            // blocks per (chain, mu) plane of the converters
  int nblk_plaq;
  size_t part_elems;        // doubles reserved for partials
};

int make_geo(Geo& g, int nb, const int dims[4], int dtype) {
  L2B_REQUIRE(dims != nullptr, L2B_ERR_INVALID, "dims is NULL");
  L2B_REQUIRE(nb > 0 && dims[0] > 0 && dims[1] > 0 && dims[2] > 0 && dims[3] > 0, L2B_ERR_INVALID,
              "non-positive size: nb=%d dims=%d,%d,%d,%d", nb, dims[0], dims[1], dims[2], dims[3]);
  L2B_REQUIRE(dtype == L2B_F64, L2B_ERR_UNSUPPORTED, "SU(3) kernels implement L2B_F64 only (got dtype=%d)", dtype);
  const long long V = (long long)dims[0] * dims[1] * dims[2] * dims[3];
  L2B_REQUIRE(V * 9 < (1ll << 31), L2B_ERR_UNSUPPORTED, "lattice volume %lld too large for 32-bit site index", V);
  L2B_REQUIRE(nb * 4 <= 65535, L2B_ERR_UNSUPPORTED, "nb=%d exceeds grid.y limit (16383 chains per call)", nb);
  g.lat = make_lat(dims[0], dims[1], dims[2], dims[3]);
  g.nb = nb;
  g.links_per_chain = (size_t)4 * V;
  g.field_elems = (size_t)nb * 4 * V * 9;
  if (const char* ev = getenv("L2B_SU3_FORCE_VARIANT")) {
    const int v = atoi(ev);
    if (v >= 0 && v < kNumForceVariants) g_force_variant = v;
  }
  g.force_variant = g_force_variant;
  {
    const ForceVariant& fv0 = kForceVariants[g.force_variant];
    const bool tiles = (dims[3] % 8 == 0) && (dims[2] % 2 == 0) && (dims[1] % 2 == 0) && (fv0.ts != 64 || dims[0] % 2 == 0);
    if (fv0.brick_fallback >= 0 && !tiles) g.force_variant = fv0.brick_fallback;
  }
  const int fts = kForceVariants[g.force_variant].ts;
  g.nblk_force = (int)((V + fts - 1) / fts);
  g.nblk_link = (int)((4 * V + NTL - 1) / NTL);
  g.nblk_conv = (int)((V + NTL - 1) / NTL);
  g.nblk_plaq = (int)((V + 127) / 128);
  size_t per_chain = (size_t)g.nblk_force * 3;   // (Re tr G, |P'|^2, |P_in|^2) of the trajectory's first step
  if ((size_t)g.nblk_link * 2 > per_chain) per_chain = (size_t)g.nblk_link * 2;
  if ((size_t)g.nblk_conv * 4 > per_chain) per_chain = (size_t)g.nblk_conv * 4;
  g.part_elems = per_chain * nb;
  return L2B_OK;
}

size_t ws_bytes_of(const Geo& g) {
  return 3 * align_up(g.field_elems * sizeof(C), 256) + align_up(g.part_elems * sizeof(double), 256);
}

struct Ws {
  C* f0;
  C* f1;
  C* f2;
  double* part;
};

int carve(Ws& w, const Geo& g, void* ws, size_t ws_bytes) {
  L2B_REQUIRE(ws != nullptr, L2B_ERR_INVALID, "workspace is NULL");
  L2B_REQUIRE(((uintptr_t)ws & 255) == 0, L2B_ERR_INVALID, "workspace must be 256-byte aligned");
  L2B_REQUIRE(ws_bytes >= ws_bytes_of(g), L2B_ERR_WORKSPACE, "workspace too small: %zu < %zu", ws_bytes, ws_bytes_of(g));
  char* p = (char*)ws;
  w.f0 = (C*)p; p += align_up(g.field_elems * sizeof(C), 256);
  w.f1 = (C*)p; p += align_up(g.field_elems * sizeof(C), 256);
  w.f2 = (C*)p; p += align_up(g.field_elems * sizeof(C), 256);
  w.part = (double*)p;
  return L2B_OK;
}

int launch_a2s(const Geo& g, const C* aos, C* soa, double* part, cudaStream_t st) {
  dim3 grid(g.nblk_conv, g.nb * 4);
  if (part) k_aos_to_soa<true><<<grid, NTL, 0, st>>>(aos, soa, g.lat.V, part);
  else k_aos_to_soa<false><<<grid, NTL, 0, st>>>(aos, soa, g.lat.V, nullptr);
  L2B_LAUNCHED("k_aos_to_soa");
  return L2B_OK;
}
int launch_s2a(const Geo& g, const C* soa, C* aos, cudaStream_t st) {
  dim3 grid(g.nblk_conv, g.nb * 4);
  k_soa_to_aos<<<grid, NTL, 0, st>>>(soa, aos, g.lat.V);
  L2B_LAUNCHED("k_soa_to_aos");
  return L2B_OK;
}
// rank-2 tensor map of a planar field for the TMA: inner = the 2 V doubles of one (chain, mu, e) row, outer = the
// nb * 36 rows; box = 32 sites x 9 rows (one link tile of one direction)
int make_plane_map(CUtensorMap* m, const void* base, const Geo& g) {
  using EncodeFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static EncodeFn encode = nullptr;
  if (encode == nullptr) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qr;
    L2B_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr));
    L2B_REQUIRE(fn != nullptr && qr == cudaDriverEntryPointSuccess, L2B_ERR_CUDA, "cuTensorMapEncodeTiled not available");
    encode = (EncodeFn)fn;
  }
  const cuuint64_t dims[2] = {(cuuint64_t)2 * g.lat.V, (cuuint64_t)g.nb * 36};
  const cuuint64_t strides[1] = {(cuuint64_t)g.lat.V * sizeof(C)};
  const cuuint32_t box[2] = {64, 9};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult rc = encode(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<void*>(base), dims, strides, box, estr,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  L2B_REQUIRE(rc == CUDA_SUCCESS, L2B_ERR_CUDA, "cuTensorMapEncodeTiled failed (CUresult %d)", (int)rc);
  return L2B_OK;
}

int launch_force(const Geo& g, const C* U, C* P, bool kick, double coef, double* part, cudaStream_t st,
                 C* Uout = nullptr, double eps_drift = 0.0) {
  const ForceVariant& fv = kForceVariants[g.force_variant];
  dim3 grid(g.nblk_force, g.nb), block(fv.ts, 4);
  if (kick && fv.tma[Uout ? 1 : 0] != nullptr && g.lat.V % 32 == 0) {
    CUtensorMap mp, mu;
    int rc = make_plane_map(&mp, P, g);
    if (rc != L2B_OK) return rc;
    rc = make_plane_map(&mu, U, g);
    if (rc != L2B_OK) return rc;
    ForceFnTma fn = fv.tma[Uout ? 1 : 0];
    constexpr int smem = 2 * 4 * 9 * 32 * (int)sizeof(C);
    fn<<<grid, block, smem, st>>>(mp, mu, U, P, g.lat, coef, part, Uout, eps_drift);
    L2B_LAUNCHED("k_force_tma");
    return L2B_OK;
  }
  ForceFn fn = Uout ? fv.kick_drift : (kick ? fv.kick : fv.nokick);
  L2B_REQUIRE(fn != nullptr, L2B_ERR_UNSUPPORTED, "force variant %d has no fused kick+drift kernel", g.force_variant);
  if (fv.smem > 48 * 1024)
    L2B_CUDA(cudaFuncSetAttribute((const void*)fn, cudaFuncAttributeMaxDynamicSharedMemorySize, fv.smem));
  if (g_force_carveout >= 0)   // register-only variants: give the whole unified cache to L1 (neighbour reuse)
    L2B_CUDA(cudaFuncSetAttribute((const void*)fn, cudaFuncAttributePreferredSharedMemoryCarveout, g_force_carveout));
  fn<<<grid, block, fv.smem, st>>>(U, P, g.lat, coef, part, Uout, eps_drift);
  L2B_LAUNCHED("k_force");
  return L2B_OK;
}
int launch_force_end(const Geo& g, int which, const C* U, C* P, double coef, double* part, cudaStream_t st, C* Uout,
                     double eps_drift, C* Paos, C* Uaos) {
  const ForceVariant& fv = kForceVariants[g.force_variant];
  ForceFnX fn = fv.ends[which];
  L2B_REQUIRE(fn != nullptr, L2B_ERR_UNSUPPORTED, "force variant %d has no fused-conversion kernels", g.force_variant);
  if (g_force_carveout >= 0)
    L2B_CUDA(cudaFuncSetAttribute((const void*)fn, cudaFuncAttributePreferredSharedMemoryCarveout, g_force_carveout));
  fn<<<dim3(g.nblk_force, g.nb), dim3(fv.ts, 4), 0, st>>>(U, P, g.lat, coef, part, Uout, eps_drift, Paos, Uaos);
  L2B_LAUNCHED("k_force_epx");
  return L2B_OK;
}
int launch_reduce(const double* part, int nblk, int ncomp, int comp, double scale, double shift, double* out,
                  int out_stride, int out_off, int nb, cudaStream_t st) {
  k_reduce_affine<<<nb, 256, 0, st>>>(part, nblk, ncomp, comp, scale, shift, out, out_stride, out_off);
  L2B_LAUNCHED("k_reduce_affine");
  return L2B_OK;
}

#define L2B_TRY(expr)            \
  do {                           \
    int rc__ = (expr);           \
    if (rc__ != L2B_OK) return rc__; \
  } while (0)

}  // namespace
}  // namespace l2b

using namespace l2b;

// ===========================================================================
// C ABI
// ===========================================================================
extern "C" {

int l2b_set_option(const char* key, int value) {
  L2B_REQUIRE(key != nullptr, L2B_ERR_INVALID, "null key");
  if (strcmp(key, "su3_force_variant") == 0) {
    L2B_REQUIRE(value >= 0 && value < kNumForceVariants, L2B_ERR_INVALID, "su3_force_variant must be in [0, %d)",
                kNumForceVariants);
    g_force_variant = value;
    return L2B_OK;
  }
  if (strcmp(key, "su3_force_carveout") == 0) {
    L2B_REQUIRE(value >= -1 && value <= 100, L2B_ERR_INVALID, "su3_force_carveout must be in [-1, 100]");
    g_force_carveout = value;
    return L2B_OK;
  }
  if (strcmp(key, "su3_gauge_bwd_minb") == 0) {
    L2B_REQUIRE(value == 2 || value == 3, L2B_ERR_INVALID, "su3_gauge_bwd_minb must be 2 or 3");
    g_gauge_bwd_minb = value;
    return L2B_OK;
  }
  if (strcmp(key, "su3_fuse_drift") == 0) {
    g_fuse_drift = value != 0;
    return L2B_OK;
  }
  if (strcmp(key, "su3_fuse_conversions") == 0) {
    g_fuse_conversions = value != 0;
    return L2B_OK;
  }
  set_error("unknown option '%s'", key);
  return L2B_ERR_INVALID;
}

size_t l2b_su3_ws_bytes(int nb, const int dims[4], int dtype) {
  Geo g;
  if (make_geo(g, nb, dims, dtype) != L2B_OK) return 0;
  return ws_bytes_of(g);
}

int l2b_su3_aos_to_soa(const void* x_aos, void* x_soa, int nb, const int dims[4], int dtype, void* stream) {
  Geo g;
  L2B_TRY(make_geo(g, nb, dims, dtype));
  L2B_REQUIRE(x_aos && x_soa, L2B_ERR_INVALID, "null field pointer");
  return launch_a2s(g, (const C*)x_aos, (C*)x_soa, nullptr, (cudaStream_t)stream);
}

int l2b_su3_soa_to_aos(const void* x_soa, void* x_aos, int nb, const int dims[4], int dtype, void* stream) {
  Geo g;
  L2B_TRY(make_geo(g, nb, dims, dtype));
  L2B_REQUIRE(x_aos && x_soa, L2B_ERR_INVALID, "null field pointer");
  return launch_s2a(g, (const C*)x_soa, (C*)x_aos, (cudaStream_t)stream);
}

int l2b_su3_wilson_loops(const void* x, void* wloops, int nb, const int dims[4], int dtype, void* ws,
                         size_t ws_bytes, void* stream) {
  Geo g;
  Ws w;
  L2B_TRY(make_geo(g, nb, dims, dtype));
  L2B_TRY(carve(w, g, ws, ws_bytes));
  L2B_REQUIRE(x && wloops, L2B_ERR_INVALID, "null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  L2B_TRY(launch_a2s(g, (const C*)x, w.f0, nullptr, st));
  k_plaq<true><<<dim3(g.nblk_plaq, nb), 128, 0, st>>>(w.f0, g.lat, nullptr, (C*)wloops, nb);
  L2B_LAUNCHED("k_plaq");
  return L2B_OK;
}

int l2b_su3_plaq_sums(const void* x, double* sums, int nb, const int dims[4], int dtype, void* ws, size_t ws_bytes,
                      void* stream) {
  Geo g;
  Ws w;
  L2B_TRY(make_geo(g, nb, dims, dtype));
  L2B_TRY(carve(w, g, ws, ws_bytes));
  L2B_REQUIRE(x && sums, L2B_ERR_INVALID, "null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  L2B_TRY(launch_a2s(g, (const C*)x, w.f0, nullptr, st));
  k_plaq<false><<<dim3(g.nblk_plaq, nb), 128, 0, st>>>(w.f0, g.lat, w.part, nullptr, nb);
  L2B_LAUNCHED("k_plaq");
  L2B_TRY(launch_reduce(w.part, g.nblk_plaq, 2, 0, 1.0, 0.0, sums, 2, 0, nb, st));
  L2B_TRY(launch_reduce(w.part, g.nblk_plaq, 2, 1, 1.0, 0.0, sums, 2, 1, nb, st));
  return L2B_OK;
}

int l2b_su3_force(const void* x, double beta, void* force, double* plaq_sum_or_null, int nb, const int dims[4],
                  int dtype, void* ws, size_t ws_bytes, void* stream) {
  Geo g;
  Ws w;
  L2B_TRY(make_geo(g, nb, dims, dtype));
  L2B_TRY(carve(w, g, ws, ws_bytes));
  L2B_REQUIRE(x && force, L2B_ERR_INVALID, "null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  L2B_TRY(launch_a2s(g, (const C*)x, w.f0, nullptr, st));
  L2B_TRY(launch_force(g, w.f0, w.f1, false, beta / 3.0, plaq_sum_or_null ? w.part : nullptr, st));
  if (plaq_sum_or_null)  // sum_links Re tr(U A) counts every plaquette four times
    L2B_TRY(launch_reduce(w.part, g.nblk_force, 2, 0, 0.25, 0.0, plaq_sum_or_null, 1, 0, nb, st));
  return launch_s2a(g, w.f1, (C*)force, st);
}

int l2b_su3_force_c1(const void* x, double beta, double c1, void* force_or_null, double* sums_or_null, int nb,
                     const int dims[4], int dtype, void* ws, size_t ws_bytes, void* stream) {
  Geo g;
  Ws w;
  L2B_TRY(make_geo(g, nb, dims, dtype));
  L2B_TRY(carve(w, g, ws, ws_bytes));
  L2B_REQUIRE(x && (force_or_null || sums_or_null), L2B_ERR_INVALID, "null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  const int nblk = (g.lat.V + 31) / 32;     // <= nblk_link, so the partials fit (make_geo)
  L2B_TRY(launch_a2s(g, (const C*)x, w.f0, nullptr, st));
  k_force_c1<32><<<dim3(nblk, nb), dim3(32, 4), 0, st>>>(w.f0, force_or_null ? w.f1 : nullptr, g.lat, beta / 3.0, c1,
                                                        sums_or_null ? w.part : nullptr);
  L2B_LAUNCHED("k_force_c1");
  if (sums_or_null) {   // every plaquette is seen from its 4 links, every rectangle from its 6
    L2B_TRY(launch_reduce(w.part, nblk, 2, 0, 0.25, 0.0, sums_or_null, 2, 0, nb, st));
    L2B_TRY(launch_reduce(w.part, nblk, 2, 1, 1.0 / 6.0, 0.0, sums_or_null, 2, 1, nb, st));
  }
  if (force_or_null) return launch_s2a(g, w.f1, (C*)force_or_null, st);
  return L2B_OK;
}

int l2b_su3_exp(const void* p, double scale, void* out, size_t nmat, int dtype, void* stream) {
  L2B_REQUIRE(dtype == L2B_F64, L2B_ERR_UNSUPPORTED, "SU(3) kernels implement L2B_F64 only");
  L2B_REQUIRE(p && out, L2B_ERR_INVALID, "null pointer");
  if (nmat == 0) return L2B_OK;
  const unsigned nblk = (unsigned)((nmat + NTL - 1) / NTL);
  k_unary<OP_EXP><<<nblk, NTL, 0, (cudaStream_t)stream>>>((const C*)p, (C*)out, nullptr, nmat, scale);
  L2B_LAUNCHED("k_unary<exp>");
  return L2B_OK;
}

int l2b_su3_tah(const void* x, void* out, size_t nmat, int dtype, void* stream) {
  L2B_REQUIRE(dtype == L2B_F64, L2B_ERR_UNSUPPORTED, "SU(3) kernels implement L2B_F64 only");
  L2B_REQUIRE(x && out, L2B_ERR_INVALID, "null pointer");
  if (nmat == 0) return L2B_OK;
  const unsigned nblk = (unsigned)((nmat + NTL - 1) / NTL);
  k_unary<OP_TAH><<<nblk, NTL, 0, (cudaStream_t)stream>>>((const C*)x, (C*)out, nullptr, nmat, 1.0);
  L2B_LAUNCHED("k_unary<tah>");
  return L2B_OK;
}

int l2b_su3_project(const void* x, void* x_proj_or_null, void* vec8_or_null, size_t nmat, int dtype, void* stream) {
  L2B_REQUIRE(dtype == L2B_F64, L2B_ERR_UNSUPPORTED, "SU(3) kernels implement L2B_F64 only");
  L2B_REQUIRE(x && (x_proj_or_null || vec8_or_null), L2B_ERR_INVALID, "null pointer");
  if (nmat == 0) return L2B_OK;
  const unsigned nblk = (unsigned)((nmat + NTL - 1) / NTL);
  k_unary<OP_PROJECT><<<nblk, NTL, 0, (cudaStream_t)stream>>>((const C*)x, (C*)x_proj_or_null, (T*)vec8_or_null, nmat, 1.0);
  L2B_LAUNCHED("k_unary<project>");
  return L2B_OK;
}

int l2b_su3_to_vec(const void* x, void* vec8, size_t nmat, int dtype, void* stream) {
  L2B_REQUIRE(dtype == L2B_F64, L2B_ERR_UNSUPPORTED, "SU(3) kernels implement L2B_F64 only");
  L2B_REQUIRE(x && vec8, L2B_ERR_INVALID, "null pointer");
  if (nmat == 0) return L2B_OK;
  const unsigned nblk = (unsigned)((nmat + NTL - 1) / NTL);
  k_unary<OP_TOVEC><<<nblk, NTL, 0, (cudaStream_t)stream>>>((const C*)x, nullptr, (T*)vec8, nmat, 1.0);
  L2B_LAUNCHED("k_unary<tovec>");
  return L2B_OK;
}

int l2b_su3_from_vec(const void* vec8, void* x, size_t nmat, int dtype, void* stream) {
  L2B_REQUIRE(dtype == L2B_F64, L2B_ERR_UNSUPPORTED, "SU(3) kernels implement L2B_F64 only");
  L2B_REQUIRE(x && vec8, L2B_ERR_INVALID, "null pointer");
  if (nmat == 0) return L2B_OK;
  const unsigned nblk = (unsigned)((nmat + NTL - 1) / NTL);
  k_from_vec<<<nblk, NTL, 0, (cudaStream_t)stream>>>((const T*)vec8, (C*)x, nmat);
  L2B_LAUNCHED("k_from_vec");
  return L2B_OK;
}

int l2b_su3_update_gauge(const void* x, const void* p, double eps, const double* eps_dev, const float* mask, int mask_complement,
                         void* x_out, int nb, const int dims[4], int dtype, void* stream) {
  Geo g;
  L2B_TRY(make_geo(g, nb, dims, dtype));
  L2B_REQUIRE(x && p && x_out, L2B_ERR_INVALID, "null pointer");
  k_update_gauge<<<dim3(g.nblk_link, nb), NTL, 0, (cudaStream_t)stream>>>((const C*)x, (const C*)p, eps, eps_dev, mask,
                                                                          mask_complement, (C*)x_out,
                                                                          g.links_per_chain);
  L2B_LAUNCHED("k_update_gauge");
  return L2B_OK;
}

int l2b_su3_kinetic(const void* p, double* ke, int nb, const int dims[4], int dtype, void* ws, size_t ws_bytes,
                    void* stream) {
  Geo g;
  Ws w;
  L2B_TRY(make_geo(g, nb, dims, dtype));
  L2B_TRY(carve(w, g, ws, ws_bytes));
  L2B_REQUIRE(p && ke, L2B_ERR_INVALID, "null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  const size_t row = g.links_per_chain * 9;
  int nblk = (int)((row + 256 * 8 - 1) / (256 * 8));
  if (nblk > g.nblk_link) nblk = g.nblk_link;
  if (nblk < 1) nblk = 1;
  k_norm2_rows<<<dim3(nblk, nb), 256, 0, st>>>((const C*)p, row, w.part);
  L2B_LAUNCHED("k_norm2_rows");
  return launch_reduce(w.part, nblk, 1, 0, 0.5, 0.0, ke, 1, 0, nb, st);
}

int l2b_su3_check(const void* x, double* avg, double* mx, int nb, const int dims[4], int dtype, void* ws,
                  size_t ws_bytes, void* stream) {
  Geo g;
  Ws w;
  L2B_TRY(make_geo(g, nb, dims, dtype));
  L2B_TRY(carve(w, g, ws, ws_bytes));
  L2B_REQUIRE(x && avg && mx, L2B_ERR_INVALID, "null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  k_check<<<dim3(g.nblk_link, nb), NTL, 0, st>>>((const C*)x, w.part, g.links_per_chain);
  L2B_LAUNCHED("k_check");
  k_check_final<<<nb, 256, 0, st>>>(w.part, g.nblk_link, (double)g.links_per_chain, avg, mx);
  L2B_LAUNCHED("k_check_final");
  return L2B_OK;
}

int l2b_su3_rand_momentum(uint64_t seed, uint64_t offset, uint64_t* offset_dev_or_null, void* p, double* ke_or_null,
                          int nb, const int dims[4], int dtype, void* ws, size_t ws_bytes, void* stream) {
  Geo g;
  Ws w;
  L2B_TRY(make_geo(g, nb, dims, dtype));
  L2B_REQUIRE(p, L2B_ERR_INVALID, "null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  double* part = nullptr;
  if (ke_or_null) {
    L2B_TRY(carve(w, g, ws, ws_bytes));
    part = w.part;
  }
  k_rand_momentum<<<dim3(g.nblk_link, nb), NTL, 0, st>>>(seed, offset, (const unsigned long long*)offset_dev_or_null,
                                                         (C*)p, part, g.links_per_chain);
  L2B_LAUNCHED("k_rand_momentum");
  if (offset_dev_or_null) {
    k_u64_add<<<1, 1, 0, st>>>((unsigned long long*)offset_dev_or_null, 1ull);
    L2B_LAUNCHED("k_u64_add");
  }
  if (ke_or_null)
    L2B_TRY(launch_reduce(part, g.nblk_link, 1, 0, 0.5, 0.0, ke_or_null, 1, 0, nb, st));
  return L2B_OK;
}

int l2b_su3_vupdate(const void* v, const void* force, const void* s, const void* t, const void* q, double eps, const double* eps_dev,
                    int sign, void* v_out, double* logdet, int nb, const int dims[4], int dtype, void* ws,
                    size_t ws_bytes, void* stream) {
  Geo g;
  Ws w;
  L2B_TRY(make_geo(g, nb, dims, dtype));
  L2B_REQUIRE(v && force && v_out, L2B_ERR_INVALID, "null pointer");
  L2B_REQUIRE(sign == 1 || sign == -1, L2B_ERR_INVALID, "sign must be +1 or -1");
  cudaStream_t st = (cudaStream_t)stream;
  double* part = nullptr;
  if (logdet) {
    L2B_TRY(carve(w, g, ws, ws_bytes));
    part = w.part;
  }
  k_vupdate<<<dim3(g.nblk_link, nb), NTL, 0, st>>>((const C*)v, (const C*)force, (const T*)s, (const T*)t,
                                                   (const T*)q, eps, eps_dev, sign, (C*)v_out, part, g.links_per_chain);
  L2B_LAUNCHED("k_vupdate");
  if (logdet) L2B_TRY(launch_reduce(part, g.nblk_link, 1, 0, 1.0, 0.0, logdet, 1, 0, nb, st));
  return L2B_OK;
}

int l2b_su3_force_kick_planar(const void* u_planar, void* p_planar, double beta, double eps_kick,
                              double* sums_or_null, int nb, const int dims[4], int dtype, void* ws, size_t ws_bytes,
                              void* stream) {
  Geo g;
  Ws w;
  L2B_TRY(make_geo(g, nb, dims, dtype));
  L2B_REQUIRE(u_planar && p_planar, L2B_ERR_INVALID, "null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  double* part = nullptr;
  if (sums_or_null) {
    L2B_TRY(carve(w, g, ws, ws_bytes));
    part = w.part;
  }
  L2B_TRY(launch_force(g, (const C*)u_planar, (C*)p_planar, true, eps_kick * beta / 3.0, part, st));
  if (sums_or_null) {
    L2B_TRY(launch_reduce(part, g.nblk_force, 2, 0, 0.25, 0.0, sums_or_null, 2, 0, nb, st));
    L2B_TRY(launch_reduce(part, g.nblk_force, 2, 1, 1.0, 0.0, sums_or_null, 2, 1, nb, st));
  }
  return L2B_OK;
}

int l2b_su3_force_kick_drift_planar(const void* u_in_planar, void* p_planar, void* u_out_planar, double beta,
                                    double eps_kick, double eps_drift, double* sums_or_null, int nb,
                                    const int dims[4], int dtype, void* ws, size_t ws_bytes, void* stream) {
  Geo g;
  Ws w;
  L2B_TRY(make_geo(g, nb, dims, dtype));
  L2B_REQUIRE(u_in_planar && p_planar && u_out_planar, L2B_ERR_INVALID, "null pointer");
  L2B_REQUIRE(u_in_planar != u_out_planar, L2B_ERR_INVALID, "u_out must not alias u_in");
  cudaStream_t st = (cudaStream_t)stream;
  double* part = nullptr;
  if (sums_or_null) {
    L2B_TRY(carve(w, g, ws, ws_bytes));
    part = w.part;
  }
  L2B_TRY(launch_force(g, (const C*)u_in_planar, (C*)p_planar, true, eps_kick * beta / 3.0, part, st,
                       (C*)u_out_planar, eps_drift));
  if (sums_or_null) {
    L2B_TRY(launch_reduce(part, g.nblk_force, 2, 0, 0.25, 0.0, sums_or_null, 2, 0, nb, st));
    L2B_TRY(launch_reduce(part, g.nblk_force, 2, 1, 1.0, 0.0, sums_or_null, 2, 1, nb, st));
  }
  return L2B_OK;
}

int l2b_su3_drift_planar(void* u_planar, const void* p_planar, double eps, int nb, const int dims[4], int dtype,
                         void* stream) {
  Geo g;
  L2B_TRY(make_geo(g, nb, dims, dtype));
  L2B_REQUIRE(u_planar && p_planar, L2B_ERR_INVALID, "null pointer");
  k_drift<<<dim3((g.lat.V + 127) / 128, nb * 4), 128, 0, (cudaStream_t)stream>>>((C*)u_planar, (const C*)p_planar,
                                                                                 g.lat.V, eps);
  L2B_LAUNCHED("k_drift");
  return L2B_OK;
}

int l2b_su3_hmc_trajectory(const void* x, const void* v, double beta, double eps, int nlf, void* x_prop,
                           void* v_prop, double* energies, int nb, const int dims[4], int dtype, void* ws,
                           size_t ws_bytes, void* stream) {
  Geo g;
  Ws w;
  L2B_TRY(make_geo(g, nb, dims, dtype));
  L2B_TRY(carve(w, g, ws, ws_bytes));
  L2B_REQUIRE(x && v && x_prop && v_prop && energies, L2B_ERR_INVALID, "null pointer");
  L2B_REQUIRE(nlf >= 1, L2B_ERR_INVALID, "nlf must be >= 1 (got %d)", nlf);
  cudaStream_t st = (cudaStream_t)stream;
  C* U = w.f0;
  C* P = w.f1;
  const double b3 = beta / 3.0;
  const double ke_shift = 0.0;  // the -8 per link is already inside the partials
  const bool fused = g_fuse_drift && kForceVariants[g.force_variant].kick_drift != nullptr;
  if (fused && g_fuse_conversions && kForceVariants[g.force_variant].ends[0] != nullptr) {
    // Only the links are converted: the first step reads the momenta straight from the caller's tensor (and sums
    // KE0), the last step writes x_prop next to its planar output, the final half kick writes v_prop.
    // 1 conversion pass + 1 extra field write instead of 4 conversion passes (8 field transfers -> 3).
    L2B_TRY(launch_a2s(g, (const C*)x, U, nullptr, st));
    C* Ua = U;
    C* Ub = w.f2;
    for (int k = 0; k < nlf; ++k) {
      const bool first = (k == 0), last = (k == nlf - 1);
      const double coef = (first ? 0.5 : 1.0) * eps * b3;
      if (first || last) {
        L2B_TRY(launch_force_end(g, first ? (last ? 2 : 0) : 1, Ua, P, coef, first ? w.part : nullptr, st, Ub, eps,
                                 first ? (C*)v : nullptr, last ? (C*)x_prop : nullptr));
      } else {
        L2B_TRY(launch_force(g, Ua, P, true, coef, nullptr, st, Ub, eps));
      }
      if (first) {
        L2B_TRY(launch_reduce(w.part, g.nblk_force, 3, 2, 0.5, ke_shift, energies, 4, 0, nb, st));
        L2B_TRY(launch_reduce(w.part, g.nblk_force, 3, 0, -b3 * 0.25, 0.0, energies, 4, 1, nb, st));
      }
      C* t = Ua; Ua = Ub; Ub = t;
    }
    L2B_TRY(launch_force_end(g, 3, Ua, P, 0.5 * eps * b3, w.part, st, nullptr, 0.0, (C*)v_prop, nullptr));
    L2B_TRY(launch_reduce(w.part, g.nblk_force, 2, 1, 0.5, ke_shift, energies, 4, 2, nb, st));
    L2B_TRY(launch_reduce(w.part, g.nblk_force, 2, 0, -b3 * 0.25, 0.0, energies, 4, 3, nb, st));
    return L2B_OK;
  }
  // in: boundary layout -> planar; KE0 rides on the momentum conversion
  L2B_TRY(launch_a2s(g, (const C*)x, U, nullptr, st));
  L2B_TRY(launch_a2s(g, (const C*)v, P, w.part, st));
  L2B_TRY(launch_reduce(w.part, g.nblk_conv * 4, 1, 0, 0.5, ke_shift, energies, 4, 0, nb, st));
  if (fused) {
    // kick + drift fused, links ping-pong between two planar buffers:
    //   K(eps/2) K(eps) ... K(eps)   [nlf launches, each: P -= c F(U); U' = exp(eps P) U]   +   final half kick
    C* Ua = U;
    C* Ub = w.f2;
    for (int k = 0; k < nlf; ++k) {
      L2B_TRY(launch_force(g, Ua, P, true, (k == 0 ? 0.5 : 1.0) * eps * b3, k == 0 ? w.part : nullptr, st, Ub, eps));
      if (k == 0) L2B_TRY(launch_reduce(w.part, g.nblk_force, 2, 0, -b3 * 0.25, 0.0, energies, 4, 1, nb, st));
      C* t = Ua; Ua = Ub; Ub = t;
    }
    L2B_TRY(launch_force(g, Ua, P, true, 0.5 * eps * b3, w.part, st));
    U = Ua;
  } else {
    // first half kick; S0 rides on the force evaluation
    L2B_TRY(launch_force(g, U, P, true, 0.5 * eps * b3, w.part, st));
    L2B_TRY(launch_reduce(w.part, g.nblk_force, 2, 0, -b3 * 0.25, 0.0, energies, 4, 1, nb, st));
    const dim3 dgrid((g.lat.V + 127) / 128, nb * 4);
    for (int k = 1; k <= nlf; ++k) {
      k_drift<<<dgrid, 128, 0, st>>>(U, P, g.lat.V, eps);
      L2B_LAUNCHED("k_drift");
      const bool last = (k == nlf);
      L2B_TRY(launch_force(g, U, P, true, (last ? 0.5 : 1.0) * eps * b3, last ? w.part : nullptr, st));
    }
  }
  L2B_TRY(launch_reduce(w.part, g.nblk_force, 2, 1, 0.5, ke_shift, energies, 4, 2, nb, st));
  L2B_TRY(launch_reduce(w.part, g.nblk_force, 2, 0, -b3 * 0.25, 0.0, energies, 4, 3, nb, st));
  L2B_TRY(launch_s2a(g, U, (C*)x_prop, st));
  L2B_TRY(launch_s2a(g, P, (C*)v_prop, st));
  return L2B_OK;
}

int l2b_su3_action_grad(const void* x, const double* coef, void* gx, int nb, const int dims[4], int dtype, void* ws,
                        size_t ws_bytes, void* stream) {
  Geo g;
  Ws w;
  L2B_TRY(make_geo(g, nb, dims, dtype));
  L2B_TRY(carve(w, g, ws, ws_bytes));
  L2B_REQUIRE(x && coef && gx, L2B_ERR_INVALID, "null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  L2B_TRY(launch_a2s(g, (const C*)x, w.f0, nullptr, st));
  k_action_grad<32><<<dim3((g.lat.V + 31) / 32, nb), dim3(32, 4), 0, st>>>(w.f0, w.f1, g.lat, coef, 0.0, nullptr);
  L2B_LAUNCHED("k_action_grad");
  return launch_s2a(g, w.f1, (C*)gx, st);
}

int l2b_su3_force_bwd(const void* x, double beta, const void* gforce, void* gx, int nb, const int dims[4], int dtype,
                      void* ws, size_t ws_bytes, void* stream) {
  Geo g;
  Ws w;
  L2B_TRY(make_geo(g, nb, dims, dtype));
  L2B_TRY(carve(w, g, ws, ws_bytes));
  L2B_REQUIRE(x && gforce && gx, L2B_ERR_INVALID, "null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  L2B_TRY(launch_a2s(g, (const C*)x, w.f0, nullptr, st));
  L2B_TRY(launch_a2s(g, (const C*)gforce, w.f2, nullptr, st));
  k_action_grad<32><<<dim3((g.lat.V + 31) / 32, nb), dim3(32, 4), 0, st>>>(w.f0, w.f1, g.lat, nullptr, -beta / 3.0, w.f2);
  L2B_LAUNCHED("k_action_grad<force_bwd>");
  return launch_s2a(g, w.f1, (C*)gx, st);
}

int l2b_su3_action_grad_c1(const void* x, const double* coef_or_null, double scale, double c1,
                           const void* gforce_or_null, void* gx, int nb, const int dims[4], int dtype, void* ws,
                           size_t ws_bytes, void* stream) {
  Geo g;
  Ws w;
  L2B_TRY(make_geo(g, nb, dims, dtype));
  L2B_TRY(carve(w, g, ws, ws_bytes));
  L2B_REQUIRE(x && gx, L2B_ERR_INVALID, "null pointer");
  L2B_REQUIRE(coef_or_null || gforce_or_null, L2B_ERR_INVALID, "need coef (action adjoint) or gforce (force adjoint)");
  cudaStream_t st = (cudaStream_t)stream;
  L2B_TRY(launch_a2s(g, (const C*)x, w.f0, nullptr, st));
  if (gforce_or_null) L2B_TRY(launch_a2s(g, (const C*)gforce_or_null, w.f2, nullptr, st));
  k_action_grad_c1<32><<<dim3((g.lat.V + 31) / 32, nb), dim3(32, 4), 0, st>>>(
      w.f0, w.f1, g.lat, gforce_or_null ? nullptr : coef_or_null, scale, c1, gforce_or_null ? w.f2 : nullptr);
  L2B_LAUNCHED("k_action_grad_c1");
  return launch_s2a(g, w.f1, (C*)gx, st);
}

int l2b_su3_wilson_loops_bwd(const void* x, const void* gwloops, void* gx, int nb, const int dims[4], int dtype,
                             void* ws, size_t ws_bytes, void* stream) {
  Geo g;
  Ws w;
  L2B_TRY(make_geo(g, nb, dims, dtype));
  L2B_TRY(carve(w, g, ws, ws_bytes));
  L2B_REQUIRE(x && gwloops && gx, L2B_ERR_INVALID, "null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  L2B_TRY(launch_a2s(g, (const C*)x, w.f0, nullptr, st));
  k_wloops_bwd<32><<<dim3((g.lat.V + 31) / 32, nb), dim3(32, 4), 0, st>>>(w.f0, (const C*)gwloops, w.f1, g.lat, nb);
  L2B_LAUNCHED("k_wloops_bwd");
  return launch_s2a(g, w.f1, (C*)gx, st);
}

int l2b_su3_vupdate_bwd(const void* v, const void* force, const void* s, const void* t, const void* q, double eps, const double* eps_dev,
                        int sign, const void* gv_out, const double* glogdet, void* gv, void* gforce, void* gs, void* gt,
                        void* gq, double* geps, int nb, const int dims[4], int dtype, void* ws, size_t ws_bytes, void* stream) {
  Geo g;
  Ws w;
  L2B_TRY(make_geo(g, nb, dims, dtype));
  L2B_TRY(carve(w, g, ws, ws_bytes));
  L2B_REQUIRE(v && force && gv_out && gv && geps, L2B_ERR_INVALID, "null pointer");
  L2B_REQUIRE(sign == 1 || sign == -1, L2B_ERR_INVALID, "sign must be +1 or -1");
  cudaStream_t st = (cudaStream_t)stream;
  k_vupdate_bwd<<<dim3(g.nblk_link, nb), NTL, 0, st>>>((const C*)v, (const C*)force, (const T*)s, (const T*)t,
                                                       (const T*)q, eps, eps_dev, sign, (const C*)gv_out, glogdet, (C*)gv,
                                                       (C*)gforce, (T*)gs, (T*)gt, (T*)gq, w.part,
                                                       g.links_per_chain);
  L2B_LAUNCHED("k_vupdate_bwd");
  return launch_reduce(w.part, g.nblk_link, 1, 0, 1.0, 0.0, geps, 1, 0, nb, st);
}

int l2b_su3_update_gauge_bwd(const void* x, const void* p, double eps, const double* eps_dev, const float* mask, int mask_complement,
                             const void* gx_out, void* gx, void* gp, double* geps, int* bad_flag, int nb,
                             const int dims[4], int dtype, void* ws, size_t ws_bytes, void* stream) {
  Geo g;
  Ws w;
  L2B_TRY(make_geo(g, nb, dims, dtype));
  L2B_TRY(carve(w, g, ws, ws_bytes));
  L2B_REQUIRE(x && p && gx_out && gx && gp && geps && bad_flag, L2B_ERR_INVALID, "null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  // register cap: 3 resident blocks per SM (168 registers, small spill) or 2 (240, none); option su3_gauge_bwd_minb
  if (g_gauge_bwd_minb == 2) k_update_gauge_bwd<2><<<dim3(g.nblk_link, nb), NTL, 0, st>>>((const C*)x, (const C*)p, eps, eps_dev, mask, mask_complement,
                                                            (const C*)gx_out, (C*)gx, (C*)gp, w.part, bad_flag,
                                                            g.links_per_chain);
  else k_update_gauge_bwd<3><<<dim3(g.nblk_link, nb), NTL, 0, st>>>((const C*)x, (const C*)p, eps, eps_dev, mask, mask_complement,
                                                            (const C*)gx_out, (C*)gx, (C*)gp, w.part, bad_flag,
                                                            g.links_per_chain);
  L2B_LAUNCHED("k_update_gauge_bwd");
  return launch_reduce(w.part, g.nblk_link, 1, 0, 1.0, 0.0, geps, 1, 0, nb, st);
}

int l2b_su3_to_vec_bwd(const void* gvec8, void* gx, size_t nmat, int dtype, void* stream) {
  L2B_REQUIRE(dtype == L2B_F64, L2B_ERR_UNSUPPORTED, "SU(3) kernels implement L2B_F64 only");
  L2B_REQUIRE(gvec8 && gx, L2B_ERR_INVALID, "null pointer");
  if (nmat == 0) return L2B_OK;
  const unsigned nblk = (unsigned)((nmat + NTL - 1) / NTL);
  k_to_vec_bwd<<<nblk, NTL, 0, (cudaStream_t)stream>>>((const T*)gvec8, (C*)gx, nmat);
  L2B_LAUNCHED("k_to_vec_bwd");
  return L2B_OK;
}

int l2b_su3_project_vec(const void* x, void* vec8, int vec_dtype, size_t nmat, int dtype, void* stream) {
  L2B_REQUIRE(dtype == L2B_F64, L2B_ERR_UNSUPPORTED, "SU(3) kernels implement L2B_F64 only");
  L2B_REQUIRE(x && vec8, L2B_ERR_INVALID, "null pointer");
  if (nmat == 0) return L2B_OK;
  const unsigned nblk = (unsigned)((nmat + NTL - 1) / NTL);
  cudaStream_t st = (cudaStream_t)stream;
  if (vec_dtype == L2B_F64) k_project_vec<double><<<nblk, NTL, 0, st>>>((const C*)x, (double*)vec8, nmat);
  else if (vec_dtype == L2B_F32) k_project_vec<float><<<nblk, NTL, 0, st>>>((const C*)x, (float*)vec8, nmat);
  else if (vec_dtype == L2B_BF16) k_project_vec<__nv_bfloat16><<<nblk, NTL, 0, st>>>((const C*)x, (__nv_bfloat16*)vec8, nmat);
  else L2B_REQUIRE(false, L2B_ERR_UNSUPPORTED, "vec_dtype must be L2B_F64, L2B_F32 or L2B_BF16");
  L2B_LAUNCHED("k_project_vec");
  return L2B_OK;
}

int l2b_su3_project_bwd(const void* x, const void* gmat_or_null, const void* gvec8_or_null, int vec_dtype, void* gx,
                        size_t nmat, int dtype, void* stream) {
  L2B_REQUIRE(dtype == L2B_F64, L2B_ERR_UNSUPPORTED, "SU(3) kernels implement L2B_F64 only");
  L2B_REQUIRE(x && gx && (gmat_or_null || gvec8_or_null), L2B_ERR_INVALID, "null pointer");
  if (nmat == 0) return L2B_OK;
  const unsigned nblk = (unsigned)((nmat + NTL - 1) / NTL);
  cudaStream_t st = (cudaStream_t)stream;
  const C* gm = (const C*)gmat_or_null;
  if (vec_dtype == L2B_F64) k_project_bwd<double><<<nblk, NTL, 0, st>>>((const C*)x, gm, (const double*)gvec8_or_null, (C*)gx, nmat);
  else if (vec_dtype == L2B_F32) k_project_bwd<float><<<nblk, NTL, 0, st>>>((const C*)x, gm, (const float*)gvec8_or_null, (C*)gx, nmat);
  else if (vec_dtype == L2B_BF16) k_project_bwd<__nv_bfloat16><<<nblk, NTL, 0, st>>>((const C*)x, gm, (const __nv_bfloat16*)gvec8_or_null, (C*)gx, nmat);
  else L2B_REQUIRE(false, L2B_ERR_UNSUPPORTED, "vec_dtype must be L2B_F64, L2B_F32 or L2B_BF16");
  L2B_LAUNCHED("k_project_bwd");
  return L2B_OK;
}

int l2b_su3_force_planar(const void* u_planar, double beta, void* f_planar, int nb, const int dims[4], int dtype,
                         void* stream) {
  Geo g;
  L2B_TRY(make_geo(g, nb, dims, dtype));
  L2B_REQUIRE(u_planar && f_planar, L2B_ERR_INVALID, "null pointer");
  return launch_force(g, (const C*)u_planar, (C*)f_planar, false, beta / 3.0, nullptr, (cudaStream_t)stream);
}

int l2b_su3_project_vec_planar(const void* x_planar, void* vec8, int vec_dtype, int nb, const int dims[4], int dtype,
                               void* stream) {
  Geo g;
  L2B_TRY(make_geo(g, nb, dims, dtype));
  L2B_REQUIRE(x_planar && vec8, L2B_ERR_INVALID, "null pointer");
  const dim3 grid((g.lat.V + NTL - 1) / NTL, nb * 4);
  cudaStream_t st = (cudaStream_t)stream;
  if (vec_dtype == L2B_F64) k_project_vec_planar<double><<<grid, NTL, 0, st>>>((const C*)x_planar, (double*)vec8, g.lat.V);
  else if (vec_dtype == L2B_F32) k_project_vec_planar<float><<<grid, NTL, 0, st>>>((const C*)x_planar, (float*)vec8, g.lat.V);
  else if (vec_dtype == L2B_BF16) k_project_vec_planar<__nv_bfloat16><<<grid, NTL, 0, st>>>((const C*)x_planar, (__nv_bfloat16*)vec8, g.lat.V);
  else L2B_REQUIRE(false, L2B_ERR_UNSUPPORTED, "vec_dtype must be L2B_F64, L2B_F32 or L2B_BF16");
  L2B_LAUNCHED("k_project_vec_planar");
  return L2B_OK;
}

int l2b_su3_update_gauge_planar_pair(const void* x_planar, const void* p_planar, double eps, const double* eps_dev,
                                     const float* mask_planar, int first_complement, void* x_out_planar, int nb,
                                     const int dims[4], int dtype, void* stream) {
  Geo g;
  L2B_TRY(make_geo(g, nb, dims, dtype));
  L2B_REQUIRE(x_planar && p_planar && x_out_planar && mask_planar, L2B_ERR_INVALID, "null pointer");
  const dim3 grid((g.lat.V + NTL - 1) / NTL, nb * 4);
  k_update_gauge_planar_pair<<<grid, NTL, 0, (cudaStream_t)stream>>>((const C*)x_planar, (const C*)p_planar, eps,
                                                                     eps_dev, mask_planar, first_complement,
                                                                     (C*)x_out_planar, g.lat.V);
  L2B_LAUNCHED("k_update_gauge_planar_pair");
  return L2B_OK;
}

int l2b_su3_project_vec_planar_lm(const void* x_planar, void* vec8_lm, int nb, int nb_pad, const int dims[4], int dtype,
                                  void* stream) {
  Geo g;
  L2B_TRY(make_geo(g, nb, dims, dtype));
  L2B_REQUIRE(x_planar && vec8_lm, L2B_ERR_INVALID, "null pointer");
  L2B_REQUIRE(nb_pad >= nb && nb_pad % 8 == 0, L2B_ERR_INVALID, "nb_pad must be a multiple of 8 >= nb");
  L2B_REQUIRE(((uintptr_t)vec8_lm & 15) == 0, L2B_ERR_INVALID, "vec8_lm must be 16-byte aligned");
  const dim3 grid((g.lat.V + NTL - 1) / NTL, nb * 4);
  k_project_vec_planar_lm<<<grid, NTL, 0, (cudaStream_t)stream>>>((const C*)x_planar, (__nv_bfloat16*)vec8_lm, g.lat.V,
                                                                  nb_pad);
  L2B_LAUNCHED("k_project_vec_planar_lm");
  return L2B_OK;
}

int l2b_su3_update_gauge_planar(const void* x_planar, const void* p_planar, double eps, const double* eps_dev,
                                const float* mask_planar, int mask_complement, void* x_out_planar, int nb,
                                const int dims[4], int dtype, void* stream) {
  Geo g;
  L2B_TRY(make_geo(g, nb, dims, dtype));
  L2B_REQUIRE(x_planar && p_planar && x_out_planar, L2B_ERR_INVALID, "null pointer");
  const dim3 grid((g.lat.V + NTL - 1) / NTL, nb * 4);
  k_update_gauge_planar<<<grid, NTL, 0, (cudaStream_t)stream>>>((const C*)x_planar, (const C*)p_planar, eps, eps_dev,
                                                                mask_planar, mask_complement, (C*)x_out_planar,
                                                                g.lat.V);
  L2B_LAUNCHED("k_update_gauge_planar");
  return L2B_OK;
}

}  // extern "C"