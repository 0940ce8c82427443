// l2b_conv.cu -- the U(1) xnet's convolution stack (reference network/pytorch/network.py:151-172 `PeriodicPadding`,
// :240-346 `ConvStack`: [PeriodicPadding(n - 1), Conv2d(f, n)] blocks, MaxPool2d after every second one, activation)
// around the tensor-core GEMM of l2b_gemm.cu.  A convolution is  col[M, K] . W2d[Cout, K]^T  with
//     M = (chain, oh, ow),  OH = H + n - 1 (the reference pads n - 1 on BOTH sides and convolves "valid"),
//     K = (ci, kh, kw)  in the order of Conv2d's own weight [Cout, Cin, n, n] -- no weight permutation (k_order 0),
//       or (kh, kw, ci), channel fastest (k_order 1, "tap-major": the weight is used as [Cout, n, n, Cin]): on NHWC
//       activations eight consecutive columns are then eight consecutive channels of ONE tap -- one wrap and one
//       32-byte load per thread instead of eight wraps and eight scattered 4-byte loads (the blocks after the first),
//     col[(b, oh, ow)][(ci, kh, kw)] = in[b, ci, (oh + kh - n + 1) mod H, (ow + kw - n + 1) mod W]:
// the periodic padding is an index wrap inside the gather, never a tensor.  The GEMM's output [M, Cout] is the next
// layer's input in NHWC.  Kernels here: the gather (k_im2col_periodic: bf16, or the bf16x3 split of fp32 nets, written
// directly), its adjoint as a gather as well (k_col2im_periodic: every input pixel collects its <= 4 n^2 contributions
// in a fixed order, no atomics), and max pooling with the activation that follows it, forward and backward.
#include <cuda_bf16.h>

#include "l2b_common.cuh"
#include "l2b_tc.cuh"

namespace l2b {
namespace {

template <typename TIN> __device__ __forceinline__ float ld_f(const TIN* p);
template <> __device__ __forceinline__ float ld_f<float>(const float* p) { return __ldg(p); }
template <> __device__ __forceinline__ float ld_f<__nv_bfloat16>(const __nv_bfloat16* p) { return __bfloat162float(*p); }

struct ConvGeo {
  int nb, C, H, W, n;                 // input channels / extent, kernel size
  long long sb, sc, sh, sw;           // element strides of the input (NCHW network input or NHWC activations)
  int OH, OW, K, K8;
  int small;                          // every flat index of the launch fits 31 bits: 32-bit index arithmetic
};

// (q, r) = (id / d, id % d).  A 64-bit division is ~100 instructions on the GPU and these kernels do three to five of
// them per thread for a handful of loads; every shipped size fits 32 bits (flag computed on the host).
__device__ __forceinline__ void divmod(long long id, int d, int small, long long& q, int& r) {
  if (small) {
    const unsigned u = (unsigned)id, qq = u / (unsigned)d;
    q = qq; r = (int)(u - qq * (unsigned)d);
  } else {
    q = id / d; r = (int)(id - q * d);
  }
}

// one thread per (row m, group of 8 columns); NT = 1: bf16 col, NT = 3: bf16x3 planes col[t][M][K8]
template <typename TIN, int NT>
__global__ void __launch_bounds__(256) k_im2col_periodic(const TIN* __restrict__ in, const ConvGeo g,
                                                         __nv_bfloat16* __restrict__ col, long long M) {
  const long long id = (long long)blockIdx.x * 256 + threadIdx.x;
  const int kg8 = g.K8 / 8;
  if (id >= M * kg8) return;
  long long m, rest, b;
  int k0, ow, oh;
  divmod(id, kg8, g.small, m, k0); k0 *= 8;
  divmod(m, g.OW, g.small, rest, ow);
  divmod(rest, g.OH, g.small, b, oh);
  const int n2 = g.n * g.n;
  __align__(16) __nv_bfloat16 h[NT][8];
  // (ci, kh, kw) of the first column by division, of the next seven by counting; the wrapped source coordinates lie
  // in (-H, 2H), so one conditional add / subtract replaces the modulo
  int ci = k0 / n2, r0 = k0 - ci * n2, kh = r0 / g.n, kw = r0 - kh * g.n;
  const TIN* src_b = in + b * g.sb;
  const int oh1 = oh - (g.n - 1), ow1 = ow - (g.n - 1);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    float v = 0.f;
    if (k0 + i < g.K) {
      int ih = oh1 + kh, iw = ow1 + kw;
      ih += ih < 0 ? g.H : 0; ih -= ih >= g.H ? g.H : 0;
      iw += iw < 0 ? g.W : 0; iw -= iw >= g.W ? g.W : 0;
      v = ld_f<TIN>(src_b + ci * g.sc + ih * g.sh + iw * g.sw);
    }
    if (++kw == g.n) { kw = 0; if (++kh == g.n) { kh = 0; ++ci; } }
    const __nv_bfloat16 a = __float2bfloat16(v);
    h[0][i] = a;
    if (NT >= 2) {
      const float r1 = v - __bfloat162float(a);
      const __nv_bfloat16 b2 = __float2bfloat16(r1);
      h[1][i] = b2;
      if (NT == 3) h[2][i] = __float2bfloat16(r1 - __bfloat162float(b2));
    }
  }
  const size_t plane = (size_t)M * g.K8;
#pragma unroll
  for (int t = 0; t < NT; ++t)
    *reinterpret_cast<uint4*>(col + t * plane + (size_t)m * g.K8 + k0) = *reinterpret_cast<const uint4*>(h[t]);
}

// tap-major columns k = (kh n + kw) C + ci.  VEC: C % 8 == 0, channel stride 1 and a 32-byte (fp32) / 16-byte (bf16)
// aligned input: the thread's eight columns are eight consecutive channels of one tap, fetched with vector loads.
template <typename TIN, int NT, bool VEC>
__global__ void __launch_bounds__(256) k_im2col_periodic_tap(const TIN* __restrict__ in, const ConvGeo g,
                                                             __nv_bfloat16* __restrict__ col, long long M) {
  const long long id = (long long)blockIdx.x * 256 + threadIdx.x;
  const int kg8 = g.K8 / 8;
  if (id >= M * kg8) return;
  long long m, rest, b;
  int k0, ow, oh;
  divmod(id, kg8, g.small, m, k0); k0 *= 8;
  divmod(m, g.OW, g.small, rest, ow);
  divmod(rest, g.OH, g.small, b, oh);
  int tap = k0 / g.C, ci = k0 - tap * g.C, kh = tap / g.n, kw = tap - kh * g.n;
  const TIN* src_b = in + b * g.sb;
  const int oh1 = oh - (g.n - 1), ow1 = ow - (g.n - 1);
  float v[8];
  if (VEC) {
    int ih = oh1 + kh, iw = ow1 + kw;
    ih += ih < 0 ? g.H : 0; ih -= ih >= g.H ? g.H : 0;
    iw += iw < 0 ? g.W : 0; iw -= iw >= g.W ? g.W : 0;
    const TIN* s = src_b + ih * g.sh + iw * g.sw + ci;
    if (sizeof(TIN) == 4) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(s)), c = __ldg(reinterpret_cast<const float4*>(s) + 1);
      v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = c.x; v[5] = c.y; v[6] = c.z; v[7] = c.w;
    } else {
      const uint4 raw = __ldg(reinterpret_cast<const uint4*>(s));
      const __nv_bfloat16* hb = reinterpret_cast<const __nv_bfloat16*>(&raw);
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = __bfloat162float(hb[i]);
    }
  } else {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      v[i] = 0.f;
      if (k0 + i < g.K) {
        int ih = oh1 + kh, iw = ow1 + kw;
        ih += ih < 0 ? g.H : 0; ih -= ih >= g.H ? g.H : 0;
        iw += iw < 0 ? g.W : 0; iw -= iw >= g.W ? g.W : 0;
        v[i] = ld_f<TIN>(src_b + ci * g.sc + ih * g.sh + iw * g.sw);
      }
      if (++ci == g.C) { ci = 0; if (++kw == g.n) { kw = 0; ++kh; } }
    }
  }
  __align__(16) __nv_bfloat16 h[NT][8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const __nv_bfloat16 a = __float2bfloat16(v[i]);
    h[0][i] = a;
    if (NT >= 2) {
      const float r1 = v[i] - __bfloat162float(a);
      const __nv_bfloat16 b2 = __float2bfloat16(r1);
      h[1][i] = b2;
      if (NT == 3) h[2][i] = __float2bfloat16(r1 - __bfloat162float(b2));
    }
  }
  const size_t plane = (size_t)M * g.K8;
#pragma unroll
  for (int t = 0; t < NT; ++t)
    *reinterpret_cast<uint4*>(col + t * plane + (size_t)m * g.K8 + k0) = *reinterpret_cast<const uint4*>(h[t]);
}

// adjoint of the gather: din[b, ci, h, w] = sum over (kh, kw) and the output pixels (oh, ow) that read (h, w) through
// that tap: oh = (h - kh + n - 1) mod H, and oh + H if that is still < OH (the padded image repeats the input).
// One thread per input element, output strides free (NCHW for the network input, NHWC between layers).
// TAP: columns in tap-major order (k = (kh n + kw) C + ci): consecutive threads (ci fastest) read consecutive columns.
template <typename TG, bool TAP>
__global__ void __launch_bounds__(256) k_col2im_periodic(const TG* __restrict__ dcol, long long ldc, const ConvGeo g,
                                                         float* __restrict__ din, long long ob, long long oc, long long oh_,
                                                         long long ow_) {
  const long long id = (long long)blockIdx.x * 256 + threadIdx.x;
  const long long total = (long long)g.nb * g.H * g.W * g.C;
  if (id >= total) return;
  long long r1, r2, b;
  int ci, w, h;
  divmod(id, g.C, g.small, r1, ci);
  divmod(r1, g.W, g.small, r2, w);
  divmod(r2, g.H, g.small, b, h);
  float acc = 0.f;
  for (int kh = 0; kh < g.n; ++kh) {
    int oh0 = h - kh + g.n - 1;                       // in (-H, 2H): the modulo is one conditional add / subtract
    oh0 += oh0 < 0 ? g.H : 0; oh0 -= oh0 >= g.H ? g.H : 0;
    for (int kw = 0; kw < g.n; ++kw) {
      int ow0 = w - kw + g.n - 1;
      ow0 += ow0 < 0 ? g.W : 0; ow0 -= ow0 >= g.W ? g.W : 0;
      const int k = TAP ? (kh * g.n + kw) * g.C + ci : (ci * g.n + kh) * g.n + kw;
      for (int oh = oh0; oh < g.OH; oh += g.H)
        for (int ow = ow0; ow < g.OW; ow += g.W)
          acc += ld_f<TG>(dcol + ((b * g.OH + oh) * g.OW + ow) * ldc + k);
    }
  }
  din[b * ob + ci * oc + h * oh_ + w * ow_] = acc;
}

// MaxPool2d(p) (floor) on NHWC + the activation that follows it in the stack; idx = winning tap (first maximum, as
// ATen) for the adjoint; `pre` keeps the pooled pre-activation when the activation's derivative needs it (swish)
template <typename T>
__global__ void __launch_bounds__(256) k_pool_act(const T* __restrict__ x, int nb, int H, int W, int C, int p, int act,
                                                  T* __restrict__ y, unsigned char* __restrict__ idx,
                                                  float* __restrict__ pre, int small) {
  const int PH = H / p, PW = W / p;
  const long long id = (long long)blockIdx.x * 256 + threadIdx.x;
  if (id >= (long long)nb * PH * PW * C) return;
  long long r1, r2, b;
  int c, pw, ph;
  divmod(id, C, small, r1, c);
  divmod(r1, PW, small, r2, pw);
  divmod(r2, PH, small, b, ph);
  float best = -INFINITY;
  int arg = 0;
  for (int i = 0; i < p; ++i)
    for (int j = 0; j < p; ++j) {
      const float v = ld_f<T>(x + ((b * H + ph * p + i) * W + pw * p + j) * C + c);
      if (v > best || (i == 0 && j == 0)) { best = v; arg = i * p + j; }
    }
  idx[id] = (unsigned char)arg;
  if (pre) pre[id] = best;
  const float o = il_act(best, act);
  if (sizeof(T) == 2) reinterpret_cast<__nv_bfloat16*>(y)[id] = __float2bfloat16(o);
  else reinterpret_cast<float*>(y)[id] = o;
}

// adjoint: gx (NHWC, zero elsewhere) gets gy * act'(.) at the winning tap
template <typename T>
__global__ void __launch_bounds__(256) k_pool_act_bwd(const float* __restrict__ gy, const T* __restrict__ y,
                                                      const float* __restrict__ pre, const unsigned char* __restrict__ idx,
                                                      int nb, int H, int W, int C, int p, int act, float* __restrict__ gx,
                                                      int small) {
  const int PH = H / p, PW = W / p;
  const long long id = (long long)blockIdx.x * 256 + threadIdx.x;
  if (id >= (long long)nb * PH * PW * C) return;
  long long r1, r2, b;
  int c, pw, ph;
  divmod(id, C, small, r1, c);
  divmod(r1, PW, small, r2, pw);
  divmod(r2, PH, small, b, ph);
  const float yo = ld_f<T>(y + id);
  float d = 1.f;
  switch (act) {
    case 1: d = 1.f - yo * yo; break;
    case 2: d = yo > 0.f ? 1.f : 0.f; break;
    case 3: { const float pz = pre[id], sg = 1.f / (1.f + __expf(-pz)); d = sg * (1.f + pz * (1.f - sg)); } break;
    case 4: d = yo > 0.f ? 1.f : 0.01f; break;
    case 5: d = yo > 0.f ? 1.f : yo + 1.f; break;
    default: break;
  }
  const int arg = idx[id], i = arg / p, j = arg - i * p;
  gx[((b * H + ph * p + i) * W + pw * p + j) * C + c] = gy[id] * d;
}

}  // namespace
}  // namespace l2b

using namespace l2b;

extern "C" {

static int conv_geo(ConvGeo& g, int nb, int C, int H, int W, int n, const long long strides[4]) {
  L2B_REQUIRE(nb > 0 && C > 0 && H > 0 && W > 0 && n > 0, L2B_ERR_INVALID, "nb, C, H, W, n must be positive");
  L2B_REQUIRE(n - 1 <= H && n - 1 <= W, L2B_ERR_UNSUPPORTED, "periodic padding n - 1 = %d exceeds the image (%d x %d)",
              n - 1, H, W);
  g.nb = nb; g.C = C; g.H = H; g.W = W; g.n = n;
  g.sb = strides[0]; g.sc = strides[1]; g.sh = strides[2]; g.sw = strides[3];
  g.OH = H + n - 1; g.OW = W + n - 1;
  g.K = C * n * n;
  g.K8 = (g.K + 7) / 8 * 8;
  // flat indices of the gather ((nb OH OW) x K8 / 8 threads) and of its adjoint (nb H W C threads)
  g.small = (long long)nb * g.OH * g.OW * (g.K8 / 8) < (1ll << 31) && (long long)nb * H * W * C < (1ll << 31);
  return L2B_OK;
}

int l2b_conv_im2col(const void* in, int in_dtype, int nb, int C, int H, int W, int n, const long long strides[4],
                    void* col, int planes, int k_order, void* stream) {
  L2B_REQUIRE(in && col && strides, L2B_ERR_INVALID, "null pointer");
  L2B_REQUIRE(planes >= 1 && planes <= 3, L2B_ERR_INVALID, "planes must be 1 (bf16), 2 (bf16x2) or 3 (bf16x3)");
  L2B_REQUIRE(k_order == 0 || k_order == 1, L2B_ERR_INVALID, "k_order must be 0 (ci, kh, kw) or 1 (kh, kw, ci)");
  L2B_REQUIRE(in_dtype == L2B_F32 || in_dtype == L2B_BF16, L2B_ERR_UNSUPPORTED, "in_dtype must be L2B_F32 or L2B_BF16");
  L2B_REQUIRE(((uintptr_t)col & 15) == 0, L2B_ERR_INVALID, "col must be 16-byte aligned");
  ConvGeo g;
  const int rc = conv_geo(g, nb, C, H, W, n, strides);
  if (rc != L2B_OK) return rc;
  const long long M = (long long)nb * g.OH * g.OW, total = M * (g.K8 / 8);
  const unsigned nblk = (unsigned)((total + 255) / 256);
  cudaStream_t st = (cudaStream_t)stream;
  __nv_bfloat16* c = (__nv_bfloat16*)col;
  if (k_order == 1) {
    // vector path: eight consecutive channels of one tap are contiguous and aligned
    const size_t esz = in_dtype == L2B_F32 ? 4 : 2;
    const bool vec = C % 8 == 0 && g.sc == 1 && ((uintptr_t)in & (8 * esz - 1)) == 0 && g.sb % 8 == 0 &&
                     g.sh % 8 == 0 && g.sw % 8 == 0;
#define L2B_TAP(TIN, NT)                                                                                       \
  do {                                                                                                         \
    if (vec) k_im2col_periodic_tap<TIN, NT, true><<<nblk, 256, 0, st>>>((const TIN*)in, g, c, M);              \
    else k_im2col_periodic_tap<TIN, NT, false><<<nblk, 256, 0, st>>>((const TIN*)in, g, c, M);                 \
  } while (0)
    if (in_dtype == L2B_F32) {
      if (planes == 3) L2B_TAP(float, 3);
      else if (planes == 2) L2B_TAP(float, 2);
      else L2B_TAP(float, 1);
    } else {
      if (planes == 3) L2B_TAP(__nv_bfloat16, 3);
      else if (planes == 2) L2B_TAP(__nv_bfloat16, 2);
      else L2B_TAP(__nv_bfloat16, 1);
    }
#undef L2B_TAP
    L2B_LAUNCHED("k_im2col_periodic_tap");
    return L2B_OK;
  }
  if (in_dtype == L2B_F32) {
    if (planes == 3) k_im2col_periodic<float, 3><<<nblk, 256, 0, st>>>((const float*)in, g, c, M);
    else if (planes == 2) k_im2col_periodic<float, 2><<<nblk, 256, 0, st>>>((const float*)in, g, c, M);
    else k_im2col_periodic<float, 1><<<nblk, 256, 0, st>>>((const float*)in, g, c, M);
  } else {
    if (planes == 3) k_im2col_periodic<__nv_bfloat16, 3><<<nblk, 256, 0, st>>>((const __nv_bfloat16*)in, g, c, M);
    else if (planes == 2) k_im2col_periodic<__nv_bfloat16, 2><<<nblk, 256, 0, st>>>((const __nv_bfloat16*)in, g, c, M);
    else k_im2col_periodic<__nv_bfloat16, 1><<<nblk, 256, 0, st>>>((const __nv_bfloat16*)in, g, c, M);
  }
  L2B_LAUNCHED("k_im2col_periodic");
  return L2B_OK;
}

int l2b_conv_col2im(const void* dcol, int dcol_dtype, long long ldc, int nb, int C, int H, int W, int n, float* din,
                    const long long out_strides[4], int k_order, void* stream) {
  L2B_REQUIRE(dcol && din && out_strides, L2B_ERR_INVALID, "null pointer");
  L2B_REQUIRE(k_order == 0 || k_order == 1, L2B_ERR_INVALID, "k_order must be 0 (ci, kh, kw) or 1 (kh, kw, ci)");
  L2B_REQUIRE(dcol_dtype == L2B_F32 || dcol_dtype == L2B_BF16, L2B_ERR_UNSUPPORTED, "dcol_dtype must be L2B_F32 or L2B_BF16");
  ConvGeo g;
  const long long dummy[4] = {0, 0, 0, 0};
  const int rc = conv_geo(g, nb, C, H, W, n, dummy);
  if (rc != L2B_OK) return rc;
  L2B_REQUIRE(ldc >= g.K, L2B_ERR_INVALID, "ldc must be >= C n^2");
  const long long total = (long long)nb * H * W * C;
  const unsigned nblk = (unsigned)((total + 255) / 256);
  cudaStream_t st = (cudaStream_t)stream;
  const long long* o = out_strides;
  if (dcol_dtype == L2B_F32) {
    if (k_order) k_col2im_periodic<float, true><<<nblk, 256, 0, st>>>((const float*)dcol, ldc, g, din, o[0], o[1], o[2], o[3]);
    else k_col2im_periodic<float, false><<<nblk, 256, 0, st>>>((const float*)dcol, ldc, g, din, o[0], o[1], o[2], o[3]);
  } else {
    if (k_order) k_col2im_periodic<__nv_bfloat16, true><<<nblk, 256, 0, st>>>((const __nv_bfloat16*)dcol, ldc, g, din, o[0], o[1], o[2], o[3]);
    else k_col2im_periodic<__nv_bfloat16, false><<<nblk, 256, 0, st>>>((const __nv_bfloat16*)dcol, ldc, g, din, o[0], o[1], o[2], o[3]);
  }
  L2B_LAUNCHED("k_col2im_periodic");
  return L2B_OK;
}

int l2b_pool_act(const void* x, int dtype, int nb, int H, int W, int C, int pool, int activation, void* y,
                 unsigned char* idx, float* pre, void* stream) {
  L2B_REQUIRE(x && y && idx, L2B_ERR_INVALID, "null pointer");
  L2B_REQUIRE(dtype == L2B_F32 || dtype == L2B_BF16, L2B_ERR_UNSUPPORTED, "dtype must be L2B_F32 or L2B_BF16");
  L2B_REQUIRE(pool >= 1 && pool <= 15 && H / pool > 0 && W / pool > 0, L2B_ERR_UNSUPPORTED, "pool window %d does not fit", pool);
  L2B_REQUIRE(activation >= 0 && activation <= 5, L2B_ERR_INVALID, "activation code must be in [0, 5]");
  L2B_REQUIRE(activation != 3 || pre != nullptr, L2B_ERR_INVALID, "swish needs the pre-activation buffer");
  const long long total = (long long)nb * (H / pool) * (W / pool) * C;
  const int small = (long long)nb * H * W * C < (1ll << 31);
  const unsigned nblk = (unsigned)((total + 255) / 256);
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == L2B_F32)
    k_pool_act<float><<<nblk, 256, 0, st>>>((const float*)x, nb, H, W, C, pool, activation, (float*)y, idx, pre, small);
  else
    k_pool_act<__nv_bfloat16><<<nblk, 256, 0, st>>>((const __nv_bfloat16*)x, nb, H, W, C, pool, activation,
                                                    (__nv_bfloat16*)y, idx, pre, small);
  L2B_LAUNCHED("k_pool_act");
  return L2B_OK;
}

int l2b_pool_act_bwd(const float* gy, const void* y, int dtype, const float* pre, const unsigned char* idx, int nb, int H,
                     int W, int C, int pool, int activation, float* gx, void* stream) {
  L2B_REQUIRE(gy && y && idx && gx, L2B_ERR_INVALID, "null pointer");
  L2B_REQUIRE(dtype == L2B_F32 || dtype == L2B_BF16, L2B_ERR_UNSUPPORTED, "dtype must be L2B_F32 or L2B_BF16");
  L2B_REQUIRE(pool >= 1 && pool <= 15 && H / pool > 0 && W / pool > 0, L2B_ERR_UNSUPPORTED, "pool window %d does not fit", pool);
  L2B_REQUIRE(activation != 3 || pre != nullptr, L2B_ERR_INVALID, "swish needs the pre-activation buffer");
  cudaStream_t st = (cudaStream_t)stream;
  L2B_CUDA(cudaMemsetAsync(gx, 0, sizeof(float) * (size_t)nb * H * W * C, st));
  const long long total = (long long)nb * (H / pool) * (W / pool) * C;
  const int small = (long long)nb * H * W * C < (1ll << 31);
  const unsigned nblk = (unsigned)((total + 255) / 256);
  if (dtype == L2B_F32)
    k_pool_act_bwd<float><<<nblk, 256, 0, st>>>(gy, (const float*)y, pre, idx, nb, H, W, C, pool, activation, gx, small);
  else
    k_pool_act_bwd<__nv_bfloat16><<<nblk, 256, 0, st>>>(gy, (const __nv_bfloat16*)y, pre, idx, nb, H, W, C, pool,
                                                        activation, gx, small);
  L2B_LAUNCHED("k_pool_act_bwd");
  return L2B_OK;
}

}  // extern "C"
