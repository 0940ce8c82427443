// l2b_conv.cu -- the U(1) xnet's convolution stack (reference network/pytorch/network.py:151-172 `PeriodicPadding`,
// :240-346 `ConvStack`: [PeriodicPadding(n - 1), Conv2d(f, n)] blocks, MaxPool2d after every second one, activation)
// around the tensor-core GEMM of l2b_gemm.cu.  A convolution is  col[M, K] . W2d[Cout, K]^T  with
//     M = (chain, oh, ow),  OH = H + n - 1 (the reference pads n - 1 on BOTH sides and convolves "valid"),
//     K = (ci, kh, kw)  in the order of Conv2d's own weight [Cout, Cin, n, n] -- no weight permutation,
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
};

// one thread per (row m, group of 8 columns); NT = 1: bf16 col, NT = 3: bf16x3 planes col[t][M][K8]
template <typename TIN, int NT>
__global__ void __launch_bounds__(256) k_im2col_periodic(const TIN* __restrict__ in, const ConvGeo g,
                                                         __nv_bfloat16* __restrict__ col, long long M) {
  const long long id = (long long)blockIdx.x * 256 + threadIdx.x;
  const int kg8 = g.K8 / 8;
  if (id >= M * kg8) return;
  const long long m = id / kg8;
  const int k0 = (int)(id % kg8) * 8;
  const int ow = (int)(m % g.OW);
  const int oh = (int)((m / g.OW) % g.OH);
  const long long b = m / ((long long)g.OW * g.OH);
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

// adjoint of the gather: din[b, ci, h, w] = sum over (kh, kw) and the output pixels (oh, ow) that read (h, w) through
// that tap: oh = (h - kh + n - 1) mod H, and oh + H if that is still < OH (the padded image repeats the input).
// One thread per input element, output strides free (NCHW for the network input, NHWC between layers).
template <typename TG>
__global__ void __launch_bounds__(256) k_col2im_periodic(const TG* __restrict__ dcol, long long ldc, const ConvGeo g,
                                                         float* __restrict__ din, long long ob, long long oc, long long oh_,
                                                         long long ow_) {
  const long long id = (long long)blockIdx.x * 256 + threadIdx.x;
  const long long total = (long long)g.nb * g.H * g.W * g.C;
  if (id >= total) return;
  const int ci = (int)(id % g.C);
  const int w = (int)((id / g.C) % g.W);
  const int h = (int)((id / ((long long)g.C * g.W)) % g.H);
  const long long b = id / ((long long)g.C * g.W * g.H);
  float acc = 0.f;
  for (int kh = 0; kh < g.n; ++kh) {
    int oh0 = h - kh + g.n - 1;                       // in (-H, 2H): the modulo is one conditional add / subtract
    oh0 += oh0 < 0 ? g.H : 0; oh0 -= oh0 >= g.H ? g.H : 0;
    for (int kw = 0; kw < g.n; ++kw) {
      int ow0 = w - kw + g.n - 1;
      ow0 += ow0 < 0 ? g.W : 0; ow0 -= ow0 >= g.W ? g.W : 0;
      const int k = (ci * g.n + kh) * g.n + kw;
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
                                                  float* __restrict__ pre) {
  const int PH = H / p, PW = W / p;
  const long long id = (long long)blockIdx.x * 256 + threadIdx.x;
  if (id >= (long long)nb * PH * PW * C) return;
  const int c = (int)(id % C);
  const int pw = (int)((id / C) % PW);
  const int ph = (int)((id / ((long long)C * PW)) % PH);
  const long long b = id / ((long long)C * PW * PH);
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
                                                      int nb, int H, int W, int C, int p, int act, float* __restrict__ gx) {
  const int PH = H / p, PW = W / p;
  const long long id = (long long)blockIdx.x * 256 + threadIdx.x;
  if (id >= (long long)nb * PH * PW * C) return;
  const int c = (int)(id % C);
  const int pw = (int)((id / C) % PW);
  const int ph = (int)((id / ((long long)C * PW)) % PH);
  const long long b = id / ((long long)C * PW * PH);
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
  return L2B_OK;
}

int l2b_conv_im2col(const void* in, int in_dtype, int nb, int C, int H, int W, int n, const long long strides[4],
                    void* col, int planes, void* stream) {
  L2B_REQUIRE(in && col && strides, L2B_ERR_INVALID, "null pointer");
  L2B_REQUIRE(planes >= 1 && planes <= 3, L2B_ERR_INVALID, "planes must be 1 (bf16), 2 (bf16x2) or 3 (bf16x3)");
  L2B_REQUIRE(in_dtype == L2B_F32 || in_dtype == L2B_BF16, L2B_ERR_UNSUPPORTED, "in_dtype must be L2B_F32 or L2B_BF16");
  L2B_REQUIRE(((uintptr_t)col & 15) == 0, L2B_ERR_INVALID, "col must be 16-byte aligned");
  ConvGeo g;
  const int rc = conv_geo(g, nb, C, H, W, n, strides);
  if (rc != L2B_OK) return rc;
  const long long M = (long long)nb * g.OH * g.OW, total = M * (g.K8 / 8);
  const unsigned nblk = (unsigned)((total + 255) / 256);
  cudaStream_t st = (cudaStream_t)stream;
  __nv_bfloat16* c = (__nv_bfloat16*)col;
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
                    const long long out_strides[4], void* stream) {
  L2B_REQUIRE(dcol && din && out_strides, L2B_ERR_INVALID, "null pointer");
  L2B_REQUIRE(dcol_dtype == L2B_F32 || dcol_dtype == L2B_BF16, L2B_ERR_UNSUPPORTED, "dcol_dtype must be L2B_F32 or L2B_BF16");
  ConvGeo g;
  const long long dummy[4] = {0, 0, 0, 0};
  const int rc = conv_geo(g, nb, C, H, W, n, dummy);
  if (rc != L2B_OK) return rc;
  L2B_REQUIRE(ldc >= g.K, L2B_ERR_INVALID, "ldc must be >= C n^2");
  const long long total = (long long)nb * H * W * C;
  const unsigned nblk = (unsigned)((total + 255) / 256);
  cudaStream_t st = (cudaStream_t)stream;
  if (dcol_dtype == L2B_F32)
    k_col2im_periodic<float><<<nblk, 256, 0, st>>>((const float*)dcol, ldc, g, din, out_strides[0], out_strides[1],
                                                   out_strides[2], out_strides[3]);
  else
    k_col2im_periodic<__nv_bfloat16><<<nblk, 256, 0, st>>>((const __nv_bfloat16*)dcol, ldc, g, din, out_strides[0],
                                                           out_strides[1], out_strides[2], out_strides[3]);
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
  const unsigned nblk = (unsigned)((total + 255) / 256);
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == L2B_F32)
    k_pool_act<float><<<nblk, 256, 0, st>>>((const float*)x, nb, H, W, C, pool, activation, (float*)y, idx, pre);
  else
    k_pool_act<__nv_bfloat16><<<nblk, 256, 0, st>>>((const __nv_bfloat16*)x, nb, H, W, C, pool, activation,
                                                    (__nv_bfloat16*)y, idx, pre);
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
  const unsigned nblk = (unsigned)((total + 255) / 256);
  if (dtype == L2B_F32)
    k_pool_act_bwd<float><<<nblk, 256, 0, st>>>(gy, (const float*)y, pre, idx, nb, H, W, C, pool, activation, gx);
  else
    k_pool_act_bwd<__nv_bfloat16><<<nblk, 256, 0, st>>>(gy, (const __nv_bfloat16*)y, pre, idx, nb, H, W, C, pool,
                                                        activation, gx);
  L2B_LAUNCHED("k_pool_act_bwd");
  return L2B_OK;
}

}  // extern "C"
