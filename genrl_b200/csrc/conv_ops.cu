// Conv encoder / transposed-conv decoder lowering (reference: Encoder dreamer_utils.py:580-621, Decoder :656-706).
// Activations are NHWC so the channel LayerNorm is a row LayerNorm and the GEMM operands are K-contiguous.
//   Conv2d(k, s=2)          : col = im2col(x)            -> y = col * Wp^T        (Wp = [Cout, (kh,kw,Cin)])
//   ConvTranspose2d(k, s=2) : col = x * Wt               -> y = col2im(col) + b   (Wt = [Cin, (kh,kw,Cout)])
// and the adjoints swap roles (conv dgrad = col2im, convT dgrad = im2col).  The uint8 NCHW -> (x/255 - 0.5)
// preprocessing (dreamer.py:294-295) is fused into the first im2col.
#include "common.cuh"
#include "rowops_fast.cuh"
#include <algorithm>

namespace genrl {

// col[(n,oh,ow), (kh,kw,c)] = x[n, 2*oh+kh, 2*ow+kw, c]       x: NHWC (T) or NCHW uint8 (preprocessed on the fly)
template <typename TI, typename TO, bool U8_NCHW>
__global__ void __launch_bounds__(256) im2col_kernel(const TI* __restrict__ x, TO* __restrict__ col, long long ldcol, int N, int H, int W, int C,
                                                     int k, int Ho, int Wo) {
  genrl::pdl_sync();
  const long long total = (long long)N * Ho * Wo * k * k * C;
  const int Kc = k * k * C;
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
    int kc = (int)(i % Kc);
    long long m = i / Kc;
    int c = kc % C; int kw = (kc / C) % k; int kh = kc / (C * k);
    int ow = (int)(m % Wo); int oh = (int)((m / Wo) % Ho); long long n = m / ((long long)Wo * Ho);
    int ih = 2 * oh + kh, iw = 2 * ow + kw;
    float v;
    if (U8_NCHW) v = (float)reinterpret_cast<const uint8_t*>(x)[((n * C + c) * H + ih) * W + iw] / 255.0f - 0.5f;
    else v = to_f32(x[((n * H + ih) * W + iw) * C + c]);
    col[m * ldcol + kc] = from_f32<TO>(v);
  }
}

// y[n, oh, ow, c] = bias[c] + sum_{kh,kw : (oh-kh),(ow-kw) even, in range} col[(n,(oh-kh)/2,(ow-kw)/2), (kh,kw,c)]
// NCHW_F32_OUT writes y as fp32 [N, C, Ho, Wo] (final decoder layer: the API returns NCHW means).
template <typename TI, typename TO, bool NCHW_OUT>
__global__ void __launch_bounds__(256) col2im_kernel(const TI* __restrict__ col, long long ldcol, const float* __restrict__ bias, TO* __restrict__ y,
                                                     int N, int Hi, int Wi, int C, int k, int Ho, int Wo) {
  genrl::pdl_sync();
  const long long total = (long long)N * Ho * Wo * C;
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
    int c, ow, oh; long long n;
    if (NCHW_OUT) { ow = (int)(i % Wo); oh = (int)((i / Wo) % Ho); c = (int)((i / ((long long)Wo * Ho)) % C); n = i / ((long long)Wo * Ho * C); }
    else { c = (int)(i % C); ow = (int)((i / C) % Wo); oh = (int)((i / ((long long)C * Wo)) % Ho); n = i / ((long long)C * Wo * Ho); }
    float s = bias ? bias[c] : 0.f;
    for (int kh = oh & 1; kh < k; kh += 2) {
      int ih = (oh - kh) >> 1;
      if (ih < 0 || ih >= Hi || oh < kh) continue;
      for (int kw = ow & 1; kw < k; kw += 2) {
        int iw = (ow - kw) >> 1;
        if (iw < 0 || iw >= Wi || ow < kw) continue;
        s += to_f32(col[((n * Hi + ih) * Wi + iw) * ldcol + (kh * k + kw) * C + c]);
      }
    }
    y[i] = from_f32<TO>(s);
  }
}

// NCHW fp32 [N,C,H,W] -> NHWC T  (gradient of the final decoder layer arrives as NCHW)
template <typename TO>
__global__ void __launch_bounds__(256) nchw_to_nhwc_kernel(const float* __restrict__ x, TO* __restrict__ y, int N, int C, int H, int W) {
  genrl::pdl_sync();
  const long long total = (long long)N * C * H * W;
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
    int c = (int)(i % C); int w = (int)((i / C) % W); int h = (int)((i / ((long long)C * W)) % H); long long n = i / ((long long)C * W * H);
    y[i] = from_f32<TO>(x[((n * C + c) * H + h) * W + w]);
  }
}

// generic [A, B, P] <-> [A, P, B] permutation with cast; forward (pack) overwrites, backward (unpack) accumulates into fp32
template <typename TI, typename TO>
__global__ void __launch_bounds__(256) permute_abp_kernel(const TI* __restrict__ src, TO* __restrict__ dst, long long A, int B, int P, int accumulate) {
  genrl::pdl_sync();
  // src is [A, B, P]; dst is [A, P, B]
  const long long total = A * B * P;
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
    int b = (int)(i % B); int p = (int)((i / B) % P); long long a = i / ((long long)B * P);
    float v = to_f32(src[(a * B + b) * P + p]);
    if (accumulate) dst[i] = from_f32<TO>(to_f32(dst[i]) + v); else dst[i] = from_f32<TO>(v);
  }
}


// ---- fast paths -------------------------------------------------------------------------------------------------
// For a fixed output pixel m and filter row kh the k*C elements (kw, c) are contiguous in NHWC memory AND in the col row,
// so im2col is a set of contiguous segment copies: one 8-element vector per thread when C % 8 == 0.
template <typename T>
__global__ void __launch_bounds__(256) im2col_vec_kernel(const T* __restrict__ x, T* __restrict__ col, long long ldcol, int N, int H, int W, int C,
                                                         int k, int Ho, int Wo) {
  genrl::pdl_sync();
  const int seg8 = k * C / 8;
  const long long total = (long long)N * Ho * Wo * k * seg8;
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
    const int q = (int)(i % seg8);
    const int kh = (int)((i / seg8) % k);
    const long long m = i / ((long long)seg8 * k);
    const int ow = (int)(m % Wo); const int oh = (int)((m / Wo) % Ho); const long long n = m / ((long long)Wo * Ho);
    float v[8];
    V8<T>::load(x + ((n * H + 2 * oh + kh) * W + 2 * ow) * C + q * 8, v);
    V8<T>::store(col + m * ldcol + (long long)kh * k * C + q * 8, v);
  }
}

// generic segment kernel: one thread per (m, kh), scalar inner loop (first layer uint8 NCHW, 3-channel images)
template <typename TI, typename TO, bool U8_NCHW>
__global__ void __launch_bounds__(256) im2col_seg_kernel(const TI* __restrict__ x, TO* __restrict__ col, long long ldcol, int N, int H, int W, int C,
                                                         int k, int Ho, int Wo) {
  genrl::pdl_sync();
  const long long total = (long long)N * Ho * Wo * k;
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
    const int kh = (int)(i % k);
    const long long m = i / k;
    const int ow = (int)(m % Wo); const int oh = (int)((m / Wo) % Ho); const long long n = m / ((long long)Wo * Ho);
    const int ih = 2 * oh + kh, iw0 = 2 * ow;
    TO* dst = col + m * ldcol + (long long)kh * k * C;
    if (U8_NCHW) {
      const uint8_t* xb = reinterpret_cast<const uint8_t*>(x);
      for (int c = 0; c < C; ++c) {
        const uint8_t* src = xb + ((n * C + c) * H + ih) * W + iw0;
        for (int kw = 0; kw < k; ++kw) dst[kw * C + c] = from_f32<TO>((float)src[kw] / 255.0f - 0.5f);
      }
    } else {
      const TI* src = x + ((n * H + ih) * W + iw0) * C;
      for (int e = 0; e < k * C; ++e) dst[e] = from_f32<TO>(to_f32(src[e]));
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(256) col2im_vec_kernel(const T* __restrict__ col, long long ldcol, const float* __restrict__ bias, T* __restrict__ y,
                                                         int N, int Hi, int Wi, int C, int k, int Ho, int Wo) {
  genrl::pdl_sync();
  const int C8 = C / 8;
  const long long total = (long long)N * Ho * Wo * C8;
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
    const int c = (int)(i % C8) * 8;
    const int ow = (int)((i / C8) % Wo); const int oh = (int)((i / ((long long)C8 * Wo)) % Ho); const long long n = i / ((long long)C8 * Wo * Ho);
    float acc[8];
    if (bias) ldf8(bias + c, acc);
    else {
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    }
    for (int kh = oh & 1; kh < k && kh <= oh; kh += 2) {
      const int ih = (oh - kh) >> 1;
      if (ih >= Hi) continue;
      for (int kw = ow & 1; kw < k && kw <= ow; kw += 2) {
        const int iw = (ow - kw) >> 1;
        if (iw >= Wi) continue;
        float v[8];
        V8<T>::load(col + ((n * Hi + ih) * Wi + iw) * ldcol + (long long)(kh * k + kw) * C + c, v);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] += v[j];
      }
    }
    V8<T>::store(y + ((n * Ho + oh) * (long long)Wo + ow) * C + c, acc);
  }
}

static inline unsigned grid_for(long long total) {
  long long g = (total + 255) / 256;
  return (unsigned)(g > 148 * 32 ? 148 * 32 : (g < 1 ? 1 : g));
}

}  // namespace genrl

using namespace genrl;
#define ST(s) reinterpret_cast<cudaStream_t>(s)

// x: uint8 NCHW (x_dtype = -1) or NHWC of x_dtype; col has col_dtype.
extern "C" int genrl_im2col(void* stream, const void* x, int x_dtype, void* col, long long ldcol, int col_dtype, int N, int H, int W, int C, int k,
                            int Ho, int Wo) {
  const long long total = (long long)N * Ho * Wo * k * k * C;
  if (total <= 0) return 0;
  GENRL_REQUIRE(2 * (Ho - 1) + k <= H && 2 * (Wo - 1) + k <= W, "im2col: output grid %dx%d exceeds input %dx%d (k=%d)", Ho, Wo, H, W, k);
  cudaStream_t st = ST(stream);
  const long long segs = (long long)N * Ho * Wo * k;
  if (x_dtype == -1) {
    if (col_dtype == GENRL_F32) genrl::launch_pdl(im2col_seg_kernel<uint8_t, float, true>, grid_for(segs), 256, 0, st, (const uint8_t*)x, (float*)col, ldcol, N, H, W, C, k, Ho, Wo);
    else genrl::launch_pdl(im2col_seg_kernel<uint8_t, __nv_bfloat16, true>, grid_for(segs), 256, 0, st, (const uint8_t*)x, (__nv_bfloat16*)col, ldcol, N, H, W, C, k, Ho, Wo);
  } else {
    GENRL_REQUIRE(x_dtype == col_dtype, "im2col: dtype mismatch");
    if (col_dtype == GENRL_F32) {
      if (C % 8 == 0 && vec_ok<float>(x, C) && vec_ok<float>(col, ldcol))
        genrl::launch_pdl(im2col_vec_kernel<float>, grid_for(total / 8), 256, 0, st, (const float*)x, (float*)col, ldcol, N, H, W, C, k, Ho, Wo);
      else genrl::launch_pdl(im2col_seg_kernel<float, float, false>, grid_for(segs), 256, 0, st, (const float*)x, (float*)col, ldcol, N, H, W, C, k, Ho, Wo);
    } else {
      if (C % 8 == 0 && vec_ok<__nv_bfloat16>(x, C) && vec_ok<__nv_bfloat16>(col, ldcol))
        genrl::launch_pdl(im2col_vec_kernel<__nv_bfloat16>, grid_for(total / 8), 256, 0, st, (const __nv_bfloat16*)x, (__nv_bfloat16*)col, ldcol, N, H, W, C, k, Ho, Wo);
      else genrl::launch_pdl(im2col_seg_kernel<__nv_bfloat16, __nv_bfloat16, false>, grid_for(segs), 256, 0, st, (const __nv_bfloat16*)x, (__nv_bfloat16*)col, ldcol, N, H, W, C, k, Ho, Wo);
    }
  }
  return check_launch("im2col");
}

// y: NHWC of y_dtype, or (nchw_out) fp32 NCHW.
extern "C" int genrl_col2im(void* stream, const void* col, long long ldcol, int col_dtype, const float* bias, void* y, int y_dtype, int nchw_out,
                            int N, int Hi, int Wi, int C, int k, int Ho, int Wo) {
  const long long total = (long long)N * Ho * Wo * C;
  if (total <= 0) return 0;
  unsigned g = grid_for(total);
  cudaStream_t st = ST(stream);
  if (nchw_out) {
    GENRL_REQUIRE(y_dtype == GENRL_F32, "col2im: NCHW output is fp32");
    if (col_dtype == GENRL_F32) genrl::launch_pdl(col2im_kernel<float, float, true>, g, 256, 0, st, (const float*)col, ldcol, bias, (float*)y, N, Hi, Wi, C, k, Ho, Wo);
    else genrl::launch_pdl(col2im_kernel<__nv_bfloat16, float, true>, g, 256, 0, st, (const __nv_bfloat16*)col, ldcol, bias, (float*)y, N, Hi, Wi, C, k, Ho, Wo);
  } else {
    GENRL_REQUIRE(y_dtype == col_dtype, "col2im: dtype mismatch");
    const bool bias_ok = bias == nullptr || (reinterpret_cast<uintptr_t>(bias) & 15) == 0;
    if (col_dtype == GENRL_F32) {
      if (C % 8 == 0 && bias_ok && vec_ok<float>(col, ldcol) && vec_ok<float>(y, C))
        genrl::launch_pdl(col2im_vec_kernel<float>, grid_for(total / 8), 256, 0, st, (const float*)col, ldcol, bias, (float*)y, N, Hi, Wi, C, k, Ho, Wo);
      else genrl::launch_pdl(col2im_kernel<float, float, false>, g, 256, 0, st, (const float*)col, ldcol, bias, (float*)y, N, Hi, Wi, C, k, Ho, Wo);
    } else {
      if (C % 8 == 0 && bias_ok && vec_ok<__nv_bfloat16>(col, ldcol) && vec_ok<__nv_bfloat16>(y, C))
        genrl::launch_pdl(col2im_vec_kernel<__nv_bfloat16>, grid_for(total / 8), 256, 0, st, (const __nv_bfloat16*)col, ldcol, bias, (__nv_bfloat16*)y, N, Hi, Wi, C, k, Ho, Wo);
      else genrl::launch_pdl(col2im_kernel<__nv_bfloat16, __nv_bfloat16, false>, g, 256, 0, st, (const __nv_bfloat16*)col, ldcol, bias, (__nv_bfloat16*)y, N, Hi, Wi, C, k, Ho, Wo);
    }
  }
  return check_launch("col2im");
}

extern "C" int genrl_nchw_to_nhwc(void* stream, const float* x, void* y, int y_dtype, int N, int C, int H, int W) {
  const long long total = (long long)N * C * H * W;
  if (total <= 0) return 0;
  if (y_dtype == GENRL_F32) genrl::launch_pdl(nchw_to_nhwc_kernel<float>, grid_for(total), 256, 0, ST(stream), x, (float*)y, N, C, H, W);
  else genrl::launch_pdl(nchw_to_nhwc_kernel<__nv_bfloat16>, grid_for(total), 256, 0, ST(stream), x, (__nv_bfloat16*)y, N, C, H, W);
  return check_launch("nchw_to_nhwc");
}

// src [A,B,P] (src_dtype) -> dst [A,P,B] (dst_dtype); accumulate adds into dst (fp32 grads).
extern "C" int genrl_permute_abp(void* stream, const void* src, int src_dtype, void* dst, int dst_dtype, long long A, int B, int P, int accumulate) {
  const long long total = A * B * P;
  if (total <= 0) return 0;
  unsigned g = grid_for(total);
  cudaStream_t st = ST(stream);
  if (src_dtype == GENRL_F32 && dst_dtype == GENRL_F32) genrl::launch_pdl(permute_abp_kernel<float, float>, g, 256, 0, st, (const float*)src, (float*)dst, A, B, P, accumulate);
  else if (src_dtype == GENRL_F32 && dst_dtype == GENRL_BF16) genrl::launch_pdl(permute_abp_kernel<float, __nv_bfloat16>, g, 256, 0, st, (const float*)src, (__nv_bfloat16*)dst, A, B, P, accumulate);
  else if (src_dtype == GENRL_BF16 && dst_dtype == GENRL_F32) genrl::launch_pdl(permute_abp_kernel<__nv_bfloat16, float>, g, 256, 0, st, (const __nv_bfloat16*)src, (float*)dst, A, B, P, accumulate);
  else genrl::launch_pdl(permute_abp_kernel<__nv_bfloat16, __nv_bfloat16>, g, 256, 0, st, (const __nv_bfloat16*)src, (__nv_bfloat16*)dst, A, B, P, accumulate);
  return check_launch("permute_abp");
}
