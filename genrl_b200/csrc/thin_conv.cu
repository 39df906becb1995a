// Direct fused kernels for the two "thin" convolutions of the Dreamer conv stacks — the 3-channel ends of the encoder and
// decoder (reference Encoder._conv_model[0..2] dreamer_utils.py:580-589, Decoder._conv_model[9] :656-671 + MSEDist :74-83).
// These layers carry 2.2-2.3 M pixels x 48 channels but only K = 48 / N = 3 of contraction depth: they are HBM-bound byte
// shuffling, not GEMMs (lowering them materialises 220-470 MB im2col buffers for 5-11 GFLOP).  Each kernel keeps the 48-wide
// pixel vector in registers, the (tiny) filter in shared memory, and fuses everything that touches the big activation:
//   conv0 fwd : uint8 NCHW -> (x/255-0.5) -> Conv2d(3->C, k4, s2) + bias -> LayerNorm(C) -> SiLU -> y, stats          (1 write)
//   conv0 bwd : dy -> [recompute conv + LN] -> LN backward -> dW, db, dgamma, dbeta                                    (1 read)
//   convT fwd : x -> ConvTranspose2d(C->3, k6, s2) + bias -> mean (fp32 NCHW) and -sum((mean - obs)^2) per frame
//   convT bwd : dlike, mean, obs -> dy on the fly -> dx, dW, db
#include "common.cuh"
#include "rowops_fast.cuh"
#include <algorithm>

namespace genrl {

constexpr int TC_K0 = 4;     // encoder first-layer kernel size
constexpr int TC_KT = 6;     // decoder last-layer kernel size
constexpr int TC_PIX = 128;  // pixels per block tile

__device__ __forceinline__ float reduce_scatter32_(float (&v)[32], int lane) {
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) {
    const bool up = (lane & o) != 0;
#pragma unroll
    for (int i = 0; i < o; ++i) {
      float send = up ? v[i] : v[i + o];
      float keep = up ? v[i + o] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
    }
  }
  return v[0];
}

// ------------------------------------------------------------------------------------------------------------------------------
// encoder layer 0
// ------------------------------------------------------------------------------------------------------------------------------
template <int CH>
__device__ __forceinline__ void conv0_pixel(const uint8_t* __restrict__ obs, const float* __restrict__ ws, const float* __restrict__ bias,
                                            long long n, int oh, int ow, int H, int W, float (&patch)[48], float (&acc)[CH]) {
  // patch index k = (c*4 + kh)*4 + kw  (the reference filter layout [co][c][kh][kw] flattened)
#pragma unroll
  for (int c = 0; c < 3; ++c)
#pragma unroll
    for (int kh = 0; kh < TC_K0; ++kh) {
      const uint8_t* src = obs + ((n * 3 + c) * H + 2 * oh + kh) * W + 2 * ow;
      const uint16_t lo = *reinterpret_cast<const uint16_t*>(src), hi = *reinterpret_cast<const uint16_t*>(src + 2);
      const int k = (c * 4 + kh) * 4;
      patch[k + 0] = (float)(lo & 0xff) / 255.0f - 0.5f; patch[k + 1] = (float)(lo >> 8) / 255.0f - 0.5f;
      patch[k + 2] = (float)(hi & 0xff) / 255.0f - 0.5f; patch[k + 3] = (float)(hi >> 8) / 255.0f - 0.5f;
    }
#pragma unroll
  for (int co = 0; co < CH; ++co) acc[co] = bias[co];
#pragma unroll
  for (int k = 0; k < 48; ++k) {
    const float p = patch[k];
#pragma unroll
    for (int c4 = 0; c4 < CH / 4; ++c4) {
      const float4 w = *reinterpret_cast<const float4*>(ws + k * CH + c4 * 4);       // broadcast read
      acc[c4 * 4 + 0] = fmaf(p, w.x, acc[c4 * 4 + 0]); acc[c4 * 4 + 1] = fmaf(p, w.y, acc[c4 * 4 + 1]);
      acc[c4 * 4 + 2] = fmaf(p, w.z, acc[c4 * 4 + 2]); acc[c4 * 4 + 3] = fmaf(p, w.w, acc[c4 * 4 + 3]);
    }
  }
}

template <int CH>
__device__ __forceinline__ void load_w0(const float* __restrict__ w, float* ws) {
  // ws[k][co] = w[co][k]   (w is [CH][48] in the reference layout)
  for (int i = threadIdx.x; i < 48 * CH; i += blockDim.x) { int co = i % CH, k = i / CH; ws[k * CH + co] = w[co * 48 + k]; }
}

template <typename T, int CH>
__global__ void __launch_bounds__(TC_PIX) conv0_fwd_kernel(const uint8_t* __restrict__ obs, const float* __restrict__ w, const float* __restrict__ bias,
                                                           const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                                                           T* __restrict__ y, float* __restrict__ stats, long long M, int H, int W, int Ho, int Wo) {
  genrl::pdl_sync();
  __shared__ __align__(16) float ws[48 * CH];
  load_w0<CH>(w, ws);                   // once per (persistent) block
  __syncthreads();
  // one tile per block: a persistent loop lets the compiler hoist the 144 loop-invariant bias/gamma/beta loads into registers (spills)
  const long long m = (long long)blockIdx.x * TC_PIX + threadIdx.x;
  if (m >= M) return;
  {
  const int ow = (int)(m % Wo), oh = (int)((m / Wo) % Ho); const long long n = m / ((long long)Wo * Ho);
  float patch[48], acc[CH];
  conv0_pixel<CH>(obs, ws, bias, n, oh, ow, H, W, patch, acc);
  float s = 0.f;
#pragma unroll
  for (int co = 0; co < CH; ++co) s += acc[co];
  const float mean = s / CH;
  float q = 0.f;
#pragma unroll
  for (int co = 0; co < CH; ++co) { float d = acc[co] - mean; q += d * d; }
  const float rstd = 1.0f / sqrtf(q / CH + eps);
#pragma unroll
  for (int c8 = 0; c8 < CH / 8; ++c8) {
    float o[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { const int co = c8 * 8 + j; o[j] = silu_t<(sizeof(T) == 2)>((acc[co] - mean) * rstd * gamma[co] + beta[co]); }
    V8<T>::store(y + m * CH + c8 * 8, o);
  }
  stats[m * 2] = mean; stats[m * 2 + 1] = rstd;
  }
}

// persistent blocks; gradients accumulated in registers across pixel tiles, one atomic flush per block
template <typename T, int CH>
__global__ void __launch_bounds__(TC_PIX) conv0_bwd_kernel(const uint8_t* __restrict__ obs, const float* __restrict__ w, const float* __restrict__ bias,
                                                           const float* __restrict__ gamma, const float* __restrict__ beta,
                                                           const float* __restrict__ stats, const T* __restrict__ dy, float* __restrict__ dW,
                                                           float* __restrict__ db, float* __restrict__ dgamma, float* __restrict__ dbeta,
                                                           long long M, int H, int W, int Ho, int Wo) {
  genrl::pdl_sync();
  extern __shared__ __align__(16) float smem0[];
  float* ws = smem0;                     // [48][CH]
  constexpr int PS = 49, DS = CH + 1;    // odd row strides: conflict-free per-thread row writes
  float* Ps = ws + 48 * CH;              // [TC_PIX][PS]   patches
  float* Ds = Ps + TC_PIX * PS;          // [TC_PIX][DS]   dpre
  load_w0<CH>(w, ws);
  const int tid = threadIdx.x, lane = tid & 31;
  constexpr int KG = 3, CG = CH / 8;     // thread owns 3 k x CG co of dW: 16 k-groups x 8 co-groups = 128 threads
  const int tk = tid / 8, tc = tid % 8;
  float aW[KG][CG];
#pragma unroll
  for (int i = 0; i < KG; ++i)
#pragma unroll
    for (int j = 0; j < CG; ++j) aW[i][j] = 0.f;
  float a_gb[(2 * CH + 31) / 32];        // after the transposing reduction lane l owns indices 32 s + l of [dgamma | dbeta]
#pragma unroll
  for (int i = 0; i < (2 * CH + 31) / 32; ++i) a_gb[i] = 0.f;
  float a_db = 0.f;                      // threads < CH: column sum of dpre
  __syncthreads();
  for (long long base = (long long)blockIdx.x * TC_PIX; base < M; base += (long long)gridDim.x * TC_PIX) {
    const long long m = base + tid;
    const bool valid = m < M;
    float xh[CH], g[CH];
    if (valid) {
      const int ow = (int)(m % Wo), oh = (int)((m / Wo) % Ho); const long long n = m / ((long long)Wo * Ho);
      float patch[48];
      conv0_pixel<CH>(obs, ws, bias, n, oh, ow, H, W, patch, xh);
#pragma unroll
      for (int k = 0; k < 48; ++k) Ps[tid * PS + k] = patch[k];
      const float mean = stats[m * 2], rstd = stats[m * 2 + 1];
      float s1 = 0.f, s2 = 0.f;
#pragma unroll
      for (int c8 = 0; c8 < CH / 8; ++c8) {
        float d[8];
        V8<T>::load(dy + m * CH + c8 * 8, d);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int co = c8 * 8 + j;
          const float h = (xh[co] - mean) * rstd;
          const float gg = d[j] * silu_grad_t<(sizeof(T) == 2)>(h * gamma[co] + beta[co]);
          xh[co] = h; g[co] = gg;
          s1 += gg * gamma[co]; s2 += gg * gamma[co] * h;
        }
      }
      s1 /= CH; s2 /= CH;
#pragma unroll
      for (int co = 0; co < CH; ++co) Ds[tid * DS + co] = rstd * (g[co] * gamma[co] - s1 - xh[co] * s2);
    } else {
#pragma unroll
      for (int co = 0; co < CH; ++co) { xh[co] = 0.f; g[co] = 0.f; Ds[tid * DS + co] = 0.f; }
#pragma unroll
      for (int k = 0; k < 48; ++k) Ps[tid * PS + k] = 0.f;
    }
    // dgamma / dbeta: warp-transposing reduction of [g*xhat (CH) | g (CH)]
#pragma unroll
    for (int sgrp = 0; sgrp < (2 * CH + 31) / 32; ++sgrp) {
      float v[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const int idx = sgrp * 32 + i;
        const int a = idx < CH ? idx : (idx < 2 * CH ? idx - CH : 0);
        v[i] = idx < CH ? g[a] * xh[a] : (idx < 2 * CH ? g[a] : 0.f);
      }
      a_gb[sgrp] += reduce_scatter32_(v, lane);
    }
    __syncthreads();
    // dW[co][k] += sum_p dpre[p][co] * patch[p][k]
    for (int p = 0; p < TC_PIX; ++p) {
      float pk[KG], dc[CG];
#pragma unroll
      for (int i = 0; i < KG; ++i) pk[i] = Ps[p * PS + tk * KG + i];
#pragma unroll
      for (int j = 0; j < CG; ++j) dc[j] = Ds[p * DS + tc * CG + j];
#pragma unroll
      for (int i = 0; i < KG; ++i)
#pragma unroll
        for (int j = 0; j < CG; ++j) aW[i][j] = fmaf(pk[i], dc[j], aW[i][j]);
    }
    if (tid < CH) { float sdb = 0.f; for (int p = 0; p < TC_PIX; ++p) sdb += Ds[p * DS + tid]; a_db += sdb; }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < KG; ++i)
#pragma unroll
    for (int j = 0; j < CG; ++j) atomicAdd(&dW[(tc * CG + j) * 48 + tk * KG + i], aW[i][j]);
  if (tid < CH) atomicAdd(&db[tid], a_db);
  // every warp holds partial [dgamma|dbeta] sums for its own pixels
#pragma unroll
  for (int sgrp = 0; sgrp < (2 * CH + 31) / 32; ++sgrp) {
    const int idx = sgrp * 32 + lane;
    if (idx < CH) atomicAdd(&dgamma[idx], a_gb[sgrp]);
    else if (idx < 2 * CH) atomicAdd(&dbeta[idx - CH], a_gb[sgrp]);
  }
}

// ------------------------------------------------------------------------------------------------------------------------------
// decoder last layer: ConvTranspose2d(CH -> 3, k6, s2) + MSE against the uint8 observation
// ------------------------------------------------------------------------------------------------------------------------------
// Work item = 4 output pixels of one parity class (oh = 2*ohh + ph, ow = pw + 2*(4*q + j), j = 0..3): they use the same filter
// taps (kh = ph + 2a, kw = pw + 2b) on 4 consecutive input pixels, so every shared-memory filter read feeds 12 FMAs.  Items are
// ordered (n, ph, pw, ohh, q): 256 consecutive items share the parity class, hence each warp reads one filter address at a time.
template <typename T, int CH>
__global__ void __launch_bounds__(256) convT3_fwd_kernel(const T* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                                                         const uint8_t* __restrict__ obs, float* __restrict__ mean_out, float* __restrict__ like,
                                                         long long items, int Hi, int Wi, int Ho, int Wo) {
  genrl::pdl_sync();
  __shared__ float4 ws[TC_KT * TC_KT * CH];          // [kh][kw][ci] -> (co0, co1, co2, 0)
  for (int i = threadIdx.x; i < TC_KT * TC_KT * CH; i += 256) {
    const int ci = i % CH, kk = i / CH, kh = kk / TC_KT, kw = kk % TC_KT;
    ws[i] = make_float4(w[((ci * 3 + 0) * TC_KT + kh) * TC_KT + kw], w[((ci * 3 + 1) * TC_KT + kh) * TC_KT + kw],
                        w[((ci * 3 + 2) * TC_KT + kh) * TC_KT + kw], 0.f);
  }
  __syncthreads();
  const int Hh = Ho / 2, Q = Wo / 8;                  // rows per parity, 4-pixel groups per row and parity
  const long long plane = (long long)Ho * Wo;
  const float b0 = bias[0], b1 = bias[1], b2 = bias[2];
#pragma unroll 1
  for (long long it = (long long)blockIdx.x * 256 + threadIdx.x; it < ((items + 31) / 32) * 32; it += (long long)gridDim.x * 256) {
    float sse = 0.f;
    long long n = 0;
    if (it < items) {
      const int q = (int)(it % Q); const int ohh = (int)((it / Q) % Hh); const int pw = (int)((it / ((long long)Q * Hh)) & 1);
      const int ph = (int)((it / ((long long)Q * Hh * 2)) & 1); n = it / ((long long)Q * Hh * 4);
      const int oh = 2 * ohh + ph, ow0 = pw + 8 * q;
      float acc[4][3];
#pragma unroll
      for (int j = 0; j < 4; ++j) { acc[j][0] = b0; acc[j][1] = b1; acc[j][2] = b2; }
#pragma unroll 1
      for (int a = 0; a < TC_KT / 2; ++a) {
        const int kh = ph + 2 * a, ih = ohh - a;
        if (ih < 0 || ih >= Hi) continue;
#pragma unroll 1
        for (int b = 0; b < TC_KT / 2; ++b) {
          const int kw = pw + 2 * b;
          const float4* wk = ws + (kh * TC_KT + kw) * CH;
          const int iw0 = 4 * q - b;                  // input column of pixel j is iw0 + j
          const T* xr = x + ((n * Hi + ih) * (long long)Wi) * CH;
#pragma unroll
          for (int c8 = 0; c8 < CH / 8; ++c8) {
            float v[4][8];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int iw = iw0 + j;
              if (iw >= 0 && iw < Wi) V8<T>::load(xr + (long long)iw * CH + c8 * 8, v[j]);
              else {
#pragma unroll
                for (int e = 0; e < 8; ++e) v[j][e] = 0.f;
              }
            }
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const float4 f = wk[c8 * 8 + e];
#pragma unroll
              for (int j = 0; j < 4; ++j) { acc[j][0] = fmaf(v[j][e], f.x, acc[j][0]); acc[j][1] = fmaf(v[j][e], f.y, acc[j][1]); acc[j][2] = fmaf(v[j][e], f.z, acc[j][2]); }
            }
          }
        }
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const long long pix = (long long)oh * Wo + ow0 + 2 * j;
#pragma unroll
        for (int co = 0; co < 3; ++co) {
          const float o = (float)obs[(n * 3 + co) * plane + pix] / 255.0f - 0.5f;
          if (mean_out) mean_out[(n * 3 + co) * plane + pix] = acc[j][co];
          const float d = acc[j][co] - o;
          sse += d * d;
        }
      }
    }
    sse = warp_sum(sse);                              // a warp's items belong to one frame (Hh*Q*4 % 32 == 0, checked by the host)
    if ((threadIdx.x & 31) == 0 && it < items) atomicAdd(&like[n], -sse);
  }
}

template <typename T, int CH>
__global__ void __launch_bounds__(TC_PIX) convT3_bwd_kernel(const T* __restrict__ x, const float* __restrict__ w, const float* __restrict__ mean,
                                                            const uint8_t* __restrict__ obs, const float* __restrict__ dlike, T* __restrict__ dx,
                                                            float* __restrict__ dW, float* __restrict__ db, long long Mi, int Hi, int Wi, int Ho, int Wo) {
  genrl::pdl_sync();
  constexpr int NJ = 3 * TC_KT * TC_KT;                // 108 = (co, kh, kw)
  constexpr int JG = 14, CG = 3;                       // thread owns 3 ci x 14 j of dW: 16 ci-groups x 8 j-groups
  extern __shared__ __align__(16) float smemT[];
  float* wsb = smemT;                                  // [NJ][CH]       wsb[j][ci] = w[ci][j]
  constexpr int XS = CH + 1, YS = NJ + 1;              // odd row strides (bank-conflict-free per-thread rows)
  float* Xs = wsb + NJ * CH;                           // [TC_PIX][XS]
  float* Ys = Xs + TC_PIX * XS;                        // [TC_PIX][YS]  dy windows
  for (int i = threadIdx.x; i < NJ * CH; i += TC_PIX) { const int ci = i % CH, j = i / CH; wsb[j * CH + ci] = w[ci * NJ + j]; }
  const int tid = threadIdx.x;
  const int tci = tid / 8, tj = tid % 8;
  float aW[CG][JG];
#pragma unroll
  for (int a = 0; a < CG; ++a)
#pragma unroll
    for (int b = 0; b < JG; ++b) aW[a][b] = 0.f;
  float a_db0 = 0.f, a_db1 = 0.f, a_db2 = 0.f;
  const long long plane = (long long)Ho * Wo;
  __syncthreads();
  for (long long base = (long long)blockIdx.x * TC_PIX; base < Mi; base += (long long)gridDim.x * TC_PIX) {
    const long long m = base + tid;
    const bool valid = m < Mi;
    if (valid) {
      const int iw = (int)(m % Wi), ih = (int)((m / Wi) % Hi); const long long n = m / ((long long)Wi * Hi);
      const float gl = -2.0f * dlike[n];
      // dy window (co, kh, kw) around output pixel (2 ih + kh, 2 iw + kw)
#pragma unroll 1
      for (int co = 0; co < 3; ++co)
#pragma unroll 1
        for (int kh = 0; kh < TC_KT; ++kh) {
          const long long off = (n * 3 + co) * plane + (long long)(2 * ih + kh) * Wo + 2 * iw;
#pragma unroll
          for (int kw = 0; kw < TC_KT; ++kw) {
            const float o = (float)obs[off + kw] / 255.0f - 0.5f;
            const float d = gl * (mean[off + kw] - o);
            Ys[tid * YS + (co * TC_KT + kh) * TC_KT + kw] = d;
            const bool own = (kh < 2 || ih == Hi - 1) && (kw < 2 || iw == Wi - 1);     // each output pixel counted once for db
            if (own) { if (co == 0) a_db0 += d; else if (co == 1) a_db1 += d; else a_db2 += d; }
          }
        }
      float acc[CH];
#pragma unroll
      for (int ci = 0; ci < CH; ++ci) acc[ci] = 0.f;
#pragma unroll 2
      for (int j = 0; j < NJ; ++j) {
        const float d = Ys[tid * YS + j];
#pragma unroll
        for (int c4 = 0; c4 < CH / 4; ++c4) {
          const float4 q = *reinterpret_cast<const float4*>(wsb + j * CH + c4 * 4);
          acc[c4 * 4 + 0] = fmaf(d, q.x, acc[c4 * 4 + 0]); acc[c4 * 4 + 1] = fmaf(d, q.y, acc[c4 * 4 + 1]);
          acc[c4 * 4 + 2] = fmaf(d, q.z, acc[c4 * 4 + 2]); acc[c4 * 4 + 3] = fmaf(d, q.w, acc[c4 * 4 + 3]);
        }
      }
#pragma unroll
      for (int c8 = 0; c8 < CH / 8; ++c8) {
        float o[8], xv[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = acc[c8 * 8 + j];
        V8<T>::store(dx + m * CH + c8 * 8, o);
        V8<T>::load(x + m * CH + c8 * 8, xv);
#pragma unroll
        for (int j = 0; j < 8; ++j) Xs[tid * XS + c8 * 8 + j] = xv[j];
      }
    } else {
      for (int j = 0; j < NJ; ++j) Ys[tid * YS + j] = 0.f;
      for (int ci = 0; ci < CH; ++ci) Xs[tid * XS + ci] = 0.f;
    }
    __syncthreads();
    // dW[ci][j] += sum_p x[p][ci] * dy[p][j]
    for (int p = 0; p < TC_PIX; ++p) {
      float xv[CG], dv[JG];
#pragma unroll
      for (int a = 0; a < CG; ++a) xv[a] = Xs[p * XS + tci * CG + a];
#pragma unroll
      for (int b = 0; b < JG; ++b) { const int j = tj * JG + b; dv[b] = j < NJ ? Ys[p * YS + j] : 0.f; }
#pragma unroll
      for (int a = 0; a < CG; ++a)
#pragma unroll
        for (int b = 0; b < JG; ++b) aW[a][b] = fmaf(xv[a], dv[b], aW[a][b]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int a = 0; a < CG; ++a)
#pragma unroll
    for (int b = 0; b < JG; ++b) { const int j = tj * JG + b; if (j < NJ) atomicAdd(&dW[(tci * CG + a) * NJ + j], aW[a][b]); }
  a_db0 = warp_sum(a_db0); a_db1 = warp_sum(a_db1); a_db2 = warp_sum(a_db2);
  if ((tid & 31) == 0) { atomicAdd(&db[0], a_db0); atomicAdd(&db[1], a_db1); atomicAdd(&db[2], a_db2); }
}

}  // namespace genrl

using namespace genrl;
#define ST(s) reinterpret_cast<cudaStream_t>(s)

static int sm_count() {
  static int n = 0;
  if (!n) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev); if (n <= 0) n = 148; }
  return n;
}

// 1 if the fused thin-conv kernels cover this channel count (the GenRL default cnn_depth = 48)
extern "C" int genrl_thin_conv_supported(int channels) { return channels == 48 ? 1 : 0; }

// Encoder layer 0: y[N*Ho*Wo, C] = SiLU(LN_C(Conv2d(obs/255-0.5; w[C,3,4,4], b))) ; stats[., 2] = (mean, rstd)
extern "C" int genrl_conv0_ln_silu_fwd(void* stream, const unsigned char* obs, const float* w, const float* bias, const float* gamma,
                                       const float* beta, float eps, void* y, float* stats, int N, int H, int W, int C, int dtype) {
  GENRL_REQUIRE(C == 48 && (W % 2) == 0, "conv0_ln_silu: unsupported shape C=%d W=%d", C, W);
  const int Ho = (H - TC_K0) / 2 + 1, Wo = (W - TC_K0) / 2 + 1;
  const long long M = (long long)N * Ho * Wo;
  if (M <= 0) return 0;
  const unsigned grid = (unsigned)((M + TC_PIX - 1) / TC_PIX);
  if (dtype == GENRL_BF16) genrl::launch_pdl(conv0_fwd_kernel<__nv_bfloat16, 48>, grid, TC_PIX, 0, ST(stream), obs, w, bias, gamma, beta, eps, (__nv_bfloat16*)y, stats, M, H, W, Ho, Wo);
  else genrl::launch_pdl(conv0_fwd_kernel<float, 48>, grid, TC_PIX, 0, ST(stream), obs, w, bias, gamma, beta, eps, (float*)y, stats, M, H, W, Ho, Wo);
  return check_launch("conv0_ln_silu_fwd");
}

// gradients are ACCUMULATED into dW [C,3,4,4], db, dgamma, dbeta (fp32)
extern "C" int genrl_conv0_ln_silu_bwd(void* stream, const unsigned char* obs, const float* w, const float* bias, const float* gamma,
                                       const float* beta, const float* stats, const void* dy, float* dW, float* db, float* dgamma, float* dbeta,
                                       int N, int H, int W, int C, int dtype) {
  GENRL_REQUIRE(C == 48 && (W % 2) == 0, "conv0_ln_silu: unsupported shape C=%d W=%d", C, W);
  const int Ho = (H - TC_K0) / 2 + 1, Wo = (W - TC_K0) / 2 + 1;
  const long long M = (long long)N * Ho * Wo;
  if (M <= 0) return 0;
  const size_t smem = (size_t)(48 * 48 + TC_PIX * 49 + TC_PIX * 49) * sizeof(float);
  const unsigned grid = (unsigned)std::min<long long>((M + TC_PIX - 1) / TC_PIX, (long long)sm_count() * 3);
  if (dtype == GENRL_BF16) {
    cudaFuncSetAttribute(conv0_bwd_kernel<__nv_bfloat16, 48>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    genrl::launch_pdl(conv0_bwd_kernel<__nv_bfloat16, 48>, grid, TC_PIX, smem, ST(stream), obs, w, bias, gamma, beta, stats, (const __nv_bfloat16*)dy, dW, db, dgamma, dbeta, M, H, W, Ho, Wo);
  } else {
    cudaFuncSetAttribute(conv0_bwd_kernel<float, 48>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    genrl::launch_pdl(conv0_bwd_kernel<float, 48>, grid, TC_PIX, smem, ST(stream), obs, w, bias, gamma, beta, stats, (const float*)dy, dW, db, dgamma, dbeta, M, H, W, Ho, Wo);
  }
  return check_launch("conv0_ln_silu_bwd");
}

// Decoder last layer: mean[N,3,Ho,Wo] (optional) = ConvT(x[N*Hi*Wi, C]; w[C,3,6,6], b);  like[n] = -sum (mean - (obs/255-0.5))^2
extern "C" int genrl_convT3_mse_fwd(void* stream, const void* x, const float* w, const float* bias, const unsigned char* obs, float* mean_out,
                                    float* like, int N, int Hi, int Wi, int C, int dtype) {
  const int Ho = (Hi - 1) * 2 + TC_KT, Wo = (Wi - 1) * 2 + TC_KT;
  GENRL_REQUIRE(C == 48 && (Wo % 8) == 0 && (Ho % 2) == 0 && (((Ho / 2) * (Wo / 8) * 4) % 32) == 0, "convT3_mse: unsupported shape C=%d Ho=%d Wo=%d", C, Ho, Wo);
  const long long total = (long long)N * Ho * (Wo / 4);          // work items of 4 pixels
  if (total <= 0) return 0;
  cudaMemsetAsync(like, 0, (size_t)N * sizeof(float), ST(stream));
  const unsigned grid = (unsigned)std::min<long long>((total + 255) / 256, (long long)sm_count() * 8);
  if (dtype == GENRL_BF16) genrl::launch_pdl(convT3_fwd_kernel<__nv_bfloat16, 48>, grid, 256, 0, ST(stream), (const __nv_bfloat16*)x, w, bias, obs, mean_out, like, total, Hi, Wi, Ho, Wo);
  else genrl::launch_pdl(convT3_fwd_kernel<float, 48>, grid, 256, 0, ST(stream), (const float*)x, w, bias, obs, mean_out, like, total, Hi, Wi, Ho, Wo);
  return check_launch("convT3_mse_fwd");
}

// dx[N*Hi*Wi, C] written; dW [C,3,6,6] and db [3] ACCUMULATED.  dlike[n] = d loss / d like[n].
extern "C" int genrl_convT3_mse_bwd(void* stream, const void* x, const float* w, const float* mean, const unsigned char* obs, const float* dlike,
                                    void* dx, float* dW, float* db, int N, int Hi, int Wi, int C, int dtype) {
  const int Ho = (Hi - 1) * 2 + TC_KT, Wo = (Wi - 1) * 2 + TC_KT;
  GENRL_REQUIRE(C == 48, "convT3_mse: unsupported C=%d", C);
  const long long Mi = (long long)N * Hi * Wi;
  if (Mi <= 0) return 0;
  const size_t smem = (size_t)(108 * 48 + TC_PIX * 49 + TC_PIX * 109) * sizeof(float);
  const unsigned grid = (unsigned)std::min<long long>((Mi + TC_PIX - 1) / TC_PIX, (long long)sm_count() * 2);
  if (dtype == GENRL_BF16) {
    cudaFuncSetAttribute(convT3_bwd_kernel<__nv_bfloat16, 48>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    genrl::launch_pdl(convT3_bwd_kernel<__nv_bfloat16, 48>, grid, TC_PIX, smem, ST(stream), (const __nv_bfloat16*)x, w, mean, obs, dlike, (__nv_bfloat16*)dx, dW, db, Mi, Hi, Wi, Ho, Wo);
  } else {
    cudaFuncSetAttribute(convT3_bwd_kernel<float, 48>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    genrl::launch_pdl(convT3_bwd_kernel<float, 48>, grid, TC_PIX, smem, ST(stream), (const float*)x, w, mean, obs, dlike, (float*)dx, dW, db, Mi, Hi, Wi, Ho, Wo);
  }
  return check_launch("convT3_mse_bwd");
}
