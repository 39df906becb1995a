// Row-wise fused kernels of the GenRL hot path (sm_100a): LayerNorm+SiLU, GRU gates, categorical
// straight-through sample, balanced KL, two-hot head, Normal actor head, lambda-return, MSE image loss.
// All are HBM-bound: one warp per row, coalesced loads, fp32 math, activations stored as T (fp32 | bf16).
// Each forward has a hand-written backward (no autograd inside the library).
#include "common.cuh"
#include "rowops_fast.cuh"
#include "ln_pipe.cuh"
#include <algorithm>
#include <cstdlib>

namespace genrl {

#define DISPATCH_T(dtype, ...)                                     \
  do {                                                             \
    if ((dtype) == GENRL_F32) { using T = float; __VA_ARGS__; }    \
    else if ((dtype) == GENRL_BF16) { using T = __nv_bfloat16; __VA_ARGS__; } \
    else { set_error("bad dtype %d", (int)(dtype)); return 1; }    \
  } while (0)

constexpr int WARPS = 8;
constexpr int THREADS = WARPS * 32;

static inline int row_grid(long long rows, int cap = 148 * 8) {
  long long g = (rows + WARPS - 1) / WARPS;
  if (g < 1) g = 1;
  return (int)(g > cap ? cap : g);
}

// =================================================================================================
// LayerNorm (+SiLU)     reference: NormLayer dreamer_utils.py:844-859 (eps 1e-5), ImgChLayerNorm :1031-1040 (eps 1e-3)
// =================================================================================================
template <typename T>
__global__ void __launch_bounds__(THREADS) ln_act_fwd_kernel(const T* __restrict__ x, long long ldx, const float* __restrict__ gamma,
                                                             const float* __restrict__ beta, float eps, int act, T* __restrict__ y,
                                                             long long ldy, float* __restrict__ stats, long long M, int D) {
  genrl::pdl_sync();
  const int lane = threadIdx.x & 31;
  for (long long row = (long long)blockIdx.x * WARPS + (threadIdx.x >> 5); row < M; row += (long long)gridDim.x * WARPS) {
    const T* xr = x + row * ldx;
    float s = 0.f;
    for (int c = lane; c < D; c += 32) s += to_f32(xr[c]);
    const float mean = warp_sum(s) / D;
    float v = 0.f;
    for (int c = lane; c < D; c += 32) { float d = to_f32(xr[c]) - mean; v += d * d; }
    const float rstd = 1.0f / sqrtf(warp_sum(v) / D + eps);
    T* yr = y + row * ldy;
    for (int c = lane; c < D; c += 32) {
      float a = (to_f32(xr[c]) - mean) * rstd * gamma[c] + beta[c];
      yr[c] = from_f32<T>(act == GENRL_ACT_SILU ? siluf_(a) : a);
    }
    if (stats && lane == 0) { stats[row * 2] = mean; stats[row * 2 + 1] = rstd; }
  }
}

// dx = rstd * (g*gamma - mean(g*gamma) - xhat * mean(g*gamma*xhat)),  g = dy * act'(gamma*xhat+beta)
// dgamma += sum_rows g*xhat ; dbeta += sum_rows g ; dlin_bias += sum_rows dx   (all optional, fp32 atomics)
template <typename T>
__global__ void __launch_bounds__(THREADS) ln_act_bwd_kernel(const T* __restrict__ dy, long long lddy, const T* __restrict__ x, long long ldx,
                                                             const float* __restrict__ gamma, const float* __restrict__ beta,
                                                             const float* __restrict__ stats, int act, T* __restrict__ dx, long long lddx,
                                                             float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ dlin_bias,
                                                             long long M, int D) {
  genrl::pdl_sync();
  extern __shared__ float sm[];        // [3][D] column partials
  float* s_dg = sm; float* s_db = sm + D; float* s_dl = sm + 2 * D;
  const bool need_cols = dgamma != nullptr || dlin_bias != nullptr;
  if (need_cols) { for (int i = threadIdx.x; i < 3 * D; i += THREADS) sm[i] = 0.f; __syncthreads(); }
  const int lane = threadIdx.x & 31;
  for (long long row = (long long)blockIdx.x * WARPS + (threadIdx.x >> 5); row < M; row += (long long)gridDim.x * WARPS) {
    const T* xr = x + row * ldx; const T* dyr = dy + row * lddy;
    const float mean = stats[row * 2], rstd = stats[row * 2 + 1];
    float s1 = 0.f, s2 = 0.f;
    for (int c = lane; c < D; c += 32) {
      float xh = (to_f32(xr[c]) - mean) * rstd;
      float g = to_f32(dyr[c]);
      if (act == GENRL_ACT_SILU) g *= silu_gradf_(xh * gamma[c] + beta[c]);
      float gh = g * gamma[c];
      s1 += gh; s2 += gh * xh;
    }
    s1 = warp_sum(s1) / D; s2 = warp_sum(s2) / D;
    T* dxr = dx + row * lddx;
    for (int c = lane; c < D; c += 32) {
      float xh = (to_f32(xr[c]) - mean) * rstd;
      float g = to_f32(dyr[c]);
      if (act == GENRL_ACT_SILU) g *= silu_gradf_(xh * gamma[c] + beta[c]);
      float d = rstd * (g * gamma[c] - s1 - xh * s2);
      dxr[c] = from_f32<T>(d);
      if (need_cols) {
        if (dgamma) { atomicAdd(&s_dg[c], g * xh); atomicAdd(&s_db[c], g); }
        if (dlin_bias) atomicAdd(&s_dl[c], d);
      }
    }
  }
  if (need_cols) {
    __syncthreads();
    for (int c = threadIdx.x; c < D; c += THREADS) {
      if (dgamma) { atomicAdd(&dgamma[c], s_dg[c]); atomicAdd(&dbeta[c], s_db[c]); }
      if (dlin_bias) atomicAdd(&dlin_bias[c], s_dl[c]);
    }
  }
}

// column sum: out[c] += sum_rows x[r,c]   (bias gradients of Linear layers not followed by LN)
template <typename T>
__global__ void __launch_bounds__(256) colsum_kernel(const T* __restrict__ x, long long ldx, float* __restrict__ out, long long M, int D, long long rows_per_block) {
  genrl::pdl_sync();
  const long long r0 = (long long)blockIdx.y * rows_per_block;
  const long long r1 = min(M, r0 + rows_per_block);
  for (int c = blockIdx.x * 256 + threadIdx.x; c < D; c += gridDim.x * 256) {
    float s = 0.f;
    for (long long r = r0; r < r1; ++r) s += to_f32(x[r * ldx + c]);
    atomicAdd(&out[c], s);
  }
}

// =================================================================================================
// GRU cell tail     reference: GRUCell.forward dreamer_utils.py:771-785 (LayerNorm(3D) eps 1e-5, update_bias -1, tanh)
//   z[M,3D] = W [x|h]   ->   p = LN(z);  r = sig(p_r); cand = tanh(r * p_c); u = sig(p_u - 1);  h' = u*cand + (1-u)*h
// =================================================================================================
template <typename T>
__global__ void __launch_bounds__(THREADS) gru_fwd_kernel(const T* __restrict__ z, long long ldz, const float* __restrict__ gamma,
                                                          const float* __restrict__ beta, const float* __restrict__ hprev, long long ldh,
                                                          float* __restrict__ hout, long long ldho, T* __restrict__ hout_t, long long ldht,
                                                          float* __restrict__ stats, long long M, int D) {
  genrl::pdl_sync();
  const int lane = threadIdx.x & 31;
  const int D3 = 3 * D;
  for (long long row = (long long)blockIdx.x * WARPS + (threadIdx.x >> 5); row < M; row += (long long)gridDim.x * WARPS) {
    const T* zr = z + row * ldz;
    float s = 0.f;
    for (int c = lane; c < D3; c += 32) s += to_f32(zr[c]);
    const float mean = warp_sum(s) / D3;
    float v = 0.f;
    for (int c = lane; c < D3; c += 32) { float d = to_f32(zr[c]) - mean; v += d * d; }
    const float rstd = 1.0f / sqrtf(warp_sum(v) / D3 + 1e-5f);
    for (int j = lane; j < D; j += 32) {
      float pr = (to_f32(zr[j]) - mean) * rstd * gamma[j] + beta[j];
      float pc = (to_f32(zr[D + j]) - mean) * rstd * gamma[D + j] + beta[D + j];
      float pu = (to_f32(zr[2 * D + j]) - mean) * rstd * gamma[2 * D + j] + beta[2 * D + j];
      float r = sigmoidf_(pr), cand = tanhf(r * pc), u = sigmoidf_(pu - 1.0f);
      float h = hprev[row * ldh + j];
      float hn = u * cand + (1.0f - u) * h;
      hout[row * ldho + j] = hn;
      if (hout_t) hout_t[row * ldht + j] = from_f32<T>(hn);
    }
    if (lane == 0) { stats[row * 2] = mean; stats[row * 2 + 1] = rstd; }
  }
}

template <typename T>
__global__ void __launch_bounds__(THREADS) gru_bwd_kernel(const float* __restrict__ dh, long long lddh, const T* __restrict__ z, long long ldz,
                                                          const float* __restrict__ gamma, const float* __restrict__ beta,
                                                          const float* __restrict__ hprev, long long ldh, const float* __restrict__ stats,
                                                          T* __restrict__ dz, long long lddz, float* __restrict__ dhprev, long long lddhp,
                                                          float* __restrict__ dgamma, float* __restrict__ dbeta, long long M, int D) {
  genrl::pdl_sync();
  extern __shared__ float sm[];        // [2][3D]
  const int D3 = 3 * D;
  float* s_dg = sm; float* s_db = sm + D3;
  if (dgamma) { for (int i = threadIdx.x; i < 2 * D3; i += THREADS) sm[i] = 0.f; __syncthreads(); }
  const int lane = threadIdx.x & 31;
  for (long long row = (long long)blockIdx.x * WARPS + (threadIdx.x >> 5); row < M; row += (long long)gridDim.x * WARPS) {
    const T* zr = z + row * ldz;
    const float mean = stats[row * 2], rstd = stats[row * 2 + 1];
    float s1 = 0.f, s2 = 0.f;
    // pass 1: gate gradients -> reductions for the LN backward
    for (int j = lane; j < D; j += 32) {
      float xr_ = (to_f32(zr[j]) - mean) * rstd, xc = (to_f32(zr[D + j]) - mean) * rstd, xu = (to_f32(zr[2 * D + j]) - mean) * rstd;
      float pr = xr_ * gamma[j] + beta[j], pc = xc * gamma[D + j] + beta[D + j], pu = xu * gamma[2 * D + j] + beta[2 * D + j];
      float r = sigmoidf_(pr), cand = tanhf(r * pc), u = sigmoidf_(pu - 1.0f);
      float h = hprev[row * ldh + j];
      float g = dh[row * lddh + j];
      float drc = g * u * (1.0f - cand * cand);
      float gpr = drc * pc * r * (1.0f - r), gpc = drc * r, gpu = g * (cand - h) * u * (1.0f - u);
      float a = gpr * gamma[j], b = gpc * gamma[D + j], c = gpu * gamma[2 * D + j];
      s1 += a + b + c; s2 += a * xr_ + b * xc + c * xu;
    }
    s1 = warp_sum(s1) / D3; s2 = warp_sum(s2) / D3;
    for (int j = lane; j < D; j += 32) {
      float xr_ = (to_f32(zr[j]) - mean) * rstd, xc = (to_f32(zr[D + j]) - mean) * rstd, xu = (to_f32(zr[2 * D + j]) - mean) * rstd;
      float pr = xr_ * gamma[j] + beta[j], pc = xc * gamma[D + j] + beta[D + j], pu = xu * gamma[2 * D + j] + beta[2 * D + j];
      float r = sigmoidf_(pr), cand = tanhf(r * pc), u = sigmoidf_(pu - 1.0f);
      float h = hprev[row * ldh + j];
      float g = dh[row * lddh + j];
      float drc = g * u * (1.0f - cand * cand);
      float gpr = drc * pc * r * (1.0f - r), gpc = drc * r, gpu = g * (cand - h) * u * (1.0f - u);
      dz[row * lddz + j] = from_f32<T>(rstd * (gpr * gamma[j] - s1 - xr_ * s2));
      dz[row * lddz + D + j] = from_f32<T>(rstd * (gpc * gamma[D + j] - s1 - xc * s2));
      dz[row * lddz + 2 * D + j] = from_f32<T>(rstd * (gpu * gamma[2 * D + j] - s1 - xu * s2));
      if (dhprev) dhprev[row * lddhp + j] = g * (1.0f - u);
      if (dgamma) {
        atomicAdd(&s_dg[j], gpr * xr_); atomicAdd(&s_dg[D + j], gpc * xc); atomicAdd(&s_dg[2 * D + j], gpu * xu);
        atomicAdd(&s_db[j], gpr); atomicAdd(&s_db[D + j], gpc); atomicAdd(&s_db[2 * D + j], gpu);
      }
    }
  }
  if (dgamma) {
    __syncthreads();
    for (int c = threadIdx.x; c < D3; c += THREADS) { atomicAdd(&dgamma[c], s_dg[c]); atomicAdd(&dbeta[c], s_db[c]); }
  }
}

// =================================================================================================
// Categorical latent   reference: OneHotDist dreamer_utils.py:177-197, get_dist :410-416
//   pi = normalise(0.99 softmax(logit) + 0.01/K);  idx = argmax(pi / q), q ~ Exp(1);  stoch = one_hot(idx)
// One warp per (row, latent); lanes stride over the K classes.
// =================================================================================================
struct CatProbs { float p, u, pi; };   // softmax prob, unimix, normalised

template <int MAXK_PER_LANE>
__device__ __forceinline__ void cat_probs(const float* lg, int K, int lane, float* p, float* pi, float& usum) {
  float mx = -INFINITY;
#pragma unroll
  for (int i = 0; i < MAXK_PER_LANE; ++i) { int k = lane + 32 * i; if (k < K) mx = fmaxf(mx, lg[k]); }
  mx = warp_max(mx);
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < MAXK_PER_LANE; ++i) { int k = lane + 32 * i; p[i] = k < K ? expf(lg[k] - mx) : 0.f; s += p[i]; }
  s = warp_sum(s);
  float us = 0.f;
#pragma unroll
  for (int i = 0; i < MAXK_PER_LANE; ++i) {
    int k = lane + 32 * i;
    p[i] = p[i] / s;
    pi[i] = k < K ? 0.99f * p[i] + (1.0f - 0.99f) * 1.0f / K : 0.f;
    us += pi[i];
  }
  usum = warp_sum(us);
#pragma unroll
  for (int i = 0; i < MAXK_PER_LANE; ++i) pi[i] = pi[i] / usum;
}

// given d(loss)/d(pi) (per lane) returns d(loss)/d(logit) through normalise -> unimix -> softmax
template <int MAXK_PER_LANE>
__device__ __forceinline__ void cat_probs_bwd(const float* p, const float* pi, float usum, const float* dpi, float* dl) {
  float dot = 0.f;
#pragma unroll
  for (int i = 0; i < MAXK_PER_LANE; ++i) dot += dpi[i] * pi[i];
  dot = warp_sum(dot);
  float dp[MAXK_PER_LANE];
  float dot2 = 0.f;
#pragma unroll
  for (int i = 0; i < MAXK_PER_LANE; ++i) { dp[i] = 0.99f * (dpi[i] - dot) / usum; dot2 += dp[i] * p[i]; }
  dot2 = warp_sum(dot2);
#pragma unroll
  for (int i = 0; i < MAXK_PER_LANE; ++i) dl[i] = p[i] * (dp[i] - dot2);
}

constexpr int KPL = 2;   // classes per lane -> K <= 64

template <typename T>
__global__ void __launch_bounds__(THREADS) cat_sample_fwd_kernel(const float* __restrict__ logit, const float* __restrict__ noise,
                                                                 unsigned long long seed, unsigned long long offset,
                                                                 const unsigned long long* __restrict__ step_dev, int sample,
                                                                 T* __restrict__ stoch, long long ld_stoch, float* __restrict__ stoch32,
                                                                 long long ld_stoch32, int* __restrict__ idx_out, long long rows, int S, int K) {
  genrl::pdl_sync();
  const int lane = threadIdx.x & 31;
  const long long total = rows * S;
  if (step_dev) seed += *step_dev * 0x9E3779B97F4A7C15ull;     // fresh stream every optimisation step (CUDA-graph safe)
  for (long long g = (long long)blockIdx.x * WARPS + (threadIdx.x >> 5); g < total; g += (long long)gridDim.x * WARPS) {
    const float* lg = logit + g * K;
    float p[KPL], pi[KPL], usum;
    cat_probs<KPL>(lg, K, lane, p, pi, usum);
    float best = -INFINITY; int bi = 0x7fffffff;
#pragma unroll
    for (int i = 0; i < KPL; ++i) {
      int k = lane + 32 * i;
      if (k < K) {
        float score;
        if (sample) {
          float q;
          if (noise) q = noise[g * K + k];
          else {
            uint4 r = philox4x32(make_uint4((uint32_t)(g * K + k), (uint32_t)((g * K + k) >> 32), (uint32_t)offset, (uint32_t)(offset >> 32)),
                                 make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
            q = -logf(u32_to_unit_open(r.x));
          }
          score = pi[i] / q;
        } else {
          score = logf(pi[i]);
        }
        if (score > best) { best = score; bi = k; }
      }
    }
    // warp argmax, first index wins on ties
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      float ob = __shfl_xor_sync(0xffffffffu, best, o);
      int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
    }
    const long long row = g / S; const int s = (int)(g % S);
#pragma unroll
    for (int i = 0; i < KPL; ++i) {
      int k = lane + 32 * i;
      if (k < K) {
        float v = k == bi ? 1.f : 0.f;
        if (stoch) stoch[row * ld_stoch + s * K + k] = from_f32<T>(v);
        if (stoch32) stoch32[row * ld_stoch32 + s * K + k] = v;
      }
    }
    if (idx_out && lane == 0) idx_out[g] = bi;
  }
}

// straight-through backward: dlogit (+)= J_pi^T dstoch   (sample) or J_logpi^T dstoch (mode)
template <typename T>
__global__ void __launch_bounds__(THREADS) cat_sample_bwd_kernel(const float* __restrict__ logit, const T* __restrict__ dstoch, long long ld_ds,
                                                                 int sample, float* __restrict__ dlogit, int accumulate, long long rows, int S, int K) {
  genrl::pdl_sync();
  const int lane = threadIdx.x & 31;
  const long long total = rows * S;
  for (long long g = (long long)blockIdx.x * WARPS + (threadIdx.x >> 5); g < total; g += (long long)gridDim.x * WARPS) {
    float p[KPL], pi[KPL], usum, dpi[KPL], dl[KPL];
    cat_probs<KPL>(logit + g * K, K, lane, p, pi, usum);
    const long long row = g / S; const int s = (int)(g % S);
#pragma unroll
    for (int i = 0; i < KPL; ++i) {
      int k = lane + 32 * i;
      float d = k < K ? to_f32(dstoch[row * ld_ds + s * K + k]) : 0.f;
      dpi[i] = sample ? d : (k < K ? d / pi[i] : 0.f);
    }
    cat_probs_bwd<KPL>(p, pi, usum, dpi, dl);
#pragma unroll
    for (int i = 0; i < KPL; ++i) {
      int k = lane + 32 * i;
      if (k < K) { float* o = dlogit + g * K + k; *o = accumulate ? *o + dl[i] : dl[i]; }
    }
  }
}


// ---- thread-per-latent fast paths for K == 32 (the GenRL latent): the 32 class values live in one thread's registers, so
//      there are no warp reductions at all and one Philox call yields 4 noise values ---------------------------------------
constexpr int CK = 32;
__device__ __forceinline__ void cat_probs_thr(const float* __restrict__ lg, float (&p)[CK], float (&pi)[CK], float& usum) {
#pragma unroll
  for (int i = 0; i < CK / 4; ++i) { float4 v = reinterpret_cast<const float4*>(lg)[i]; p[4 * i] = v.x; p[4 * i + 1] = v.y; p[4 * i + 2] = v.z; p[4 * i + 3] = v.w; }
  float mx = p[0];
#pragma unroll
  for (int k = 1; k < CK; ++k) mx = fmaxf(mx, p[k]);
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < CK; ++k) { p[k] = expf(p[k] - mx); s += p[k]; }
  float us = 0.f;
#pragma unroll
  for (int k = 0; k < CK; ++k) { p[k] = p[k] / s; pi[k] = 0.99f * p[k] + 0.01f / CK; us += pi[k]; }
  usum = us;
#pragma unroll
  for (int k = 0; k < CK; ++k) pi[k] = pi[k] / us;
}
__device__ __forceinline__ void cat_probs_bwd_thr(const float (&p)[CK], const float (&pi)[CK], float usum, const float (&dpi)[CK], float (&dl)[CK]) {
  float dot = 0.f;
#pragma unroll
  for (int k = 0; k < CK; ++k) dot += dpi[k] * pi[k];
  float dot2 = 0.f;
#pragma unroll
  for (int k = 0; k < CK; ++k) { dl[k] = 0.99f * (dpi[k] - dot) / usum; dot2 += dl[k] * p[k]; }
#pragma unroll
  for (int k = 0; k < CK; ++k) dl[k] = p[k] * (dl[k] - dot2);
}

// ---- four lanes per latent (K == 32: 8 classes per lane) ------------------------------------------------------------------------------------
// A thread-per-latent kernel reads / writes 128 contiguous bytes per THREAD (32 lines per warp instruction) and run 8 Philox calls and
// 32 logarithms in one dependent chain; at the rollout's 2 400 rows that is 17 us per launch.  Here a quad of lanes shares one latent: loads and
// stores are coalesced (a warp covers 8 latents = 1 KB), each lane does 2 Philox calls.  Sums run in the SAME sequential class order as the
// thread kernels (lane 0 adds its 8 values, hands the partial to lane 1, ...), so probabilities, noise mapping and the sampled index are
// bit-identical to the thread-per-latent form (cat_probs_thr, still used by the KL kernels and the generic backward).
__device__ __forceinline__ float quad_seq_sum(const float (&v)[8], int ql, int lane) {
  float acc = 0.f;
#pragma unroll
  for (int step = 0; step < 4; ++step) {
    if (ql == step) {
#pragma unroll
      for (int k = 0; k < 8; ++k) acc += v[k];
    }
    acc = __shfl_sync(0xffffffffu, acc, (lane & ~3) + step);
  }
  return acc;
}
__device__ __forceinline__ float quad_seq_dot(const float (&a)[8], const float (&b)[8], int ql, int lane) {
  float acc = 0.f;
#pragma unroll
  for (int step = 0; step < 4; ++step) {
    if (ql == step) {
#pragma unroll
      for (int k = 0; k < 8; ++k) acc += a[k] * b[k];
    }
    acc = __shfl_sync(0xffffffffu, acc, (lane & ~3) + step);
  }
  return acc;
}
__device__ __forceinline__ void cat_probs_quad(const float* __restrict__ lg8, float (&p)[8], float (&pi)[8], float& usum, int ql, int lane) {
  const float4 v0 = reinterpret_cast<const float4*>(lg8)[0], v1 = reinterpret_cast<const float4*>(lg8)[1];
  p[0] = v0.x; p[1] = v0.y; p[2] = v0.z; p[3] = v0.w; p[4] = v1.x; p[5] = v1.y; p[6] = v1.z; p[7] = v1.w;
  float mx = p[0];
#pragma unroll
  for (int k = 1; k < 8; ++k) mx = fmaxf(mx, p[k]);
  mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
  mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
#pragma unroll
  for (int k = 0; k < 8; ++k) p[k] = expf(p[k] - mx);
  const float s = quad_seq_sum(p, ql, lane);
#pragma unroll
  for (int k = 0; k < 8; ++k) { p[k] = p[k] / s; pi[k] = 0.99f * p[k] + 0.01f / CK; }
  const float us = quad_seq_sum(pi, ql, lane);
  usum = us;
#pragma unroll
  for (int k = 0; k < 8; ++k) pi[k] = pi[k] / us;
}

template <typename T>
__global__ void __launch_bounds__(128) cat_sample_fwd_quad_kernel(const float* __restrict__ logit, const float* __restrict__ noise,
                                                                  unsigned long long seed, unsigned long long offset,
                                                                  const unsigned long long* __restrict__ step_dev, int sample,
                                                                  T* __restrict__ stoch, long long ld_stoch, float* __restrict__ stoch32,
                                                                  long long ld_stoch32, int* __restrict__ idx_out, long long rows, int S) {
  genrl::pdl_sync();
  const long long t = (long long)blockIdx.x * 128 + threadIdx.x;
  const int lane = threadIdx.x & 31, ql = threadIdx.x & 3;
  const bool active = (t >> 2) < rows * S;
  const long long g = active ? (t >> 2) : 0;                  // inactive quads compute on latent 0 and store nothing (full-warp shuffles)
  if (step_dev) seed += *step_dev * 0x9E3779B97F4A7C15ull;
  float p[8], pi[8], usum;
  cat_probs_quad(logit + g * CK + ql * 8, p, pi, usum, ql, lane);
  float best = -INFINITY; int bi = 0;
  if (sample) {
#pragma unroll
    for (int jj = 0; jj < 2; ++jj) {
      const int j = ql * 2 + jj;
      float q[4];
      if (noise) { const float4 v = reinterpret_cast<const float4*>(noise + g * CK)[j]; q[0] = v.x; q[1] = v.y; q[2] = v.z; q[3] = v.w; }
      else {
        const unsigned long long c = (unsigned long long)g * (CK / 4) + j;
        const uint4 r = philox4x32(make_uint4((uint32_t)c, (uint32_t)(c >> 32), (uint32_t)offset, (uint32_t)(offset >> 32)), make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
        q[0] = -logf(u32_to_unit_open(r.x)); q[1] = -logf(u32_to_unit_open(r.y)); q[2] = -logf(u32_to_unit_open(r.z)); q[3] = -logf(u32_to_unit_open(r.w));
      }
#pragma unroll
      for (int e = 0; e < 4; ++e) { const float sc = pi[4 * jj + e] / q[e]; if (sc > best) { best = sc; bi = ql * 8 + 4 * jj + e; } }
    }
  } else {
#pragma unroll
    for (int k = 0; k < 8; ++k) { const float sc = logf(pi[k]); if (sc > best) { best = sc; bi = ql * 8 + k; } }
  }
  // first maximum in class order: on ties the lower class index wins (what the sequential strict-'>' scan of the thread kernel picks)
#pragma unroll
  for (int o = 1; o <= 2; o <<= 1) {
    const float ob = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
  }
  if (!active) return;
  const long long row = g / S; const int sidx = (int)(g % S);
  float oh[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) oh[k] = (ql * 8 + k) == bi ? 1.f : 0.f;
  if (stoch32) {
    float4* o = reinterpret_cast<float4*>(stoch32 + row * ld_stoch32 + sidx * CK + ql * 8);
    o[0] = make_float4(oh[0], oh[1], oh[2], oh[3]); o[1] = make_float4(oh[4], oh[5], oh[6], oh[7]);
  }
  if (stoch) V8<T>::store(stoch + row * ld_stoch + sidx * CK + ql * 8, oh);
  if (idx_out && ql == 0) idx_out[g] = bi;
}

// rollout backward: two gradient contributions (the feat slice of step i and the next step's input projection) summed on the fly, the logit
// gradient written in the activation dtype so that it is the dgrad GEMM's A operand as is (no accumulate pass, no cast pass)
template <typename T>
__global__ void __launch_bounds__(128) cat_sample_bwd2_quad_kernel(const float* __restrict__ logit, const T* __restrict__ ds1, long long ld1,
                                                                   const T* __restrict__ ds2, long long ld2, T* __restrict__ dl_out, long long ld_dl,
                                                                   long long rows, int S) {
  genrl::pdl_sync();
  const long long t = (long long)blockIdx.x * 128 + threadIdx.x;
  const int lane = threadIdx.x & 31, ql = threadIdx.x & 3;
  const bool active = (t >> 2) < rows * S;
  const long long g = active ? (t >> 2) : 0;
  float p[8], pi[8], usum, dpi[8], dl[8];
  cat_probs_quad(logit + g * CK + ql * 8, p, pi, usum, ql, lane);
  const long long row = g / S; const int sidx = (int)(g % S);
  V8<T>::load(ds1 + row * ld1 + sidx * CK + ql * 8, dpi);
  if (ds2) {
    float v8[8];
    V8<T>::load(ds2 + row * ld2 + sidx * CK + ql * 8, v8);
#pragma unroll
    for (int e = 0; e < 8; ++e) dpi[e] += v8[e];
  }
  const float dot = quad_seq_dot(dpi, pi, ql, lane);
#pragma unroll
  for (int k = 0; k < 8; ++k) dl[k] = 0.99f * (dpi[k] - dot) / usum;
  const float dot2 = quad_seq_dot(dl, p, ql, lane);
#pragma unroll
  for (int k = 0; k < 8; ++k) dl[k] = p[k] * (dl[k] - dot2);
  if (active) V8<T>::store(dl_out + row * ld_dl + sidx * CK + ql * 8, dl);
}

template <typename T>
__global__ void __launch_bounds__(128) cat_sample_bwd_thr_kernel(const float* __restrict__ logit, const T* __restrict__ dstoch, long long ld_ds,
                                                                 int sample, float* __restrict__ dlogit, int accumulate, long long rows, int S) {
  genrl::pdl_sync();
  const long long g = (long long)blockIdx.x * 128 + threadIdx.x;
  if (g >= rows * S) return;
  float p[CK], pi[CK], usum, dpi[CK], dl[CK];
  cat_probs_thr(logit + g * CK, p, pi, usum);
  const long long row = g / S; const int sidx = (int)(g % S);
#pragma unroll
  for (int i = 0; i < CK / 8; ++i) { float v8[8]; V8<T>::load(dstoch + row * ld_ds + sidx * CK + 8 * i, v8);
#pragma unroll
    for (int e = 0; e < 8; ++e) dpi[8 * i + e] = sample ? v8[e] : v8[e] / pi[8 * i + e]; }
  cat_probs_bwd_thr(p, pi, usum, dpi, dl);
  float4* o = reinterpret_cast<float4*>(dlogit + g * CK);
#pragma unroll
  for (int i = 0; i < CK / 4; ++i) {
    float4 v = make_float4(dl[4 * i], dl[4 * i + 1], dl[4 * i + 2], dl[4 * i + 3]);
    if (accumulate) { float4 a = o[i]; v.x += a.x; v.y += a.y; v.z += a.z; v.w += a.w; }
    o[i] = v;
  }
}

// S == 32, K == 32: one warp per row, one latent per lane
__global__ void __launch_bounds__(128) kl_fwd_thr_kernel(const float* __restrict__ lhs, const float* __restrict__ rhs, float free_bits,
                                                         float* __restrict__ value, float* __restrict__ out, long long rows) {
  genrl::pdl_sync();
  const long long g = (long long)blockIdx.x * 128 + threadIdx.x;
  const long long row = g >> 5;
  float kl = 0.f, el = 0.f, er = 0.f;
  if (row < rows) {
    float p[CK], P[CK], q[CK], Q[CK], us;
    cat_probs_thr(lhs + g * CK, p, P, us);
    cat_probs_thr(rhs + g * CK, q, Q, us);
#pragma unroll
    for (int k = 0; k < CK; ++k) { float lp = logf(P[k]), lq = logf(Q[k]); kl += P[k] * (lp - lq); el -= P[k] * lp; er -= Q[k] * lq; }
  }
  kl = warp_sum(kl); el = warp_sum(el); er = warp_sum(er);
  if ((threadIdx.x & 31) == 0 && row < rows) {
    value[row] = kl;
    const float inv = 1.0f / rows;
    atomicAdd(&out[0], fmaxf(kl, free_bits) * inv); atomicAdd(&out[1], kl * inv); atomicAdd(&out[2], el * inv); atomicAdd(&out[3], er * inv);
  }
}

__global__ void __launch_bounds__(128) kl_bwd_thr_kernel(const float* __restrict__ lhs, const float* __restrict__ rhs, const float* __restrict__ value,
                                                         float free_bits, float w_lhs, float w_rhs, const float* __restrict__ gscale,
                                                         float* __restrict__ dlhs, float* __restrict__ drhs, long long rows) {
  genrl::pdl_sync();
  const long long g = (long long)blockIdx.x * 128 + threadIdx.x;
  const long long row = g >> 5;
  if (row >= rows) return;
  const float gs = gscale ? *gscale : 1.0f;
  const float v = value[row];
  const float gate = v > free_bits ? 1.f : (v == free_bits ? 0.5f : 0.f);
  const float cl = gs * w_lhs * gate / rows, cr = gs * w_rhs * gate / rows;
  float p[CK], P[CK], q[CK], Q[CK], usp, usq, d[CK], dl[CK];
  cat_probs_thr(lhs + g * CK, p, P, usp);
  cat_probs_thr(rhs + g * CK, q, Q, usq);
#pragma unroll
  for (int k = 0; k < CK; ++k) d[k] = cl * (logf(P[k]) - logf(Q[k]) + 1.0f);
  cat_probs_bwd_thr(p, P, usp, d, dl);
  float4* o = reinterpret_cast<float4*>(dlhs + g * CK);
#pragma unroll
  for (int i = 0; i < CK / 4; ++i) o[i] = make_float4(dl[4 * i], dl[4 * i + 1], dl[4 * i + 2], dl[4 * i + 3]);
#pragma unroll
  for (int k = 0; k < CK; ++k) d[k] = -cr * P[k] / Q[k];
  cat_probs_bwd_thr(q, Q, usq, d, dl);
  o = reinterpret_cast<float4*>(drhs + g * CK);
#pragma unroll
  for (int i = 0; i < CK / 4; ++i) o[i] = make_float4(dl[4 * i], dl[4 * i + 1], dl[4 * i + 2], dl[4 * i + 3]);
}

// =================================================================================================
// Balanced KL with free bits + entropies   reference: EnsembleRSSM.kl_loss dreamer_utils.py:534-555, dreamer.py:249-250
//   value[r] = KL(lhs || rhs) summed over S;  loss = mix*mean(max(value,free)) + (1-mix)*mean(max(value,free))
//   (the two terms differ only in where the gradient flows).  out[0]=loss, out[1]=mean(value), out[2]=mean ent(lhs), out[3]=mean ent(rhs)
// One warp per row; loops the S latents.
// =================================================================================================
__global__ void __launch_bounds__(THREADS) kl_fwd_kernel(const float* __restrict__ lhs, const float* __restrict__ rhs, float free_bits,
                                                         float* __restrict__ value, float* __restrict__ out, long long rows, int S, int K) {
  genrl::pdl_sync();
  const int lane = threadIdx.x & 31;
  float acc_loss = 0.f, acc_val = 0.f, acc_el = 0.f, acc_er = 0.f;
  for (long long row = (long long)blockIdx.x * WARPS + (threadIdx.x >> 5); row < rows; row += (long long)gridDim.x * WARPS) {
    float kl = 0.f, el = 0.f, er = 0.f;
    for (int s = 0; s < S; ++s) {
      float p[KPL], P[KPL], q[KPL], Q[KPL], us;
      cat_probs<KPL>(lhs + (row * S + s) * K, K, lane, p, P, us);
      cat_probs<KPL>(rhs + (row * S + s) * K, K, lane, q, Q, us);
#pragma unroll
      for (int i = 0; i < KPL; ++i) {
        if (lane + 32 * i < K) {
          float lp = logf(P[i]), lq = logf(Q[i]);
          kl += P[i] * (lp - lq); el -= P[i] * lp; er -= Q[i] * lq;
        }
      }
    }
    kl = warp_sum(kl); el = warp_sum(el); er = warp_sum(er);
    if (lane == 0) { value[row] = kl; acc_loss += fmaxf(kl, free_bits); acc_val += kl; acc_el += el; acc_er += er; }
  }
  if (lane == 0) {
    const float inv = 1.0f / rows;
    atomicAdd(&out[0], acc_loss * inv); atomicAdd(&out[1], acc_val * inv);
    atomicAdd(&out[2], acc_el * inv); atomicAdd(&out[3], acc_er * inv);
  }
}

// dlhs = gscale*w_lhs/rows * 1[value>free] * dKL/dlhs ; drhs likewise with w_rhs.  (w = mix / 1-mix)
__global__ void __launch_bounds__(THREADS) kl_bwd_kernel(const float* __restrict__ lhs, const float* __restrict__ rhs, const float* __restrict__ value,
                                                         float free_bits, float w_lhs, float w_rhs, const float* __restrict__ gscale,
                                                         float* __restrict__ dlhs, float* __restrict__ drhs, long long rows, int S, int K) {
  genrl::pdl_sync();
  const int lane = threadIdx.x & 31;
  const float gs = gscale ? *gscale : 1.0f;
  for (long long row = (long long)blockIdx.x * WARPS + (threadIdx.x >> 5); row < rows; row += (long long)gridDim.x * WARPS) {
    const float v = value[row];
    // torch.maximum passes the gradient where value > free (half of it on exact ties)
    const float gate = v > free_bits ? 1.f : (v == free_bits ? 0.5f : 0.f);
    const float cl = gs * w_lhs * gate / rows, cr = gs * w_rhs * gate / rows;
    for (int s = 0; s < S; ++s) {
      float p[KPL], P[KPL], q[KPL], Q[KPL], usp, usq, dP[KPL], dQ[KPL], dl[KPL];
      const long long base = (row * S + s) * K;
      cat_probs<KPL>(lhs + base, K, lane, p, P, usp);
      cat_probs<KPL>(rhs + base, K, lane, q, Q, usq);
#pragma unroll
      for (int i = 0; i < KPL; ++i) {
        bool ok = lane + 32 * i < K;
        dP[i] = ok ? cl * (logf(P[i]) - logf(Q[i]) + 1.0f) : 0.f;
        dQ[i] = ok ? -cr * P[i] / Q[i] : 0.f;
      }
      cat_probs_bwd<KPL>(p, P, usp, dP, dl);
#pragma unroll
      for (int i = 0; i < KPL; ++i) if (lane + 32 * i < K) dlhs[base + lane + 32 * i] = dl[i];
      cat_probs_bwd<KPL>(q, Q, usq, dQ, dl);
#pragma unroll
      for (int i = 0; i < KPL; ++i) if (lane + 32 * i < K) drhs[base + lane + 32 * i] = dl[i];
    }
  }
}

// =================================================================================================
// Two-hot symlog head   reference: TwoHotDist dreamer_utils.py:120-171 (255 buckets on linspace(-20,20))
// =================================================================================================
constexpr int NB = 255;
__device__ __forceinline__ float bucket(int i) {
  // torch.linspace(-20, 20, 255): start + i*step for the first half, end - (n-1-i)*step for the second (symmetric fill)
  const float step = 40.0f / 254.0f;
  return i < NB / 2 ? -20.0f + step * i : 20.0f - step * (NB - 1 - i);
}
__device__ __forceinline__ float symlogf_(float x) { return copysignf(logf(fabsf(x) + 1.0f), x) * (x != 0.f); }
__device__ __forceinline__ float symexpf_(float y) { return copysignf(expf(fabsf(y)) - 1.0f, y) * (y != 0.f); }

// mode 0: out[r] = log_prob(x[r]) ; mode 1: out[r] = mean
__global__ void __launch_bounds__(THREADS) twohot_fwd_kernel(const float* __restrict__ logits, long long ldl, const float* __restrict__ x, int mode,
                                                             float* __restrict__ out, long long rows) {
  genrl::pdl_sync();
  const int lane = threadIdx.x & 31;
  for (long long row = (long long)blockIdx.x * WARPS + (threadIdx.x >> 5); row < rows; row += (long long)gridDim.x * WARPS) {
    const float* lg = logits + row * ldl;
    float v[8], mx = -INFINITY;
#pragma unroll
    for (int i = 0; i < 8; ++i) { int k = lane + 32 * i; v[i] = k < NB ? lg[k] : -INFINITY; mx = fmaxf(mx, v[i]); }
    mx = warp_max(mx);
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += (lane + 32 * i < NB) ? expf(v[i] - mx) : 0.f;
    s = warp_sum(s);
    if (mode == 1) {
      float m = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) { int k = lane + 32 * i; if (k < NB) m += expf(v[i] - mx) / s * bucket(k); }
      m = warp_sum(m);
      if (lane == 0) out[row] = symexpf_(m);
    } else {
      const float lse = mx + logf(s);
      const float xs = symlogf_(x[row]);
      int cnt_le = 0, cnt_gt = 0;
#pragma unroll
      for (int i = 0; i < 8; ++i) { int k = lane + 32 * i; if (k < NB) { cnt_le += bucket(k) <= xs; cnt_gt += bucket(k) > xs; } }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) { cnt_le += __shfl_xor_sync(0xffffffffu, cnt_le, o); cnt_gt += __shfl_xor_sync(0xffffffffu, cnt_gt, o); }
      int below = min(max(cnt_le - 1, 0), NB - 1), above = min(max(NB - cnt_gt, 0), NB - 1);
      float wb, wa;
      if (below == above) { wb = 0.5f; wa = 0.5f; }
      else { float db = fabsf(bucket(below) - xs), da = fabsf(bucket(above) - xs); wb = da / (db + da); wa = db / (db + da); }
      float acc = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        int k = lane + 32 * i;
        if (k < NB) { float t = (k == below ? wb : 0.f) + (k == above ? wa : 0.f); if (t != 0.f) acc += t * (v[i] - lse); }
      }
      acc = warp_sum(acc);
      if (lane == 0) out[row] = acc;
    }
  }
}

// dlogits = dout[r] * d out / d logits
__global__ void __launch_bounds__(THREADS) twohot_bwd_kernel(const float* __restrict__ logits, long long ldl, const float* __restrict__ x, int mode,
                                                             const float* __restrict__ dout, float scale, float* __restrict__ dlogits, long long lddl,
                                                             long long rows) {
  genrl::pdl_sync();
  const int lane = threadIdx.x & 31;
  for (long long row = (long long)blockIdx.x * WARPS + (threadIdx.x >> 5); row < rows; row += (long long)gridDim.x * WARPS) {
    const float* lg = logits + row * ldl;
    float v[8], mx = -INFINITY;
#pragma unroll
    for (int i = 0; i < 8; ++i) { int k = lane + 32 * i; v[i] = k < NB ? lg[k] : -INFINITY; mx = fmaxf(mx, v[i]); }
    mx = warp_max(mx);
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) { v[i] = (lane + 32 * i < NB) ? expf(v[i] - mx) : 0.f; s += v[i]; }
    s = warp_sum(s);
    const float g = dout[row] * scale;
    float* dl = dlogits + row * lddl;
    if (mode == 1) {
      float m = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) { int k = lane + 32 * i; if (k < NB) m += v[i] / s * bucket(k); }
      m = warp_sum(m);
      const float gm = g * expf(fabsf(m));        // d symexp(y)/dy = exp(|y|)
#pragma unroll
      for (int i = 0; i < 8; ++i) { int k = lane + 32 * i; if (k < NB) dl[k] = gm * (v[i] / s) * (bucket(k) - m); }
    } else {
      const float xs = symlogf_(x[row]);
      int cnt_le = 0, cnt_gt = 0;
#pragma unroll
      for (int i = 0; i < 8; ++i) { int k = lane + 32 * i; if (k < NB) { cnt_le += bucket(k) <= xs; cnt_gt += bucket(k) > xs; } }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) { cnt_le += __shfl_xor_sync(0xffffffffu, cnt_le, o); cnt_gt += __shfl_xor_sync(0xffffffffu, cnt_gt, o); }
      int below = min(max(cnt_le - 1, 0), NB - 1), above = min(max(NB - cnt_gt, 0), NB - 1);
      float wb, wa;
      if (below == above) { wb = 0.5f; wa = 0.5f; }
      else { float db = fabsf(bucket(below) - xs), da = fabsf(bucket(above) - xs); wb = da / (db + da); wa = db / (db + da); }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        int k = lane + 32 * i;
        if (k < NB) { float t = (k == below ? wb : 0.f) + (k == above ? wa : 0.f); dl[k] = g * (t - v[i] / s); }
      }
    }
  }
}

// =================================================================================================
// Normal actor head   reference: DistLayer 'normal' dreamer_utils.py:814-819; Independent(Normal).entropy
//   out[r, 0:A] -> mean = tanh ; out[r, A:2A] -> std = 0.9*sigmoid(.+2)+0.1 ; action = mean + std*eps
// =================================================================================================
template <typename T>
__global__ void actor_head_fwd_kernel(const float* __restrict__ out, long long ldo, const float* __restrict__ eps, int use_mean,
                                      float* __restrict__ action, long long lda, T* __restrict__ action_t, long long ldat,
                                      float* __restrict__ ent, long long rows, int A) {
  genrl::pdl_sync();
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows) return;
  float e = 0.f;
  for (int a = 0; a < A; ++a) {
    float mean = tanhf(out[i * ldo + a]);
    float std = 0.9f * sigmoidf_(out[i * ldo + A + a] + 2.0f) + 0.1f;
    float act = use_mean ? mean : mean + std * eps[i * A + a];
    if (action) action[i * lda + a] = act;
    if (action_t) action_t[i * ldat + a] = from_f32<T>(act);
    e += 0.5f + 0.9189385332046727f + logf(std);
  }
  if (ent) ent[i] = e;
}

template <typename T>
__global__ void actor_head_bwd_kernel(const float* __restrict__ out, long long ldo, const float* __restrict__ eps, const T* __restrict__ daction,
                                      long long ldda, const float* __restrict__ dent, float dent_scale, void* __restrict__ dout, long long lddo,
                                      int dout_dtype, long long rows, int A) {
  genrl::pdl_sync();
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows) return;
  const float de = dent ? dent[i] * dent_scale : dent_scale;
  for (int a = 0; a < A; ++a) {
    float mean = tanhf(out[i * ldo + a]);
    float sg = sigmoidf_(out[i * ldo + A + a] + 2.0f);
    float std = 0.9f * sg + 0.1f;
    float da = daction ? to_f32(daction[i * ldda + a]) : 0.f;
    float dstd = da * eps[i * A + a] + de / std;
    st_any(dout, dout_dtype, (size_t)(i * lddo + a), da * (1.0f - mean * mean));
    st_any(dout, dout_dtype, (size_t)(i * lddo + A + a), dstd * 0.9f * sg * (1.0f - sg));
  }
}

// =================================================================================================
// lambda-return   reference: lambda_return dreamer_utils.py:228-253 (reverse scan), ActorCritic.target dreamer.py:440-453
//   R_t = r_t + g*((1-l)*v_{t+1} + l*R_{t+1}),  R_H = v_H.   reward, value: [H+1, N]; target: [H, N]
// =================================================================================================
__global__ void lambda_return_fwd_kernel(const float* __restrict__ reward, const float* __restrict__ value, float gamma, float lambda,
                                         float* __restrict__ target, int H, long long N) {
  genrl::pdl_sync();
  long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  float last = value[(long long)H * N + n];
  for (int t = H - 1; t >= 0; --t) {
    float inp = reward[t * N + n] + gamma * value[(long long)(t + 1) * N + n] * (1.0f - lambda);
    last = inp + gamma * lambda * last;
    target[t * N + n] = last;
  }
}

// dtarget given as coef[t] (same for every n: the actor objective is a plain mean) or per-element array
__global__ void lambda_return_bwd_kernel(const float* __restrict__ dtarget, const float* __restrict__ dt_scale, float gamma, float lambda,
                                         float* __restrict__ dreward, float* __restrict__ dvalue, int H, long long N) {
  genrl::pdl_sync();
  long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  const float sc = dt_scale ? *dt_scale : 1.0f;
  float acc = 0.f;
  dvalue[n] = 0.f;
  for (int t = 0; t < H; ++t) {
    acc = dtarget[t * N + n] * sc + gamma * lambda * acc;
    dreward[t * N + n] = acc;
    float dv = gamma * (1.0f - lambda) * acc;
    if (t == H - 1) dv += gamma * lambda * acc;
    dvalue[(long long)(t + 1) * N + n] = dv;
  }
  dreward[(long long)H * N + n] = 0.f;
}

// =================================================================================================
// MSE image likelihood   reference: MSEDist.log_prob dreamer_utils.py:74-83 on (obs/255 - 0.5) dreamer.py:294-295
//   like[r] = -sum_{chw} (mean - (u8/255 - 0.5))^2 ;  dmean = -2 (mean - x) * dlike[r] * scale
// =================================================================================================
__global__ void __launch_bounds__(256) mse_fwd_kernel(const float* __restrict__ mean, const uint8_t* __restrict__ obs, float* __restrict__ like, int E) {
  genrl::pdl_sync();
  const long long r = blockIdx.x;
  float s = 0.f;
  for (int i = threadIdx.x; i < E; i += 256) {
    float x = (float)obs[r * E + i] / 255.0f - 0.5f;
    float d = mean[r * E + i] - x;
    s += d * d;
  }
  __shared__ float red[8];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) { float t = 0.f; for (int i = 0; i < 8; ++i) t += red[i]; like[r] = -t; }
}
__global__ void __launch_bounds__(256) mse_bwd_kernel(const float* __restrict__ mean, const uint8_t* __restrict__ obs, const float* __restrict__ dlike,
                                                      const float* __restrict__ gscale, float scale, float* __restrict__ dmean, int E) {
  genrl::pdl_sync();
  const long long r = blockIdx.x;
  const float g = (dlike ? dlike[r] : 1.0f) * scale * (gscale ? *gscale : 1.0f);
  for (int i = threadIdx.x; i < E; i += 256) {
    float x = (float)obs[r * E + i] / 255.0f - 0.5f;
    dmean[r * E + i] = -2.0f * (mean[r * E + i] - x) * g;
  }
}

// =================================================================================================
// small utilities
// =================================================================================================
// out[0] = mean, out[1] = unbiased std (torch .std()), single block
__global__ void __launch_bounds__(1024) mean_std_kernel(const float* __restrict__ x, long long n, float* __restrict__ out) {
  genrl::pdl_sync();
  __shared__ float red[32];
  __shared__ float s_mean;
  float s = 0.f;
  for (long long i = threadIdx.x; i < n; i += 1024) s += x[i];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) { float t = 0.f; for (int i = 0; i < 32; ++i) t += red[i]; s_mean = t / n; }
  __syncthreads();
  const float m = s_mean;
  float v = 0.f;
  for (long long i = threadIdx.x; i < n; i += 1024) { float d = x[i] - m; v += d * d; }
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x == 0) { float t = 0.f; for (int i = 0; i < 32; ++i) t += red[i]; out[0] = m; out[1] = n > 1 ? sqrtf(t / (n - 1)) : 0.f; }
}

// y[i] = a * x[i] + b   with optional device-side scalars (a_dev / b_dev) -> used for normalisation by RewardEMA
template <typename TI, typename TO>
__global__ void cast_kernel(const TI* __restrict__ x, TO* __restrict__ y, long long n) {
  genrl::pdl_sync();
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) y[i] = from_f32<TO>(to_f32(x[i]));
}

// strided 2-D copy with cast: dst[r, c] = src[r, c]   (used to assemble [stoch | action] / [x | h] operand buffers)
template <typename TI, typename TO>
__global__ void copy2d_kernel(const TI* __restrict__ src, long long lds, TO* __restrict__ dst, long long ldd, long long rows, int cols,
                              const float* __restrict__ row_scale) {
  genrl::pdl_sync();
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * cols) return;
  long long r = i / cols; int c = (int)(i % cols);
  float v = to_f32(src[r * lds + c]);
  if (row_scale) v *= row_scale[r];
  dst[r * ldd + c] = from_f32<TO>(v);
}

// dst[r, 0:cols] (+)= scale[r] * src[idx[r], 0:cols]   (idx < 0 -> zeros).  Time shifts of the RSSM inputs
// ("prev stoch" = posterior of t-1, masked by 1 - is_first: obs_step dreamer_utils.py:432-436) and their adjoints.
template <typename TI, typename TO>
__global__ void gather_rows_kernel(const TI* __restrict__ src, long long lds, const int* __restrict__ idx, const float* __restrict__ scale,
                                   TO* __restrict__ dst, long long ldd, long long rows, int cols, int accumulate) {
  genrl::pdl_sync();
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * cols) return;
  long long r = i / cols; int c = (int)(i % cols);
  long long sr = idx ? (long long)idx[r] : r;
  float v = sr < 0 ? 0.f : to_f32(src[sr * lds + c]) * (scale ? scale[r] : 1.f);
  TO* o = dst + r * ldd + c;
  *o = from_f32<TO>(accumulate ? to_f32(*o) + v : v);
}

// standard-normal fill (Box-Muller on Philox4x32-10): actor rsample noise, LAFITE noise of the connector
__global__ void __launch_bounds__(256) fill_normal_kernel(float* __restrict__ out, long long n, unsigned long long seed, unsigned long long offset,
                                                          const unsigned long long* __restrict__ step_dev) {
  genrl::pdl_sync();
  if (step_dev) seed += *step_dev * 0x9E3779B97F4A7C15ull;
  const long long quads = (n + 3) / 4;
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < quads; i += (long long)gridDim.x * 256) {
    uint4 r = philox4x32(make_uint4((uint32_t)i, (uint32_t)(i >> 32), (uint32_t)offset, (uint32_t)(offset >> 32) ^ 0x5EEDu),
                         make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
    float u0 = u32_to_unit_open(r.x), u1 = u32_to_unit_open(r.y), u2 = u32_to_unit_open(r.z), u3 = u32_to_unit_open(r.w);
    float m0 = sqrtf(-2.0f * logf(u0)), m1 = sqrtf(-2.0f * logf(u2));
    float z[4] = {m0 * cospif(2.0f * u1), m0 * sinpif(2.0f * u1), m1 * cospif(2.0f * u3), m1 * sinpif(2.0f * u3)};
#pragma unroll
    for (int k = 0; k < 4; ++k) if (i * 4 + k < n) out[i * 4 + k] = z[k];
  }
}
// fp32 -> three bf16 terms with x = hi + mid + lo to ~2^-24 relative (the operands of the "bf16x3" tensor-core GEMM of the fp32 parity mode)
__global__ void __launch_bounds__(256) split3_bf16_kernel(const float* __restrict__ src, long long lds, __nv_bfloat16* __restrict__ hi,
                                                          __nv_bfloat16* __restrict__ mid, __nv_bfloat16* __restrict__ lo, long long ldd,
                                                          long long rows, int cols) {
  genrl::pdl_sync();
  const long long total = rows * cols;
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
    const long long r = i / cols; const int c = (int)(i - r * cols);
    const float x = src[r * lds + c];
    const __nv_bfloat16 h = __float2bfloat16_rn(x);
    const float r1 = x - __bfloat162float(h);
    const __nv_bfloat16 m = __float2bfloat16_rn(r1);
    const float r2 = r1 - __bfloat162float(m);
    hi[r * ldd + c] = h; mid[r * ldd + c] = m; lo[r * ldd + c] = __float2bfloat16_rn(r2);
  }
}

__global__ void counter_add_kernel(unsigned long long* c, unsigned long long inc) {
  genrl::pdl_sync(); *c += inc; }

}  // namespace genrl

using namespace genrl;
#define ST(s) reinterpret_cast<cudaStream_t>(s)

static int num_sms_rows() {
  static int n = 0;
  if (n == 0) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev); if (n <= 0) n = 148; }
  return n;
}
static inline int ln_lpr(int D, int nch) { int l = 1; while (l * 8 * nch < D) l *= 2; return l; }
// the bulk-copy pipelines of ln_pipe.cuh pay from ~4 row tiles per SM (below that the register kernels' finer grid wins); GENRL_LN_PIPE=0 turns
// them off (A/B measurements)
static bool ln_pipe_enabled() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("GENRL_LN_PIPE"); v = (e && e[0] == '0') ? 0 : 1; }
  return v == 1;
}
static bool ln_pipe_fwd_on() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("GENRL_LN_PIPE_FWD"); v = (e && e[0] == '0') ? 0 : 1; }
  return v == 1;
}
// launchers of the compile-time-geometry pipelines: one function-local flag per instantiation for the opt-in shared-memory size
template <typename T, int LPR, int CPL, bool COLG>
static void ln_bwd_pipe_launch(cudaStream_t st, const T* dy, long long lddy, const T* x, long long ldx, const float* gamma, const float* beta, const float* stats,
                               int act, T* dx, long long lddx, float* dgamma, float* dbeta, float* dlb, long long M) {
  using G = LnpGeom<T, LPR, CPL>;
  static bool attr_set = false;
  if (!attr_set) { cudaFuncSetAttribute(ln_bwd_pipe_kernel<T, LPR, CPL, COLG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G::BWD_SMEM); attr_set = true; }
  const long long ntiles = (M + G::R - 1) / G::R;
  genrl::launch_pdl(ln_bwd_pipe_kernel<T, LPR, CPL, COLG>, (int)std::min<long long>(ntiles, (long long)LNP_BLOCKS_PER_SM * num_sms_rows()), LNP_THREADS, G::BWD_SMEM, st, dy, lddy, x, ldx, gamma, beta,
                    stats, act, dx, lddx, dgamma, dbeta, dlb, M);
}
template <typename T, int LPR, int CPL>
static void ln_fwd_pipe_launch(cudaStream_t st, const T* x, long long ldx, const float* gamma, const float* beta, float eps, int act, T* y, long long ldy,
                               float* stats, long long M) {
  using G = LnpGeom<T, LPR, CPL>;
  static bool attr_set = false;
  if (!attr_set) { cudaFuncSetAttribute(ln_fwd_pipe_kernel<T, LPR, CPL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G::FWD_SMEM); attr_set = true; }
  const long long ntiles = (M + G::FR - 1) / G::FR;
  genrl::launch_pdl(ln_fwd_pipe_kernel<T, LPR, CPL>, (int)std::min<long long>(ntiles, (long long)LNP_BLOCKS_PER_SM * num_sms_rows()), LNP_THREADS, G::FWD_SMEM, st, x, ldx, gamma, beta, eps, act, y, ldy,
                    stats, M);
}
// eligible when the width has an instantiation and there are at least ~4 row tiles per SM (16 384 elements per tile)
template <typename T> static bool ln_pipe_ok(long long M, int D) {
  int lpr, cpl;
  return ln_pipe_enabled() && lnp_geom(D, &lpr, &cpl) && M * D >= (long long)4 * num_sms_rows() * (sizeof(T) == 2 ? 16384 : 8192);
}

extern "C" int genrl_ln_act_fwd(void* stream, const void* x, long long ldx, const float* gamma, const float* beta, float eps, int act,
                                void* y, long long ldy, float* stats, long long M, int D, int dtype) {
  if (M <= 0) return 0;
  DISPATCH_T(dtype, {
    if (ln_pipe_fwd_on() && ln_pipe_ok<T>(M, D) && vec_ok<T>(x, ldx) && vec_ok<T>(y, ldy)) {
      LNP_DISPATCH(D, (ln_fwd_pipe_launch<T, LPR, CPL>(ST(stream), (const T*)x, ldx, gamma, beta, eps, act, (T*)y, ldy, stats, M)));
    } else if (D % 8 == 0 && D <= 1024 && vec_ok<T>(x, ldx) && vec_ok<T>(y, ldy) && f32_vec_ok(gamma) && f32_vec_ok(beta)) {
      const int nch = D <= 384 ? 6 : 4;
      const int lpr = ln_lpr(D, nch);
      const long long rpb = 8 * (32 / lpr);
      const int grid = (int)std::min<long long>((M + rpb - 1) / rpb, 148 * 16);
      if (nch == 6) genrl::launch_pdl(ln_fwd_fast_kernel<T, 6>, grid, 256, 0, ST(stream), (const T*)x, ldx, gamma, beta, eps, act, (T*)y, ldy, stats, M, D, lpr);
      else genrl::launch_pdl(ln_fwd_fast_kernel<T, 4>, grid, 256, 0, ST(stream), (const T*)x, ldx, gamma, beta, eps, act, (T*)y, ldy, stats, M, D, lpr);
    } else {
      genrl::launch_pdl(ln_act_fwd_kernel<T>, row_grid(M, 148 * 16), THREADS, 0, ST(stream), (const T*)x, ldx, gamma, beta, eps, act, (T*)y, ldy, stats, M, D);
    }
  });
  return check_launch("ln_act_fwd");
}

extern "C" int genrl_ln_act_bwd(void* stream, const void* dy, long long lddy, const void* x, long long ldx, const float* gamma, const float* beta,
                                const float* stats, int act, void* dx, long long lddx, float* dgamma, float* dbeta, float* dlin_bias,
                                long long M, int D, int dtype) {
  if (M <= 0) return 0;
  GENRL_REQUIRE(3 * D * 4 <= 200 * 1024, "ln_act_bwd: D too large");
  size_t smem = (size_t)3 * D * sizeof(float);
  DISPATCH_T(dtype, {
    if (ln_pipe_ok<T>(M, D) && vec_ok<T>(x, ldx) && vec_ok<T>(dy, lddy) && vec_ok<T>(dx, lddx) && (reinterpret_cast<uintptr_t>(stats) & 7) == 0) {
      // large row counts: bulk-copy pipeline through shared memory (ln_pipe.cuh), one persistent 512-thread block per SM
      if (dgamma || dlin_bias) {
        LNP_DISPATCH(D, (ln_bwd_pipe_launch<T, LPR, CPL, true>(ST(stream), (const T*)dy, lddy, (const T*)x, ldx, gamma, beta, stats, act, (T*)dx, lddx, dgamma, dbeta,
                                                               dlin_bias, M)));
      } else {
        LNP_DISPATCH(D, (ln_bwd_pipe_launch<T, LPR, CPL, false>(ST(stream), (const T*)dy, lddy, (const T*)x, ldx, gamma, beta, stats, act, (T*)dx, lddx, (float*)nullptr,
                                                                (float*)nullptr, (float*)nullptr, M)));
      }
    } else if (D % 8 == 0 && D <= 1024 && vec_ok<T>(x, ldx) && vec_ok<T>(dy, lddy) && vec_ok<T>(dx, lddx) && f32_vec_ok(gamma) && f32_vec_ok(beta)) {
      if (dgamma || dlin_bias) {
        // one pass: dx and the parameter gradients (narrow rows use 2 chunks per lane: 128 registers, two blocks per SM)
        const int nch = D <= 384 ? 2 : 4;
        const int lpr = ln_lpr(D, nch);
        const long long rpb = 8 * (32 / lpr);
        static int bps2 = 0, bps4 = 0;
        int& bps = nch == 2 ? bps2 : bps4;
        if (bps == 0) {
          if (nch == 2) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, ln_bwd_colg_kernel<T, 2>, 256, 0);
          else cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, ln_bwd_colg_kernel<T, 4>, 256, 0);
          if (bps <= 0) bps = 1;
        }
        const int grid = (int)std::min<long long>((M + rpb - 1) / rpb, (long long)num_sms_rows() * bps);
        if (nch == 2) genrl::launch_pdl(ln_bwd_colg_kernel<T, 2>, grid, 256, 0, ST(stream), (const T*)dy, lddy, (const T*)x, ldx, gamma, beta, stats, act, (T*)dx, lddx,
                                                                             dgamma, dbeta, dlin_bias, M, D, lpr);
        else genrl::launch_pdl(ln_bwd_colg_kernel<T, 4>, grid, 256, 0, ST(stream), (const T*)dy, lddy, (const T*)x, ldx, gamma, beta, stats, act, (T*)dx, lddx,
                                                                    dgamma, dbeta, dlin_bias, M, D, lpr);
      } else {
        const int nch = D <= 384 ? 6 : 4;
        const int lpr = ln_lpr(D, nch);
        const long long rpb = 8 * (32 / lpr);
        const int grid = (int)std::min<long long>((M + rpb - 1) / rpb, 148 * 16);
        const long long rows_wave4 = (long long)num_sms_rows() * 5 * 4 * (32 / lpr);      // rows one wave of the 4-warp variant covers
        if (nch == 6) genrl::launch_pdl(ln_bwd_fast_kernel<T, 6, 8>, grid, 256, 0, ST(stream), (const T*)dy, lddy, (const T*)x, ldx, gamma, beta, stats, act, (T*)dx, lddx, M, D, lpr);
        else if (M > (long long)num_sms_rows() * 2 * rpb && M <= rows_wave4)
          genrl::launch_pdl(ln_bwd_fast_kernel<T, 4, 4>, (int)((M + rpb / 2 - 1) / (rpb / 2)), 128, 0, ST(stream), (const T*)dy, lddy, (const T*)x, ldx, gamma, beta, stats, act,
                            (T*)dx, lddx, M, D, lpr);
        else genrl::launch_pdl(ln_bwd_fast_kernel<T, 4, 8>, grid, 256, 0, ST(stream), (const T*)dy, lddy, (const T*)x, ldx, gamma, beta, stats, act, (T*)dx, lddx, M, D, lpr);
      }
    } else {
      if (smem > 48 * 1024) cudaFuncSetAttribute(ln_act_bwd_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      genrl::launch_pdl(ln_act_bwd_kernel<T>, row_grid(M, 148 * 2), THREADS, smem, ST(stream), (const T*)dy, lddy, (const T*)x, ldx, gamma, beta, stats, act, (T*)dx,
                                                                                lddx, dgamma, dbeta, dlin_bias, M, D);
    }
  });
  return check_launch("ln_act_bwd");
}

// column pass alone (deferred parameter gradients: one launch over the rows of all imagination steps)
extern "C" int genrl_ln_colgrad(void* stream, const void* dy, long long lddy, const void* x, long long ldx, const void* dx, long long lddx,
                                const float* gamma, const float* beta, const float* stats, int act, float* dgamma, float* dbeta, float* dlin_bias,
                                long long M, int D, int dtype) {
  if (M <= 0) return 0;
  DISPATCH_T(dtype, {
    if (D % 8 == 0 && vec_ok<T>(x, ldx) && vec_ok<T>(dy, lddy) && vec_ok<T>(dx, lddx) && f32_vec_ok(gamma) && f32_vec_ok(beta)) {
      dim3 g2(ceil_div(M, COLV_ROWS), ceil_div(D, 128));
      genrl::launch_pdl(ln_colgrad_vec_kernel<T>, g2, 256, 0, ST(stream), (const T*)dy, lddy, (const T*)x, ldx, (const T*)dx, lddx, gamma, beta, stats, act,
                                                          dgamma, dbeta, dlin_bias, M, D);
    } else {
      dim3 g2(ceil_div(M, COL_RC), ceil_div(D, 32));
      genrl::launch_pdl(ln_colgrad_kernel<T>, g2, 256, 0, ST(stream), (const T*)dy, lddy, (const T*)x, ldx, (const T*)dx, lddx, gamma, beta, stats, act, dgamma,
                                                      dbeta, dlin_bias, M, D);
    }
  });
  return check_launch("ln_colgrad");
}

extern "C" int genrl_colsum(void* stream, const void* x, long long ldx, float* out, long long M, int D, int dtype) {
  if (M <= 0) return 0;
  DISPATCH_T(dtype, {
    if (D % 8 == 0 && vec_ok<T>(x, ldx)) {
      dim3 grid(ceil_div(M, COLV_ROWS), ceil_div(D, 128));
      genrl::launch_pdl(colsum_vec_kernel<T>, grid, 256, 0, ST(stream), (const T*)x, ldx, out, M, D);
    } else {
      dim3 grid(ceil_div(M, COL_RC), ceil_div(D, 32));
      genrl::launch_pdl(colsum_fast_kernel<T>, grid, 256, 0, ST(stream), (const T*)x, ldx, out, M, D);
    }
  });
  return check_launch("colsum");
}

extern "C" int genrl_gru_fwd(void* stream, const void* z, long long ldz, const float* gamma, const float* beta, const float* hprev, long long ldh,
                             float* hout, long long ldho, void* hout_t, long long ldht, float* stats, long long M, int D, int dtype) {
  if (M <= 0) return 0;
  DISPATCH_T(dtype, {
    if (D % 8 == 0 && vec_ok<T>(z, ldz) && vec_ok<float>(hprev, ldh) && vec_ok<float>(hout, ldho) && (!hout_t || vec_ok<T>(hout_t, ldht)) &&
        f32_vec_ok(gamma) && f32_vec_ok(beta)) {
      genrl::launch_pdl(gru_fwd_fast_kernel<T>, row_grid(M, 148 * 16), 256, 0, ST(stream), (const T*)z, ldz, gamma, beta, hprev, ldh, hout, ldho, (T*)hout_t, ldht, stats, M, D);
    } else {
      genrl::launch_pdl(gru_fwd_kernel<T>, row_grid(M, 148 * 16), THREADS, 0, ST(stream), (const T*)z, ldz, gamma, beta, hprev, ldh, hout, ldho, (T*)hout_t, ldht, stats, M, D);
    }
  });
  return check_launch("gru_fwd");
}

extern "C" int genrl_gru_bwd(void* stream, const float* dh, long long lddh, const void* z, long long ldz, const float* gamma, const float* beta,
                             const float* hprev, long long ldh, const float* stats, void* dz, long long lddz, float* dhprev, long long lddhp,
                             float* dgamma, float* dbeta, long long M, int D, int dtype) {
  if (M <= 0) return 0;
  size_t smem = (size_t)6 * D * sizeof(float);
  GENRL_REQUIRE(smem <= 200 * 1024, "gru_bwd: D too large");
  DISPATCH_T(dtype, {
    if (!dgamma && D % 8 == 0 && vec_ok<T>(z, ldz) && vec_ok<T>(dz, lddz) && vec_ok<float>(hprev, ldh) && vec_ok<float>(dh, lddh) &&
        (!dhprev || vec_ok<float>(dhprev, lddhp)) && f32_vec_ok(gamma) && f32_vec_ok(beta)) {
      if (D == 1024)
        genrl::launch_pdl(gru_bwd_row1024_kernel<T>, (int)std::min<long long>(M, 148 * 16), 128, 0, ST(stream), dh, lddh, (const T*)z, ldz, gamma, beta, hprev, ldh, stats,
                                                                                               (T*)dz, lddz, dhprev, lddhp, M, (const T*)nullptr, 0LL, (const T*)nullptr, 0LL, (const T*)nullptr, 0LL);
      else
        genrl::launch_pdl(gru_bwd_fast_kernel<T>, row_grid(M, 148 * 16), 256, 0, ST(stream), dh, lddh, (const T*)z, ldz, gamma, beta, hprev, ldh, stats, (T*)dz, lddz,
                                                                             dhprev, lddhp, M, D);
    } else {
      if (smem > 48 * 1024) cudaFuncSetAttribute(gru_bwd_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      genrl::launch_pdl(gru_bwd_kernel<T>, row_grid(M, 148 * 2), THREADS, smem, ST(stream), dh, lddh, (const T*)z, ldz, gamma, beta, hprev, ldh, stats, (T*)dz, lddz,
                                                                             dhprev, lddhp, dgamma, dbeta, M, D);
    }
  });
  return check_launch("gru_bwd");
}

// frozen-weights GRU backward at D = 1024 whose incoming gradient is dh + e1 + e2 + e3 (e* optional, activation dtype).  Returns 2 when the
// shape / alignment is not eligible (the caller then sums the contributions itself and calls genrl_gru_bwd).
extern "C" int genrl_gru_bwd_sum(void* stream, const float* dh, long long lddh, const void* e1, long long lde1, const void* e2, long long lde2,
                                 const void* e3, long long lde3,
                                 const void* z, long long ldz, const float* gamma, const float* beta, const float* hprev, long long ldh,
                                 const float* stats, void* dz, long long lddz, float* dhprev, long long lddhp, long long M, int D, int dtype) {
  if (M <= 0) return 0;
  if (D != 1024) return 2;
  DISPATCH_T(dtype, {
    if (!(vec_ok<T>(z, ldz) && vec_ok<T>(dz, lddz) && vec_ok<float>(hprev, ldh) && vec_ok<float>(dh, lddh) && (!dhprev || vec_ok<float>(dhprev, lddhp)) &&
          (!e1 || vec_ok<T>(e1, lde1)) && (!e2 || vec_ok<T>(e2, lde2)) && (!e3 || vec_ok<T>(e3, lde3)) && f32_vec_ok(gamma) && f32_vec_ok(beta)))
      return 2;
    genrl::launch_pdl(gru_bwd_row1024_kernel<T>, (int)std::min<long long>(M, 148 * 16), 128, 0, ST(stream), dh, lddh, (const T*)z, ldz, gamma, beta, hprev, ldh, stats,
                      (T*)dz, lddz, dhprev, lddhp, M, (const T*)e1, lde1, (const T*)e2, lde2, (const T*)e3, lde3);
  });
  return check_launch("gru_bwd_sum");
}

extern "C" int genrl_cat_sample_fwd(void* stream, const float* logit, const float* noise, unsigned long long seed, unsigned long long offset,
                                    const unsigned long long* step_dev, int sample, void* stoch, long long ld_stoch, float* stoch32, long long ld_stoch32, int* idx, long long rows,
                                    int S, int K, int dtype) {
  if (rows <= 0) return 0;
  GENRL_REQUIRE(K <= 32 * KPL, "cat_sample: K=%d > %d unsupported", K, 32 * KPL);
  const bool fast_ok = K == CK && f32_vec_ok(logit) && (!noise || f32_vec_ok(noise)) && (!stoch32 || vec_ok<float>(stoch32, ld_stoch32));
  if (fast_ok) {
    DISPATCH_T(dtype, {
      if (!stoch || vec_ok<T>(stoch, ld_stoch)) {
        genrl::launch_pdl(cat_sample_fwd_quad_kernel<T>, ceil_div(rows * S * 4, 128), 128, 0, ST(stream), logit, noise, seed, offset, step_dev, sample, (T*)stoch, ld_stoch,
                                                                                      stoch32, ld_stoch32, idx, rows, S);
        return check_launch("cat_sample_fwd");
      }
    });
  }
  DISPATCH_T(dtype, (genrl::launch_pdl(cat_sample_fwd_kernel<T>, row_grid(rows * S, 148 * 16), THREADS, 0, ST(stream), logit, noise, seed, offset, step_dev, sample, (T*)stoch, ld_stoch,
                                                                                                     stoch32, ld_stoch32, idx, rows, S, K)));
  return check_launch("cat_sample_fwd");
}

extern "C" int genrl_cat_sample_bwd(void* stream, const float* logit, const void* dstoch, long long ld_ds, int sample, float* dlogit, int accumulate,
                                    long long rows, int S, int K, int dtype) {
  if (rows <= 0) return 0;
  GENRL_REQUIRE(K <= 32 * KPL, "cat_sample: K=%d > %d unsupported", K, 32 * KPL);
  if (K == CK && f32_vec_ok(logit) && f32_vec_ok(dlogit)) {
    DISPATCH_T(dtype, {
      if (vec_ok<T>(dstoch, ld_ds)) {
        genrl::launch_pdl(cat_sample_bwd_thr_kernel<T>, ceil_div(rows * S, 128), 128, 0, ST(stream), logit, (const T*)dstoch, ld_ds, sample, dlogit, accumulate, rows, S);
        return check_launch("cat_sample_bwd");
      }
    });
  }
  DISPATCH_T(dtype, (genrl::launch_pdl(cat_sample_bwd_kernel<T>, row_grid(rows * S, 148 * 16), THREADS, 0, ST(stream), logit, (const T*)dstoch, ld_ds, sample, dlogit, accumulate, rows, S, K)));
  return check_launch("cat_sample_bwd");
}

// sampled one-hot (K = 32), straight-through backward of ds1 + ds2 (ds2 may be NULL), output in the same dtype as the inputs
extern "C" int genrl_cat_sample_bwd2(void* stream, const float* logit, const void* ds1, long long ld1, const void* ds2, long long ld2, void* dlogit,
                                     long long ld_dl, long long rows, int S, int K, int dtype) {
  if (rows <= 0) return 0;
  GENRL_REQUIRE(K == CK && f32_vec_ok(logit), "cat_sample_bwd2: K must be %d and logit 16-byte aligned", CK);
  DISPATCH_T(dtype, {
    GENRL_REQUIRE(vec_ok<T>(ds1, ld1) && (!ds2 || vec_ok<T>(ds2, ld2)) && vec_ok<T>(dlogit, ld_dl), "cat_sample_bwd2: operands must be 16-byte aligned");
    genrl::launch_pdl(cat_sample_bwd2_quad_kernel<T>, ceil_div(rows * S * 4, 128), 128, 0, ST(stream), logit, (const T*)ds1, ld1, (const T*)ds2, ld2, (T*)dlogit, ld_dl, rows, S);
  });
  return check_launch("cat_sample_bwd2");
}

extern "C" int genrl_kl_fwd(void* stream, const float* lhs, const float* rhs, float free_bits, float* value, float* out4, long long rows, int S, int K) {
  if (rows <= 0) return 0;
  GENRL_REQUIRE(K <= 32 * KPL, "kl: K=%d unsupported", K);
  cudaMemsetAsync(out4, 0, 4 * sizeof(float), ST(stream));
  if (S == 32 && K == CK && f32_vec_ok(lhs) && f32_vec_ok(rhs)) {
    genrl::launch_pdl(kl_fwd_thr_kernel, ceil_div(rows * 32, 128), 128, 0, ST(stream), lhs, rhs, free_bits, value, out4, rows);
    return check_launch("kl_fwd");
  }
  genrl::launch_pdl(kl_fwd_kernel, row_grid(rows, 148 * 4), THREADS, 0, ST(stream), lhs, rhs, free_bits, value, out4, rows, S, K);
  return check_launch("kl_fwd");
}

extern "C" int genrl_kl_bwd(void* stream, const float* lhs, const float* rhs, const float* value, float free_bits, float w_lhs, float w_rhs,
                            const float* gscale, float* dlhs, float* drhs, long long rows, int S, int K) {
  if (rows <= 0) return 0;
  if (S == 32 && K == CK && f32_vec_ok(lhs) && f32_vec_ok(rhs) && f32_vec_ok(dlhs) && f32_vec_ok(drhs)) {
    genrl::launch_pdl(kl_bwd_thr_kernel, ceil_div(rows * 32, 128), 128, 0, ST(stream), lhs, rhs, value, free_bits, w_lhs, w_rhs, gscale, dlhs, drhs, rows);
    return check_launch("kl_bwd");
  }
  genrl::launch_pdl(kl_bwd_kernel, row_grid(rows, 148 * 8), THREADS, 0, ST(stream), lhs, rhs, value, free_bits, w_lhs, w_rhs, gscale, dlhs, drhs, rows, S, K);
  return check_launch("kl_bwd");
}

extern "C" int genrl_twohot_fwd(void* stream, const float* logits, long long ldl, const float* x, int mode, float* out, long long rows) {
  if (rows <= 0) return 0;
  genrl::launch_pdl(twohot_fwd_kernel, row_grid(rows, 148 * 16), THREADS, 0, ST(stream), logits, ldl, x, mode, out, rows);
  return check_launch("twohot_fwd");
}

extern "C" int genrl_twohot_bwd(void* stream, const float* logits, long long ldl, const float* x, int mode, const float* dout, float scale,
                                float* dlogits, long long lddl, long long rows) {
  if (rows <= 0) return 0;
  genrl::launch_pdl(twohot_bwd_kernel, row_grid(rows, 148 * 16), THREADS, 0, ST(stream), logits, ldl, x, mode, dout, scale, dlogits, lddl, rows);
  return check_launch("twohot_bwd");
}

extern "C" int genrl_actor_head_fwd(void* stream, const float* out, long long ldo, const float* eps, int use_mean, float* action, long long lda,
                                    void* action_t, long long ldat, float* ent, long long rows, int A, int dtype) {
  if (rows <= 0) return 0;
  DISPATCH_T(dtype, (genrl::launch_pdl(actor_head_fwd_kernel<T>, ceil_div(rows, 128), 128, 0, ST(stream), out, ldo, eps, use_mean, action, lda, (T*)action_t, ldat, ent, rows, A)));
  return check_launch("actor_head_fwd");
}

extern "C" int genrl_actor_head_bwd(void* stream, const float* out, long long ldo, const float* eps, const void* daction, long long ldda,
                                    const float* dent, float dent_scale, void* dout, long long lddo, int dout_dtype, long long rows, int A, int dtype) {
  if (rows <= 0) return 0;
  GENRL_REQUIRE(dout_dtype == GENRL_F32 || dout_dtype == GENRL_BF16, "actor_head_bwd: bad output dtype %d", dout_dtype);
  DISPATCH_T(dtype, (genrl::launch_pdl(actor_head_bwd_kernel<T>, ceil_div(rows, 128), 128, 0, ST(stream), out, ldo, eps, (const T*)daction, ldda, dent, dent_scale, dout, lddo, dout_dtype, rows, A)));
  return check_launch("actor_head_bwd");
}

extern "C" int genrl_lambda_return_fwd(void* stream, const float* reward, const float* value, float gamma, float lambda, float* target, int H, long long N) {
  if (N <= 0) return 0;
  genrl::launch_pdl(lambda_return_fwd_kernel, ceil_div(N, 128), 128, 0, ST(stream), reward, value, gamma, lambda, target, H, N);
  return check_launch("lambda_return_fwd");
}

extern "C" int genrl_lambda_return_bwd(void* stream, const float* dtarget, const float* dt_scale, float gamma, float lambda, float* dreward,
                                       float* dvalue, int H, long long N) {
  if (N <= 0) return 0;
  genrl::launch_pdl(lambda_return_bwd_kernel, ceil_div(N, 128), 128, 0, ST(stream), dtarget, dt_scale, gamma, lambda, dreward, dvalue, H, N);
  return check_launch("lambda_return_bwd");
}

extern "C" int genrl_mse_fwd(void* stream, const float* mean, const unsigned char* obs, float* like, long long rows, int E) {
  if (rows <= 0) return 0;
  genrl::launch_pdl(mse_fwd_kernel, (unsigned)rows, 256, 0, ST(stream), mean, obs, like, E);
  return check_launch("mse_fwd");
}

extern "C" int genrl_mse_bwd(void* stream, const float* mean, const unsigned char* obs, const float* dlike, const float* gscale, float scale,
                             float* dmean, long long rows, int E) {
  if (rows <= 0) return 0;
  genrl::launch_pdl(mse_bwd_kernel, (unsigned)rows, 256, 0, ST(stream), mean, obs, dlike, gscale, scale, dmean, E);
  return check_launch("mse_bwd");
}

extern "C" int genrl_mean_std(void* stream, const float* x, long long n, float* out2) {
  genrl::launch_pdl(mean_std_kernel, 1, 1024, 0, ST(stream), x, n, out2);
  return check_launch("mean_std");
}

// dst[r, 0:cols] = cast(src[r, 0:cols]) * row_scale[r]      dtypes: GENRL_F32 / GENRL_BF16 each side
template <typename TI, typename TO>
static bool try_fast_gather(cudaStream_t st, const void* src, long long lds, const int* idx, const float* scale, void* dst, long long ldd, long long rows,
                            int cols, int accumulate) {
  if (cols % 8 != 0 || !vec_ok<TI>(src, lds) || !vec_ok<TO>(dst, ldd)) return false;
  const long long total = rows * (cols / 8);
  const unsigned grid = (unsigned)std::min<long long>((total + 255) / 256, 148 * 32);
  genrl::launch_pdl(gather_rows_fast_kernel<TI, TO>, grid, 256, 0, st, (const TI*)src, lds, idx, scale, (TO*)dst, ldd, rows, cols / 8, accumulate);
  return true;
}

template <typename TI, typename TO>
static void run_gather(cudaStream_t st, const void* src, long long lds, const int* idx, const float* scale, void* dst, long long ldd, long long rows,
                       int cols, int accumulate) {
  if (try_fast_gather<TI, TO>(st, src, lds, idx, scale, dst, ldd, rows, cols, accumulate)) return;
  unsigned grid = (unsigned)ceil_div(rows * cols, 256);
  genrl::launch_pdl(gather_rows_kernel<TI, TO>, grid, 256, 0, st, (const TI*)src, lds, idx, scale, (TO*)dst, ldd, rows, cols, accumulate);
}

extern "C" int genrl_gather_rows(void* stream, const void* src, long long lds, int src_dtype, const int* idx, const float* scale, void* dst,
                                 long long ldd, int dst_dtype, long long rows, int cols, int accumulate) {
  if (rows <= 0 || cols <= 0) return 0;
  cudaStream_t st = ST(stream);
  if (src_dtype == GENRL_F32 && dst_dtype == GENRL_F32) run_gather<float, float>(st, src, lds, idx, scale, dst, ldd, rows, cols, accumulate);
  else if (src_dtype == GENRL_F32 && dst_dtype == GENRL_BF16) run_gather<float, __nv_bfloat16>(st, src, lds, idx, scale, dst, ldd, rows, cols, accumulate);
  else if (src_dtype == GENRL_BF16 && dst_dtype == GENRL_F32) run_gather<__nv_bfloat16, float>(st, src, lds, idx, scale, dst, ldd, rows, cols, accumulate);
  else run_gather<__nv_bfloat16, __nv_bfloat16>(st, src, lds, idx, scale, dst, ldd, rows, cols, accumulate);
  return check_launch("gather_rows");
}

// dst[r, 0:cols] = cast(src[r, 0:cols]) * row_scale[r]      dtypes: GENRL_F32 / GENRL_BF16 each side
extern "C" int genrl_copy2d(void* stream, const void* src, long long lds, int src_dtype, void* dst, long long ldd, int dst_dtype, long long rows,
                            int cols, const float* row_scale) {
  return genrl_gather_rows(stream, src, lds, src_dtype, nullptr, row_scale, dst, ldd, dst_dtype, rows, cols, 0);
}

extern "C" int genrl_split3_bf16(void* stream, const float* src, long long lds, void* hi, void* mid, void* lo, long long ldd, long long rows, int cols) {
  if (rows <= 0 || cols <= 0) return 0;
  const unsigned grid = (unsigned)std::min<long long>((rows * cols + 255) / 256, 148 * 32);
  genrl::launch_pdl(split3_bf16_kernel, grid, 256, 0, ST(stream), src, lds, (__nv_bfloat16*)hi, (__nv_bfloat16*)mid, (__nv_bfloat16*)lo, ldd, rows, cols);
  return check_launch("split3_bf16");
}

extern "C" int genrl_fill_normal(void* stream, float* out, long long n, unsigned long long seed, unsigned long long offset,
                                 const unsigned long long* step_dev) {
  if (n <= 0) return 0;
  unsigned grid = (unsigned)std::min<long long>(((n + 3) / 4 + 255) / 256, 148 * 16);
  genrl::launch_pdl(fill_normal_kernel, grid, 256, 0, ST(stream), out, n, seed, offset, step_dev);
  return check_launch("fill_normal");
}

extern "C" int genrl_counter_add(void* stream, unsigned long long* counter, unsigned long long inc) {
  genrl::launch_pdl(counter_add_kernel, 1, 1, 0, ST(stream), counter, inc);
  return check_launch("counter_add");
}
