// Optimizer step on flat parameter buffers + compute-layout weight shadows + RewardEMA quantiles.
//   reference: Optimizer.__call__ dreamer_utils.py:892-932 (clip_grad_norm_ -> multiplicative weight decay -> Adam),
//              RewardEMA dreamer_utils.py:1014-1029 (5%/95% torch.quantile, EMA), update_slow_target dreamer.py:455-462.
#include "common.cuh"

namespace genrl {

__global__ void __launch_bounds__(256) sumsq_kernel(const float* __restrict__ g, long long n, float* __restrict__ out) {
  genrl::pdl_sync();
  float s = 0.f;
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256) { float v = g[i]; s += v * v; }
  __shared__ float red[8];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) { float t = 0.f; for (int i = 0; i < 8; ++i) t += red[i]; atomicAdd(out, t); }
}

__global__ void __launch_bounds__(256) scale_kernel(float* __restrict__ p, long long n, float s) {
  genrl::pdl_sync();
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256) p[i] *= s;
}

// state[0] = step count (as float), state[1] = grad norm of the last call (metric)
__global__ void __launch_bounds__(256) adam_kernel(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                                                   long long n, float lr, float eps, float clip, const float* __restrict__ sumsq,
                                                   const float* __restrict__ state, float decay) {
  genrl::pdl_sync();
  const float norm = sqrtf(*sumsq);
  const float coef = clip > 0.f ? fminf(clip / (norm + 1e-6f), 1.0f) : 1.0f;
  const float t = state[0] + 1.0f;
  const float bc1 = 1.0f - powf(0.9f, t), bc2 = 1.0f - powf(0.999f, t);
  const float step_size = lr / bc1, inv_bc2_sqrt = 1.0f / sqrtf(bc2);
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256) {
    float gi = g[i] * coef;
    float mi = m[i] + (gi - m[i]) * (1.0f - 0.9f);            // exp_avg.lerp_(grad, 1-beta1)
    float vi = v[i] * 0.999f + gi * gi * (1.0f - 0.999f);
    m[i] = mi; v[i] = vi;
    float denom = sqrtf(vi) * inv_bc2_sqrt + eps;
    const float pi = decay != 1.0f ? __fmul_rn(p[i], decay) : p[i];     // multiplicative weight decay first (dreamer_utils.py:926-932), own rounding
    p[i] = pi - step_size * (mi / denom);
    g[i] = 0.f;
  }
}

__global__ void adam_finish_kernel(float* state, const float* sumsq) {
  genrl::pdl_sync();
  state[0] += 1.0f;
  state[1] = sqrtf(*sumsq);
}

// -------------------------------------------------------------------------------------------------
// multi-tensor weight packer: one launch refreshes every compute-layout shadow after an optimizer step
// -------------------------------------------------------------------------------------------------
struct PackEntry {
  const float* src;
  void* dst;
  long long n;       // elements in src
  int kind;          // 0: cast copy, 1: [A,B,P] -> [A,P,B], 2: rows x cols -> rows x ld (zero padded)
  int B, P;          // kind 1
  int cols, ld;      // kind 2
};

template <typename TO>
__global__ void __launch_bounds__(256) pack_kernel(const PackEntry* __restrict__ table) {
  genrl::pdl_sync();
  const PackEntry e = table[blockIdx.y];
  TO* dst = reinterpret_cast<TO*>(e.dst);
  if (e.kind == 0) {
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < e.n; i += (long long)gridDim.x * 256) dst[i] = from_f32<TO>(e.src[i]);
  } else if (e.kind == 1) {
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < e.n; i += (long long)gridDim.x * 256) {
      int b = (int)(i % e.B); int pp = (int)((i / e.B) % e.P); long long a = i / ((long long)e.B * e.P);
      dst[i] = from_f32<TO>(e.src[(a * e.B + b) * e.P + pp]);
    }
  } else {
    const long long rows = e.n / e.cols;
    const long long total = rows * e.ld;
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
      long long r = i / e.ld; int c = (int)(i % e.ld);
      dst[i] = from_f32<TO>(c < e.cols ? e.src[r * e.cols + c] : 0.f);
    }
  }
}

// -------------------------------------------------------------------------------------------------
// RewardEMA: exact k-th order statistics by 4-pass radix select on order-preserving uint32 keys, 4 ranks at once
// -------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t f2key(float f) { uint32_t u = __float_as_uint(f); return (u & 0x80000000u) ? ~u : (u | 0x80000000u); }
__device__ __forceinline__ float key2f(uint32_t k) { uint32_t u = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k; return __uint_as_float(u); }

// ema[0:2] <- alpha*q + (1-alpha)*ema ; out[0]=offset(ema[0]) out[1]=scale=max(ema[1]-ema[0],1) out[2..3] = raw quantiles
__global__ void __launch_bounds__(1024) quantile_ema_kernel(const float* __restrict__ x, long long n, float q_lo, float q_hi, float alpha,
                                                            float* __restrict__ ema, float* __restrict__ out) {
  genrl::pdl_sync();
  __shared__ unsigned hist[4][256];
  __shared__ uint32_t prefix[4];
  __shared__ long long rank[4];
  __shared__ float vals[4];
  const int tid = threadIdx.x;
  // torch.quantile 'linear': pos = q*(n-1); lerp(x[floor], x[ceil], frac)
  const float pos_lo = q_lo * (float)(n - 1), pos_hi = q_hi * (float)(n - 1);
  if (tid == 0) {
    rank[0] = (long long)floorf(pos_lo); rank[1] = (long long)ceilf(pos_lo);
    rank[2] = (long long)floorf(pos_hi); rank[3] = (long long)ceilf(pos_hi);
    for (int r = 0; r < 4; ++r) prefix[r] = 0;
  }
  __syncthreads();
  for (int pass = 0; pass < 4; ++pass) {
    const int shift = 24 - 8 * pass;
    const uint32_t mask = pass == 0 ? 0u : (0xffffffffu << (shift + 8));
    for (int i = tid; i < 4 * 256; i += 1024) (&hist[0][0])[i] = 0;
    __syncthreads();
    for (long long i = tid; i < n; i += 1024) {
      uint32_t k = f2key(x[i]);
#pragma unroll
      for (int r = 0; r < 4; ++r) if ((k & mask) == prefix[r]) atomicAdd(&hist[r][(k >> shift) & 255], 1u);
    }
    __syncthreads();
    if (tid < 4) {
      long long rk = rank[tid];
      unsigned cum = 0; int b = 0;
      for (; b < 256; ++b) { if (rk < (long long)cum + hist[tid][b]) break; cum += hist[tid][b]; }
      rank[tid] = rk - cum;
      prefix[tid] |= (uint32_t)b << shift;
    }
    __syncthreads();
  }
  if (tid < 4) vals[tid] = key2f(prefix[tid]);
  __syncthreads();
  if (tid == 0) {
    float fl = pos_lo - floorf(pos_lo), fh = pos_hi - floorf(pos_hi);
    float qlo = vals[0] + (vals[1] - vals[0]) * fl;
    float qhi = vals[2] + (vals[3] - vals[2]) * fh;
    float e0 = alpha * qlo + (1.0f - alpha) * ema[0];
    float e1 = alpha * qhi + (1.0f - alpha) * ema[1];
    ema[0] = e0; ema[1] = e1;
    out[0] = e0; out[1] = fmaxf(e1 - e0, 1.0f); out[2] = qlo; out[3] = qhi;
  }
}

// actor objective: loss = -mean_{t=1..H-1,n}((R_t - off)/scale) - ent_scale * mean_{t<H-1}(ent)  (dreamer.py:392-429, weights == 1)
// out[0]=loss, out[1]=mean normed target (all t), out[2]=std normed target, out[3]=mean ent ; dcoef[0] = -1/((H-1) N scale)
__global__ void __launch_bounds__(1024) actor_objective_kernel(const float* __restrict__ target, const float* __restrict__ ent, const float* __restrict__ offscale,
                                                               float ent_scale, int H, long long N, float* __restrict__ out, float* __restrict__ dcoef) {
  genrl::pdl_sync();
  __shared__ double red[3][32];
  const float off = offscale[0], sc = offscale[1];
  double s_all = 0.0, s_obj = 0.0, s_ent = 0.0;
  const long long tot = (long long)H * N;
  for (long long i = threadIdx.x; i < tot; i += 1024) {
    float nt = (target[i] - off) / sc;
    s_all += nt;
    if (i >= N) s_obj += nt;
    if (i < (long long)(H - 1) * N) s_ent += ent[i];
  }
  for (int o = 16; o > 0; o >>= 1) {
    s_all += __shfl_xor_sync(0xffffffffu, s_all, o); s_obj += __shfl_xor_sync(0xffffffffu, s_obj, o); s_ent += __shfl_xor_sync(0xffffffffu, s_ent, o);
  }
  if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = s_all; red[1][threadIdx.x >> 5] = s_obj; red[2][threadIdx.x >> 5] = s_ent; }
  __syncthreads();
  __shared__ float mean_all;
  if (threadIdx.x == 0) {
    double a = 0, b = 0, c = 0;
    for (int i = 0; i < 32; ++i) { a += red[0][i]; b += red[1][i]; c += red[2][i]; }
    const double cnt = (double)(H - 1) * N;
    out[0] = (float)(-(b / cnt) - ent_scale * (c / cnt));
    out[1] = (float)(a / tot);
    out[3] = (float)(c / cnt);
    mean_all = (float)(a / tot);
    dcoef[0] = (float)(-1.0 / (cnt * sc));
  }
  __syncthreads();
  double v = 0.0;
  for (long long i = threadIdx.x; i < tot; i += 1024) { float d = (target[i] - off) / sc - mean_all; v += (double)d * d; }
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[0][threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x == 0) { double a = 0; for (int i = 0; i < 32; ++i) a += red[0][i]; out[2] = tot > 1 ? (float)sqrt(a / (tot - 1)) : 0.f; }
}

// out[0] (+)= scale * sum(x) / n    (weighted means of per-row log-likelihoods -> scalar losses)
__global__ void __launch_bounds__(1024) mean_scale_kernel(const float* __restrict__ x, long long n, float scale, float* __restrict__ out, int accumulate) {
  genrl::pdl_sync();
  __shared__ double red[32];
  double s = 0.0;
  for (long long i = threadIdx.x; i < n; i += 1024) s += x[i];
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) { double t = 0; for (int i = 0; i < 32; ++i) t += red[i]; float r = (float)(t / n) * scale; out[0] = accumulate ? out[0] + r : r; }
}

static inline unsigned grid_for(long long total, int cap = 148 * 8) {
  long long g = (total + 255) / 256;
  return (unsigned)(g > cap ? cap : (g < 1 ? 1 : g));
}

}  // namespace genrl

using namespace genrl;
#define ST(s) reinterpret_cast<cudaStream_t>(s)

// sumsq_out[0] (+)= sum g^2
extern "C" int genrl_grad_sumsq(void* stream, const float* g, long long n, float* sumsq_out, int zero_first) {
  if (zero_first) cudaMemsetAsync(sumsq_out, 0, sizeof(float), ST(stream));
  if (n > 0) genrl::launch_pdl(sumsq_kernel, grid_for(n), 256, 0, ST(stream), g, n, sumsq_out);
  return check_launch("grad_sumsq");
}

extern "C" int genrl_scale(void* stream, float* p, long long n, float s) {
  if (n <= 0) return 0;
  genrl::launch_pdl(scale_kernel, grid_for(n), 256, 0, ST(stream), p, n, s);
  return check_launch("scale");
}

// p,g,m,v: flat fp32 [n]; sumsq: device scalar from genrl_grad_sumsq (norm over the whole optimizer call);
// state: device float[2] {step, last grad norm}; `finish` bumps the step (set on the last segment of a call).
extern "C" int genrl_adam_step(void* stream, float* p, float* g, float* m, float* v, long long n, float lr, float eps, float clip,
                               const float* sumsq, float* state, int finish, float decay) {
  if (n > 0) genrl::launch_pdl(adam_kernel, grid_for(n), 256, 0, ST(stream), p, g, m, v, n, lr, eps, clip, sumsq, state, decay);
  if (finish) genrl::launch_pdl(adam_finish_kernel, 1, 1, 0, ST(stream), state, sumsq);
  return check_launch("adam_step");
}

// table: device array of PackEntry (see genrl_b200.h for the field layout), n_entries launches in one grid
extern "C" int genrl_pack_weights(void* stream, const void* table, int n_entries, int dst_dtype) {
  if (n_entries <= 0) return 0;
  dim3 grid(148 * 2, n_entries);
  if (dst_dtype == GENRL_F32) genrl::launch_pdl(pack_kernel<float>, grid, 256, 0, ST(stream), (const PackEntry*)table);
  else genrl::launch_pdl(pack_kernel<__nv_bfloat16>, grid, 256, 0, ST(stream), (const PackEntry*)table);
  return check_launch("pack_weights");
}
extern "C" int genrl_pack_entry_size() { return (int)sizeof(PackEntry); }

extern "C" int genrl_quantile_ema(void* stream, const float* x, long long n, float q_lo, float q_hi, float alpha, float* ema, float* out4) {
  GENRL_REQUIRE(n > 0, "quantile_ema: empty input");
  genrl::launch_pdl(quantile_ema_kernel, 1, 1024, 0, ST(stream), x, n, q_lo, q_hi, alpha, ema, out4);
  return check_launch("quantile_ema");
}

extern "C" int genrl_actor_objective(void* stream, const float* target, const float* ent, const float* offscale, float ent_scale, int H, long long N,
                                     float* out4, float* dcoef) {
  genrl::launch_pdl(actor_objective_kernel, 1, 1024, 0, ST(stream), target, ent, offscale, ent_scale, H, N, out4, dcoef);
  return check_launch("actor_objective");
}

extern "C" int genrl_mean_scale(void* stream, const float* x, long long n, float scale, float* out, int accumulate) {
  GENRL_REQUIRE(n > 0, "mean_scale: empty input");
  genrl::launch_pdl(mean_scale_kernel, 1, 1024, 0, ST(stream), x, n, scale, out, accumulate);
  return check_launch("mean_scale");
}
