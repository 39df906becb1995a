// Video-text reward of the imagination behaviour, fused:
//   reference tools/genrl_utils.py:250-277 compute_reward(score_fn = 'max_cosine') and :279-370 video_text_reward with align_sequence.
//
// Inputs are the decoder-input projections (decoder._conv_in[0], a GEMM upstream) of the agent's imagined latents, pa[t, n, :], and of the
// cached target sequence, pt[t, n, :], both [T, N, E] time-major.
//   maxcos(u, v) = <u, v> / max(|u|, |v|)^2                                                              (genrl_utils.py:240-243)
//   align_sequence: score[o, n] = mean_{j < F} maxcos(pa[o + j, n], pt[j, n]) for o in [0, T - F);  b_n = argmax_o score (first maximum);
//                   the target is delayed by b_n frames: idx[t, n] = max(t - b_n, 0)                    (genrl_utils.py:344-366)
//   reward[t, n] = maxcos(pa[t, n], pt[idx[t, n], n])
// One CTA per rollout row n: its 2T vectors (E = 1536: 2T x 3 KB in bf16) are read once for the norms and then come from L1/L2 for the
// (T - F) F alignment dot products and the T reward dot products.  HBM-bound: 2 T N E s bytes in, T N 4 bytes out.
#include "common.cuh"

namespace genrl {

#define ST(s) reinterpret_cast<cudaStream_t>(s)
constexpr int VR_THREADS = 256;
constexpr int VR_MAX_T = 64;

template <typename T>
__device__ __forceinline__ void warp_dot3(const T* __restrict__ u, const T* __restrict__ v, int E, float& uv, float& uu, float& vv) {
  const int lane = threadIdx.x & 31;
  float a = 0.f, b = 0.f, c = 0.f;
  for (int i = lane; i < E; i += 32) {
    const float x = to_f32<T>(u[i]), y = to_f32<T>(v[i]);
    a = fmaf(x, y, a); b = fmaf(x, x, b); c = fmaf(y, y, c);
  }
  uv = warp_sum(a); uu = warp_sum(b); vv = warp_sum(c);
}

__device__ __forceinline__ float maxcos_from(float uv, float uu, float vv) {
  // torch: max_norm = max(|u|, |v|); sum((u / max_norm) * (v / max_norm))
  const float m = fmaxf(sqrtf(uu), sqrtf(vv));
  return uv / (m * m);
}

template <typename T>
__global__ void __launch_bounds__(VR_THREADS) vlm_reward_fwd_kernel(const T* __restrict__ pa, const T* __restrict__ pt, float* __restrict__ reward,
                                                                    int* __restrict__ shift, int Tn, long long N, int E, int F, int align) {
  genrl::pdl_sync();
  const long long n = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = VR_THREADS / 32;
  __shared__ float score[VR_MAX_T];
  __shared__ int best;
  const int offs = Tn - F;
  if (align) {
    for (int o = warp; o < offs; o += nw) {
      float s = 0.f;
      for (int j = 0; j < F; ++j) {
        float uv, uu, vv;
        warp_dot3<T>(pt + ((long long)j * N + n) * E, pa + ((long long)(o + j) * N + n) * E, E, uv, uu, vv);
        s += maxcos_from(uv, uu, vv);
      }
      if (lane == 0) score[o] = s / (float)F;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      int b = 0;
      float m = score[0];
      for (int o = 1; o < offs; ++o)
        if (score[o] > m) { m = score[o]; b = o; }       // first maximum wins (torch.argmax)
      best = b;
      shift[n] = b;
    }
  } else if (threadIdx.x == 0) {
    best = 0;
    shift[n] = 0;
  }
  __syncthreads();
  const int b = best;
  for (int t = warp; t < Tn; t += nw) {
    const int src = max(t - b, 0);
    float uv, uu, vv;
    warp_dot3<T>(pt + ((long long)src * N + n) * E, pa + ((long long)t * N + n) * E, E, uv, uu, vv);
    if (lane == 0) reward[(long long)t * N + n] = maxcos_from(uv, uu, vv);
  }
}

// d reward / d pa (the target is a constant): with u = pt[idx], v = pa[t], m = max(|u|, |v|):
//   |v| >= |u|:  r = <u,v>/|v|^2   ->  dr/dv = u/|v|^2 - 2 <u,v> v/|v|^4
//   |v| <  |u|:  r = <u,v>/|u|^2   ->  dr/dv = u/|u|^2
template <typename T>
__global__ void __launch_bounds__(VR_THREADS) vlm_reward_bwd_kernel(const T* __restrict__ pa, const T* __restrict__ pt, const int* __restrict__ shift,
                                                                    const float* __restrict__ dreward, T* __restrict__ dpa, int Tn, long long N, int E) {
  genrl::pdl_sync();
  const long long n = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = VR_THREADS / 32;
  const int b = shift[n];
  for (int t = warp; t < Tn; t += nw) {
    const int src = max(t - b, 0);
    const T* u = pt + ((long long)src * N + n) * E;
    const T* v = pa + ((long long)t * N + n) * E;
    T* d = dpa + ((long long)t * N + n) * E;
    float uv, uu, vv;
    warp_dot3<T>(u, v, E, uv, uu, vv);
    const float g = dreward[(long long)t * N + n];
    float cu, cv;
    if (vv >= uu) { cu = g / vv; cv = -2.f * g * uv / (vv * vv); }
    else { cu = g / uu; cv = 0.f; }
    for (int i = lane; i < E; i += 32) d[i] = from_f32<T>(fmaf(cu, to_f32<T>(u[i]), cv * to_f32<T>(v[i])));
  }
}

}  // namespace genrl

using namespace genrl;

extern "C" int genrl_vlm_reward_fwd(void* stream, const void* pa, const void* pt, float* reward, int* shift, int T, long long N, int E, int n_frames,
                                    int align, int dtype) {
  if (N <= 0 || T <= 0) return 0;
  GENRL_REQUIRE(T <= VR_MAX_T, "vlm_reward: T = %d exceeds %d", T, VR_MAX_T);
  GENRL_REQUIRE(!align || T > n_frames, "vlm_reward: align_sequence needs more than n_frames = %d steps, got %d", n_frames, T);
  if (dtype == GENRL_F32)
    genrl::launch_pdl(vlm_reward_fwd_kernel<float>, (unsigned)N, VR_THREADS, 0, ST(stream), (const float*)pa, (const float*)pt, reward, shift, T, N, E, n_frames, align);
  else
    genrl::launch_pdl(vlm_reward_fwd_kernel<__nv_bfloat16>, (unsigned)N, VR_THREADS, 0, ST(stream), (const __nv_bfloat16*)pa, (const __nv_bfloat16*)pt, reward, shift,
                                                                                    T, N, E, n_frames, align);
  return check_launch("vlm_reward_fwd");
}

extern "C" int genrl_vlm_reward_bwd(void* stream, const void* pa, const void* pt, const int* shift, const float* dreward, void* dpa, int T, long long N,
                                    int E, int dtype) {
  if (N <= 0 || T <= 0) return 0;
  if (dtype == GENRL_F32)
    genrl::launch_pdl(vlm_reward_bwd_kernel<float>, (unsigned)N, VR_THREADS, 0, ST(stream), (const float*)pa, (const float*)pt, shift, dreward, (float*)dpa, T, N, E);
  else
    genrl::launch_pdl(vlm_reward_bwd_kernel<__nv_bfloat16>, (unsigned)N, VR_THREADS, 0, ST(stream), (const __nv_bfloat16*)pa, (const __nv_bfloat16*)pt, shift, dreward,
                                                                                    (__nv_bfloat16*)dpa, T, N, E);
  return check_launch("vlm_reward_bwd");
}
