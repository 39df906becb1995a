// Shared helpers for the genrl_b200 CUDA library (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#define GENRL_F32 0
#define GENRL_BF16 1

#define GENRL_ACT_NONE 0
#define GENRL_ACT_SILU 1

namespace genrl {

// ---- error plumbing (C-ABI: int status + genrl_last_error()) ------------------------------
void set_error(const char* fmt, ...);
int check_launch(const char* what);

#define GENRL_REQUIRE(cond, ...)                 \
  do {                                           \
    if (!(cond)) {                               \
      genrl::set_error(__VA_ARGS__);             \
      return 1;                                  \
    }                                            \
  } while (0)

// ---- dtype helpers ---------------------------------------------------------------------------
template <typename T> __device__ __forceinline__ float to_f32(T v);
template <> __device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

__device__ __forceinline__ float ld_any(const void* p, int dtype, size_t i) {
  return dtype == GENRL_F32 ? reinterpret_cast<const float*>(p)[i]
                            : __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p)[i]);
}
__device__ __forceinline__ void st_any(void* p, int dtype, size_t i, float v) {
  if (dtype == GENRL_F32) reinterpret_cast<float*>(p)[i] = v;
  else reinterpret_cast<__nv_bfloat16*>(p)[i] = __float2bfloat16_rn(v);
}

// ---- warp / block reductions -----------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }
__device__ __forceinline__ float siluf_(float x) { return x * sigmoidf_(x); }
// d/dx silu(x) = s + x s (1 - s)
__device__ __forceinline__ float silu_gradf_(float x) {
  float s = sigmoidf_(x);
  return s * (1.0f + x * (1.0f - s));
}

// bf16 mode uses the SFU fast paths (ex2/rcp approximations, ~2 ulp): activations are rounded to bf16 right after anyway.
// One MUFU per activation: sigmoid(x) = 0.5 + 0.5 tanh(x/2) with tanh.approx.f32 (rel. error 2^-11, abs. error of the sigmoid <= 2.5e-4,
// below the bf16 rounding of the O(1) activations stored next).  The IEEE forms (__frcp_rn: slow-path call per element; ex2 + rcp: two
// MUFU + two FP ops) made the row kernels instruction-issue / MUFU bound (profiles/r01_notes.md §2).
__device__ __forceinline__ float tanh_fast(float x) { float y; asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
template <bool FAST> __device__ __forceinline__ float sigm_t(float x) {
  if (FAST) return fmaf(0.5f, tanh_fast(0.5f * x), 0.5f);
  return 1.0f / (1.0f + expf(-x));
}
template <bool FAST> __device__ __forceinline__ float silu_t(float x) {
  if (FAST) { const float h = 0.5f * x; return fmaf(h, tanh_fast(h), h); }
  return x * sigm_t<FAST>(x);
}
template <bool FAST> __device__ __forceinline__ float silu_grad_t(float x) {
  float s = sigm_t<FAST>(x);
  return s * (1.0f + x * (1.0f - s));
}
template <bool FAST> __device__ __forceinline__ float tanh_t(float x) {
  if (FAST) return tanh_fast(x);
  return tanhf(x);
}

// Philox4x32-10 counter-based generator (fast-mode noise; parity mode takes explicit noise pointers).
__device__ __forceinline__ uint4 philox4x32(uint4 ctr, uint2 key) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
    uint32_t hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += W0; key.y += W1;
  }
  return ctr;
}
__device__ __forceinline__ float u32_to_unit_open(uint32_t x) {   // (0,1]
  return (float)(x >> 8) * (1.0f / 16777216.0f) + (1.0f / 16777216.0f);
}

inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

// ---- programmatic dependent launch (PDL) ------------------------------------------------------------------------------------------
// One update is ~1 600 launches of mostly 4-20 us kernels replayed from a CUDA graph: the drain -> launch -> ramp bubble between two
// dependent kernels is a sizeable share of the step.  Every kernel of the library is launched with
// cudaLaunchAttributeProgrammaticStreamSerialization and calls pdl_sync() before its first access to global memory:
//   griddepcontrol.wait               blocks until the preceding grid has completed and its memory operations are visible
//   griddepcontrol.launch_dependents  lets the NEXT grid's CTAs become resident (they park in their own wait)
// so a kernel's launch latency, CTA scheduling and (for the tcgen05 kernels) barrier init / TMEM allocation / descriptor prefetch
// overlap the tail of its predecessor.  The trigger comes after the wait, so at most one successor is parked at any time.
// A predecessor that is not ours (an ATen kernel) never triggers early: the wait then returns at once, ordinary stream semantics.
__device__ __forceinline__ void pdl_sync() {
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

bool pdl_enabled();      // GENRL_PDL=0 turns the launch attribute off (runtime.cu)

template <typename... KArgs, typename... Args>
inline void launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
  static_assert(sizeof...(KArgs) == sizeof...(Args), "kernel argument count");
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = pdl_enabled() ? 1 : 0;
  cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

}  // namespace genrl
