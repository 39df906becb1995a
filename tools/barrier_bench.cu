// Micro-benchmark: cost of one grid-wide barrier among 128 co-resident CTAs (one per SM) on sm_100a, for the variants
// considered for the persistent RSSM scan kernel.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o barrier_bench barrier_bench.cu
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>

__device__ __forceinline__ unsigned ld_acquire(const unsigned* p) { unsigned v; asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ unsigned ld_relaxed(const unsigned* p) { unsigned v; asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ uint4 ld_relaxed4(const unsigned* p) { uint4 v; asm volatile("ld.relaxed.gpu.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ void st_release(unsigned* p, unsigned v) { asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ void red_release(unsigned* p, unsigned v) { asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ void fence_ar() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }

template <int V>
__device__ __forceinline__ void barrier(unsigned* ws, unsigned& epoch, int G) {
  __syncthreads();
  epoch += 1;
  const int tid = threadIdx.x;
  if (V == 0) {            // atomic counter, one poller, SC fences (first implementation)
    if (tid == 0) { __threadfence(); atomicAdd(ws, 1u); while (ld_acquire(ws) < epoch * G) {} __threadfence(); }
  } else if (V == 1) {     // flags 256 B apart, warp 0 polls 4 per lane
    if (tid == 0) st_release(ws + blockIdx.x * 64, epoch);
    if (tid < 32) {
      for (;;) { bool ok = true;
#pragma unroll
        for (int k = 0; k < 4; ++k) { int c = tid + 32 * k; if (c < G) ok = ok && ld_relaxed(ws + c * 64) >= epoch; }
        if (__all_sync(~0u, ok)) break; }
      fence_ar();
    }
  } else if (V == 2) {     // contiguous flags, warp 0 polls with one 16-byte load per lane
    if (tid == 0) st_release(ws + blockIdx.x, epoch);
    if (tid < 32) {
      for (;;) { uint4 f = ld_relaxed4(ws + tid * 4); int c = tid * 4;
        bool ok = (c >= G || f.x >= epoch) && (c + 1 >= G || f.y >= epoch) && (c + 2 >= G || f.z >= epoch) && (c + 3 >= G || f.w >= epoch);
        if (__all_sync(~0u, ok)) break; }
      fence_ar();
    }
  } else if (V == 3) {     // 8 counters 256 B apart (16 CTAs each), release reductions, lanes 0..7 poll
    if (tid == 0) red_release(ws + (blockIdx.x & 7) * 64, 1u);
    if (tid < 32) {
      for (;;) { bool ok = true; if (tid < 8) { unsigned n = (G - tid + 7) / 8; ok = ld_relaxed(ws + tid * 64) >= epoch * n; }
        if (__all_sync(~0u, ok)) break; }
      fence_ar();
    }
  } else if (V == 4) {     // single counter, release reduction + relaxed polling by one thread + acq_rel fence
    if (tid == 0) { red_release(ws, 1u); while (ld_relaxed(ws) < epoch * G) {} fence_ar(); }
  } else if (V == 5) {     // contiguous flags, 128 polling threads (one flag each), acq_rel fence each
    if (tid == 0) st_release(ws + blockIdx.x, epoch);
    if (tid < G) { while (ld_relaxed(ws + tid) < epoch) {} fence_ar(); }
  } else if (V == 6) {     // like 2 but the arriving thread is in warp 1, so warp 0 starts polling immediately
    if (tid == 32) st_release(ws + blockIdx.x, epoch);
    if (tid < 32) {
      for (;;) { uint4 f = ld_relaxed4(ws + tid * 4); int c = tid * 4;
        bool ok = (c >= G || f.x >= epoch) && (c + 1 >= G || f.y >= epoch) && (c + 2 >= G || f.z >= epoch) && (c + 3 >= G || f.w >= epoch);
        if (__all_sync(~0u, ok)) break; }
      fence_ar();
    }
  }
  __syncthreads();
}

template <int V>
__global__ void __launch_bounds__(512, 1) bench_kernel(unsigned* ws, float* scratch, int iters, int stores, int G) {
  extern __shared__ float smem[];
  unsigned epoch = 0;
  for (int it = 0; it < iters; ++it) {
    // mimic the scan step: a few global stores per thread that the barrier has to publish
    for (int s = 0; s < stores; ++s) scratch[((size_t)(it & 1) * 4 + s) * G * 512 + blockIdx.x * 512 + threadIdx.x] = (float)it;
    barrier<V>(ws, epoch, G);
  }
  if (threadIdx.x == 0) smem[0] = (float)epoch;
}

template <int V>
static void run(const char* name, unsigned* ws, float* scratch, int G, int stores) {
  const int iters = 2000;
  const size_t smem = 120 * 1024;
  cudaFuncSetAttribute(bench_kernel<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e9f;
  for (int rep = 0; rep < 5; ++rep) {
    cudaMemset(ws, 0, 64 * 1024);
    int it = iters;
    void* args[] = {&ws, &scratch, &it, &stores, &G};
    cudaEventRecord(e0);
    cudaError_t e = cudaLaunchCooperativeKernel((void*)bench_kernel<V>, dim3(G), dim3(512), args, smem, 0);
    cudaEventRecord(e1);
    if (e != cudaSuccess || cudaEventSynchronize(e1) != cudaSuccess) { printf("%s: launch failed: %s\n", name, cudaGetErrorString(cudaGetLastError())); return; }
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  printf("%-60s G=%d stores/thread=%d : %.3f us per barrier\n", name, G, stores, best * 1000.f / iters);
}

int main() {
  unsigned* ws; float* scratch;
  cudaMalloc(&ws, 64 * 1024);
  cudaMalloc(&scratch, (size_t)8 * 148 * 512 * 4);
  for (int stores = 0; stores <= 2; stores += 2) {
    run<0>("V0 atomic counter, 1 poller, SC fences", ws, scratch, 128, stores);
    run<4>("V4 red.release counter, 1 relaxed poller, acq_rel fence", ws, scratch, 128, stores);
    run<3>("V3 8 spread counters (red.release), 8 pollers", ws, scratch, 128, stores);
    run<1>("V1 flags 256 B apart, warp polls 4/lane", ws, scratch, 128, stores);
    run<2>("V2 contiguous flags, warp polls one 16 B load/lane", ws, scratch, 128, stores);
    run<6>("V6 = V2 with the arrival store issued from warp 1", ws, scratch, 128, stores);
    run<5>("V5 contiguous flags, 128 polling threads", ws, scratch, 128, stores);
  }
  return 0;
}
