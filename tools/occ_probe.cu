// Occupancy probe: how many CTAs of a given (threads, regs-bound, dynamic smem) shape does the runtime admit per SM, with and without the
// max-shared carve-out?  (Debugging aid for the 2-CTAs/SM assumptions of the GEMM / conv kernels.)
#include <cstdio>
#include <cuda_runtime.h>
template <int T, int MINB> __global__ void __launch_bounds__(T, MINB) k(float* p) { extern __shared__ float s[]; if (p) p[threadIdx.x] = s[threadIdx.x]; }
template <int T, int MINB> void probe(int smem) {
  auto kern = k<T, MINB>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  int a = -1, b = -1;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&a, kern, T, smem);
  cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, kern, T, smem);
  printf("threads %d minblocks %d smem %d: blocks/SM default carve-out %d, max-shared %d\n", T, MINB, smem, a, b);
}
int main() {
  cudaDeviceProp pr; cudaGetDeviceProperties(&pr, 0);
  printf("%s: smem/SM %zu, smem/block optin %zu, reserved/block %zu, regs/SM %d\n", pr.name, pr.sharedMemPerMultiprocessor, pr.sharedMemPerBlockOptin,
         pr.reservedSharedMemPerBlock, pr.regsPerMultiprocessor);
  for (int smem : {60000, 80000, 87296, 99584, 103168, 110000, 113000, 115000}) { probe<192, 2>(smem); probe<320, 2>(smem); probe<448, 2>(smem); }
  return 0;
}
