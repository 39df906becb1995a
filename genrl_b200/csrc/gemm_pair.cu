// Persistent 2-CTA ("pair") tcgen05 GEMM: C[M,N] = A[M,K] op(W)^T + bias, bf16 operands, fp32 accumulation in TMEM.
//
// Why: round 2 measured the library bar for the step's GEMM shapes (tools/gemm_bar.py, profiles/r02_gemm_bar.txt): cuBLASLt runs
// 38400x1024x1024 at 1.23 PFLOP/s and 2400x1024x1024 at 0.80 where the single-CTA kernels of gemm_sm100.cu reach 0.93 / 0.50.  Those
// kernels are bound by operand bytes through the L2->SM fabric (~6 300 B/clk chip-wide): 128x128 tiles move 0.0156 B/FLOP, the
// 256x128 persistent kernel 0.0117.  With tcgen05.mma.cta_group::2 a PAIR of CTAs on two SMs computes one 256 x 256 tile: each CTA
// loads its own 128 rows of A and HALF of the B tile (128 of the 256 weight rows), the tensor cores of both SMs read both halves —
// 32 KB per CTA per 128x256x64 MACs = 0.0078 B/FLOP, which moves the kernel from the fabric limit towards the tensor pipe.
//
// Structure (cluster of 2, CTA rank 0 = leader):
//   warp 0 (both CTAs)  TMA producer: waits the CTA's own empty[stage]; the leader arms full[stage] in ITS shared memory with the bytes of
//                       BOTH CTAs; both issue cp.async.bulk.tensor ... .cta_group::2 loads that complete on the leader's barrier
//   warp 1 (leader)     one thread issues tcgen05.mma.cta_group::2 (UMMA 256 x 256 x 16); tcgen05.commit ... .multicast::cluster frees the
//                       stage in both CTAs and, after the last k-block, raises tmem_full in both CTAs
//   warps 2..9 (both)   epilogue: each CTA drains its 128 accumulator rows (2 x 256 TMEM columns, double-buffered against the next tile's
//                       main loop), one arrival per warp on the leader's tmem_empty barrier
// Persistent over 256x256 tiles, one pair per 2 SMs (74 pairs on B200).
#include "tc_ptx.cuh"
#include <algorithm>
#include <cstdlib>

namespace genrl {

constexpr int PR_BM = 128;           // rows per CTA (256 per pair)
constexpr int PR_BN = 256;           // columns per pair tile
constexpr int PR_BK = 64;
constexpr int PR_STAGES = 6;
constexpr int PR_THREADS = 320;
constexpr int PR_A_BYTES = PR_BM * PR_BK * 2;            // 16 KB
constexpr int PR_B_BYTES = (PR_BN / 2) * PR_BK * 2;      // 16 KB: this CTA's half of the B tile
constexpr int PR_STAGE_BYTES = PR_A_BYTES + PR_B_BYTES;
constexpr int PR_SMEM = PR_STAGES * PR_STAGE_BYTES + 1024 + 512;

struct PairParams {
  int M, N, K, num_kb;
  int splits, kb_per_split;      // split-K: work item w = tile * splits + split covers k-blocks [split * kb_per_split, ...); splits > 1 -> atomic epilogue
  int atomic;                    // fp32 atomicAdd epilogue (C += result: the wgrad GEMMs accumulate into the flat gradient buffers)
  int ldc, c_dtype;
  const float* bias;
  void* C;
};

__device__ __forceinline__ uint32_t pr_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void pr_cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t pr_mapa(uint32_t local, uint32_t cta) {
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local), "r"(cta));
  return remote;
}
// TMA load whose completion bytes go to an mbarrier given as a shared::cluster address (the pair leader's barrier)
__device__ __forceinline__ void pr_tma_load_2d(const CUtensorMap* map, uint32_t bar_cluster_addr, void* dst, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void pr_umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum) : "memory");
}
// arrives (count 1) on the barrier at the same shared-memory offset in every CTA of the mask once the MMAs issued so far have completed
__device__ __forceinline__ void pr_umma_commit_mc(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"(mask) : "memory");
}
__device__ __forceinline__ void pr_mbar_arrive_cluster(uint32_t bar_cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(bar_cluster_addr) : "memory");
}

template <bool A_MN, bool B_MN>
__global__ void __launch_bounds__(PR_THREADS, 1)
gemm_pair_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const PairParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + PR_STAGES * PR_STAGE_BYTES);       // used in the leader
  uint64_t* empty_bar = full_bar + PR_STAGES;                                                // per CTA
  uint64_t* tfull_bar = empty_bar + PR_STAGES;                                               // [2] per CTA
  uint64_t* tempty_bar = tfull_bar + 2;                                                      // [2] used in the leader
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = pr_ctarank();
  const bool leader = rank == 0;
  const int pair = blockIdx.x >> 1;
  const int npairs = gridDim.x >> 1;
  const int nkb = p.num_kb;
  const int tiles_n = (p.N + PR_BN - 1) / PR_BN;
  const int tiles = ((p.M + 2 * PR_BM - 1) / (2 * PR_BM)) * tiles_n;
  const int items = tiles * p.splits;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < PR_STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&tfull_bar[s], 1); mbar_init(&tempty_bar[s], 16); }      // 8 epilogue warps x 2 CTAs
    fence_barrier_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  genrl::pdl_sync();
  pr_cluster_sync();                    // both CTAs' barriers are initialised before any remote arrival / completion

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int w = pair; w < items; w += npairs) {
        const int t = w / p.splits, sp = w - t * p.splits;
        const int m0 = (t / tiles_n) * (2 * PR_BM) + (int)rank * PR_BM;
        const int n0 = (t % tiles_n) * PR_BN + (int)rank * (PR_BN / 2);
        const int kb0 = sp * p.kb_per_split, kb1 = min(nkb, kb0 + p.kb_per_split);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait_sleep(&empty_bar[stage], phase ^ 1, 32);
          uint8_t* sa = smem + stage * PR_STAGE_BYTES;
          uint8_t* sb = sa + PR_A_BYTES;
          if (leader) mbar_expect_tx(&full_bar[stage], 2 * PR_STAGE_BYTES);
          const uint32_t fb = pr_mapa(smem_u32(&full_bar[stage]), 0);
          if (!A_MN) {
            pr_tma_load_2d(&tmA, fb, sa, kb * PR_BK, m0);                                     // box {64 k, 128 rows}
          } else {
#pragma unroll
            for (int j = 0; j < PR_BM / 64; ++j)                                              // box {64 m, 64 k}
              pr_tma_load_2d(&tmA, fb, sa + j * (64 * PR_BK * 2), m0 + 64 * j, kb * PR_BK);
          }
          if (!B_MN) {
            pr_tma_load_2d(&tmB, fb, sb, kb * PR_BK, n0);                                     // box {64 k, 128 weight rows}
          } else {
#pragma unroll
            for (int j = 0; j < (PR_BN / 2) / 64; ++j)                                        // box {64 n, 64 k}
              pr_tma_load_2d(&tmB, fb, sb + j * (64 * PR_BK * 2), n0 + 64 * j, kb * PR_BK);
          }
          if (++stage == PR_STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (leader && lane == 0) {
      // M = 256 (pair), N = 256: c = F32, a = b = BF16, B major, N >> 3, M >> 4
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((A_MN ? 1u : 0u) << 15) | ((B_MN ? 1u : 0u) << 16) | ((uint32_t)(PR_BN >> 3) << 17) |
                             ((uint32_t)(256 >> 4) << 24);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int w = pair; w < items; w += npairs, ++it) {
        const int sp = w % p.splits;
        const int kb0 = sp * p.kb_per_split, kb1 = min(nkb, kb0 + p.kb_per_split);
        const int buf = it & 1;
        mbar_wait_sleep(&tempty_bar[buf], ((it >> 1) & 1) ^ 1, 32);
        tc_fence_after();
        const uint32_t acc = tmem_base + (uint32_t)(buf * PR_BN);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * PR_STAGE_BYTES);
          const uint32_t sb = sa + PR_A_BYTES;
#pragma unroll
          for (int k = 0; k < PR_BK / 16; ++k) {
            const uint64_t ad = A_MN ? make_smem_desc(sa + k * 2048, 64 * PR_BK * 2, 1024) : make_smem_desc(sa + k * 32, 16, 1024);
            const uint64_t bd = B_MN ? make_smem_desc(sb + k * 2048, 64 * PR_BK * 2, 1024) : make_smem_desc(sb + k * 32, 16, 1024);
            pr_umma_bf16(acc, ad, bd, idesc, (kb > kb0 || k != 0) ? 1u : 0u);
          }
          pr_umma_commit_mc(&empty_bar[stage], 3);
          if (kb == kb1 - 1) pr_umma_commit_mc(&tfull_bar[buf], 3);
          if (++stage == PR_STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else {
    // ------------------------------ epilogue (warps 2..9 of both CTAs) ------------------------------
    const int quarter = warp & 3;
    const int half = warp >= 6 ? 1 : 0;                   // column half of the 256-wide tile
    const uint32_t tempty_leader0 = pr_mapa(smem_u32(&tempty_bar[0]), 0);
    const uint32_t tempty_leader1 = pr_mapa(smem_u32(&tempty_bar[1]), 0);
    int it = 0;
    for (int w = pair; w < items; w += npairs, ++it) {
      const int t = w / p.splits, sp = w - t * p.splits;
      const int buf = it & 1;
      const int m0 = (t / tiles_n) * (2 * PR_BM) + (int)rank * PR_BM;
      const int n0 = (t % tiles_n) * PR_BN;
      mbar_wait_sleep(&tfull_bar[buf], (it >> 1) & 1, 64);
      tc_fence_after();
      const uint32_t acc = tmem_base + (uint32_t)(buf * PR_BN + half * 128) + ((uint32_t)(quarter * 32) << 16);
      const int row = m0 + quarter * 32 + lane;
      const bool row_ok = row < p.M;
#pragma unroll 1
      for (int c = 0; c < 128; c += 32) {
        const int col0 = n0 + half * 128 + c;
        if (col0 >= p.N) break;
        uint32_t v[32];
        tmem_ld32(acc + (uint32_t)c, v);
        if (!row_ok) continue;
        const int ncols = min(32, p.N - col0);
        float f[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
        if (p.bias != nullptr && sp == 0) {
#pragma unroll
          for (int j = 0; j < 32; ++j) if (j < ncols) f[j] += __ldg(p.bias + col0 + j);
        }
        if (p.atomic) {
          float* out = reinterpret_cast<float*>(p.C) + (size_t)row * p.ldc + col0;
#pragma unroll
          for (int j = 0; j < 32; ++j) if (j < ncols) atomicAdd(out + j, f[j]);
        } else if (p.c_dtype == GENRL_F32) {
          float* out = reinterpret_cast<float*>(p.C) + (size_t)row * p.ldc + col0;
          if (ncols == 32 && ((reinterpret_cast<uintptr_t>(out) & 15) == 0)) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(out + j) = make_float4(f[j], f[j + 1], f[j + 2], f[j + 3]);
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) if (j < ncols) out[j] = f[j];
          }
        } else {
          __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(p.C) + (size_t)row * p.ldc + col0;
          if (ncols == 32 && ((reinterpret_cast<uintptr_t>(out) & 15) == 0)) {
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
              uint4 pk;
              __nv_bfloat162 a0 = __floats2bfloat162_rn(f[j], f[j + 1]), a1 = __floats2bfloat162_rn(f[j + 2], f[j + 3]);
              __nv_bfloat162 a2 = __floats2bfloat162_rn(f[j + 4], f[j + 5]), a3 = __floats2bfloat162_rn(f[j + 6], f[j + 7]);
              pk.x = *reinterpret_cast<uint32_t*>(&a0); pk.y = *reinterpret_cast<uint32_t*>(&a1);
              pk.z = *reinterpret_cast<uint32_t*>(&a2); pk.w = *reinterpret_cast<uint32_t*>(&a3);
              *reinterpret_cast<uint4*>(out + j) = pk;
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) if (j < ncols) out[j] = __float2bfloat16_rn(f[j]);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) pr_mbar_arrive_cluster(buf ? tempty_leader1 : tempty_leader0);
    }
  }
  __syncwarp();
  tc_fence_before();
  pr_cluster_sync();                    // no CTA leaves while its peer may still signal its barriers or read its shared memory
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512));
  }
}

static int g_pair_sms = 0;

}  // namespace genrl

using namespace genrl;

// C[M,N] = A[M,K] op(W)^T + bias on the 2-CTA kernel.  transW = 0: W is [N,K] (K-major); 1: W is [K,N] (dgrad).  Returns 2 when the shape
// is not eligible (the caller falls back to the single-CTA kernels).
extern "C" int genrl_gemm_pair(void* stream, int M, int N, int K, const void* A, long long lda, const void* W, long long ldw, int transW, void* C,
                               long long ldc, int c_dtype, const float* bias) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (M <= 0 || N <= 0) return 0;
  if (N % 128 != 0 || K <= 0 || (lda % 8) || (ldw % 8) || ((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(W)) & 15)) return 2;
  CUtensorMap ta, tb;
  if (make_map(&ta, A, M, K, lda, PR_BK, PR_BM)) return 1;
  if (!transW) { if (make_map(&tb, W, N, K, ldw, PR_BK, PR_BN / 2)) return 1; }
  else         { if (make_map(&tb, W, K, N, ldw, 64, 64)) return 1; }
  PairParams p;
  p.M = M; p.N = N; p.K = K; p.num_kb = ceil_div(K, PR_BK); p.ldc = (int)ldc; p.c_dtype = c_dtype; p.bias = bias; p.C = C;
  p.splits = 1; p.kb_per_split = p.num_kb; p.atomic = 0;
  if (g_pair_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_pair_sms, cudaDevAttrMultiProcessorCount, dev);
    if (g_pair_sms <= 0) g_pair_sms = 148;
  }
  const int tiles = ceil_div(M, 2 * PR_BM) * ceil_div(N, PR_BN);
  const int npairs = std::min(tiles, g_pair_sms / 2);
  static bool attr_set[2] = {false, false};
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * npairs, 1, 1);
  cfg.blockDim = dim3(PR_THREADS, 1, 1);
  cfg.dynamicSmemBytes = PR_SMEM;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = pdl_enabled() ? 2 : 1;
  cudaError_t e;
  if (!transW) {
    if (!attr_set[0]) { cudaFuncSetAttribute(gemm_pair_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, PR_SMEM); attr_set[0] = true; }
    e = cudaLaunchKernelEx(&cfg, gemm_pair_kernel<false, false>, ta, tb, p);
  } else {
    if (!attr_set[1]) { cudaFuncSetAttribute(gemm_pair_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, PR_SMEM); attr_set[1] = true; }
    e = cudaLaunchKernelEx(&cfg, gemm_pair_kernel<false, true>, ta, tb, p);
  }
  GENRL_REQUIRE(e == cudaSuccess, "gemm_pair launch failed: %s", cudaGetErrorString(e));
  return check_launch("gemm_pair_kernel");
}

// Weight-gradient form on the 2-CTA kernel:  C[M,N] += A^T W  with A stored [K, M] and W stored [K, N] (K = rows of the activation slabs: both
// operands MN-major), fp32 C, split over K so that every CTA pair has a work item; partial tiles are combined with fp32 atomics.
// Returns 2 when the shape is not eligible.
extern "C" int genrl_gemm_pair_wgrad(void* stream, int M, int N, int K, const void* A, long long lda, const void* W, long long ldw, float* C,
                                     long long ldc) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (M <= 0 || N <= 0 || K <= 0) return 0;
  if (M % 128 != 0 || N % 128 != 0 || (lda % 8) || (ldw % 8) || ((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(W)) & 15)) return 2;
  CUtensorMap ta, tb;
  if (make_map(&ta, A, K, M, lda, 64, 64)) return 1;
  if (make_map(&tb, W, K, N, ldw, 64, 64)) return 1;
  if (g_pair_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_pair_sms, cudaDevAttrMultiProcessorCount, dev);
    if (g_pair_sms <= 0) g_pair_sms = 148;
  }
  PairParams p;
  p.M = M; p.N = N; p.K = K; p.num_kb = ceil_div(K, PR_BK); p.ldc = (int)ldc; p.c_dtype = GENRL_F32; p.bias = nullptr; p.C = C;
  const int tiles = ceil_div(M, 2 * PR_BM) * ceil_div(N, PR_BN);
  const int max_pairs = g_pair_sms / 2;
  // as many K splits as keep every pair busy, each split at least 16 k-blocks long
  int splits = std::max(1, std::min(max_pairs / std::max(tiles, 1), p.num_kb / 16));
  p.kb_per_split = ceil_div(p.num_kb, splits);
  splits = ceil_div(p.num_kb, p.kb_per_split);
  p.splits = splits; p.atomic = 1;
  const int npairs = std::min(tiles * splits, max_pairs);
  static bool attr_set = false;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * npairs, 1, 1);
  cfg.blockDim = dim3(PR_THREADS, 1, 1);
  cfg.dynamicSmemBytes = PR_SMEM;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = pdl_enabled() ? 2 : 1;
  if (!attr_set) { cudaFuncSetAttribute(gemm_pair_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, PR_SMEM); attr_set = true; }
  cudaError_t e = cudaLaunchKernelEx(&cfg, gemm_pair_kernel<true, true>, ta, tb, p);
  GENRL_REQUIRE(e == cudaSuccess, "gemm_pair_wgrad launch failed: %s", cudaGetErrorString(e));
  return check_launch("gemm_pair_kernel(wgrad)");
}
