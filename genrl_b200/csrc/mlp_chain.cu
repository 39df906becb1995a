// Persistent cluster kernel for CHAINS of  y_l = SiLU(LayerNorm(y_{l-1} W_l^T + b_l))  layers (bf16 operands, fp32 accumulation):
//   reference MLP dreamer_utils.py:718-747 (dense{l} -> norm{l} -> act), the RSSM input / output blocks :339-346, NormLayer :844-859.
//
// Why a new kernel.  Every hidden layer of the model is a 1024-wide Linear + LayerNorm + SiLU.  As separate launches a layer costs a GEMM
// (14 us at the 2 400 rows of the imagination rollout: ~8 us of launch / prologue / pipeline-fill / epilogue around a ~6 us L2-bound
// main loop) plus a LayerNorm kernel (6 us, a second pass over the activations); the 8-CTA cluster kernel of round 1 (gemm_ln_kernel)
// fuses the two but needs 19 clusters of 8 where only 15 are co-resident, so it runs in two waves and wins nothing at N = 1024.
// Here
//   * a cluster is CL = N/256 CTAs (4 at N = 1024); CTA r owns output columns [256 r, 256 r + 256) of a 128-row stripe: one
//     UMMA 128x256x16 per k-step, fp32 accumulator = 256 TMEM columns.  19 stripes x 4 CTAs = 76 CTAs at 2 400 rows: ONE wave
//     (33 clusters of 4 are co-resident); at 38 400 rows the 33 clusters loop over the 300 stripes;
//   * the cluster walks the WHOLE chain for its stripe: layer l+1 reads the stripe's y_l (written to global memory by the 4 CTAs,
//     so still in L2) with TMA; pipeline barriers, TMEM and the tensor-map prefetches are set up once per launch instead of once
//     per layer, and the next layer's weight tiles are already streaming while the previous epilogue drains;
//   * LayerNorm statistics: each epilogue thread owns one row of its CTA's accumulator, publishes (sum, sum of squares) of its 128
//     columns in shared memory; after a cluster barrier every thread merges the 2 CL partials of its row through distributed shared
//     memory, normalises from TMEM, writes pre-activation (kept for the backward), activation and (mean, rstd).
// Operand bytes per FLOP equal the 256x128 persistent GEMM (48 KB per 128x256x64 MACs).
#include "tc_ptx.cuh"
#include <algorithm>

namespace genrl {

constexpr int CH_BM = 128;
constexpr int CH_BN = 256;
constexpr int CH_BK = 64;
constexpr int CH_STAGES = 4;
constexpr int CH_THREADS = 320;      // warp 0: TMA, warp 1: MMA, warps 2..9: epilogue (lane quarter = warp & 3, column half = warp >= 6)
constexpr int CH_MAXL = 6;
constexpr int CH_A_BYTES = CH_BM * CH_BK * 2;      // 16 KB
constexpr int CH_B_BYTES = CH_BN * CH_BK * 2;      // 32 KB
constexpr int CH_STAGE_BYTES = CH_A_BYTES + CH_B_BYTES;
// epilogue staging: an epilogue thread owns one ROW of the accumulator, so a direct 16-byte store per lane touches 32 different 128-byte lines
// per instruction (one L1 tag cycle each: 8 warps x 32 stores x 32 lines = 8 192 cycles = 4 us per layer).  Each warp therefore transposes its
// 32 rows x 32 columns chunk through a private 32 x 80-byte buffer and writes 8 rows x 64 bytes per instruction (8 lines).
constexpr int CH_STG_ROW = 80;                     // 64 data bytes + 16 padding (conflict-free 16-byte accesses in both directions)
constexpr int CH_STG_BYTES = 32 * CH_STG_ROW;      // per epilogue warp
constexpr int CH_SMEM = CH_STAGES * CH_STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/ + 2 * 128 * 8 /*row partials*/ + 3 * CH_BN * 4 /*column params*/ +
                        8 * CH_STG_BYTES /*store staging*/;

struct ChainLayer {
  const float* bias; const float* gamma; const float* beta;
  float eps;
  int num_kb;
  __nv_bfloat16* pre; long long ldpre;
  __nv_bfloat16* y; long long ldy;
  float* stats;
};
struct ChainParams {
  int M, N, L, n_stripes, act;
  ChainLayer layer[CH_MAXL];
};
struct ChainMaps {
  CUtensorMap a[CH_MAXL];      // layer input:  x for l = 0, y_{l-1} otherwise      box {64 k, 128 rows}
  CUtensorMap b[CH_MAXL];      // layer weight W_l [N, K_l]                         box {64 k, 256 rows}
};

__device__ __forceinline__ uint32_t ch_cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t ch_cluster_id() { uint32_t r; asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t ch_num_clusters() { uint32_t r; asm volatile("mov.u32 %0, %%nclusterid.x;" : "=r"(r)); return r; }
__device__ __forceinline__ void ch_cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ float2 ch_ld_dsmem_f2(const float2* local, uint32_t cta) {
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(local)), "r"(cta));
  float2 v;
  asm volatile("ld.shared::cluster.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(remote) : "memory");
  return v;
}
// generic-proxy global writes (the epilogue's y_l stores) -> async-proxy reads (the next layer's TMA loads)
__device__ __forceinline__ void ch_fence_proxy_async_global() { asm volatile("fence.proxy.async.global;" ::: "memory"); }
__device__ __forceinline__ void ch_bar_epilogue() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

__global__ void __launch_bounds__(CH_THREADS, 1)
mlp_chain_fwd_kernel(const __grid_constant__ ChainMaps maps, const ChainParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + CH_STAGES * CH_STAGE_BYTES);
  uint64_t* empty_bar = full_bar + CH_STAGES;
  uint64_t* tmem_full_bar = empty_bar + CH_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);
  float2* row_part = reinterpret_cast<float2*>(smem + CH_STAGES * CH_STAGE_BYTES + 256);      // [2 halves][128 rows]
  float* s_bias = reinterpret_cast<float*>(row_part + 256);                                    // [256] each
  float* s_gamma = s_bias + CH_BN;
  float* s_beta = s_gamma + CH_BN;
  uint8_t* s_stage = reinterpret_cast<uint8_t*>(s_beta + CH_BN);                              // [8 epilogue warps][32][80]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = ch_cluster_ctarank();
  const int ncta = p.N / CH_BN;                       // cluster size
  const int n0 = (int)rank * CH_BN;
  const int cluster = (int)ch_cluster_id();
  const int nclusters = (int)ch_num_clusters();

  if (threadIdx.x == 0) {
    for (int l = 0; l < p.L; ++l) { tma_prefetch_desc(&maps.a[l]); tma_prefetch_desc(&maps.b[l]); }
    for (int s = 0; s < CH_STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(tmem_full_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(CH_BN));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  genrl::pdl_sync();
  ch_cluster_sync();                                  // every CTA of the cluster is running (its shared memory may be read remotely)

  const int quarter = warp & 3;
  const int half = warp >= 6 ? 1 : 0;
  const int r = quarter * 32 + lane;                  // accumulator row owned by an epilogue thread
  const uint32_t tcol = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(half * 128);

  int stage = 0;                                      // pipeline position (producer lane and MMA lane keep their own copy)
  uint32_t phase = 0;
  uint32_t it = 0;                                    // (stripe, layer) iterations done: parity of tmem_full_bar

  for (int stripe = cluster; stripe < p.n_stripes; stripe += nclusters) {
    const int m0 = stripe * CH_BM;
    for (int l = 0; l < p.L; ++l, ++it) {
      const ChainLayer& ly = p.layer[l];
      const int nkb = ly.num_kb;
      if (warp == 0) {
        if (lane == 0) {
          for (int kb = 0; kb < nkb; ++kb) {
            mbar_wait(&empty_bar[stage], phase ^ 1);
            uint8_t* sa = smem + stage * CH_STAGE_BYTES;
            mbar_expect_tx(&full_bar[stage], CH_STAGE_BYTES);
            tma_load_2d(&maps.a[l], &full_bar[stage], sa, kb * CH_BK, m0);
            tma_load_2d(&maps.b[l], &full_bar[stage], sa + CH_A_BYTES, kb * CH_BK, n0);
            if (++stage == CH_STAGES) { stage = 0; phase ^= 1; }
          }
        }
      } else if (warp == 1) {
        if (lane == 0) {
          const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(CH_BN >> 3) << 17) | ((uint32_t)(CH_BM >> 4) << 24);
          for (int kb = 0; kb < nkb; ++kb) {
            mbar_wait(&full_bar[stage], phase);
            tc_fence_after();
            const uint32_t sa = smem_u32(smem + stage * CH_STAGE_BYTES);
            const uint32_t sb = sa + CH_A_BYTES;
#pragma unroll
            for (int k = 0; k < CH_BK / 16; ++k)
              umma_bf16(tmem_base, make_smem_desc(sa + k * 32, 16, 1024), make_smem_desc(sb + k * 32, 16, 1024), idesc, (kb | k) != 0 ? 1u : 0u);
            umma_commit(&empty_bar[stage]);
            if (kb == nkb - 1) umma_commit(tmem_full_bar);
            if (++stage == CH_STAGES) { stage = 0; phase ^= 1; }
          }
        }
      } else {
        // column parameters of this layer for the CTA's 256 columns (the previous layer's readers passed the cluster barrier)
        const int c = threadIdx.x - 64;
        s_bias[c] = ly.bias ? ly.bias[n0 + c] : 0.f;
        s_gamma[c] = ly.gamma[n0 + c];
        s_beta[c] = ly.beta[n0 + c];
        ch_bar_epilogue();
        // ---- part A: (sum, sum of squares) over this thread's 128 columns ----
        mbar_wait_sleep(tmem_full_bar, it & 1u, 32);
        tc_fence_after();
        float s = 0.f, ss = 0.f;
#pragma unroll
        for (int cb = 0; cb < 128; cb += 32) {
          uint32_t v[32];
          tmem_ld32(tcol + (uint32_t)cb, v);
#pragma unroll
          for (int q = 0; q < 32; ++q) { const float x = __uint_as_float(v[q]) + s_bias[half * 128 + cb + q]; s += x; ss = fmaf(x, x, ss); }
        }
        row_part[half * 128 + r] = make_float2(s, ss);
      }
      __syncwarp();
      ch_cluster_sync();                              // partials of every CTA of the stripe are published
      if (warp >= 2) {
        float2 t0[4], t1[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const uint32_t src = q < ncta ? (uint32_t)q : 0u;
          t0[q] = ch_ld_dsmem_f2(&row_part[r], src); t1[q] = ch_ld_dsmem_f2(&row_part[128 + r], src);
        }
        float s = 0.f, ss = 0.f;
#pragma unroll
        for (int q = 0; q < 4; ++q) if (q < ncta) { s += t0[q].x + t1[q].x; ss += t0[q].y + t1[q].y; }
        const float mean = s / p.N;
        const float rstd = 1.0f / sqrtf(fmaxf(ss / p.N - mean * mean, 0.f) + ly.eps);
        const int row = m0 + r;
        const bool row_ok = row < p.M;
        if (row_ok && rank == 0 && half == 0) { ly.stats[(size_t)row * 2] = mean; ly.stats[(size_t)row * 2 + 1] = rstd; }
        uint8_t* stg = s_stage + (warp - 2) * CH_STG_BYTES;
        const int srow = lane >> 2, spiece = lane & 3;              // copy-out mapping: 8 rows x four 16-byte pieces per instruction
#pragma unroll 1
        for (int cb = 0; cb < 128; cb += 32) {
          uint32_t v[32];
          tmem_ld32(tcol + (uint32_t)cb, v);
          const int cc = half * 128 + cb;
          uint4 pk[4], yk[4];
#pragma unroll
          for (int q = 0; q < 32; q += 8) {
            float x[8], o[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              x[e] = __uint_as_float(v[q + e]) + s_bias[cc + q + e];
              const float a = (x[e] - mean) * rstd * s_gamma[cc + q + e] + s_beta[cc + q + e];
              o[e] = p.act == GENRL_ACT_SILU ? silu_t<true>(a) : a;
            }
            __nv_bfloat162 a0 = __floats2bfloat162_rn(x[0], x[1]), a1 = __floats2bfloat162_rn(x[2], x[3]);
            __nv_bfloat162 a2 = __floats2bfloat162_rn(x[4], x[5]), a3 = __floats2bfloat162_rn(x[6], x[7]);
            pk[q >> 3].x = *reinterpret_cast<uint32_t*>(&a0); pk[q >> 3].y = *reinterpret_cast<uint32_t*>(&a1);
            pk[q >> 3].z = *reinterpret_cast<uint32_t*>(&a2); pk[q >> 3].w = *reinterpret_cast<uint32_t*>(&a3);
            __nv_bfloat162 b0 = __floats2bfloat162_rn(o[0], o[1]), b1 = __floats2bfloat162_rn(o[2], o[3]);
            __nv_bfloat162 b2 = __floats2bfloat162_rn(o[4], o[5]), b3 = __floats2bfloat162_rn(o[6], o[7]);
            yk[q >> 3].x = *reinterpret_cast<uint32_t*>(&b0); yk[q >> 3].y = *reinterpret_cast<uint32_t*>(&b1);
            yk[q >> 3].z = *reinterpret_cast<uint32_t*>(&b2); yk[q >> 3].w = *reinterpret_cast<uint32_t*>(&b3);
          }
          // transpose through the warp's staging buffer: row-per-lane in, 8 rows x 64 bytes per instruction out
#pragma unroll
          for (int arr = 0; arr < 2; ++arr) {
#pragma unroll
            for (int q = 0; q < 4; ++q) *reinterpret_cast<uint4*>(stg + lane * CH_STG_ROW + q * 16) = arr == 0 ? pk[q] : yk[q];
            __syncwarp();
            __nv_bfloat16* base = arr == 0 ? ly.pre : ly.y;
            const long long ld = arr == 0 ? ly.ldpre : ly.ldy;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int rr = i * 8 + srow;
              const uint4 w = *reinterpret_cast<const uint4*>(stg + rr * CH_STG_ROW + spiece * 16);
              const int grow = m0 + quarter * 32 + rr;
              if (grow < p.M) *reinterpret_cast<uint4*>(base + (size_t)grow * ld + n0 + cc + spiece * 8) = w;
            }
            __syncwarp();
          }
        }
        ch_fence_proxy_async_global();                // the next layer's TMA loads (any CTA of the cluster) must see these stores
      }
      __syncwarp();
      tc_fence_before();
      // y_l complete and visible cluster-wide; row partials, column parameters and the TMEM accumulator may be overwritten
      ch_cluster_sync();
      tc_fence_after();
    }
  }
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(CH_BN));
  }
}

}  // namespace genrl

using namespace genrl;

// Chain of L layers  y_l = act(LN(y_{l-1} W_l^T + b_l)), y_{-1} = x.  All layers N wide (N = 256, 512, 768 or 1024), bf16 activations.
//   x [M, K0] (ldx); W[l] [N, K_l] bf16 (ldw[l]), K_l = N for l > 0; bias/gamma/beta[l] fp32 [N] (bias may be NULL);
//   pre[l], y[l] bf16 [M, N] (ldpre, ldy, multiples of 8); stats[l] fp32 [M, 2] = (mean, rstd).
extern "C" int genrl_mlp_chain_fwd(void* stream, int M, int N, int L, const void* x, long long ldx, int K0, const void* const* W, const long long* ldw,
                                   const float* const* bias, const float* const* gamma, const float* const* beta, float eps, int act,
                                   void* const* pre, long long ldpre, void* const* y, long long ldy, float* const* stats) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (M <= 0 || L <= 0) return 0;
  GENRL_REQUIRE(L <= CH_MAXL, "mlp_chain_fwd: at most %d layers (got %d)", CH_MAXL, L);
  GENRL_REQUIRE(N % CH_BN == 0 && N >= CH_BN && N <= 4 * CH_BN, "mlp_chain_fwd: N must be 256, 512, 768 or 1024 (got %d)", N);
  GENRL_REQUIRE(K0 > 0 && ldx % 8 == 0 && ldpre % 8 == 0 && ldy % 8 == 0, "mlp_chain_fwd: leading dimensions must be multiples of 8");
  GENRL_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0, "mlp_chain_fwd: x must be 16-byte aligned");
  ChainMaps maps;
  ChainParams p;
  memset(&p, 0, sizeof(p));
  p.M = M; p.N = N; p.L = L; p.n_stripes = ceil_div(M, CH_BM); p.act = act;
  for (int l = 0; l < L; ++l) {
    const int K = l == 0 ? K0 : N;
    const void* a = l == 0 ? x : y[l - 1];
    const long long lda = l == 0 ? ldx : ldy;
    GENRL_REQUIRE(ldw[l] % 8 == 0 && ((reinterpret_cast<uintptr_t>(W[l]) | reinterpret_cast<uintptr_t>(pre[l]) | reinterpret_cast<uintptr_t>(y[l])) & 15) == 0,
                  "mlp_chain_fwd: layer %d operands must be 16-byte aligned with leading dimensions that are multiples of 8", l);
    if (make_map(&maps.a[l], a, M, K, lda, CH_BK, CH_BM)) return 1;
    if (make_map(&maps.b[l], W[l], N, K, ldw[l], CH_BK, CH_BN)) return 1;
    ChainLayer& ly = p.layer[l];
    ly.bias = bias ? bias[l] : nullptr; ly.gamma = gamma[l]; ly.beta = beta[l]; ly.eps = eps; ly.num_kb = ceil_div(K, CH_BK);
    ly.pre = reinterpret_cast<__nv_bfloat16*>(pre[l]); ly.ldpre = ldpre;
    ly.y = reinterpret_cast<__nv_bfloat16*>(y[l]); ly.ldy = ldy;
    ly.stats = stats[l];
  }
  for (int l = L; l < CH_MAXL; ++l) { maps.a[l] = maps.a[0]; maps.b[l] = maps.b[0]; }
  static bool attr_set = false;
  static int max_clusters = 0;
  const int cl = N / CH_BN;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(mlp_chain_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, CH_SMEM);
    GENRL_REQUIRE(e == cudaSuccess, "cudaFuncSetAttribute(mlp_chain) failed: %s", cudaGetErrorString(e));
    cudaFuncSetAttribute(mlp_chain_fwd_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    attr_set = true;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.blockDim = dim3(CH_THREADS, 1, 1);
  cfg.dynamicSmemBytes = CH_SMEM;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cl; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = pdl_enabled() ? 2 : 1;
  // persistent: as many clusters as are co-resident (one CTA per SM), each looping over its share of the 128-row stripes
  static int cached_cl = 0;
  if (max_clusters == 0 || cached_cl != cl) {
    cfg.gridDim = dim3(cl * 32, 1, 1);
    int ncl = 0;
    cudaError_t eo = cudaOccupancyMaxActiveClusters(&ncl, mlp_chain_fwd_kernel, &cfg);
    max_clusters = (eo == cudaSuccess && ncl > 0) ? ncl : std::max(1, 120 / cl);
    cached_cl = cl;
    if (getenv("GENRL_GEMM_DEBUG")) fprintf(stderr, "mlp_chain_fwd: cluster of %d, %d clusters co-resident (%s)\n", cl, max_clusters, cudaGetErrorString(eo));
  }
  const int nclusters = std::min(p.n_stripes, max_clusters);
  cfg.gridDim = dim3(cl * nclusters, 1, 1);
  cudaError_t e = cudaLaunchKernelEx(&cfg, mlp_chain_fwd_kernel, maps, p);
  GENRL_REQUIRE(e == cudaSuccess, "mlp_chain_fwd launch failed: %s", cudaGetErrorString(e));
  return check_launch("mlp_chain_fwd_kernel");
}
