// GEMM for the GenRL hot path on sm_100a.
//
//   C[M,N] (+)= op(A)[M,K] * op(W)[N,K]^T + bias[N]
//
// * bf16 operands -> tcgen05.mma (UMMA 128xBNx16, fp32 accumulators in TMEM), operands staged by TMA
//   (cp.async.bulk.tensor, 128B swizzle) through an mbarrier ring; warp-specialised:
//   warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer, warps 2..5 = epilogue (tcgen05.ld).
//   Both K-major and MN-major operands are supported (transA / transW) so forward, dgrad and wgrad
//   of every Linear / lowered conv in the path run on the same kernel without transposed copies.
// * fp32 operands -> register-tiled SIMT kernel (the fp32 parity mode; 1e-4 parity needs full fp32).
//
// Replaces: every nn.Linear / nn.Conv2d / nn.ConvTranspose2d call of reference agent/dreamer_utils.py
// (e.g. :339-346 RSSM projections, :760 GRU layer, :733-735 MLP dense layers, :586,:668 conv stacks).
#include "common.cuh"
#include "tc_ptx.cuh"
#include <algorithm>
#include <cstdlib>

namespace genrl {

// =================================================================================================
// tcgen05 GEMM kernel
// =================================================================================================
constexpr int TC_BM = 128;
constexpr int TC_BK = 64;
constexpr int TC_THREADS = 192;

struct TcParams {
  int M, N, K;
  int ldc;
  int c_dtype;          // GENRL_F32 / GENRL_BF16
  int accumulate;       // C += result (fp32 C only)
  int atomic;           // use atomicAdd (split-K)
  int kb_per_split;     // k-blocks handled per blockIdx.z
  int num_kb;
  const float* bias;
  void* C;
};

template <int BN, bool A_MN, bool B_MN, int STAGES>
__global__ void __launch_bounds__(TC_THREADS, (BN <= 128 ? 2 : 1))
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, TcParams p) {
  extern __shared__ uint8_t smem_raw[];
  // 1024B alignment for the 128B swizzle atoms
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  constexpr int A_BYTES = TC_BM * TC_BK * 2;
  constexpr int B_BYTES = BN * TC_BK * 2;
  constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int m0 = blockIdx.y * TC_BM;
  const int n0 = blockIdx.x * BN;
  const int kb_begin = blockIdx.z * p.kb_per_split;
  const int kb_end = min(p.num_kb, kb_begin + p.kb_per_split);
  const int nkb = kb_end - kb_begin;   // host guarantees >= 1

  constexpr int TMEM_COLS = BN <= 32 ? 32 : BN <= 64 ? 64 : BN <= 128 ? 128 : 256;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tmem_full_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  genrl::pdl_sync();      // everything above (barrier init, TMEM allocation, descriptor prefetch) overlaps the previous kernel's tail

  if (warp == 0) {
    // ------------------------------ TMA producer ------------------------------
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int kb = kb_begin; kb < kb_end; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        uint8_t* sa = smem + stage * STAGE_BYTES;
        uint8_t* sb = sa + A_BYTES;
        mbar_expect_tx(&full_bar[stage], STAGE_BYTES);
        const int k0 = kb * TC_BK;
        if (!A_MN) {
          tma_load_2d(&tmA, &full_bar[stage], sa, k0, m0);                       // box {64 k, 128 m}
        } else {
#pragma unroll
          for (int j = 0; j < TC_BM / 64; ++j)                                    // box {64 m, 64 k}
            tma_load_2d(&tmA, &full_bar[stage], sa + j * (64 * TC_BK * 2), m0 + 64 * j, k0);
        }
        if (!B_MN) {
          tma_load_2d(&tmB, &full_bar[stage], sb, k0, n0);                       // box {64 k, BN n}
        } else {
#pragma unroll
          for (int j = 0; j < BN / 64; ++j)                                       // box {64 n, 64 k}
            tma_load_2d(&tmB, &full_bar[stage], sb + j * (64 * TC_BK * 2), n0 + 64 * j, k0);
        }
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ------------------------------ MMA issuer ------------------------------
    if (lane == 0) {
      // instruction descriptor (cute InstrDescriptor): c=F32, a=b=BF16, majors, N>>3, M>>4
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((A_MN ? 1u : 0u) << 15) | ((B_MN ? 1u : 0u) << 16) |
                             ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
      int stage = 0;
      uint32_t phase = 0;
      for (int kb = 0; kb < nkb; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + stage * STAGE_BYTES);
        const uint32_t sb = sa + A_BYTES;
#pragma unroll
        for (int k = 0; k < TC_BK / 16; ++k) {
          // K-major SW128: rows 128B apart, 8-row groups 1024B apart (SBO); +32B per K=16 step.
          // MN-major SW128: 64-element MN chunks 8192B apart (LBO), 8-k-row groups 1024B apart (SBO); +2048B per K=16.
          const uint64_t ad = A_MN ? make_smem_desc(sa + k * 2048, 64 * TC_BK * 2, 1024) : make_smem_desc(sa + k * 32, 16, 1024);
          const uint64_t bd = B_MN ? make_smem_desc(sb + k * 2048, 64 * TC_BK * 2, 1024) : make_smem_desc(sb + k * 32, 16, 1024);
          umma_bf16(tmem_base, ad, bd, idesc, (kb | k) != 0 ? 1u : 0u);
        }
        umma_commit(&empty_bar[stage]);           // frees the smem slot once the MMAs have read it
        if (kb == nkb - 1) umma_commit(tmem_full_bar);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else {
    // ------------------------------ epilogue (warps 2..5) ------------------------------
    const int quarter = warp & 3;                 // TMEM lane quarter this warp may access
    mbar_wait_sleep(tmem_full_bar, 0, 64);      // idle during the whole main loop: do not compete with the TMA / MMA lanes for issue slots
    tc_fence_after();
    const int row = m0 + quarter * 32 + lane;
    const bool row_ok = row < p.M;
    const bool add_bias = p.bias != nullptr && blockIdx.z == 0;
    // Coalesced stores: an epilogue thread owns one ROW of the accumulator, so a direct 16-byte store per lane touches 32 different 128-byte
    // lines per instruction (one L1 tag cycle each).  The main loop is over (every pipeline stage has been consumed), so each warp transposes
    // its 32 rows x 32 columns chunk through a private piece of the pipeline's shared memory and writes 4 rows x 128 bytes (fp32) or
    // 8 rows x 64 bytes (bf16) per instruction.  Ragged column tiles, unaligned outputs and the split-K atomics keep the per-lane path.
    uint8_t* stg = smem + (warp - 2) * (32 * 144);
    const size_t esz = p.c_dtype == GENRL_F32 ? 4 : 2;
    const bool ld_ok = ((size_t)p.ldc * esz) % 16 == 0 && !p.atomic;
#pragma unroll 1
    for (int c = 0; c < BN; c += 32) {
      if (n0 + c >= p.N) break;                  // warp-uniform
      uint32_t v[32];
      tmem_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)c, v);
      const int col0 = n0 + c;
      const int ncols = min(32, p.N - col0);
      float f[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
      if (add_bias) {
#pragma unroll
        for (int j = 0; j < 32; ++j) if (j < ncols) f[j] += __ldg(p.bias + col0 + j);
      }
      const bool staged = ld_ok && ncols == 32 && ((reinterpret_cast<uintptr_t>(p.C) + (size_t)col0 * esz) & 15) == 0;      // warp-uniform
      if (staged && p.c_dtype == GENRL_F32) {
#pragma unroll
        for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(stg + lane * 144 + j * 4) = make_float4(f[j], f[j + 1], f[j + 2], f[j + 3]);
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int rr = i * 4 + (lane >> 3), piece = lane & 7;
          float4 o = *reinterpret_cast<const float4*>(stg + rr * 144 + piece * 16);
          const int grow = m0 + quarter * 32 + rr;
          if (grow < p.M) {
            float* out = reinterpret_cast<float*>(p.C) + (size_t)grow * p.ldc + col0 + piece * 4;
            if (p.accumulate) {
              const float4 prev = *reinterpret_cast<const float4*>(out);
              o.x += prev.x; o.y += prev.y; o.z += prev.z; o.w += prev.w;
            }
            *reinterpret_cast<float4*>(out) = o;
          }
        }
        __syncwarp();
        continue;
      }
      if (staged) {
#pragma unroll
        for (int j = 0; j < 32; j += 8) {
          uint4 pk;
          __nv_bfloat162 a0 = __floats2bfloat162_rn(f[j], f[j + 1]);
          __nv_bfloat162 a1 = __floats2bfloat162_rn(f[j + 2], f[j + 3]);
          __nv_bfloat162 a2 = __floats2bfloat162_rn(f[j + 4], f[j + 5]);
          __nv_bfloat162 a3 = __floats2bfloat162_rn(f[j + 6], f[j + 7]);
          pk.x = *reinterpret_cast<uint32_t*>(&a0); pk.y = *reinterpret_cast<uint32_t*>(&a1);
          pk.z = *reinterpret_cast<uint32_t*>(&a2); pk.w = *reinterpret_cast<uint32_t*>(&a3);
          *reinterpret_cast<uint4*>(stg + lane * 80 + j * 2) = pk;
        }
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int rr = i * 8 + (lane >> 2), piece = lane & 3;
          const uint4 w = *reinterpret_cast<const uint4*>(stg + rr * 80 + piece * 16);
          const int grow = m0 + quarter * 32 + rr;
          if (grow < p.M) *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.C) + (size_t)grow * p.ldc + col0 + piece * 8) = w;
        }
        __syncwarp();
        continue;
      }
      if (!row_ok) continue;
      if (p.c_dtype == GENRL_F32) {
        float* out = reinterpret_cast<float*>(p.C) + (size_t)row * p.ldc + col0;
        if (p.atomic) {
#pragma unroll
          for (int j = 0; j < 32; ++j) if (j < ncols) atomicAdd(out + j, f[j]);
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) if (j < ncols) out[j] = p.accumulate ? out[j] + f[j] : f[j];
        }
      } else {
        __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(p.C) + (size_t)row * p.ldc + col0;
#pragma unroll
        for (int j = 0; j < 32; ++j) if (j < ncols) out[j] = __float2bfloat16_rn(f[j]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS));
  }
}

// =================================================================================================
// Persistent 256x128 GEMM for the large-M shapes (the 36 000 / 38 400-row head and critic GEMMs: 60 % of the step's GEMM FLOPs).
// The 128x128 kernel above is bound by per-SM operand ingest there (two co-resident CTAs pull 64 KB from L2 per 4.2 MFLOP, measured
// ~90 GB/s per SM -> 800-870 TFLOP/s).  Here one CTA per SM computes a 256-row tile with TWO UMMA_M=128 instructions per k-step that
// share the B tile (48 KB per 4.2 MFLOP), walks tiles persistently, and double-buffers the 2x128-column accumulator pair in TMEM
// (2 x 256 of the 512 columns) so the 8 epilogue warps drain tile i while TMA/MMA already run tile i+1.
// A is K-major; B is K-major (forward) or MN-major (dgrad).  Plain stores only (no split-K, no accumulate).
// =================================================================================================
constexpr int PG_BM = 256;
constexpr int PG_BN = 128;
constexpr int PG_THREADS = 320;      // TMA warp, MMA warp, 8 epilogue warps (rows 0..127: warps 2..5, rows 128..255: warps 6..9)
constexpr int PG_STAGES = 4;

template <bool B_MN>
__global__ void __launch_bounds__(PG_THREADS, 1)
gemm_tc_persist_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const TcParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  constexpr int A_BYTES = PG_BM * TC_BK * 2;          // 32 KB: rows 0..127 at +0, rows 128..255 at +16 KB
  constexpr int B_BYTES = PG_BN * TC_BK * 2;          // 16 KB
  constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + PG_STAGES * STAGE_BYTES);
  uint64_t* empty_bar = full_bar + PG_STAGES;
  uint64_t* tfull_bar = empty_bar + PG_STAGES;        // [2]
  uint64_t* tempty_bar = tfull_bar + 2;               // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int nkb = p.num_kb;
  const int tiles_n = (p.N + PG_BN - 1) / PG_BN;
  const int tiles = ((p.M + PG_BM - 1) / PG_BM) * tiles_n;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < PG_STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], 256);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  genrl::pdl_sync();      // everything above (barrier init, TMEM allocation, descriptor prefetch) overlaps the previous kernel's tail

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = blockIdx.x; t < tiles; t += gridDim.x) {
        const int m0 = (t / tiles_n) * PG_BM, n0 = (t % tiles_n) * PG_BN;
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait_sleep(&empty_bar[stage], phase ^ 1, 32);
          uint8_t* sa = smem + stage * STAGE_BYTES;
          uint8_t* sb = sa + A_BYTES;
          mbar_expect_tx(&full_bar[stage], STAGE_BYTES);
          tma_load_2d(&tmA, &full_bar[stage], sa, kb * TC_BK, m0);                          // box {64 k, 256 m}
          if (!B_MN) {
            tma_load_2d(&tmB, &full_bar[stage], sb, kb * TC_BK, n0);                        // box {64 k, 128 n}
          } else {
#pragma unroll
            for (int j = 0; j < PG_BN / 64; ++j)                                            // box {64 n, 64 k}
              tma_load_2d(&tmB, &full_bar[stage], sb + j * (64 * TC_BK * 2), n0 + 64 * j, kb * TC_BK);
          }
          if (++stage == PG_STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((B_MN ? 1u : 0u) << 16) | ((uint32_t)(PG_BN >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int t = blockIdx.x; t < tiles; t += gridDim.x, ++it) {
        const int buf = it & 1;
        mbar_wait_sleep(&tempty_bar[buf], ((it >> 1) & 1) ^ 1, 32);
        tc_fence_after();
        const uint32_t acc = tmem_base + (uint32_t)(buf * 256);
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * STAGE_BYTES);
          const uint32_t sb = sa + A_BYTES;
#pragma unroll
          for (int k = 0; k < TC_BK / 16; ++k) {
            const uint64_t bd = B_MN ? make_smem_desc(sb + k * 2048, 64 * TC_BK * 2, 1024) : make_smem_desc(sb + k * 32, 16, 1024);
            const uint32_t accum = (kb | k) != 0 ? 1u : 0u;
            umma_bf16(acc, make_smem_desc(sa + k * 32, 16, 1024), bd, idesc, accum);                       // rows 0..127
            umma_bf16(acc + 128, make_smem_desc(sa + 128 * TC_BK * 2 + k * 32, 16, 1024), bd, idesc, accum); // rows 128..255
          }
          umma_commit(&empty_bar[stage]);
          if (kb == nkb - 1) umma_commit(&tfull_bar[buf]);
          if (++stage == PG_STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else {
    // ------------------------------ epilogue (warps 2..9) ------------------------------
    const int quarter = warp & 3;
    const int half = warp >= 6 ? 1 : 0;
    int it = 0;
    for (int t = blockIdx.x; t < tiles; t += gridDim.x, ++it) {
      const int buf = it & 1;
      const int m0 = (t / tiles_n) * PG_BM, n0 = (t % tiles_n) * PG_BN;
      mbar_wait_sleep(&tfull_bar[buf], (it >> 1) & 1, 64);
      tc_fence_after();
      const uint32_t acc = tmem_base + (uint32_t)(buf * 256 + half * 128) + ((uint32_t)(quarter * 32) << 16);
      const int row = m0 + half * 128 + quarter * 32 + lane;
      const bool row_ok = row < p.M;
#pragma unroll 1
      for (int c = 0; c < PG_BN; c += 32) {
        if (n0 + c >= p.N) break;
        uint32_t v[32];
        tmem_ld32(acc + (uint32_t)c, v);
        if (!row_ok) continue;
        const int col0 = n0 + c;
        const int ncols = min(32, p.N - col0);
        float f[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
        if (p.bias != nullptr) {
#pragma unroll
          for (int j = 0; j < 32; ++j) if (j < ncols) f[j] += __ldg(p.bias + col0 + j);
        }
        if (p.c_dtype == GENRL_F32) {
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
      mbar_arrive_local(&tempty_bar[buf]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512));
  }
}

// =================================================================================================
// Linear + LayerNorm (+SiLU) in one kernel:  pre = A W^T + b ;  y = act(LN(pre))           (N = 128 .. 1024, multiple of 128)
// The N/128 CTAs that own the column tiles of one 128-row stripe form a thread-block CLUSTER.  Each epilogue thread owns one row of
// its CTA's 128 fp32 accumulator columns in TMEM, computes the local (mean, M2) of that row slice, publishes it in shared memory,
// and after a cluster barrier reads the other CTAs' partials through distributed shared memory (mapa + ld.shared::cluster) and merges
// them (Chan's parallel variance), so the full-row statistics never touch HBM and the separate LayerNorm launch (a read of `pre`
// and a ~8 us latency-bound kernel at 2 400 rows) disappears.  Replaces Linear -> NormLayer -> SiLU (dreamer_utils.py:738-744, :339-346).
// =================================================================================================
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ float2 ld_dsmem_f2(const float2* local, uint32_t cta) {
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(local)), "r"(cta));
  float2 v;
  asm volatile("ld.shared::cluster.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(remote) : "memory");
  return v;
}

struct LnParams {
  int M, N, K, num_kb;
  const float* bias;
  const float* gamma;
  const float* beta;
  float eps;
  int act;
  __nv_bfloat16* pre; int ldpre;
  __nv_bfloat16* y; int ldy;
  float* stats;                 // [M][2] mean, rstd
};

constexpr int LN_THREADS = 320;     // TMA warp, MMA warp, 8 epilogue warps (two per TMEM lane quarter, 64 columns each)

// 320 threads = 10 warps occupy 12 warp slots of the register file (allocation granularity 4 warps): two CTAs per SM need <= 85 registers
template <int STAGES>
__global__ void __maxnreg__(80)
gemm_ln_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const LnParams p) {
  constexpr int BN = 128;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  constexpr int A_BYTES = TC_BM * TC_BK * 2;
  constexpr int B_BYTES = BN * TC_BK * 2;
  constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);
  float2* row_part = reinterpret_cast<float2*>(smem + STAGES * STAGE_BYTES + 256);      // [2][128] (sum, sum of squares) per column half
  float* s_bias = reinterpret_cast<float*>(row_part + 256);                             // [128] each: bias, gamma, beta of this CTA's columns
  float* s_gamma = s_bias + BN;
  float* s_beta = s_gamma + BN;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int m0 = blockIdx.y * TC_BM;
  const int n0 = blockIdx.x * BN;
  const int nkb = p.num_kb;
  const int ncta = gridDim.x;                   // cluster = all column tiles of the stripe

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tmem_full_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(128));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (threadIdx.x >= 64 && threadIdx.x < 64 + BN) {          // stage the column parameters while the main loop runs
    const int c = threadIdx.x - 64;
    s_bias[c] = p.bias ? p.bias[n0 + c] : 0.f;
    s_gamma[c] = p.gamma[n0 + c];
    s_beta[c] = p.beta[n0 + c];
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  genrl::pdl_sync();      // everything above (barrier init, TMEM allocation, descriptor prefetch) overlaps the previous kernel's tail

  const int quarter = warp & 3;                 // TMEM lane quarter this warp may access
  const int half = warp >= 6 ? 1 : 0;           // warps 2..5: columns 0..63, warps 6..9: columns 64..127
  const int r = quarter * 32 + lane;            // row of the tile owned by an epilogue thread
  const uint32_t tcol = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(half * 64);

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int kb = 0; kb < nkb; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        uint8_t* sa = smem + stage * STAGE_BYTES;
        mbar_expect_tx(&full_bar[stage], STAGE_BYTES);
        tma_load_2d(&tmA, &full_bar[stage], sa, kb * TC_BK, m0);
        tma_load_2d(&tmB, &full_bar[stage], sa + A_BYTES, kb * TC_BK, n0);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
      int stage = 0;
      uint32_t phase = 0;
      for (int kb = 0; kb < nkb; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + stage * STAGE_BYTES);
        const uint32_t sb = sa + A_BYTES;
#pragma unroll
        for (int k = 0; k < TC_BK / 16; ++k)
          umma_bf16(tmem_base, make_smem_desc(sa + k * 32, 16, 1024), make_smem_desc(sb + k * 32, 16, 1024), idesc, (kb | k) != 0 ? 1u : 0u);
        umma_commit(&empty_bar[stage]);
        if (kb == nkb - 1) umma_commit(tmem_full_bar);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else {
    // ---- epilogue part A: (sum, sum of squares) of this thread's 64 columns ----
    mbar_wait_sleep(tmem_full_bar, 0, 64);
    tc_fence_after();
    float s = 0.f, ss = 0.f;
#pragma unroll
    for (int c = 0; c < 64; c += 32) {
      uint32_t v[32];
      tmem_ld32(tcol + (uint32_t)c, v);
#pragma unroll
      for (int q = 0; q < 32; ++q) { const float x = __uint_as_float(v[q]) + s_bias[half * 64 + c + q]; s += x; ss = fmaf(x, x, ss); }
    }
    row_part[half * 128 + r] = make_float2(s, ss);
  }
  __syncwarp();
  cluster_sync_all();                            // every CTA of the stripe has published its row partials
  if (warp >= 2) {
    // ---- epilogue part B: merge the partials over the cluster, normalise, activate, store ----
    float2 t0[8], t1[8];                          // all remote loads in flight together (~215 cycles each), then the sums
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const uint32_t src = q < ncta ? (uint32_t)q : 0u;
      t0[q] = ld_dsmem_f2(&row_part[r], src); t1[q] = ld_dsmem_f2(&row_part[128 + r], src);
    }
    float s = 0.f, ss = 0.f;
#pragma unroll
    for (int q = 0; q < 8; ++q) if (q < ncta) { s += t0[q].x + t1[q].x; ss += t0[q].y + t1[q].y; }
    const float mean = s / p.N;
    const float rstd = 1.0f / sqrtf(fmaxf(ss / p.N - mean * mean, 0.f) + p.eps);
    const int row = m0 + r;
    const bool row_ok = row < p.M;
    if (row_ok && n0 == 0 && half == 0) { p.stats[(size_t)row * 2] = mean; p.stats[(size_t)row * 2 + 1] = rstd; }
#pragma unroll
    for (int c = 0; c < 64; c += 32) {
      uint32_t v[32];
      tmem_ld32(tcol + (uint32_t)c, v);
      if (!row_ok) continue;
      const int cc = half * 64 + c;
      __nv_bfloat16* po = p.pre + (size_t)row * p.ldpre + n0 + cc;
      __nv_bfloat16* yo = p.y + (size_t)row * p.ldy + n0 + cc;
#pragma unroll
      for (int q = 0; q < 32; q += 8) {
        float x[8], o[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          x[e] = __uint_as_float(v[q + e]) + s_bias[cc + q + e];
          const float a = (x[e] - mean) * rstd * s_gamma[cc + q + e] + s_beta[cc + q + e];
          o[e] = p.act == GENRL_ACT_SILU ? silu_t<true>(a) : a;
        }
        uint4 pk, yk;
        __nv_bfloat162 a0 = __floats2bfloat162_rn(x[0], x[1]), a1 = __floats2bfloat162_rn(x[2], x[3]);
        __nv_bfloat162 a2 = __floats2bfloat162_rn(x[4], x[5]), a3 = __floats2bfloat162_rn(x[6], x[7]);
        pk.x = *reinterpret_cast<uint32_t*>(&a0); pk.y = *reinterpret_cast<uint32_t*>(&a1);
        pk.z = *reinterpret_cast<uint32_t*>(&a2); pk.w = *reinterpret_cast<uint32_t*>(&a3);
        __nv_bfloat162 b0 = __floats2bfloat162_rn(o[0], o[1]), b1 = __floats2bfloat162_rn(o[2], o[3]);
        __nv_bfloat162 b2 = __floats2bfloat162_rn(o[4], o[5]), b3 = __floats2bfloat162_rn(o[6], o[7]);
        yk.x = *reinterpret_cast<uint32_t*>(&b0); yk.y = *reinterpret_cast<uint32_t*>(&b1);
        yk.z = *reinterpret_cast<uint32_t*>(&b2); yk.w = *reinterpret_cast<uint32_t*>(&b3);
        *reinterpret_cast<uint4*>(po + q) = pk;
        *reinterpret_cast<uint4*>(yo + q) = yk;
      }
    }
  }
  __syncwarp();
  tc_fence_before();
  cluster_sync_all();                            // no CTA leaves (and frees its shared memory) while a peer may still read its partials
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(128));
  }
}

// =================================================================================================
// fp32 SIMT GEMM (parity mode): generic strides, 64x64x16 tiles, 4x4 outputs per thread
// =================================================================================================
struct SimtParams {
  int M, N, K;
  const float* A; long long sam, sak;     // A(m,k) = A[m*sam + k*sak]
  const float* W; long long swn, swk;     // W(n,k) = W[n*swn + k*swk]
  float* C; int ldc;
  const float* bias;
  int accumulate;
};

__global__ void __launch_bounds__(256) gemm_simt_kernel(SimtParams p) {
  genrl::pdl_sync();
  __shared__ float As[16][64 + 4];
  __shared__ float Ws[16][64 + 4];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  // loader mapping: pick the contiguous direction to be the fast thread index
  const bool a_kfast = (p.sak == 1);
  const bool w_kfast = (p.swk == 1);
  for (int k0 = 0; k0 < p.K; k0 += 16) {
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      int e = tid + it * 256;                 // 1024 elements per tile
      int mm, kk;
      if (a_kfast) { kk = e & 15; mm = e >> 4; } else { mm = e & 63; kk = e >> 6; }
      int gm = m0 + mm, gk = k0 + kk;
      As[kk][mm] = (gm < p.M && gk < p.K) ? p.A[gm * p.sam + gk * p.sak] : 0.f;
      int nn, k2;
      if (w_kfast) { k2 = e & 15; nn = e >> 4; } else { nn = e & 63; k2 = e >> 6; }
      int gn = n0 + nn; gk = k0 + k2;
      Ws[k2][nn] = (gn < p.N && gk < p.K) ? p.W[gn * p.swn + gk * p.swk] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      float a[4], w[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) w[j] = Ws[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], w[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int gm = m0 + ty * 4 + i;
    if (gm >= p.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int gn = n0 + tx * 4 + j;
      if (gn >= p.N) continue;
      float v = acc[i][j] + (p.bias ? p.bias[gn] : 0.f);
      float* o = p.C + (size_t)gm * p.ldc + gn;
      *o = p.accumulate ? *o + v : v;
    }
  }
}

// =================================================================================================
// host side
// =================================================================================================
template <int BN, bool A_MN, bool B_MN>
static int launch_tc(const CUtensorMap& ta, const CUtensorMap& tb, TcParams p, int splits, cudaStream_t st) {
  constexpr int STAGE_BYTES = (TC_BM + BN) * TC_BK * 2;
  // BN <= 128: 3 stages x <=32 KB -> two CTAs per SM (one CTA's prologue/epilogue overlaps the other's main loop, 6 stages in
  // flight per SM); wider tiles: 4 stages, one CTA per SM.
  constexpr int STAGES = (BN <= 128) ? 3 : 4;
  constexpr int SMEM = STAGES * STAGE_BYTES + 1024 + 256;
  static bool attr_set = false;
  auto kern = gemm_tc_kernel<BN, A_MN, B_MN, STAGES>;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
    GENRL_REQUIRE(e == cudaSuccess, "cudaFuncSetAttribute(gemm_tc) failed: %s", cudaGetErrorString(e));
    attr_set = true;
  }
  dim3 grid(ceil_div(p.N, BN), ceil_div(p.M, TC_BM), splits);
  genrl::launch_pdl(kern, grid, TC_THREADS, SMEM, st, ta, tb, p);
  return check_launch("gemm_tc_kernel");
}

template <bool A_MN, bool B_MN>
static int dispatch_bn(int bn, const CUtensorMap& ta, const CUtensorMap& tb, TcParams p, int splits, cudaStream_t st) {
  switch (bn) {
    case 64: return launch_tc<64, A_MN, B_MN>(ta, tb, p, splits, st);
    case 128: return launch_tc<128, A_MN, B_MN>(ta, tb, p, splits, st);
    case 160: if constexpr (!B_MN) return launch_tc<160, A_MN, B_MN>(ta, tb, p, splits, st); else return launch_tc<192, A_MN, B_MN>(ta, tb, p, splits, st);
    case 192: return launch_tc<192, A_MN, B_MN>(ta, tb, p, splits, st);
    default: return launch_tc<256, A_MN, B_MN>(ta, tb, p, splits, st);
  }
}

static int g_num_sms = 0;
static int num_sms() {
  if (g_num_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (g_num_sms <= 0) g_num_sms = 148;
  }
  return g_num_sms;
}

// Tile policy, from measurements on B200 (tools/tune_gemm.py, profiles/r01_kineto_step_bf16.txt "GEMM by shape", profiles/r01_notes.md):
// at the sizes of this path the kernel is bound by operand-feed LATENCY (a 16..38-deep K loop fed from L2/HBM), not by the tensor
// pipe or by bandwidth, so what matters is bytes in flight per SM: 128x128 tiles with two co-resident CTAs per SM (2 x 3 stages,
// and one CTA's prologue/epilogue hidden behind the other's main loop) beat wider single-CTA tiles by 2-3x on cold operands.
// Split-K only for few-tile / very deep-K problems (conv wgrads), keeping >= 32 k-blocks per split.
static void choose_tiling(int M, int N, int K, bool allow_split, bool b_mn, int* bn_out, int* splits_out) {
  (void)b_mn;
  const int sms = num_sms();
  const int bn = N <= 64 ? 64 : 128;
  const int num_kb = ceil_div(K, TC_BK);
  const long long tiles = (long long)ceil_div(M, TC_BM) * ceil_div(N, bn);
  int sp = 1;
  if (allow_split) {
    while ((long long)sp * 2 * tiles <= 2LL * sms && sp < 512) sp *= 2;
    while (sp > 1 && num_kb / sp < 32) sp /= 2;
  }
  int best_bn = bn, best_s = sp;
  *bn_out = best_bn;
  *splits_out = best_s;
  // tuning hook (tools/tune_gemm.py): GENRL_GEMM_BN / GENRL_GEMM_SPLITS force a configuration
  static const char* fb = getenv("GENRL_GEMM_FORCE");
  if (fb) {
    const char* e1 = getenv("GENRL_GEMM_BN"); const char* e2 = getenv("GENRL_GEMM_SPLITS");
    if (e1 && atoi(e1) > 0) { int v = atoi(e1); if (!(b_mn && (v % 64))) *bn_out = v; }
    if (e2 && atoi(e2) > 0 && allow_split) *splits_out = std::min(atoi(e2), num_kb);
  }
}

}  // namespace genrl

using namespace genrl;

// C-ABI -------------------------------------------------------------------------------------------
// C[M,N] (+)= op(A) * op(W)^T + bias.  A: [M,K] row-major (lda) or, if transA, [K,M] row-major (lda).
// W: [N,K] row-major (ldw) or, if transW, [K,N] row-major (ldw).  in_dtype: GENRL_F32 (SIMT) or
// GENRL_BF16 (tcgen05).  C: [M,N] row-major (ldc) of c_dtype; accumulate requires fp32 C.
extern "C" int genrl_gemm_pair(void* stream, int M, int N, int K, const void* A, long long lda, const void* W, long long ldw, int transW, void* C,
                               long long ldc, int c_dtype, const float* bias);

extern "C" int genrl_gemm_pair_wgrad(void* stream, int M, int N, int K, const void* A, long long lda, const void* W, long long ldw, float* C,
                                     long long ldc);
extern "C" int genrl_gemm(void* stream, int M, int N, int K, const void* A, long long lda, int transA, const void* W,
                          long long ldw, int transW, int in_dtype, void* C, long long ldc, int c_dtype, const float* bias,
                          int accumulate) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (M <= 0 || N <= 0) return 0;
  GENRL_REQUIRE(K > 0, "genrl_gemm: K must be positive");
  if (in_dtype == GENRL_F32) {
    GENRL_REQUIRE(c_dtype == GENRL_F32, "genrl_gemm: fp32 operands need fp32 output");
    SimtParams p;
    p.M = M; p.N = N; p.K = K;
    p.A = reinterpret_cast<const float*>(A);
    p.sam = transA ? 1 : lda; p.sak = transA ? lda : 1;
    p.W = reinterpret_cast<const float*>(W);
    p.swn = transW ? 1 : ldw; p.swk = transW ? ldw : 1;
    p.C = reinterpret_cast<float*>(C); p.ldc = (int)ldc; p.bias = bias; p.accumulate = accumulate;
    dim3 grid(ceil_div(N, 64), ceil_div(M, 64));
    genrl::launch_pdl(gemm_simt_kernel, grid, 256, 0, st, p);
    return check_launch("gemm_simt_kernel");
  }
  GENRL_REQUIRE(in_dtype == GENRL_BF16, "genrl_gemm: unknown operand dtype %d", in_dtype);
  GENRL_REQUIRE(!(accumulate && c_dtype != GENRL_F32), "genrl_gemm: accumulate needs fp32 C");
  GENRL_REQUIRE((lda % 8) == 0 && (ldw % 8) == 0, "genrl_gemm: bf16 leading dimensions must be multiples of 8 (lda=%lld ldw=%lld)", lda, ldw);
  GENRL_REQUIRE((reinterpret_cast<uintptr_t>(A) & 15) == 0 && (reinterpret_cast<uintptr_t>(W) & 15) == 0, "genrl_gemm: operands must be 16B aligned");
  // large-M forward / dgrad shapes: persistent 256x128 kernel (see gemm_tc_persist_kernel)
  // GENRL_GEMM_PERSIST: 0 = never, 2 = whenever the kernel supports the call (tests), default = large-M shapes only
  const char* pm = getenv("GENRL_GEMM_PERSIST");
  const int persist_mode = pm ? atoi(pm) : 1;
  // 2-CTA kernel (gemm_pair.cu) first: measured 1.18-1.27 PFLOP/s on the 38 400-row shapes against 0.98 for the 256x128 single-CTA
  // persistent kernel (cuBLASLt: 1.21-1.23), and ahead of the 128x128 kernel from ~64 pair tiles up (2400x3072x2048: 28.7 vs 34.6 us);
  // below that a launch has fewer 256x256 tiles than CTA pairs and the 128x128 kernel's two CTAs per SM win (tools/gemm_bar.py)
  static int pair_mode = -1;
  if (pair_mode < 0) { const char* e = getenv("GENRL_GEMM_PAIR"); pair_mode = e ? atoi(e) : 1; }
  if (pair_mode && !transA && !accumulate && N % 128 == 0) {
    const long long pair_tiles = (long long)ceil_div(M, 256) * ceil_div(N, 256);
    if (pair_mode == 2 || pair_tiles >= 64) {
      const int rc = genrl_gemm_pair(stream, M, N, K, A, lda, W, ldw, transW, C, ldc, c_dtype, bias);
      if (rc != 2) return rc;
    }
  }
  // weight gradients over long row slabs (C += A^T W, both operands MN-major, K = rows): the 128x128 split-K kernel moves 0.0156 operand
  // bytes per FLOP through the L2 -> SM fabric (1024x1024x36000: 93 us = 0.81 PFLOP/s); the 2-CTA kernel needs half of that
  static int pair_wgrad = -1;
  if (pair_wgrad < 0) { const char* e = getenv("GENRL_GEMM_PAIR_WGRAD"); pair_wgrad = e ? atoi(e) : 1; }
  if (pair_mode && pair_wgrad && transA && transW && accumulate && c_dtype == GENRL_F32 && bias == nullptr && M % 128 == 0 && N % 128 == 0 &&
      (pair_wgrad == 2 || (K >= 8192 && (long long)M * N >= 512 * 512))) {
    const int rc = genrl_gemm_pair_wgrad(stream, M, N, K, A, lda, W, ldw, reinterpret_cast<float*>(C), ldc);
    if (rc != 2) return rc;
  }
  const bool persist_ok = !transA && !accumulate;
  // measured (tools/debug_pgemm.py): 1.02-1.15 PFLOP/s vs 0.79-1.0 for N >= 1024; N = 255 (300 tiles = 2.03 waves) loses to the 128x128 kernel
  const bool persist_auto = M >= 8192 && N >= 512 && (long long)ceil_div(M, PG_BM) * ceil_div(N, PG_BN) >= 4 * num_sms();
  if (persist_ok && (persist_mode == 2 || (persist_mode == 1 && persist_auto))) {
    CUtensorMap pa, pb;
    if (make_map(&pa, A, M, K, lda, 64, PG_BM)) return 1;
    if (!transW) { if (make_map(&pb, W, N, K, ldw, 64, PG_BN)) return 1; }
    else         { if (make_map(&pb, W, K, N, ldw, 64, 64)) return 1; }
    TcParams q;
    q.M = M; q.N = N; q.K = K; q.ldc = (int)ldc; q.c_dtype = c_dtype; q.accumulate = 0; q.atomic = 0;
    q.num_kb = ceil_div(K, TC_BK); q.kb_per_split = q.num_kb; q.bias = bias; q.C = C;
    constexpr int SMEM = PG_STAGES * (PG_BM + PG_BN) * TC_BK * 2 + 1024 + 256;
    static bool attr_set[2] = {false, false};
    const int tiles = ceil_div(M, PG_BM) * ceil_div(N, PG_BN);
    const int grid = std::min(tiles, num_sms());
    if (!transW) {
      if (!attr_set[0]) { cudaFuncSetAttribute(gemm_tc_persist_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM); attr_set[0] = true; }
      genrl::launch_pdl(gemm_tc_persist_kernel<false>, grid, PG_THREADS, SMEM, st, pa, pb, q);
    } else {
      if (!attr_set[1]) { cudaFuncSetAttribute(gemm_tc_persist_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM); attr_set[1] = true; }
      genrl::launch_pdl(gemm_tc_persist_kernel<true>, grid, PG_THREADS, SMEM, st, pa, pb, q);
    }
    return check_launch("gemm_tc_persist_kernel");
  }
  int bn, splits;
  choose_tiling(M, N, K, c_dtype == GENRL_F32, transW != 0, &bn, &splits);
  CUtensorMap ta, tb;
  if (!transA) { if (make_map(&ta, A, M, K, lda, 64, TC_BM)) return 1; }
  else         { if (make_map(&ta, A, K, M, lda, 64, 64)) return 1; }
  if (!transW) { if (make_map(&tb, W, N, K, ldw, 64, bn)) return 1; }
  else         { if (make_map(&tb, W, K, N, ldw, 64, 64)) return 1; }
  TcParams p;
  p.M = M; p.N = N; p.K = K; p.ldc = (int)ldc; p.c_dtype = c_dtype; p.accumulate = accumulate;
  p.num_kb = ceil_div(K, TC_BK);
  p.kb_per_split = ceil_div(p.num_kb, splits);
  splits = ceil_div(p.num_kb, p.kb_per_split);
  p.atomic = splits > 1;
  if (p.atomic && !accumulate) {           // split-K partial sums are added atomically: start from zero
    cudaError_t e = cudaMemset2DAsync(C, (size_t)ldc * 4, 0, (size_t)N * 4, (size_t)M, st);
    GENRL_REQUIRE(e == cudaSuccess, "genrl_gemm: memset failed: %s", cudaGetErrorString(e));
  }
  p.bias = bias; p.C = C;
  if (!transA && !transW) return dispatch_bn<false, false>(bn, ta, tb, p, splits, st);
  if (!transA && transW) return dispatch_bn<false, true>(bn, ta, tb, p, splits, st);
  if (transA && !transW) return dispatch_bn<true, false>(bn, ta, tb, p, splits, st);
  return dispatch_bn<true, true>(bn, ta, tb, p, splits, st);
}

// pre[M,N] = A[M,K] W[N,K]^T + bias ; y = act(LayerNorm_N(pre) * gamma + beta) ; stats[M][2] = (mean, rstd).  bf16 operands and outputs,
// N a multiple of 128 up to 1024 (the column tiles of a row stripe form one thread-block cluster).
extern "C" int genrl_gemm_ln_fwd(void* stream, int M, int N, int K, const void* A, long long lda, const void* W, long long ldw, const float* bias,
                                 const float* gamma, const float* beta, float eps, int act, void* pre, long long ldpre, void* y, long long ldy,
                                 float* stats) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (M <= 0) return 0;
  GENRL_REQUIRE(N % 128 == 0 && N >= 128 && N <= 1024, "gemm_ln_fwd: N must be a multiple of 128 in [128, 1024] (got %d)", N);
  GENRL_REQUIRE(K > 0 && (lda % 8) == 0 && (ldw % 8) == 0 && (ldpre % 8) == 0 && (ldy % 8) == 0, "gemm_ln_fwd: leading dimensions must be multiples of 8");
  GENRL_REQUIRE(((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(W) | reinterpret_cast<uintptr_t>(pre) | reinterpret_cast<uintptr_t>(y)) & 15) == 0,
                "gemm_ln_fwd: operands must be 16-byte aligned");
  CUtensorMap ta, tb;
  if (make_map(&ta, A, M, K, lda, 64, TC_BM)) return 1;
  if (make_map(&tb, W, N, K, ldw, 64, 128)) return 1;
  LnParams p;
  p.M = M; p.N = N; p.K = K; p.num_kb = ceil_div(K, TC_BK);
  p.bias = bias; p.gamma = gamma; p.beta = beta; p.eps = eps; p.act = act;
  p.pre = reinterpret_cast<__nv_bfloat16*>(pre); p.ldpre = (int)ldpre;
  p.y = reinterpret_cast<__nv_bfloat16*>(y); p.ldy = (int)ldy; p.stats = stats;
  constexpr int STAGES = 3;
  constexpr int SMEM = STAGES * (TC_BM + 128) * TC_BK * 2 + 1024 + 256 + 2048 + 1536;
  auto kern = gemm_ln_kernel<STAGES>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
    GENRL_REQUIRE(e == cudaSuccess, "cudaFuncSetAttribute(gemm_ln) failed: %s", cudaGetErrorString(e));
    // cluster launches are admitted by the occupancy the driver computes for the *current* carve-out: ask for the shared-memory-heavy
    // split explicitly, otherwise only one CTA per SM (15 clusters of 8) is co-scheduled and a 19-cluster grid takes two waves
    cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    attr_set = true;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(N / 128, ceil_div(M, TC_BM), 1);
  cfg.blockDim = dim3(LN_THREADS, 1, 1);
  cfg.dynamicSmemBytes = SMEM;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = N / 128; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = pdl_enabled() ? 2 : 1;
  if (getenv("GENRL_GEMM_DEBUG")) {
    int ncl = -1;
    cudaError_t eo = cudaOccupancyMaxActiveClusters(&ncl, kern, &cfg);
    fprintf(stderr, "gemm_ln_fwd: M=%d N=%d K=%d cluster=%d grid=(%d,%d) max active clusters %d (%s)\n", M, N, K, N / 128, N / 128, ceil_div(M, TC_BM), ncl,
            cudaGetErrorString(eo));
  }
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, ta, tb, p);
  GENRL_REQUIRE(e == cudaSuccess, "gemm_ln_fwd launch failed: %s", cudaGetErrorString(e));
  return check_launch("gemm_ln_kernel");
}

// Debug aid: function attributes and the occupancy the runtime computes for the tensor-core kernels (printed to stderr).
template <typename K> static void dbg_occ(const char* name, K kern, int threads, int smem) {
  cudaFuncAttributes fa;
  cudaFuncGetAttributes(&fa, kern);
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  int occ = -1;
  cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, threads, smem);
  fprintf(stderr, "%-28s regs %3d static smem %5zu local %4zu maxThreads %4d | threads %d dyn smem %6d -> blocks/SM %d (%s)\n", name, fa.numRegs,
          fa.sharedSizeBytes, fa.localSizeBytes, fa.maxThreadsPerBlock, threads, smem, occ, cudaGetErrorString(e));
}
extern "C" int genrl_debug_occupancy() {
  dbg_occ("gemm_tc_kernel<128,3>", gemm_tc_kernel<128, false, false, 3>, TC_THREADS, 3 * 32768 + 1280);
  dbg_occ("gemm_tc_persist_kernel", gemm_tc_persist_kernel<false>, PG_THREADS, PG_STAGES * 49152 + 1280);
  dbg_occ("gemm_ln_kernel<3>", gemm_ln_kernel<3>, LN_THREADS, 3 * 32768 + 1024 + 256 + 2048 + 1536);
  dbg_occ("gemm_ln_kernel<3> 192thr", gemm_ln_kernel<3>, 192, 3 * 32768 + 1024 + 256 + 2048 + 1536);
  dbg_occ("gemm_ln_kernel<3> 64KB", gemm_ln_kernel<3>, LN_THREADS, 65536);
  return 0;
}
