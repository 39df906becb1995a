// LayerNorm(+SiLU) backward / forward for LARGE row counts as a bulk-copy (TMA) pipeline through shared memory (round 2, second half).
//   reference: NormLayer dreamer_utils.py:844-859 + act of the MLPs (:738-744) and ImgChLayerNorm of the conv stacks (:862-871)
//
// Why.  The register kernels of rowops_fast.cuh keep a whole row (and, with parameter gradients, 96 column partials) per lane: 128-254
// registers, 8-16 resident warps per SM, every warp stalled on its own global loads (ncu, profiles/r01_ncu_ln_digest.txt: long-scoreboard
// 4-5 warps per issue, 18-31 % of HBM bandwidth; 124-133 us for 38 400 x 1024 where the bytes need 36 us).  The first staged attempt of this
// round (removed) loaded a tile, waited, and recomputed the SiLU-gradient math in both of its passes: twice the instructions, no overlap.
// Here
//   * one elected warp streams row tiles (R rows of dy and x, 32 KB) into a two-stage shared-memory ring with cp.async.bulk and an
//     mbarrier per stage: the next tile is in flight while the current one is computed, no register is tied up by a load;
//     256 threads x 2 blocks per SM, so that one block computes while the other sits in a row reduction or a barrier;
//   * phase 1 (warp per row, 8 warps per block): g = dy * act'(a) ONCE per element, row sums c1 = mean(g gamma), c2 = mean(g gamma xhat); g is parked
//     in shared memory as fp32 (two 16-byte planes: conflict-free for both phases);
//   * phase 2 (thread per 8-column chunk, fixed for the whole launch): dx = rstd (g gamma - c1 - xhat c2) with 16-byte coalesced stores and
//     the chunk's dgamma / dbeta / dbias partials in 24 registers — no recomputation of the activation gradient (measured: 33 instructions
//     per element in total, issue slots ~60 % busy at 16 warps per SM);
//   * column partials leave the block once, as fp32 atomics.
// HBM traffic is the minimum: dy and x read once, dx written once.
#pragma once
#include "rowops_fast.cuh"
#include "tc_ptx.cuh"
#include <algorithm>

namespace genrl {

constexpr int LNP_THREADS = 256;      // two blocks per SM: while one block sits in a row reduction or a barrier the other one computes
constexpr int LNP_BLOCKS_PER_SM = 2;
constexpr int LNP_WARPS = LNP_THREADS / 32;

__device__ __forceinline__ void lnp_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// d/da silu(a) = s (1 + a (1 - s)); with t = tanh(a / 2): s = (1 + t) / 2 and the product collapses to 0.5 (1 + t) + 0.5 (a / 2) (1 - t^2):
// one MUFU + 5 fp32 operations (the generic form takes 6)
template <bool FAST> __device__ __forceinline__ float lnp_silu_grad(float a) {
  if (!FAST) return silu_grad_t<false>(a);
  const float ha = 0.5f * a, t = tanh_fast(ha);
  return fmaf(0.5f * ha, fmaf(-t, t, 1.0f), fmaf(0.5f, t, 0.5f));
}

// 8 consecutive fp32 parameters of chunk c8 from the two-plane layout [2][CPR] float4
__device__ __forceinline__ void lnp_ld_planes(const float* base, int cpr, uint32_t c8, float (&v)[8]) {
  const float4 lo = reinterpret_cast<const float4*>(base)[c8], hi = reinterpret_cast<const float4*>(base)[cpr + c8];
  v[0] = lo.x; v[1] = lo.y; v[2] = lo.z; v[3] = lo.w; v[4] = hi.x; v[5] = hi.y; v[6] = hi.z; v[7] = hi.w;
}

// one warp: rows [r0, r0 + nr) of a (and b) -> shared memory; completion (bytes) on `bar`
template <typename T>
__device__ __forceinline__ void lnp_issue(const T* __restrict__ a, long long lda, const T* __restrict__ b, long long ldb, T* sa, T* sb,
                                          long long r0, int nr, int D, uint64_t* bar, int lane) {
  const uint32_t row_bytes = (uint32_t)D * sizeof(T);
  if (lane == 0) mbar_expect_tx(bar, (b ? 2u : 1u) * (uint32_t)nr * row_bytes);
  __syncwarp();
  if (lda == D && (!b || ldb == D)) {
    if (lane == 0) {
      lnp_bulk_g2s(sa, a + r0 * lda, (uint32_t)nr * row_bytes, bar);
      if (b) lnp_bulk_g2s(sb, b + r0 * ldb, (uint32_t)nr * row_bytes, bar);
    }
  } else {
    for (int r = lane; r < nr; r += 32) {
      lnp_bulk_g2s(sa + (size_t)r * D, a + (r0 + r) * lda, row_bytes, bar);
      if (b) lnp_bulk_g2s(sb + (size_t)r * D, b + (r0 + r) * ldb, row_bytes, bar);
    }
  }
}

// Geometry (all compile-time: D = 8 * LPR * CPL, so every loop unrolls and every shared-memory offset is a 32-bit constant expression).
//   LPR  lanes per row in the warp-per-row phase (power of two), CPL chunks of 8 elements per lane:
//        1024 = (32,4)  512 = (32,2)  256 = (32,1)  384 = (16,3)  192 = (8,3)  96 = (4,3)  48 = (2,3)   (the MLP width and the conv channel counts)
//   rows per tile R: a multiple of the phase-2 row groups (256 threads / chunks per row) that also fills the 8 * 32 / LPR phase-1 row slots
template <typename T, int LPR, int CPL> struct LnpGeom {
  static constexpr int D = 8 * LPR * CPL;
  static constexpr int CPR = LPR * CPL;                       // 8-element chunks per row
  static constexpr int RPW = 32 / LPR;                        // rows per warp pass
  static constexpr int SLOTS = LNP_WARPS * RPW;               // rows per block pass (phase 1)
  static constexpr int RG = LNP_THREADS / CPR;                // row groups (phase 2)
  static constexpr int ELEMS = sizeof(T) == 2 ? 8192 : 4096;  // <= 32 KB per stage (dy + x) and 32 KB of parked gradients per block
  static constexpr int RMAX = ELEMS / D;
  // largest R <= RMAX that is a multiple of RG and wastes the fewest phase-1 slots: try multiples of RG downwards from RMAX
  static constexpr int pick() {
    int best = RG, best_waste = 1 << 30;
    for (int r = (RMAX / RG) * RG; r >= RG; r -= RG) {
      const int passes = (r + SLOTS - 1) / SLOTS;
      const int waste = (passes * SLOTS - r) * 1000 / (passes * SLOTS) + (RMAX - r) * 200 / RMAX;     // idle slots + (mildly) smaller tiles
      if (waste < best_waste) { best_waste = waste; best = r; }
    }
    return best;
  }
  static constexpr int R = pick();
  static constexpr size_t TILE = (size_t)R * D;               // elements
  static constexpr size_t BWD_SMEM = 4 * TILE * sizeof(T) + TILE * 4 + 2 * D * 4 + 3 * R * 8 + 64;
  // forward: x only, three stages of up to 64 KB, R a multiple of the phase-1 slots
  static constexpr int FR_MAX = (sizeof(T) == 2 ? 16384 : 8192) / D;
  static constexpr int FR = FR_MAX >= SLOTS ? (FR_MAX / SLOTS) * SLOTS : FR_MAX;
  static constexpr size_t FTILE = (size_t)FR * D;
  static constexpr size_t FWD_SMEM = 3 * FTILE * sizeof(T) + 2 * D * 4 + 64;
};

template <typename T, int LPR, int CPL, bool COLG>
__global__ void __launch_bounds__(LNP_THREADS, LNP_BLOCKS_PER_SM) ln_bwd_pipe_kernel(const T* __restrict__ dy, long long lddy, const T* __restrict__ x, long long ldx,
                                                                     const float* __restrict__ gamma, const float* __restrict__ beta,
                                                                     const float* __restrict__ stats, int act, T* __restrict__ dx, long long lddx,
                                                                     float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ dlb,
                                                                     long long M) {
  using G = LnpGeom<T, LPR, CPL>;
  constexpr int D = G::D, CPR = G::CPR, R = G::R, RG = G::RG, RPW = G::RPW;
  constexpr uint32_t TILE = (uint32_t)G::TILE;
  constexpr bool FAST = sizeof(T) == 2;
  genrl::pdl_sync();
  extern __shared__ __align__(128) unsigned char lnp_smem[];
  T* const s_tiles = reinterpret_cast<T*>(lnp_smem);                                       // stage s: dy at s * 2 * TILE, x at s * 2 * TILE + TILE
  float4* const s_glo = reinterpret_cast<float4*>(lnp_smem + 4 * TILE * sizeof(T));        // g[.., 0..3] of every chunk
  float4* const s_ghi = s_glo + R * CPR;                                                   // g[.., 4..7]
  float* const s_gamma = reinterpret_cast<float*>(s_ghi + R * CPR);
  float* const s_beta = s_gamma + D;
  float2* const s_stat0 = reinterpret_cast<float2*>(s_beta + D);                           // stage s at s * R
  float2* const s_c = s_stat0 + 2 * R;
  uint64_t* const full = reinterpret_cast<uint64_t*>(s_c + R);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long ntiles = (M + R - 1) / R;
  if (threadIdx.x == 0) { mbar_init(&full[0], 1); mbar_init(&full[1], 1); fence_barrier_init(); }
  // gamma / beta as two 16-byte planes per vector (elements 0..3 and 4..7 of every 8-column chunk): the lanes of a warp read consecutive
  // 16-byte words (conflict-free) instead of 32-byte-strided float4 pairs (two wavefronts per quarter warp)
  for (int i = threadIdx.x; i < D; i += LNP_THREADS) {
    const int pl = ((i >> 2) & 1) * (CPR * 4) + (i >> 3) * 4 + (i & 3);
    s_gamma[pl] = gamma[i]; s_beta[pl] = beta[i];
  }
  long long t = blockIdx.x;
  if (t < ntiles) {
    const long long r0 = t * R;
    const int nr = (int)min((long long)R, M - r0);
    for (int r = threadIdx.x; r < nr; r += LNP_THREADS) s_stat0[r] = *reinterpret_cast<const float2*>(stats + (r0 + r) * 2);
  }
  __syncthreads();
  if (warp == 0 && t < ntiles) {
    const long long r0 = t * R;
    lnp_issue<T>(dy, lddy, x, ldx, s_tiles, s_tiles + TILE, r0, (int)min((long long)R, M - r0), D, &full[0], lane);
  }
  // phase-1 mapping: LPR lanes per row; phase-2 mapping: thread = (8-column chunk cc, row group rg)
  const int sub = lane % LPR, rwi = lane / LPR;
  const int cc = threadIdx.x % CPR, rg = threadIdx.x / CPR;
  float gam[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) gam[j] = s_gamma[(j >> 2) * (CPR * 4) + cc * 4 + (j & 3)];
  float a0[8], a1[8], a2[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { a0[j] = 0.f; a1[j] = 0.f; a2[j] = 0.f; }

  for (int k = 0; t < ntiles; t += gridDim.x, ++k) {
    const uint32_t s = (uint32_t)k & 1u;
    const long long r0 = t * R;
    const int nr = (int)min((long long)R, M - r0);
    const long long tn = t + gridDim.x;
    if (tn < ntiles) {
      // stage s^1 (tile k-1) was released by the barrier that ended the previous iteration
      const long long rn = tn * R;
      const int nn = (int)min((long long)R, M - rn);
      if (warp == 0) lnp_issue<T>(dy, lddy, x, ldx, s_tiles + (s ^ 1u) * 2u * TILE, s_tiles + (s ^ 1u) * 2u * TILE + TILE, rn, nn, D, &full[s ^ 1u], lane);
      else for (int r = threadIdx.x - 32; r < nn; r += LNP_THREADS - 32) s_stat0[(s ^ 1u) * R + r] = *reinterpret_cast<const float2*>(stats + (rn + r) * 2);
    }
    mbar_wait(&full[s], (uint32_t)(k >> 1) & 1u);
    // ---- phase 1: g = dy act'(a), row sums ----
    const T* tdy = s_tiles + s * 2u * TILE;
    const T* tx = tdy + TILE;
    const float2* tstat = s_stat0 + s * R;
#pragma unroll 1
    for (int rb = warp * RPW; rb < nr; rb += G::SLOTS) {
      const int r = rb + rwi;
      float s1 = 0.f, s2 = 0.f;
      if (r < nr) {
        const float2 st = tstat[r];
        const float rstd = st.y, nmr = -st.x * st.y;
        // two chunks in flight per lane (enough independent work to cover the shared-memory latency, half the registers of all CPL at once)
#pragma unroll
        for (int i0 = 0; i0 < CPL; i0 += 2) {
          constexpr int NB = 2;
          float d[NB][8], xv[NB][8];
#pragma unroll
          for (int u = 0; u < NB; ++u) {
            if (i0 + u < CPL) {
              const uint32_t off = (uint32_t)r * D + (uint32_t)(sub + (i0 + u) * LPR) * 8u;
              V8<T>::load(tdy + off, d[u]);
              V8<T>::load(tx + off, xv[u]);
            }
          }
#pragma unroll
          for (int u = 0; u < NB; ++u) {
            if (i0 + u < CPL) {
              const uint32_t c8 = (uint32_t)(sub + (i0 + u) * LPR);
              float g8[8], b8[8];
              lnp_ld_planes(s_gamma, CPR, c8, g8);
              if (act == GENRL_ACT_SILU) lnp_ld_planes(s_beta, CPR, c8, b8);
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const float h = FAST ? fmaf(xv[u][j], rstd, nmr) : (xv[u][j] - st.x) * rstd;
                float gg = d[u][j];
                if (act == GENRL_ACT_SILU) gg *= lnp_silu_grad<FAST>(fmaf(h, g8[j], b8[j]));
                d[u][j] = gg;
                const float gw = gg * g8[j];
                s1 += gw; s2 = fmaf(gw, h, s2);
              }
              s_glo[(uint32_t)r * CPR + c8] = make_float4(d[u][0], d[u][1], d[u][2], d[u][3]);
              s_ghi[(uint32_t)r * CPR + c8] = make_float4(d[u][4], d[u][5], d[u][6], d[u][7]);
            }
          }
        }
      }
      s1 = group_sum(s1, LPR); s2 = group_sum(s2, LPR);
      if (r < nr && sub == 0) s_c[r] = make_float2(s1 / D, s2 / D);
    }
    __syncthreads();
    // ---- phase 2: dx + column partials of chunk cc ----
    if (rg < RG) {
#pragma unroll 2
      for (int r = rg; r < nr; r += RG) {
        const float2 st = tstat[r];
        const float2 c = s_c[r];
        const float rstd = st.y, nmr = -st.x * st.y;
        const float4 lo = s_glo[(uint32_t)r * CPR + cc], hi = s_ghi[(uint32_t)r * CPR + cc];
        const float gg[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
        float xv[8], o[8];
        V8<T>::load(tx + (uint32_t)r * D + (uint32_t)cc * 8u, xv);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float h = FAST ? fmaf(xv[j], rstd, nmr) : (xv[j] - st.x) * rstd;
          if (COLG) { a0[j] = fmaf(gg[j], h, a0[j]); a1[j] += gg[j]; }
          o[j] = rstd * (gg[j] * gam[j] - c.x - h * c.y);
          if (COLG) a2[j] += o[j];
        }
        V8<T>::store(dx + (r0 + r) * lddx + cc * 8, o);
      }
    }
    __syncthreads();                       // g planes, row sums and stage s may be overwritten
  }
  if (COLG) {
    float* s_red = reinterpret_cast<float*>(s_glo);          // [3][D]
    for (int i = threadIdx.x; i < 3 * D; i += LNP_THREADS) s_red[i] = 0.f;
    __syncthreads();
    if (rg < RG) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        atomicAdd(&s_red[cc * 8 + j], a0[j]); atomicAdd(&s_red[D + cc * 8 + j], a1[j]); atomicAdd(&s_red[2 * D + cc * 8 + j], a2[j]);
      }
    }
    __syncthreads();
    for (int c = threadIdx.x; c < D; c += LNP_THREADS) {
      if (dgamma) { atomicAdd(&dgamma[c], s_red[c]); atomicAdd(&dbeta[c], s_red[D + c]); }
      if (dlb) atomicAdd(&dlb[c], s_red[2 * D + c]);
    }
  }
}

// ---- forward: y = act(LN(x)), (mean, rstd) per row; tiles of x through a three-stage ring, warp per row, row chunk in registers ---------
template <typename T, int LPR, int CPL>
__global__ void __launch_bounds__(LNP_THREADS, LNP_BLOCKS_PER_SM) ln_fwd_pipe_kernel(const T* __restrict__ x, long long ldx, const float* __restrict__ gamma,
                                                                     const float* __restrict__ beta, float eps, int act, T* __restrict__ y, long long ldy,
                                                                     float* __restrict__ stats, long long M) {
  using G = LnpGeom<T, LPR, CPL>;
  constexpr int D = G::D, R = G::FR;
  constexpr uint32_t TILE = (uint32_t)G::FTILE;
  constexpr bool FAST = sizeof(T) == 2;
  genrl::pdl_sync();
  extern __shared__ __align__(128) unsigned char lnp_smem[];
  T* const s_x0 = reinterpret_cast<T*>(lnp_smem);                                         // stage s at s * TILE
  float* const s_gamma = reinterpret_cast<float*>(lnp_smem + 3 * TILE * sizeof(T));
  float* const s_beta = s_gamma + D;
  uint64_t* const full = reinterpret_cast<uint64_t*>(s_beta + D);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long ntiles = (M + R - 1) / R;
  if (threadIdx.x == 0) { mbar_init(&full[0], 1); mbar_init(&full[1], 1); mbar_init(&full[2], 1); fence_barrier_init(); }
  for (int i = threadIdx.x; i < D; i += LNP_THREADS) {
    const int pl = ((i >> 2) & 1) * (G::CPR * 4) + (i >> 3) * 4 + (i & 3);
    s_gamma[pl] = gamma[i]; s_beta[pl] = beta[i];
  }
  __syncthreads();
  long long t = blockIdx.x;
  if (warp == 0) {
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      const long long tt = t + (long long)s * gridDim.x;
      if (tt < ntiles) lnp_issue<T>(x, ldx, nullptr, 0, s_x0 + (uint32_t)s * TILE, nullptr, tt * R, (int)min((long long)R, M - tt * R), D, &full[s], lane);
    }
  }
  const int sub = lane % LPR, rwi = lane / LPR;
  uint32_t s = 0, ph = 0;
  for (; t < ntiles; t += gridDim.x) {
    const long long r0 = t * R;
    const int nr = (int)min((long long)R, M - r0);
    const long long t2 = t + 2LL * gridDim.x;
    if (warp == 0 && t2 < ntiles) {
      const uint32_t s2 = s == 0 ? 2u : s - 1u;       // the stage consumed in the previous iteration (all warps passed its trailing barrier)
      lnp_issue<T>(x, ldx, nullptr, 0, s_x0 + s2 * TILE, nullptr, t2 * R, (int)min((long long)R, M - t2 * R), D, &full[s2], lane);
    }
    mbar_wait(&full[s], ph);
    const T* tx = s_x0 + s * TILE;
#pragma unroll 1
    for (int rb = warp * G::RPW; rb < nr; rb += G::SLOTS) {
      const int r = rb + rwi;
      const bool valid = r < nr;
      float v[CPL][8];
      float sum = 0.f, sq = 0.f;
#pragma unroll
      for (int i = 0; i < CPL; ++i) {
        if (valid) V8<T>::load(tx + (uint32_t)r * D + (uint32_t)(sub + i * LPR) * 8u, v[i]);
        else {
#pragma unroll
          for (int j = 0; j < 8; ++j) v[i][j] = 0.f;
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) sum += v[i][j];
      }
      const float mean = group_sum(sum, LPR) / D;
#pragma unroll
      for (int i = 0; i < CPL; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) { const float d = v[i][j] - mean; sq = fmaf(d, d, sq); }
      const float rstd = 1.0f / sqrtf(group_sum(sq, LPR) / D + eps);
      if (valid) {
#pragma unroll
        for (int i = 0; i < CPL; ++i) {
          const uint32_t c8 = (uint32_t)(sub + i * LPR);
          float g8[8], b8[8], o[8];
          lnp_ld_planes(s_gamma, G::CPR, c8, g8); lnp_ld_planes(s_beta, G::CPR, c8, b8);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float a = (v[i][j] - mean) * rstd * g8[j] + b8[j];
            o[j] = act == GENRL_ACT_SILU ? silu_t<FAST>(a) : a;
          }
          V8<T>::store(y + (r0 + r) * ldy + c8 * 8u, o);
        }
        if (stats && sub == 0) { stats[(r0 + r) * 2] = mean; stats[(r0 + r) * 2 + 1] = rstd; }
      }
    }
    __syncthreads();
    if (++s == 3) { s = 0; ph ^= 1u; }
  }
}

// host side: (LPR, CPL) of a row width, or false when the width has no instantiation (the register kernels of rowops_fast.cuh run)
static inline bool lnp_geom(int D, int* lpr, int* cpl) {
  switch (D) {
    case 1024: *lpr = 32; *cpl = 4; return true;
    case 512: *lpr = 32; *cpl = 2; return true;
    case 256: *lpr = 32; *cpl = 1; return true;
    case 384: *lpr = 16; *cpl = 3; return true;
    case 192: *lpr = 8; *cpl = 3; return true;
    case 96: *lpr = 4; *cpl = 3; return true;
    case 48: *lpr = 2; *cpl = 3; return true;
    default: return false;
  }
}
#define LNP_DISPATCH(D, ...)                                                        \
  switch (D) {                                                                      \
    case 1024: { constexpr int LPR = 32, CPL = 4; __VA_ARGS__; } break;             \
    case 512: { constexpr int LPR = 32, CPL = 2; __VA_ARGS__; } break;              \
    case 256: { constexpr int LPR = 32, CPL = 1; __VA_ARGS__; } break;              \
    case 384: { constexpr int LPR = 16, CPL = 3; __VA_ARGS__; } break;              \
    case 192: { constexpr int LPR = 8, CPL = 3; __VA_ARGS__; } break;               \
    case 96: { constexpr int LPR = 4, CPL = 3; __VA_ARGS__; } break;                \
    default: { constexpr int LPR = 2, CPL = 3; __VA_ARGS__; } break;                \
  }

}  // namespace genrl
