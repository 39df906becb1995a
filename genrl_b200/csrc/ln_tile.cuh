// LayerNorm(+SiLU) backward for LARGE row counts, staged through shared memory (round 2).
//
// Round-1 kernels keep a whole row (and, for the parameter gradients, 96 column partials) in registers: 128-254 registers per thread,
// 8-14 resident warps per SM, 18-31 % of HBM bandwidth on the 36 000 x 1024 / 2 306 400 x 48 shapes (profiles/r01_ncu_ln_digest.txt:
// long-scoreboard stalls, 12-22 % achieved occupancy).  Here a block copies a TILE of rows of dy and x into shared memory with
// cp.async (48 KB in flight per block, three blocks per SM, no registers tied up while the data is in flight) and then makes two passes
// over the tile out of shared memory:
//   row pass     lanes-per-row mapping: c1 = mean(g), c2 = mean(g * xhat) with g = dy * act'(a) * gamma          -> smem [rows][2]
//   column pass  thread (row lane, 8-column chunk) with a FIXED chunk: recomputes g and xhat, writes dx = rstd (g - c1 - xhat c2)
//                with 16-byte coalesced stores and keeps the chunk's dgamma / dbeta / dbias partials in 24 registers
// HBM traffic is the minimum (read dy, x once, write dx once); the column partials leave the block as one fp32 atomic per column.
#pragma once
#include "rowops_fast.cuh"

namespace genrl {

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;" ::: "memory");
  asm volatile("cp.async.wait_group 0;" ::: "memory");
}

struct LnTileCfg {
  int threads;      // multiple of chunks-per-row (and of 32)
  int rows;         // tile rows, multiple of threads / chunks-per-row
  int lpr;          // lanes per row in the row pass (power of two)
  size_t smem;
};

// T: activation dtype.  D % 8 == 0.  dynamic smem: [2][R][D] T (dy, x) | gamma[D] | beta[D] | red[3][D] | rowc[R][4] (mean, rstd, c1, c2)
template <typename T>
__global__ void __launch_bounds__(256) ln_bwd_tile_kernel(const T* __restrict__ dy, long long lddy, const T* __restrict__ x, long long ldx,
                                                          const float* __restrict__ gamma, const float* __restrict__ beta,
                                                          const float* __restrict__ stats, int act, T* __restrict__ dx, long long lddx,
                                                          float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ dlb,
                                                          long long M, int D, int R, int lpr) {
  genrl::pdl_sync();
  extern __shared__ __align__(16) unsigned char ln_smem[];
  const int CPR = D >> 3;
  const int NT = blockDim.x;
  T* s_dy = reinterpret_cast<T*>(ln_smem);
  T* s_x = s_dy + (size_t)R * D;
  float* s_gamma = reinterpret_cast<float*>(s_x + (size_t)R * D);
  float* s_beta = s_gamma + D;
  float* s_red = s_beta + D;                 // [3][D]
  float* s_row = s_red + 3 * D;              // [R][4]
  const bool colg = dgamma != nullptr || dlb != nullptr;
  for (int i = threadIdx.x; i < D; i += NT) { s_gamma[i] = gamma[i]; s_beta[i] = beta[i]; }
  for (int i = threadIdx.x; i < 3 * D; i += NT) s_red[i] = 0.f;
  const int RPP = NT / CPR;                  // rows per column-pass sweep
  const int cc = threadIdx.x % CPR, rl = threadIdx.x / CPR;
  float a0[8], a1[8], a2[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { a0[j] = 0.f; a1[j] = 0.f; a2[j] = 0.f; }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = NT >> 5;
  const int sub = lane % lpr, rwi = lane / lpr, rpw = 32 / lpr;
  const long long ntiles = (M + R - 1) / R;
  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const long long r0 = tile * R;
    const int nr = (int)min((long long)R, M - r0);
    __syncthreads();                         // previous tile fully consumed
    // ---- stage the tile ----
    const int nchunks = nr * CPR;
    for (int i = threadIdx.x; i < nchunks; i += NT) {
      const int r = i / CPR, c = (i - r * CPR) << 3;
      cp_async16(s_dy + (size_t)r * D + c, dy + (r0 + r) * lddy + c);
      cp_async16(s_x + (size_t)r * D + c, x + (r0 + r) * ldx + c);
    }
    for (int r = threadIdx.x; r < nr; r += NT) { s_row[r * 4] = stats[(r0 + r) * 2]; s_row[r * 4 + 1] = stats[(r0 + r) * 2 + 1]; }
    cp_async_wait_all();
    __syncthreads();
    // ---- row pass: c1, c2 ----
    for (int rb = warp * rpw; rb < nr; rb += nwarps * rpw) {
      const int r = rb + rwi;
      float s1 = 0.f, s2 = 0.f;
      if (r < nr) {
        const float mean = s_row[r * 4], rstd = s_row[r * 4 + 1];
        for (int c8 = sub; c8 < CPR; c8 += lpr) {
          float d[8], xv[8], g8[8], b8[8];
          V8<T>::load(s_dy + (size_t)r * D + (c8 << 3), d);
          V8<T>::load(s_x + (size_t)r * D + (c8 << 3), xv);
          ldf8(s_gamma + (c8 << 3), g8);
          ldf8(s_beta + (c8 << 3), b8);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float h = (xv[j] - mean) * rstd;
            float gg = d[j];
            if (act == GENRL_ACT_SILU) gg *= silu_grad_t<(sizeof(T) == 2)>(h * g8[j] + b8[j]);
            gg *= g8[j];
            s1 += gg; s2 += gg * h;
          }
        }
      }
      s1 = group_sum(s1, lpr); s2 = group_sum(s2, lpr);
      if (r < nr && sub == 0) { s_row[r * 4 + 2] = s1 / D; s_row[r * 4 + 3] = s2 / D; }
    }
    __syncthreads();
    // ---- column pass: dx and the column partials of chunk cc ----
    if (rl < RPP) {
      float g8[8], b8[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) { g8[j] = s_gamma[(cc << 3) + j]; b8[j] = s_beta[(cc << 3) + j]; }
      for (int r = rl; r < nr; r += RPP) {
        const float mean = s_row[r * 4], rstd = s_row[r * 4 + 1], c1 = s_row[r * 4 + 2], c2 = s_row[r * 4 + 3];
        float d[8], xv[8], o[8];
        V8<T>::load(s_dy + (size_t)r * D + (cc << 3), d);
        V8<T>::load(s_x + (size_t)r * D + (cc << 3), xv);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float h = (xv[j] - mean) * rstd;
          float gg = d[j];
          if (act == GENRL_ACT_SILU) gg *= silu_grad_t<(sizeof(T) == 2)>(h * g8[j] + b8[j]);
          a0[j] += gg * h; a1[j] += gg;
          gg *= g8[j];
          o[j] = rstd * (gg - c1 - h * c2);
          a2[j] += o[j];
        }
        V8<T>::store(dx + (r0 + r) * lddx + (cc << 3), o);
      }
    }
  }
  if (colg) {
    __syncthreads();
    if (rl < RPP) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        atomicAdd(&s_red[(cc << 3) + j], a0[j]); atomicAdd(&s_red[D + (cc << 3) + j], a1[j]); atomicAdd(&s_red[2 * D + (cc << 3) + j], a2[j]);
      }
    }
    __syncthreads();
    for (int c = threadIdx.x; c < D; c += NT) {
      if (dgamma) { atomicAdd(&dgamma[c], s_red[c]); atomicAdd(&dbeta[c], s_red[D + c]); }
      if (dlb) atomicAdd(&dlb[c], s_red[2 * D + c]);
    }
  }
}

// tile geometry for D columns of `esize`-byte elements; returns false when no thread count <= 256 is a multiple of both D/8 and 32
static inline bool ln_tile_config(int D, int esize, LnTileCfg* cfg) {
  if (D % 8 != 0 || D > 1024) return false;
  const int cpr = D / 8;
  int a = cpr, b = 32;
  while (b) { int t = a % b; a = b; b = t; }
  const int lcm = cpr / a * 32;
  if (lcm > 256) return false;
  cfg->threads = lcm * (256 / lcm);
  const int rpp = cfg->threads / cpr;
  int rows = (48 * 1024) / (2 * D * esize);
  rows = rows / rpp * rpp;
  if (rows < rpp) rows = rpp;
  if (rows > 1024) rows = 1024 / rpp * rpp;
  cfg->rows = rows;
  int lpr = 1;
  while (lpr < 32 && lpr * 4 < cpr) lpr *= 2;      // <= 4 chunks per lane in the row pass
  cfg->lpr = lpr;
  cfg->smem = (size_t)2 * rows * D * esize + (size_t)5 * D * 4 + (size_t)rows * 16;
  return cfg->smem <= 100 * 1024;
}

}  // namespace genrl
