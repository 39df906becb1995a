// Weight-STREAMING RSSM time loop for hidden sizes whose recurrent matrix cannot stay resident in shared memory
// (BASELINE configs[4]: deter = hidden = 4096, 16 sequences per GPU; reference dreamer_utils.py:362-371 static_scan over
//  obs_step :432-440 -> img_step :459-473 -> GRUCell :771-785).
//
// The resident-weight scan (rssm_scan.cu) keeps 24 rows of W_h per CTA in shared memory: 3 D^2 bf16 bytes over the grid, 6.3 MB at
// D = 1024 but 100 MB at D = 4096.  Here the 100 MB are streamed once per time step — this is the regime the HBM roofline of the
// north star is about: per step 3 D^2 * 2 bytes of weights against 2 * B * 3 D^2 FLOP (B = 16: 16 FLOP per byte, far below the
// tensor pipe's balance point), so the kernel is built as a bandwidth pipeline:
//   * G = D/32 CTAs (128 at D = 4096), one cooperative launch loops T.  CTA c owns hidden units [32c, 32c+32): 96 rows of W_h
//     (reset / candidate / update gates) = six m16 tiles;
//   * swap-AB mma.sync m16n8k16: the WEIGHT rows are the M operand and go global -> registers -> tensor core with 16-byte
//     __ldcg loads (no shared-memory staging: every byte is used exactly once per step), the <= 16 batch rows of m_t * h_{t-1} are the
//     N operand (two n8 tiles) and sit in shared memory as bf16 (128 KB at D = 4096);
//   * 24 warps = 6 m-tiles x 4 K quarters, 8 independent 16-byte loads in flight per thread (96 KB per SM): enough bytes in flight
//     for the HBM latency without any software pipeline beyond the unrolled prefetch;
//   * LayerNorm(3D) statistics via per-CTA partials + grid barrier (as in rssm_scan.cu), gates for the owned units, h_t written
//     to L2 as bf16 (already multiplied by the next step's reset mask), second grid barrier, every CTA reloads the 16 x D operand.
// The backward (second kernel below) streams the transposed matrix the same way; the weight gradient is one GEMM over all B*T rows afterwards.
#include "common.cuh"
#include <stdlib.h>
#include <algorithm>

namespace genrl {

constexpr int SS_U = 32;                    // hidden units per CTA
constexpr int SS_R = 3 * SS_U;              // rows of W_h per CTA
constexpr int SS_MT = SS_R / 16;            // m16 tiles
constexpr int SS_KP = 4;                    // K parts
constexpr int SS_WARPS = SS_MT * SS_KP;
constexpr int SS_THREADS = SS_WARPS * 32;   // 768
constexpr int SS_NB = 16;                   // batch rows (two n8 tiles)
constexpr int SS_HPAD = 32;                 // bf16 padding of an operand row: stride = 64 B mod 128 B -> conflict-free LDS.128
constexpr int SS_RLD = SS_R + 1;            // row stride of the split-K partials
constexpr int SS_MAX_CTAS = 160;
constexpr int SS_FLAG_BYTES = 256;

__device__ __forceinline__ unsigned ss_ld_relaxed(const unsigned* p) {
  unsigned v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// grid barrier of rssm_scan.cu (one release reduction per CTA on a monotonically growing counter, thread 0 polls)
__device__ __forceinline__ void ss_grid_barrier(unsigned* counter, unsigned& epoch, int nblocks) {
  __syncthreads();
  epoch += 1;
  if (threadIdx.x == 0) {
    asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(counter), "r"(1u) : "memory");
    const unsigned target = epoch * (unsigned)nblocks;
    while (ss_ld_relaxed(counter) < target) { }
    asm volatile("fence.acq_rel.gpu;" ::: "memory");
  }
  __syncthreads();
}

__device__ __forceinline__ void ss_mma(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// 16-byte weight load with an L2 eviction policy: the first `res` K blocks of every row are kept with evict_last (they stay in the 126 MB L2
// from one time step to the next), the rest is streamed with evict_first so that it cannot push the resident part out
__device__ __forceinline__ uint4 ss_ld_policy(const uint4* p, uint64_t pol) {
  uint4 v;
  asm volatile("ld.global.L2::cache_hint.v4.u32 {%0, %1, %2, %3}, [%4], %5;" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p), "l"(pol));
  return v;
}

struct StreamFwdParams {
  int B, T, D, G;
  int res_blk;                              // K blocks (of 32) per K part loaded with the evict_last policy
  const float* gx;                          // [B*T, 3D] x-part of the GRU pre-activation (row n = b*T + t)
  const __nv_bfloat16* Wh; long long ldw;   // W_gru[:, Hd:] in bf16: 3D rows, row stride ldw
  const float* gamma; const float* beta;
  const float* mask;                        // [B*T] 1 - is_first, or null
  const float* h0;                          // [B, D] or null
  float* deter;                             // [B*T, D]
  __nv_bfloat16* hprev; long long ld_hp;    // [B*T, D] m_t * h_{t-1} (the recurrent operand; also the W_h wgrad operand of the backward)
  __nv_bfloat16* z; long long ld_z;         // [B*T, 3D] pre-LayerNorm activations (saved for the backward)
  float* stats;                             // [B*T, 2] (mean, rstd)
  float2* part;                             // workspace [B][G]
  unsigned* barrier;
};

__global__ void __launch_bounds__(SS_THREADS, 1) rssm_scan_stream_fwd_kernel(StreamFwdParams p) {
  genrl::pdl_sync();
  extern __shared__ __align__(16) unsigned char ss_smem[];
  const int D = p.D, D3 = 3 * p.D, B = p.B, T = p.T, G = p.G;
  const int HLD = D + SS_HPAD;
  __nv_bfloat16* hsm = reinterpret_cast<__nv_bfloat16*>(ss_smem);                  // [16][HLD] m_t * h_{t-1}
  float* red = reinterpret_cast<float*>(ss_smem + (size_t)SS_NB * HLD * 2);       // [KP][16][RLD]
  float* zs = red + SS_KP * SS_NB * SS_RLD;                                         // [16][RLD]
  float* st = zs + SS_NB * SS_RLD;                                                  // [16][2]
  float* hs = st + SS_NB * 2;                                                       // [16][U] own units of h_{t-1} (fp32, unmasked)
  float* ms = hs + SS_NB * SS_U;                                                    // [2][16]
  float* gb = ms + 2 * SS_NB;                                                       // [2][R]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int j0 = blockIdx.x * SS_U;
  unsigned epoch = 0;

  for (int r = tid; r < SS_R; r += SS_THREADS) {
    const int gate = r / SS_U, u = r % SS_U;
    gb[r] = p.gamma[gate * D + j0 + u];
    gb[SS_R + r] = p.beta[gate * D + j0 + u];
  }
  for (int i = tid; i < SS_NB * HLD; i += SS_THREADS) hsm[i] = __float2bfloat16_rn(0.f);        // rows >= B and the padding stay zero
  for (int i = tid; i < SS_NB * SS_U; i += SS_THREADS) {
    const int b = i / SS_U, u = i % SS_U;
    float v = 0.f;
    if (b < B) {
      v = p.h0 ? p.h0[(size_t)b * D + j0 + u] : 0.f;
      const float m = p.mask ? p.mask[(size_t)b * T] : 1.f;
      p.hprev[(size_t)(b * (long long)T) * p.ld_hp + j0 + u] = __float2bfloat16_rn(v * m);
    }
    hs[i] = v;
  }
  ss_grid_barrier(p.barrier, epoch, G);

  const int mt = warp % SS_MT, kp = warp / SS_MT, g = lane >> 2, q = lane & 3;
  const int KQ = D / SS_KP;
  // global rows of this lane's two weight rows (local rows mt*16 + g and + 8; local row r <-> W_h[(r / U) * D + j0 + r % U])
  const int r_lo = mt * 16 + g, r_hi = r_lo + 8;
  const uint4* w_lo = reinterpret_cast<const uint4*>(p.Wh + (size_t)((r_lo / SS_U) * D + j0 + (r_lo % SS_U)) * p.ldw + kp * KQ + q * 8);
  const uint4* w_hi = reinterpret_cast<const uint4*>(p.Wh + (size_t)((r_hi / SS_U) * D + j0 + (r_hi % SS_U)) * p.ldw + kp * KQ + q * 8);
  const __nv_bfloat16* hrow = hsm + (size_t)g * HLD + kp * KQ + q * 8;
  const int nblk = KQ >> 5;
  uint64_t pol_keep, pol_stream;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol_keep));
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol_stream));

  for (int t = 0; t < T; ++t) {
    // ---- operand: all batch rows of m_t * h_{t-1} (bf16, written by every CTA in the previous step) -> shared memory ----
    {
      const int cpr = D >> 3;                                       // 16-byte chunks per row
      for (int i = tid; i < B * cpr; i += SS_THREADS) {
        const int b = i / cpr, c = i - b * cpr;
        const uint4 v = __ldcg(reinterpret_cast<const uint4*>(p.hprev + ((size_t)b * T + t) * p.ld_hp) + c);
        *reinterpret_cast<uint4*>(hsm + (size_t)b * HLD + c * 8) = v;
      }
    }
    for (int b = tid; b < SS_NB; b += SS_THREADS) {
      ms[b] = (p.mask && b < B) ? p.mask[(size_t)b * T + t] : 1.f;
      ms[SS_NB + b] = (p.mask && b < B && t + 1 < T) ? p.mask[(size_t)b * T + t + 1] : 1.f;
    }
    // x-part for the combine step: issued before the product so its latency hides behind the weight stream
    float gxv[2];
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const int i = tid + k * SS_THREADS, b = i / SS_R, r = i % SS_R;
      gxv[k] = (i < SS_NB * SS_R && b < B) ? p.gx[((size_t)b * T + t) * D3 + (r / SS_U) * D + j0 + (r % SS_U)] : 0.f;
    }
    __syncthreads();
    // ---- z_h = W_h (m_t h_{t-1}): weight rows streamed from global memory straight into the A fragments ----
    float acc[2][4];
#pragma unroll
    for (int i = 0; i < 2; ++i) { acc[i][0] = 0.f; acc[i][1] = 0.f; acc[i][2] = 0.f; acc[i][3] = 0.f; }
    constexpr int PF = 4;
    for (int blk0 = 0; blk0 < nblk; blk0 += PF) {
      uint4 lo[PF], hi[PF];
#pragma unroll
      for (int i = 0; i < PF; ++i) {
        if (blk0 + i < nblk) {
          const uint64_t pol = (blk0 + i) < p.res_blk ? pol_keep : pol_stream;
          lo[i] = ss_ld_policy(w_lo + (blk0 + i) * 4, pol); hi[i] = ss_ld_policy(w_hi + (blk0 + i) * 4, pol);
        }
      }
#pragma unroll
      for (int i = 0; i < PF; ++i) {
        if (blk0 + i >= nblk) break;
#pragma unroll
        for (int nt = 0; nt < 2; ++nt) {
          const uint4 h = *reinterpret_cast<const uint4*>(hrow + (size_t)nt * 8 * HLD + (blk0 + i) * 32);
          ss_mma(acc[nt], lo[i].x, hi[i].x, lo[i].y, hi[i].y, h.x, h.y);
          ss_mma(acc[nt], lo[i].z, hi[i].z, lo[i].w, hi[i].w, h.z, h.w);
        }
      }
    }
    // accumulator (weight row, batch column) -> red[kp][batch][row]
#pragma unroll
    for (int nt = 0; nt < 2; ++nt) {
      float* rb0 = red + ((size_t)kp * SS_NB + nt * 8 + q * 2) * SS_RLD;
      rb0[r_lo] = acc[nt][0]; rb0[SS_RLD + r_lo] = acc[nt][1]; rb0[r_hi] = acc[nt][2]; rb0[SS_RLD + r_hi] = acc[nt][3];
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const int i = tid + k * SS_THREADS, b = i / SS_R, r = i % SS_R;
      if (i < SS_NB * SS_R) {
        const int o = b * SS_RLD + r;
        const float zv = (red[o] + red[SS_NB * SS_RLD + o]) + (red[2 * SS_NB * SS_RLD + o] + red[3 * SS_NB * SS_RLD + o]) + gxv[k];
        zs[o] = zv;
        if (b < B) p.z[((size_t)b * T + t) * p.ld_z + (r / SS_U) * D + j0 + (r % SS_U)] = __float2bfloat16_rn(zv);
      }
    }
    __syncthreads();
    // ---- LayerNorm partial sums of this CTA's 96 values per batch row ----
    if (warp < SS_NB && warp < B) {
      float s1 = 0.f, s2 = 0.f;
#pragma unroll
      for (int k = 0; k < SS_R / 32; ++k) { const float v = zs[warp * SS_RLD + lane + 32 * k]; s1 += v; s2 = fmaf(v, v, s2); }
      s1 = warp_sum(s1); s2 = warp_sum(s2);
      if (lane == 0) p.part[(size_t)warp * G + blockIdx.x] = make_float2(s1, s2);
    }
    ss_grid_barrier(p.barrier, epoch, G);
    // ---- full-row statistics: warp b adds the G partials of row b in a fixed order ----
    if (warp < SS_NB && warp < B) {
      float s1 = 0.f, s2 = 0.f;
      constexpr int NC = SS_MAX_CTAS / 32;
      float2 v[NC];
#pragma unroll
      for (int k = 0; k < NC; ++k) { const int c = lane + 32 * k; v[k] = c < G ? __ldcg(p.part + (size_t)warp * G + c) : make_float2(0.f, 0.f); }
#pragma unroll
      for (int k = 0; k < NC; ++k) { s1 += v[k].x; s2 += v[k].y; }
      s1 = warp_sum(s1); s2 = warp_sum(s2);
      if (lane == 0) { st[warp * 2] = s1 / D3; st[warp * 2 + 1] = s2 / D3; }
    }
    __syncthreads();
    // ---- gates for the owned units ----
    for (int i = tid; i < B * SS_U; i += SS_THREADS) {
      const int b = i / SS_U, u = i % SS_U, j = j0 + u;
      const float mean = st[b * 2];
      const float rstd = 1.0f / sqrtf(fmaxf(st[b * 2 + 1] - mean * mean, 0.f) + 1e-5f);
      const size_t n = (size_t)b * T + t;
      if (blockIdx.x == 0 && u == 0) { p.stats[n * 2] = mean; p.stats[n * 2 + 1] = rstd; }
      const float* zr = zs + b * SS_RLD;
      const float pr = (zr[u] - mean) * rstd * gb[u] + gb[SS_R + u];
      const float pc = (zr[SS_U + u] - mean) * rstd * gb[SS_U + u] + gb[SS_R + SS_U + u];
      const float pu = (zr[2 * SS_U + u] - mean) * rstd * gb[2 * SS_U + u] + gb[SS_R + 2 * SS_U + u];
      const float r = sigmoidf_(pr), cand = tanhf(r * pc), uu = sigmoidf_(pu - 1.0f);
      const float hn = uu * cand + (1.0f - uu) * (hs[i] * ms[b]);
      hs[i] = hn;
      p.deter[n * D + j] = hn;
      if (t + 1 < T) p.hprev[(n + 1) * p.ld_hp + j] = __float2bfloat16_rn(hn * ms[SS_NB + b]);
    }
    ss_grid_barrier(p.barrier, epoch, G);
  }
}

// =================================================================================================
// Backward of the same time loop, weights streamed once per step as well.  CTA c owns the same 32 hidden units.  Per step t = T-1..0:
//   A. gate + LayerNorm backward for the owned units (dh_t = external gradient + carry): the three gate gradients times gamma, x-hat, the
//      direct path (1 - u) dh; per-CTA partial sums of the LayerNorm backward's two row reductions -> grid barrier -> dz_t (bf16, global:
//      all CTAs need all of it, and the deferred W_h / W_x weight gradients read it afterwards);
//   B. carry[b, k] = m_t ((1 - u) dh + sum_j dz_t[b, j] W_h[j, k]) for the owned k: the TRANSPOSED recurrent matrix (bf16 [D, 3D], built
//      once per backward by the caller) is the streamed M operand (32 rows = two m16 tiles, K = 3D split over 16 warps), dz_t rows are the
//      N operand, read as fragments straight from L2.
// =================================================================================================
constexpr int SB_WARPS = 16;
constexpr int SB_THREADS = SB_WARPS * 32;   // 512
constexpr int SB_ULD = SS_U + 1;

struct StreamBwdParams {
  int B, T, D, G;
  int res_blk;
  const float* dh_ext;                          // [B*T, D]
  const __nv_bfloat16* z; long long ld_z;       // [B*T, 3D]
  const float* stats;                           // [B*T, 2]
  const __nv_bfloat16* WhT; long long ldwt;     // [D, 3D]: WhT[k][j] = W_h[j][k]
  const float* gamma; const float* beta;
  const float* mask;                            // [B*T] or null
  const float* deter;                           // [B*T, D]
  const float* h0;                              // [B, D] or null
  __nv_bfloat16* dz; long long ld_dz;           // [B*T, 3D] out
  float* dh0;                                   // [B, D] or null
  float* dgamma; float* dbeta;                  // [3D] accumulated (+=), or null
  float2* part;
  unsigned* barrier;
};

__global__ void __launch_bounds__(SB_THREADS, 1) rssm_scan_stream_bwd_kernel(StreamBwdParams p) {
  genrl::pdl_sync();
  extern __shared__ __align__(16) unsigned char ss_smem[];
  const int D = p.D, D3 = 3 * p.D, B = p.B, T = p.T, G = p.G;
  float* red = reinterpret_cast<float*>(ss_smem);              // [16 K parts][16 b][ULD]
  float* gs = red + SB_WARPS * SS_NB * SB_ULD;                 // [16][RLD] gamma * gate gradient
  float* xs = gs + SS_NB * SS_RLD;                             // [16][RLD] normalised z
  float* carry = xs + SS_NB * SS_RLD;                          // [16][U]
  float* keep = carry + SS_NB * SS_U;                          // [16][U] (1 - u) dh
  float* st = keep + SS_NB * SS_U;                             // [16][2] row sums / 3D
  float* rs = st + SS_NB * 2;                                  // [16] rstd, [16] mask
  float* dgl = rs + 2 * SS_NB;                                 // [2][R]
  float* gb = dgl + 2 * SS_R;                                  // [2][R]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int j0 = blockIdx.x * SS_U;
  unsigned epoch = 0;
  for (int r = tid; r < SS_R; r += SB_THREADS) {
    const int gate = r / SS_U, u = r % SS_U;
    gb[r] = p.gamma[gate * D + j0 + u];
    gb[SS_R + r] = p.beta[gate * D + j0 + u];
    dgl[r] = 0.f; dgl[SS_R + r] = 0.f;
  }
  for (int i = tid; i < SS_NB * SS_U; i += SB_THREADS) { carry[i] = 0.f; keep[i] = 0.f; }
  for (int i = tid; i < SS_NB * SS_RLD; i += SB_THREADS) { gs[i] = 0.f; xs[i] = 0.f; }
  __syncthreads();

  const int g = lane >> 2, q = lane & 3;
  const int KP = D3 / SB_WARPS;                               // K range of a warp (multiple of 32)
  const int nblk = KP >> 5;
  const uint4* w_r[4];                                        // rows g, g + 8, 16 + g, 24 + g of the owned slice of W_h^T
#pragma unroll
  for (int i = 0; i < 4; ++i) w_r[i] = reinterpret_cast<const uint4*>(p.WhT + (size_t)(j0 + i * 8 + g) * p.ldwt + warp * KP + q * 8);
  uint64_t pol_keep, pol_stream;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol_keep));
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol_stream));
  const uint4 zero4 = make_uint4(0u, 0u, 0u, 0u);

  for (int t = T - 1; t >= 0; --t) {
    // ---- A. gate + LayerNorm backward of the owned units ----
    {
      const int b = tid / SS_U, u = tid % SS_U, j = j0 + u;     // 512 threads = 16 rows x 32 units
      if (b < B) {
        const size_t n = (size_t)b * T + t;
        const float mean = p.stats[n * 2], rstd = p.stats[n * 2 + 1];
        const float m = p.mask ? p.mask[n] : 1.f;
        const float hp = (t == 0 ? (p.h0 ? p.h0[(size_t)b * D + j] : 0.f) : p.deter[(n - 1) * D + j]) * m;
        const float dh = p.dh_ext[n * D + j] + carry[tid];
        const __nv_bfloat16* zr = p.z + n * p.ld_z;
        const float xr = (__bfloat162float(zr[j]) - mean) * rstd, xc = (__bfloat162float(zr[D + j]) - mean) * rstd,
                    xu = (__bfloat162float(zr[2 * D + j]) - mean) * rstd;
        const float pr = xr * gb[u] + gb[SS_R + u], pc = xc * gb[SS_U + u] + gb[SS_R + SS_U + u], pu = xu * gb[2 * SS_U + u] + gb[SS_R + 2 * SS_U + u];
        const float r = sigmoidf_(pr), cand = tanhf(r * pc), uu = sigmoidf_(pu - 1.0f);
        const float drc = dh * uu * (1.0f - cand * cand);
        const float gpr = drc * pc * r * (1.0f - r), gpc = drc * r, gpu = dh * (cand - hp) * uu * (1.0f - uu);
        float* gr = gs + b * SS_RLD; float* xr_ = xs + b * SS_RLD;
        gr[u] = gpr * gb[u]; gr[SS_U + u] = gpc * gb[SS_U + u]; gr[2 * SS_U + u] = gpu * gb[2 * SS_U + u];
        xr_[u] = xr; xr_[SS_U + u] = xc; xr_[2 * SS_U + u] = xu;
        keep[tid] = dh * (1.0f - uu);
        if (u == 0) { rs[b] = rstd; rs[SS_NB + b] = m; }
        if (p.dgamma) {
          atomicAdd(&dgl[u], gpr * xr); atomicAdd(&dgl[SS_U + u], gpc * xc); atomicAdd(&dgl[2 * SS_U + u], gpu * xu);
          atomicAdd(&dgl[SS_R + u], gpr); atomicAdd(&dgl[SS_R + SS_U + u], gpc); atomicAdd(&dgl[SS_R + 2 * SS_U + u], gpu);
        }
      }
    }
    __syncthreads();
    if (warp < B) {
      float s1 = 0.f, s2 = 0.f;
#pragma unroll
      for (int k = 0; k < SS_R / 32; ++k) { const float gv = gs[warp * SS_RLD + lane + 32 * k]; s1 += gv; s2 = fmaf(gv, xs[warp * SS_RLD + lane + 32 * k], s2); }
      s1 = warp_sum(s1); s2 = warp_sum(s2);
      if (lane == 0) p.part[(size_t)warp * G + blockIdx.x] = make_float2(s1, s2);
    }
    ss_grid_barrier(p.barrier, epoch, G);
    if (warp < B) {
      constexpr int NC = SS_MAX_CTAS / 32;
      float2 v[NC];
#pragma unroll
      for (int k = 0; k < NC; ++k) { const int c = lane + 32 * k; v[k] = c < G ? __ldcg(p.part + (size_t)warp * G + c) : make_float2(0.f, 0.f); }
      float s1 = 0.f, s2 = 0.f;
#pragma unroll
      for (int k = 0; k < NC; ++k) { s1 += v[k].x; s2 += v[k].y; }
      s1 = warp_sum(s1); s2 = warp_sum(s2);
      if (lane == 0) { st[warp * 2] = s1 / D3; st[warp * 2 + 1] = s2 / D3; }
    }
    __syncthreads();
    for (int i = tid; i < SS_NB * SS_R; i += SB_THREADS) {
      const int b = i / SS_R, r = i % SS_R;
      if (b < B) {
        const float dzv = rs[b] * (gs[b * SS_RLD + r] - st[b * 2] - xs[b * SS_RLD + r] * st[b * 2 + 1]);
        p.dz[((size_t)b * T + t) * p.ld_dz + (r / SS_U) * D + j0 + (r % SS_U)] = __float2bfloat16_rn(dzv);
      }
    }
    ss_grid_barrier(p.barrier, epoch, G);
    // ---- B. carry = m_t ((1 - u) dh + dz_t W_h[:, owned units]) ----
    {
      float acc[2][2][4];
#pragma unroll
      for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int c = 0; c < 2; ++c) { acc[a][c][0] = 0.f; acc[a][c][1] = 0.f; acc[a][c][2] = 0.f; acc[a][c][3] = 0.f; }
      const uint4* d_r[2];
#pragma unroll
      for (int nt = 0; nt < 2; ++nt) {
        const int b = nt * 8 + g;
        d_r[nt] = b < B ? reinterpret_cast<const uint4*>(p.dz + ((size_t)b * T + t) * p.ld_dz + warp * KP + q * 8) : nullptr;
      }
      constexpr int PF = 3;
      for (int blk0 = 0; blk0 < nblk; blk0 += PF) {
        uint4 w[PF][4], dv[PF][2];
#pragma unroll
        for (int i = 0; i < PF; ++i) {
          if (blk0 + i < nblk) {
            const uint64_t pol = (blk0 + i) < p.res_blk ? pol_keep : pol_stream;
#pragma unroll
            for (int k = 0; k < 4; ++k) w[i][k] = ss_ld_policy(w_r[k] + (blk0 + i) * 4, pol);
#pragma unroll
            for (int nt = 0; nt < 2; ++nt) dv[i][nt] = d_r[nt] ? __ldcg(d_r[nt] + (blk0 + i) * 4) : zero4;
          }
        }
#pragma unroll
        for (int i = 0; i < PF; ++i) {
          if (blk0 + i >= nblk) break;
#pragma unroll
          for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int nt = 0; nt < 2; ++nt) {
              ss_mma(acc[mt][nt], w[i][2 * mt].x, w[i][2 * mt + 1].x, w[i][2 * mt].y, w[i][2 * mt + 1].y, dv[i][nt].x, dv[i][nt].y);
              ss_mma(acc[mt][nt], w[i][2 * mt].z, w[i][2 * mt + 1].z, w[i][2 * mt].w, w[i][2 * mt + 1].w, dv[i][nt].z, dv[i][nt].w);
            }
        }
      }
      // accumulator (unit row, batch column) -> red[warp][batch][unit]
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = 0; nt < 2; ++nt) {
          float* rb0 = red + ((size_t)warp * SS_NB + nt * 8 + q * 2) * SB_ULD + mt * 16 + g;
          rb0[0] = acc[mt][nt][0]; rb0[SB_ULD] = acc[mt][nt][1]; rb0[8] = acc[mt][nt][2]; rb0[SB_ULD + 8] = acc[mt][nt][3];
        }
    }
    __syncthreads();
    {
      const int b = tid / SS_U, u = tid % SS_U;
      float sum = keep[tid];
#pragma unroll
      for (int k = 0; k < SB_WARPS; ++k) sum += red[((size_t)k * SS_NB + b) * SB_ULD + u];
      const float cv = b < B ? sum * rs[SS_NB + b] : 0.f;
      carry[tid] = cv;
      if (t == 0 && p.dh0 && b < B) p.dh0[(size_t)b * D + j0 + u] = cv;
    }
    __syncthreads();
  }
  if (p.dgamma) {
    for (int r = tid; r < SS_R; r += SB_THREADS) {
      const int gate = r / SS_U, u = r % SS_U;
      atomicAdd(&p.dgamma[gate * D + j0 + u], dgl[r]);
      atomicAdd(&p.dbeta[gate * D + j0 + u], dgl[SS_R + r]);
    }
  }
}

}  // namespace genrl

using namespace genrl;

// 1 when the streaming scan can run this shape on this device (bf16 activations): D a multiple of 128 with D/32 CTAs co-resident,
// at most 16 sequences, the operand rows within shared memory
extern "C" int genrl_rssm_scan_stream_supported(int B, int D) {
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (sms <= 0) sms = 148;
  const size_t smem = (size_t)SS_NB * (D + SS_HPAD) * 2 + (size_t)(SS_KP + 1) * SS_NB * SS_RLD * 4 + 4096;
  return (B >= 1 && B <= SS_NB && D % 128 == 0 && D / SS_U <= sms && D / SS_U <= SS_MAX_CTAS && smem <= 227 * 1024) ? 1 : 0;
}

// deter[b,t] = GRU(gx[b,t] + W_h (m[b,t] deter[b,t-1])), t = 0..T-1, weights streamed per step.  Wh: bf16 [3D, D] (row stride ldw,
// 16-byte aligned rows); hprev / z: bf16 outputs [B*T, D] / [B*T, 3D]; workspace: genrl_rssm_scan_workspace_bytes(B, T, D).
extern "C" int genrl_rssm_scan_stream_fwd(void* stream, int B, int T, int D, const float* gx, const void* Wh, long long ldw, const float* gamma,
                                          const float* beta, const float* mask, const float* h0, float* deter, void* hprev, long long ld_hp,
                                          void* z, long long ld_z, float* stats, void* workspace) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  GENRL_REQUIRE(genrl_rssm_scan_stream_supported(B, D) == 1, "rssm_scan_stream_fwd: unsupported shape B=%d D=%d", B, D);
  GENRL_REQUIRE(T >= 1 && ldw % 8 == 0 && ld_hp % 8 == 0 && ((reinterpret_cast<uintptr_t>(Wh) | reinterpret_cast<uintptr_t>(hprev)) & 15) == 0,
                "rssm_scan_stream_fwd: W_h / hprev rows must be 16-byte aligned");
  StreamFwdParams p;
  p.B = B; p.T = T; p.D = D; p.G = D / SS_U;
  {
    // share of the weight bytes kept in L2 between steps (GENRL_SCAN_STREAM_RES = percent; default min(100, 75 MB / |W_h|): measured on B200 at
    // D = 4096, us per step for 0 / 30 / 50 / 62 / 75 / 100 percent: 26.0 / 23.9 / 23.4 / 22.6 / 22.0 / 22.9)
    const double wbytes = 3.0 * D * D * 2.0;
    int pct = (int)std::min(100.0, 100.0 * 75e6 / wbytes);
    if (const char* e = getenv("GENRL_SCAN_STREAM_RES")) pct = atoi(e);
    p.res_blk = (int)((long long)(D / SS_KP / 32) * pct / 100);
  }
  p.gx = gx; p.Wh = reinterpret_cast<const __nv_bfloat16*>(Wh); p.ldw = ldw; p.gamma = gamma; p.beta = beta; p.mask = mask; p.h0 = h0;
  p.deter = deter; p.hprev = reinterpret_cast<__nv_bfloat16*>(hprev); p.ld_hp = ld_hp; p.z = reinterpret_cast<__nv_bfloat16*>(z); p.ld_z = ld_z;
  p.stats = stats;
  p.barrier = reinterpret_cast<unsigned*>(workspace);
  p.part = reinterpret_cast<float2*>(reinterpret_cast<char*>(workspace) + SS_FLAG_BYTES);
  const size_t smem = (size_t)SS_NB * (D + SS_HPAD) * 2 + ((size_t)(SS_KP + 1) * SS_NB * SS_RLD + SS_NB * 2 + SS_NB * SS_U + 2 * SS_NB + 2 * SS_R) * sizeof(float);
  cudaError_t e = cudaFuncSetAttribute(rssm_scan_stream_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  GENRL_REQUIRE(e == cudaSuccess, "rssm_scan_stream_fwd: smem attribute: %s", cudaGetErrorString(e));
  cudaMemsetAsync(workspace, 0, SS_FLAG_BYTES, st);
  void* args[] = {&p};
  e = cudaLaunchCooperativeKernel((void*)rssm_scan_stream_fwd_kernel, dim3(p.G), dim3(SS_THREADS), args, smem, st);
  GENRL_REQUIRE(e == cudaSuccess, "rssm_scan_stream_fwd: cooperative launch failed: %s", cudaGetErrorString(e));
  return check_launch("rssm_scan_stream_fwd");
}

// Backward of genrl_rssm_scan_stream_fwd: dz (bf16 [B*T, 3D]) = gradient of the pre-LayerNorm GRU activations, dh0, dgamma / dbeta (+=).
// WhT: the TRANSPOSED recurrent matrix in bf16, [D, 3D] with row stride ldwt.  Requires D % 512 == 0 (returns 2 otherwise: the caller then
// walks the steps with genrl_gru_bwd + genrl_gemm).
extern "C" int genrl_rssm_scan_stream_bwd(void* stream, int B, int T, int D, const float* dh_ext, const void* z, long long ld_z, const float* stats,
                                          const void* WhT, long long ldwt, const float* gamma, const float* beta, const float* mask,
                                          const float* deter, const float* h0, void* dz, long long ld_dz, float* dh0, float* dgamma, float* dbeta,
                                          void* workspace) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (genrl_rssm_scan_stream_supported(B, D) != 1 || (3 * D / SB_WARPS) % 32 != 0) return 2;
  GENRL_REQUIRE(T >= 1 && ldwt % 8 == 0 && ld_dz % 8 == 0 && ((reinterpret_cast<uintptr_t>(WhT) | reinterpret_cast<uintptr_t>(dz)) & 15) == 0,
                "rssm_scan_stream_bwd: W_h^T / dz rows must be 16-byte aligned");
  StreamBwdParams p;
  p.B = B; p.T = T; p.D = D; p.G = D / SS_U;
  {
    const double wbytes = 3.0 * D * D * 2.0;
    int pct = (int)std::min(100.0, 100.0 * 75e6 / wbytes);
    if (const char* e = getenv("GENRL_SCAN_STREAM_RES")) pct = atoi(e);
    p.res_blk = (int)((long long)(3 * D / SB_WARPS / 32) * pct / 100);
  }
  p.dh_ext = dh_ext; p.z = reinterpret_cast<const __nv_bfloat16*>(z); p.ld_z = ld_z; p.stats = stats;
  p.WhT = reinterpret_cast<const __nv_bfloat16*>(WhT); p.ldwt = ldwt; p.gamma = gamma; p.beta = beta; p.mask = mask; p.deter = deter; p.h0 = h0;
  p.dz = reinterpret_cast<__nv_bfloat16*>(dz); p.ld_dz = ld_dz; p.dh0 = dh0; p.dgamma = dgamma; p.dbeta = dbeta;
  p.barrier = reinterpret_cast<unsigned*>(workspace);
  p.part = reinterpret_cast<float2*>(reinterpret_cast<char*>(workspace) + SS_FLAG_BYTES);
  const size_t smem = ((size_t)SB_WARPS * SS_NB * SB_ULD + 2 * SS_NB * SS_RLD + 2 * SS_NB * SS_U + SS_NB * 2 + 2 * SS_NB + 4 * SS_R) * sizeof(float);
  cudaError_t e = cudaFuncSetAttribute(rssm_scan_stream_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  GENRL_REQUIRE(e == cudaSuccess, "rssm_scan_stream_bwd: smem attribute: %s", cudaGetErrorString(e));
  cudaMemsetAsync(workspace, 0, SS_FLAG_BYTES, st);
  void* args[] = {&p};
  e = cudaLaunchCooperativeKernel((void*)rssm_scan_stream_bwd_kernel, dim3(p.G), dim3(SB_THREADS), args, smem, st);
  GENRL_REQUIRE(e == cudaSuccess, "rssm_scan_stream_bwd: cooperative launch failed: %s", cudaGetErrorString(e));
  return check_launch("rssm_scan_stream_bwd");
}
