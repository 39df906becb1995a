// Persistent RSSM time-loop kernels (sm_100a): the only truly sequential part of EnsembleRSSM.observe
// (reference dreamer_utils.py:362-371 static_scan over obs_step :432-440 -> img_step :459-473 -> GRUCell :771-785).
//
// With single_obs_posterior the posterior, the GRU input projection x_t and the prior head are batched GEMMs over all
// B*T rows; what remains per step is   z_t = Gx_t + W_h (m_t * h_{t-1});  h_t = gates(LN_3D(z_t), m_t * h_{t-1}).
// One cooperative launch loops the time axis:
//   * G = D/8 CTAs, CTA c owns hidden units [8c, 8c+8): the 24 rows of W_h for its (reset,cand,update) gates live in
//     shared memory for the whole sequence (D=1024: 48 KB bf16 / 96 KB fp32 per CTA, read from HBM once);
//   * per step: the recurrent product for all B rows (bf16 mode: mma.sync tensor cores, split-K over 16 warps; fp32 mode:
//     FFMA), LayerNorm(3D) statistics exchanged through a [B][G] array of per-CTA partial sums, gates, h_t written to
//     HBM/L2 for the other CTAs; two grid barriers per step.
// The backward kernel walks t = T-1..0 with the transposed slice W_h[:, units] resident and produces dz_t (the gradient
// of the pre-LayerNorm GRU activations) for the batched wgrad/dgrad GEMMs that follow.
#include "common.cuh"
#include <stdlib.h>

namespace genrl {

constexpr int SCAN_U = 8;        // hidden units per CTA
constexpr int SCAN_RB = 4;       // batch rows processed together by one warp (FFMA path)
constexpr int SCAN_THREADS = 512;
constexpr int SCAN_WARPS = SCAN_THREADS / 32;
constexpr int SCAN_MAX_CTAS = 160;      // >= SM count of the target (148), multiple of 32
constexpr int SCAN_FLAG_BYTES = 256;    // workspace head: barrier counter

__device__ __forceinline__ unsigned ld_relaxed_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_u32(unsigned* p, unsigned v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void fence_acq_rel_gpu() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }

__device__ __forceinline__ void red_release_add_u32(unsigned* p, unsigned v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// Grid barrier: one release reduction on a monotonically growing counter per CTA, thread 0 polls it with relaxed loads
// and issues one acquire fence (1.26 us for 128 CTAs on B200 vs 1.35 us with __threadfence + atomicAdd, and 1.8-7 us for
// the per-CTA-flag variants tried: tools/barrier_bench.cu). The host zeroes the counter before every launch.
__device__ __forceinline__ void grid_barrier(unsigned* counter, unsigned& epoch, int nblocks) {
  __syncthreads();
  epoch += 1;
  if (threadIdx.x == 0) {
    red_release_add_u32(counter, 1u);
    const unsigned target = epoch * (unsigned)nblocks;
    while (ld_relaxed_u32(counter) < target) { }
    fence_acq_rel_gpu();
  }
  __syncthreads();
}

// LayerNorm statistics: every CTA stores its partial (sum a, sum b) of a row as one float2 into part[b][cta]; after the
// barrier each CTA re-reads the G partials of every row from L2 and adds them in a fixed order (deterministic run to run).
// All loads of a pass are issued before the first add (in-order issue would otherwise serialise the L2 round trips).
// out[b] = (sum a, sum b) * inv_n
__device__ __forceinline__ void reduce_row_partials(const float2* part, int B, int G, float inv_n, float* out, int warp, int lane) {
  constexpr int NC = SCAN_MAX_CTAS / 32;
  for (int b0 = warp; b0 < B; b0 += 4 * SCAN_WARPS) {
    float2 v[4][NC];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int b = b0 + r * SCAN_WARPS;
#pragma unroll
      for (int k = 0; k < NC; ++k) {
        const int c = lane + 32 * k;
        v[r][k] = (b < B && c < G) ? __ldcg(part + (size_t)b * G + c) : make_float2(0.f, 0.f);
      }
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int b = b0 + r * SCAN_WARPS;
      float s1 = 0.f, s2 = 0.f;
#pragma unroll
      for (int k = 0; k < NC; ++k) { s1 += v[r][k].x; s2 += v[r][k].y; }
      s1 = warp_sum(s1); s2 = warp_sum(s2);
      if (b < B && lane == 0) { out[b * 2] = s1 * inv_n; out[b * 2 + 1] = s2 * inv_n; }
    }
  }
}

// after the call lane l holds sum over lanes of v[l]
__device__ __forceinline__ float reduce_scatter32(float (&v)[32], int lane) {
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) {
    const bool up = (lane & o) != 0;
#pragma unroll
    for (int i = 0; i < o; ++i) {
      float send = up ? v[i] : v[i + o];
      float keep = up ? v[i + o] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
    }
  }
  return v[0];
}

// ---- bf16 tensor-core path (mma.sync m16n8k16, fp32 accumulate) for the per-step recurrent product ----------------------
// Both operands are K-contiguous, so every thread fetches natural-order 16-byte chunks (8 bf16 = elements [8q, 8q+8) of a
// 32-wide K block) from its A rows (g, g+8) and its B row (g); words {0,1} of a chunk feed the first k16 step and words
// {2,3} the second. A and B use the same element->fragment mapping, so the contraction is exact whatever the order.
constexpr int SCAN_KPAD = 32;    // bf16 padding per resident weight row: row stride = 64 B mod 128 B -> conflict-free LDS.128
__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint4 ldcg_u4(const uint4* p) { return __ldcg(p); }

// acc[nt][:] += A[16 rows of this warp's m-tile][K slice] * Wb[nt*8 .. nt*8+8][K slice]^T
//   alo / ahi: global bf16 rows g and g+8 of the m-tile, already offset to (K slice start + 8q); null -> zero rows
//   wrow:      shared bf16 row g of n-tile 0, same offset; n-tile stride 8*wld
template <int NT, int PF>
__device__ __forceinline__ void warp_mma_slice(const uint4* alo, const uint4* ahi, const __nv_bfloat16* wrow, int wld, int nblk, float (&acc)[NT][4]) {
  const uint4 zero = make_uint4(0u, 0u, 0u, 0u);
  for (int blk0 = 0; blk0 < nblk; blk0 += PF) {
    uint4 lo[PF], hi[PF];
#pragma unroll
    for (int i = 0; i < PF; ++i) {
      const int blk = blk0 + i;
      lo[i] = (alo && blk < nblk) ? ldcg_u4(alo + blk * 4) : zero;
      hi[i] = (ahi && blk < nblk) ? ldcg_u4(ahi + blk * 4) : zero;
    }
#pragma unroll
    for (int i = 0; i < PF; ++i) {
      const int blk = blk0 + i;
      if (blk < nblk) {
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
          const uint4 w = *reinterpret_cast<const uint4*>(wrow + (size_t)nt * 8 * wld + blk * 32);
          mma_bf16_16816(acc[nt], lo[i].x, hi[i].x, lo[i].y, hi[i].y, w.x, w.y);
          mma_bf16_16816(acc[nt], lo[i].z, hi[i].z, lo[i].w, hi[i].w, w.z, w.w);
        }
      }
    }
  }
}

// acc[rb*R + r] += sum_k Ws[r][k] * (scale[rb] * hrow[rb][k]),  k strided over the lanes of one warp
template <int R>
__device__ __forceinline__ void warp_matvec(const float* __restrict__ Ws, int Kd, const float* const (&hrow)[SCAN_RB], const float (&scale)[SCAN_RB],
                                            float (&acc)[SCAN_RB * R], int lane) {
  if ((Kd & 127) == 0) {
    float4 hn[SCAN_RB];
#pragma unroll
    for (int rb = 0; rb < SCAN_RB; ++rb) hn[rb] = hrow[rb] ? __ldcg(reinterpret_cast<const float4*>(hrow[rb] + 4 * lane)) : make_float4(0.f, 0.f, 0.f, 0.f);
    for (int k = 4 * lane; k < Kd; k += 128) {
      float4 h[SCAN_RB];
#pragma unroll
      for (int rb = 0; rb < SCAN_RB; ++rb) {
        h[rb] = hn[rb];
        h[rb].x *= scale[rb]; h[rb].y *= scale[rb]; h[rb].z *= scale[rb]; h[rb].w *= scale[rb];
      }
      if (k + 128 < Kd) {                        // prefetch the next slice while this one is multiplied
#pragma unroll
        for (int rb = 0; rb < SCAN_RB; ++rb) if (hrow[rb]) hn[rb] = __ldcg(reinterpret_cast<const float4*>(hrow[rb] + k + 128));
      }
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const float4 w = *reinterpret_cast<const float4*>(Ws + (size_t)r * Kd + k);
#pragma unroll
        for (int rb = 0; rb < SCAN_RB; ++rb)
          acc[rb * R + r] = fmaf(w.x, h[rb].x, fmaf(w.y, h[rb].y, fmaf(w.z, h[rb].z, fmaf(w.w, h[rb].w, acc[rb * R + r]))));
      }
    }
  } else {
    for (int k = lane; k < Kd; k += 32) {
      float h[SCAN_RB];
#pragma unroll
      for (int rb = 0; rb < SCAN_RB; ++rb) h[rb] = hrow[rb] ? __ldcg(hrow[rb] + k) * scale[rb] : 0.f;
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const float w = Ws[(size_t)r * Kd + k];
#pragma unroll
        for (int rb = 0; rb < SCAN_RB; ++rb) acc[rb * R + r] = fmaf(w, h[rb], acc[rb * R + r]);
      }
    }
  }
}

struct ScanFwdParams {
  int B, T, D, G;
  const float* gx;            // [B*T, 3D]  x-part of the GRU pre-activation (row n = b*T + t)
  const float* Wh; long long ldw;   // W_gru[:, Hd:], 3D rows, row stride ldw
  const float* gamma; const float* beta;
  const float* mask;          // [B*T] 1 - is_first, or null
  const float* h0;            // [B, D] or null (zeros)
  float* deter;               // [B*T, D]
  void* deter_t; long long ld_dt; int t_dtype;      // optional copy in activation dtype (feat buffer slice)
  void* hprev_t; long long ld_hp;                   // optional masked h_{t-1} in activation dtype ([x|h] operand slice)
  float* z;                   // [B*T, 3D] pre-LayerNorm activations (saved for backward)
  float* stats;               // [B*T, 2]
  float2* part;               // workspace [B][G] per-CTA partial (sum, sum of squares)
  unsigned* barrier;          // workspace: barrier counter, zeroed by the host
  long long* dbg;             // optional [T][6] globaltimer stamps of block 0 (profiling aid), or null
};

__device__ __forceinline__ long long gtime() { long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
#define SCAN_STAMP(i) do { if (p.dbg && blockIdx.x == 0 && threadIdx.x == 0) p.dbg[t * 6 + (i)] = gtime(); } while (0)

// MMA = true: bf16 tensor-core mat-vec (weights resident as bf16, h read from the bf16 hprev_t rows); false: fp32 FFMA.
template <bool MMA>
__global__ void __launch_bounds__(SCAN_THREADS, 1) rssm_scan_fwd_kernel(ScanFwdParams p) {
  genrl::pdl_sync();
  extern __shared__ float smem[];
  const int D = p.D, D3 = 3 * p.D, B = p.B, T = p.T, G = p.G;
  constexpr int R = 3 * SCAN_U;
  const int WLD = D + SCAN_KPAD;          // bf16 row stride of the resident slice (MMA mode)
  float* Ws = smem;                       // [R][D] fp32, or [R][WLD] bf16 in MMA mode
  __nv_bfloat16* Wb = reinterpret_cast<__nv_bfloat16*>(smem);
  float* red = MMA ? smem + (size_t)R * WLD / 2 : nullptr;             // [4][64][R] split-K partials (MMA mode)
  float* zs = MMA ? red + 4 * 64 * R : Ws + (size_t)R * D;             // [B][R]  this CTA's pre-LN values for the current step
  float* st = zs + (size_t)B * R;         // [B][2]  mean / rstd
  float* hs = st + (size_t)B * 2;         // [B][U]  h_{t-1} of the owned units (fp32, unmasked)
  float* ms = hs + (size_t)B * SCAN_U;    // [2][B]  mask at t and t+1
  float* gb = ms + (size_t)2 * B;         // [2][R]  gamma / beta of the owned rows
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int j0 = blockIdx.x * SCAN_U;
  unsigned epoch = 0;

  // resident weight slice: row r = gate*U + u  <->  W_h[gate*D + j0 + u, :]
  for (int i = tid; i < R * D; i += SCAN_THREADS) {
    int r = i / D, k = i % D;
    int gate = r / SCAN_U, u = r % SCAN_U;
    float w = p.Wh[(size_t)(gate * D + j0 + u) * p.ldw + k];
    if (MMA) Wb[(size_t)r * WLD + k] = __float2bfloat16_rn(w);
    else Ws[i] = w;
  }
  for (int r = tid; r < R; r += SCAN_THREADS) {
    int gate = r / SCAN_U, u = r % SCAN_U;
    gb[r] = p.gamma[gate * D + j0 + u];
    gb[R + r] = p.beta[gate * D + j0 + u];
  }
  // owned slice of the initial state; masked copy for the [x|h] operand (row t = 0)
  for (int i = tid; i < B * SCAN_U; i += SCAN_THREADS) {
    int b = i / SCAN_U, u = i % SCAN_U;
    float v = p.h0 ? p.h0[(size_t)b * D + j0 + u] : 0.f;
    hs[i] = v;
    if (p.hprev_t) {
      float m = p.mask ? p.mask[(size_t)b * T] : 1.f;
      st_any(p.hprev_t, p.t_dtype, (size_t)(b * (long long)T) * p.ld_hp + j0 + u, v * m);
    }
  }
  if (MMA) grid_barrier(p.barrier, epoch, G);      // every CTA reads all of row t = 0
  else __syncthreads();

  const int ngroups = (B + SCAN_RB - 1) / SCAN_RB;
  for (int t = 0; t < T; ++t) {
    // ---- phase 1: zh = W_h (m * h_{t-1}) for every batch row ----
    SCAN_STAMP(0);
    for (int b = tid; b < B; b += SCAN_THREADS) {
      ms[b] = p.mask ? p.mask[(size_t)b * T + t] : 1.f;
      ms[B + b] = (p.mask && t + 1 < T) ? p.mask[(size_t)b * T + t + 1] : 1.f;
    }
    if (MMA) {
      // warp (kq, mq): K quarter kq, batch rows [16 mq, 16 mq + 16) of the current 64-row chunk, all 24 weight rows
      const __nv_bfloat16* hb = reinterpret_cast<const __nv_bfloat16*>(p.hprev_t);
      const int kq = warp >> 2, mq = warp & 3, g = lane >> 2, q = lane & 3;
      const int KQ = D >> 2;
      for (int bc = 0; bc < B; bc += 64) {
        // x-part for the combine step below: issued before the MMA loop so the HBM latency hides behind it
        float gxv[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          const int i = tid + k * SCAN_THREADS, bl = i / R, r = i % R, b = bc + bl;
          gxv[k] = b < B ? p.gx[((size_t)b * T + t) * D3 + (r / SCAN_U) * D + j0 + (r % SCAN_U)] : 0.f;
        }
        float acc[3][4];
#pragma unroll
        for (int i = 0; i < 3; ++i) { acc[i][0] = 0.f; acc[i][1] = 0.f; acc[i][2] = 0.f; acc[i][3] = 0.f; }
        const int b_lo = bc + mq * 16 + g, b_hi = b_lo + 8;
        const uint4* alo = b_lo < B ? reinterpret_cast<const uint4*>(hb + ((size_t)b_lo * T + t) * p.ld_hp + kq * KQ + q * 8) : nullptr;
        const uint4* ahi = b_hi < B ? reinterpret_cast<const uint4*>(hb + ((size_t)b_hi * T + t) * p.ld_hp + kq * KQ + q * 8) : nullptr;
        warp_mma_slice<3, 8>(alo, ahi, Wb + (size_t)g * WLD + kq * KQ + q * 8, WLD, KQ >> 5, acc);
#pragma unroll
        for (int nt = 0; nt < 3; ++nt) {
          float* r = red + ((size_t)(kq * 64 + mq * 16 + g) * R) + nt * 8 + q * 2;
          r[0] = acc[nt][0]; r[1] = acc[nt][1]; r[8 * R] = acc[nt][2]; r[8 * R + 1] = acc[nt][3];
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          const int i = tid + k * SCAN_THREADS, bl = i / R, r = i % R, b = bc + bl;
          if (b < B) {
            const float zv = (red[i] + red[64 * R + i]) + (red[2 * 64 * R + i] + red[3 * 64 * R + i]) + gxv[k];
            zs[b * R + r] = zv;
            p.z[((size_t)b * T + t) * D3 + (r / SCAN_U) * D + j0 + (r % SCAN_U)] = zv;
          }
        }
        if (bc + 64 < B) __syncthreads();
      }
    } else {
      for (int g = warp; g < ngroups; g += SCAN_WARPS) {
        const float* hrow[SCAN_RB];
        float scale[SCAN_RB];
        float acc[SCAN_RB * R];
#pragma unroll
        for (int i = 0; i < SCAN_RB * R; ++i) acc[i] = 0.f;
#pragma unroll
        for (int rb = 0; rb < SCAN_RB; ++rb) {
          int b = g * SCAN_RB + rb;
          if (b < B) {
            scale[rb] = p.mask ? p.mask[(size_t)b * T + t] : 1.f;
            hrow[rb] = t == 0 ? (p.h0 ? p.h0 + (size_t)b * D : nullptr) : p.deter + ((size_t)b * T + t - 1) * D;
          } else { scale[rb] = 0.f; hrow[rb] = nullptr; }
        }
        warp_matvec<R>(Ws, D, hrow, scale, acc, lane);
        // 96 partial sums per lane -> 3 transposing reductions; lane l of set s owns flat index 32 s + l = rb*R + r
#pragma unroll
        for (int s = 0; s < (SCAN_RB * R) / 32; ++s) {
          float v[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = acc[s * 32 + i];
          float tot = reduce_scatter32(v, lane);
          int idx = s * 32 + lane;
          int rb = idx / R, r = idx % R;
          int b = g * SCAN_RB + rb;
          if (b < B) {
            int gate = r / SCAN_U, u = r % SCAN_U;
            size_t n = (size_t)b * T + t;
            float zv = tot + p.gx[n * D3 + gate * D + j0 + u];
            zs[b * R + r] = zv;
            p.z[n * D3 + gate * D + j0 + u] = zv;
          }
        }
      }
    }
    __syncthreads();
    // ---- phase 2: LayerNorm partial sums of this CTA's 24 values per row ----
    SCAN_STAMP(1);
    for (int b = tid; b < B; b += SCAN_THREADS) {
      float s1 = 0.f, s2 = 0.f;
#pragma unroll
      for (int r = 0; r < R; ++r) { float v = zs[b * R + r]; s1 += v; s2 += v * v; }
      p.part[(size_t)b * G + blockIdx.x] = make_float2(s1, s2);
    }
    SCAN_STAMP(2);
    grid_barrier(p.barrier, epoch, G);
    SCAN_STAMP(3);
    // ---- phase 3: full-row statistics ----
    reduce_row_partials(p.part, B, G, 1.0f / D3, st, warp, lane);
    __syncthreads();
    // ---- phase 4: gates for the owned units ----
    SCAN_STAMP(4);
    for (int i = tid; i < B * SCAN_U; i += SCAN_THREADS) {
      int b = i / SCAN_U, u = i % SCAN_U, j = j0 + u;
      float mean = st[b * 2];
      float rstd = 1.0f / sqrtf(fmaxf(st[b * 2 + 1] - mean * mean, 0.f) + 1e-5f);
      if (blockIdx.x == 0 && u == 0) { p.stats[((size_t)b * T + t) * 2] = mean; p.stats[((size_t)b * T + t) * 2 + 1] = rstd; }
      float pr = (zs[b * R + u] - mean) * rstd * gb[u] + gb[R + u];
      float pc = (zs[b * R + SCAN_U + u] - mean) * rstd * gb[SCAN_U + u] + gb[R + SCAN_U + u];
      float pu = (zs[b * R + 2 * SCAN_U + u] - mean) * rstd * gb[2 * SCAN_U + u] + gb[R + 2 * SCAN_U + u];
      float r = sigmoidf_(pr), cand = tanhf(r * pc), uu = sigmoidf_(pu - 1.0f);
      float hn = uu * cand + (1.0f - uu) * (hs[i] * ms[b]);
      hs[i] = hn;
      size_t n = (size_t)b * T + t;
      p.deter[n * D + j] = hn;
      if (p.deter_t) st_any(p.deter_t, p.t_dtype, n * p.ld_dt + j, hn);
      if (p.hprev_t && t + 1 < T) st_any(p.hprev_t, p.t_dtype, (n + 1) * p.ld_hp + j, hn * ms[B + b]);
    }
    SCAN_STAMP(5);
    grid_barrier(p.barrier, epoch, G);
  }
}

struct ScanBwdParams {
  int B, T, D, G;
  const float* dh_ext;        // [B*T, D] gradient reaching each h_t from outside the recurrence
  const float* z;             // [B*T, 3D]
  const float* stats;         // [B*T, 2]
  const float* Wh; long long ldw;
  const float* gamma; const float* beta;
  const float* mask;          // [B*T] or null
  const float* deter;         // [B*T, D]
  const float* h0;            // [B, D] or null
  void* dz_t; long long ld_dz; int t_dtype;      // [B*T, 3D] gradient of pre-LN activations (activation dtype)
  float* dh0;                 // [B, D] or null
  float* dgamma; float* dbeta;                   // [3D] accumulated (+=), or null
  float* dzf;                 // workspace [2][B][3D] fp32 (FFMA mode)
  float2* part;               // workspace [B][G] per-CTA partial sums
  unsigned* barrier;          // workspace: barrier counter, zeroed by the host
};

// inputs of one (batch row, owned unit) gate-gradient item; loaded one step ahead so HBM latency stays off the serial chain
struct ScanBwdIn { float mean, rstd, zr, zc, zu, hp, dh, m; };
__device__ __forceinline__ ScanBwdIn scan_bwd_load(const ScanBwdParams& p, int b, int j, int t) {
  ScanBwdIn v;
  const size_t n = (size_t)b * p.T + t;
  const int D = p.D;
  v.mean = p.stats[n * 2]; v.rstd = p.stats[n * 2 + 1];
  v.zr = p.z[n * 3 * D + j]; v.zc = p.z[n * 3 * D + D + j]; v.zu = p.z[n * 3 * D + 2 * D + j];
  v.m = p.mask ? p.mask[n] : 1.f;
  v.hp = t == 0 ? (p.h0 ? p.h0[(size_t)b * D + j] : 0.f) : p.deter[(n - 1) * D + j];
  v.dh = p.dh_ext[n * D + j];
  return v;
}

template <bool MMA>
__global__ void __launch_bounds__(SCAN_THREADS, 1) rssm_scan_bwd_kernel(ScanBwdParams p) {
  genrl::pdl_sync();
  extern __shared__ float smem[];
  const int D = p.D, D3 = 3 * p.D, B = p.B, T = p.T, G = p.G;
  constexpr int R = 3 * SCAN_U;
  const int WLD = D3 + SCAN_KPAD;            // bf16 row stride of the resident transposed slice (MMA mode)
  float* Wc = smem;                          // [U][3D]   Wc[u][i] = W_h[i, j0+u]   (fp32, or [U][WLD] bf16 in MMA mode)
  __nv_bfloat16* Wcb = reinterpret_cast<__nv_bfloat16*>(smem);
  float* red = MMA ? smem + (size_t)SCAN_U * WLD / 2 : nullptr;        // [4][64][U] split-K partials (MMA mode)
  float* gs = MMA ? red + 4 * 64 * SCAN_U : Wc + (size_t)SCAN_U * D3;  // [B][R]    gamma * gate-gradient
  float* xs = gs + (size_t)B * R;            // [B][R]    normalised z
  float* carry = xs + (size_t)B * R;         // [B][U]    gradient flowing into h_t from step t+1
  float* gate_keep = carry + (size_t)B * SCAN_U;   // [B][U]  (1-u_t) * dh_total
  float* st = gate_keep + (size_t)B * SCAN_U;      // [B][2]
  float* rs = st + (size_t)B * 2;            // [B] rstd, [B] mask of the current step
  float* dgl = rs + (size_t)2 * B;           // [2][R] local dgamma / dbeta accumulators
  float* gb = dgl + 2 * R;                   // [2][R] gamma / beta of the owned rows
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int j0 = blockIdx.x * SCAN_U;
  unsigned epoch = 0;

  for (int i = tid; i < SCAN_U * D3; i += SCAN_THREADS) {
    int u = i / D3, row = i % D3;
    float w = p.Wh[(size_t)row * p.ldw + j0 + u];
    if (MMA) Wcb[(size_t)u * WLD + row] = __float2bfloat16_rn(w);
    else Wc[i] = w;
  }
  for (int i = tid; i < B * SCAN_U; i += SCAN_THREADS) carry[i] = 0.f;
  for (int r = tid; r < R; r += SCAN_THREADS) {
    int gate = r / SCAN_U, u = r % SCAN_U;
    dgl[r] = 0.f; dgl[R + r] = 0.f;
    gb[r] = p.gamma[gate * D + j0 + u];
    gb[R + r] = p.beta[gate * D + j0 + u];
  }
  ScanBwdIn nxt;
  if (tid < B * SCAN_U) nxt = scan_bwd_load(p, tid / SCAN_U, j0 + tid % SCAN_U, T - 1);
  __syncthreads();

  const int ngroups = (B + SCAN_RB - 1) / SCAN_RB;
  for (int t = T - 1; t >= 0; --t) {
    float* dzf = p.dzf + (size_t)(t & 1) * B * D3;
    // ---- phase 1: gate gradients for owned units, LayerNorm-backward partial sums ----
    for (int i = tid; i < B * SCAN_U; i += SCAN_THREADS) {
      int b = i / SCAN_U, u = i % SCAN_U, j = j0 + u;
      ScanBwdIn in = i == tid ? nxt : scan_bwd_load(p, b, j, t);
      if (i == tid && t > 0) nxt = scan_bwd_load(p, b, j, t - 1);       // consumed one step later
      float xr = (in.zr - in.mean) * in.rstd, xc = (in.zc - in.mean) * in.rstd, xu = (in.zu - in.mean) * in.rstd;
      float pr = xr * gb[u] + gb[R + u], pc = xc * gb[SCAN_U + u] + gb[R + SCAN_U + u], pu = xu * gb[2 * SCAN_U + u] + gb[R + 2 * SCAN_U + u];
      float r = sigmoidf_(pr), cand = tanhf(r * pc), uu = sigmoidf_(pu - 1.0f);
      float hp = in.hp * in.m;
      float g = in.dh + carry[i];
      float drc = g * uu * (1.0f - cand * cand);
      float gpr = drc * pc * r * (1.0f - r), gpc = drc * r, gpu = g * (cand - hp) * uu * (1.0f - uu);
      gs[b * R + u] = gpr * gb[u]; gs[b * R + SCAN_U + u] = gpc * gb[SCAN_U + u]; gs[b * R + 2 * SCAN_U + u] = gpu * gb[2 * SCAN_U + u];
      xs[b * R + u] = xr; xs[b * R + SCAN_U + u] = xc; xs[b * R + 2 * SCAN_U + u] = xu;
      gate_keep[i] = g * (1.0f - uu);
      if (u == 0) { rs[b] = in.rstd; rs[B + b] = in.m; }
      if (p.dgamma) {
        atomicAdd(&dgl[u], gpr * xr); atomicAdd(&dgl[SCAN_U + u], gpc * xc); atomicAdd(&dgl[2 * SCAN_U + u], gpu * xu);
        atomicAdd(&dgl[R + u], gpr); atomicAdd(&dgl[R + SCAN_U + u], gpc); atomicAdd(&dgl[R + 2 * SCAN_U + u], gpu);
      }
    }
    __syncthreads();
    for (int b = tid; b < B; b += SCAN_THREADS) {
      float s1 = 0.f, s2 = 0.f;
#pragma unroll
      for (int r = 0; r < R; ++r) { s1 += gs[b * R + r]; s2 += gs[b * R + r] * xs[b * R + r]; }
      p.part[(size_t)b * G + blockIdx.x] = make_float2(s1, s2);
    }
    grid_barrier(p.barrier, epoch, G);
    // ---- phase 2: dz for owned rows ----
    reduce_row_partials(p.part, B, G, 1.0f / D3, st, warp, lane);
    __syncthreads();
    for (int i = tid; i < B * R; i += SCAN_THREADS) {
      int b = i / R, r = i % R;
      int gate = r / SCAN_U, u = r % SCAN_U;
      size_t n = (size_t)b * T + t;
      float dzv = rs[b] * (gs[i] - st[b * 2] - xs[i] * st[b * 2 + 1]);
      int col = gate * D + j0 + u;
      if (!MMA) dzf[(size_t)b * D3 + col] = dzv;
      st_any(p.dz_t, p.t_dtype, n * p.ld_dz + col, dzv);
    }
    grid_barrier(p.barrier, epoch, G);
    // ---- phase 3: carry_{t-1}[b, units] = m_t * (W_h[:, units]^T dz_t + (1-u_t) dh_t) ----
    if (MMA) {
      // warp (kq, mq): quarter kq of the 3D contraction, batch rows [16 mq, 16 mq + 16) of the 64-row chunk, 8 units
      const __nv_bfloat16* dzb = reinterpret_cast<const __nv_bfloat16*>(p.dz_t);
      const int kq = warp >> 2, mq = warp & 3, g = lane >> 2, q = lane & 3;
      const int KQ = D3 >> 2;
      for (int bc = 0; bc < B; bc += 64) {
        float acc[1][4] = {{0.f, 0.f, 0.f, 0.f}};
        const int b_lo = bc + mq * 16 + g, b_hi = b_lo + 8;
        const uint4* alo = b_lo < B ? reinterpret_cast<const uint4*>(dzb + ((size_t)b_lo * T + t) * p.ld_dz + kq * KQ + q * 8) : nullptr;
        const uint4* ahi = b_hi < B ? reinterpret_cast<const uint4*>(dzb + ((size_t)b_hi * T + t) * p.ld_dz + kq * KQ + q * 8) : nullptr;
        warp_mma_slice<1, 8>(alo, ahi, Wcb + (size_t)g * WLD + kq * KQ + q * 8, WLD, KQ >> 5, acc);
        float* r = red + ((size_t)(kq * 64 + mq * 16 + g) * SCAN_U) + q * 2;
        r[0] = acc[0][0]; r[1] = acc[0][1]; r[8 * SCAN_U] = acc[0][2]; r[8 * SCAN_U + 1] = acc[0][3];
        __syncthreads();
        for (int i = tid; i < 64 * SCAN_U; i += SCAN_THREADS) {
          const int bl = i / SCAN_U, u = i % SCAN_U, b = bc + bl;
          if (b < B) {
            const float tot = (red[i] + red[64 * SCAN_U + i]) + (red[2 * 64 * SCAN_U + i] + red[3 * 64 * SCAN_U + i]);
            const float c = rs[B + b] * (tot + gate_keep[b * SCAN_U + u]);
            carry[b * SCAN_U + u] = c;
            if (t == 0 && p.dh0) p.dh0[(size_t)b * D + j0 + u] = c;
          }
        }
        if (bc + 64 < B) __syncthreads();
      }
    } else {
      for (int g = warp; g < ngroups; g += SCAN_WARPS) {
        const float* drow[SCAN_RB];
        float scale[SCAN_RB];
        float acc[SCAN_RB * SCAN_U];
#pragma unroll
        for (int i = 0; i < SCAN_RB * SCAN_U; ++i) acc[i] = 0.f;
#pragma unroll
        for (int rb = 0; rb < SCAN_RB; ++rb) {
          int b = g * SCAN_RB + rb;
          drow[rb] = b < B ? dzf + (size_t)b * D3 : nullptr;
          scale[rb] = 1.f;
        }
        warp_matvec<SCAN_U>(Wc, D3, drow, scale, acc, lane);
        float tot = reduce_scatter32(acc, lane);          // lane = rb*U + u
        int rb = lane / SCAN_U, u = lane % SCAN_U;
        int b = g * SCAN_RB + rb;
        if (b < B) {
          float c = rs[B + b] * (tot + gate_keep[b * SCAN_U + u]);
          carry[b * SCAN_U + u] = c;
          if (t == 0 && p.dh0) p.dh0[(size_t)b * D + j0 + u] = c;
        }
      }
    }
    __syncthreads();
  }
  if (p.dgamma) {
    for (int r = tid; r < R; r += SCAN_THREADS) {
      int gate = r / SCAN_U, u = r % SCAN_U;
      p.dgamma[gate * D + j0 + u] += dgl[r];
      p.dbeta[gate * D + j0 + u] += dgl[R + r];
    }
  }
}

}  // namespace genrl

using namespace genrl;

static size_t scan_part_bytes(int B, int D) { return (((size_t)B * (D / SCAN_U) * sizeof(float2)) + 255) / 256 * 256; }

extern "C" long long genrl_rssm_scan_workspace_bytes(int B, int T, int D) {
  // [256 B barrier counter][B*G float2 LayerNorm partials][2*B*3D fp32 dz staging (backward, fp32 mode)]
  (void)T;
  return SCAN_FLAG_BYTES + (long long)scan_part_bytes(B, D) + (long long)2 * B * 3 * D * 4;
}

static long long* g_scan_dbg = nullptr;
// profiling aid: when set, block 0 of the forward scan records 6 globaltimer stamps per step into dbg[T][6]
extern "C" int genrl_rssm_scan_set_debug(long long* dbg) { g_scan_dbg = dbg; return 0; }

// GENRL_SCAN_FFMA=1 keeps the fp32 FFMA mat-vec even for bf16 activations (A/B measurements)
static bool scan_force_ffma() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("GENRL_SCAN_FFMA"); v = (e && e[0] == '1') ? 1 : 0; }
  return v == 1;
}

static int scan_check(int B, int T, int D, const char* who) {
  GENRL_REQUIRE(B > 0 && T > 0, "%s: empty batch", who);
  GENRL_REQUIRE(D % 32 == 0, "%s: deter size %d must be a multiple of 32", who, D);
  int sms = 0, dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  GENRL_REQUIRE(D / SCAN_U <= sms, "%s: deter size %d needs %d co-resident CTAs > %d SMs (use the per-step path)", who, D, D / SCAN_U, sms);
  GENRL_REQUIRE(D / SCAN_U <= SCAN_MAX_CTAS, "%s: %d CTAs exceed the statistics exchange capacity %d", who, D / SCAN_U, SCAN_MAX_CTAS);
  return 0;
}

// workspace: genrl_rssm_scan_workspace_bytes() bytes; the barrier counter at its head is zeroed here.
extern "C" int genrl_rssm_scan_fwd(void* stream, int B, int T, int D, const float* gx, const float* Wh, long long ldw, const float* gamma,
                                   const float* beta, const float* mask, const float* h0, float* deter, void* deter_t, long long ld_dt,
                                   void* hprev_t, long long ld_hp, int t_dtype, float* z, float* stats, void* workspace) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (scan_check(B, T, D, "rssm_scan_fwd")) return 1;
  ScanFwdParams p;
  p.B = B; p.T = T; p.D = D; p.G = D / SCAN_U;
  p.gx = gx; p.Wh = Wh; p.ldw = ldw; p.gamma = gamma; p.beta = beta; p.mask = mask; p.h0 = h0;
  p.deter = deter; p.deter_t = deter_t; p.ld_dt = ld_dt; p.t_dtype = t_dtype; p.hprev_t = hprev_t; p.ld_hp = ld_hp;
  p.z = z; p.stats = stats;
  p.dbg = g_scan_dbg;
  p.barrier = reinterpret_cast<unsigned*>(workspace);
  p.part = reinterpret_cast<float2*>(reinterpret_cast<char*>(workspace) + SCAN_FLAG_BYTES);
  const int R = 3 * SCAN_U;
  // tensor-core path: bf16 activations with the masked h_{t-1} rows available as 16-byte-aligned bf16
  const bool mma = t_dtype == GENRL_BF16 && hprev_t && D % 128 == 0 && ld_hp % 8 == 0 && (reinterpret_cast<uintptr_t>(hprev_t) & 15) == 0 && !scan_force_ffma();
  const size_t tail = ((size_t)B * R + (size_t)B * 2 + (size_t)B * SCAN_U + (size_t)2 * B + 2 * R) * sizeof(float);
  size_t smem = mma ? (size_t)R * (D + SCAN_KPAD) * 2 + (size_t)4 * 64 * R * sizeof(float) + tail : (size_t)R * D * sizeof(float) + tail;
  GENRL_REQUIRE(smem <= 227 * 1024, "rssm_scan_fwd: %zu bytes of shared memory needed (B=%d D=%d)", smem, B, D);
  void* kern = mma ? (void*)rssm_scan_fwd_kernel<true> : (void*)rssm_scan_fwd_kernel<false>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  GENRL_REQUIRE(e == cudaSuccess, "rssm_scan_fwd: smem attribute: %s", cudaGetErrorString(e));
  cudaMemsetAsync(workspace, 0, SCAN_FLAG_BYTES, st);
  void* args[] = {&p};
  e = cudaLaunchCooperativeKernel(kern, dim3(p.G), dim3(SCAN_THREADS), args, smem, st);
  GENRL_REQUIRE(e == cudaSuccess, "rssm_scan_fwd: cooperative launch failed: %s", cudaGetErrorString(e));
  return check_launch("rssm_scan_fwd");
}

extern "C" int genrl_rssm_scan_bwd(void* stream, int B, int T, int D, const float* dh_ext, const float* z, const float* stats, const float* Wh,
                                   long long ldw, const float* gamma, const float* beta, const float* mask, const float* deter, const float* h0,
                                   void* dz_t, long long ld_dz, int t_dtype, float* dh0, float* dgamma, float* dbeta, void* workspace) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (scan_check(B, T, D, "rssm_scan_bwd")) return 1;
  ScanBwdParams p;
  p.B = B; p.T = T; p.D = D; p.G = D / SCAN_U;
  p.dh_ext = dh_ext; p.z = z; p.stats = stats; p.Wh = Wh; p.ldw = ldw; p.gamma = gamma; p.beta = beta; p.mask = mask;
  p.deter = deter; p.h0 = h0; p.dz_t = dz_t; p.ld_dz = ld_dz; p.t_dtype = t_dtype; p.dh0 = dh0; p.dgamma = dgamma; p.dbeta = dbeta;
  p.barrier = reinterpret_cast<unsigned*>(workspace);
  p.part = reinterpret_cast<float2*>(reinterpret_cast<char*>(workspace) + SCAN_FLAG_BYTES);
  p.dzf = reinterpret_cast<float*>(reinterpret_cast<char*>(workspace) + SCAN_FLAG_BYTES + scan_part_bytes(B, D));
  const int R = 3 * SCAN_U;
  const bool mma = t_dtype == GENRL_BF16 && D % 128 == 0 && ld_dz % 8 == 0 && (reinterpret_cast<uintptr_t>(dz_t) & 15) == 0 && !scan_force_ffma();
  const size_t tail = ((size_t)2 * B * R + (size_t)2 * B * SCAN_U + (size_t)B * 2 + (size_t)2 * B + 4 * R) * sizeof(float);
  size_t smem = mma ? (size_t)SCAN_U * (3 * D + SCAN_KPAD) * 2 + (size_t)4 * 64 * SCAN_U * sizeof(float) + tail
                    : (size_t)SCAN_U * 3 * D * sizeof(float) + tail;
  GENRL_REQUIRE(smem <= 227 * 1024, "rssm_scan_bwd: %zu bytes of shared memory needed (B=%d D=%d)", smem, B, D);
  void* kern = mma ? (void*)rssm_scan_bwd_kernel<true> : (void*)rssm_scan_bwd_kernel<false>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  GENRL_REQUIRE(e == cudaSuccess, "rssm_scan_bwd: smem attribute: %s", cudaGetErrorString(e));
  cudaMemsetAsync(workspace, 0, SCAN_FLAG_BYTES, st);
  void* args[] = {&p};
  e = cudaLaunchCooperativeKernel(kern, dim3(p.G), dim3(SCAN_THREADS), args, smem, st);
  GENRL_REQUIRE(e == cudaSuccess, "rssm_scan_bwd: cooperative launch failed: %s", cudaGetErrorString(e));
  return check_launch("rssm_scan_bwd");
}
