// Persistent RSSM time-loop kernels (sm_100a): the only truly sequential part of EnsembleRSSM.observe
// (reference dreamer_utils.py:362-371 static_scan over obs_step :432-440 -> img_step :459-473 -> GRUCell :771-785).
//
// With single_obs_posterior the posterior, the GRU input projection x_t and the prior head are batched GEMMs over all
// B*T rows; what remains per step is   z_t = Gx_t + W_h (m_t * h_{t-1});  h_t = gates(LN_3D(z_t), m_t * h_{t-1}).
// One cooperative launch loops the time axis:
//   * G = D/8 CTAs, CTA c owns hidden units [8c, 8c+8): the 24 rows of W_h for its (reset,cand,update) gates live in
//     shared memory for the whole sequence (D=1024: 96 KB fp32 per CTA, 12.6 MB across the grid, read from HBM once);
//   * per step: FFMA mat-vec for all B rows, LayerNorm(3D) statistics exchanged through a [B][G] partial array,
//     gates, h_t written to HBM/L2 for the other CTAs; two grid barriers per step.
// The backward kernel walks t = T-1..0 with the transposed slice W_h[:, units] resident and produces dz_t (the gradient
// of the pre-LayerNorm GRU activations) for the batched wgrad/dgrad GEMMs that follow.
#include "common.cuh"

namespace genrl {

constexpr int SCAN_U = 8;        // hidden units per CTA
constexpr int SCAN_RB = 4;       // batch rows processed together by one warp
constexpr int SCAN_THREADS = 512;
constexpr int SCAN_WARPS = SCAN_THREADS / 32;

__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

__device__ __forceinline__ void grid_barrier(unsigned* counter, unsigned& epoch, unsigned nblocks) {
  __syncthreads();
  if (threadIdx.x == 0) {
    epoch += 1;
    __threadfence();
    atomicAdd(counter, 1u);
    const unsigned target = epoch * nblocks;
    while (ld_acquire_u32(counter) < target) { }
    __threadfence();
  }
  __syncthreads();
}

// LayerNorm statistics are exchanged as 64-bit fixed-point atomics: integer addition is associative, so the sum over the
// grid is deterministic (bit-identical run to run, unlike float atomics) and needs one 16-byte read per row afterwards.
__device__ __forceinline__ void fx_add(unsigned long long* p, float v, double scale) {
  const long long q = __double2ll_rn((double)v * scale);
  atomicAdd(p, (unsigned long long)q);
}
__device__ __forceinline__ float fx_get(const unsigned long long* p, double inv_scale) {
  unsigned long long u;
  asm volatile("ld.global.cg.u64 %0, [%1];" : "=l"(u) : "l"(p) : "memory");
  return (float)((double)(long long)u * inv_scale);
}
constexpr double FX_FWD = 17179869184.0;        // 2^34: |sum| < 5e8, resolution 6e-11
constexpr double FX_BWD = 17592186044416.0;     // 2^44: |sum| < 5e5, resolution 6e-14 (gradient-sized sums)

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
  unsigned long long* fx;     // workspace [T][B][2] fixed-point (sum, sum of squares), zeroed by the host
  unsigned* barrier;          // workspace counter, zeroed by the host
  long long* dbg;             // optional [T][6] globaltimer stamps of block 0 (profiling aid), or null
};

__device__ __forceinline__ long long gtime() { long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
#define SCAN_STAMP(i) do { if (p.dbg && blockIdx.x == 0 && threadIdx.x == 0) p.dbg[t * 6 + (i)] = gtime(); } while (0)

__global__ void __launch_bounds__(SCAN_THREADS, 1) rssm_scan_fwd_kernel(ScanFwdParams p) {
  extern __shared__ float smem[];
  const int D = p.D, D3 = 3 * p.D, B = p.B, T = p.T, G = p.G;
  constexpr int R = 3 * SCAN_U;
  float* Ws = smem;                       // [R][D]
  float* zs = Ws + (size_t)R * D;         // [B][R]  this CTA's pre-LN values for the current step
  float* st = zs + (size_t)B * R;         // [B][2]  mean / rstd
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int j0 = blockIdx.x * SCAN_U;
  unsigned epoch = 0;

  // resident weight slice: row r = gate*U + u  <->  W_h[gate*D + j0 + u, :]
  for (int i = tid; i < R * D; i += SCAN_THREADS) {
    int r = i / D, k = i % D;
    int gate = r / SCAN_U, u = r % SCAN_U;
    Ws[i] = p.Wh[(size_t)(gate * D + j0 + u) * p.ldw + k];
  }
  // masked initial state copy for the [x|h] operand (row t = 0)
  if (p.hprev_t) {
    for (int i = tid; i < B * SCAN_U; i += SCAN_THREADS) {
      int b = i / SCAN_U, u = i % SCAN_U;
      float m = p.mask ? p.mask[(size_t)b * T] : 1.f;
      float v = p.h0 ? p.h0[(size_t)b * D + j0 + u] * m : 0.f;
      st_any(p.hprev_t, p.t_dtype, (size_t)(b * (long long)T) * p.ld_hp + j0 + u, v);
    }
  }
  __syncthreads();

  const int ngroups = (B + SCAN_RB - 1) / SCAN_RB;
  for (int t = 0; t < T; ++t) {
    // ---- phase 1: zh = W_h (m * h_{t-1}) for every batch row ----
    SCAN_STAMP(0);
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
    __syncthreads();
    // ---- phase 2: LayerNorm partial sums of this CTA's 24 values per row ----
    SCAN_STAMP(1);
    unsigned long long* fx = p.fx + (size_t)t * B * 2;
    for (int b = tid; b < B; b += SCAN_THREADS) {
      float s1 = 0.f, s2 = 0.f;
#pragma unroll
      for (int r = 0; r < R; ++r) { float v = zs[b * R + r]; s1 += v; s2 += v * v; }
      fx_add(fx + b * 2, s1, FX_FWD);
      fx_add(fx + b * 2 + 1, s2, FX_FWD);
    }
    SCAN_STAMP(2);
    grid_barrier(p.barrier, epoch, G);
    SCAN_STAMP(3);
    // ---- phase 3: full-row statistics ----
    for (int b = tid; b < B; b += SCAN_THREADS) {
      float mean = fx_get(fx + b * 2, 1.0 / FX_FWD) / D3;
      float var = fmaxf(fx_get(fx + b * 2 + 1, 1.0 / FX_FWD) / D3 - mean * mean, 0.f);
      float rstd = 1.0f / sqrtf(var + 1e-5f);
      st[b * 2] = mean; st[b * 2 + 1] = rstd;
      if (blockIdx.x == 0) { p.stats[((size_t)b * T + t) * 2] = mean; p.stats[((size_t)b * T + t) * 2 + 1] = rstd; }
    }
    __syncthreads();
    // ---- phase 4: gates for the owned units ----
    SCAN_STAMP(4);
    for (int i = tid; i < B * SCAN_U; i += SCAN_THREADS) {
      int b = i / SCAN_U, u = i % SCAN_U, j = j0 + u;
      float mean = st[b * 2], rstd = st[b * 2 + 1];
      float pr = (zs[b * R + u] - mean) * rstd * p.gamma[j] + p.beta[j];
      float pc = (zs[b * R + SCAN_U + u] - mean) * rstd * p.gamma[D + j] + p.beta[D + j];
      float pu = (zs[b * R + 2 * SCAN_U + u] - mean) * rstd * p.gamma[2 * D + j] + p.beta[2 * D + j];
      float r = sigmoidf_(pr), cand = tanhf(r * pc), uu = sigmoidf_(pu - 1.0f);
      float m = p.mask ? p.mask[(size_t)b * T + t] : 1.f;
      float hp = t == 0 ? (p.h0 ? p.h0[(size_t)b * D + j] : 0.f) : __ldcg(p.deter + ((size_t)b * T + t - 1) * D + j);
      float hn = uu * cand + (1.0f - uu) * (hp * m);
      size_t n = (size_t)b * T + t;
      p.deter[n * D + j] = hn;
      if (p.deter_t) st_any(p.deter_t, p.t_dtype, n * p.ld_dt + j, hn);
      if (p.hprev_t && t + 1 < T) {
        float m1 = p.mask ? p.mask[(size_t)b * T + t + 1] : 1.f;
        st_any(p.hprev_t, p.t_dtype, (n + 1) * p.ld_hp + j, hn * m1);
      }
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
  float* dzf;                 // workspace [2][B][3D] fp32
  unsigned long long* fx;     // workspace [T][B][2] fixed-point, zeroed by the host
  unsigned* barrier;
};

__global__ void __launch_bounds__(SCAN_THREADS, 1) rssm_scan_bwd_kernel(ScanBwdParams p) {
  extern __shared__ float smem[];
  const int D = p.D, D3 = 3 * p.D, B = p.B, T = p.T, G = p.G;
  constexpr int R = 3 * SCAN_U;
  float* Wc = smem;                          // [U][3D]   Wc[u][i] = W_h[i, j0+u]
  float* gs = Wc + (size_t)SCAN_U * D3;      // [B][R]    gamma * gate-gradient
  float* xs = gs + (size_t)B * R;            // [B][R]    normalised z
  float* carry = xs + (size_t)B * R;         // [B][U]    gradient flowing into h_t from step t+1
  float* gate_keep = carry + (size_t)B * SCAN_U;   // [B][U]  (1-u_t) * dh_total
  float* st = gate_keep + (size_t)B * SCAN_U;      // [B][2]
  float* dgl = st + (size_t)B * 2;           // [2][R] local dgamma / dbeta accumulators
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int j0 = blockIdx.x * SCAN_U;
  unsigned epoch = 0;

  for (int i = tid; i < SCAN_U * D3; i += SCAN_THREADS) {
    int u = i / D3, row = i % D3;
    Wc[i] = p.Wh[(size_t)row * p.ldw + j0 + u];
  }
  for (int i = tid; i < B * SCAN_U; i += SCAN_THREADS) carry[i] = 0.f;
  for (int i = tid; i < 2 * R; i += SCAN_THREADS) dgl[i] = 0.f;
  __syncthreads();

  const int ngroups = (B + SCAN_RB - 1) / SCAN_RB;
  for (int t = T - 1; t >= 0; --t) {
    unsigned long long* fx = p.fx + (size_t)t * B * 2;
    float* dzf = p.dzf + (size_t)(t & 1) * B * D3;
    // ---- phase 1: gate gradients for owned units, LayerNorm-backward partial sums ----
    for (int i = tid; i < B * SCAN_U; i += SCAN_THREADS) {
      int b = i / SCAN_U, u = i % SCAN_U, j = j0 + u;
      size_t n = (size_t)b * T + t;
      float mean = p.stats[n * 2], rstd = p.stats[n * 2 + 1];
      float xr = (p.z[n * D3 + j] - mean) * rstd, xc = (p.z[n * D3 + D + j] - mean) * rstd, xu = (p.z[n * D3 + 2 * D + j] - mean) * rstd;
      float pr = xr * p.gamma[j] + p.beta[j], pc = xc * p.gamma[D + j] + p.beta[D + j], pu = xu * p.gamma[2 * D + j] + p.beta[2 * D + j];
      float r = sigmoidf_(pr), cand = tanhf(r * pc), uu = sigmoidf_(pu - 1.0f);
      float m = p.mask ? p.mask[n] : 1.f;
      float hp = (t == 0 ? (p.h0 ? p.h0[(size_t)b * D + j] : 0.f) : p.deter[(n - 1) * D + j]) * m;
      float g = p.dh_ext[n * D + j] + carry[i];
      float drc = g * uu * (1.0f - cand * cand);
      float gpr = drc * pc * r * (1.0f - r), gpc = drc * r, gpu = g * (cand - hp) * uu * (1.0f - uu);
      gs[b * R + u] = gpr * p.gamma[j]; gs[b * R + SCAN_U + u] = gpc * p.gamma[D + j]; gs[b * R + 2 * SCAN_U + u] = gpu * p.gamma[2 * D + j];
      xs[b * R + u] = xr; xs[b * R + SCAN_U + u] = xc; xs[b * R + 2 * SCAN_U + u] = xu;
      gate_keep[i] = g * (1.0f - uu);
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
      fx_add(fx + b * 2, s1, FX_BWD);
      fx_add(fx + b * 2 + 1, s2, FX_BWD);
    }
    grid_barrier(p.barrier, epoch, G);
    // ---- phase 2: dz for owned rows ----
    for (int b = tid; b < B; b += SCAN_THREADS) {
      st[b * 2] = fx_get(fx + b * 2, 1.0 / FX_BWD) / D3;
      st[b * 2 + 1] = fx_get(fx + b * 2 + 1, 1.0 / FX_BWD) / D3;
    }
    __syncthreads();
    for (int i = tid; i < B * R; i += SCAN_THREADS) {
      int b = i / R, r = i % R;
      int gate = r / SCAN_U, u = r % SCAN_U;
      size_t n = (size_t)b * T + t;
      float rstd = p.stats[n * 2 + 1];
      float dzv = rstd * (gs[i] - st[b * 2] - xs[i] * st[b * 2 + 1]);
      int col = gate * D + j0 + u;
      dzf[(size_t)b * D3 + col] = dzv;
      st_any(p.dz_t, p.t_dtype, n * p.ld_dz + col, dzv);
    }
    grid_barrier(p.barrier, epoch, G);
    // ---- phase 3: carry_{t-1}[b, units] = m_t * (W_h[:, units]^T dz_t + (1-u_t) dh_t) ----
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
        float m = p.mask ? p.mask[(size_t)b * T + t] : 1.f;
        float c = m * (tot + gate_keep[b * SCAN_U + u]);
        carry[b * SCAN_U + u] = c;
        if (t == 0 && p.dh0) p.dh0[(size_t)b * D + j0 + u] = c;
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

extern "C" long long genrl_rssm_scan_workspace_bytes(int B, int T, int D) {
  // [256 B barrier][T*B*2 u64 fixed-point statistics][2*B*3D fp32 dz staging (backward)]
  return 256 + (long long)T * B * 2 * 8 + (long long)2 * B * 3 * D * 4;
}

static long long* g_scan_dbg = nullptr;
// profiling aid: when set, block 0 of the forward scan records 6 globaltimer stamps per step into dbg[T][6]
extern "C" int genrl_rssm_scan_set_debug(long long* dbg) { g_scan_dbg = dbg; return 0; }

static int scan_check(int B, int T, int D, const char* who) {
  GENRL_REQUIRE(B > 0 && T > 0, "%s: empty batch", who);
  GENRL_REQUIRE(D % 32 == 0, "%s: deter size %d must be a multiple of 32", who, D);
  int sms = 0, dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  GENRL_REQUIRE(D / SCAN_U <= sms, "%s: deter size %d needs %d co-resident CTAs > %d SMs (use the per-step path)", who, D, D / SCAN_U, sms);
  return 0;
}

// workspace: genrl_rssm_scan_workspace_bytes() bytes; the first 256 bytes (barrier) are zeroed here.
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
  p.fx = reinterpret_cast<unsigned long long*>(reinterpret_cast<char*>(workspace) + 256);
  size_t smem = ((size_t)3 * SCAN_U * D + (size_t)B * 3 * SCAN_U + (size_t)B * 2) * sizeof(float);
  GENRL_REQUIRE(smem <= 227 * 1024, "rssm_scan_fwd: %zu bytes of shared memory needed (B=%d D=%d)", smem, B, D);
  cudaError_t e = cudaFuncSetAttribute(rssm_scan_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  GENRL_REQUIRE(e == cudaSuccess, "rssm_scan_fwd: smem attribute: %s", cudaGetErrorString(e));
  cudaMemsetAsync(workspace, 0, 256 + (size_t)T * B * 16, st);
  void* args[] = {&p};
  e = cudaLaunchCooperativeKernel((void*)rssm_scan_fwd_kernel, dim3(p.G), dim3(SCAN_THREADS), args, smem, st);
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
  p.fx = reinterpret_cast<unsigned long long*>(reinterpret_cast<char*>(workspace) + 256);
  p.dzf = reinterpret_cast<float*>(reinterpret_cast<char*>(workspace) + 256 + (size_t)T * B * 16);
  const int R = 3 * SCAN_U;
  size_t smem = ((size_t)SCAN_U * 3 * D + (size_t)2 * B * R + (size_t)2 * B * SCAN_U + (size_t)B * 2 + 2 * R) * sizeof(float);
  GENRL_REQUIRE(smem <= 227 * 1024, "rssm_scan_bwd: %zu bytes of shared memory needed (B=%d D=%d)", smem, B, D);
  cudaError_t e = cudaFuncSetAttribute(rssm_scan_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  GENRL_REQUIRE(e == cudaSuccess, "rssm_scan_bwd: smem attribute: %s", cudaGetErrorString(e));
  cudaMemsetAsync(workspace, 0, 256 + (size_t)T * B * 16, st);
  void* args[] = {&p};
  e = cudaLaunchCooperativeKernel((void*)rssm_scan_bwd_kernel, dim3(p.G), dim3(SCAN_THREADS), args, smem, st);
  GENRL_REQUIRE(e == cudaSuccess, "rssm_scan_bwd: cooperative launch failed: %s", cudaGetErrorString(e));
  return check_launch("rssm_scan_bwd");
}
