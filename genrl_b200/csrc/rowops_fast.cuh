// Vectorised fast paths of the HBM-bound row kernels (16-byte loads/stores, rows cached in registers, no atomics in
// the row pass; column reductions for parameter gradients run in a separate coalesced column kernel).
// Conditions for the fast paths are checked by the C entry points; otherwise the generic scalar kernels run.
#pragma once
#include "common.cuh"

namespace genrl {

// ---- 8-element vector access ----------------------------------------------------------------------------------
template <typename T> struct V8;
template <> struct V8<__nv_bfloat16> {
  static __device__ __forceinline__ void load(const __nv_bfloat16* p, float (&v)[8]) {
    uint4 u = *reinterpret_cast<const uint4*>(p);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
    for (int i = 0; i < 4; ++i) { float2 f = __bfloat1622float2(h[i]); v[2 * i] = f.x; v[2 * i + 1] = f.y; }
  }
  static __device__ __forceinline__ void store(__nv_bfloat16* p, const float (&v)[8]) {
    uint4 u;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
    *reinterpret_cast<uint4*>(p) = u;
  }
};
template <> struct V8<float> {
  static __device__ __forceinline__ void load(const float* p, float (&v)[8]) {
    float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  }
  static __device__ __forceinline__ void store(float* p, const float (&v)[8]) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
  }
};
__device__ __forceinline__ void ldf8(const float* p, float (&v)[8]) { V8<float>::load(p, v); }

template <typename T> inline bool vec_ok(const void* p, long long ld) {
  const int a = sizeof(T) == 2 ? 8 : 4;
  return (reinterpret_cast<uintptr_t>(p) & 15) == 0 && (ld % a) == 0;
}

static inline bool f32_vec_ok(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

__device__ __forceinline__ float group_sum(float v, int lpr) {
  for (int o = lpr >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}


// ---- LayerNorm (+SiLU) forward: LPR lanes per row, row in registers ------------------------------------------------
// NCH = chunks (of 8 elements) per lane: 4 for wide rows (up to 32*4*8 = 1024), 6 for the narrow conv rows (48..384 channels,
// 1..8 lanes per row: more rows per warp and more bytes in flight per lane)
template <typename T, int FAST_NCH>
__global__ void __launch_bounds__(256, (FAST_NCH == 4 ? 4 : 2)) ln_fwd_fast_kernel(const T* __restrict__ x, long long ldx, const float* __restrict__ gamma,
                                                          const float* __restrict__ beta, float eps, int act, T* __restrict__ y, long long ldy,
                                                          float* __restrict__ stats, long long M, int D, int lpr) {
  genrl::pdl_sync();
  const int lane = threadIdx.x & 31, sub = lane % lpr, rwi = lane / lpr, rpw = 32 / lpr;
  const long long wstride = (long long)gridDim.x * 8 * rpw;
  for (long long base = ((long long)blockIdx.x * 8 + (threadIdx.x >> 5)) * rpw; base < M; base += wstride) {
    const long long row = base + rwi;
    const bool valid = row < M;
    float v[FAST_NCH][8];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < FAST_NCH; ++i) {
      const int c = (i * lpr + sub) * 8;
      if (valid && c < D) {
        V8<T>::load(x + row * ldx + c, v[i]);
#pragma unroll
        for (int j = 0; j < 8; ++j) s += v[i][j];
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) v[i][j] = 0.f;
      }
    }
    const float mean = group_sum(s, lpr) / D;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < FAST_NCH; ++i) {
      const int c = (i * lpr + sub) * 8;
      if (c < D) {
#pragma unroll
        for (int j = 0; j < 8; ++j) { float d = v[i][j] - mean; q += d * d; }
      }
    }
    const float rstd = 1.0f / sqrtf(group_sum(q, lpr) / D + eps);
    if (valid) {
#pragma unroll
      for (int i = 0; i < FAST_NCH; ++i) {
        const int c = (i * lpr + sub) * 8;
        if (c < D) {
          float g[8], b[8], o[8];
          ldf8(gamma + c, g); ldf8(beta + c, b);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float a = (v[i][j] - mean) * rstd * g[j] + b[j];
            o[j] = act == GENRL_ACT_SILU ? silu_t<(sizeof(T) == 2)>(a) : a;
          }
          V8<T>::store(y + row * ldy + c, o);
        }
      }
      if (stats && sub == 0) { stats[row * 2] = mean; stats[row * 2 + 1] = rstd; }
    }
  }
}

// ---- LayerNorm (+SiLU) backward, row pass: dx only ------------------------------------------------------------------
// WPB = warps per block: 8 (two blocks per SM) for large row counts; 4 with five blocks per SM (20 resident rows-in-flight per SM at one
// row per warp) so that the 2 400-row launches of the imagination rollout's reverse loop are ONE wave instead of 296 + 4 blocks.
template <typename T, int FAST_NCH, int WPB>
__global__ void __launch_bounds__(WPB * 32, (WPB == 8 ? 2 : 5)) ln_bwd_fast_kernel(const T* __restrict__ dy, long long lddy, const T* __restrict__ x, long long ldx,
                                                          const float* __restrict__ gamma, const float* __restrict__ beta,
                                                          const float* __restrict__ stats, int act, T* __restrict__ dx, long long lddx,
                                                          long long M, int D, int lpr) {
  genrl::pdl_sync();
  const int lane = threadIdx.x & 31, sub = lane % lpr, rwi = lane / lpr, rpw = 32 / lpr;
  const long long wstride = (long long)gridDim.x * WPB * rpw;
  for (long long base = ((long long)blockIdx.x * WPB + (threadIdx.x >> 5)) * rpw; base < M; base += wstride) {
    const long long row = base + rwi;
    const bool valid = row < M;
    const float mean = valid ? stats[row * 2] : 0.f, rstd = valid ? stats[row * 2 + 1] : 0.f;
    float gh[FAST_NCH][8], xh[FAST_NCH][8];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < FAST_NCH; ++i) {
      const int c = (i * lpr + sub) * 8;
      if (valid && c < D) {
        float d[8], g[8], b[8];
        V8<T>::load(dy + row * lddy + c, d);
        V8<T>::load(x + row * ldx + c, xh[i]);
        ldf8(gamma + c, g);
        if (act == GENRL_ACT_SILU) ldf8(beta + c, b);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float h = (xh[i][j] - mean) * rstd;
          float gg = d[j];
          if (act == GENRL_ACT_SILU) gg *= silu_grad_t<(sizeof(T) == 2)>(h * g[j] + b[j]);
          gg *= g[j];
          xh[i][j] = h; gh[i][j] = gg;
          s1 += gg; s2 += gg * h;
        }
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) { gh[i][j] = 0.f; xh[i][j] = 0.f; }
      }
    }
    s1 = group_sum(s1, lpr) / D; s2 = group_sum(s2, lpr) / D;
    if (valid) {
#pragma unroll
      for (int i = 0; i < FAST_NCH; ++i) {
        const int c = (i * lpr + sub) * 8;
        if (c < D) {
          float o[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j] = rstd * (gh[i][j] - s1 - xh[i][j] * s2);
          V8<T>::store(dx + row * lddx + c, o);
        }
      }
    }
  }
}

// ---- LayerNorm (+SiLU) backward, row pass + parameter gradients in ONE pass ------------------------------------------
// Same row mapping as ln_bwd_fast_kernel; every lane additionally keeps the column partials of the columns it owns
// (dgamma += g*xhat, dbeta += g, dlin_bias += dx) in registers across its grid-stride rows; they are reduced over the rows
// of a warp by shuffles, over the warps of a block in shared memory, and leave the block as one fp32 atomic per column.
// Replaces the separate column kernel (a second read of dy, x, dx): traffic 3 tensors instead of 6.
template <typename T, int FAST_NCH>
__global__ void __launch_bounds__(256, (FAST_NCH <= 2 ? 2 : 1)) ln_bwd_colg_kernel(const T* __restrict__ dy, long long lddy, const T* __restrict__ x, long long ldx,
                                                          const float* __restrict__ gamma, const float* __restrict__ beta,
                                                          const float* __restrict__ stats, int act, T* __restrict__ dx, long long lddx,
                                                          float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ dlb,
                                                          long long M, int D, int lpr) {
  genrl::pdl_sync();
  __shared__ float red[3 * 1024];
  for (int i = threadIdx.x; i < 3 * D; i += 256) red[i] = 0.f;
  __syncthreads();
  const int lane = threadIdx.x & 31, sub = lane % lpr, rwi = lane / lpr, rpw = 32 / lpr;
  const long long wstride = (long long)gridDim.x * 8 * rpw;
  float a0[FAST_NCH][8], a1[FAST_NCH][8], a2[FAST_NCH][8];
#pragma unroll
  for (int i = 0; i < FAST_NCH; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) { a0[i][j] = 0.f; a1[i][j] = 0.f; a2[i][j] = 0.f; }
  for (long long base = ((long long)blockIdx.x * 8 + (threadIdx.x >> 5)) * rpw; base < M; base += wstride) {
    const long long row = base + rwi;
    const bool valid = row < M;
    const float mean = valid ? stats[row * 2] : 0.f, rstd = valid ? stats[row * 2 + 1] : 0.f;
    float gh[FAST_NCH][8], xh[FAST_NCH][8];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < FAST_NCH; ++i) {
      const int c = (i * lpr + sub) * 8;
      if (valid && c < D) {
        float d[8], g[8], b[8];
        V8<T>::load(dy + row * lddy + c, d);
        V8<T>::load(x + row * ldx + c, xh[i]);
        ldf8(gamma + c, g);
        if (act == GENRL_ACT_SILU) ldf8(beta + c, b);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float h = (xh[i][j] - mean) * rstd;
          float gg = d[j];
          if (act == GENRL_ACT_SILU) gg *= silu_grad_t<(sizeof(T) == 2)>(h * g[j] + b[j]);
          a0[i][j] += gg * h; a1[i][j] += gg;
          gg *= g[j];
          xh[i][j] = h; gh[i][j] = gg;
          s1 += gg; s2 += gg * h;
        }
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) { gh[i][j] = 0.f; xh[i][j] = 0.f; }
      }
    }
    s1 = group_sum(s1, lpr) / D; s2 = group_sum(s2, lpr) / D;
    if (valid) {
#pragma unroll
      for (int i = 0; i < FAST_NCH; ++i) {
        const int c = (i * lpr + sub) * 8;
        if (c < D) {
          float o[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) { o[j] = rstd * (gh[i][j] - s1 - xh[i][j] * s2); a2[i][j] += o[j]; }
          V8<T>::store(dx + row * lddx + c, o);
        }
      }
    }
  }
  // rows of one warp -> lanes rwi == 0
  for (int o = lpr; o < 32; o <<= 1) {
#pragma unroll
    for (int i = 0; i < FAST_NCH; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        a0[i][j] += __shfl_xor_sync(0xffffffffu, a0[i][j], o);
        a1[i][j] += __shfl_xor_sync(0xffffffffu, a1[i][j], o);
        a2[i][j] += __shfl_xor_sync(0xffffffffu, a2[i][j], o);
      }
  }
  if (rwi == 0) {
#pragma unroll
    for (int i = 0; i < FAST_NCH; ++i) {
      const int c = (i * lpr + sub) * 8;
      if (c < D) {
#pragma unroll
        for (int j = 0; j < 8; ++j) { atomicAdd(&red[c + j], a0[i][j]); atomicAdd(&red[D + c + j], a1[i][j]); atomicAdd(&red[2 * D + c + j], a2[i][j]); }
      }
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < D; c += 256) {
    if (dgamma) { atomicAdd(&dgamma[c], red[c]); atomicAdd(&dbeta[c], red[D + c]); }
    if (dlb) atomicAdd(&dlb[c], red[2 * D + c]);
  }
}

// ---- column pass of the LayerNorm backward: dgamma/dbeta/dlin_bias += column sums over a chunk of rows --------------
constexpr int COL_RC = 256;   // rows per block
template <typename T>
__global__ void __launch_bounds__(256) ln_colgrad_kernel(const T* __restrict__ dy, long long lddy, const T* __restrict__ x, long long ldx,
                                                         const T* __restrict__ dx, long long lddx, const float* __restrict__ gamma,
                                                         const float* __restrict__ beta, const float* __restrict__ stats, int act,
                                                         float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ dlb, long long M, int D) {
  genrl::pdl_sync();
  __shared__ float red[3][8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c = blockIdx.y * 32 + tx;
  const long long r0 = (long long)blockIdx.x * COL_RC, r1 = min(M, r0 + COL_RC);
  float a0 = 0.f, a1 = 0.f, a2 = 0.f;
  if (c < D) {
    const float g = gamma[c], b = beta[c];
#pragma unroll 8
    for (long long r = r0 + ty; r < r1; r += 8) {
      if (dgamma) {
        float h = (to_f32(x[r * ldx + c]) - stats[r * 2]) * stats[r * 2 + 1];
        float gg = to_f32(dy[r * lddy + c]);
        if (act == GENRL_ACT_SILU) gg *= silu_grad_t<(sizeof(T) == 2)>(h * g + b);
        a0 += gg * h; a1 += gg;
      }
      if (dlb) a2 += to_f32(dx[r * lddx + c]);
    }
  }
  red[0][ty][tx] = a0; red[1][ty][tx] = a1; red[2][ty][tx] = a2;
  __syncthreads();
  if (ty == 0 && c < D) {
    float s0 = 0.f, s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) { s0 += red[0][k][tx]; s1 += red[1][k][tx]; s2 += red[2][k][tx]; }
    if (dgamma) { atomicAdd(&dgamma[c], s0); atomicAdd(&dbeta[c], s1); }
    if (dlb) atomicAdd(&dlb[c], s2);
  }
}

template <typename T>
__global__ void __launch_bounds__(256) colsum_fast_kernel(const T* __restrict__ x, long long ldx, float* __restrict__ out, long long M, int D) {
  genrl::pdl_sync();
  __shared__ float red[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c = blockIdx.y * 32 + tx;
  const long long r0 = (long long)blockIdx.x * COL_RC, r1 = min(M, r0 + COL_RC);
  float a = 0.f;
  if (c < D) {
#pragma unroll 8
    for (long long r = r0 + ty; r < r1; r += 8) a += to_f32(x[r * ldx + c]);
  }
  red[ty][tx] = a;
  __syncthreads();
  if (ty == 0 && c < D) {
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += red[k][tx];
    atomicAdd(&out[c], s);
  }
}


// vectorised column pass: a block covers 128 rows x 128 columns, a thread 8 rows x 8 adjacent columns (16-byte loads)
constexpr int COLV_ROWS = 128;
template <typename T>
__global__ void __launch_bounds__(256) ln_colgrad_vec_kernel(const T* __restrict__ dy, long long lddy, const T* __restrict__ x, long long ldx,
                                                             const T* __restrict__ dx, long long lddx, const float* __restrict__ gamma,
                                                             const float* __restrict__ beta, const float* __restrict__ stats, int act,
                                                             float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ dlb, long long M, int D) {
  genrl::pdl_sync();
  __shared__ float red[3][16][129];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int c = blockIdx.y * 128 + tx * 8;
  const long long r0 = (long long)blockIdx.x * COLV_ROWS, r1 = min(M, r0 + COLV_ROWS);
  float a0[8], a1[8], a2[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { a0[j] = 0.f; a1[j] = 0.f; a2[j] = 0.f; }
  if (c < D) {
    float g[8], b[8];
    ldf8(gamma + c, g); ldf8(beta + c, b);
#pragma unroll 2
    for (long long r = r0 + ty; r < r1; r += 16) {
      if (dgamma) {
        float xv[8], dv[8];
        V8<T>::load(x + r * ldx + c, xv); V8<T>::load(dy + r * lddy + c, dv);
        const float mean = stats[r * 2], rstd = stats[r * 2 + 1];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float h = (xv[j] - mean) * rstd;
          float gg = dv[j];
          if (act == GENRL_ACT_SILU) gg *= silu_grad_t<(sizeof(T) == 2)>(h * g[j] + b[j]);
          a0[j] += gg * h; a1[j] += gg;
        }
      }
      if (dlb) { float v[8]; V8<T>::load(dx + r * lddx + c, v);
#pragma unroll
        for (int j = 0; j < 8; ++j) a2[j] += v[j]; }
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) { red[0][ty][tx * 8 + j] = a0[j]; red[1][ty][tx * 8 + j] = a1[j]; red[2][ty][tx * 8 + j] = a2[j]; }
  __syncthreads();
  for (int i = threadIdx.x; i < 3 * 128; i += 256) {
    const int a = i >> 7, cc = i & 127, col = blockIdx.y * 128 + cc;
    if (col >= D) continue;
    float sum = 0.f;
#pragma unroll
    for (int k = 0; k < 16; ++k) sum += red[a][k][cc];
    if (a == 0 && dgamma) atomicAdd(&dgamma[col], sum);
    if (a == 1 && dgamma) atomicAdd(&dbeta[col], sum);
    if (a == 2 && dlb) atomicAdd(&dlb[col], sum);
  }
}

template <typename T>
__global__ void __launch_bounds__(256) colsum_vec_kernel(const T* __restrict__ x, long long ldx, float* __restrict__ out, long long M, int D) {
  genrl::pdl_sync();
  __shared__ float red[16][129];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int c = blockIdx.y * 128 + tx * 8;
  const long long r0 = (long long)blockIdx.x * COLV_ROWS, r1 = min(M, r0 + COLV_ROWS);
  float a[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) a[j] = 0.f;
  if (c < D) {
#pragma unroll 2
    for (long long r = r0 + ty; r < r1; r += 16) { float v[8]; V8<T>::load(x + r * ldx + c, v);
#pragma unroll
      for (int j = 0; j < 8; ++j) a[j] += v[j]; }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) red[ty][tx * 8 + j] = a[j];
  __syncthreads();
  for (int i = threadIdx.x; i < 128; i += 256) {
    const int col = blockIdx.y * 128 + i;
    if (col >= D) continue;
    float sum = 0.f;
#pragma unroll
    for (int k = 0; k < 16; ++k) sum += red[k][i];
    atomicAdd(&out[col], sum);
  }
}

// ---- GRU tail, vectorised (one warp per row; gate parts re-read from L1) ---------------------------------------------
template <bool FAST>
__device__ __forceinline__ void gru_gate8(const float (&zr)[8], const float (&zc)[8], const float (&zu)[8], float mean, float rstd,
                                          const float* __restrict__ gamma, const float* __restrict__ beta, int D, int j,
                                          float (&r)[8], float (&cand)[8], float (&u)[8], float (&pc)[8], float (&xr)[8], float (&xc)[8], float (&xu)[8]) {
  float gr[8], gc[8], gu[8], br[8], bc[8], bu[8];
  ldf8(gamma + j, gr); ldf8(gamma + D + j, gc); ldf8(gamma + 2 * D + j, gu);
  ldf8(beta + j, br); ldf8(beta + D + j, bc); ldf8(beta + 2 * D + j, bu);
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    xr[k] = (zr[k] - mean) * rstd; xc[k] = (zc[k] - mean) * rstd; xu[k] = (zu[k] - mean) * rstd;
    float pr = xr[k] * gr[k] + br[k];
    pc[k] = xc[k] * gc[k] + bc[k];
    float pu = xu[k] * gu[k] + bu[k];
    r[k] = sigm_t<FAST>(pr); cand[k] = tanh_t<FAST>(r[k] * pc[k]); u[k] = sigm_t<FAST>(pu - 1.0f);
  }
}

template <typename T>
__global__ void __launch_bounds__(256) gru_fwd_fast_kernel(const T* __restrict__ z, long long ldz, const float* __restrict__ gamma,
                                                           const float* __restrict__ beta, const float* __restrict__ hprev, long long ldh,
                                                           float* __restrict__ hout, long long ldho, T* __restrict__ hout_t, long long ldht,
                                                           float* __restrict__ stats, long long M, int D) {
  genrl::pdl_sync();
  const int lane = threadIdx.x & 31;
  const int D3 = 3 * D;
  for (long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5); row < M; row += (long long)gridDim.x * 8) {
    const T* zr_ = z + row * ldz;
    float s = 0.f;
    for (int c = lane * 8; c < D3; c += 256) { float v[8]; V8<T>::load(zr_ + c, v);
#pragma unroll
      for (int k = 0; k < 8; ++k) s += v[k]; }
    const float mean = warp_sum(s) / D3;
    float q = 0.f;
    for (int c = lane * 8; c < D3; c += 256) { float v[8]; V8<T>::load(zr_ + c, v);
#pragma unroll
      for (int k = 0; k < 8; ++k) { float d = v[k] - mean; q += d * d; } }
    const float rstd = 1.0f / sqrtf(warp_sum(q) / D3 + 1e-5f);
    for (int j = lane * 8; j < D; j += 256) {
      float a[8], b[8], c[8], r[8], cand[8], u[8], pc[8], xr[8], xc[8], xu[8], h[8], o[8];
      V8<T>::load(zr_ + j, a); V8<T>::load(zr_ + D + j, b); V8<T>::load(zr_ + 2 * D + j, c);
      gru_gate8<(sizeof(T) == 2)>(a, b, c, mean, rstd, gamma, beta, D, j, r, cand, u, pc, xr, xc, xu);
      ldf8(hprev + row * ldh + j, h);
#pragma unroll
      for (int k = 0; k < 8; ++k) o[k] = u[k] * cand[k] + (1.0f - u[k]) * h[k];
      V8<float>::store(hout + row * ldho + j, o);
      if (hout_t) V8<T>::store(hout_t + row * ldht + j, o);
    }
    if (lane == 0) { stats[row * 2] = mean; stats[row * 2 + 1] = rstd; }
  }
}

// frozen-weights backward (no dgamma/dbeta): dz, dhprev
template <typename T>
__global__ void __launch_bounds__(256) gru_bwd_fast_kernel(const float* __restrict__ dh, long long lddh, const T* __restrict__ z, long long ldz,
                                                           const float* __restrict__ gamma, const float* __restrict__ beta,
                                                           const float* __restrict__ hprev, long long ldh, const float* __restrict__ stats,
                                                           T* __restrict__ dz, long long lddz, float* __restrict__ dhprev, long long lddhp, long long M, int D) {
  genrl::pdl_sync();
  const int lane = threadIdx.x & 31;
  const int D3 = 3 * D;
  for (long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5); row < M; row += (long long)gridDim.x * 8) {
    const T* zr_ = z + row * ldz;
    const float mean = stats[row * 2], rstd = stats[row * 2 + 1];
    float s1 = 0.f, s2 = 0.f;
    for (int pass = 0; pass < 2; ++pass) {
      const float m1 = s1, m2 = s2;          // valid in pass 1
      for (int j = lane * 8; j < D; j += 256) {
        float a[8], b[8], c[8], r[8], cand[8], u[8], pc[8], xr[8], xc[8], xu[8], h[8], g[8];
        float gr[8], gc[8], gu[8];
        V8<T>::load(zr_ + j, a); V8<T>::load(zr_ + D + j, b); V8<T>::load(zr_ + 2 * D + j, c);
        gru_gate8<(sizeof(T) == 2)>(a, b, c, mean, rstd, gamma, beta, D, j, r, cand, u, pc, xr, xc, xu);
        ldf8(hprev + row * ldh + j, h); ldf8(dh + row * lddh + j, g);
        ldf8(gamma + j, gr); ldf8(gamma + D + j, gc); ldf8(gamma + 2 * D + j, gu);
        float o1[8], o2[8], o3[8], o4[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          float drc = g[k] * u[k] * (1.0f - cand[k] * cand[k]);
          float ga = drc * pc[k] * r[k] * (1.0f - r[k]) * gr[k];
          float gb = drc * r[k] * gc[k];
          float gcu = g[k] * (cand[k] - h[k]) * u[k] * (1.0f - u[k]) * gu[k];
          if (pass == 0) { s1 += ga + gb + gcu; s2 += ga * xr[k] + gb * xc[k] + gcu * xu[k]; }
          else {
            o1[k] = rstd * (ga - m1 - xr[k] * m2); o2[k] = rstd * (gb - m1 - xc[k] * m2); o3[k] = rstd * (gcu - m1 - xu[k] * m2);
            o4[k] = g[k] * (1.0f - u[k]);
          }
        }
        if (pass == 1) {
          V8<T>::store(dz + row * lddz + j, o1); V8<T>::store(dz + row * lddz + D + j, o2); V8<T>::store(dz + row * lddz + 2 * D + j, o3);
          if (dhprev) V8<float>::store(dhprev + row * lddhp + j, o4);
        }
      }
      if (pass == 0) { s1 = warp_sum(s1) / D3; s2 = warp_sum(s2) / D3; }
    }
  }
}

// frozen-weights GRU backward, one 128-thread block per row of D = 1024: each thread owns 8 hidden units, gates and partial
// gradients stay in registers (z, h, dh are read once), the two LayerNorm row sums go through a 4-warp shared-memory reduction
template <typename T>
__global__ void __launch_bounds__(128) gru_bwd_row1024_kernel(const float* __restrict__ dh, long long lddh, const T* __restrict__ z, long long ldz,
                                                              const float* __restrict__ gamma, const float* __restrict__ beta,
                                                              const float* __restrict__ hprev, long long ldh, const float* __restrict__ stats,
                                                              T* __restrict__ dz, long long lddz, float* __restrict__ dhprev, long long lddhp, long long M,
                                                              const T* __restrict__ e1, long long lde1, const T* __restrict__ e2, long long lde2,
                                                              const T* __restrict__ e3, long long lde3) {
  // e1 / e2 / e3 (optional, activation dtype): further contributions to d loss / d h summed on the fly (the rollout's feat-slice gradient, the
  // h half of the previous step's GRU dgrad, the output block's dgrad) instead of accumulate-copy launches / fp32 accumulate epilogues
  genrl::pdl_sync();
  constexpr int D = 1024, D3 = 3072;
  __shared__ float red[2][4];
  const int j = threadIdx.x * 8, lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  float gr[8], gc[8], gu[8];
  ldf8(gamma + j, gr); ldf8(gamma + D + j, gc); ldf8(gamma + 2 * D + j, gu);
  for (long long row = blockIdx.x; row < M; row += gridDim.x) {
    const T* zr_ = z + row * ldz;
    const float mean = stats[row * 2], rstd = stats[row * 2 + 1];
    float a[8], b[8], c[8], r[8], cand[8], u[8], pc[8], xr[8], xc[8], xu[8], h[8], g[8];
    V8<T>::load(zr_ + j, a); V8<T>::load(zr_ + D + j, b); V8<T>::load(zr_ + 2 * D + j, c);
    ldf8(hprev + row * ldh + j, h); ldf8(dh + row * lddh + j, g);
    if (e1) { float t[8]; V8<T>::load(e1 + row * lde1 + j, t);
#pragma unroll
      for (int k = 0; k < 8; ++k) g[k] += t[k]; }
    if (e2) { float t[8]; V8<T>::load(e2 + row * lde2 + j, t);
#pragma unroll
      for (int k = 0; k < 8; ++k) g[k] += t[k]; }
    if (e3) { float t[8]; V8<T>::load(e3 + row * lde3 + j, t);
#pragma unroll
      for (int k = 0; k < 8; ++k) g[k] += t[k]; }
    gru_gate8<(sizeof(T) == 2)>(a, b, c, mean, rstd, gamma, beta, D, j, r, cand, u, pc, xr, xc, xu);
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float drc = g[k] * u[k] * (1.0f - cand[k] * cand[k]);
      a[k] = drc * pc[k] * r[k] * (1.0f - r[k]) * gr[k];              // reuse a,b,c as the three gate gradients
      b[k] = drc * r[k] * gc[k];
      c[k] = g[k] * (cand[k] - h[k]) * u[k] * (1.0f - u[k]) * gu[k];
      s1 += a[k] + b[k] + c[k]; s2 += a[k] * xr[k] + b[k] * xc[k] + c[k] * xu[k];
      h[k] = g[k] * (1.0f - u[k]);                                     // dhprev
    }
    s1 = warp_sum(s1); s2 = warp_sum(s2);
    __syncthreads();                                                   // red[] of the previous row is no longer read
    if (lane == 0) { red[0][w] = s1; red[1][w] = s2; }
    __syncthreads();
    const float m1 = (red[0][0] + red[0][1] + red[0][2] + red[0][3]) / D3, m2 = (red[1][0] + red[1][1] + red[1][2] + red[1][3]) / D3;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      a[k] = rstd * (a[k] - m1 - xr[k] * m2); b[k] = rstd * (b[k] - m1 - xc[k] * m2); c[k] = rstd * (c[k] - m1 - xu[k] * m2);
    }
    V8<T>::store(dz + row * lddz + j, a); V8<T>::store(dz + row * lddz + D + j, b); V8<T>::store(dz + row * lddz + 2 * D + j, c);
    if (dhprev) V8<float>::store(dhprev + row * lddhp + j, h);
  }
}

// ---- vectorised strided copy / row gather with cast -------------------------------------------------------------------
template <typename TI, typename TO>
__global__ void __launch_bounds__(256) gather_rows_fast_kernel(const TI* __restrict__ src, long long lds, const int* __restrict__ idx,
                                                               const float* __restrict__ scale, TO* __restrict__ dst, long long ldd, long long rows,
                                                               int cols8, int accumulate) {
  genrl::pdl_sync();
  const long long total = rows * cols8;
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
    const long long r = i / cols8; const int c = (int)(i % cols8) * 8;
    const long long sr = idx ? (long long)idx[r] : r;
    float v[8];
    if (sr < 0) {
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] = 0.f;
    } else {
      V8<TI>::load(src + sr * lds + c, v);
      if (scale) { const float s = scale[r];
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] *= s; }
    }
    if (accumulate) { float o[8]; V8<TO>::load(dst + r * ldd + c, o);
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] += o[k]; }
    V8<TO>::store(dst + r * ldd + c, v);
  }
}

}  // namespace genrl
