// Implicit-GEMM convolutions on tcgen05 for the Dreamer conv stacks (reference Encoder dreamer_utils.py:580-621: Conv2d k4 s2;
// Decoder :656-706: ConvTranspose2d k5/k6 s2), NHWC bf16 activations.  No im2col / col2im buffers: the "virtual im2col" tile
//     col[row, kidx] = in[n, yy*sy + a*ts, xx*sy + b*ts, c]        row = (n, yy, xx),  kidx = ((a*tw + b)*C + c)
// is gathered straight into the 128B-swizzled shared-memory layout UMMA expects by cp.async (16-byte channel chunks, zero fill
// outside the image), while the (small) filter matrix arrives by TMA.  One geometry covers every pass of both layer types:
//   gather form  (sy = 2, ts = +1, taps k x k)       : Conv2d forward, ConvTranspose2d dgrad           -> conv_fwd_kernel
//   sub-pixel form (sy = 1, ts = -1, taps T x T,     : ConvTranspose2d forward, Conv2d dgrad           -> conv_fwd_kernel
//      T = ceil(k/2), 4 output parity classes stacked along N: out[n, 2yy+py, 2xx+px, co])
//   weight gradients (gather form, K = pixels)       : dW[s, kidx] = sum_rows S[row, s] * col[row, kidx] -> conv_wgrad_kernel
// Warp roles: warp 0 = TMA (filters / S matrix), warp 1 = TMEM allocator + MMA issuer, warps 2..5 = cp.async gather producers;
// epilogue (tcgen05.ld -> bias -> NHWC bf16 or NCHW fp32 stores / fp32 atomics for split-K wgrad) on warps 6..9 of the persistent forward
// kernel (overlapped with the next tile through a second TMEM accumulator) and on the gather warps of the wgrad kernel.
#include "common.cuh"
#include "tc_ptx.cuh"
#include <algorithm>
#include <cstdlib>

namespace genrl {

constexpr int CV_BM = 128;
constexpr int CV_BK = 64;
constexpr int CV_THREADS = 192;
constexpr int CV_GATHER = 128;     // gather threads of the wgrad kernel (warps 2..5)
constexpr int CVF_GATHER = 256;    // gather threads of the forward kernel (warps 2..9)
constexpr int CV_LAG = 2;          // cp.async groups kept in flight per gather thread before the stage is published
constexpr int CVF_KTAB_KB = 32;    // k-blocks covered by the forward kernel's shared-memory k table (8 entries of 8 bytes per k-block)

// .ca: every input pixel is gathered by several taps of the same CTA (4 for k4/s2, 9 for the 3x3 sub-pixel taps); the CTA's input
// footprint (~20 KB) stays in L1, so the re-reads do not go back to L2
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// exact x / d for x < 2^24, d < 2^12 as one 64-bit multiply: q = (x * M) >> 40, M = floor(2^40 / d) + 1   (M = 0: plain division)
__device__ __forceinline__ uint32_t fast_div(uint32_t x, uint32_t d, unsigned long long M) {
  return M ? (uint32_t)(((unsigned long long)x * M) >> 40) : x / d;
}
static unsigned long long div_magic(long long xmax, int d) {
  return (xmax < (1LL << 24) && d > 0 && d < (1 << 12)) ? ((1ULL << 40) / (unsigned long long)d + 1ULL) : 0ULL;
}

struct ConvGeom {
  const __nv_bfloat16* in;   // [Nimg, IH, IW, C] NHWC
  int Nimg, IH, IW, C;
  int RH, RW;                // row grid per image
  int sy, ts, th, tw;        // input stride of the row grid, tap direction (+1/-1), taps along y / x
  int K;                     // th*tw*C
  long long rows;            // Nimg*RH*RW
  unsigned long long m_rhw, m_rw, m_c, m_tw;   // fast_div magics for RH*RW, RW, C, tw
  int all_inside;            // gather form whose every tap of every valid row lies inside the image (stride-2 convs without padding): constant masks
};

// row -> (n, yy, xx) of the row grid
__device__ __forceinline__ void decode_row(const ConvGeom& g, long long row, int rhw, long long& n, int& yy, int& xx) {
  const uint32_t r = (uint32_t)row;
  const uint32_t nn = g.m_rhw ? fast_div(r, (uint32_t)rhw, g.m_rhw) : (uint32_t)(row / rhw);
  const uint32_t rem = (uint32_t)(row - (long long)nn * rhw);
  const uint32_t y = fast_div(rem, (uint32_t)g.RW, g.m_rw);
  n = nn; yy = (int)y; xx = (int)(rem - y * g.RW);
}
// per-row gather state: 32-bit element offset of the row's base pixel and the bitmask of taps that fall inside the image
// (bits 0..7: tap rows a, bits 8..15: tap columns b); off < 0 marks a row past the end of the row grid
__device__ __forceinline__ void row_state(const ConvGeom& g, long long row, int rhw, int& off, int& mask) {
  off = -1; mask = 0;
  if (row < g.rows) {
    long long n; int yy, xx;
    decode_row(g, row, rhw, n, yy, xx);
    const int iy0 = yy * g.sy, ix0 = xx * g.sy;
    off = (int)(((n * g.IH + iy0) * g.IW + ix0) * g.C);
    if (g.all_inside) { mask = ((1 << g.th) - 1) | (((1 << g.tw) - 1) << 8); return; }
    // taps a with 0 <= iy0 + a*ts < IH form a contiguous range [lo, hi]
    const int ylo = g.ts > 0 ? 0 : max(0, iy0 - g.IH + 1), yhi = g.ts > 0 ? min(g.th, g.IH - iy0) - 1 : min(g.th - 1, iy0);
    const int xlo = g.ts > 0 ? 0 : max(0, ix0 - g.IW + 1), xhi = g.ts > 0 ? min(g.tw, g.IW - ix0) - 1 : min(g.tw - 1, ix0);
    const int my = yhi >= ylo ? (int)(((2u << yhi) - 1u) & ~((1u << ylo) - 1u)) : 0;
    const int mx = xhi >= xlo ? (int)(((2u << xhi) - 1u) & ~((1u << xlo) - 1u)) : 0;
    mask = my | (mx << 8);
  }
}
// kidx -> (tap row a, tap col b, channel c)
__device__ __forceinline__ void decode_k(const ConvGeom& g, int kidx, int& a, int& b, int& c) {
  const uint32_t tap = fast_div((uint32_t)kidx, (uint32_t)g.C, g.m_c);
  c = kidx - (int)tap * g.C;
  a = (int)fast_div(tap, (uint32_t)g.tw, g.m_tw);
  b = (int)tap - a * g.tw;
}

struct ConvFwdParams {
  ConvGeom g;
  int N, Cpad, Cout, ncls;   // N = ncls*Cpad output columns; class = (py, px) of the sub-pixel form
  int OH, OW;
  int out_mode;              // 0: NHWC bf16 rows [Nimg*OH*OW, Cout]; 1: NCHW fp32 [Nimg, Cout, OH, OW]
  void* out;
  const float* bias;
  int num_kb;
};

struct ConvWgParams {
  ConvGeom g;
  int Cs;                    // channels of the row-aligned matrix S [rows, Cs]
  float* dw; int ldw;        // [Cs, K] fp32, accumulated atomically
  int num_kb, kb_per_split;  // k-blocks of 64 rows
};

// publish cp.async-filled stages: the writes of one group are complete for this thread after wait_group, become visible to the
// tensor-core (async) proxy after the proxy fence, then the thread arrives on the stage's full barrier
__device__ __forceinline__ void publish_stage(uint64_t* bar) {
  fence_proxy_async();
  mbar_arrive(bar);
}

// after the last issued k-block: publish the D = LAG .. 1 most recent stages in order
template <int D, int STAGES> struct DrainStages {
  static __device__ __forceinline__ void run(long long gs, uint64_t* full_bar) {
    if (gs >= D) { cp_async_wait<D - 1>(); publish_stage(&full_bar[(int)((gs - D) % STAGES)]); }
    DrainStages<D - 1, STAGES>::run(gs, full_bar);
  }
};
template <int STAGES> struct DrainStages<0, STAGES> {
  static __device__ __forceinline__ void run(long long, uint64_t*) {}
};

// shared-memory stages of the forward kernel.  The L1 that is left over matters: every input pixel is re-read by 4..9 taps, and
// the 5/6-stage variants measured 20-30 % slower on the wide tiles than 4 stages (profiles/r01_notes.md §3)
template <int BN> struct FwdCfg {
  static constexpr int STAGES = BN <= 96 ? 3 : 4;      // measured: deeper rings do not help (the gather is not latency bound) and cost L1 capacity
  static constexpr int LAG = STAGES - 1;
  static constexpr int CTAS = BN <= 96 ? 2 : 1;
};

// =================================================================================================
// forward-type kernel: D[row, col] = sum_k col[row, k] * Wm[col, k]
// Persistent: each CTA walks tiles t = blockIdx.x, +gridDim.x, ...; two TMEM accumulators so the epilogue warps (10..13) drain tile i
// while the gather warps (2..9), the TMA lane and the MMA lane already work on tile i+1 (most of these layers have only 4..27 k-blocks
// per tile, so an un-overlapped epilogue + CTA prologue would cost as much as the main loop).
// =================================================================================================
constexpr int CVF_THREADS = 448;     // TMA warp, MMA warp, 8 gather warps, 4 epilogue warps

// Registers are allocated per 4 warps: a 448-thread CTA occupies 16 warp slots, so two CTAs per SM need <= 64 registers per thread
// (tools/occ_probe.cu; __launch_bounds__(448, 2) alone lets ptxas use 72 and silently drops to one CTA per SM).
template <int BN, int STAGES>
__global__ void __launch_bounds__(CVF_THREADS) __maxnreg__(BN <= 96 ? 64 : 128)
conv_fwd_kernel(const __grid_constant__ CUtensorMap tmW, const ConvFwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  constexpr int LAG = STAGES - 1;
  constexpr int A_BYTES = CV_BM * CV_BK * 2;
  constexpr int B_BYTES = BN * CV_BK * 2;
  constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  constexpr int ACC_COLS = BN <= 32 ? 32 : BN <= 64 ? 64 : BN <= 128 ? 128 : 256;      // TMEM columns of one accumulator
  constexpr int TMEM_COLS = 2 * ACC_COLS;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;          // [2] accumulator ready for the epilogue
  uint64_t* tempty_bar = tfull_bar + 2;              // [2] accumulator drained
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
  // (element offset of the tap + channel chunk, tap bit pair) of every 16-byte chunk of every k-block: decoded ONCE per CTA instead of once
  // per thread and k-block (two multiply-high divisions + a dozen integer operations in the gather warps' inner loop, which is issue-bound)
  int2* s_ktab = reinterpret_cast<int2*>(smem + STAGES * STAGE_BYTES + 256);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int nkb = p.num_kb;
  const int tiles_n = (p.N + BN - 1) / BN;
  const long long tiles = ((p.g.rows + CV_BM - 1) / CV_BM) * tiles_n;
  const bool use_ktab = nkb <= CVF_KTAB_KB;
  if (use_ktab) {
    for (int i = threadIdx.x; i < nkb * 8; i += CVF_THREADS) {
      const int kidx = (i >> 3) * CV_BK + (i & 7) * 8;
      int a, b, c;
      decode_k(p.g, kidx, a, b, c);
      s_ktab[i] = make_int2((a * p.g.IW + b) * p.g.ts * p.g.C + c, kidx < p.g.K ? ((1 << a) | (256 << b)) : 0x10000);
    }
  }

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmW);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], CVF_GATHER + 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], 128);
    }
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
    // ------------------------------ TMA: filter k-blocks ------------------------------
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (long long t = blockIdx.x; t < tiles; t += gridDim.x) {
        const int n0 = (int)(t % tiles_n) * BN;
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait_sleep(&empty_bar[stage], phase ^ 1, 64);
          mbar_expect_tx(&full_bar[stage], B_BYTES);
          tma_load_2d(&tmW, &full_bar[stage], smem + stage * STAGE_BYTES + A_BYTES, kb * CV_BK, n0);     // box {64 k, BN n}
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------ MMA issuer ------------------------------
    if (lane == 0) {
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(CV_BM >> 4) << 24);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (long long t = blockIdx.x; t < tiles; t += gridDim.x, ++it) {
        const int buf = it & 1;
        mbar_wait_sleep(&tempty_bar[buf], ((it >> 1) & 1) ^ 1, 32);              // epilogue has drained this accumulator (first use passes)
        tc_fence_after();
        const uint32_t acc = tmem_base + (uint32_t)(buf * ACC_COLS);
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait_sleep(&full_bar[stage], phase, 20);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * STAGE_BYTES);
          const uint32_t sb = sa + A_BYTES;
#pragma unroll
          for (int k = 0; k < CV_BK / 16; ++k)
            umma_bf16(acc, make_smem_desc(sa + k * 32, 16, 1024), make_smem_desc(sb + k * 32, 16, 1024), idesc, (kb | k) != 0 ? 1u : 0u);
          umma_commit(&empty_bar[stage]);
          if (kb == nkb - 1) umma_commit(&tfull_bar[buf]);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp < 10) {
    // ------------------------------ gather producers (warps 2..9): thread = (16-byte chunk j, rows rg + 32 i) ------------------------------
    const ConvGeom& g = p.g;
    const int t256 = threadIdx.x - 64;
    const int j = t256 & 7;                  // 16-byte chunk of the 64-wide k block
    const int rg = t256 >> 3;                // rows rg + 32*i, i < 4
    const int rhw = g.RH * g.RW;
    const uint32_t dst_off = (uint32_t)(rg * 128 + ((j ^ (rg & 7)) << 4));
    int stage = 0;
    uint32_t phase = 0;
    long long gs = 0;                        // k-blocks issued so far (across tiles): publishing lags LAG behind issuing
    for (long long t = blockIdx.x; t < tiles; t += gridDim.x) {
      const long long m0 = (t / tiles_n) * CV_BM;
      int off[4], mask[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) row_state(g, m0 + rg + 32 * i, rhw, off[i], mask[i]);
      for (int kb = 0; kb < nkb; ++kb, ++gs) {
        mbar_wait_sleep(&empty_bar[stage], phase ^ 1, 32);
        int koff, kbit;
        if (use_ktab) {
          const int2 kt = s_ktab[kb * 8 + j];
          koff = kt.x; kbit = kt.y;
        } else {
          const int kidx = kb * CV_BK + j * 8;
          int a, b, c;
          decode_k(g, kidx, a, b, c);
          koff = (a * g.IW + b) * g.ts * g.C + c;
          kbit = kidx < g.K ? ((1 << a) | (256 << b)) : 0x10000;       // bit 16 is never set in a row mask
        }
        const uint32_t sa = smem_u32(smem + stage * STAGE_BYTES) + dst_off;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const bool ok = (mask[i] & kbit) == kbit;
          cp_async16(sa + i * (32 * 128), g.in + (ok ? off[i] + koff : 0), ok ? 16u : 0u);
        }
        cp_async_commit();
        if (gs >= LAG) {
          cp_async_wait<LAG>();
          publish_stage(&full_bar[(int)((gs - LAG) % STAGES)]);
        }
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
    DrainStages<LAG, STAGES>::run(gs, full_bar);
  } else {
    // ------------------------------ epilogue (warps 10..13) ------------------------------
    const ConvGeom& g = p.g;
    const int quarter = warp & 3;            // TMEM lane quarter this warp may access
    const int rhw = g.RH * g.RW;
    int it = 0;
    for (long long t = blockIdx.x; t < tiles; t += gridDim.x, ++it) {
      const int buf = it & 1;
      const long long m0 = (t / tiles_n) * CV_BM;
      const int n0 = (int)(t % tiles_n) * BN;
      mbar_wait_sleep(&tfull_bar[buf], (it >> 1) & 1, 128);
      tc_fence_after();
      const uint32_t acc = tmem_base + (uint32_t)(buf * ACC_COLS) + ((uint32_t)(quarter * 32) << 16);
      const long long row = m0 + quarter * 32 + lane;
      const bool row_ok = row < g.rows;
      long long n = 0; int yy = 0, xx = 0;
      if (row_ok) decode_row(g, row, rhw, n, yy, xx);
      const long long obase = p.ncls > 1 ? (n * p.OH + 2 * yy) * p.OW + 2 * xx : (n * p.OH + yy) * p.OW + xx;      // output pixel of class 0
#pragma unroll 1
      for (int cc = 0; cc < BN; cc += 32) {
        if (n0 + cc >= p.N) break;
        uint32_t v[32];
        tmem_ld32(acc + (uint32_t)cc, v);
        if (!row_ok) continue;
        if (p.ncls == 1 && p.out_mode == 0 && (p.Cout & 7) == 0 && p.Cpad == p.Cout && p.OH == g.RH && p.OW == g.RW) {
          // plain NHWC rows (gather form): output row = GEMM row, 8-channel groups are 16-byte aligned
          __nv_bfloat16* orow = reinterpret_cast<__nv_bfloat16*>(p.out) + row * p.Cout + n0 + cc;
#pragma unroll
          for (int h = 0; h < 4; ++h) {
            if (n0 + cc + 8 * h >= p.Cout || cc + 8 * h >= BN) continue;
            float f[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) f[q] = __uint_as_float(v[8 * h + q]);
            if (p.bias != nullptr) {
#pragma unroll
              for (int q = 0; q < 8; ++q) f[q] += __ldg(p.bias + n0 + cc + 8 * h + q);
            }
            uint4 pk;
            __nv_bfloat162 a0 = __floats2bfloat162_rn(f[0], f[1]), a1 = __floats2bfloat162_rn(f[2], f[3]);
            __nv_bfloat162 a2 = __floats2bfloat162_rn(f[4], f[5]), a3 = __floats2bfloat162_rn(f[6], f[7]);
            pk.x = *reinterpret_cast<uint32_t*>(&a0); pk.y = *reinterpret_cast<uint32_t*>(&a1);
            pk.z = *reinterpret_cast<uint32_t*>(&a2); pk.w = *reinterpret_cast<uint32_t*>(&a3);
            *reinterpret_cast<uint4*>(orow + 8 * h) = pk;
          }
          continue;
        }
        // generic path (sub-pixel classes stacked along N, NCHW fp32 output, ragged channel counts): class / channel of the chunk's first
        // column by ONE division, then carried incrementally over the four 8-column groups (Cpad is a multiple of 8); the pixel index of
        // class 0 is computed once per tile (obase)
        int cls = (n0 + cc) / p.Cpad;
        int co0 = (n0 + cc) - cls * p.Cpad;
#pragma unroll
        for (int h = 0; h < 4; ++h, co0 += 8) {
          if (n0 + cc + 8 * h >= p.N || cc + 8 * h >= BN) continue;
          if (co0 >= p.Cpad) { co0 -= p.Cpad; ++cls; }
          const int nco = min(8, p.Cout - co0);
          if (nco <= 0) continue;
          int oy = yy, ox = xx;
          long long pix = obase;
          if (p.ncls > 1) { oy = 2 * yy + (cls >> 1); ox = 2 * xx + (cls & 1); pix += (long long)(cls >> 1) * p.OW + (cls & 1); }
          if (oy >= p.OH || ox >= p.OW) continue;
          float f[8];
#pragma unroll
          for (int q = 0; q < 8; ++q) f[q] = __uint_as_float(v[8 * h + q]);
          if (p.bias != nullptr) {
            if (nco == 8) {
#pragma unroll
              for (int q = 0; q < 8; ++q) f[q] += __ldg(p.bias + co0 + q);
            } else {
#pragma unroll
              for (int q = 0; q < 8; ++q) if (q < nco) f[q] += __ldg(p.bias + co0 + q);
            }
          }
          if (p.out_mode == 0) {
            __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(p.out) + pix * p.Cout + co0;
            if (nco == 8 && (reinterpret_cast<uintptr_t>(o) & 15) == 0) {
              uint4 pk;
              __nv_bfloat162 a0 = __floats2bfloat162_rn(f[0], f[1]), a1 = __floats2bfloat162_rn(f[2], f[3]);
              __nv_bfloat162 a2 = __floats2bfloat162_rn(f[4], f[5]), a3 = __floats2bfloat162_rn(f[6], f[7]);
              pk.x = *reinterpret_cast<uint32_t*>(&a0); pk.y = *reinterpret_cast<uint32_t*>(&a1);
              pk.z = *reinterpret_cast<uint32_t*>(&a2); pk.w = *reinterpret_cast<uint32_t*>(&a3);
              *reinterpret_cast<uint4*>(o) = pk;
            } else {
#pragma unroll
              for (int q = 0; q < 8; ++q) if (q < nco) o[q] = __float2bfloat16_rn(f[q]);
            }
          } else {
            float* o = reinterpret_cast<float*>(p.out) + ((n * p.Cout + co0) * p.OH + oy) * p.OW + ox;
#pragma unroll
            for (int q = 0; q < 8; ++q) if (q < nco) o[(long long)q * p.OH * p.OW] = f[q];
          }
        }
      }
      tc_fence_before();
      mbar_arrive(&tempty_bar[buf]);           // 128 arrivals: the accumulator may be overwritten
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
// weight-gradient kernel: dW[s, kidx] += sum_{rows of this split} S[row, s] * col[row, kidx]
// A = S^T (MN-major, TMA), B = col (MN-major, gathered: 64 pixel rows x BN k-columns per stage)
// =================================================================================================
template <int BN, int STAGES>
__global__ void __launch_bounds__(CV_THREADS, 2)
conv_wgrad_kernel(const __grid_constant__ CUtensorMap tmS, const ConvWgParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  constexpr int A_BYTES = CV_BM * CV_BK * 2;
  constexpr int B_BYTES = BN * CV_BK * 2;
  constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  constexpr int NG = BN / 64;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n0 = blockIdx.x * BN;
  const int m0 = blockIdx.y * CV_BM;
  const int kb_begin = blockIdx.z * p.kb_per_split;
  const int kb_end = min(p.num_kb, kb_begin + p.kb_per_split);
  const int nkb = kb_end - kb_begin;
  constexpr int TMEM_COLS = BN <= 64 ? 64 : BN <= 128 ? 128 : 256;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmS);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], CV_GATHER + 1);
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
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int kb = kb_begin; kb < kb_end; ++kb) {
        mbar_wait_sleep(&empty_bar[stage], phase ^ 1, 64);
        uint8_t* sa = smem + stage * STAGE_BYTES;
        mbar_expect_tx(&full_bar[stage], A_BYTES);
#pragma unroll
        for (int q = 0; q < CV_BM / 64; ++q)                                     // box {64 s-channels, 64 rows}
          tma_load_2d(&tmS, &full_bar[stage], sa + q * (64 * CV_BK * 2), m0 + 64 * q, kb * CV_BK);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(CV_BM >> 4) << 24);
      int stage = 0;
      uint32_t phase = 0;
      for (int kb = 0; kb < nkb; ++kb) {
        mbar_wait_sleep(&full_bar[stage], phase, 20);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + stage * STAGE_BYTES);
        const uint32_t sb = sa + A_BYTES;
#pragma unroll
        for (int k = 0; k < CV_BK / 16; ++k)
          umma_bf16(tmem_base, make_smem_desc(sa + k * 2048, 64 * CV_BK * 2, 1024), make_smem_desc(sb + k * 2048, 64 * CV_BK * 2, 1024), idesc,
                    (kb | k) != 0 ? 1u : 0u);
        umma_commit(&empty_bar[stage]);
        if (kb == nkb - 1) umma_commit(tmem_full_bar);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else {
    // ------------------------------ gather producers ------------------------------
    const ConvGeom& g = p.g;
    const int t = threadIdx.x - 64;
    const int j = t & 7;
    const int rg = t >> 3;                   // rows rg + 16*i, i < 4
    int koff[NG];                            // element offset of this thread's k-chunk relative to the row's base pixel
    bool kval[NG];
#pragma unroll
    for (int q = 0; q < NG; ++q) {
      const int kidx = n0 + q * 64 + j * 8;
      kval[q] = kidx < g.K;
      int a, b, c;
      decode_k(g, kidx, a, b, c);
      koff[q] = (a * g.IW + b) * g.C + c;    // gather geometry (ts = +1, no padding): every tap of a valid row lies inside the image
    }
    const int rhw = g.RH * g.RW;
    const uint32_t dst_off = (uint32_t)(rg * 128 + ((j ^ (rg & 7)) << 4));
    int stage = 0;
    uint32_t phase = 0;
    for (int kb = 0; kb < nkb; ++kb) {
      mbar_wait_sleep(&empty_bar[stage], phase ^ 1, 32);
      const uint32_t sb = smem_u32(smem + stage * STAGE_BYTES + A_BYTES) + dst_off;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const long long row = (long long)(kb_begin + kb) * CV_BK + rg + 16 * i;
        int base = 0;
        const bool rok = row < g.rows;
        if (rok) {
          long long n; int yy, xx;
          decode_row(g, row, rhw, n, yy, xx);
          base = (int)(((n * g.IH + yy * g.sy) * g.IW + xx * g.sy) * g.C);
        }
#pragma unroll
        for (int q = 0; q < NG; ++q) {
          const bool ok = rok && kval[q];
          cp_async16(sb + q * 8192 + i * (16 * 128), g.in + (ok ? base + koff[q] : 0), ok ? 16u : 0u);
        }
      }
      cp_async_commit();
      if (kb >= CV_LAG) {
        cp_async_wait<CV_LAG>();
        publish_stage(&full_bar[(kb - CV_LAG) % STAGES]);
      }
      if (++stage == STAGES) { stage = 0; phase ^= 1; }
    }
    if (nkb >= 2) { cp_async_wait<1>(); publish_stage(&full_bar[(nkb - 2) % STAGES]); }
    cp_async_wait<0>();
    publish_stage(&full_bar[(nkb - 1) % STAGES]);

    // ------------------------------ epilogue: fp32 atomics into dW ------------------------------
    const int quarter = warp & 3;
    mbar_wait_sleep(tmem_full_bar, 0, 64);
    tc_fence_after();
    const int m = m0 + quarter * 32 + lane;
    const bool m_ok = m < p.Cs;
#pragma unroll 1
    for (int cc = 0; cc < BN; cc += 32) {
      if (n0 + cc >= g.K) break;
      uint32_t v[32];
      tmem_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)cc, v);
      if (!m_ok) continue;
      float* o = p.dw + (long long)m * p.ldw + n0 + cc;
      const int ncols = min(32, g.K - (n0 + cc));
#pragma unroll
      for (int q = 0; q < 32; ++q) if (q < ncols) atomicAdd(o + q, __uint_as_float(v[q]));
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
// filter packing
// =================================================================================================
// sub-pixel form: Wq[(cls*Cpad + b), ((a_*T + b_)*A + a)] = W4[a][b][py + 2a_][px + 2b_]   (0 outside the filter / for b >= B)
// W4 is [A][B][k][k] fp32: Conv2d weight [Cout][Cin] (A = Cout, used by its dgrad) or ConvTranspose2d weight [Cin][Cout] (A = Cin, forward)
__global__ void __launch_bounds__(256) conv_pack_parity_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ dst, int A, int B, int k, int T,
                                                               int Cpad) {
  genrl::pdl_sync();
  const long long K = (long long)T * T * A;
  const long long total = 4LL * Cpad * K;
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
    const int a = (int)(i % A);
    const int tap = (int)((i / A) % (T * T));
    const long long r = i / K;
    const int b = (int)(r % Cpad), cls = (int)(r / Cpad);
    const int ky = (cls >> 1) + 2 * (tap / T), kx = (cls & 1) + 2 * (tap % T);
    float v = 0.f;
    if (b < B && ky < k && kx < k) v = w[(((long long)a * B + b) * k + ky) * k + kx];
    dst[i] = __float2bfloat16_rn(v);
  }
}

// gather form with channel padding: Wg[a, (ky*k + kx)*Cp + b] = W4[a][b][ky][kx]  (0 for b >= B)
__global__ void __launch_bounds__(256) conv_pack_gather_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ dst, int A, int B, int k, int Cp) {
  genrl::pdl_sync();
  const long long total = (long long)A * k * k * Cp;
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
    const int b = (int)(i % Cp);
    const int tap = (int)((i / Cp) % (k * k));
    const long long a = i / ((long long)Cp * k * k);
    dst[i] = __float2bfloat16_rn(b < B ? w[((a * B + b) * k * k) + tap] : 0.f);
  }
}

// grad[a][b][tap] += dwp[a, tap*Cp + b]
__global__ void __launch_bounds__(256) conv_unpack_wgrad_kernel(const float* __restrict__ dwp, long long ld, float* __restrict__ grad, int A, int B, int kk,
                                                                int Cp) {
  genrl::pdl_sync();
  const long long total = (long long)A * B * kk;
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
    const int tap = (int)(i % kk);
    const int b = (int)((i / kk) % B);
    const long long a = i / ((long long)kk * B);
    grad[i] += dwp[a * ld + (long long)tap * Cp + b];
  }
}

// NCHW fp32 [N,C,H,W] -> NHWC bf16 with the channel dimension zero-padded to Cp
__global__ void __launch_bounds__(256) nchw_to_nhwc_pad_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ y, long long N, int C, int H, int W,
                                                               int Cp) {
  genrl::pdl_sync();
  const long long total = N * H * W;
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
    const long long n = i / ((long long)H * W);
    const long long hw = i - n * H * W;
    for (int c = 0; c < Cp; ++c) y[i * Cp + c] = __float2bfloat16_rn(c < C ? x[(n * C + c) * H * W + hw] : 0.f);
  }
}

// uint8 NCHW observation -> (x/255 - 0.5) bf16 NHWC with the channel dimension zero-padded to 8 (one 16-byte store per pixel):
// the encoder's first layer then runs on the implicit-GEMM path like every other layer (preprocess dreamer.py:294-295 fused here)
__global__ void __launch_bounds__(256) u8_nchw_to_nhwc8_kernel(const uint8_t* __restrict__ x, uint4* __restrict__ y, long long N, int C, int HW) {
  genrl::pdl_sync();
  const long long total = N * HW;
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
    const long long n = i / HW;
    const long long hw = i - n * HW;
    float v[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) v[c] = c < C ? (float)x[(n * C + c) * HW + hw] / 255.0f - 0.5f : 0.f;
    uint4 pk;
    __nv_bfloat162 a0 = __floats2bfloat162_rn(v[0], v[1]), a1 = __floats2bfloat162_rn(v[2], v[3]);
    __nv_bfloat162 a2 = __floats2bfloat162_rn(v[4], v[5]), a3 = __floats2bfloat162_rn(v[6], v[7]);
    pk.x = *reinterpret_cast<uint32_t*>(&a0); pk.y = *reinterpret_cast<uint32_t*>(&a1);
    pk.z = *reinterpret_cast<uint32_t*>(&a2); pk.w = *reinterpret_cast<uint32_t*>(&a3);
    y[i] = pk;
  }
}

template <int BN>
static int launch_fwd(const CUtensorMap& tw, const ConvFwdParams& p, cudaStream_t st) {
  constexpr int STAGES = FwdCfg<BN>::STAGES;
  constexpr int SMEM = STAGES * (CV_BM + BN) * CV_BK * 2 + 1024 + 256 + CVF_KTAB_KB * 8 * 8;
  static bool attr_set = false;
  static int ctas_per_sm = 1, sms = 148;
  auto kern = conv_fwd_kernel<BN, STAGES>;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
    GENRL_REQUIRE(e == cudaSuccess, "cudaFuncSetAttribute(conv_fwd) failed: %s", cudaGetErrorString(e));
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) sms = 148;
    int occ = 0;
    {
      const int need = FwdCfg<BN>::CTAS * (SMEM + 1024);
      cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, std::min(100, (need * 100 + 228 * 1024 - 1) / (228 * 1024)));
    }
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, CVF_THREADS, SMEM) != cudaSuccess) occ = -1;
    ctas_per_sm = FwdCfg<BN>::CTAS;          // 2 CTAs x 2 accumulators of <= 128 TMEM columns fit the 512-column TMEM; smem 2 x 87 KB
    if (getenv("GENRL_CONV_DEBUG")) fprintf(stderr, "conv_fwd<%d>: occupancy query %d, using %d CTAs/SM x %d SMs, smem %d\n", BN, occ, ctas_per_sm, sms, SMEM);
    attr_set = true;
  }
  const long long tiles = ((p.g.rows + CV_BM - 1) / CV_BM) * ceil_div(p.N, BN);
  dim3 grid((unsigned)std::min<long long>(tiles, (long long)sms * ctas_per_sm), 1, 1);
  genrl::launch_pdl(kern, grid, CVF_THREADS, SMEM, st, tw, p);
  return check_launch("conv_fwd_kernel");
}

}  // namespace genrl

using namespace genrl;

static int fill_geom(ConvGeom* g, const void* in, int Nimg, int IH, int IW, int C, int RH, int RW, int sy, int ts, int th, int tw) {
  GENRL_REQUIRE(C % 8 == 0 && C > 0, "conv_igemm: channels must be a multiple of 8 (got %d)", C);
  GENRL_REQUIRE((reinterpret_cast<uintptr_t>(in) & 15) == 0, "conv_igemm: input must be 16-byte aligned");
  GENRL_REQUIRE(IH < 4096 && IW < 4096 && RH * sy + th < 8192 && RW * sy + tw < 8192 && th <= 8 && tw <= 8, "conv_igemm: image / filter too large");
  GENRL_REQUIRE((long long)Nimg * IH * IW * C < (1LL << 31), "conv_igemm: input has more than 2^31 elements");
  g->in = reinterpret_cast<const __nv_bfloat16*>(in);
  g->Nimg = Nimg; g->IH = IH; g->IW = IW; g->C = C; g->RH = RH; g->RW = RW; g->sy = sy; g->ts = ts; g->th = th; g->tw = tw;
  g->K = th * tw * C;
  g->rows = (long long)Nimg * RH * RW;
  g->m_rhw = div_magic(g->rows, RH * RW);
  g->m_rw = div_magic(RH * RW, RW);
  g->m_c = div_magic(g->K + 64 + 256, C);
  g->m_tw = div_magic(th * tw + 64, tw);
  g->all_inside = (ts > 0 && (long long)(RH - 1) * sy + th <= IH && (long long)(RW - 1) * sy + tw <= IW) ? 1 : 0;
  return 0;
}

// Forward-type implicit GEMM (see the file header).  in: bf16 NHWC [Nimg,IH,IW,C]; Wm: bf16 [ncls*Cpad, th*tw*C] (ld = ldw);
// output pixel of (row, class): ncls == 1 -> (yy, xx); ncls == 4 -> (2yy + py, 2xx + px).
extern "C" int genrl_conv_igemm_fwd(void* stream, const void* in, int Nimg, int IH, int IW, int C, int RH, int RW, int sy, int ts, int th, int tw,
                                    const void* Wm, long long ldw, int Cpad, int Cout, int ncls, int OH, int OW, void* out, int out_mode,
                                    const float* bias) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (Nimg <= 0) return 0;
  ConvFwdParams p;
  if (fill_geom(&p.g, in, Nimg, IH, IW, C, RH, RW, sy, ts, th, tw)) return 1;
  GENRL_REQUIRE(ncls == 1 || ncls == 4, "conv_igemm_fwd: ncls must be 1 or 4");
  GENRL_REQUIRE(Cpad % 8 == 0 && Cout <= Cpad && (ncls * Cpad) % 16 == 0, "conv_igemm_fwd: Cpad must be a multiple of 8 covering Cout, N a multiple of 16");
  GENRL_REQUIRE(out_mode == 1 || Cout % 8 == 0, "conv_igemm_fwd: NHWC output needs Cout %% 8 == 0");
  GENRL_REQUIRE((ldw % 8) == 0 && (reinterpret_cast<uintptr_t>(Wm) & 15) == 0, "conv_igemm_fwd: filter matrix must be 16-byte aligned rows");
  p.N = ncls * Cpad; p.Cpad = Cpad; p.Cout = Cout; p.ncls = ncls; p.OH = OH; p.OW = OW; p.out_mode = out_mode; p.out = out; p.bias = bias;
  p.num_kb = ceil_div(p.g.K, CV_BK);
  int bn;
  if (p.N <= 16) bn = 16;
  else if (p.N <= 32) bn = 32;
  else if (p.N <= 48) bn = 48;
  else if (p.N <= 64) bn = 64;
  else if (p.N <= 96) bn = 96;
  else if (p.N % 192 == 0 || p.N > 128) bn = 192;
  else bn = 128;
  CUtensorMap tm;
  if (make_map(&tm, Wm, ncls == 1 ? Cout : p.N, p.g.K, ldw, 64, bn)) return 1;      // rows past the matrix are zero-filled by TMA
  switch (bn) {
    case 16: return launch_fwd<16>(tm, p, st);
    case 32: return launch_fwd<32>(tm, p, st);
    case 48: return launch_fwd<48>(tm, p, st);
    case 64: return launch_fwd<64>(tm, p, st);
    case 96: return launch_fwd<96>(tm, p, st);
    case 128: return launch_fwd<128>(tm, p, st);
    default: return launch_fwd<192>(tm, p, st);
  }
}

// Weight gradients in gather form: dw[Cs, th*tw*C] (fp32, ld = lddw) += S^T col, S: bf16 [rows, Cs] (ld = lds) aligned with the row grid.
extern "C" int genrl_conv_igemm_wgrad(void* stream, const void* in, int Nimg, int IH, int IW, int C, int RH, int RW, int sy, int th, int tw,
                                      const void* S, long long lds, int Cs, float* dw, long long lddw) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (Nimg <= 0) return 0;
  ConvWgParams p;
  if (fill_geom(&p.g, in, Nimg, IH, IW, C, RH, RW, sy, 1, th, tw)) return 1;
  GENRL_REQUIRE((RH - 1) * sy + th <= IH && (RW - 1) * sy + tw <= IW, "conv_igemm_wgrad: taps must lie inside the image (gather geometry)");
  GENRL_REQUIRE((lds % 8) == 0 && (reinterpret_cast<uintptr_t>(S) & 15) == 0, "conv_igemm_wgrad: S must have 16-byte aligned rows");
  p.Cs = Cs; p.dw = dw; p.ldw = (int)lddw;
  p.num_kb = ceil_div(p.g.rows, CV_BK);
  constexpr int BN = 128, STAGES = 3;
  const int ntiles = ceil_div(p.g.K, BN) * ceil_div(Cs, CV_BM);
  int sms = 148;
  { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); if (sms <= 0) sms = 148; }
  int splits = std::max(1, (2 * sms) / ntiles);
  splits = std::min(splits, std::max(1, p.num_kb / 8));
  p.kb_per_split = ceil_div(p.num_kb, splits);
  splits = ceil_div(p.num_kb, p.kb_per_split);
  CUtensorMap tm;
  if (make_map(&tm, S, p.g.rows, Cs, lds, 64, 64)) return 1;
  constexpr int SMEM = STAGES * (CV_BM + BN) * CV_BK * 2 + 1024 + 256;
  static bool attr_set = false;
  auto kern = conv_wgrad_kernel<BN, STAGES>;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
    GENRL_REQUIRE(e == cudaSuccess, "cudaFuncSetAttribute(conv_wgrad) failed: %s", cudaGetErrorString(e));
    attr_set = true;
  }
  dim3 grid(ceil_div(p.g.K, BN), ceil_div(Cs, CV_BM), splits);
  genrl::launch_pdl(kern, grid, CV_THREADS, SMEM, st, tm, p);
  return check_launch("conv_wgrad_kernel");
}

extern "C" int genrl_conv_pack_parity(void* stream, const float* w, void* dst, int A, int B, int k, int Cpad) {
  const int T = (k + 1) / 2;
  const long long total = 4LL * Cpad * T * T * A;
  genrl::launch_pdl(conv_pack_parity_kernel, (int)std::min<long long>((total + 255) / 256, 148 * 8), 256, 0, reinterpret_cast<cudaStream_t>(stream), 
      w, reinterpret_cast<__nv_bfloat16*>(dst), A, B, k, T, Cpad);
  return check_launch("conv_pack_parity");
}

extern "C" int genrl_conv_pack_gather(void* stream, const float* w, void* dst, int A, int B, int k, int Cp) {
  const long long total = (long long)A * k * k * Cp;
  genrl::launch_pdl(conv_pack_gather_kernel, (int)std::min<long long>((total + 255) / 256, 148 * 8), 256, 0, reinterpret_cast<cudaStream_t>(stream), 
      w, reinterpret_cast<__nv_bfloat16*>(dst), A, B, k, Cp);
  return check_launch("conv_pack_gather");
}

extern "C" int genrl_conv_unpack_wgrad(void* stream, const float* dwp, long long ld, float* grad, int A, int B, int k, int Cp) {
  const long long total = (long long)A * B * k * k;
  genrl::launch_pdl(conv_unpack_wgrad_kernel, (int)std::min<long long>((total + 255) / 256, 148 * 8), 256, 0, reinterpret_cast<cudaStream_t>(stream), dwp, ld, grad, A, B,
                                                                                                                                k * k, Cp);
  return check_launch("conv_unpack_wgrad");
}

extern "C" int genrl_nchw_to_nhwc_pad(void* stream, const float* x, void* y, long long N, int C, int H, int W, int Cp) {
  const long long total = N * H * W;
  genrl::launch_pdl(nchw_to_nhwc_pad_kernel, (int)std::min<long long>((total + 255) / 256, 148 * 16), 256, 0, reinterpret_cast<cudaStream_t>(stream), 
      x, reinterpret_cast<__nv_bfloat16*>(y), N, C, H, W, Cp);
  return check_launch("nchw_to_nhwc_pad");
}

extern "C" int genrl_u8_nchw_to_nhwc8(void* stream, const void* x, void* y, long long N, int C, int H, int W) {
  GENRL_REQUIRE(C <= 8 && (reinterpret_cast<uintptr_t>(y) & 15) == 0, "u8_nchw_to_nhwc8: at most 8 channels, 16-byte aligned output");
  const long long total = N * H * W;
  genrl::launch_pdl(u8_nchw_to_nhwc8_kernel, (int)std::min<long long>((total + 255) / 256, 148 * 16), 256, 0, reinterpret_cast<cudaStream_t>(stream), 
      reinterpret_cast<const uint8_t*>(x), reinterpret_cast<uint4*>(y), N, C, H * W);
  return check_launch("u8_nchw_to_nhwc8");
}
