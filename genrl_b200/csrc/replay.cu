// On-GPU episode store: the window gather behind `next(replay_iter)`.
//   reference tools/replay.py:223-236 — per batch, for every key, B Python-level slices episode[b][t : t+L] are stacked on the host and the
//   ~30 MB result is copied to the device from pageable memory.  Here episodes live in HBM (one flat [steps, row_bytes] store per key) and a
//   batch is B contiguous spans of L rows copied store -> batch by one kernel: 2 * B * L * row_bytes of HBM traffic, nothing on the host
//   but sampling the B (episode, offset) pairs.
#include "common.cuh"
#include <algorithm>

namespace genrl {

#define ST(s) reinterpret_cast<cudaStream_t>(s)

// dst[b, :] (span_bytes) = src[start[b] * row_bytes : ... + span_bytes];  16-byte vector path when everything is aligned
__global__ void __launch_bounds__(256) replay_gather_kernel(const unsigned char* __restrict__ src, const long long* __restrict__ start, long long row_bytes,
                                                            long long span_bytes, unsigned char* __restrict__ dst, int chunks_per_span, int vec) {
  genrl::pdl_sync();
  const int b = blockIdx.x / chunks_per_span;
  const int c = blockIdx.x % chunks_per_span;
  const unsigned char* s = src + start[b] * row_bytes;
  unsigned char* d = dst + (long long)b * span_bytes;
  const long long per = (span_bytes + chunks_per_span - 1) / chunks_per_span;
  long long lo = (long long)c * per, hi = min(lo + per, span_bytes);
  if (vec) {
    lo = lo / 16 * 16;                                  // chunk boundaries on 16 bytes (per is a multiple of 16 in this mode)
    const uint4* s4 = reinterpret_cast<const uint4*>(s + lo);
    uint4* d4 = reinterpret_cast<uint4*>(d + lo);
    const long long n4 = (hi - lo) / 16;
    for (long long i = threadIdx.x; i < n4; i += 256) d4[i] = s4[i];
    for (long long i = lo + n4 * 16 + threadIdx.x; i < hi; i += 256) d[i] = s[i];
  } else {
    for (long long i = lo + threadIdx.x; i < hi; i += 256) d[i] = s[i];
  }
}

}  // namespace genrl

using namespace genrl;

extern "C" int genrl_replay_gather(void* stream, const void* store, const long long* start, long long row_bytes, int B, int L, void* batch) {
  if (B <= 0 || L <= 0 || row_bytes <= 0) return 0;
  const long long span = (long long)L * row_bytes;
  int chunks = (int)std::min<long long>(std::max<long long>(span / (64 * 1024), 1), 64);      // ~64 KB per CTA, B * chunks CTAs
  long long per = (span + chunks - 1) / chunks;
  const bool aligned = (row_bytes % 16 == 0) && (reinterpret_cast<uintptr_t>(store) % 16 == 0) && (reinterpret_cast<uintptr_t>(batch) % 16 == 0);
  if (aligned && per % 16 != 0) { per = (per + 15) / 16 * 16; chunks = (int)((span + per - 1) / per); }
  genrl::launch_pdl(replay_gather_kernel, (unsigned)(B * chunks), 256, 0, ST(stream), (const unsigned char*)store, start, row_bytes, span, (unsigned char*)batch, chunks,
                                                                      aligned ? 1 : 0);
  return check_launch("replay_gather");
}
