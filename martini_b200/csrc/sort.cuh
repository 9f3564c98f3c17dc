// Stable LSD radix sort of 64-bit (brick key << 32 | record index) pairs by brick key.
//
// Pairs are emitted in particle order and the sort is stable, so inside every brick the
// particles stay in index order -- the summation order of the reference's per-pixel
// np.sum (martini.py:281).  Hand-written: 8-bit digits; per pass (1) per-warp-chunk digit
// histograms, (2) one exclusive scan over the digit-major histogram matrix, (3) a stable
// scatter in which each warp walks its chunk in order and ranks equal digits with
// __match_any_sync.  HBM-bound integer work: 24 B moved per pair per pass.
#pragma once

#include "common.cuh"
#include "scan.cuh"

namespace mtn {

constexpr int RADIX_BITS = 8;
constexpr int RADIX = 1 << RADIX_BITS;
constexpr int SORT_WARPS = 8;
constexpr int SORT_THREADS = SORT_WARPS * 32;
constexpr int SORT_IPW = 1024;  // pairs per warp chunk

inline int64_t sort_num_chunks(int64_t n) { return (n + SORT_IPW - 1) / SORT_IPW; }
inline size_t sort_hist_bytes(int64_t n) {
  return align_up((size_t)sort_num_chunks(n) * RADIX * sizeof(uint32_t) + 16);
}

__global__ void __launch_bounds__(SORT_THREADS) radix_count_kernel(
    const uint64_t* __restrict__ in, int64_t n, int shift, int64_t n_chunks,
    uint32_t* __restrict__ hist) {
  __shared__ uint32_t cnt[SORT_WARPS][RADIX];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t chunk = (int64_t)blockIdx.x * SORT_WARPS + warp;
  if (chunk >= n_chunks) return;  // whole warp leaves together; no block barrier below
  for (int d = lane; d < RADIX; d += 32) cnt[warp][d] = 0;
  __syncwarp();
  const int64_t base = chunk * SORT_IPW;
#pragma unroll 4
  for (int s = 0; s < SORT_IPW / 32; ++s) {
    const int64_t i = base + s * 32 + lane;
    if (i < n) atomicAdd(&cnt[warp][(uint32_t)(in[i] >> shift) & (RADIX - 1)], 1u);
  }
  __syncwarp();
  for (int d = lane; d < RADIX; d += 32) hist[(int64_t)d * n_chunks + chunk] = cnt[warp][d];
}

__global__ void __launch_bounds__(SORT_THREADS) radix_scatter_kernel(
    const uint64_t* __restrict__ in, uint64_t* __restrict__ out, int64_t n, int shift,
    int64_t n_chunks, const uint32_t* __restrict__ offs) {
  __shared__ uint32_t pos[SORT_WARPS][RADIX];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t chunk = (int64_t)blockIdx.x * SORT_WARPS + warp;
  if (chunk >= n_chunks) return;
  for (int d = lane; d < RADIX; d += 32) pos[warp][d] = offs[(int64_t)d * n_chunks + chunk];
  __syncwarp();
  const uint32_t lt = (1u << lane) - 1u;
  const int64_t base = chunk * SORT_IPW;
  for (int s = 0; s < SORT_IPW / 32; ++s) {
    const int64_t i = base + s * 32 + lane;
    const bool valid = i < n;
    const uint64_t kv = valid ? in[i] : 0ull;
    // invalid lanes get private digits so they never match a real one
    const uint32_t d = valid ? ((uint32_t)(kv >> shift) & (RADIX - 1)) : (uint32_t)(RADIX + lane);
    const uint32_t peers = __match_any_sync(0xffffffffu, d);
    const uint32_t rank = __popc(peers & lt);
    uint32_t p = 0;
    if (valid) p = pos[warp][d] + rank;
    __syncwarp();
    if (valid) {
      out[p] = kv;
      if (rank == 0) pos[warp][d] += __popc(peers);
    }
    __syncwarp();
  }
}

// Sort `n` pairs by the low `key_bits` bits of their upper word.  `a` holds the input;
// returns (in *sorted) whichever of a / b holds the result.
inline int radix_sort_pairs(uint64_t* a, uint64_t* b, int64_t n, int key_bits, uint32_t* hist,
                            void* scan_temp, uint64_t** sorted, cudaStream_t st) {
  *sorted = a;
  if (n <= 1) return MTN_OK;
  const int64_t n_chunks = sort_num_chunks(n);
  const unsigned grid = (unsigned)((n_chunks + SORT_WARPS - 1) / SORT_WARPS);
  uint64_t* src = a;
  uint64_t* dst = b;
  for (int bit = 0; bit < key_bits; bit += RADIX_BITS) {
    const int shift = 32 + bit;
    MTN_LAUNCH(radix_count_kernel, grid, SORT_THREADS, 0, st, src, n, shift, n_chunks, hist);
    MTN_LAUNCH_CHECK();
    int rc = exclusive_scan<uint32_t, uint32_t>(hist, hist, n_chunks * RADIX, scan_temp, nullptr, st);
    if (rc) return rc;
    MTN_LAUNCH(radix_scatter_kernel, grid, SORT_THREADS, 0, st, src, dst, n, shift, n_chunks, hist);
    MTN_LAUNCH_CHECK();
    uint64_t* t = src;
    src = dst;
    dst = t;
  }
  *sorted = src;
  return MTN_OK;
}

}  // namespace mtn
