// Stable LSD radix sort of 64-bit (brick key << 32 | record index) pairs by brick key.
//
// Pairs are emitted in particle order and the sort is stable, so inside every brick the
// particles stay in index order -- the summation order of the reference's per-pixel
// np.sum (martini.py:281).  Hand-written.  The key's bits are split evenly over the fewest
// passes of at most SORT_MAX_BITS each (18-bit pixel keys: 2 x 9, 20-bit (tile, channel) keys:
// 2 x 10).  Per pass: (1) one digit histogram per block tile of SORT_TILE pairs, (2) one
// exclusive scan over the digit-major histogram matrix, (3) a scatter in which the block first
// orders its tile by digit in shared memory -- every warp walks its contiguous share in order
// and ranks equal digits with one ballot per digit bit, the warps' counts are prefix-summed per digit
// -- and then writes each digit's run to its place in one piece: runs of SORT_TILE / 2^bits
// pairs (64 to 512 bytes) instead of the single 8-byte stores of rounds 1-2, which cost four
// times their bytes in 32-byte sectors (131 us against 32 us per pass of 9.3e6 pairs, the
// same instructions, depending only on how scattered the digit was).  HBM-bound integer work:
// 24 B moved per pair per pass.
#pragma once

#include "common.cuh"
#include "scan.cuh"

namespace mtn {

constexpr int SORT_MAX_BITS = 10;
constexpr int SORT_WARPS = 8;
constexpr int SORT_THREADS = SORT_WARPS * 32;
constexpr int SORT_PER_THREAD = 32;                       // pairs a thread holds in registers
constexpr int SORT_IPW = 32 * SORT_PER_THREAD;            // pairs per warp (contiguous)
constexpr int SORT_TILE = SORT_WARPS * SORT_IPW;          // pairs per block: 8192

inline int64_t sort_num_tiles(int64_t n) { return (n + SORT_TILE - 1) / SORT_TILE; }
// passes and bits per pass for a key of `key_bits` bits
inline int sort_passes(int key_bits) { return (key_bits + SORT_MAX_BITS - 1) / SORT_MAX_BITS; }
inline int sort_bits_per_pass(int key_bits) {
  const int p = sort_passes(key_bits);
  return p ? (key_bits + p - 1) / p : 0;
}
inline int64_t sort_hist_entries(int64_t n) { return sort_num_tiles(n) << SORT_MAX_BITS; }
inline size_t sort_hist_bytes(int64_t n) {
  return align_up((size_t)sort_hist_entries(n) * sizeof(uint32_t) + 16);
}
// dynamic shared memory of the scatter kernel: the staged tile, per-warp digit counters,
// per-digit (global offset - local start)
inline size_t sort_scatter_smem(int bits) {
  return (size_t)SORT_TILE * 8 + ((size_t)SORT_WARPS + 1) * ((size_t)4 << bits);
}

__global__ void __launch_bounds__(SORT_THREADS) radix_count_kernel(
    const uint64_t* __restrict__ in, int64_t n, int shift, int bits, int64_t n_tiles,
    uint32_t* __restrict__ hist) {
  __shared__ uint32_t cnt[1 << SORT_MAX_BITS];
  const int bins = 1 << bits;
  const uint32_t mask = (uint32_t)bins - 1u;
  for (int d = threadIdx.x; d < bins; d += SORT_THREADS) cnt[d] = 0;
  __syncthreads();
  const int64_t base = (int64_t)blockIdx.x * SORT_TILE;
#pragma unroll 8
  for (int s = 0; s < SORT_TILE / SORT_THREADS; ++s) {
    const int64_t i = base + s * SORT_THREADS + threadIdx.x;
    if (i < n) atomicAdd(&cnt[(uint32_t)(in[i] >> shift) & mask], 1u);
  }
  __syncthreads();
  for (int d = threadIdx.x; d < bins; d += SORT_THREADS) hist[(int64_t)d * n_tiles + blockIdx.x] = cnt[d];
}

__global__ void __launch_bounds__(SORT_THREADS, 2) radix_scatter_kernel(
    const uint64_t* __restrict__ in, uint64_t* __restrict__ out, int64_t n, int shift, int bits,
    int64_t n_tiles, const uint32_t* __restrict__ offs) {
  MTN_DYN_SMEM(unsigned char, smem_raw);
  const int bins = 1 << bits;
  const uint32_t mask = (uint32_t)bins - 1u;
  uint64_t* stage = reinterpret_cast<uint64_t*>(smem_raw);                  // [SORT_TILE]
  uint32_t* cnt = reinterpret_cast<uint32_t*>(stage + SORT_TILE);          // [SORT_WARPS][bins]
  uint32_t* delta = cnt + SORT_WARPS * bins;                               // [bins]
  __shared__ uint32_t scan_sm[33];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t base = (int64_t)blockIdx.x * SORT_TILE;
  const int n_valid = (int)min((int64_t)SORT_TILE, n - base);

  for (int k = tid; k < SORT_WARPS * bins; k += SORT_THREADS) cnt[k] = 0;
  // this thread's pairs: element s of lane l of warp w is pair  w * SORT_IPW + s * 32 + l  of the
  // tile (coalesced loads, all in flight together)
  uint64_t kv[SORT_PER_THREAD];
#pragma unroll
  for (int s = 0; s < SORT_PER_THREAD; ++s) {
    const int loc = warp * SORT_IPW + s * 32 + lane;
    kv[s] = loc < n_valid ? in[base + loc] : 0ull;
  }
  __syncthreads();

  // rank inside the warp's share: pairs of equal digit, in order
  uint32_t rank2[SORT_PER_THREAD / 2];  // two 16-bit ranks per register
#pragma unroll
  for (int s = 0; s < SORT_PER_THREAD / 2; ++s) rank2[s] = 0;
  uint32_t* wcnt = cnt + warp * bins;
  const uint32_t lt = (1u << lane) - 1u;
#pragma unroll
  for (int s = 0; s < SORT_PER_THREAD; ++s) {
    const bool valid = warp * SORT_IPW + s * 32 + lane < n_valid;
    // invalid lanes get private digits so they never match a real one
    const uint32_t d = valid ? ((uint32_t)(kv[s] >> shift) & mask) : (uint32_t)(bins + lane);
    // lanes holding my digit, one ballot per digit bit (a fixed ~2 instructions per bit:
    // __match_any_sync takes one pass per DISTINCT value in the warp, ~30 of them for a random
    // 9-bit digit, and was most of this kernel)
    uint32_t peers = __ballot_sync(0xffffffffu, valid);
    for (int k = 0; k < bits; ++k) {
      const uint32_t bk = __ballot_sync(0xffffffffu, (d >> k) & 1u);
      peers &= ((d >> k) & 1u) ? bk : ~bk;
    }
    if (!valid) peers = 1u << lane;
    const uint32_t r = __popc(peers & lt);
    uint32_t before = 0;
    if (valid) before = wcnt[d];
    __syncwarp();
    if (valid && r == 0) wcnt[d] = before + __popc(peers);
    __syncwarp();
    rank2[s >> 1] |= (before + r) << (16 * (s & 1));
  }
  __syncthreads();

  // per digit: exclusive prefix over the warps, then over the digits (the tile's local order)
  {
    const int per = (bins + SORT_THREADS - 1) / SORT_THREADS;  // consecutive digits per thread (<= 4)
    uint32_t tot[4] = {0, 0, 0, 0};
    uint32_t mine = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int d = tid * per + k;
      if (k < per && d < bins) {
        uint32_t run = 0;
        for (int w = 0; w < SORT_WARPS; ++w) {
          const uint32_t c = cnt[w * bins + d];
          cnt[w * bins + d] = run;
          run += c;
        }
        tot[k] = run;
        mine += run;
      }
    }
    uint32_t total;
    uint32_t start = block_excl_scan(mine, scan_sm, &total);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int d = tid * per + k;
      if (k < per && d < bins) {
        for (int w = 0; w < SORT_WARPS; ++w) cnt[w * bins + d] += start;
        delta[d] = offs[(int64_t)d * n_tiles + blockIdx.x] - start;  // (mod 2^32)
        start += tot[k];
      }
    }
  }
  __syncthreads();

#pragma unroll
  for (int s = 0; s < SORT_PER_THREAD; ++s) {
    if (warp * SORT_IPW + s * 32 + lane < n_valid) {
      const uint32_t d = (uint32_t)(kv[s] >> shift) & mask;
      stage[wcnt[d] + ((rank2[s >> 1] >> (16 * (s & 1))) & 0xffffu)] = kv[s];
    }
  }
  __syncthreads();
  for (int i = tid; i < n_valid; i += SORT_THREADS) {
    const uint64_t v = stage[i];
    out[(uint32_t)i + delta[(uint32_t)(v >> shift) & mask]] = v;
  }
}

// Sort `n` pairs by the low `key_bits` bits of their upper word.  `a` holds the input;
// returns (in *sorted) whichever of a / b holds the result.
inline int radix_sort_pairs(uint64_t* a, uint64_t* b, int64_t n, int key_bits, uint32_t* hist,
                            void* scan_temp, uint64_t** sorted, cudaStream_t st) {
  *sorted = a;
  if (n <= 1) return MTN_OK;
  const int64_t n_tiles = sort_num_tiles(n);
  const int bits = sort_bits_per_pass(key_bits);
  const size_t smem = sort_scatter_smem(bits);
#ifndef MTN_HOST_EMU
  static thread_local int attr_dev = -1;  // (per host thread and device, like the projection kernels')
  int dev = 0;
  MTN_CUDA(cudaGetDevice(&dev));
  if (attr_dev != dev) {
    MTN_CUDA(cudaFuncSetAttribute(radix_scatter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)sort_scatter_smem(SORT_MAX_BITS)));
    attr_dev = dev;
  }
#endif
  uint64_t* src = a;
  uint64_t* dst = b;
  for (int bit = 0; bit < key_bits; bit += bits) {
    const int shift = 32 + bit;
    MTN_LAUNCH(radix_count_kernel, (unsigned)n_tiles, SORT_THREADS, 0, st, src, n, shift, bits, n_tiles, hist);
    MTN_LAUNCH_CHECK();
    int rc = exclusive_scan<uint32_t, uint32_t>(hist, hist, n_tiles << bits, scan_temp, nullptr, st);
    if (rc) return rc;
    MTN_LAUNCH(radix_scatter_kernel, (unsigned)n_tiles, SORT_THREADS, smem, st, src, dst, n, shift, bits, n_tiles,
               hist);
    MTN_LAUNCH_CHECK();
    uint64_t* t = src;
    src = dst;
    dst = t;
  }
  *sorted = src;
  return MTN_OK;
}

}  // namespace mtn
