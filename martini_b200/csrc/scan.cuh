// Device-wide exclusive prefix sum (hand-written; no CUB/thrust).
//
// Three launches: per-block sums -> one CTA scans the block sums -> per-block rescan
// with the block's base (short in-place scans take a single-CTA shortcut).  Blocks cover SCAN_ITEMS elements; the middle step loops, so any
// length works.  HBM-bound: reads the input twice, writes it once.
#pragma once

#include "common.cuh"

namespace mtn {

constexpr int SCAN_THREADS = 512;
constexpr int SCAN_PER_THREAD = 8;
constexpr int SCAN_ITEMS = SCAN_THREADS * SCAN_PER_THREAD;
constexpr int64_t SCAN_SMALL = 16384;  // up to here a single CTA scans in place

template <typename T>
__device__ __forceinline__ T warp_incl_scan(T x) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    T y = __shfl_up_sync(0xffffffffu, x, d);
    if (lane >= d) x += y;
  }
  return x;
}

// Exclusive scan of one value per thread across the block; returns the exclusive prefix
// and leaves the block total in *total.  `smem` must hold 33 T's.
template <typename T>
__device__ __forceinline__ T block_excl_scan(T x, T* smem, T* total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nwarps = (blockDim.x + 31) >> 5;
  const T incl = warp_incl_scan(x);
  if (lane == 31) smem[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    T w = lane < nwarps ? smem[lane] : T(0);
    const T wi = warp_incl_scan(w);
    smem[lane] = wi - w;
    if (lane == 31) smem[32] = wi;
  }
  __syncthreads();
  const T res = smem[warp] + incl - x;
  *total = smem[32];
  __syncthreads();
  return res;
}

// Exclusive scans of three values per thread across the block in one pass (x[k] is replaced by
// its exclusive prefix).  `smem` holds 3 x 33 values; smem[k][32] is left holding the block
// total of quantity k.  Two barriers.
template <typename T>
__device__ __forceinline__ void block_excl_scan3(T x[3], T (*smem)[33]) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nwarps = (blockDim.x + 31) >> 5;
  T incl[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    incl[k] = warp_incl_scan(x[k]);
    if (lane == 31) smem[k][warp] = incl[k];
  }
  __syncthreads();
  if (warp < 3) {  // warp k scans the warp totals of quantity k
    const T w = lane < nwarps ? smem[warp][lane] : T(0);
    const T wi = warp_incl_scan(w);
    smem[warp][lane] = wi - w;
    if (lane == 31) smem[warp][32] = wi;
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 3; ++k) x[k] = smem[k][warp] + incl[k] - x[k];
}

template <typename TIn, typename T>
__global__ void __launch_bounds__(SCAN_THREADS) scan_block_sums(const TIn* __restrict__ in,
                                                                int64_t n, T* __restrict__ sums) {
  __shared__ T sm[33];
  const int64_t base = (int64_t)blockIdx.x * SCAN_ITEMS + (int64_t)threadIdx.x * SCAN_PER_THREAD;
  T s = 0;
#pragma unroll
  for (int k = 0; k < SCAN_PER_THREAD; ++k)
    if (base + k < n) s += (T)in[base + k];
  T total;
  block_excl_scan(s, sm, &total);
  if (threadIdx.x == 0) sums[blockIdx.x] = total;
}

// One CTA: exclusive scan of `m` block sums in place; grand total to *total_out.
template <typename T>
__global__ void __launch_bounds__(1024) scan_sums_inplace(T* __restrict__ sums, int64_t m,
                                                          T* __restrict__ total_out) {
  __shared__ T sm[33];
  T carry = 0;
  for (int64_t base = 0; base < m; base += blockDim.x) {
    const int64_t i = base + threadIdx.x;
    const T x = i < m ? sums[i] : T(0);
    T total;
    const T ex = block_excl_scan(x, sm, &total);
    if (i < m) sums[i] = carry + ex;
    carry += total;
  }
  if (threadIdx.x == 0 && total_out) *total_out = carry;
}

// Three arrays of block sums scanned by the three CTAs of ONE launch (mtn_plan's kept /
// pairs / second-stream pairs); totals to total_out[blockIdx.x].
__global__ void __launch_bounds__(1024) scan3_sums_inplace(int64_t* __restrict__ a, int64_t* __restrict__ b,
                                                           int64_t* __restrict__ c, int64_t m,
                                                           int64_t* __restrict__ total_out) {
  __shared__ int64_t sm[33];
  constexpr int PER = 4;  // consecutive entries per thread and round, staged through shared memory
  __shared__ int64_t stage[1024 * PER];  // so that the global loads and stores of a round are coalesced
  int64_t* sums = blockIdx.x == 0 ? a : (blockIdx.x == 1 ? b : c);
  int64_t carry = 0;
  for (int64_t base = 0; base < m; base += (int64_t)blockDim.x * PER) {
#pragma unroll
    for (int k = 0; k < PER; ++k) {
      const int64_t i = base + (int64_t)k * blockDim.x + threadIdx.x;
      stage[k * blockDim.x + threadIdx.x] = i < m ? sums[i] : 0;
    }
    __syncthreads();
    int64_t v[PER], s = 0;
#pragma unroll
    for (int k = 0; k < PER; ++k) {
      v[k] = stage[threadIdx.x * PER + k];
      s += v[k];
    }
    int64_t total;
    int64_t run = carry + block_excl_scan(s, sm, &total);
#pragma unroll
    for (int k = 0; k < PER; ++k) {
      stage[threadIdx.x * PER + k] = run;
      run += v[k];
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < PER; ++k) {
      const int64_t i = base + (int64_t)k * blockDim.x + threadIdx.x;
      if (i < m) sums[i] = stage[k * blockDim.x + threadIdx.x];
    }
    carry += total;
    __syncthreads();
  }
  if (threadIdx.x == 0) total_out[blockIdx.x] = carry;
}

template <typename TIn, typename T>
__global__ void __launch_bounds__(SCAN_THREADS) scan_apply(const TIn* in, int64_t n,
                                                           const T* __restrict__ sums, T* out) {
  __shared__ T sm[33];
  const int64_t base = (int64_t)blockIdx.x * SCAN_ITEMS + (int64_t)threadIdx.x * SCAN_PER_THREAD;
  T v[SCAN_PER_THREAD];
  T s = 0;
#pragma unroll
  for (int k = 0; k < SCAN_PER_THREAD; ++k) {
    v[k] = base + k < n ? (T)in[base + k] : T(0);
    s += v[k];
  }
  T total;
  T run = block_excl_scan(s, sm, &total) + sums[blockIdx.x];
#pragma unroll
  for (int k = 0; k < SCAN_PER_THREAD; ++k) {
    if (base + k < n) out[base + k] = run;
    run += v[k];
  }
}

inline size_t scan_temp_bytes(int64_t n, size_t elem) {
  return align_up(((size_t)((n + SCAN_ITEMS - 1) / SCAN_ITEMS) + 1) * elem);
}

// out[i] = sum_{j<i} in[j]; *total_dev (device, optional) = sum of all.  in == out allowed.
template <typename TIn, typename T>
int exclusive_scan(const TIn* in, T* out, int64_t n, void* temp, T* total_dev, cudaStream_t st) {
  if (n <= 0) {
    if (total_dev) MTN_CUDA(cudaMemsetAsync(total_dev, 0, sizeof(T), st));
    return MTN_OK;
  }
  // short in-place scans (brick tables of ordinary cubes): one CTA, one launch
  if (sizeof(TIn) == sizeof(T) && (const void*)in == (const void*)out && n <= SCAN_SMALL) {
    MTN_LAUNCH(scan_sums_inplace<T>, 1, 1024, 0, st, out, n, total_dev);
    MTN_LAUNCH_CHECK();
    return MTN_OK;
  }
  const int64_t nb = (n + SCAN_ITEMS - 1) / SCAN_ITEMS;
  T* sums = reinterpret_cast<T*>(temp);
  auto k_sums = scan_block_sums<TIn, T>;
  MTN_LAUNCH(k_sums, (unsigned)nb, SCAN_THREADS, 0, st, in, n, sums);
  MTN_LAUNCH_CHECK();
  MTN_LAUNCH(scan_sums_inplace<T>, 1, 1024, 0, st, sums, nb, total_dev);
  MTN_LAUNCH_CHECK();
  auto k_apply = scan_apply<TIn, T>;
  MTN_LAUNCH(k_apply, (unsigned)nb, SCAN_THREADS, 0, st, in, n, sums, out);
  MTN_LAUNCH_CHECK();
  return MTN_OK;
}

}  // namespace mtn
