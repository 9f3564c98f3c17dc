// Particle routing for the multi-GPU path: the exchange step of the input side, fused with its
// bucketing.  Every rank starts with a contiguous share of the particle list; a particle is
// needed by every rank whose x-slab its candidate box [px - r, px + r] can reach (halo
// particles go to both neighbours).  Instead of packing send buffers and calling an all-to-all,
// the scatter kernel stores each particle's quantities STRAIGHT INTO THE DESTINATION RANK'S
// inbox through peer-mapped pointers (NVLink): the transfer is the kernel's own store stream.
//
//   route_count_kernel    per block and destination: how many of the block's particles go there;
//   (exclusive scans of the block counts per destination; an all-gather of the W totals gives
//    every source its offset in every inbox -- device-side, no host involved)
//   route_scatter_kernel  position = offset of this source in the destination's inbox + block
//                         prefix + rank inside the block (ballots, in particle order), then F
//                         stores of 8 bytes per (particle, destination).
//
// Inboxes are filled in (source rank, particle index) order = ascending global index: the
// summation order of the single-GPU run.  The destination test is conservative (one pixel of
// slack); the exact candidate-box predicate is applied by mtn_plan on the receiving rank.
#pragma once

#include "common.cuh"

namespace mtn {

constexpr int ROUTE_THREADS = 256;
constexpr int ROUTE_MAX_WORLD = 16;
constexpr int ROUTE_MAX_FIELDS = 12;

struct RouteArgs {
  int64_t n;
  const double* px;
  const double* sm_range;
  int world;
  int bounds[ROUTE_MAX_WORLD + 1];  // slab boundaries (cube rows)
};

// Destination ranks of particle i: [d0, d1] (empty if d0 > d1).
__device__ __forceinline__ void route_dests(const RouteArgs& a, int64_t i, int& d0, int& d1) {
  d0 = 1;
  d1 = 0;
  const double x = a.px[i];
  if (isnan(x)) return;
  double r = a.sm_range[i];
  if (isnan(r)) r = 0.0;
  // first / last cube row the box may reach, one pixel of slack; clamp before the int cast
  const double top = (double)a.bounds[a.world] + 2.0;
  const double lo = fmin(fmax(floor(x - r) - 1.0, -2.0), top), hi = fmin(fmax(ceil(x + r) + 1.0, -2.0), top);
  int f = a.world, l = -1;
  for (int d = 0; d < a.world; ++d) {  // slab d = rows [bounds[d], bounds[d+1]); skip empty slabs
    const bool hit = a.bounds[d + 1] > a.bounds[d] && lo < (double)a.bounds[d + 1] && hi >= (double)a.bounds[d];
    if (hit) {
      f = min(f, d);
      l = max(l, d);
    }
  }
  d0 = f;
  d1 = l;
}

__global__ void __launch_bounds__(ROUTE_THREADS) route_count_kernel(RouteArgs a, int64_t nblk,
                                                                    uint32_t* __restrict__ blk_cnt) {
  __shared__ uint32_t cnt[ROUTE_MAX_WORLD];
  if (threadIdx.x < ROUTE_MAX_WORLD) cnt[threadIdx.x] = 0;
  __syncthreads();
  const int64_t i = (int64_t)blockIdx.x * ROUTE_THREADS + threadIdx.x;
  int d0 = 1, d1 = 0;
  if (i < a.n) route_dests(a, i, d0, d1);
  for (int d = 0; d < a.world; ++d) {
    const uint32_t m = __ballot_sync(0xffffffffu, d >= d0 && d <= d1);
    if ((threadIdx.x & 31) == 0 && m) atomicAdd(&cnt[d], (uint32_t)__popc(m));
  }
  __syncthreads();
  if (threadIdx.x < a.world) blk_cnt[(int64_t)threadIdx.x * nblk + blockIdx.x] = cnt[threadIdx.x];
}

struct ScatterArgs {
  int n_fields;
  const double* src[ROUTE_MAX_FIELDS];           // this rank's arrays
  double* dst[ROUTE_MAX_WORLD];                  // inbox of every rank (peer-mapped), [field][capacity]
  int64_t capacity;
  const uint32_t* blk_off;                       // [world][nblk] exclusive prefix of blk_cnt
  const int64_t* src_off;                        // [world]: where this source starts in each inbox
};

__global__ void __launch_bounds__(ROUTE_THREADS) route_scatter_kernel(RouteArgs a, ScatterArgs s, int64_t nblk) {
  __shared__ uint32_t warp_cnt[ROUTE_MAX_WORLD][ROUTE_THREADS / 32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t i = (int64_t)blockIdx.x * ROUTE_THREADS + threadIdx.x;
  int d0 = 1, d1 = 0;
  if (i < a.n) route_dests(a, i, d0, d1);
  uint32_t before[ROUTE_MAX_WORLD];  // lanes of this warp ahead of me going to d (registers: world <= 16)
#pragma unroll
  for (int d = 0; d < ROUTE_MAX_WORLD; ++d) {
    before[d] = 0;
    if (d < a.world) {
      const uint32_t m = __ballot_sync(0xffffffffu, d >= d0 && d <= d1);
      before[d] = __popc(m & ((1u << lane) - 1u));
      if (lane == 0) warp_cnt[d][warp] = __popc(m);
    }
  }
  __syncthreads();
  if (d0 > d1) return;
  double v[ROUTE_MAX_FIELDS];
#pragma unroll
  for (int f = 0; f < ROUTE_MAX_FIELDS; ++f) v[f] = f < s.n_fields ? s.src[f][i] : 0.0;
#pragma unroll
  for (int d = 0; d < ROUTE_MAX_WORLD; ++d) {
    if (d < a.world && d >= d0 && d <= d1) {
      uint32_t pos = before[d];
      for (int w = 0; w < warp; ++w) pos += warp_cnt[d][w];
      const int64_t at = s.src_off[d] + (int64_t)s.blk_off[(int64_t)d * nblk + blockIdx.x] + pos;
      if (at < s.capacity) {  // (the host checks the totals against the capacity before reading)
        double* base = s.dst[d];
#pragma unroll
        for (int f = 0; f < ROUTE_MAX_FIELDS; ++f)
          if (f < s.n_fields) base[(int64_t)f * s.capacity + at] = v[f];
      }
    }
  }
}

__global__ void widen_totals_kernel(const uint32_t* __restrict__ t32, int world, int64_t* __restrict__ t64) {
  if ((int)threadIdx.x < world) t64[threadIdx.x] = (int64_t)t32[threadIdx.x];
}

}  // namespace mtn
