// K0 smoothing setup, K1 prune, and the footprint / binning front half of the projection.
#pragma once

#include "common.cuh"
#include "scan.cuh"

namespace mtn {

// 256-thread blocks: at ~50 registers five of them share an SM (a 1024-thread block was alone on
// it: half the warp slots idle while every particle waits on its loads)
constexpr int PLAN_THREADS = 256;
constexpr int TILE_STAT_STRIDE = 32;

// ---------------------------------------------------------------------------------------
// K0: per-particle kernel choice, sm_range, h_eff.
//   sph_kernels.py:257-262 (_init_sm_ranges), :1254-1272 (_AdaptiveKernel._init_sm_lengths),
//   :116 (rescaled_h), :121-138 (_confirm_validation's `valid`).
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) smoothing_setup_kernel(
    int64_t n, const double* __restrict__ sm_length, KernelTableDev t,
    uint8_t* __restrict__ kid_out, uint8_t* __restrict__ valid_out,
    double* __restrict__ range_out, double* __restrict__ heff_out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double s = sm_length[i];
  int kid = 0, valid = 0;
  if (t.adaptive) {
    // first kernel whose _validate(sm_lengths * K._rescale) passes; none -> entry 0
    for (int k = 0; k < t.n; ++k) {
      const double x = __dmul_rn(s, t.rescale[k]);
      const bool ok = t.valid_is_max[k] ? (x <= t.valid_size[k]) : (x >= t.valid_size[k]);
      if (ok) {
        kid = k;
        valid = 1;
        break;
      }
    }
  } else {
    // a simple kernel validates the raw smoothing length (:138)
    valid = t.valid_is_max[0] ? (s <= t.valid_size[0]) : (s >= t.valid_size[0]);
  }
  if (kid_out) kid_out[i] = (uint8_t)kid;
  if (valid_out) valid_out[i] = (uint8_t)valid;
  if (range_out) range_out[i] = ceil(__dmul_rn(s, t.size_in_fwhm[kid]));
  if (heff_out) heff_out[i] = __dmul_rn(s, t.rescale[kid]);
}

// ---------------------------------------------------------------------------------------
// K1: prune mask, martini.py:198-232 (bit-exact: same IEEE comparisons as numpy).
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) prune_kernel(
    int64_t n, const double* __restrict__ px, const double* __restrict__ py,
    const double* __restrict__ pz, const double* __restrict__ sm_range,
    const double* __restrict__ mHI, double mHI_scalar, const double* __restrict__ half_width,
    double hw_scalar, double max_abs_dv, double nx_tot, double ny_tot, double n_channels,
    int flags, uint8_t* __restrict__ accept, unsigned long long* __restrict__ n_accept) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  bool keep = false;
  if (i < n) {
    bool reject = false;
    if (flags & MTN_PRUNE_SPATIAL) {
      const double x = px[i], y = py[i], r = sm_range[i];
      reject |= __dadd_rn(x, r) < 0.0;
      reject |= __dadd_rn(y, r) < 0.0;
      reject |= __dsub_rn(x, r) > nx_tot;
      reject |= __dsub_rn(y, r) > ny_tot;
      reject |= isnan(x) || isnan(y);
    }
    if (flags & MTN_PRUNE_SPECTRAL) {
      const double hw = half_width ? half_width[i] : hw_scalar;
      const double w4 = __dmul_rn(4.0, __ddiv_rn(hw, max_abs_dv));
      const double z = pz[i];
      reject |= __dadd_rn(z, w4) < 0.0;
      reject |= __dsub_rn(z, w4) > n_channels;
    }
    if (flags & MTN_PRUNE_MASS) {
      const double m = mHI ? mHI[i] : mHI_scalar;
      reject |= (m == 0.0);
    }
    keep = !reject;
    accept[i] = keep ? 1 : 0;
  }
  if (n_accept) {  // one atomic per block (one per warp cost 0.3 ms of serialised L2 atomics at 1e7 particles)
    __shared__ unsigned int warp_cnt[8];
    const unsigned b = __ballot_sync(0xffffffffu, keep);
    if ((threadIdx.x & 31) == 0) warp_cnt[threadIdx.x >> 5] = __popc(b);
    __syncthreads();
    if (threadIdx.x == 0) {
      unsigned int c = 0;
      for (int w = 0; w < 8; ++w) c += warp_cnt[w];
      if (c) atomicAdd(n_accept, (unsigned long long)c);
    }
  }
}

// ---------------------------------------------------------------------------------------
// Footprint of one particle in the slab.
// ---------------------------------------------------------------------------------------
struct Foot {
  int i0, i1, j0, j1;  // inclusive pixel bounds (full-cube coordinates), clipped to the slab
  int c0, c1;          // inclusive live-channel window
  bool box;            // candidate box reaches the slab (counts towards U_dense)
  bool live;           // box && a live channel exists (&& a pixel with non-zero weight, COLUMN)
  int route;           // Route: which kernel computes this particle
  int nbx, nby;        // extent of the reference's candidate box in the slab (U_dense)
  double px, py, s2;   // position and squared support radius [px^2] (tile culling, brick route)
};

// What the count pass leaves for the emit pass (24 bytes per particle), so the footprint --
// two exact pixel-bound searches and two binary searches over the channel edges -- is
// evaluated once.
struct __align__(8) PackedFoot {
  int32_t i0, i1, j0, j1;
  uint16_t c0, c1;
  uint8_t live, route;
  uint16_t pad;
};
__device__ __forceinline__ PackedFoot pack_foot(const Foot& f) {
  PackedFoot p;
  p.i0 = f.i0; p.i1 = f.i1; p.j0 = f.j0; p.j1 = f.j1;
  p.c0 = (uint16_t)(f.live ? f.c0 : 0);
  p.c1 = (uint16_t)(f.live ? f.c1 : 0);
  p.live = f.live ? 1 : 0;
  p.route = (uint8_t)f.route;
  p.pad = 0;
  return p;
}
// Smallest and largest integer i in [lim_lo, lim_hi] with |i - p| <= r, the candidate test
// of martini.py:272-274 evaluated exactly as numpy does (fl(i - p), then compare).
__device__ __forceinline__ bool pixel_bounds(double p, double r, int lim_lo, int lim_hi, int& lo,
                                             int& hi) {
  if (!(r >= 0.0) || isnan(p) || lim_lo > lim_hi) return false;
  auto in = [&](int i) { return fabs(__dsub_rn((double)i, p)) <= r; };
  double a0 = ceil(p - r), b0 = floor(p + r);
  a0 = fmin(fmax(a0, (double)lim_lo - 1.0), (double)lim_hi + 1.0);
  b0 = fmin(fmax(b0, (double)lim_lo - 1.0), (double)lim_hi + 1.0);
  int a = (int)a0, b = (int)b0;
  if (a - 1 >= lim_lo && in(a - 1)) --a;
  if (a < lim_lo) a = lim_lo;
  if (a <= lim_hi && !in(a)) ++a;
  if (a <= lim_hi && !in(a)) ++a;
  if (b + 1 <= lim_hi && in(b + 1)) ++b;
  if (b > lim_hi) b = lim_hi;
  if (b >= lim_lo && !in(b)) --b;
  if (b >= lim_lo && !in(b)) --b;
  if (a > b || a > lim_hi || b < lim_lo) return false;
  if (!in(a) || !in(b)) return false;
  lo = a;
  hi = b;
  return true;
}

// Live channel window.  g(e) = sgn * (edges[e] - v) * inv_s is non-decreasing in e;
// channel c can be non-zero only if g(c+1) > -T and g(c) < T (T = ERF_SAT for the Gaussian
// line, closed at 0 for the Dirac line).  Returns false if no channel is live.
// Smallest e in [0, C + 1) with pred(e) (pred is monotone: false ... false true ... true);
// C + 1 if there is none.  `guess` is where a uniform channel grid puts the answer: two probes
// around it bracket the answer to five candidates (on a uniform grid always), otherwise the
// search falls back to plain bisection of what is left -- the result never depends on the guess.
template <typename Pred>
__device__ __forceinline__ int first_true(Pred pred, int C, int guess) {
  int lo = 0, hi = C + 1;
  const int a = min(max(guess - 2, 0), C), b = min(a + 4, C);
  if (a > 0) {
    if (pred(a - 1)) hi = a - 1; else lo = a;
  }
  if (lo == a) {
    if (pred(b)) hi = b; else lo = b + 1;
  }
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (pred(mid)) hi = mid; else lo = mid + 1;
  }
  return lo;
}

// Where a uniform channel grid puts an edge value: e = (x - e0) * inv_step.  The same for every
// particle, so a block computes it once (one division).
struct ChanGrid {
  double e0, inv_step;
};
__device__ __forceinline__ ChanGrid chan_grid(const double* __restrict__ edges, int C) {
  ChanGrid cg;
  cg.e0 = __ldg(edges);
  cg.inv_step = (double)C / (__ldg(edges + C) - cg.e0);
  return cg;
}

// `sg`: sigma of the Gaussian line (only steers the guesses); inv_s = 1 / (sqrt(2) sigma).
__device__ __forceinline__ bool channel_window(const double* __restrict__ edges, int C, int sgn,
                                               int spectrum, double v, double inv_s, double sg,
                                               const ChanGrid& cg, int& c0, int& c1) {
  const bool dirac = spectrum == MTN_SPECTRUM_DIRACDELTA;
  const double scale = dirac ? (double)sgn : (double)sgn * inv_s;
  auto g = [&](int e) { return (__ldg(edges + e) - v) * scale; };
  // where a uniform grid has g(e) = x:  e = (x / scale + v - edges[0]) / mean channel width
  const double half = dirac ? 0.0 : (double)sgn * (ERF_SAT * 1.4142135623730951) * sg;  // = ERF_SAT / scale
  auto guess = [&](double x_over_scale) {
    const double e = (x_over_scale + v - cg.e0) * cg.inv_step;
    return (int)fmin(fmax(e, -1.0), (double)C + 1.0);  // (NaN -> 0 on the device: any guess is fine)
  };
  // first edge e in [0, C] with g(e) > -T (>= 0 for dirac)
  const int e_first = first_true([&](int e) { const double x = g(e); return dirac ? (x >= 0.0) : (x > -ERF_SAT); },
                                 C, guess(-half));  // C+1 if none
  // last edge e in [0, C] with g(e) < T (<= 0 for dirac): first e failing, minus one
  const int e_last = first_true([&](int e) { const double x = g(e); return !(dirac ? (x <= 0.0) : (x < ERF_SAT)); },
                                C, guess(half) + 1) - 1;  // -1 if none
  if (e_first > C || e_last < 0) return false;
  c0 = max(e_first - 1, 0);
  c1 = min(e_last, C - 1);
  return c0 <= c1;
}

struct PlanIn {
  int64_t n;
  const double* px;
  const double* py;
  const double* h_eff;
  const double* sm_range;
  const uint8_t* kernel_id;
  const double* v;
  const double* sigma;
  double sigma_scalar;
  const double* mHI;
  double mHI_scalar;
  const double* D;
  double D_scalar;
  const uint8_t* accept;
  const double* edges;
};

// The one pixel a DiracDelta-kernel particle reaches: |p - i| < 0.5 on both axes, strict
// (sph_kernels.py:1165), evaluated as numpy does (fl(p - i)); false if there is none (a
// particle exactly on a pixel edge lands nowhere) or it lies outside [lo, hi].
__device__ __forceinline__ bool dirac_pixel(double p, int lo, int hi, int& i) {
  if (!(fabs(p) < 1.0e9)) return false;
  const int k = (int)rint(p);
  if (k < lo || k > hi || !(fabs(__dsub_rn(p, (double)k)) < 0.5)) return false;
  i = k;
  return true;
}

__device__ __forceinline__ Foot footprint(const PlanIn& in, const Geo& g, int64_t i, const ChanGrid& cg) {
  Foot f;
  f.box = f.live = false;
  f.route = ROUTE_BRICK;
  // every input up front: independent loads, one exposed memory latency instead of five in a
  // row behind the early returns (all in bounds for a rejected particle too)
  const bool accepted = in.accept ? in.accept[i] != 0 : true;
  const double r = in.sm_range[i], px = in.px[i], py = in.py[i], h = in.h_eff[i], v = in.v[i];
  const int kid = in.kernel_id ? in.kernel_id[i] : 0;
  const bool gauss = g.spectrum == MTN_SPECTRUM_GAUSSIAN;
  const double sg = gauss ? (in.sigma ? in.sigma[i] : in.sigma_scalar) : 1.0;
  if (!accepted) return f;
  if (!pixel_bounds(px, r, g.x_lo, g.x_hi - 1, f.i0, f.i1)) return f;
  if (!pixel_bounds(py, r, 0, g.ny - 1, f.j0, f.j1)) return f;
  f.box = true;
  f.px = px;
  f.py = py;
  {
    const double s = h * g.support[kid];
    f.s2 = s * s * (1.0 + 1.0e-9);  // (a tile is culled only if it is clearly outside)
  }
  f.nbx = f.i1 - f.i0 + 1;
  f.nby = f.j1 - f.j0 + 1;
  const double inv_s = gauss ? 1.0 / (1.4142135623730951 * sg) : 1.0;
  f.live = channel_window(in.edges, g.C, g.edges_increasing ? 1 : -1, g.spectrum, v, inv_s, sg, cg, f.c0, f.c1);
  if (!f.live) return f;
  if (g.route2 == ROUTE_SPLAT) {
    f.route = ROUTE_SPLAT;
  } else if (g.route2 == ROUTE_COLUMN && g.kind[kid] == MTN_KERNEL_DIRACDELTA) {
    // the candidate box shrinks to the single pixel with non-zero weight (or to nothing)
    f.route = ROUTE_COLUMN;
    int x, y;
    if (dirac_pixel(px, f.i0, f.i1, x) && dirac_pixel(py, f.j0, f.j1, y)) {
      f.i0 = f.i1 = x;
      f.j0 = f.j1 = y;
    } else {
      f.live = false;
    }
  }
  return f;
}

// Tile range of a footprint, and the channel blocks it reaches in one tile.  Channel blocks
// are CB wide with a per-tile phase ph in [0, CB): block k holds channels
// [ph + (k-1) CB, ph + k CB), so block 0 is the (possibly empty) part below ph.  The phase is
// chosen per tile (tile_phase_kernel) so that the tile's mean line centre sits mid-block:
// most particles of a tile then need a single brick instead of straddling a fixed boundary.
__device__ __forceinline__ void tile_range(const Foot& f, const Geo& g, int& tx0, int& tx1,
                                           int& ty0, int& ty1) {
  tx0 = (f.i0 - g.x_lo) / TILE_X;
  tx1 = (f.i1 - g.x_lo) / TILE_X;
  ty0 = f.j0 / TILE_Y;
  ty1 = f.j1 / TILE_Y;
}
__device__ __forceinline__ void block_range(const Foot& f, int ph, int& k0, int& k1) {
  k0 = (f.c0 - ph + CB) / CB;
  k1 = (f.c1 - ph + CB) / CB;
}

// Pass 0a: per tile, sum and count of the line centres (in half channels) of the particles
// that reach it.  Integer atomics: the result does not depend on the order of arrival.
__global__ void __launch_bounds__(PLAN_THREADS) tile_stats_kernel(
    PlanIn in, Geo g, unsigned long long* __restrict__ tile_sum, unsigned int* __restrict__ tile_cnt) {
  // the phase is a heuristic: a 1-in-TILE_STAT_STRIDE sample of the particles decides it
  const int64_t i = ((int64_t)blockIdx.x * PLAN_THREADS + threadIdx.x) * TILE_STAT_STRIDE;
  if (i >= in.n) return;
  const Foot f = footprint(in, g, i, chan_grid(in.edges, g.C));
  if (!f.live || f.route != ROUTE_BRICK) return;
  int tx0, tx1, ty0, ty1;
  tile_range(f, g, tx0, tx1, ty0, ty1);
  const unsigned long long centre2 = (unsigned long long)(f.c0 + f.c1 + 1);
  for (int tx = tx0; tx <= tx1; ++tx)
    for (int ty = ty0; ty <= ty1; ++ty) {
      atomicAdd(tile_sum + tx * g.nty + ty, centre2);
      atomicAdd(tile_cnt + tx * g.nty + ty, 1u);
    }
}

// Pass 0b: phase = (mean centre - CB/2) mod CB, even (keeps 16-byte store alignment).
__global__ void __launch_bounds__(256) tile_phase_kernel(int n_tiles,
                                                         const unsigned long long* __restrict__ tile_sum,
                                                         const unsigned int* __restrict__ tile_cnt,
                                                         int* __restrict__ phase) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_tiles) return;
  int ph = 0;
  if (tile_cnt[t]) {
    const long long mean = (long long)(tile_sum[t] / (2ull * tile_cnt[t]));
    ph = (int)(((mean - CB / 2) % CB + CB) % CB) & ~1;
  }
  phase[t] = ph;
}

// Whether any pixel of tile (tx, ty) inside the particle's box can have a non-zero weight: the
// square candidate box of the reference (martini.py:272-274) has corners the kernel's round
// support does not reach, and W = 0 exactly there -- 18 % of the (particle, tile) pairs of
// config 2 hold nothing but such zeros.  Distance from the particle to the nearest pixel centre
// of the tile's share of the box, against the support radius (with a relative margin of 1e-9
// on the safe side).
__device__ __forceinline__ bool tile_reached(const Foot& f, const Geo& g, int tx, int ty) {
  const double xl = (double)max(g.x_lo + tx * TILE_X, f.i0), xh = (double)min(g.x_lo + tx * TILE_X + TILE_X - 1, f.i1);
  const double yl = (double)max(ty * TILE_Y, f.j0), yh = (double)min(ty * TILE_Y + TILE_Y - 1, f.j1);
  const double dx = fmax(fmax(xl - f.px, f.px - xh), 0.0), dy = fmax(fmax(yl - f.py, f.py - yh), 0.0);
  return !(dx * dx + dy * dy >= f.s2);  // (NaN / inf support: reached)
}

// Pairs a particle contributes to its stream: (particle, brick) for the brick kernel,
// (particle, pixel x channel superblock) for COLUMN, (particle, tile x channel) for SPLAT.
__device__ __forceinline__ int64_t count_pairs(const Foot& f, const Geo& g) {
  if (f.route == ROUTE_COLUMN) return f.c1 / CSB - f.c0 / CSB + 1;
  int tx0, tx1, ty0, ty1;
  tile_range(f, g, tx0, tx1, ty0, ty1);
  if (f.route == ROUTE_SPLAT) {
    int64_t tiles = 0;
    for (int tx = tx0; tx <= tx1; ++tx)
      for (int ty = ty0; ty <= ty1; ++ty) tiles += tile_reached(f, g, tx, ty) ? 1 : 0;
    return tiles * (f.c1 - f.c0 + 1);
  }
  int64_t n = 0;
  for (int tx = tx0; tx <= tx1; ++tx)
    for (int ty = ty0; ty <= ty1; ++ty) {
      if (!tile_reached(f, g, tx, ty)) continue;
      int k0, k1;
      block_range(f, g.phase[tx * g.nty + ty], k0, k1);
      n += k1 - k0 + 1;
    }
  return n;
}

// Pass 1: per-block totals of (kept particles, bricks overlapped) and the slab's U_dense.
__global__ void __launch_bounds__(PLAN_THREADS) plan_count_kernel(
    PlanIn in, Geo g, int64_t* __restrict__ blk_kept, int64_t* __restrict__ blk_pairs,
    int64_t* __restrict__ blk_pairs2, unsigned long long* __restrict__ updates,
    PackedFoot* __restrict__ feet) {
  const int64_t i = (int64_t)blockIdx.x * PLAN_THREADS + threadIdx.x;
  __shared__ ChanGrid s_cg;
  __shared__ unsigned long long wtot[PLAN_THREADS / 32][4];
  if (threadIdx.x == 0) s_cg = chan_grid(in.edges, g.C);
  __syncthreads();
  // per particle: kept 0/1, pairs of either stream (< 2^24: tiles x channel blocks), candidate
  // box area (U_dense / C)
  uint32_t kept = 0, pairs = 0, pairs2 = 0;
  uint64_t area = 0;
  if (i < in.n) {
    const Foot f = footprint(in, g, i, s_cg);
    feet[i] = pack_foot(f);
    // U_dense counts the reference's candidate box, whatever kernel computes the particle
    if (f.box) area = (uint64_t)f.nbx * (uint64_t)f.nby;
    if (f.live) {
      kept = 1;
      (f.route == ROUTE_BRICK ? pairs : pairs2) = (uint32_t)count_pairs(f, g);
    }
  }
  // block totals: one warp-wide integer reduction (REDUX) per quantity -- the area in three
  // 16-bit limbs so that 32 of them cannot overflow --, per-warp partials in shared memory, one
  // barrier.  (Round 2 reduced four 64-bit values by shuffles and added them with shared-memory
  // 64-bit atomics, a CAS loop under contention: a tenth of the kernel's instructions and most
  // of its barrier stalls.)
  {
    const uint32_t wk = __popc(__ballot_sync(0xffffffffu, kept != 0));
    const uint32_t wp = __reduce_add_sync(0xffffffffu, pairs);
    const uint32_t wp2 = __reduce_add_sync(0xffffffffu, pairs2);
    const uint32_t a0 = __reduce_add_sync(0xffffffffu, (uint32_t)(area & 0xffffu));
    const uint32_t a1 = __reduce_add_sync(0xffffffffu, (uint32_t)((area >> 16) & 0xffffu));
    const uint32_t a2 = __reduce_add_sync(0xffffffffu, (uint32_t)(area >> 32));
    if ((threadIdx.x & 31) == 0) {
      unsigned long long* w = wtot[threadIdx.x >> 5];
      w[0] = wk;
      w[1] = wp;
      w[2] = wp2;
      w[3] = (unsigned long long)a0 + ((unsigned long long)a1 << 16) + ((unsigned long long)a2 << 32);
    }
  }
  __syncthreads();
  if (threadIdx.x < 4) {
    unsigned long long t = 0;
    for (int w = 0; w < PLAN_THREADS / 32; ++w) t += wtot[w][threadIdx.x];
    if (threadIdx.x == 0) blk_kept[blockIdx.x] = (int64_t)t;
    if (threadIdx.x == 1) blk_pairs[blockIdx.x] = (int64_t)t;
    if (threadIdx.x == 2) blk_pairs2[blockIdx.x] = (int64_t)t;
    if (threadIdx.x == 3 && t) atomicAdd(updates, t * (unsigned long long)g.C);
  }
}

// Pass 3 (after the block sums were scanned): write one record (REC_BYTES = 80) per kept particle
// and one (brick key << 32 | record index) pair per brick it overlaps, in particle order.
__global__ void __launch_bounds__(PLAN_THREADS) plan_emit_kernel(
    PlanIn in, Geo g, const int64_t* __restrict__ blk_kept, const int64_t* __restrict__ blk_pairs,
    const int64_t* __restrict__ blk_pairs2, const PackedFoot* __restrict__ feet,
    Record* __restrict__ records, uint64_t* __restrict__ pairs_out, uint64_t* __restrict__ pairs2_out) {
  __shared__ uint32_t sm3[3][33];
  // the block's records, in order: staged here and copied out as one contiguous run of 16-byte
  // stores (a thread storing its own 80-byte record touches five half-used sectors per
  // instruction; that store was 17 % of the kernel's stall samples, queue-throttled)
  __shared__ Record srec[PLAN_THREADS];
  const int64_t i = (int64_t)blockIdx.x * PLAN_THREADS + threadIdx.x;
  uint32_t kept = 0, npair = 0, npair2 = 0;
  Foot f;
  f.live = false;
  f.route = ROUTE_BRICK;
  double px = 0.0, py = 0.0, h = 0.0, v = 0.0, m = 0.0, d = 1.0, sg = 1.0;
  int kid = 0;
  if (i < in.n) {
    const PackedFoot pf = feet[i];  // (computed by plan_count_kernel)
    if (pf.live) {  // the particle's quantities, all loads in flight together
      px = in.px[i];
      py = in.py[i];
      h = in.h_eff[i];
      v = in.v[i];
      kid = in.kernel_id ? in.kernel_id[i] : 0;
      m = in.mHI ? in.mHI[i] : in.mHI_scalar;
      d = in.D ? in.D[i] : in.D_scalar;
      if (g.spectrum == MTN_SPECTRUM_GAUSSIAN) sg = in.sigma ? in.sigma[i] : in.sigma_scalar;
      f.px = px;
      f.py = py;
      const double s = h * g.support[kid];
      f.s2 = s * s * (1.0 + 1.0e-9);
      f.i0 = pf.i0; f.i1 = pf.i1; f.j0 = pf.j0; f.j1 = pf.j1;
      f.c0 = pf.c0; f.c1 = pf.c1;
      f.live = f.box = true;
      f.route = pf.route;
      f.nbx = f.nby = 0;
      kept = 1;
      (f.route == ROUTE_BRICK ? npair : npair2) = (uint32_t)count_pairs(f, g);
    }
  }
  // (32-bit prefixes inside a block: a slab holds fewer than 2^32 pairs per stream, mtn_plan
  // refuses more)
  uint32_t ex[3] = {kept, npair, npair2};
  block_excl_scan3(ex, sm3);  // one pass, two barriers, for the three prefixes
  const int64_t ridx = blk_kept[blockIdx.x] + ex[0];
  int64_t off = blk_pairs[blockIdx.x] + ex[1];
  int64_t off2 = blk_pairs2[blockIdx.x] + ex[2];
  if (f.live) {
    Record rec;
    rec.px = px;
    rec.py = py;
    rec.h = h;
    rec.inv_h2 = f.route == ROUTE_COLUMN ? 0.0 : 1.0 / (h * h);  // (a DiracDelta kernel has no scale)
    rec.v = v;
    // A = mHI * D^-2 (spectral_models.py:94), / 2.36e5 (:139); the Gaussian line's 0.5 is
    // folded in here (exact: a power of two)
    const double amp = m * (1.0 / (d * d)) / 2.36e5;
    if (g.spectrum == MTN_SPECTRUM_GAUSSIAN) {
      rec.inv_s = 1.0 / (1.4142135623730951 * sg);  // (as footprint() formed it for the window)
      rec.amp = 0.5 * amp;
    } else {
      rec.inv_s = 1.0;
      rec.amp = amp;
    }
    rec.i0 = f.i0;
    rec.i1 = f.i1;
    rec.j0 = f.j0;
    rec.j1 = f.j1;
    rec.c_first = (uint16_t)f.c0;  // mtn_plan refuses cubes with more than 65535 channels
    rec.c_last = (uint16_t)f.c1;
    rec.kid = (uint8_t)kid;
    for (int k = 0; k < 3; ++k) rec.pad[k] = 0;
    srec[ex[0]] = rec;
  }
  __syncthreads();
  {
    // records [blk_kept[b], blk_kept[b] + kept in this block) as 16-byte words
    const uint32_t n_words = (sm3[0][32]) * (REC_BYTES / 16);
    const int4* src = reinterpret_cast<const int4*>(srec);
    int4* dst = reinterpret_cast<int4*>(records + blk_kept[blockIdx.x]);
    for (uint32_t w = threadIdx.x; w < n_words; w += PLAN_THREADS) dst[w] = src[w];
  }
  if (!f.live) return;
  if (f.route == ROUTE_COLUMN) {
    const int64_t pixel = (int64_t)(f.i0 - g.x_lo) * g.ny + f.j0;
    for (int sb = f.c0 / CSB; sb <= f.c1 / CSB; ++sb)
      pairs2_out[off2++] = ((uint64_t)(uint32_t)(pixel * g.nsb + sb) << 32) | (uint64_t)(uint32_t)ridx;
    return;
  }
  int tx0, tx1, ty0, ty1;
  tile_range(f, g, tx0, tx1, ty0, ty1);
  for (int tx = tx0; tx <= tx1; ++tx)
    for (int ty = ty0; ty <= ty1; ++ty) {
      const int tile = tx * g.nty + ty;
      if (!tile_reached(f, g, tx, ty)) continue;  // (all pixels outside the kernel's support)
      if (f.route == ROUTE_SPLAT) {
        for (int c = f.c0; c <= f.c1; ++c)
          pairs2_out[off2++] = ((uint64_t)(uint32_t)((int64_t)tile * g.C + c) << 32) | (uint64_t)(uint32_t)ridx;
        continue;
      }
      int k0, k1;
      block_range(f, g.phase[tile], k0, k1);
      for (int k = k0; k <= k1; ++k) {
        const uint32_t key = (uint32_t)(tile * g.ncb + k);
        pairs_out[off++] = ((uint64_t)key << 32) | (uint64_t)(uint32_t)ridx;
      }
    }
}

}  // namespace mtn
