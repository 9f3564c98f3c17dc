// The two degenerate particle shapes, each with a kernel whose lanes lie along the one axis
// that is left (common.cuh: Route).  Both run AFTER the brick kernel on the same stream and
// add their sums onto the voxels it has written (read-modify-write; no two work items of a
// kernel own the same voxel, so there are no atomics on the data path and the result does not
// depend on scheduling).  Work items are cut from the stably sorted pair array exactly like
// the brick kernel's (project.cuh: brick_bounds / item_count / item_fill): a key's particles
// stay in index order, a key with more than `chunk` particles is cut into chunks whose partial
// sums are added in chunk order by the reduce kernels below.
//
//   column_kernel  DiracDelta SPH KERNEL + Gaussian line (BASELINE config 3: 95 % of the
//                  particles are smaller than a pixel).  One warp = one (pixel, channel
//                  superblock); lanes = channel edges of the particle's live window: two
//                  independent erf chains per lane, adjacent edges differenced by shuffle, the
//                  line profile added to a per-warp shared-memory accumulator of the pixel's
//                  spectrum.  The erfs are evaluated once per PARTICLE -- not once per
//                  (particle, brick) -- and no (particle, pixel) work exists at all.
//   splat_kernel   DiracDelta SPECTRUM (BASELINE config 4): one warp = one (tile, channel);
//                  lanes = the tile's 64 pixels, two per lane; the kernel integral is evaluated
//                  by the lane that accumulates it (particle warp-uniform: no enumeration, no
//                  shared-memory round trip of W, no kernel-kind divergence) and the whole
//                  spectrum is one scalar.
#pragma once

#include "common.cuh"
#include "kernel_integrals.cuh"
#include "project.cuh"

namespace mtn {

constexpr int STREAM_WARPS = 4;
constexpr int STREAM_THREADS = STREAM_WARPS * 32;

// 1 / |channel width| of every channel, once per insertion (the stream kernels read it per
// particle; the brick kernel builds its own 64-channel copy in shared memory).
__global__ void __launch_bounds__(256) inv_dv_kernel(const double* __restrict__ edges, int C,
                                                     double* __restrict__ inv_dv) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < C) inv_dv[c] = 1.0 / fabs(edges[c + 1] - edges[c]);
}

struct StreamArgs {
  Geo geo;
  KernelTableDev table;
  const Record* records;
  const uint64_t* pairs;  // sorted by key
  const Item* items;
  const uint32_t* n_items;  // device scalar
  unsigned int* counter;    // device work counter, zeroed before launch
  const double* edges;
  const double* inv_dv;
  double* slab;
  double* partials;  // column: [slot][CSB]; splat: [slot][TILE_PIX]
  double px_area;
  unsigned long long* exec_counts;  // COUNT instantiation only: [updates, weights, erfs]
};

// Next work item of this warp (warp-uniform).
__device__ __forceinline__ bool next_item(const StreamArgs& a, int lane, Item& it) {
  uint32_t idx = 0;
  if (lane == 0) idx = atomicAdd(a.counter, 1u);
  idx = __shfl_sync(0xffffffffu, idx, 0);
  if (idx >= *a.n_items) return false;
  it = a.items[idx];
  return true;
}

// ------------------------------------------------------------------------------- column
__host__ __device__ inline int column_acc_stride(int C) { return C >= CSB ? CSB : (C + 63) / 64 * 64; }
// what the column kernel needs of a particle, staged per warp for its batch of 32: 32 bytes,
// read back as two 16-byte loads
struct __align__(16) ColRec {
  double v, sc;    // line centre, sgn / (sqrt(2) sigma)
  double amp;
  uint32_t cw;     // live channels of the superblock: cs | ce << 16  (channels [cs, ce), edges cs..ce)
  uint32_t start;  // first item of the particle's edge run in the batch's enumeration
};
inline size_t column_smem_bytes(int C) {
  return ((size_t)STREAM_WARPS * column_acc_stride(C) + ERFC_DOUBLES) * sizeof(double) +
         (size_t)STREAM_WARPS * 32 * sizeof(ColRec) + (size_t)STREAM_WARPS * 32;  // + owner lists
}

template <bool COUNT>
__global__ void __launch_bounds__(STREAM_THREADS) column_kernel(const StreamArgs a) {
  // [STREAM_WARPS][acc_stride] accumulators (acc_stride = channels of a superblock, rounded up
  // to 64: a 256-channel cube needs 2 KB per warp, not the 8 KB of a full superblock -- shared
  // memory is what limits the resident warps here), then a copy of the compact erf table
  MTN_DYN_SMEM(double, col_smem);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const Geo& g = a.geo;
  const int acc_stride = column_acc_stride(g.C);
  double* acc = col_smem + warp * acc_stride;
  double* erf_table = col_smem + STREAM_WARPS * acc_stride;
  ColRec* crec = reinterpret_cast<ColRec*>(erf_table + ERFC_DOUBLES) + warp * 32;  // (16-byte aligned: even counts)
  uint8_t* owner = reinterpret_cast<uint8_t*>(reinterpret_cast<ColRec*>(erf_table + ERFC_DOUBLES) + STREAM_WARPS * 32) + warp * 32;
  const double sgn = g.edges_increasing ? 1.0 : -1.0;
  for (int c = lane; c < acc_stride; c += 32) acc[c] = 0.0;
  for (int k = threadIdx.x; k < ERFC_DOUBLES; k += STREAM_THREADS) erf_table[k] = g_erf_table_compact[k];
  __syncthreads();
  unsigned long long n_upd = 0, n_w = 0, n_erf = 0;
  Item it;
  while (next_item(a, lane, it)) {
    const int64_t pixel = it.brick / g.nsb;
    const int cbase = (int)(it.brick % g.nsb) * CSB;
    const int nchs = min(CSB, g.C - cbase);  // channels of this superblock
    const double* edge = a.edges + cbase;
    const double* idv = a.inv_dv + cbase;
    int lo = CSB, hi = 0;  // touched range of the accumulator
    // The edges of the batch's particles are enumerated as ONE run of items (particle 0's live
    // edges, particle 1's, ...) and handed to the lanes 32 at a time, so every erf evaluation
    // of a warp has 32 (31) live lanes whatever the windows' lengths: the ~42-edge windows of
    // config 3 used to take two 32-lane chains each, the second a third full.  A step advances
    // by 31 items: lane 31 only supplies the upper edge of lane 30's channel and comes back as
    // lane 0 of the next step.  Lanes of different particles may hold the same channel; they
    // add to the accumulator particle by particle, in index order (deterministic, the order of
    // the reference's sum).
    const uint32_t lt = (1u << lane) - 1u;
    for (uint32_t base = it.begin; base < it.end; base += 32) {
      const int nb = (int)min(32u, it.end - base);
      __syncwarp();  // (the previous batch's records are no longer read)
      uint32_t ne = 0, my_start;
      {
        int cs = 0, ce = 0;
        ColRec c;
        c.v = c.sc = c.amp = 0.0;
        if (lane < nb) {
          const Record* r = a.records + (uint32_t)a.pairs[base + lane];
          c.v = r->v;
          c.sc = sgn * r->inv_s;
          c.amp = r->amp;
          // the particle's live window (plan.cuh: channel_window, exact predicates) cut to the
          // superblock: channels [cs, ce)
          cs = max((int)r->c_first - cbase, 0);
          ce = min((int)r->c_last + 1 - cbase, nchs);
          if (cs >= ce) cs = ce = 0;
        }
        ne = ce > cs ? (uint32_t)(ce - cs + 1) : 0u;
        const uint32_t incl = warp_incl_scan_u32(ne, lane);
        my_start = incl - ne;
        c.cw = (uint32_t)cs | ((uint32_t)ce << 16);
        c.start = my_start;
        crec[lane] = c;
        const uint32_t nonempty = __ballot_sync(0xffffffffu, ne != 0);
        if (ne != 0) {
          owner[__popc(nonempty & lt)] = (uint8_t)lane;
          lo = min(lo, cs);
          hi = max(hi, ce);
        }
        if (COUNT) n_w += ne != 0;
      }
      const uint32_t total = __shfl_sync(0xffffffffu, my_start + ne, 31);
      lo = (int)__reduce_min_sync(0xffffffffu, (unsigned)lo);
      hi = (int)__reduce_max_sync(0xffffffffu, (unsigned)hi);
      __syncwarp();
      const bool my_nonempty = ne != 0;
      for (uint32_t q0 = 0; q0 + 1 < total; q0 += 31) {
        const uint32_t q = q0 + lane;
        const bool valid = q < total;
        const int ord = owner_ordinal(q0, my_start, my_nonempty, lane);
        const int p = owner[valid ? ord : 0];
        const double2 vs = *reinterpret_cast<const double2*>(&crec[p].v);
        const int4 aw = *reinterpret_cast<const int4*>(&crec[p].amp);
        const int cs = aw.z & 0xffff, ce = (int)((uint32_t)aw.z >> 16);
        const int e = valid ? cs + (int)(q - (uint32_t)aw.w) : 0;  // edge (= channel below it) in the superblock
        // g orientation (sign folded in): S[c] = E[c+1] - E[c] >= 0; saturated edges at the ends
        // of the window come out as exactly -1 / +1
        const double t = (__ldg(edge + e) - vs.x) * vs.y;
        const double E = erf_tab_compact(erf_table, t);
        if (COUNT) n_erf += valid && (lane < 31 || q + 1 == total) && fabs(t) < ERF_SAT;
        const double U = __shfl_down_sync(0xffffffffu, E, 1);  // upper edge of my channel: the next item
        const bool active = valid && lane < 31 && e < ce;      // (e == ce: the particle's last edge, no channel)
        const double sval = (U - E) * __hiloint2double(aw.y, aw.x);
        // DiracDelta kernel: the weight of its one pixel is exactly 1
        uint32_t todo = __ballot_sync(0xffffffffu, active);
        while (todo) {
          const int pl = __shfl_sync(0xffffffffu, p, __ffs(todo) - 1);
          const bool mine = active && p == pl;
          if (mine) {
            acc[e] += sval;
            if (COUNT) n_upd += sval != 0.0;
          }
          todo &= ~__ballot_sync(0xffffffffu, mine);
          __syncwarp();
        }
      }
    }
    __syncwarp();
    // ---- flush the touched range, and clear it for the next item
    if (it.slot >= 0) {
      double* dst = a.partials + (size_t)it.slot * CSB;
      for (int c = lane; c < nchs; c += 32) {
        dst[c] = acc[c] * __ldg(idv + c);
        acc[c] = 0.0;
      }
    } else if (hi > lo) {
      double* dst = a.slab + (size_t)pixel * g.C + cbase;
      for (int c = (lo & ~31) + lane; c < hi; c += 32) {
        if (c >= lo) dst[c] += acc[c] * __ldg(idv + c) / a.px_area;  // 1 / dv of the channel (spectral_models.py:139)
        acc[c] = 0.0;
      }
    }
    __syncwarp();
  }
  if (COUNT) {
    atomicAdd(a.exec_counts + 0, n_upd);
    atomicAdd(a.exec_counts + 1, n_w);
    atomicAdd(a.exec_counts + 2, n_erf);
  }
}

// Multi-chunk columns: partial spectra summed in chunk order, then added onto the cube.
__global__ void __launch_bounds__(256) column_reduce_kernel(
    Geo g, const MultiBrick* __restrict__ multis, const uint32_t* __restrict__ n_multi,
    const double* __restrict__ partials, double* __restrict__ slab, double px_area) {
  if (blockIdx.x >= *n_multi) return;
  const MultiBrick m = multis[blockIdx.x];
  const int64_t pixel = m.brick / g.nsb;
  const int cbase = (int)(m.brick % g.nsb) * CSB;
  const int nchs = min(CSB, g.C - cbase);
  for (int c = threadIdx.x; c < nchs; c += blockDim.x) {
    double s = 0.0;
    for (uint32_t k = 0; k < m.n; ++k) s += partials[(size_t)(m.slot0 + k) * CSB + c];  // chunk order
    if (s != 0.0) slab[(size_t)pixel * g.C + cbase + c] += s / px_area;
  }
}

// ------------------------------------------------------------------------------- splat
// The Gaussian SPH kernel's pixel integral (sph_kernels.py:1024-1044) for the two pixels of a
// lane, with the separable erf factors taken once per pixel column / row of the tile: the nine
// x edges and nine y edges of the 8 x 8 tile are one erf each (lanes 0-8 and 16-24), adjacent
// pixels share an edge.  ez = erf(zmax / sqrt 2), zmax^2 = t^2 - (d / h / sig)^2, vanishes
// where the reference's truncation predicates zero the weight, continuously -- values agree
// with the closed form (kernel_integrals.cuh: w_gaussian) to a few ulp of the kernel's peak.
__device__ __forceinline__ void gaussian_pair(const double* __restrict__ erf_table, const Record& r,
                                              double truncate, double norm, double gx0, double gy,
                                              double x0d, double y0d, int lane, double& wA, double& wB) {
  const double sig = 0.42466090014400953;  // 1 / (2 sqrt(2 ln 2))
  const double c = 1.0 / (r.h * 1.4142135623730951 * sig);
  // edge k of an axis sits at pixel-centre k - 1/2: E_k = erf((p - (o + k) + 1/2) c)
  const int k = lane & 15;
  const double o = lane < 16 ? x0d : y0d, p = lane < 16 ? r.px : r.py;
  const double E = erf_tab_compact(erf_table, (p - (o + (double)k) + 0.5) * c);
  const double F = E - __shfl_down_sync(0xffffffffu, E, 1);  // lanes 0-7: columns, 16-23: rows
  const double exA = __shfl_sync(0xffffffffu, F, lane >> 3);
  const double exB = __shfl_sync(0xffffffffu, F, (lane >> 3) + 4);
  const double ey = __shfl_sync(0xffffffffu, F, 16 + (lane & 7));
  const double dy = __dsub_rn(r.py, gy);
  const double k2 = r.inv_h2 * (1.0 / (sig * sig)), t2 = truncate * truncate;
  const double q = 0.25 / norm;
  {
    const double dx = __dsub_rn(r.px, gx0);
    const double z2 = t2 - sq_dist(dx, dy) * k2;
    wA = z2 > 0.0 ? erf_tab_compact(erf_table, sqrt(0.5 * z2)) * exA * ey * q : 0.0;
  }
  {
    const double dx = __dsub_rn(r.px, gx0 + 4.0);
    const double z2 = t2 - sq_dist(dx, dy) * k2;
    wB = z2 > 0.0 ? erf_tab_compact(erf_table, sqrt(0.5 * z2)) * exB * ey * q : 0.0;
  }
}

struct SplatSmem {
  Record rec[STREAM_WARPS][PBATCH];
  double erf_table[ERFC_DOUBLES];  // (shared-memory copy of the compact table, tables.cuh)
};

template <bool COUNT>
__global__ void __launch_bounds__(STREAM_THREADS) splat_kernel(const StreamArgs a) {
  __shared__ SplatSmem sm;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  Record* rec = sm.rec[warp];
  const Geo& g = a.geo;
  for (int k = threadIdx.x; k < ERFC_DOUBLES; k += STREAM_THREADS) sm.erf_table[k] = g_erf_table_compact[k];
  __syncthreads();
  unsigned long long n_upd = 0, n_w = 0;
  Item it;
  while (next_item(a, lane, it)) {
    const int c = (int)(it.brick % g.C);
    const int tile = (int)(it.brick / g.C);
    const int x0 = g.x_lo + (tile / g.nty) * TILE_X, y0 = (tile % g.nty) * TILE_Y;
    // this lane's two pixels: (gxA, gy) and (gxA + 4, gy)
    const int gxA = x0 + (lane >> 3), gxB = gxA + 4, gy = y0 + (lane & 7);
    const double gxd = (double)gxA, gyd = (double)gy;
    const double idv = a.inv_dv[c];
    double acc0 = 0.0, acc1 = 0.0;
    for (uint32_t base = it.begin; base < it.end; base += 32) {
      const int nb = (int)min(32u, it.end - base);
      if (lane < nb) {  // stage the batch's records: one 80-byte record per lane
        const uint4* src = reinterpret_cast<const uint4*>(a.records + (uint32_t)a.pairs[base + lane]);
        uint4* dst = reinterpret_cast<uint4*>(rec + lane);
#pragma unroll
        for (int q = 0; q < REC_BYTES / 16; ++q) dst[q] = __ldg(src + q);
      }
      __syncwarp();
      for (int k = 0; k < nb; ++k) {
        const Record& r = rec[k];  // (warp-uniform)
        const double s = r.amp * idv;  // Dirac line: all of the particle's flux in this channel
        const int kind = g.kind[r.kid];
        const bool inY = gy >= r.j0 && gy <= r.j1;
        const bool inA = inY && gxA >= r.i0 && gxA <= r.i1, inB = inY && gxB >= r.i0 && gxB <= r.i1;
        double wA, wB;
        if (kind == MTN_KERNEL_GAUSSIAN) {
          gaussian_pair(sm.erf_table, r, a.table.truncate[r.kid], a.table.norm[r.kid], gxd, gyd, (double)x0, (double)y0,
                        lane, wA, wB);
        } else {
          // dij = pixcoords - ij (martini.py:276)
          const double dy = __dsub_rn(r.py, gyd);
          const double dxA = __dsub_rn(r.px, gxd), dxB = __dsub_rn(r.px, gxd + 4.0);
          if (wtab_has(kind)) {
            const double RA = sq_dist(dxA, dy) * r.inv_h2, RB = sq_dist(dxB, dy) * r.inv_h2;
            wA = wtab_eval(kind, RA) * r.inv_h2;  // straight-line: the two look-ups interleave
            wB = wtab_eval(kind, RB) * r.inv_h2;
          } else {
            wA = kernel_weight_closed(kind, dxA, dy, r.h, r.inv_h2, a.table.truncate[r.kid], a.table.norm[r.kid]);
            wB = kernel_weight_closed(kind, dxB, dy, r.h, r.inv_h2, a.table.truncate[r.kid], a.table.norm[r.kid]);
          }
        }
        wA = inA ? wA : 0.0;
        wB = inB ? wB : 0.0;
        acc0 = fma(wA, s, acc0);
        acc1 = fma(wB, s, acc1);
        if (COUNT) {
          n_w += (int)inA + (int)inB;
          n_upd += (wA != 0.0 && s != 0.0) + (wB != 0.0 && s != 0.0);
        }
      }
      __syncwarp();  // the records are free again
    }
    if (it.slot >= 0) {
      double* dst = a.partials + (size_t)it.slot * TILE_PIX;
      dst[lane] = acc0;
      dst[lane + 32] = acc1;
    } else {
      if (gy < g.ny) {
        if (gxA < g.x_hi && acc0 != 0.0) a.slab[((size_t)(gxA - g.x_lo) * g.ny + gy) * g.C + c] += acc0 / a.px_area;
        if (gxB < g.x_hi && acc1 != 0.0) a.slab[((size_t)(gxB - g.x_lo) * g.ny + gy) * g.C + c] += acc1 / a.px_area;
      }
    }
  }
  if (COUNT) {
    atomicAdd(a.exec_counts + 0, n_upd);
    atomicAdd(a.exec_counts + 1, n_w);
  }
}

// Multi-chunk (tile, channel) keys: partial maps summed in chunk order, then added onto the cube.
__global__ void __launch_bounds__(TILE_PIX) splat_reduce_kernel(
    Geo g, const MultiBrick* __restrict__ multis, const uint32_t* __restrict__ n_multi,
    const double* __restrict__ partials, double* __restrict__ slab, double px_area) {
  if (blockIdx.x >= *n_multi) return;
  const MultiBrick m = multis[blockIdx.x];
  const int c = (int)(m.brick % g.C), tile = (int)(m.brick / g.C);
  const int x0 = g.x_lo + (tile / g.nty) * TILE_X, y0 = (tile % g.nty) * TILE_Y;
  const int l = threadIdx.x & 31, hi = threadIdx.x >> 5;  // partial layout: [half][lane]
  const int gx = x0 + (l >> 3) + 4 * hi, gy = y0 + (l & 7);
  double s = 0.0;
  for (uint32_t k = 0; k < m.n; ++k) s += partials[(size_t)(m.slot0 + k) * TILE_PIX + threadIdx.x];  // chunk order
  if (s != 0.0 && gx < g.x_hi && gy < g.ny) slab[((size_t)(gx - g.x_lo) * g.ny + gy) * g.C + c] += s / px_area;
}

}  // namespace mtn
