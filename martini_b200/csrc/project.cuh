// The tile projection kernel and its small companions (work items, partial reduce).
//
// One CTA works on one brick (16 x 16 pixels x 32 channels) at a time.  Thread = pixel;
// each thread keeps its pixel's 32 channel sums in registers.  Particle records of the
// brick are gathered into shared memory with cp.async.bulk (one 64-byte bulk copy per
// record, completion on an mbarrier, double buffered).  Per batch the CTA evaluates each
// particle's channel spectrum once (edge erfs shared by adjacent channels) into shared
// memory; then every warp walks the batch: a warp-uniform box test skips particles that
// miss the warp's 4 x 8 pixel sub-block, each lane evaluates the kernel integral of its own
// pixel in registers, and the rank-1 update acc[c] += W * S[c] runs over the live channel
// groups only.  No atomics on the data path; every voxel is stored once.
#pragma once

#include "common.cuh"
#include "kernel_integrals.cuh"

namespace mtn {

// ----------------------------------------------------------------------------- PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  // bounded spin: a lost transaction traps instead of hanging the GPU
  for (uint32_t it = 0; it < (1u << 26); ++it)
    if (mbar_try_wait(bar, parity)) return;
  __trap();
}
// global -> shared bulk copy (TMA engine, SASS UBLKCP), completion bytes on `bar`
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes,
                                         uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
          "r"(smem_u32(dst_smem)),
      "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}

// ------------------------------------------------------------------------------ work items
// counts[b] = chunks of brick b; multi[b] = same if > 1 else 0; ismulti[b] = 0/1
__global__ void __launch_bounds__(256) item_count_kernel(const uint32_t* __restrict__ brick_count,
                                                         int n_bricks, uint32_t chunk,
                                                         uint32_t* __restrict__ counts,
                                                         uint32_t* __restrict__ multi,
                                                         uint32_t* __restrict__ ismulti) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= n_bricks) return;
  const uint32_t c = brick_count[b];
  const uint32_t k = (c + chunk - 1) / chunk;
  counts[b] = k;
  multi[b] = k > 1 ? k : 0;
  ismulti[b] = k > 1 ? 1 : 0;
}

struct MultiBrick {
  uint32_t brick, slot0, n;
  uint32_t pad;
};

__global__ void __launch_bounds__(256) item_fill_kernel(
    const uint32_t* __restrict__ brick_count, const uint32_t* __restrict__ brick_start,
    int n_bricks, uint32_t chunk, const uint32_t* __restrict__ item_start,
    const uint32_t* __restrict__ slot_start, const uint32_t* __restrict__ multi_idx,
    Item* __restrict__ items, MultiBrick* __restrict__ multis) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= n_bricks) return;
  const uint32_t c = brick_count[b];
  if (c == 0) return;
  const uint32_t k = (c + chunk - 1) / chunk;
  const uint32_t s0 = brick_start[b];
  for (uint32_t j = 0; j < k; ++j) {
    Item it;
    it.begin = s0 + j * chunk;
    it.end = min(s0 + c, it.begin + chunk);
    it.brick = (uint32_t)b;
    it.slot = k > 1 ? (int32_t)(slot_start[b] + j) : -1;
    items[item_start[b] + j] = it;
  }
  if (k > 1) {
    MultiBrick m;
    m.brick = (uint32_t)b;
    m.slot0 = slot_start[b];
    m.n = k;
    m.pad = 0;
    multis[multi_idx[b]] = m;
  }
}

// ---------------------------------------------------------------------- projection kernel
struct ProjArgs {
  Geo geo;
  KernelTableDev table;
  const Record* records;
  const uint64_t* pairs;  // sorted
  const Item* items;
  const uint32_t* n_items;  // device scalar
  unsigned int* counter;    // device work counter, zeroed before launch
  const double* edges;
  double* slab;
  double* partials;  // [slot][PROJ_THREADS][CB]
  double px_area;    // px_size_arcsec^2
  int zeroed;        // MTN_CUBE_ZEROED
  unsigned long long* exec_counts;  // COUNT instantiation only: [updates, weights, erfs]
};

struct ProjSmem {
  Record rec[2][PBATCH];
  double S[PBATCH][CB];
  double E[PBATCH][CB + 1];
  double edge[CB + 1];
  double inv_dv[CB];
  uint32_t gmask[PBATCH];
  uint64_t bar[2];
  uint32_t item;
};

// store one thread's CB channel sums: out = (in + acc) / px_area   (martini.py:338,364-366)
__device__ __forceinline__ void store_pixel(double* __restrict__ dst, const double* acc, int nch,
                                            double px_area, bool add_in, bool vec_ok) {
  if (vec_ok && nch == CB) {
#pragma unroll
    for (int c = 0; c < CB; c += 2) {
      double2 o;
      if (add_in) {
        const double2 i2 = *reinterpret_cast<const double2*>(dst + c);
        o.x = (i2.x + acc[c]) / px_area;
        o.y = (i2.y + acc[c + 1]) / px_area;
      } else {
        o.x = acc[c] / px_area;
        o.y = acc[c + 1] / px_area;
      }
      *reinterpret_cast<double2*>(dst + c) = o;
    }
  } else {
#pragma unroll
    for (int c = 0; c < CB; ++c)
      if (c < nch) dst[c] = ((add_in ? dst[c] : 0.0) + acc[c]) / px_area;
  }
}

// COUNT = true is a diagnostic instantiation that additionally tallies the executed
// algorithmic work (non-zero weight x non-zero spectrum terms, kernel integrals, edge erfs);
// it is never the timed kernel.
template <bool COUNT>
__global__ void __launch_bounds__(PROJ_THREADS, 2) project_kernel(const ProjArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  ProjSmem& sm = *reinterpret_cast<ProjSmem*>(smem_raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // warp -> 4 x 8 pixel sub-block of the tile, lane -> pixel inside it
  const int sx = (warp >> 1) * SUB_X, sy = (warp & 1) * SUB_Y;
  const int lx = lane >> 3, ly = lane & 7;
  const Geo& g = a.geo;
  const bool gaussian_line = g.spectrum == MTN_SPECTRUM_GAUSSIAN;
  const double sgn = g.edges_increasing ? 1.0 : -1.0;

  if (tid == 0) {
    mbar_init(&sm.bar[0], 1);
    mbar_init(&sm.bar[1], 1);
    mbar_fence_init();
  }
  __syncthreads();
  uint32_t phase = 0;  // bit k: parity to wait for on bar[k]
  unsigned long long n_upd = 0, n_w = 0, n_erf = 0;

  for (;;) {
    if (tid == 0) sm.item = atomicAdd(a.counter, 1u);
    __syncthreads();
    const uint32_t item_idx = sm.item;
    if (item_idx >= *a.n_items) break;
    const Item it = a.items[item_idx];
    const int cb = it.brick % g.ncb;
    const int tile = it.brick / g.ncb;
    const int ty = tile % g.nty, tx = tile / g.nty;
    const int x0 = g.x_lo + tx * TILE_X, y0 = ty * TILE_Y, c0 = cb * CB;
    const int nch = min(CB, g.C - c0);
    // this thread's pixel and the warp's sub-block bounds (full-cube pixel coordinates)
    const int gx = x0 + sx + lx, gy = y0 + sy + ly;
    const double wx0 = (double)(x0 + sx), wx1 = (double)(x0 + sx + SUB_X - 1);
    const double wy0 = (double)(y0 + sy), wy1 = (double)(y0 + sy + SUB_Y - 1);
    const double fx = (double)gx, fy = (double)gy;

    if (tid <= CB) sm.edge[tid] = a.edges[min(c0 + tid, g.C)];
    if (tid < CB)
      sm.inv_dv[tid] = tid < nch ? 1.0 / fabs(a.edges[c0 + tid + 1] - a.edges[c0 + tid]) : 0.0;

    double acc[CB];
#pragma unroll
    for (int c = 0; c < CB; ++c) acc[c] = 0.0;

    const uint32_t n_part = it.end - it.begin;
    const uint32_t n_batch = (n_part + PBATCH - 1) / PBATCH;

    auto issue = [&](uint32_t b) {
      const uint32_t buf = b & 1u;
      const uint32_t nb = min((uint32_t)PBATCH, n_part - b * PBATCH);
      if (tid < (int)nb) {
        const uint32_t ridx = (uint32_t)a.pairs[it.begin + b * PBATCH + tid];
        bulk_g2s(&sm.rec[buf][tid], a.records + ridx, REC_BYTES, &sm.bar[buf]);
      }
      if (tid == 0) mbar_arrive_expect_tx(&sm.bar[buf], nb * REC_BYTES);
    };

    issue(0);
    for (uint32_t b = 0; b < n_batch; ++b) {
      const uint32_t buf = b & 1u;
      const int nb = (int)min((uint32_t)PBATCH, n_part - b * PBATCH);
      if (b + 1 < n_batch) issue(b + 1);
      if (tid < PBATCH) sm.gmask[tid] = 0;
      mbar_wait(&sm.bar[buf], (phase >> buf) & 1u);
      phase ^= 1u << buf;
      __syncthreads();  // gmask zeroed, edge table visible, records landed for everyone

      // ---- spectra of the batch, once per CTA ----------------------------------------
      if (gaussian_line) {
        for (int idx = tid; idx < nb * (CB + 1); idx += PROJ_THREADS) {
          const int p = idx / (CB + 1), e = idx - p * (CB + 1);
          const Record& r = sm.rec[buf][p];
          sm.E[p][e] = edge_erf(sm.edge[e], r.v, r.inv_s);
          if (COUNT) n_erf += fabs((sm.edge[e] - r.v) * r.inv_s) < ERF_SAT;
        }
        __syncthreads();
        for (int idx = tid; idx < nb * CB; idx += PROJ_THREADS) {
          const int p = idx / CB, c = idx % CB;
          // 0.5*[erf(hi) - erf(lo)] * A / dv / 2.36e5; the 0.5 lives in amp
          const double s = sgn * (sm.E[p][c + 1] - sm.E[p][c]) * (sm.rec[buf][p].amp * sm.inv_dv[c]);
          sm.S[p][c] = s;
          if (s != 0.0) atomicOr(&sm.gmask[p], 1u << (c >> 3));
        }
      } else {
        for (int idx = tid; idx < nb * CB; idx += PROJ_THREADS) {
          const int p = idx / CB, c = idx % CB;
          const double e0 = sm.edge[c], e1 = sm.edge[c + 1];
          const double f = c < nch ? dirac_channel(fmin(e0, e1), fmax(e0, e1), sm.rec[buf][p].v) : 0.0;
          const double s = f * (sm.rec[buf][p].amp * sm.inv_dv[c]);
          sm.S[p][c] = s;
          if (s != 0.0) atomicOr(&sm.gmask[p], 1u << (c >> 3));
        }
      }
      __syncthreads();

      // ---- weights + rank-1 accumulate, per warp ---------------------------------------
      for (int p = 0; p < nb; ++p) {
        const uint32_t gm = sm.gmask[p];
        if (gm == 0) continue;
        const Record& r = sm.rec[buf][p];
        const double ppx = r.px, ppy = r.py, rr = (double)r.r;
        // candidate box of martini.py:272-274 vs the warp's sub-block (warp-uniform)
        if (!(fabs(__dsub_rn(wx0, ppx)) <= rr || fabs(__dsub_rn(wx1, ppx)) <= rr ||
              (wx0 < ppx && ppx < wx1)))
          continue;
        if (!(fabs(__dsub_rn(wy0, ppy)) <= rr || fabs(__dsub_rn(wy1, ppy)) <= rr ||
              (wy0 < ppy && ppy < wy1)))
          continue;
        double w = 0.0;
        if (fabs(__dsub_rn(fx, ppx)) <= rr && fabs(__dsub_rn(fy, ppy)) <= rr) {
          const int kid = r.kid;
          // dij = pixcoords - ij (martini.py:276)
          w = kernel_weight(a.table.kind[kid], __dsub_rn(ppx, fx), __dsub_rn(ppy, fy), r.h,
                            r.inv_h2, a.table.truncate[kid], a.table.norm[kid]);
        }
        if (COUNT && fabs(__dsub_rn(fx, ppx)) <= rr && fabs(__dsub_rn(fy, ppy)) <= rr) ++n_w;
        if (!__any_sync(0xffffffffu, w != 0.0)) continue;
        const double* Sp = sm.S[p];
        if (COUNT && w != 0.0) {
          for (int c = 0; c < CB; ++c) n_upd += Sp[c] != 0.0;
        }
#pragma unroll
        for (int gq = 0; gq < CB / 8; ++gq) {
          if (gm & (1u << gq)) {
#pragma unroll
            for (int c = 0; c < 8; c += 2) {
              const double2 s2 = *reinterpret_cast<const double2*>(Sp + gq * 8 + c);
              acc[gq * 8 + c] = fma(w, s2.x, acc[gq * 8 + c]);
              acc[gq * 8 + c + 1] = fma(w, s2.y, acc[gq * 8 + c + 1]);
            }
          }
        }
      }
      __syncthreads();  // S, gmask and rec[buf] are free again
    }

    // ---- one store per voxel -----------------------------------------------------------
    if (it.slot >= 0) {
      double* dst = a.partials + ((size_t)it.slot * PROJ_THREADS + tid) * CB;
#pragma unroll
      for (int c = 0; c < CB; c += 2)
        *reinterpret_cast<double2*>(dst + c) = make_double2(acc[c], acc[c + 1]);
    } else if (gx < g.x_hi && gy < g.ny) {
      double* dst = a.slab + ((size_t)(gx - g.x_lo) * g.ny + gy) * g.C + c0;
      store_pixel(dst, acc, nch, a.px_area, !a.zeroed, (g.C & 1) == 0);
    }
  }
  if (COUNT) {
    atomicAdd(a.exec_counts + 0, n_upd);
    atomicAdd(a.exec_counts + 1, n_w);
    atomicAdd(a.exec_counts + 2, n_erf);
  }
}

// Sum the partial bricks of a multi-chunk brick in chunk order, then store.
__global__ void __launch_bounds__(PROJ_THREADS) reduce_partials_kernel(
    Geo g, const MultiBrick* __restrict__ multis, const uint32_t* __restrict__ n_multi,
    const double* __restrict__ partials, double* __restrict__ slab, double px_area, int zeroed) {
  if (blockIdx.x >= *n_multi) return;
  const MultiBrick m = multis[blockIdx.x];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int sx = (warp >> 1) * SUB_X, sy = (warp & 1) * SUB_Y;
  const int cb = m.brick % g.ncb, tile = m.brick / g.ncb;
  const int gx = g.x_lo + (tile / g.nty) * TILE_X + sx + (lane >> 3);
  const int gy = (tile % g.nty) * TILE_Y + sy + (lane & 7);
  const int c0 = cb * CB, nch = min(CB, g.C - c0);
  double acc[CB];
#pragma unroll
  for (int c = 0; c < CB; ++c) acc[c] = 0.0;
  for (uint32_t j = 0; j < m.n; ++j) {
    const double* src = partials + ((size_t)(m.slot0 + j) * PROJ_THREADS + tid) * CB;
#pragma unroll
    for (int c = 0; c < CB; c += 2) {
      const double2 v = *reinterpret_cast<const double2*>(src + c);
      acc[c] += v.x;
      acc[c + 1] += v.y;
    }
  }
  if (gx < g.x_hi && gy < g.ny) {
    double* dst = slab + ((size_t)(gx - g.x_lo) * g.ny + gy) * g.C + c0;
    store_pixel(dst, acc, nch, px_area, !zeroed, (g.C & 1) == 0);
  }
}

// Accumulate mode only: voxels of bricks no particle reaches still get in / px_area.
__global__ void __launch_bounds__(PROJ_THREADS) empty_brick_kernel(
    Geo g, const uint32_t* __restrict__ brick_count, double* __restrict__ slab, double px_area) {
  const uint32_t brick = blockIdx.x;
  if (brick_count[brick] != 0) return;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int sx = (warp >> 1) * SUB_X, sy = (warp & 1) * SUB_Y;
  const int cb = brick % g.ncb, tile = brick / g.ncb;
  const int gx = g.x_lo + (tile / g.nty) * TILE_X + sx + (lane >> 3);
  const int gy = (tile % g.nty) * TILE_Y + sy + (lane & 7);
  const int c0 = cb * CB, nch = min(CB, g.C - c0);
  if (gx >= g.x_hi || gy >= g.ny) return;
  double* dst = slab + ((size_t)(gx - g.x_lo) * g.ny + gy) * g.C + c0;
  for (int c = 0; c < nch; ++c) dst[c] = dst[c] / px_area;
}

}  // namespace mtn
