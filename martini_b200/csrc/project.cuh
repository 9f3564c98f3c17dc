// The brick projection kernel and its small companions (work items, partial reduce).
//
// One CTA works on one brick -- 8 x 8 pixels x 64 channels -- at a time and keeps the brick's
// sums in registers.  A thread owns ONE PIXEL and NCH consecutive channels of it: warp
// (ph, cw) covers pixel half ph (rows 4 ph .. 4 ph + 3 of the tile, lane = (row, column)) and
// channels [cw NCH, (cw + 1) NCH) of the brick.  This is the layout the two costs of the sum
//
//     acc[pixel][c] += W_p(pixel) * S_p(c)
//
// want: the kernel integral W_p(pixel) is evaluated by the very lane that accumulates it
// (particle warp-uniform, one lane per pixel: no enumeration of (particle, pixel) items, no
// per-lane kernel-kind divergence, no W broadcast loads), and the spectrum S_p(c), which
// every pixel of the brick shares, is a warp-uniform shared-memory broadcast: two 16-byte
// loads feed four FMAs of every lane.
//
// The particle records of the brick are gathered into shared memory with cp.async.bulk (one
// 80-byte bulk copy per record, completion on an mbarrier, double buffered).  Per batch of 32
// staged particles:
//
//   S   the line spectrum of every particle over the brick's 64 channels, ONCE per
//       (particle, brick): lanes = channel edges, erf from the table, the difference of
//       adjacent edges by shuffle, result (exact zeros outside the live window) in shared
//       memory; the warps share the particles out.
//   W   the kernel integrals of every particle over the brick's 64 pixels, once per
//       (particle, pixel): the channel warps of a pixel half share the particles out and
//       leave W[ph][p][lane] plus the ballot of its non-zero lanes in shared memory.
//   C   each warp walks the particles that reach its pixels AND its channels (one ballot):
//       w = W[ph][p][lane], then for every live group of four channels two broadcast loads of
//       S and four FMAs.
//
// Two block barriers per batch.  No atomics on the data path; every voxel is stored exactly
// once (16-byte vector stores, 32 x NCH x 8 contiguous bytes per thread).
#pragma once

#include "common.cuh"
#include "kernel_integrals.cuh"
#include "plan.cuh"

namespace mtn {

// ----------------------------------------------------------------------------- PTX helpers
#ifndef MTN_HOST_EMU  // (the CPU test suite's emulator, tests/emu/cuda_emu.h, supplies its own)
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
// global -> shared bulk copy (TMA engine, SASS UBLKCP), completion bytes on `bar`
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes,
                                         uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
          "r"(smem_u32(dst_smem)),
      "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
#endif  // MTN_HOST_EMU
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  // bounded spin: a lost transaction traps instead of hanging the GPU
#pragma unroll 1
  for (uint32_t it = 0; it < (1u << 26); ++it)
    if (mbar_try_wait(bar, parity)) return;
  __trap();
}

// ------------------------------------------------------------------------------ work items
// Bricks are found in the sorted pair array by their key boundaries: start[] gets the first
// index, end_or_count[] the one-past-last index (rewritten to a count by item_count_kernel).
__global__ void __launch_bounds__(256) brick_bounds_kernel(const uint64_t* __restrict__ pairs,
                                                           int64_t n, uint32_t* __restrict__ start,
                                                           uint32_t* __restrict__ end_or_count) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t k = (uint32_t)(pairs[i] >> 32);
  if (i == 0 || (uint32_t)(pairs[i - 1] >> 32) != k) start[k] = (uint32_t)i;
  if (i == n - 1 || (uint32_t)(pairs[i + 1] >> 32) != k) end_or_count[k] = (uint32_t)i + 1u;
}

// counts[b] = chunks of brick b; multi[b] = same if > 1 else 0; ismulti[b] = 0/1.
__global__ void __launch_bounds__(256) item_count_kernel(uint32_t* __restrict__ brick_count,
                                                         const uint32_t* __restrict__ brick_start,
                                                         int n_bricks, uint32_t chunk,
                                                         uint32_t* __restrict__ counts,
                                                         uint32_t* __restrict__ multi,
                                                         uint32_t* __restrict__ ismulti) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= n_bricks) return;
  const uint32_t end = brick_count[b];
  const uint32_t c = end ? end - brick_start[b] : 0u;
  brick_count[b] = c;
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
  double* partials;  // [slot][TILE_PIX][CB]
  double px_area;    // px_size_arcsec^2
  int zeroed;        // MTN_CUBE_ZEROED
  unsigned long long* exec_counts;  // COUNT instantiation only: [updates, weights, erfs]
};

struct ProjSmem {
  Record rec[2][PBATCH];
  double S[PBATCH][CB];            // line spectra over the brick's channels, zero outside the window
  double W[N_PH][PBATCH][32];      // kernel integrals, [pixel half][particle][lane = pixel]
  double inv_dv[CB];               // (16-byte aligned: read as double2)
  double edge[CB + 1];
  uint32_t nz[N_PH][PBATCH];       // ballot of the lanes with W != 0
  uint32_t quads[PBATCH];          // bit q: channels [4q, 4q+4) of the brick hold a live channel
  uint64_t bar[2];
  uint32_t item;
};

// One thread stores two adjacent channels of one pixel: out = (in + acc) / px_area
// (martini.py:338, 364-366).  `nvalid` = how many of the two channels exist.
__device__ __forceinline__ void store2(double* __restrict__ dst, double a0, double a1, int nvalid,
                                       double px_area, bool add_in, bool vec_ok) {
  if (nvalid == 2 && vec_ok) {
    double2 o;
    if (add_in) {
      const double2 i2 = *reinterpret_cast<const double2*>(dst);
      o.x = (i2.x + a0) / px_area;
      o.y = (i2.y + a1) / px_area;
    } else {
      o.x = a0 / px_area;
      o.y = a1 / px_area;
    }
    *reinterpret_cast<double2*>(dst) = o;
  } else {
    if (nvalid >= 1) dst[0] = ((add_in ? dst[0] : 0.0) + a0) / px_area;
    if (nvalid >= 2) dst[1] = ((add_in ? dst[1] : 0.0) + a1) / px_area;
  }
}

// COUNT = true is a diagnostic instantiation that additionally tallies the executed
// algorithmic work (non-zero weight x non-zero spectrum terms, kernel integrals, edge erfs);
// it is never the timed kernel.  KIND >= 0: entry 0 of the kernel table -- the SPH kernel proper
// of the reference's adaptive kernels; the other entries are its small-h fallbacks -- is that
// tabulated kind, so the weight evaluation is the bare table look-up with compile-time zone
// bounds and only particles on another entry take the closed forms; KIND = -1: general case.
template <bool COUNT, int KIND>
__global__ void __launch_bounds__(PROJ_THREADS, PROJ_CTAS_PER_SM) project_kernel(const ProjArgs a) {
  MTN_DYN_SMEM(unsigned char, smem_raw);
  ProjSmem& sm = *reinterpret_cast<ProjSmem*>(smem_raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int ph = warp / N_CW, cw = warp % N_CW;  // this warp's pixel half and channel group
  const int tpx = ph * (TILE_X / N_PH) + (lane >> 3), tpy = lane & 7;  // this thread's tile pixel
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
    // channels [c0, c0 + CB) of the cube; c0 may be negative (block 0 of a phased tile)
    const int x0 = g.x_lo + tx * TILE_X, y0 = ty * TILE_Y, c0 = g.phase[tile] + (cb - 1) * CB;
    const int clo = max(0, -c0), nch = min(CB, g.C - c0);  // valid brick channels: [clo, nch)
    // pixels of the tile that exist in the slab / cube
    const int x_last = min(x0 + TILE_X, g.x_hi) - 1, y_last = min(y0 + TILE_Y, g.ny) - 1;
    // rows of this warp's pixel half
    const int hx0 = x0 + ph * (TILE_X / N_PH), hx1 = min(hx0 + TILE_X / N_PH - 1, x_last);
    const int gx = x0 + tpx, gy = y0 + tpy;
    const double gxd = (double)gx, gyd = (double)gy;

    for (int e = tid; e <= CB; e += PROJ_THREADS) sm.edge[e] = a.edges[min(max(c0 + e, 0), g.C)];
    for (int c = tid; c < CB; c += PROJ_THREADS)
      sm.inv_dv[c] = (c >= clo && c < nch) ? 1.0 / fabs(a.edges[c0 + c + 1] - a.edges[c0 + c]) : 0.0;

    double acc[NCH];
#pragma unroll
    for (int k = 0; k < NCH; ++k) acc[k] = 0.0;

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
    if (n_batch > 1) issue(1);
    __syncthreads();  // edge table visible before the first spectrum step reads it

    for (uint32_t b = 0; b < n_batch; ++b) {
      const uint32_t buf = b & 1u;
      const int nb = (int)min((uint32_t)PBATCH, n_part - b * PBATCH);
      // every thread observes the completion itself (visibility of rec[buf])
      mbar_wait(&sm.bar[buf], (phase >> buf) & 1u);
      phase ^= 1u << buf;
      const Record* rec = sm.rec[buf];

      // ---- S: line spectra, once per (particle, brick).  Warp w takes particles w, w + NW, ...
      // The record carries the particle's live channel window (plan.cuh: channel_window, the
      // exact predicates); its share of the brick is [cs, ce).  Lane l evaluates the edges
      // cs + l and cs + 32 + l (two independent erf chains), the rare 65th edge is lane 0's.
      for (int p = warp; p < nb; p += PROJ_WARPS) {
        const Record& r = rec[p];
        int cs = max((int)r.c_first - c0, clo), ce = min((int)r.c_last + 1 - c0, nch);
        if (cs >= ce) {
          if (lane == 0) sm.quads[p] = 0u;
          continue;
        }
        const int c1 = cs + lane, c2 = cs + 32 + lane;  // this lane's channels (= their lower edges)
        double s1 = 0.0, s2 = 0.0;
        if (gaussian_line) {
          const int ne = ce - cs + 1;  // edges cs .. ce
          const double v = r.v, sc = sgn * r.inv_s;
          // g orientation (sign folded in): S[c] = E[c+1] - E[c] >= 0; saturated edges at the
          // ends of the window come out as exactly -1 / +1
          const double t1 = (sm.edge[min(c1, CB)] - v) * sc;
          double E1 = erf_tab(t1), E2 = 0.0, E3 = 0.0;
          if (COUNT) n_erf += (lane < ne) && fabs(t1) < ERF_SAT;
          if (ne > 32) {
            const double t2 = (sm.edge[min(c2, CB)] - v) * sc;
            E2 = erf_tab(t2);
            if (COUNT) n_erf += (lane + 32 < ne) && fabs(t2) < ERF_SAT;
            if (ne > 64) {  // the window covers the whole brick: edge 64 (cs == 0)
              const double t3 = (sm.edge[CB] - v) * sc;
              E3 = erf_tab(t3);
              if (COUNT) n_erf += lane == 0 && fabs(t3) < ERF_SAT;
            }
          }
          // upper edge of each channel: the next lane's value; lane 31 wraps into the next chain
          double U1 = __shfl_down_sync(0xffffffffu, E1, 1);
          double U2 = __shfl_down_sync(0xffffffffu, E2, 1);
          const double E2_0 = __shfl_sync(0xffffffffu, E2, 0);
          if (lane == 31) {
            U1 = E2_0;
            U2 = E3;
          }
          const double amp = r.amp;
          if (c1 < ce) s1 = (U1 - E1) * (amp * sm.inv_dv[c1]);
          if (c2 < ce) s2 = (U2 - E2) * (amp * sm.inv_dv[c2]);
        } else {  // Dirac line: the live channels are exactly those with lo <= v <= hi
          const double amp = r.amp;
          if (c1 < ce) s1 = amp * sm.inv_dv[c1];
          if (c2 < ce) s2 = amp * sm.inv_dv[c2];
        }
        // rotate into place: lane l holds channels cs + l and cs + 32 + l; the row is written
        // whole (zeros outside the window), so phase C may read any group of four it visits
        {
          double* row = sm.S[p];
          // channels below cs and from ce on: zero.  Lane l clears l (< cs) and the tail.
          if (lane < cs) row[lane] = 0.0;
          if (c1 < CB) row[c1] = s1;
          if (c2 < CB) row[c2] = s2;
          if (cs > 32 && lane + 32 < cs) row[lane + 32] = 0.0;
        }
        if (lane == 0) {
          const uint32_t hi = ((ce + 3) >> 2) >= 32 ? 0xffffffffu : ((1u << ((ce + 3) >> 2)) - 1u);
          sm.quads[p] = hi & ~((1u << (cs >> 2)) - 1u);
        }
      }

      // ---- W: kernel integrals, once per (particle, pixel).  The channel warps of a pixel half
      // share the particles out; two particles per step (two independent evaluation chains).
      for (int p0 = cw; p0 < nb; p0 += 2 * N_CW) {
        double wv[2];
        bool live[2];
        int kidv[2];
        double dxv[2], dyv[2], R2v[2], ih2v[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const int p = p0 + u * N_CW;
          live[u] = false;
          kidv[u] = 0;
          dxv[u] = dyv[u] = R2v[u] = ih2v[u] = 0.0;
          if (p < nb) {
            const Record& r = rec[p];
            // the candidate box of martini.py:272-274 (computed once by the plan kernels with
            // the exact predicate), cut to this pixel half; live channels in the brick
            const bool any = max(r.i0, hx0) <= min(r.i1, hx1) && max(r.j0, y0) <= min(r.j1, y_last) &&
                             max((int)r.c_first - c0, clo) < min((int)r.c_last + 1 - c0, nch);
            if (any) {
              live[u] = true;
              kidv[u] = r.kid;
              // dij = pixcoords - ij (martini.py:276)
              dxv[u] = __dsub_rn(r.px, gxd);
              dyv[u] = __dsub_rn(r.py, gyd);
              ih2v[u] = r.inv_h2;
              R2v[u] = sq_dist(dxv[u], dyv[u]) * ih2v[u];
            } else if (lane == 0) {
              sm.nz[ph][p] = 0u;
            }
          }
        }
        if (!live[0] && !live[1]) continue;
#pragma unroll
        for (int u = 0; u < 2; ++u) {  // straight-line, both chains in flight together
          const int kind = (KIND >= 0 && kidv[u] == 0) ? KIND : a.table.kind[kidv[u]];
          wv[u] = wtab_eval(KIND >= 0 ? KIND : (wtab_has(kind) ? kind : MTN_KERNEL_WENDLANDC2), R2v[u]) * ih2v[u];
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          if (!live[u]) continue;  // (warp-uniform)
          const int p = p0 + u * N_CW;
          const Record& r = rec[p];
          const int kind = (KIND >= 0 && kidv[u] == 0) ? KIND : a.table.kind[kidv[u]];
          double w = wv[u];
          // closed form: kernels without a table; with KIND, every entry but the first
          if (KIND >= 0 ? kidv[u] != 0 : !wtab_has(kind))
            w = kernel_weight_closed(kind, dxv[u], dyv[u], r.h, r.inv_h2, a.table.truncate[kidv[u]],
                                     a.table.norm[kidv[u]]);
          const bool inbox = gx >= r.i0 && gx <= r.i1 && gy >= r.j0 && gy <= r.j1;
          w = inbox ? w : 0.0;
          sm.W[ph][p][lane] = w;
          const uint32_t m = __ballot_sync(0xffffffffu, w != 0.0);
          if (lane == 0) sm.nz[ph][p] = m;
          if (COUNT) n_w += inbox && cw == (p % N_CW);
        }
      }
      __syncthreads();

      // ---- C: acc[c] += W * S for the particles that reach this warp's pixels and channels ---
      {
        constexpr uint32_t QMASK = (NQ >= 32) ? 0xffffffffu : ((1u << NQ) - 1u);
        bool mine = false;
        if (lane < nb) mine = sm.nz[ph][lane] != 0u && ((sm.quads[lane] >> (cw * NQ)) & QMASK) != 0u;
        uint32_t act = __ballot_sync(0xffffffffu, mine);
        if (act) {
          int p = __ffs(act) - 1;
          act &= act - 1;
          double w = sm.W[ph][p][lane];
          uint32_t qm = (sm.quads[p] >> (cw * NQ)) & QMASK;
          for (;;) {
            const bool more = act != 0;
            const int pn = more ? __ffs(act) - 1 : p;
            act &= act - 1;
            // the next particle's loads go out ahead of this one's FMAs
            const double wn = sm.W[ph][pn][lane];
            const uint32_t qn = (sm.quads[pn] >> (cw * NQ)) & QMASK;
            const double* Sp = &sm.S[p][cw * NCH];
#pragma unroll
            for (int q = 0; q < NQ; ++q) {
              if (qm & (1u << q)) {
                const double2 sa = *reinterpret_cast<const double2*>(Sp + 4 * q);
                const double2 sb = *reinterpret_cast<const double2*>(Sp + 4 * q + 2);
                acc[4 * q + 0] = fma(w, sa.x, acc[4 * q + 0]);
                acc[4 * q + 1] = fma(w, sa.y, acc[4 * q + 1]);
                acc[4 * q + 2] = fma(w, sb.x, acc[4 * q + 2]);
                acc[4 * q + 3] = fma(w, sb.y, acc[4 * q + 3]);
                if (COUNT)
                  n_upd += w != 0.0 ? (sa.x != 0.0) + (sa.y != 0.0) + (sb.x != 0.0) + (sb.y != 0.0) : 0;
              }
            }
            if (!more) break;
            p = pn;
            w = wn;
            qm = qn;
          }
        }
      }
      __syncthreads();  // W, S, masks and rec[buf] are free again
      if (b + 2 < n_batch) issue(b + 2);
    }

    // ---- one store per voxel: NCH consecutive channels of this thread's pixel ---------------
    const int cl0 = cw * NCH;  // first brick channel of this thread
    if (it.slot >= 0) {
      double* dst = a.partials + ((size_t)it.slot * TILE_PIX + tpx * TILE_Y + tpy) * CB + cl0;
#pragma unroll
      for (int k = 0; k < NCH; k += 2)
        *reinterpret_cast<double2*>(dst + k) = make_double2(acc[k], acc[k + 1]);
    } else if (gx <= x_last && gy <= y_last) {
      double* dst = a.slab + ((size_t)(gx - g.x_lo) * g.ny + gy) * g.C + c0 + cl0;
      const bool vec_ok = (g.C & 1) == 0;
#pragma unroll
      for (int k = 0; k < NCH; k += 2) {
        const int cl = cl0 + k;
        const int nvalid = cl < clo ? 0 : max(0, min(2, nch - cl));  // (clo, c0 are even)
        if (nvalid > 0) store2(dst + k, acc[k], acc[k + 1], nvalid, a.px_area, !a.zeroed, vec_ok);
      }
    }
  }
  if (COUNT) {
    atomicAdd(a.exec_counts + 0, n_upd);
    atomicAdd(a.exec_counts + 1, n_w);
    atomicAdd(a.exec_counts + 2, n_erf);
  }
}

// Sum the partial bricks of a multi-chunk brick in chunk order, then store.  One block per
// (multi-chunk brick, 1/REDUCE_PARTS of its voxels): thread = one channel pair of one pixel,
// so the hottest brick (most partials) is spread over REDUCE_PARTS blocks.
constexpr int REDUCE_THREADS = 128;
constexpr int REDUCE_PARTS = TILE_PIX * CB / 2 / REDUCE_THREADS;
__global__ void __launch_bounds__(REDUCE_THREADS) reduce_partials_kernel(
    Geo g, const MultiBrick* __restrict__ multis, const uint32_t* __restrict__ n_multi,
    const double* __restrict__ partials, double* __restrict__ slab, double px_area, int zeroed) {
  if (blockIdx.x >= *n_multi) return;
  const MultiBrick m = multis[blockIdx.x];
  const int cb = m.brick % g.ncb, tile = m.brick / g.ncb;
  const int x0 = g.x_lo + (tile / g.nty) * TILE_X, y0 = (tile % g.nty) * TILE_Y;
  const int c0 = g.phase[tile] + (cb - 1) * CB;
  const int e = blockIdx.y * REDUCE_THREADS + threadIdx.x;  // double2 index within the brick
  const int pix = e / (CB / 2), cl = 2 * (e % (CB / 2));
  const int nvalid = c0 + cl < 0 ? 0 : max(0, min(2, min(CB, g.C - c0) - cl));
  const double* src = partials + ((size_t)m.slot0 * TILE_PIX + pix) * CB + cl;
  double a0 = 0.0, a1 = 0.0;
#pragma unroll 8
  for (uint32_t k = 0; k < m.n; ++k) {  // chunk order: deterministic
    const double2 v = *reinterpret_cast<const double2*>(src + (size_t)k * TILE_PIX * CB);
    a0 += v.x;
    a1 += v.y;
  }
  const int gx = x0 + pix / TILE_Y, gy = y0 + pix % TILE_Y;
  if (nvalid > 0 && gx < g.x_hi && gy < g.ny) {
    double* dst = slab + ((size_t)(gx - g.x_lo) * g.ny + gy) * g.C + c0 + cl;
    store2(dst, a0, a1, nvalid, px_area, !zeroed, (g.C & 1) == 0);
  }
}

// Accumulate mode only: voxels of bricks no particle reaches still get in / px_area.
__global__ void __launch_bounds__(REDUCE_THREADS) empty_brick_kernel(
    Geo g, const uint32_t* __restrict__ brick_count, double* __restrict__ slab, double px_area) {
  const uint32_t brick = blockIdx.x;
  if (brick_count[brick] != 0) return;
  const int cb = brick % g.ncb, tile = brick / g.ncb;
  const int x0 = g.x_lo + (tile / g.nty) * TILE_X, y0 = (tile % g.nty) * TILE_Y;
  const int c0 = g.phase[tile] + (cb - 1) * CB;
  for (int e = threadIdx.x; e < TILE_PIX * CB / 2; e += REDUCE_THREADS) {
    const int pix = e / (CB / 2), cl = 2 * (e % (CB / 2));
    const int nvalid = c0 + cl < 0 ? 0 : max(0, min(2, min(CB, g.C - c0) - cl));
    const int gx = x0 + pix / TILE_Y, gy = y0 + pix % TILE_Y;
    if (nvalid > 0 && gx < g.x_hi && gy < g.ny) {
      double* dst = slab + ((size_t)(gx - g.x_lo) * g.ny + gy) * g.C + c0 + cl;
      store2(dst, 0.0, 0.0, nvalid, px_area, true, (g.C & 1) == 0);
    }
  }
}

}  // namespace mtn
