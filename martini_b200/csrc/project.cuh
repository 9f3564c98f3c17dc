// The tile projection kernel and its small companions (work items, partial reduce).
//
// One CTA works on one brick -- TILE x TILE pixels x 64 channels -- at a time and keeps the
// brick's sums in registers: warp w owns a 4 x 4 pixel sub-block; lane l owns one pixel row of
// it (l / 8) and four channel pairs (2 (l % 8) + 16 h, + 1), i.e. a 4 pixel x 8 channel register
// tile = 32 float64 accumulators per thread.  With the default 8 x 8 tile a CTA is 4 warps and
// four CTAs share an SM, so one CTA's barrier waits are covered by the others.  The particle records of the brick are gathered
// into shared memory with cp.async.bulk (one 80-byte bulk copy per record, completion on an
// mbarrier, double buffered).  Per batch of 32 staged particles:
//
//   setup  one lane per particle: the record carries the particle's footprint (candidate box
//          of martini.py:272-274 in the slab, live channel window -- computed once by the
//          plan kernels with the exact predicates), so warps 0/1 clip it to the brick with
//          integer arithmetic and turn it into prefix sums of box areas and edge-run lengths,
//          for batch b+1 while batch b is in phase A;
//   A      every (particle, box pixel) pair and every (particle, live edge) pair is handed to
//          one thread through those prefix sums -- all lanes hold real work, two items per
//          lane so two dependency chains are in flight -- which evaluates the SPH-kernel
//          pixel integral (tabulated, tables.cuh; times the particle's amplitude) or the edge
//          erf ONCE into shared memory; the edges outside a particle's live window are filled
//          with the saturated values -1 / +1 they have in the reference, so that phase C needs
//          no channel predicate;
//   C      each warp builds, with one ballot, the list of particles that touch its sub-block
//          (zeroing the weights of its sub-block that lie outside the particle's box), then
//          walks it: two 16-byte loads bring the lane's four weights, four 16-byte and four
//          8-byte loads the three edge erfs of each of its channel pairs (all conflict-free by
//          layout: ~10 shared-memory wavefronts per 32 FMAs, against 22 for the 16 pixel x
//          2 channel tile of rounds 1-2), eight subtractions form E[c+1] - E[c], then
//          acc[pixel][channel] += (W amp) * (E[c+1] - E[c]); the factor 1 / dv of the channel
//          is applied once, at the store.
//
// No atomics on the data path; every voxel is stored exactly once, as 16-byte vector stores
// (per pixel and channel-pair index the eight channel groups of a warp write 128 contiguous
// bytes).
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

__device__ __forceinline__ uint32_t warp_incl_scan_u32(uint32_t x, int lane) {
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const uint32_t y = __shfl_up_sync(0xffffffffu, x, d);
    if (lane >= d) x += y;
  }
  return x;
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

// Shared-memory layouts of the per-batch weights and edge erfs, chosen for phase C's vector
// loads.  W: the 16 weights of a sub-block are contiguous (row of the sub-block major), so the
// four pixel rows a warp's lanes read are one 128-byte line; rows of W are 16-byte aligned.
// ES: the CB + 1 edge erfs of a particle in order.  Lane l of a warp owns pixel row l / 8 of the
// sub-block and the channel pairs (2g + 16h, 2g + 16h + 1), g = l % 8, h = 0..3: for every h the
// eight channel groups of a warp read one contiguous 128-byte line of edges (plus the pair's
// third edge, 16 bytes apart: conflict-free), and at the end store one contiguous 128-byte
// line of the cube per pixel and h -- whole 32-byte sectors.
constexpr int W_STRIDE = TILE_PIX + 2;
#ifndef MTN_TILE_IL
#define MTN_TILE_IL 1
#endif
#if MTN_TILE_IL
constexpr int ES_STRIDE = CB + 2;
__host__ __device__ constexpr int es_pos(int e) { return e; }
#else
// (experiment: lane owns eight ADJACENT channels; edge e sits at e + 2 (e / 8) so that the eight
// channel groups start 80 bytes apart -- fewer loads per visit, but 16-byte stores 64 bytes apart)
constexpr int ES_STRIDE = CB + 2 * (CB / 8) + 2;
__host__ __device__ constexpr int es_pos(int e) { return e + 2 * (e >> 3); }
#endif
static_assert(SUB_X == 4 && SUB_Y == 4 && CB == 64, "phase C's register tile is 4 pixels x 8 channels");
__host__ __device__ constexpr int w_index(int tpx, int tpy) {
  return (((tpx >> 2) * SUBS_Y + (tpy >> 2)) << 4) | ((tpx & 3) << 2) | (tpy & 3);
}

// What the per-batch set-up leaves for the evaluation and accumulation phases.
struct SetupBuf {
  uint32_t wprefix[PBATCH + 1];  // exclusive prefix of box areas
  uint32_t eprefix[PBATCH + 1];  // exclusive prefix of edge-run lengths
  float rny[PBATCH];             // 1 / (box height)
  uint8_t wowner[PBATCH];        // particles with a non-empty box, in order
  uint8_t eowner[PBATCH];        // particles with a non-empty edge run, in order
  uint8_t box[PBATCH][4];        // tile-local box: x0, nx, y0, ny
  uint8_t erun[PBATCH][2];       // first edge to evaluate, number of edges
  uint8_t chan[PBATCH][2];       // live channels of the brick: [cs, ce)
  uint8_t hlive[PBATCH];         // 1: the particle reaches a pixel and a live channel of the brick
};

struct ProjSmem {
  Record rec[2][PBATCH];
  double W[PBATCH][W_STRIDE];   // kernel integrals x amplitude at w_index(pixel): zero outside the
                                // particle's box (every warp clears its sub-block's share after use)
  double ES[PBATCH][ES_STRIDE]; // the CB + 1 edge erfs of every live particle
  double erf_table[ERFC_DOUBLES];  // copy of the compact erf table (tables.cuh): the edge erfs read it here
                                   // instead of the degree-9 table through L1, whose five loads per
                                   // evaluation, each lane in its own row, kept the L1 tag stage busy for
                                   // most of the kernel (22 tag look-ups per load instruction)
  double inv_dv[CB];            // (16-byte aligned: read as double2)
  double edge[CB + 1];
  SetupBuf sb[2];  // (two: batch b+1 is set up while batch b is evaluated)
  uint64_t bar[2];
  uint32_t item;
  uint32_t ctr[2][2];  // [batch parity][kernel integrals / edge erfs]: next chunk of phase-A items
};

// tile pixel (x, y) of pixel j of sub-block s
__device__ __forceinline__ int sub_x(int s, int j) { return (s / SUBS_Y) * SUB_X + j / SUB_Y; }
__device__ __forceinline__ int sub_y(int s, int j) { return (s % SUBS_Y) * SUB_Y + j % SUB_Y; }

// One thread stores its two channels of one pixel: out = (in + acc) / px_area
// (martini.py:338, 364-366).  `nvalid` = how many of the lane's two channels exist.
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

// Cooperative owner lookup for 32 consecutive enumerated items starting at q0: lane k knows
// where run k starts (`my_start`, `my_nonempty`); returns for this lane the ordinal (among
// non-empty runs) of the run that holds item q0 + lane.
__device__ __forceinline__ int owner_ordinal(uint32_t q0, uint32_t my_start, bool my_nonempty,
                                             int lane) {
  const uint32_t before = __popc(__ballot_sync(0xffffffffu, my_nonempty && my_start < q0));
  const uint32_t rel = my_start - q0;
  const uint32_t H = __reduce_or_sync(0xffffffffu, (my_nonempty && rel < 32u) ? (1u << rel) : 0u);
  return (int)(before + __popc(H & ((2u << lane) - 1u))) - 1;
}

// COUNT = true is a diagnostic instantiation that additionally tallies the executed
// algorithmic work (non-zero weight x non-zero spectrum terms, kernel integrals, edge erfs);
// it is never the timed kernel.  KIND >= 0: entry 0 of the kernel table -- the SPH kernel proper
// of the reference's adaptive kernels; the other entries are its small-h fallbacks -- is that
// tabulated kind, so the weight evaluation is the bare table look-up with compile-time zone
// bounds and only the lanes on another entry take the closed forms; KIND = -1: general case.
template <bool COUNT, int KIND>
__global__ void __launch_bounds__(PROJ_THREADS, PROJ_CTAS_PER_SM) project_kernel(const ProjArgs a) {
  MTN_DYN_SMEM(unsigned char, smem_raw);
  ProjSmem& sm = *reinterpret_cast<ProjSmem*>(smem_raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int sub = warp;  // this warp's sub-block of the tile
  const Geo& g = a.geo;
  const bool gaussian_line = g.spectrum == MTN_SPECTRUM_GAUSSIAN;
  const double sgn = g.edges_increasing ? 1.0 : -1.0;

  if (tid == 0) {
    mbar_init(&sm.bar[0], 1);
    mbar_init(&sm.bar[1], 1);
    mbar_fence_init();
  }
  for (int k = tid; k < PBATCH * W_STRIDE; k += PROJ_THREADS) (&sm.W[0][0])[k] = 0.0;
  for (int k = tid; k < ERFC_DOUBLES; k += PROJ_THREADS) sm.erf_table[k] = g_erf_table_compact[k];
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

    if (tid < 4) sm.ctr[tid >> 1][tid & 1] = 0u;
    for (int e = tid; e <= CB; e += PROJ_THREADS) sm.edge[e] = a.edges[min(max(c0 + e, 0), g.C)];
    for (int c = tid; c < CB; c += PROJ_THREADS)
      sm.inv_dv[c] = (c >= clo && c < nch) ? 1.0 / fabs(a.edges[c0 + c + 1] - a.edges[c0 + c]) : 0.0;

    double acc[4][8];  // [pixel of my sub-block row][channel of my group]
#pragma unroll
    for (int k = 0; k < 4; ++k)
#pragma unroll
      for (int q = 0; q < 8; ++q) acc[k][q] = 0.0;

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

    // ---- setup, lane = particle: the record carries the particle's footprint (candidate
    // box in the slab, live channel window; computed once by the plan kernels with the exact
    // predicates of martini.py:272-274), so the brick's share of it is integer clipping.
    // Warp 0 turns the box areas into the enumeration prefix / owner list of the kernel
    // integrals, warp 1 does the same for the edge runs; one barrier.
    auto setup_from_records = [&](uint32_t bb) {
      const int nbb = (int)min((uint32_t)PBATCH, n_part - bb * PBATCH);
      SetupBuf& S = sm.sb[bb & 1u];
      if (warp < 2) {
        int bx0 = 0, bnx = 0, by0 = 0, bny = 0, cs = 0, ce = 0;
        if (lane < nbb) {
          const Record& r = sm.rec[bb & 1u][lane];
          int xa = max(r.i0, x0), xb = min(r.i1, x_last);
          int ya = max(r.j0, y0), yb = min(r.j1, y_last);
          // the candidate box is wider than the kernel's round support by up to a pixel on every
          // side (sm_range = ceil(...)): columns / rows of it with |i - px| >= support hold exact
          // zeros (R >= support there; the margin of plan.cuh: tile_reached), W stays zero where
          // nothing is written, so they need not be enumerated
          const double sup = r.h * g.support[r.kid] * (1.0 + 1.0e-9);
          if (sup < 1.0e9) {
            xa = max(xa, (int)ceil(r.px - sup));
            xb = min(xb, (int)floor(r.px + sup));
            ya = max(ya, (int)ceil(r.py - sup));
            yb = min(yb, (int)floor(r.py + sup));
          }
          if (xa <= xb && ya <= yb) {
            bx0 = xa - x0;
            bnx = xb - xa + 1;
            by0 = ya - y0;
            bny = yb - ya + 1;
          }
          // live channels of the brick [cs, ce): the particle's window cut to the brick
          cs = max((int)r.c_first - c0, clo);
          ce = min((int)r.c_last + 1 - c0, nch);
          if (cs >= ce) cs = ce = 0;
        }
        const uint32_t area = (ce > cs) ? (uint32_t)(bnx * bny) : 0u;
        const uint32_t ne = (area && gaussian_line) ? (uint32_t)(ce - cs + 1) : 0u;
        const uint32_t lt = (1u << lane) - 1u;
        if (warp == 0) {
          S.box[lane][0] = (uint8_t)bx0;
          S.box[lane][1] = (uint8_t)bnx;
          S.box[lane][2] = (uint8_t)by0;
          S.box[lane][3] = (uint8_t)bny;
          S.rny[lane] = bny ? 1.0f / (float)bny : 0.0f;
          S.chan[lane][0] = (uint8_t)cs;
          S.chan[lane][1] = (uint8_t)ce;
          S.hlive[lane] = (uint8_t)(area ? 1u : 0u);
          const uint32_t wi = warp_incl_scan_u32(area, lane);
          S.wprefix[lane] = wi - area;
          if (lane == 31) S.wprefix[32] = wi;
          const uint32_t wm = __ballot_sync(0xffffffffu, area != 0);
          if (area != 0) S.wowner[__popc(wm & lt)] = (uint8_t)lane;
        }
        if (warp == 1) {
          S.erun[lane][0] = (uint8_t)cs;
          S.erun[lane][1] = (uint8_t)ne;
          const uint32_t ei = warp_incl_scan_u32(ne, lane);
          S.eprefix[lane] = ei - ne;
          if (lane == 31) S.eprefix[32] = ei;
          const uint32_t em = __ballot_sync(0xffffffffu, ne != 0);
          if (ne != 0) S.eowner[__popc(em & lt)] = (uint8_t)lane;
        }
      }
    };

    // Two barriers per batch: batch b+1 is set up (by warps 0 and 1, from the records that
    // landed while batch b-1 was processed) during batch b's evaluation phase, into the other
    // set-up buffer; the bulk copies of batch b+2 are issued once batch b has released rec[buf].
    issue(0);
    if (n_batch > 1) issue(1);
    __syncthreads();  // edge table visible before the first setup step reads it
    mbar_wait(&sm.bar[0], phase & 1u);
    phase ^= 1u;
    setup_from_records(0);
    __syncthreads();
    for (uint32_t b = 0; b < n_batch; ++b) {
      const uint32_t buf = b & 1u;
      const int nb = (int)min((uint32_t)PBATCH, n_part - b * PBATCH);
      SetupBuf& S = sm.sb[buf];
      if (b > 0) {  // every thread observes the completion itself (visibility of rec[buf])
        mbar_wait(&sm.bar[buf], (phase >> buf) & 1u);
        phase ^= 1u << buf;
      }
      if (warp < 2 && b + 1 < n_batch) {
        mbar_wait(&sm.bar[buf ^ 1u], (phase >> (buf ^ 1u)) & 1u);  // (parity toggled next iteration)
        setup_from_records(b + 1);
      }

      // ---- phase A: kernel integrals (once per pair) and edge erfs (once per live edge) ---
      // Items are enumerated through the prefix sums; a warp takes 2 x 32 consecutive items
      // per step, so its lanes mostly share a particle (coherent branches, conflict-free
      // rows) and every lane carries two independent dependency chains (the tabulated
      // evaluators are straight-line code, so the two interleave).
      {
        const uint32_t total = S.wprefix[PBATCH];
        const uint32_t my_start = S.wprefix[lane];
        const bool my_nonempty = S.wprefix[lane + 1] > my_start;
        constexpr int NW = 2;  // independent evaluation chains per lane
        // chunks of 32 NW items are handed out through a shared counter: warps 0 and 1 come late
        // (they set up the next batch first) and chunks differ in cost, so a fixed striping left
        // the other warps waiting at the barrier below
        for (;;) {
          uint32_t q0 = 0;
          if (lane == 0) q0 = atomicAdd(&sm.ctr[buf][0], 1u) * (32u * NW);
          q0 = __shfl_sync(0xffffffffu, q0, 0);
          if (q0 >= total) break;
          int ord[NW];
#pragma unroll
          for (int u = 0; u < NW; ++u) ord[u] = owner_ordinal(q0 + 32 * u, my_start, my_nonempty, lane);
          bool ok[NW];
          int pp[NW], pix[NW], kind[NW], kid[NW];
          double dx[NW], dy[NW], R2[NW], ih2[NW], tv[NW], amp[NW];
#pragma unroll
          for (int u = 0; u < NW; ++u) {
            const uint32_t q = q0 + 32 * u + lane;
            ok[u] = q < total;
            const int p = ok[u] ? S.wowner[ord[u]] : S.wowner[0];
            const uint32_t local = ok[u] ? q - S.wprefix[p] : 0u;
            const int ix = (int)(((float)local + 0.5f) * S.rny[p]);
            const int iy = (int)local - ix * S.box[p][3];
            const int tpx = S.box[p][0] + ix, tpy = S.box[p][2] + iy;
            const Record& r = sm.rec[buf][p];
            pp[u] = p;
            pix[u] = w_index(tpx, tpy);
            kid[u] = r.kid;
            amp[u] = r.amp;
            kind[u] = (KIND >= 0 && kid[u] == 0) ? KIND : a.table.kind[kid[u]];
            // dij = pixcoords - ij (martini.py:276)
            dx[u] = __dsub_rn(r.px, (double)(x0 + tpx));
            dy[u] = __dsub_rn(r.py, (double)(y0 + tpy));
            ih2[u] = r.inv_h2;
            R2[u] = sq_dist(dx[u], dy[u]) * ih2[u];
          }
#pragma unroll
          for (int u = 0; u < NW; ++u)  // straight-line, both chains in flight together
            tv[u] = wtab_eval(KIND >= 0 ? KIND : (wtab_has(kind[u]) ? kind[u] : MTN_KERNEL_WENDLANDC2), R2[u]) * ih2[u];
#pragma unroll
          for (int u = 0; u < NW; ++u) {
            if (ok[u]) {
              double w = tv[u];
              // closed form: kernels without a table; with KIND, every entry but the first
              if (KIND >= 0 ? kid[u] != 0 : !wtab_has(kind[u])) {
                const Record& r = sm.rec[buf][pp[u]];
                w = kernel_weight_closed(kind[u], dx[u], dy[u], r.h, r.inv_h2, a.table.truncate[r.kid],
                                         a.table.norm[r.kid]);
              }
              sm.W[pp[u]][pix[u]] = w * amp[u];
              if (COUNT) ++n_w;
            }
          }
        }
      }
      if (gaussian_line) {
        const uint32_t total = S.eprefix[PBATCH];
        const uint32_t my_start = S.eprefix[lane];
        const bool my_nonempty = S.eprefix[lane + 1] > my_start;
        for (;;) {
          uint32_t q0 = 0;
          if (lane == 0) q0 = atomicAdd(&sm.ctr[buf][1], 1u) * 64u;
          q0 = __shfl_sync(0xffffffffu, q0, 0);
          if (q0 >= total) break;
          const int ord[2] = {owner_ordinal(q0, my_start, my_nonempty, lane),
                              owner_ordinal(q0 + 32, my_start, my_nonempty, lane)};
          bool ok[2];
          int pp[2], ee[2];
          double t[2], ev[2];
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            const uint32_t q = q0 + 32 * u + lane;
            ok[u] = q < total;
            const int p = ok[u] ? S.eowner[ord[u]] : S.eowner[0];
            const int e = ok[u] ? S.erun[p][0] + (int)(q - S.eprefix[p]) : 0;
            const Record& r = sm.rec[buf][p];
            pp[u] = p;
            ee[u] = e;
            // g orientation (sign folded in): S[c] = E[c+1] - E[c] >= 0; saturated edges at
            // the ends of the run come out as exactly -1 / +1
            t[u] = (sm.edge[e] - r.v) * (sgn * r.inv_s);
          }
#pragma unroll
          for (int u = 0; u < 2; ++u) ev[u] = erf_tab_compact(sm.erf_table, t[u]);
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            if (ok[u]) {
              sm.ES[pp[u]][es_pos(ee[u])] = ev[u];
              if (COUNT) n_erf += fabs(t[u]) < ERF_SAT;
            }
          }
        }
      }
      // ---- edges outside the live window: thread = (particle, 16-edge quarter).  The reference
      // evaluates erf at every edge; outside [cs, ce] it is saturated, -1 below and +1 above in
      // the g orientation (plan.cuh: channel_window), so the differences E[c+1] - E[c] of the dead
      // channels are exactly zero without a predicate in phase C.  The Dirac line is written as
      // the step function whose differences are 1 on the live channels [cs, ce).
      {
        const int p = tid >> 2, qtr = tid & 3;
        if (p < nb && S.hlive[p]) {
          const int cs = S.chan[p][0], ce = S.chan[p][1];
          double* row = sm.ES[p];
          if (gaussian_line) {
            // eight aligned edge pairs per thread: a pair wholly outside the window is one
            // 16-byte store, a pair the window's end cuts gets its outside edge alone
#pragma unroll
            for (int m = 0; m < 8; ++m) {
              const int e = 16 * qtr + 2 * m;
              if (e + 1 < cs) {
                *reinterpret_cast<double2*>(row + es_pos(e)) = make_double2(-1.0, -1.0);
              } else if (e > ce) {
                *reinterpret_cast<double2*>(row + es_pos(e)) = make_double2(1.0, 1.0);
              } else {
                if (e < cs) row[es_pos(e)] = -1.0;
                if (e + 1 > ce) row[es_pos(e + 1)] = 1.0;
              }
            }
            if (qtr == 3 && ce < CB) row[es_pos(CB)] = 1.0;
          } else {
            const int e_end = qtr == 3 ? CB + 1 : 16 * qtr + 16;
            for (int e = 16 * qtr; e < e_end; ++e) row[es_pos(e)] = (double)(min(max(e, cs), ce) - cs);
          }
        }
      }
      __syncthreads();

      if (tid < 2) sm.ctr[buf ^ 1u][tid] = 0u;  // (last used by batch b - 1; batch b + 1 starts after the barrier below)
      // ---- warp-private list: which particles touch my sub-block.  Lane = particle: W is zero
      // wherever phase A did not write (outside the box) and where the kernel's support ends, so
      // "any of my sub-block's 16 weights non-zero" is the whole test.
      bool visit = false;
      if (lane < nb && S.hlive[lane]) {
        const double2* wsub = reinterpret_cast<const double2*>(&sm.W[lane][sub * SUB_PIX]);
#pragma unroll
        for (int j = 0; j < SUB_PIX / 2; ++j) {
          const double2 w2 = wsub[j];
          visit |= (w2.x != 0.0) | (w2.y != 0.0);
        }
      }
      uint32_t rel = __ballot_sync(0xffffffffu, visit);

      // ---- phase C: acc[4 pixels][8 channels] += (W amp) (E[c+1] - E[c]) ------------------------
      // (A version that issued the next half visit's loads ahead of the FMAs measured 2.5 % slower:
      // the register rotation costs more than the latency it hides.)
      {
        const int pr = lane >> 3, cg = lane & 7;
        const double* wbase = &sm.W[0][sub * SUB_PIX + pr * SUB_Y];
#if MTN_TILE_IL
        const double* ebase = &sm.ES[0][2 * cg];
#else
        const double* ebase = &sm.ES[0][es_pos(8 * cg)];
#endif
        while (rel) {
          const int p = __ffs(rel) - 1;
          rel &= rel - 1;
          const double2* wp = reinterpret_cast<const double2*>(wbase + p * W_STRIDE);
          const double* ep = ebase + p * ES_STRIDE;
          const double2 w01 = wp[0], w23 = wp[1];
          double d[8];
#if MTN_TILE_IL
          double2 e01[4];
          double e2[4];
#pragma unroll
          for (int h = 0; h < 4; ++h) {
            e01[h] = *reinterpret_cast<const double2*>(ep + 16 * h);
            e2[h] = ep[16 * h + 2];
          }
#pragma unroll
          for (int h = 0; h < 4; ++h) {
            d[2 * h] = e01[h].y - e01[h].x;
            d[2 * h + 1] = e2[h] - e01[h].y;
          }
#else
          {
            const double2* e2p = reinterpret_cast<const double2*>(ep);
            const double2 ea = e2p[0], eb = e2p[1], ec = e2p[2], ed = e2p[3];
            const double e8 = ep[10];  // = es_pos(8 cg + 8): first edge of the next group
            d[0] = ea.y - ea.x; d[1] = eb.x - ea.y; d[2] = eb.y - eb.x; d[3] = ec.x - eb.y;
            d[4] = ec.y - ec.x; d[5] = ed.x - ec.y; d[6] = ed.y - ed.x; d[7] = e8 - ed.y;
          }
#endif
          const double w[4] = {w01.x, w01.y, w23.x, w23.y};
#pragma unroll
          for (int k = 0; k < 4; ++k)
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              acc[k][q] = fma(w[k], d[q], acc[k][q]);
              if (COUNT) n_upd += (w[k] != 0.0) && (d[q] != 0.0);
            }
        }
        // my sub-block's share of W back to zero (the invariant the mask above relies on): 32
        // particles x 128 bytes, eight 16-byte stores per lane, after my own last read of it
        __syncwarp();
#pragma unroll
        for (int m = 0; m < 8; ++m) {
          const int idx = m * 32 + lane;
          *reinterpret_cast<double2*>(&sm.W[idx >> 3][sub * SUB_PIX + 2 * (idx & 7)]) = make_double2(0.0, 0.0);
        }
      }
      __syncthreads();  // W, ES, boxes and rec[buf] are free again
      if (b + 2 < n_batch) issue(b + 2);
    }

    // ---- one store per voxel: x 1 / dv of the channel (spectral_models.py:139), then
    // out = (in + acc) / px_area; accumulators (2h, 2h + 1) of a lane are the channel pair starting
    // at pair_ch(h): per pixel and h the eight channel groups of a warp write one contiguous
    // 128-byte line -----------------------------------------------------------------------
    {
      const int pr = lane >> 3, cg = lane & 7;
      const int tpx = (sub / SUBS_Y) * SUB_X + pr, tpy0 = (sub % SUBS_Y) * SUB_Y;
#if MTN_TILE_IL
      auto pair_ch = [&](int h) { return 2 * cg + 16 * h; };
#else
      auto pair_ch = [&](int h) { return 8 * cg + 2 * h; };
#endif
#pragma unroll
      for (int h = 0; h < 4; ++h) {
        const double2 i2 = *reinterpret_cast<const double2*>(&sm.inv_dv[pair_ch(h)]);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          acc[k][2 * h] *= i2.x;
          acc[k][2 * h + 1] *= i2.y;
        }
      }
      if (it.slot >= 0) {
        double* dst = a.partials + (size_t)it.slot * TILE_PIX * CB;
#pragma unroll
        for (int k = 0; k < 4; ++k)
#pragma unroll
          for (int h = 0; h < 4; ++h)
            *reinterpret_cast<double2*>(dst + (size_t)(tpx * TILE_Y + tpy0 + k) * CB + pair_ch(h)) =
                make_double2(acc[k][2 * h], acc[k][2 * h + 1]);
      } else {
        const int gx = x0 + tpx;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int gy = y0 + tpy0 + k;
          if (gx < g.x_hi && gy < g.ny) {
            double* dst = a.slab + ((size_t)(gx - g.x_lo) * g.ny + gy) * g.C + c0;
#pragma unroll
            for (int h = 0; h < 4; ++h) {
              const int cl = pair_ch(h);  // first channel of the pair within the brick
              const int nvalid = cl < clo ? 0 : max(0, min(2, nch - cl));
              if (nvalid > 0)
                store2(dst + cl, acc[k][2 * h], acc[k][2 * h + 1], nvalid, a.px_area, !a.zeroed,
                       (g.C & 1) == 0);
            }
          }
        }
      }
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
  const int cb = m.brick % g.ncb, tile = m.brick / g.ncb;
  const int x0 = g.x_lo + (tile / g.nty) * TILE_X, y0 = (tile % g.nty) * TILE_Y;
  const int c0 = g.phase[tile] + (cb - 1) * CB;
  const int sub = warp, cl = 2 * lane;
  const int nvalid = c0 + cl < 0 ? 0 : max(0, min(2, min(CB, g.C - c0) - cl));
  {
    const int j = blockIdx.y;  // one pixel of every sub-block per grid row: the hottest brick
                               // (most partials) is spread over SUB_PIX blocks
    const int tpx = sub_x(sub, j), tpy = sub_y(sub, j);
    const double* src = partials + ((size_t)m.slot0 * TILE_PIX + tpx * TILE_Y + tpy) * CB + cl;
    double a0 = 0.0, a1 = 0.0;
#pragma unroll 8
    for (uint32_t k = 0; k < m.n; ++k) {  // chunk order: deterministic
      const double2 v = *reinterpret_cast<const double2*>(src + (size_t)k * TILE_PIX * CB);
      a0 += v.x;
      a1 += v.y;
    }
    const int gx = x0 + tpx, gy = y0 + tpy;
    if (nvalid > 0 && gx < g.x_hi && gy < g.ny) {
      double* dst = slab + ((size_t)(gx - g.x_lo) * g.ny + gy) * g.C + c0 + cl;
      store2(dst, a0, a1, nvalid, px_area, !zeroed, (g.C & 1) == 0);
    }
  }
}

// Accumulate mode only: voxels of bricks no particle reaches still get in / px_area.
__global__ void __launch_bounds__(PROJ_THREADS) empty_brick_kernel(
    Geo g, const uint32_t* __restrict__ brick_count, double* __restrict__ slab, double px_area) {
  const uint32_t brick = blockIdx.x;
  if (brick_count[brick] != 0) return;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int cb = brick % g.ncb, tile = brick / g.ncb;
  const int x0 = g.x_lo + (tile / g.nty) * TILE_X, y0 = (tile % g.nty) * TILE_Y;
  const int c0 = g.phase[tile] + (cb - 1) * CB;
  const int sub = warp, cl = 2 * lane;
  const int nvalid = c0 + cl < 0 ? 0 : max(0, min(2, min(CB, g.C - c0) - cl));
  if (nvalid == 0) return;
  for (int j = 0; j < SUB_PIX; ++j) {
    const int gx = x0 + sub_x(sub, j), gy = y0 + sub_y(sub, j);
    if (gx < g.x_hi && gy < g.ny) {
      double* dst = slab + ((size_t)(gx - g.x_lo) * g.ny + gy) * g.C + c0 + cl;
      store2(dst, 0.0, 0.0, nvalid, px_area, true, (g.C & 1) == 0);
    }
  }
}

}  // namespace mtn
