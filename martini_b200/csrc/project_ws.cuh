// Warp-specialised projection kernel.
//
// Same decomposition as project.cuh (one brick of 8 x 8 pixels x 64 channels per work item,
// sums in registers, records staged with cp.async.bulk), but the two halves of the work run
// on different warps of the CTA and overlap in time:
//
//   producers  stage the particle records of a batch, find each particle's tile-local box and
//              live channel run (lane = particle), then evaluate -- once per (particle, pixel)
//              and once per (particle, live channel edge), handed out to lanes through prefix
//              sums -- the kernel integrals W and the line spectrum S = (E[c+1] - E[c]) amp / dv
//              (the difference of neighbouring edge erfs is formed with a warp shuffle, so
//              the erfs themselves never leave the registers) into one of two shared-memory
//              stages; they hold no accumulators, so they have the registers to keep several
//              evaluation chains in flight;
//   consumers  one warp per 4 x 4 pixel sub-block, lane = 2 channels, 32 float64 accumulators
//              per thread: for every particle of a full stage whose box meets the sub-block,
//              acc[pixel] += W * S with W a shared-memory broadcast (two pixels per 16-byte
//              load) and S one 16-byte load per lane.
//
// Stages are handed over through mbarriers (full: producers -> consumers, empty: back), so
// the consumers never meet a CTA-wide barrier and the producers only synchronise among
// themselves (named barrier 1).  The last stage of a work item carries the brick to store.
#pragma once

#include "project.cuh"

namespace mtn {

static_assert(N_HALF == 1 && SUB_X == 4 && SUB_Y == 4, "project_ws assumes 64-channel bricks, 4x4 sub-blocks");

constexpr int WS_NC = N_SUB;  // consumer warps: one per sub-block
#ifndef MTN_WS_NP
#define MTN_WS_NP 4
#endif
constexpr int WS_NP = MTN_WS_NP;  // producer warps
constexpr int WS_THREADS = (WS_NC + WS_NP) * 32;
#ifndef MTN_WS_CTAS
#define MTN_WS_CTAS 2
#endif
constexpr int WS_CTAS_PER_SM = MTN_WS_CTAS;
// With 8 producer warps the CTA is launched at 80 registers per thread and re-splits its
// register file by role (setmaxnreg acts on aligned groups of 4 warps).
#ifndef MTN_WS_CONSUMER_REGS
#define MTN_WS_CONSUMER_REGS 112
#endif
#ifndef MTN_WS_PRODUCER_REGS
#define MTN_WS_PRODUCER_REGS 64
#endif
#ifndef MTN_WS_RESPLIT
#define MTN_WS_RESPLIT (MTN_WS_NP > 4)
#endif
constexpr bool WS_RESPLIT_REGS = MTN_WS_RESPLIT;
static_assert(WS_NC == 4 && WS_NP % 4 == 0, "roles must be whole warpgroups");
constexpr int WS_STAGES = 2;
constexpr int WS_W_STRIDE = TILE_PIX + 2;  // even: a pixel pair is one aligned 16-byte load
constexpr uint32_t WS_LAST = 1u, WS_TERMINATE = 2u;

struct WsStage {
  double W[PBATCH][WS_W_STRIDE];  // kernel integrals, valid inside the particle's box
  double S[PBATCH][CB];           // line spectrum on the brick's channels (zero outside the line)
  uint32_t cbox[PBATCH];          // x0 | nx << 8 | y0 << 16 | ny << 24, 0: contributes nothing
  uint32_t nb;                    // particles in the batch
  uint32_t flags;                 // WS_LAST: store the brick after this batch
  uint32_t brick;
  int32_t slot;
};

struct WsBatch {  // producer-private, double buffered by batch parity
  uint32_t wprefix[PBATCH + 1];  // exclusive prefix of box areas
  uint32_t eprefix[PBATCH + 1];  // exclusive prefix of edge-run lengths
  float rny[PBATCH];             // 1 / (box height)
  uint8_t wowner[PBATCH];        // particles with a non-empty box, in order
  uint8_t eowner[PBATCH];        // particles with a non-empty edge run, in order
  uint8_t box[PBATCH][4];        // tile-local box: x0, nx, y0, ny
  uint8_t e01[PBATCH][2];        // first / one-past-last live edge
  uint8_t erun[PBATCH][2];       // first item (edge or channel) to evaluate, number of items
  uint8_t chan[PBATCH][2];       // live channels of the brick: [cs, ce)
};

struct WsSmem {
  WsStage st[WS_STAGES];
  Record rec[2][PBATCH];
  double inv_dv[CB];
  double edge[CB + 1];
  WsBatch bt[2];
  uint64_t full[WS_STAGES], empty[WS_STAGES], recbar[2];
  uint32_t item;
};

// a / b for a fixed positive normal b, given rb = RN(1 / b): q = RN(a rb), then one exact
// residual correction (Markstein) -- the correctly rounded quotient, like the reference's own
// division, but straight-line code: no call into the division subroutine, which would keep
// the compiler from giving the consumer branch its larger register budget.
__device__ __forceinline__ double div_by(double a, double b, double rb) {
  const double q = a * rb;
  const double q1 = fma(fma(-b, q, a), rb, q);
  return (fabs(q) < CUDART_INF) ? q1 : q;  // inf / nan pass through
}
// 1 / x to within an ulp, straight-line (x positive and normal)
__device__ __forceinline__ double rcp_newton(double x) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  r = fma(fma(-x, r, 1.0), r, r);
  r = fma(fma(-x, r, 1.0), r, r);
  r = fma(fma(-x, r, 1.0), r, r);
  return r;
}
// 1 / x for a small positive integer-valued x (<= 1 ulp; no slow-path subroutine)
__device__ __forceinline__ float rcp_approx_f32(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
// store2 of project.cuh with the division by the pixel area done by div_by
__device__ __forceinline__ void store2_ws(double* __restrict__ dst, double a0, double a1, int nvalid,
                                          double px_area, double inv_area, bool add_in, bool vec_ok) {
  if (nvalid == 2 && vec_ok) {
    double2 o = make_double2(a0, a1);
    if (add_in) {
      const double2 i2 = *reinterpret_cast<const double2*>(dst);
      o.x += i2.x;
      o.y += i2.y;
    }
    o.x = div_by(o.x, px_area, inv_area);
    o.y = div_by(o.y, px_area, inv_area);
    *reinterpret_cast<double2*>(dst) = o;
  } else {
    if (nvalid >= 1) dst[0] = div_by((add_in ? dst[0] : 0.0) + a0, px_area, inv_area);
    if (nvalid >= 2) dst[1] = div_by((add_in ? dst[1] : 0.0) + a1, px_area, inv_area);
  }
}

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// barrier among the producer warps only
__device__ __forceinline__ void producer_sync() {
  asm volatile("bar.sync 1, %0;" ::"n"(WS_NP * 32) : "memory");
}

// --------------------------------------------------------------------------------- producers
template <bool COUNT, int KIND>
__device__ __forceinline__ void ws_producer(const ProjArgs& a, WsSmem& sm, const int ptid) {
  constexpr int NPT = WS_NP * 32;
  const int pwarp = ptid >> 5, lane = ptid & 31;
  const Geo& g = a.geo;
  const bool gaussian_line = g.spectrum == MTN_SPECTRUM_GAUSSIAN;
  const double sgn = g.edges_increasing ? 1.0 : -1.0;
  uint32_t rec_phase = 0;               // bit k: parity to wait for on recbar[k]
  uint32_t empty_phase = (1u << WS_STAGES) - 1u;  // fresh barriers: the "previous" phase is complete
  uint32_t stage = 0, batch_no = 0;
  unsigned long long n_w = 0, n_erf = 0;

  for (;;) {
    if (ptid == 0) sm.item = atomicAdd(a.counter, 1u);
    producer_sync();  // also: every producer is done with the previous item
    const uint32_t item_idx = sm.item;
    if (item_idx >= *a.n_items) break;
    const Item it = a.items[item_idx];
    const int cb = it.brick % g.ncb;
    const int tile = it.brick / g.ncb;
    const int ty = tile % g.nty, tx = tile / g.nty;
    // channels [c0, c0 + CB) of the cube; c0 may be negative (block 0 of a phased tile)
    const int x0 = g.x_lo + tx * TILE_X, y0 = ty * TILE_Y, c0 = g.phase[tile] + (cb - 1) * CB;
    const int clo = max(0, -c0), nch = min(CB, g.C - c0);  // valid brick channels: [clo, nch)
    const int x_last = min(x0 + TILE_X, g.x_hi) - 1, y_last = min(y0 + TILE_Y, g.ny) - 1;

    for (int e = ptid; e <= CB; e += NPT) sm.edge[e] = a.edges[min(max(c0 + e, 0), g.C)];
    for (int c = ptid; c < CB; c += NPT)
      sm.inv_dv[c] = (c >= clo && c < nch) ? rcp_newton(fabs(a.edges[c0 + c + 1] - a.edges[c0 + c])) : 0.0;

    const uint32_t n_part = it.end - it.begin;
    const uint32_t n_batch = (n_part + PBATCH - 1) / PBATCH;
    auto issue = [&](uint32_t b) {
      const uint32_t buf = b & 1u;
      const uint32_t nb = min((uint32_t)PBATCH, n_part - b * PBATCH);
      if (ptid < (int)nb) {
        const uint32_t ridx = (uint32_t)a.pairs[it.begin + b * PBATCH + ptid];
        bulk_g2s(&sm.rec[buf][ptid], a.records + ridx, REC_BYTES, &sm.recbar[buf]);
      }
      if (ptid == 0) mbar_arrive_expect_tx(&sm.recbar[buf], nb * REC_BYTES);
    };
    issue(0);
    producer_sync();  // edge table visible

    for (uint32_t b = 0; b < n_batch; ++b, ++batch_no) {
      const uint32_t buf = b & 1u;
      const int nb = (int)min((uint32_t)PBATCH, n_part - b * PBATCH);
      WsBatch& bt = sm.bt[batch_no & 1u];
      WsStage& st = sm.st[stage];
      mbar_wait(&sm.recbar[buf], (rec_phase >> buf) & 1u);
      rec_phase ^= 1u << buf;

      // ---- setup stage 1, lane = particle: the four independent searches on four warps ----
      for (int task = pwarp; task < 4; task += WS_NP) {
        int v0 = 0, v1 = 0;
        if (lane < nb) {
          const Record& r = sm.rec[buf][lane];
          if (task < 2) {  // candidate box of martini.py:272-274 along x (task 0) or y (task 1)
            int lo, hi;
            const bool any = task == 0 ? pixel_bounds(r.px, (double)r.r, x0, x_last, lo, hi)
                                       : pixel_bounds(r.py, (double)r.r, y0, y_last, lo, hi);
            if (any) {
              v0 = lo - (task == 0 ? x0 : y0);
              v1 = hi - lo + 1;
            }
          } else {
            // g(e) = sgn * (edge[e] - v) * inv_s is non-decreasing in the edge index e.
            // task 2: e0 = first edge with g > -SAT (Gaussian) / g >= 0 (Dirac)
            // task 3: e1 = first edge with g >= SAT (Gaussian) / g > 0 (Dirac)
            const double v = r.v, sc = gaussian_line ? sgn * r.inv_s : sgn;
            const double thr = gaussian_line ? (task == 2 ? -ERF_SAT : ERF_SAT) : 0.0;
            const bool strict = gaussian_line ? task == 2 : task == 3;  // '>' vs '>='
            int l = 0, h = CB + 1;
            while (l < h) {
              const int m = (l + h) >> 1;
              const double x = (sm.edge[m] - v) * sc;
              if (strict ? (x > thr) : (x >= thr)) h = m; else l = m + 1;
            }
            v0 = l;
          }
        }
        if (task < 2) {
          bt.box[lane][2 * task] = (uint8_t)v0;
          bt.box[lane][2 * task + 1] = (uint8_t)v1;
          if (task == 1) bt.rny[lane] = v1 ? rcp_approx_f32((float)v1) : 0.0f;
        } else {
          bt.e01[lane][task - 2] = (uint8_t)v0;
        }
      }
      producer_sync();  // every producer is past phase A of the previous batch
      if (b + 1 < n_batch) issue(b + 1);  // its record buffer is free now

      // ---- stage 2: prefix sums, owner lists, what the consumers need to know -------------
      mbar_wait(&sm.empty[stage], (empty_phase >> stage) & 1u);
      empty_phase ^= 1u << stage;
      if (pwarp < 2) {
        // channel c can be non-zero only if edge c+1 >= e0 and edge c < e1:
        //   Gaussian: some edge of the channel is unsaturated, or the saturation flips in it
        //   Dirac   : lo <= v <= hi, both closed (spectral_models.py:564-569)
        const int e0 = bt.e01[lane][0], e1 = bt.e01[lane][1];
        int cs = max(e0 - 1, clo), ce = min(e1, nch);
        if (cs >= ce || lane >= nb) cs = ce = 0;
        const int bnx = bt.box[lane][1], bny = bt.box[lane][3];
        const uint32_t area = (ce > cs && bnx && bny) ? (uint32_t)(bnx * bny) : 0u;
        // items of the spectrum pass: the edges of the live channels (Gaussian) or the live
        // channels themselves (Dirac); none if no pixel is reached
        const uint32_t ne = area ? (uint32_t)(ce - cs + (gaussian_line ? 1 : 0)) : 0u;
        const uint32_t lt = (1u << lane) - 1u;
        if (pwarp == 0) {
          bt.chan[lane][0] = (uint8_t)cs;
          bt.chan[lane][1] = (uint8_t)ce;
          st.cbox[lane] = area ? ((uint32_t)bt.box[lane][0] | (uint32_t)bnx << 8 |
                                  (uint32_t)bt.box[lane][2] << 16 | (uint32_t)bny << 24)
                               : 0u;
          const uint32_t wi = warp_incl_scan_u32(area, lane);
          bt.wprefix[lane] = wi - area;
          if (lane == 31) bt.wprefix[32] = wi;
          const uint32_t wm = __ballot_sync(0xffffffffu, area != 0);
          if (area != 0) bt.wowner[__popc(wm & lt)] = (uint8_t)lane;
          if (lane == 0) {
            st.nb = (uint32_t)nb;
            st.flags = b + 1 == n_batch ? WS_LAST : 0u;
            st.brick = it.brick;
            st.slot = it.slot;
          }
        }
        if (pwarp == WS_NP - 1 || pwarp == 1) {  // warp 1, or warp 0 again with one producer warp
          bt.erun[lane][0] = (uint8_t)cs;
          bt.erun[lane][1] = (uint8_t)ne;
          const uint32_t ei = warp_incl_scan_u32(ne, lane);
          bt.eprefix[lane] = ei - ne;
          if (lane == 31) bt.eprefix[32] = ei;
          const uint32_t em = __ballot_sync(0xffffffffu, ne != 0);
          if (ne != 0) bt.eowner[__popc(em & lt)] = (uint8_t)lane;
        }
      }
      producer_sync();

      // ---- kernel integrals: one per (particle, box pixel), 2 x 32 consecutive pairs per
      // warp step so that two straight-line evaluation chains interleave ---------------------
      {
        const uint32_t total = bt.wprefix[PBATCH];
        const uint32_t my_start = bt.wprefix[lane];
        const bool my_nonempty = bt.wprefix[lane + 1] > my_start;
        for (uint32_t q0 = pwarp * 64; q0 < total; q0 += WS_NP * 64) {
          const int ord[2] = {owner_ordinal(q0, my_start, my_nonempty, lane),
                              owner_ordinal(q0 + 32, my_start, my_nonempty, lane)};
          bool ok[2];
          int pp[2], pix[2], kind[2], kid[2];
          double dx[2], dy[2], R2[2], ih2[2], tv[2];
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            const uint32_t q = q0 + 32 * u + lane;
            ok[u] = q < total;
            const int p = ok[u] ? bt.wowner[ord[u]] : bt.wowner[0];
            const uint32_t local = ok[u] ? q - bt.wprefix[p] : 0u;
            const int ix = (int)(((float)local + 0.5f) * bt.rny[p]);
            const int iy = (int)local - ix * bt.box[p][3];
            const int tpx = bt.box[p][0] + ix, tpy = bt.box[p][2] + iy;
            const Record& r = sm.rec[buf][p];
            pp[u] = p;
            pix[u] = tpx * TILE_Y + tpy;
            kid[u] = r.kid;
            kind[u] = (KIND >= 0 && kid[u] == 0) ? KIND : a.table.kind[kid[u]];
            // dij = pixcoords - ij (martini.py:276)
            dx[u] = __dsub_rn(r.px, (double)(x0 + tpx));
            dy[u] = __dsub_rn(r.py, (double)(y0 + tpy));
            ih2[u] = r.inv_h2;
            R2[u] = sq_dist(dx[u], dy[u]) * ih2[u];
          }
#pragma unroll
          for (int u = 0; u < 2; ++u)
            tv[u] = wtab_eval(KIND >= 0 ? KIND : (wtab_has(kind[u]) ? kind[u] : MTN_KERNEL_WENDLANDC2), R2[u]) * ih2[u];
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            if (ok[u]) {
              double w = tv[u];
              // closed form: kernels without a table; with KIND, every entry but the first
              if (KIND >= 0 ? kid[u] != 0 : !wtab_has(kind[u])) {
                const Record& r = sm.rec[buf][pp[u]];
                w = kernel_weight_closed(kind[u], dx[u], dy[u], r.h, r.inv_h2, a.table.truncate[r.kid],
                                         a.table.norm[r.kid]);
              }
              st.W[pp[u]][pix[u]] = w;
              if (COUNT) ++n_w;
            }
          }
        }
      }
      // ---- line spectra.  Gaussian: the items are the edges of the live channels; a warp
      // step evaluates 64 consecutive edge erfs and forms the 63 differences between
      // neighbours (steps overlap by one edge), S = 0.5 [erf(hi) - erf(lo)] A / dv / 2.36e5
      // with the 0.5 and 2.36e5 inside amp; saturated edges come out as exactly -1 / +1.
      // Dirac: the items are the live channels, exactly those with lo <= v <= hi. -----------
      {
        const uint32_t total = bt.eprefix[PBATCH];
        const uint32_t my_start = bt.eprefix[lane];
        const bool my_nonempty = bt.eprefix[lane + 1] > my_start;
        const uint32_t stride = gaussian_line ? 63u : 64u;
        for (uint32_t q0 = pwarp * stride; q0 < total; q0 += WS_NP * stride) {
          const int ord[2] = {owner_ordinal(q0, my_start, my_nonempty, lane),
                              owner_ordinal(q0 + 32, my_start, my_nonempty, lane)};
          bool ok[2], more[2];
          int pp[2], ee[2];
          double t[2], ev[2], scale[2];
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            const uint32_t q = q0 + 32 * u + lane;
            ok[u] = q < total;
            const int p = ok[u] ? bt.eowner[ord[u]] : bt.eowner[0];
            const uint32_t local = ok[u] ? q - bt.eprefix[p] : 0u;
            const int e = ok[u] ? bt.erun[p][0] + (int)local : 0;
            more[u] = local + 1u < bt.erun[p][1];  // not the last edge of its run
            const Record& r = sm.rec[buf][p];
            pp[u] = p;
            ee[u] = e;
            t[u] = (sm.edge[e] - r.v) * (sgn * r.inv_s);
            scale[u] = r.amp * sm.inv_dv[min(e, CB - 1)];
          }
          if (gaussian_line) {
#pragma unroll
            for (int u = 0; u < 2; ++u) ev[u] = erf_tab(t[u]);
            const double first1 = __shfl_sync(0xffffffffu, ev[1], 0);
            double en[2];
            en[0] = __shfl_down_sync(0xffffffffu, ev[0], 1);
            en[1] = __shfl_down_sync(0xffffffffu, ev[1], 1);
            if (lane == 31) en[0] = first1;
#pragma unroll
            for (int u = 0; u < 2; ++u) {
              if (COUNT && ok[u] && !(u == 1 && lane == 31)) n_erf += fabs(t[u]) < ERF_SAT;
              if (ok[u] && more[u] && !(u == 1 && lane == 31))
                st.S[pp[u]][ee[u]] = (en[u] - ev[u]) * scale[u];
            }
          } else {
#pragma unroll
            for (int u = 0; u < 2; ++u)
              if (ok[u]) st.S[pp[u]][ee[u]] = scale[u];
          }
        }
      }
      // channels outside the line are exact zeros
      for (int p = pwarp; p < nb; p += WS_NP) {
        if (st.cbox[p] == 0u) continue;  // the consumers skip it
        const int cs = bt.chan[p][0], ce = bt.chan[p][1];
#pragma unroll
        for (int k = 0; k < CB / 32; ++k) {
          const int c = lane + 32 * k;
          if (c < cs || c >= ce) st.S[p][c] = 0.0;
        }
      }
      mbar_arrive(&sm.full[stage]);
      stage = (stage + 1) % WS_STAGES;
    }
  }
  // tell the consumers to stop
  mbar_wait(&sm.empty[stage], (empty_phase >> stage) & 1u);
  if (ptid == 0) sm.st[stage].flags = WS_TERMINATE;
  mbar_arrive(&sm.full[stage]);
  if (COUNT) {
    atomicAdd(a.exec_counts + 1, n_w);
    atomicAdd(a.exec_counts + 2, n_erf);
  }
}

// --------------------------------------------------------------------------------- consumers
template <bool COUNT>
__device__ __forceinline__ void ws_consumer(const ProjArgs& a, WsSmem& sm, const int sub, const int lane) {
  const Geo& g = a.geo;
  const int sx0 = (sub / SUBS_Y) * SUB_X, sy0 = (sub % SUBS_Y) * SUB_Y;
  // an opaque zero: keeps the compiler from hoisting the accumulator initialisation above the
  // role branch, i.e. above the setmaxnreg that gives this branch its registers
  double zero;
  asm volatile("mov.f64 %0, 0d0000000000000000;" : "=d"(zero));
  double acc[SUB_PIX][2];
#pragma unroll
  for (int j = 0; j < SUB_PIX; ++j) acc[j][0] = acc[j][1] = zero;
  uint32_t full_phase = 0, stage = 0;
  unsigned long long n_upd = 0;

  for (;;) {
    mbar_wait(&sm.full[stage], (full_phase >> stage) & 1u);
    full_phase ^= 1u << stage;
    const WsStage& st = sm.st[stage];
    const uint32_t flags = st.flags;
    if (flags & WS_TERMINATE) break;
    const int nb = (int)st.nb;

    // which pixels of my sub-block does particle `lane` reach: its box cut to the sub-block
    uint32_t mymask = 0;
    if (lane < nb) {
      const uint32_t cbx = st.cbox[lane];
      const int bx0 = cbx & 0xff, bx1 = bx0 + ((cbx >> 8) & 0xff);
      const int by0 = (cbx >> 16) & 0xff, by1 = by0 + (cbx >> 24);
      const int jx0 = max(bx0 - sx0, 0), jx1 = min(bx1 - sx0, SUB_X);
      const int jy0 = max(by0 - sy0, 0), jy1 = min(by1 - sy0, SUB_Y);
      if (jx1 > jx0 && jy1 > jy0) {
        const uint32_t cols = ((1u << (jy1 - jy0)) - 1u) << jy0;                  // pixel j = jx * 4 + jy
        const uint32_t rows = (0x1111u >> (4 * (SUB_X - (jx1 - jx0)))) << (4 * jx0);
        mymask = rows * cols;
      }
    }
    uint32_t rel = __ballot_sync(0xffffffffu, mymask != 0);
    while (rel) {
      const int p = __ffs(rel) - 1;
      rel &= rel - 1;
      const uint32_t m = __shfl_sync(0xffffffffu, mymask, p);
      const double2 s2 = *reinterpret_cast<const double2*>(&st.S[p][2 * lane]);
      const double* Wp = &st.W[p][sx0 * TILE_Y + sy0];
      // a pixel pair is loaded only if the particle reaches it: the kernel is bound by the
      // shared-memory / L1 data pipe, so the extra branches are cheaper than the extra loads
#pragma unroll
      for (int j = 0; j < SUB_PIX; j += 2) {
        if (m & (3u << j)) {
          const double2 w2 = *reinterpret_cast<const double2*>(Wp + (j / SUB_Y) * TILE_Y + (j % SUB_Y));
          if (m & (1u << j)) {
            acc[j][0] = fma(w2.x, s2.x, acc[j][0]);
            acc[j][1] = fma(w2.x, s2.y, acc[j][1]);
            if (COUNT && w2.x != 0.0) n_upd += (s2.x != 0.0) + (s2.y != 0.0);
          }
          if (m & (2u << j)) {
            acc[j + 1][0] = fma(w2.y, s2.x, acc[j + 1][0]);
            acc[j + 1][1] = fma(w2.y, s2.y, acc[j + 1][1]);
            if (COUNT && w2.y != 0.0) n_upd += (s2.x != 0.0) + (s2.y != 0.0);
          }
        }
      }
    }
    const uint32_t brick = st.brick;
    const int32_t slot = st.slot;
    mbar_arrive(&sm.empty[stage]);  // the stage can be refilled
    stage = (stage + 1) % WS_STAGES;

    if (flags & WS_LAST) {  // ---- one store per voxel --------------------------------------
      // Every path stores and clears the accumulators in one pass, and clears them with the
      // opaque zero (with a literal 0.0 the merge of these paths makes ptxas spill the
      // accumulators around the whole epilogue).
      const int cl = 2 * lane;  // this lane's first channel within the brick
      if (slot >= 0) {
        double* dst = a.partials + (size_t)slot * TILE_PIX * CB + cl;
#pragma unroll
        for (int j = 0; j < SUB_PIX; ++j) {
          *reinterpret_cast<double2*>(dst + (size_t)(sub_x(sub, j) * TILE_Y + sub_y(sub, j)) * CB) =
              make_double2(acc[j][0], acc[j][1]);
          acc[j][0] = acc[j][1] = zero;
        }
      } else {
        const int cb = brick % g.ncb, tile = brick / g.ncb;
        const int x0 = g.x_lo + (tile / g.nty) * TILE_X, y0 = (tile % g.nty) * TILE_Y;
        const int c0 = g.phase[tile] + (cb - 1) * CB;
        const int clo = max(0, -c0), nch = min(CB, g.C - c0);
        const int nvalid = cl < clo ? 0 : max(0, min(2, nch - cl));
        double* dst = a.slab + ((size_t)(x0 + sx0 - g.x_lo) * g.ny + (y0 + sy0)) * g.C + c0 + cl;
        const size_t row = (size_t)g.ny * g.C;
        if (nvalid == 2 && a.zeroed && (g.C & 1) == 0 && x0 + TILE_X <= g.x_hi && y0 + TILE_Y <= g.ny) {
          // the common case, straight-line: whole tile inside the slab, fresh cube, both channels
#pragma unroll
          for (int j = 0; j < SUB_PIX; ++j) {
            *reinterpret_cast<double2*>(dst + (j / SUB_Y) * row + (size_t)(j % SUB_Y) * g.C) =
                make_double2(div_by(acc[j][0], a.px_area, a.inv_px_area),
                             div_by(acc[j][1], a.px_area, a.inv_px_area));
            acc[j][0] = acc[j][1] = zero;
          }
        } else {
#pragma unroll
          for (int j = 0; j < SUB_PIX; ++j) {
            const int gx = x0 + sx0 + j / SUB_Y, gy = y0 + sy0 + j % SUB_Y;
            if (nvalid > 0 && gx < g.x_hi && gy < g.ny)
              store2_ws(dst + (j / SUB_Y) * row + (size_t)(j % SUB_Y) * g.C, acc[j][0], acc[j][1], nvalid,
                        a.px_area, a.inv_px_area, !a.zeroed, (g.C & 1) == 0);
            acc[j][0] = acc[j][1] = zero;
          }
        }
      }
    }
  }
  if (COUNT) atomicAdd(a.exec_counts + 0, n_upd);
}

// COUNT = true additionally tallies the executed algorithmic work (diagnostic, never timed);
// KIND >= 0: every particle uses that tabulated SPH kernel; KIND = -1: the general case.
template <bool COUNT, int KIND>
__global__ void __launch_bounds__(WS_THREADS, WS_CTAS_PER_SM) project_ws_kernel(const ProjArgs a) {
  MTN_DYN_SMEM(unsigned char, smem_raw);
  WsSmem& sm = *reinterpret_cast<WsSmem*>(smem_raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int s = 0; s < WS_STAGES; ++s) {
      mbar_init(&sm.full[s], WS_NP * 32);
      mbar_init(&sm.empty[s], WS_NC * 32);
    }
    mbar_init(&sm.recbar[0], 1);
    mbar_init(&sm.recbar[1], 1);
    mbar_fence_init();
  }
  __syncthreads();
  // role = warpgroup, made warp-uniform for the compiler (it sizes each branch's registers
  // by the setmaxnreg that dominates it)
  const int wgroup = __shfl_sync(0xffffffffu, tid >> 7, 0);
  if (wgroup == 0) {
    if (WS_RESPLIT_REGS) asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(MTN_WS_CONSUMER_REGS));
    ws_consumer<COUNT>(a, sm, warp, lane);
  } else {
    if (WS_RESPLIT_REGS) asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(MTN_WS_PRODUCER_REGS));
    ws_producer<COUNT, KIND>(a, sm, tid - WS_NC * 32);
  }
}

}  // namespace mtn
