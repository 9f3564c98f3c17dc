// C ABI of libmartini_b200.so -- see include/martini_b200.h for the contract.
#include <algorithm>
#include <cstdlib>
#include <mutex>

#include "common.cuh"
#include "convolve.cuh"
#include "frontend.cuh"
#include "kernel_integrals.cuh"
#include "plan.cuh"
#include "project.cuh"
#include "route.cuh"
#include "scan.cuh"
#include "sort.cuh"
#include "streams.cuh"
#include "tables_host.hpp"

namespace mtn {
thread_local char g_err[512] = "";
thread_local int g_launches = 0;

// optional per-stage timing of mtn_project (CUDA events on the caller's stream)
constexpr int N_STAGES = 7;  // emit, sort, items, project, reduce, finalize, stream2
thread_local int g_timing = 0;
thread_local int g_count_exec = 0;
thread_local cudaEvent_t g_ev[N_STAGES + 1];
thread_local bool g_ev_init = false;
thread_local bool g_ev_valid = false;
thread_local unsigned long long g_exec_counts[3] = {0, 0, 0};

static void mark(int i, cudaStream_t st) {
  if (!g_timing) return;
  if (!g_ev_init) {
    for (int k = 0; k <= N_STAGES; ++k) cudaEventCreate(&g_ev[k]);
    g_ev_init = true;
  }
  cudaEventRecord(g_ev[i], st);
}

static int to_dev_table(const MtnKernelTable* t, KernelTableDev* d) {
  if (!t || t->n < 1 || t->n > MTN_MAX_KERNELS) return fail(MTN_ERR_INVALID, "kernel table: bad n%s", "");
  memset(d, 0, sizeof(*d));
  d->n = t->n;
  d->adaptive = t->adaptive;
  for (int i = 0; i < t->n; ++i) {
    const MtnKernelEntry& e = t->k[i];
    if (e.kind < 0 || e.kind > MTN_KERNEL_QUARTICSPLINE)
      return fail(MTN_ERR_INVALID, "kernel table: unknown kind%s %lld", "", (long long)e.kind);
    d->kind[i] = e.kind;
    d->valid_is_max[i] = e.valid_is_max;
    d->rescale[i] = e.rescale;
    d->size_in_fwhm[i] = e.size_in_fwhm;
    d->valid_size[i] = e.valid_size;
    d->truncate[i] = e.truncate;
    d->norm[i] = e.norm;
  }
  return MTN_OK;
}

// Per-device one-time state (a process may drive several GPUs, one host thread each):
// function attributes and __device__ / __constant__ tables belong to the device that was
// current when they were set.
constexpr int MAX_DEVICES = 64;
static std::mutex g_once_mutex;
static int current_device() {
  int dev = 0;
  cudaGetDevice(&dev);
  return (dev >= 0 && dev < MAX_DEVICES) ? dev : 0;
}

static int sm_count() {
  static int n[MAX_DEVICES] = {0};
  const int dev = current_device();
  std::lock_guard<std::mutex> lock(g_once_mutex);
  if (!n[dev]) {
    cudaDeviceGetAttribute(&n[dev], cudaDevAttrMultiProcessorCount, dev);
    if (n[dev] <= 0) n[dev] = 148;
  }
  return n[dev];
}

// Fill the device tables (erf, kernel integrals) once per device; see tables_host.hpp.
static double g_table_err[WT_KINDS] = {0};
static int ensure_tables() {
  static bool done[MAX_DEVICES] = {false};
  const int dev = current_device();
  std::lock_guard<std::mutex> lock(g_once_mutex);
  if (done[dev]) return MTN_OK;
  static double erf_tab_host[ERF_NINT * ERF_NCOEF];
  static double erf_tabc_host[ERFC_DOUBLES];
  static HostTables T;
  static bool built = false;
  if (!built) {
    build_erf_table(erf_tab_host);
    build_erf_table_compact(erf_tabc_host);
    T = build_kernel_tables();
    for (int k = 0; k < WT_KINDS; ++k) g_table_err[k] = T.max_err[k];
    built = true;
  }
  if (T.rows.size() > (size_t)WT_MAX_ROWS * WT_ROW)
    return fail(MTN_ERR_LIMIT, "kernel tables exceed WT_MAX_ROWS%s", "");
  MTN_CUDA(cudaMemcpyToSymbol(g_erf_table, erf_tab_host, sizeof(erf_tab_host)));
  MTN_CUDA(cudaMemcpyToSymbol(g_erf_table_compact, erf_tabc_host, sizeof(erf_tabc_host)));
  MTN_CUDA(cudaMemcpyToSymbol(c_wzone, T.zone, sizeof(T.zone)));
  MTN_CUDA(cudaMemcpyToSymbol(c_wnz, T.nz, sizeof(T.nz)));
  MTN_CUDA(cudaMemcpyToSymbol(c_wscale, T.scale, sizeof(T.scale)));
  MTN_CUDA(cudaMemcpyToSymbol(c_wend, T.end, sizeof(T.end)));
  MTN_CUDA(cudaMemcpyToSymbol(g_wtab_rows, T.rows.data(), T.rows.size() * sizeof(double)));
  done[dev] = true;
  return MTN_OK;
}

// One instantiation per (diagnostic counting, uniform kernel kind).
template <bool COUNT, int KIND>
static int launch_project_as(const ProjArgs& a, int64_t max_items, cudaStream_t st) {
  static bool attr_set_dev[MAX_DEVICES] = {false};
  const int dev = current_device();
  {
    std::lock_guard<std::mutex> lock(g_once_mutex);
    if (!attr_set_dev[dev]) {
      MTN_CUDA(cudaFuncSetAttribute(project_kernel<COUNT, KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)sizeof(ProjSmem)));
      attr_set_dev[dev] = true;
    }
  }
  const unsigned grid = (unsigned)std::min<int64_t>(max_items, (int64_t)sm_count() * PROJ_CTAS_PER_SM);
  auto kfn = project_kernel<COUNT, KIND>;
  MTN_LAUNCH(kfn, grid, PROJ_THREADS, sizeof(ProjSmem), st, a);
  return MTN_OK;
}

template <bool COUNT>
static int launch_project_count(const ProjArgs& a, int primary_kind, int64_t max_items, cudaStream_t st) {
  switch (primary_kind) {
    case MTN_KERNEL_WENDLANDC2: return launch_project_as<COUNT, MTN_KERNEL_WENDLANDC2>(a, max_items, st);
    case MTN_KERNEL_CUBICSPLINE: return launch_project_as<COUNT, MTN_KERNEL_CUBICSPLINE>(a, max_items, st);
    case MTN_KERNEL_WENDLANDC6: return launch_project_as<COUNT, MTN_KERNEL_WENDLANDC6>(a, max_items, st);
    case MTN_KERNEL_QUARTICSPLINE: return launch_project_as<COUNT, MTN_KERNEL_QUARTICSPLINE>(a, max_items, st);
    default: return launch_project_as<COUNT, -1>(a, max_items, st);
  }
}

// The column / splat kernel of an insertion's second stream.
static int launch_stream(const StreamArgs& a, int route, bool count, int64_t max_items, cudaStream_t st) {
  const int64_t want = (max_items + STREAM_WARPS - 1) / STREAM_WARPS;
  if (route == ROUTE_COLUMN) {
    static bool attr_set[MAX_DEVICES][2] = {{false, false}};
    const int dev = current_device();
    const size_t smem = column_smem_bytes(a.geo.C);
    {
      std::lock_guard<std::mutex> lock(g_once_mutex);
      if (!attr_set[dev][count]) {
        const int smem_max = (int)column_smem_bytes(CSB);
        if (count)
          MTN_CUDA(cudaFuncSetAttribute(column_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_max));
        else
          MTN_CUDA(cudaFuncSetAttribute(column_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_max));
        attr_set[dev][count] = true;
      }
    }
    // resident CTAs per SM: shared memory (227 KB per SM, 1 KB reserved per CTA), at most 12 (48 warps)
    const int per_sm = (int)std::max<size_t>(1, std::min<size_t>(12, (227 * 1024) / (smem + 1024)));
    const unsigned grid = (unsigned)std::min<int64_t>(want, (int64_t)sm_count() * per_sm);
    if (count) {
      MTN_LAUNCH(column_kernel<true>, grid, STREAM_THREADS, smem, st, a);
    } else {
      MTN_LAUNCH(column_kernel<false>, grid, STREAM_THREADS, smem, st, a);
    }
  } else {
    const unsigned grid = (unsigned)std::min<int64_t>(want, (int64_t)sm_count() * 5);  // 96 registers
    if (count) {
      MTN_LAUNCH(splat_kernel<true>, grid, STREAM_THREADS, 0, st, a);
    } else {
      MTN_LAUNCH(splat_kernel<false>, grid, STREAM_THREADS, 0, st, a);
    }
  }
  return MTN_OK;
}

static int launch_project(const ProjArgs& a, int primary_kind, bool count, int64_t max_items, cudaStream_t st) {
  return count ? launch_project_count<true>(a, primary_kind, max_items, st)
               : launch_project_count<false>(a, primary_kind, max_items, st);
}

static int make_geo(const MtnCube* c, const MtnKernelTable* table, Geo* g, int edges_increasing) {
  if (!c || c->nx <= 0 || c->ny <= 0 || c->n_channels <= 0)
    return fail(MTN_ERR_INVALID, "cube: bad shape%s", "");
  if (c->x_lo < 0 || c->x_hi > c->nx || c->x_lo >= c->x_hi)
    return fail(MTN_ERR_INVALID, "cube: bad slab rows%s", "");
  if (c->spectrum != MTN_SPECTRUM_GAUSSIAN && c->spectrum != MTN_SPECTRUM_DIRACDELTA)
    return fail(MTN_ERR_INVALID, "cube: unknown spectrum kind%s", "");
  g->nx = c->nx;
  g->ny = c->ny;
  g->C = c->n_channels;
  g->x_lo = c->x_lo;
  g->x_hi = c->x_hi;
  g->ntx = (c->x_hi - c->x_lo + TILE_X - 1) / TILE_X;
  g->nty = (c->ny + TILE_Y - 1) / TILE_Y;
  g->ncb = (c->n_channels + CB - 1) / CB + 1;  // + the partial block below the tile's phase
  g->phase = nullptr;
  const int64_t nb = (int64_t)g->ntx * g->nty * g->ncb;
  if (nb >= (1ll << 31)) return fail(MTN_ERR_LIMIT, "cube: too many bricks%s", "");
  g->n_bricks = (int)nb;
  g->spectrum = c->spectrum;
  g->edges_increasing = edges_increasing;
  g->nsb = (c->n_channels + CSB - 1) / CSB;
  // the second stream (common.cuh: Route): a DiracDelta spectrum sends every particle to the
  // splat kernel; otherwise the particles on a DiracDelta SPH kernel go to the column kernel.
  // Keys must fit the sort's 32-bit key; if they do not, everything stays with the bricks.
  g->route2 = ROUTE_BRICK;
  g->n_keys2 = 0;
  for (int i = 0; i < MTN_MAX_KERNELS; ++i) {
    g->kind[i] = (table && i < table->n) ? table->k[i].kind : -1;
    switch (g->kind[i]) {
      case MTN_KERNEL_WENDLANDC2: case MTN_KERNEL_WENDLANDC6: case MTN_KERNEL_CUBICSPLINE:
      case MTN_KERNEL_QUARTICSPLINE: g->support[i] = 1.0; break;
      case MTN_KERNEL_GAUSSIAN: g->support[i] = table->k[i].truncate * 0.42466090014400953; break;
      default: g->support[i] = INFINITY;
    }
  }
  static int streams = -1;  // MTN_STREAMS=0: developer switch, everything through the brick kernel
  if (streams < 0) {
    const char* e = getenv("MTN_STREAMS");
    streams = (e && atoi(e) == 0) ? 0 : 1;
  }
  if (table && streams) {
    const int64_t n_tiles = (int64_t)g->ntx * g->nty;
    const int64_t n_pix = (int64_t)(c->x_hi - c->x_lo) * c->ny;
    bool dirac_kernel = false;
    for (int i = 0; i < table->n; ++i) dirac_kernel |= table->k[i].kind == MTN_KERNEL_DIRACDELTA;
    if (c->spectrum == MTN_SPECTRUM_DIRACDELTA && n_tiles * c->n_channels < (1ll << 31)) {
      g->route2 = ROUTE_SPLAT;
      g->n_keys2 = n_tiles * c->n_channels;
    } else if (c->spectrum == MTN_SPECTRUM_GAUSSIAN && dirac_kernel && n_pix * g->nsb < (1ll << 31)) {
      g->route2 = ROUTE_COLUMN;
      g->n_keys2 = n_pix * g->nsb;
    }
  }
  return MTN_OK;
}

static PlanIn make_plan_in(const MtnParticles* p, const MtnCube* c) {
  PlanIn in;
  in.n = p->n;
  in.px = p->px;
  in.py = p->py;
  in.h_eff = p->h_eff;
  in.sm_range = p->sm_range;
  in.kernel_id = p->kernel_id;
  in.v = p->v;
  in.sigma = p->sigma;
  in.sigma_scalar = p->sigma_scalar;
  in.mHI = p->mHI;
  in.mHI_scalar = p->mHI_scalar;
  in.D = p->D;
  in.D_scalar = p->D_scalar;
  in.accept = p->accept;
  in.edges = c->edges;
  return in;
}

// plan scratch: [blk_kept | blk_pairs | blk_pairs2 | totals(8 x u64) | tile_sum | tile_cnt | tile_phase | feet]
struct PlanScratch {
  int64_t nblk;
  int64_t* blk_kept;
  int64_t* blk_pairs;
  int64_t* blk_pairs2;
  unsigned long long* totals;
  unsigned long long* tile_sum;
  unsigned int* tile_cnt;
  int* tile_phase;
  PackedFoot* feet;
};
static int64_t num_tiles(const MtnCube* c) {
  if (!c || c->x_hi <= c->x_lo || c->ny <= 0) return 1;
  return (int64_t)((c->x_hi - c->x_lo + TILE_X - 1) / TILE_X) * ((c->ny + TILE_Y - 1) / TILE_Y);
}
static size_t plan_scratch_layout(int64_t n, int64_t n_tiles, void* base, PlanScratch* s) {
  const int64_t nblk = std::max<int64_t>(1, (n + PLAN_THREADS - 1) / PLAN_THREADS);
  size_t off = 0;
  auto take = [&](size_t bytes) {
    char* p = base ? (char*)base + off : nullptr;
    off += align_up(bytes);
    return p;
  };
  char* a = take(nblk * sizeof(int64_t));
  char* b = take(nblk * sizeof(int64_t));
  char* b2 = take(nblk * sizeof(int64_t));
  char* t = take(8 * sizeof(unsigned long long));
  char* ts = take((size_t)n_tiles * sizeof(unsigned long long));
  char* tc = take((size_t)n_tiles * sizeof(unsigned int));
  char* tp = take((size_t)n_tiles * sizeof(int));
  char* ft = take((size_t)std::max<int64_t>(n, 1) * sizeof(PackedFoot));
  if (s) {
    s->nblk = nblk;
    s->blk_kept = (int64_t*)a;
    s->blk_pairs = (int64_t*)b;
    s->blk_pairs2 = (int64_t*)b2;
    s->totals = (unsigned long long*)t;
    s->tile_sum = (unsigned long long*)ts;
    s->tile_cnt = (unsigned int*)tc;
    s->tile_phase = (int*)tp;
    s->feet = (PackedFoot*)ft;
  }
  return off;
}

// project workspace.  A "stream" is one sorted pair array with its work items: the bricks
// (stream 0) and, if the insertion has one, the column or splat stream (stream 1).
struct StreamWs {
  uint64_t* pairs_a;
  uint64_t* pairs_b;
  uint32_t* key_count;  // per key: particles (brick_bounds writes one-past-last, item_count a count)
  uint32_t* key_start;
  uint32_t* counts;
  uint32_t* multi;
  uint32_t* ismulti;
  uint32_t* scalars;  // [n_items, n_slots, n_multi, counter, ...exec counts from +8]
  Item* items;
  MultiBrick* multis;
  double* partials;
  int64_t max_items, max_multi, max_slots;
};
struct Workspace {
  Record* records;
  double* inv_dv;
  uint32_t* hist;
  void* scan_temp;
  StreamWs s[2];
};
static size_t workspace_layout(int64_t n_kept, const int64_t n_pairs[2], const int64_t n_keys[2],
                               const int64_t chunk[2], const int64_t slot_doubles[2], int n_channels,
                               void* base, Workspace* w) {
  size_t off = 0;
  auto take = [&](size_t bytes) {
    char* p = base ? (char*)base + off : nullptr;
    off += align_up(bytes ? bytes : 1);
    return p;
  };
  Workspace ws;
  ws.records = (Record*)take((size_t)n_kept * sizeof(Record));
  ws.inv_dv = (double*)take((size_t)n_channels * sizeof(double));
  const int64_t np_max = std::max(n_pairs[0], n_pairs[1]), nk_max = std::max(n_keys[0], n_keys[1]);
  ws.hist = (uint32_t*)take(sort_hist_bytes(np_max));
  ws.scan_temp = take(std::max(scan_temp_bytes(sort_hist_entries(np_max), 4), scan_temp_bytes(nk_max, 4)));
  for (int k = 0; k < 2; ++k) {
    StreamWs& t = ws.s[k];
    const int64_t np = n_pairs[k], nk = np > 0 || k == 0 ? n_keys[k] : 0, ch = std::max<int64_t>(1, chunk[k]);
    t.max_multi = np / ch + 1;
    t.max_slots = 2 * (np / ch) + 2;
    t.max_items = std::min<int64_t>(nk, np) + np / ch + 1;
    t.pairs_a = (uint64_t*)take((size_t)np * 8);
    t.pairs_b = (uint64_t*)take((size_t)np * 8);
    t.key_count = (uint32_t*)take((size_t)nk * 4);
    t.key_start = (uint32_t*)take((size_t)nk * 4);
    t.counts = (uint32_t*)take((size_t)nk * 4);
    t.multi = (uint32_t*)take((size_t)nk * 4);
    t.ismulti = (uint32_t*)take((size_t)nk * 4);
    t.scalars = (uint32_t*)take(64);
    t.items = (Item*)take((size_t)t.max_items * sizeof(Item));
    t.multis = (MultiBrick*)take((size_t)t.max_multi * sizeof(MultiBrick));
    t.partials = (double*)take((size_t)t.max_slots * slot_doubles[k] * sizeof(double));
  }
  if (w) *w = ws;
  return off;
}

// The plan's sizes in the form workspace_layout wants them.
struct PlanSizes {
  int64_t n_pairs[2], n_keys[2], chunk[2], slot_doubles[2];
};
static PlanSizes plan_sizes(const MtnPlan* plan, const Geo& g) {
  PlanSizes z;
  z.n_pairs[0] = plan->n_pairs;
  z.n_pairs[1] = plan->n_pairs2;
  z.n_keys[0] = g.n_bricks;
  z.n_keys[1] = g.n_keys2;
  z.chunk[0] = plan->chunk;
  z.chunk[1] = plan->chunk2;
  z.slot_doubles[0] = (int64_t)TILE_PIX * CB;
  z.slot_doubles[1] = g.route2 == ROUTE_COLUMN ? CSB : TILE_PIX;
  return z;
}

static int64_t choose_chunk(int64_t n_pairs) {
  // enough work items for ~8 rounds over the resident CTAs, never smaller than 8 batches
  static int rounds = 0;
  if (!rounds) {  // MTN_ITEM_ROUNDS: tuning knob, work items per resident CTA
    const char* e = getenv("MTN_ITEM_ROUNDS");
    rounds = e ? std::max(1, atoi(e)) : 8;
  }
  const int64_t target = (int64_t)sm_count() * PROJ_CTAS_PER_SM * rounds;
  int64_t chunk = std::max<int64_t>(8 * PBATCH, (n_pairs + target - 1) / target);
  return (chunk + PBATCH - 1) / PBATCH * PBATCH;
}

// Work items of the column / splat kernels are taken by warps: ~8 rounds over the resident
// warps, never smaller than 8 batches.
static int64_t choose_chunk2(int64_t n_pairs) {
  const int64_t target = (int64_t)sm_count() * 16 * 8;
  int64_t chunk = std::max<int64_t>(8 * PBATCH, (n_pairs + target - 1) / target);
  return (chunk + PBATCH - 1) / PBATCH * PBATCH;
}

// FP64 FMA microbenchmark: 8 independent chains per thread, all in registers.
__global__ void __launch_bounds__(256) fp64_peak_kernel(double* out, int iters, double a, double b) {
  double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5,
         x6 = x0 + 6, x7 = x0 + 7;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
      x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
    }
  }
  const double s = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
  if (s == 123.456) out[0] = s;  // never true; keeps the chains alive
}

__global__ void __launch_bounds__(256) probe_kernel_integral_kernel(
    int kind, double truncate, double norm, int closed_form, int64_t n, const double* __restrict__ dx,
    const double* __restrict__ dy, const double* __restrict__ h, double* __restrict__ w) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double hh = h[i];
  w[i] = closed_form ? kernel_weight_closed(kind, dx[i], dy[i], hh, 1.0 / (hh * hh), truncate, norm)
                     : kernel_weight(kind, dx[i], dy[i], hh, 1.0 / (hh * hh), truncate, norm);
}

__global__ void __launch_bounds__(256) probe_spectra_kernel(
    int spectrum, int64_t n, const double* __restrict__ v, const double* __restrict__ sigma,
    double sigma_scalar, const double* __restrict__ amp, int C, const double* __restrict__ edges,
    double* __restrict__ out) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n * C) return;
  const int64_t i = idx / C;
  const int c = (int)(idx - i * C);
  const double e0 = edges[c], e1 = edges[c + 1];
  const double lo = fmin(e0, e1), hi = fmax(e0, e1);
  const double inv_dv = 1.0 / fabs(e1 - e0);
  double f;
  if (spectrum == MTN_SPECTRUM_GAUSSIAN) {
    const double inv_s = 1.0 / (1.4142135623730951 * (sigma ? sigma[i] : sigma_scalar));
    f = (edge_erf(hi, v[i], inv_s) - edge_erf(lo, v[i], inv_s)) * (0.5 * amp[i] * inv_dv);
  } else {
    f = dirac_channel(lo, hi, v[i]) * (amp[i] * inv_dv);
  }
  out[idx] = f;
}

}  // namespace mtn

using namespace mtn;

extern "C" {

int mtn_version(void) { return MTN_VERSION; }
const char* mtn_last_error(void) { return g_err; }
int mtn_last_launch_count(void) { return g_launches; }

int mtn_device_info(int* sms, int* cc_major, int* cc_minor) {
  int dev = 0;
  MTN_CUDA(cudaGetDevice(&dev));
  if (sms) MTN_CUDA(cudaDeviceGetAttribute(sms, cudaDevAttrMultiProcessorCount, dev));
  if (cc_major) MTN_CUDA(cudaDeviceGetAttribute(cc_major, cudaDevAttrComputeCapabilityMajor, dev));
  if (cc_minor) MTN_CUDA(cudaDeviceGetAttribute(cc_minor, cudaDevAttrComputeCapabilityMinor, dev));
  return MTN_OK;
}

int mtn_smoothing_setup(int64_t n, const double* sm_length, const MtnKernelTable* table,
                        uint8_t* kernel_id_out, uint8_t* valid_out, double* sm_range_out,
                        double* h_eff_out, void* stream) {
  KernelTableDev t;
  if (int rc = to_dev_table(table, &t)) return rc;
  if (n < 0 || (n > 0 && !sm_length)) return fail(MTN_ERR_INVALID, "smoothing_setup: bad input%s", "");
  if (n == 0) return MTN_OK;
  MTN_LAUNCH(smoothing_setup_kernel, (unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream, n,
             sm_length, t, kernel_id_out, valid_out, sm_range_out, h_eff_out);
  MTN_LAUNCH_CHECK();
  return MTN_OK;
}

int mtn_prune(int64_t n0, const double* px, const double* py, const double* pz,
              const double* sm_range, const double* mHI, double mHI_scalar,
              const double* half_width, double half_width_scalar, double max_abs_dv,
              int32_t nx_tot, int32_t ny_tot, int32_t n_channels, int32_t flags,
              uint8_t* accept_out, int64_t* n_accept_out, void* stream) {
  if (n0 < 0 || (n0 > 0 && !accept_out)) return fail(MTN_ERR_INVALID, "prune: bad input%s", "");
  cudaStream_t st = (cudaStream_t)stream;
  if (n_accept_out) MTN_CUDA(cudaMemsetAsync(n_accept_out, 0, sizeof(int64_t), st));
  if (n0 == 0) return MTN_OK;
  if ((flags & MTN_PRUNE_SPATIAL) && (!px || !py || !sm_range))
    return fail(MTN_ERR_INVALID, "prune: spatial pruning needs px, py, sm_range%s", "");
  if ((flags & MTN_PRUNE_SPECTRAL) && !pz)
    return fail(MTN_ERR_INVALID, "prune: spectral pruning needs pz%s", "");
  MTN_LAUNCH(prune_kernel, (unsigned)((n0 + 255) / 256), 256, 0, st, n0, px, py, pz, sm_range, mHI,
             mHI_scalar, half_width, half_width_scalar, max_abs_dv, (double)nx_tot, (double)ny_tot,
             (double)n_channels, flags, accept_out, (unsigned long long*)n_accept_out);
  MTN_LAUNCH_CHECK();
  return MTN_OK;
}

size_t mtn_plan_scratch_bytes(int64_t n, const MtnCube* cube) {
  return plan_scratch_layout(n, num_tiles(cube), nullptr, nullptr);
}

static int check_particles(const MtnParticles* p) {
  if (!p || p->n < 0) return fail(MTN_ERR_INVALID, "particles: bad n%s", "");
  if (p->n > 0 && (!p->px || !p->py || !p->h_eff || !p->sm_range || !p->v))
    return fail(MTN_ERR_INVALID, "particles: px, py, h_eff, sm_range and v are required%s", "");
  return MTN_OK;
}

static int edges_direction(const MtnCube* cube, cudaStream_t st, int* increasing) {
  // the host mirror validates monotonicity (spectral_models.py:180-187); here only the
  // direction is needed.  A caller that knows it says so (no device read-back, no sync).
  if (cube->edges_direction > 0 || cube->edges_direction < 0) {
    *increasing = cube->edges_direction > 0 ? 1 : 0;
    return MTN_OK;
  }
  double e[2];
  MTN_CUDA(cudaMemcpyAsync(e, cube->edges, sizeof(e), cudaMemcpyDeviceToHost, st));
  MTN_CUDA(cudaStreamSynchronize(st));
  *increasing = e[1] > e[0] ? 1 : 0;
  return MTN_OK;
}

int mtn_plan(const MtnParticles* p, const MtnKernelTable* table, const MtnCube* cube, void* scratch,
             size_t scratch_bytes, MtnPlan* plan, void* stream) {
  if (int rc = check_particles(p)) return rc;
  if (!cube || !cube->edges || !plan || !table) return fail(MTN_ERR_INVALID, "plan: bad arguments%s", "");
  cudaStream_t st = (cudaStream_t)stream;
  PlanScratch ps;
  if (plan_scratch_layout(p->n, num_tiles(cube), scratch, &ps) > scratch_bytes || !scratch)
    return fail(MTN_ERR_WORKSPACE, "plan: scratch too small%s", "");
  int inc = 0;
  if (int rc = edges_direction(cube, st, &inc)) return rc;
  Geo g;
  if (int rc = make_geo(cube, table, &g, inc)) return rc;
  if (cube->n_channels > 65535) return fail(MTN_ERR_LIMIT, "plan: more than 65535 channels%s", "");
  g.phase = ps.tile_phase;
  const int n_tiles = g.ntx * g.nty;
  MTN_CUDA(cudaMemsetAsync(ps.totals, 0, 8 * sizeof(unsigned long long), st));
  MTN_CUDA(cudaMemsetAsync(ps.tile_sum, 0, (size_t)n_tiles * sizeof(unsigned long long), st));
  MTN_CUDA(cudaMemsetAsync(ps.tile_cnt, 0, (size_t)n_tiles * sizeof(unsigned int), st));
  if (p->n > 0 && g.route2 != ROUTE_SPLAT) {  // (the splat stream has no channel blocks to phase)
    MTN_LAUNCH(tile_stats_kernel, (unsigned)((ps.nblk + TILE_STAT_STRIDE - 1) / TILE_STAT_STRIDE),
               PLAN_THREADS, 0, st, make_plan_in(p, cube), g, ps.tile_sum, ps.tile_cnt);
    MTN_LAUNCH_CHECK();
  }
  MTN_LAUNCH(tile_phase_kernel, (unsigned)((n_tiles + 255) / 256), 256, 0, st, n_tiles, ps.tile_sum,
             ps.tile_cnt, ps.tile_phase);
  MTN_LAUNCH_CHECK();
  if (p->n > 0) {
    MTN_LAUNCH(plan_count_kernel, (unsigned)ps.nblk, PLAN_THREADS, 0, st, make_plan_in(p, cube), g,
               ps.blk_kept, ps.blk_pairs, ps.blk_pairs2, ps.totals + 3, ps.feet);
    MTN_LAUNCH_CHECK();
    MTN_LAUNCH(scan3_sums_inplace, 3, 1024, 0, st, ps.blk_kept, ps.blk_pairs, ps.blk_pairs2, ps.nblk,
               (int64_t*)ps.totals);
    MTN_LAUNCH_CHECK();
  }
  unsigned long long tot[4];
  MTN_CUDA(cudaMemcpyAsync(tot, ps.totals, sizeof(tot), cudaMemcpyDeviceToHost, st));
  MTN_CUDA(cudaStreamSynchronize(st));
  plan->n_kept = (int64_t)tot[0];
  plan->n_pairs = (int64_t)tot[1];
  plan->n_pairs2 = (int64_t)tot[2];
  plan->updates_dense = (int64_t)tot[3];
  plan->n_bricks = g.n_bricks;
  plan->route2 = g.route2;
  if (std::max(plan->n_pairs, plan->n_pairs2) >= (1ll << 32) - 1 || plan->n_kept >= (1ll << 32) - 1)
    return fail(MTN_ERR_LIMIT, "plan: %s%lld pairs exceed the 32-bit sort index; split the slab", "",
                (long long)std::max(plan->n_pairs, plan->n_pairs2));
  plan->chunk = choose_chunk(plan->n_pairs);
  plan->chunk2 = choose_chunk2(plan->n_pairs2);
  plan->edges_increasing = inc;
  const PlanSizes z = plan_sizes(plan, g);
  plan->workspace_bytes = workspace_layout(plan->n_kept, z.n_pairs, z.n_keys, z.chunk, z.slot_doubles,
                                           cube->n_channels, nullptr, nullptr);
  return MTN_OK;
}

// Sort one stream's pairs by key and cut the sorted runs into work items.
static int build_items(StreamWs& t, int64_t n_pairs, int64_t n_keys, int64_t chunk, uint32_t* hist,
                       void* scan_temp, uint64_t** sorted, cudaStream_t st) {
  int key_bits = 1;
  while ((1ll << key_bits) < n_keys) ++key_bits;
  if (int rc = radix_sort_pairs(t.pairs_a, t.pairs_b, n_pairs, key_bits, hist, scan_temp, sorted, st)) return rc;
  mark(2, st);
  MTN_LAUNCH(brick_bounds_kernel, (unsigned)((n_pairs + 255) / 256), 256, 0, st, *sorted, n_pairs, t.key_start,
             t.key_count);
  MTN_LAUNCH_CHECK();
  const unsigned bgrid = (unsigned)((n_keys + 255) / 256);
  MTN_LAUNCH(item_count_kernel, bgrid, 256, 0, st, t.key_count, t.key_start, (int)n_keys, (uint32_t)chunk,
             t.counts, t.multi, t.ismulti);
  MTN_LAUNCH_CHECK();
  if (int rc = exclusive_scan<uint32_t, uint32_t>(t.counts, t.counts, n_keys, scan_temp, t.scalars + 0, st)) return rc;
  if (int rc = exclusive_scan<uint32_t, uint32_t>(t.multi, t.multi, n_keys, scan_temp, t.scalars + 1, st)) return rc;
  if (int rc = exclusive_scan<uint32_t, uint32_t>(t.ismulti, t.ismulti, n_keys, scan_temp, t.scalars + 2, st))
    return rc;
  MTN_LAUNCH(item_fill_kernel, bgrid, 256, 0, st, t.key_count, t.key_start, (int)n_keys, (uint32_t)chunk, t.counts,
             t.multi, t.ismulti, t.items, t.multis);
  MTN_LAUNCH_CHECK();
  return MTN_OK;
}

int mtn_project(const MtnParticles* p, const MtnKernelTable* table, const MtnCube* cube,
                const MtnPlan* plan, void* scratch, size_t scratch_bytes, void* workspace,
                size_t workspace_bytes, void* stream) {
  g_launches = 0;
  if (int rc = ensure_tables()) return rc;
  if (int rc = check_particles(p)) return rc;
  if (!cube || !cube->edges || !cube->slab || !plan)
    return fail(MTN_ERR_INVALID, "project: bad arguments%s", "");
  if (!(cube->px_size_arcsec > 0.0)) return fail(MTN_ERR_INVALID, "project: px_size must be > 0%s", "");
  KernelTableDev t;
  if (int rc = to_dev_table(table, &t)) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  PlanScratch ps;
  if (plan_scratch_layout(p->n, num_tiles(cube), scratch, &ps) > scratch_bytes || !scratch)
    return fail(MTN_ERR_WORKSPACE, "project: scratch too small%s", "");
  Geo g;
  if (int rc = make_geo(cube, table, &g, plan->edges_increasing)) return rc;
  if (g.n_bricks != plan->n_bricks || g.route2 != plan->route2)
    return fail(MTN_ERR_INVALID, "project: plan/cube mismatch%s", "");
  g.phase = ps.tile_phase;  // written by mtn_plan into the caller's scratch
  const PlanSizes z = plan_sizes(plan, g);
  Workspace ws;
  if (workspace_layout(plan->n_kept, z.n_pairs, z.n_keys, z.chunk, z.slot_doubles, cube->n_channels, workspace,
                       &ws) > workspace_bytes ||
      !workspace)
    return fail(MTN_ERR_WORKSPACE, "project: workspace too small%s", "");
  const double px_area = cube->px_size_arcsec * cube->px_size_arcsec;
  const int zeroed = (cube->flags & MTN_CUBE_ZEROED) ? 1 : 0;
  const bool stream2 = plan->n_pairs2 > 0;

  MTN_CUDA(cudaMemsetAsync(ws.s[0].key_count, 0, (size_t)g.n_bricks * 4, st));
  MTN_CUDA(cudaMemsetAsync(ws.s[0].scalars, 0, 64, st));
  if (stream2) {
    MTN_CUDA(cudaMemsetAsync(ws.s[1].key_count, 0, (size_t)g.n_keys2 * 4, st));
    MTN_CUDA(cudaMemsetAsync(ws.s[1].scalars, 0, 64, st));
  }
  g_ev_valid = false;
  mark(0, st);
  for (int k = 1; k <= N_STAGES; ++k) mark(k, st);  // stages skipped below read as 0 ms
  for (int k = 0; k < 3; ++k) g_exec_counts[k] = 0;

  if (plan->n_pairs + plan->n_pairs2 > 0) {
    MTN_LAUNCH(plan_emit_kernel, (unsigned)ps.nblk, PLAN_THREADS, 0, st, make_plan_in(p, cube), g,
               ps.blk_kept, ps.blk_pairs, ps.blk_pairs2, ps.feet, ws.records, ws.s[0].pairs_a, ws.s[1].pairs_a);
    MTN_LAUNCH_CHECK();
  }
  mark(1, st);
  uint64_t* sorted[2] = {nullptr, nullptr};
  if (plan->n_pairs > 0)
    if (int rc = build_items(ws.s[0], plan->n_pairs, g.n_bricks, plan->chunk, ws.hist, ws.scan_temp, &sorted[0], st))
      return rc;
  if (stream2) {
    if (int rc = build_items(ws.s[1], plan->n_pairs2, g.n_keys2, plan->chunk2, ws.hist, ws.scan_temp, &sorted[1], st))
      return rc;
    MTN_LAUNCH(inv_dv_kernel, (unsigned)((g.C + 255) / 256), 256, 0, st, cube->edges, g.C, ws.inv_dv);
    MTN_LAUNCH_CHECK();
  }
  mark(3, st);
  if (plan->n_pairs > 0) {
    ProjArgs a;
    a.geo = g;
    a.table = t;
    a.records = ws.records;
    a.pairs = sorted[0];
    a.items = ws.s[0].items;
    a.n_items = ws.s[0].scalars + 0;
    a.counter = ws.s[0].scalars + 3;
    a.edges = cube->edges;
    a.slab = cube->slab;
    a.partials = ws.s[0].partials;
    a.px_area = px_area;
    a.zeroed = zeroed;
    a.exec_counts = (unsigned long long*)(ws.s[0].scalars + 8);  // zeroed with the scalars
    // the instantiation specialised on the kind of table entry 0 (if that kind is tabulated)
    if (int rc = launch_project(a, t.kind[0], g_count_exec != 0, ws.s[0].max_items, st)) return rc;
    MTN_LAUNCH_CHECK();
    mark(4, st);
    MTN_LAUNCH(reduce_partials_kernel, dim3((unsigned)ws.s[0].max_multi, SUB_PIX), PROJ_THREADS, 0, st, g,
               ws.s[0].multis, ws.s[0].scalars + 2, ws.s[0].partials, cube->slab, px_area, zeroed);
    MTN_LAUNCH_CHECK();
  } else {
    mark(4, st);
  }
  mark(5, st);
  if (!zeroed) {  // voxels of bricks no brick-kernel particle reaches still get in / px_area
    MTN_LAUNCH(empty_brick_kernel, (unsigned)g.n_bricks, PROJ_THREADS, 0, st, g, ws.s[0].key_count, cube->slab,
               px_area);
    MTN_LAUNCH_CHECK();
  }
  mark(6, st);
  if (stream2) {  // the column / splat kernel adds onto what the brick kernel has written
    StreamArgs a;
    a.geo = g;
    a.table = t;
    a.records = ws.records;
    a.pairs = sorted[1];
    a.items = ws.s[1].items;
    a.n_items = ws.s[1].scalars + 0;
    a.counter = ws.s[1].scalars + 3;
    a.edges = cube->edges;
    a.inv_dv = ws.inv_dv;
    a.slab = cube->slab;
    a.partials = ws.s[1].partials;
    a.px_area = px_area;
    a.exec_counts = (unsigned long long*)(ws.s[1].scalars + 8);
    if (int rc = launch_stream(a, g.route2, g_count_exec != 0, ws.s[1].max_items, st)) return rc;
    MTN_LAUNCH_CHECK();
    if (g.route2 == ROUTE_COLUMN) {
      MTN_LAUNCH(column_reduce_kernel, (unsigned)ws.s[1].max_multi, 256, 0, st, g, ws.s[1].multis,
                 ws.s[1].scalars + 2, ws.s[1].partials, cube->slab, px_area);
    } else {
      MTN_LAUNCH(splat_reduce_kernel, (unsigned)ws.s[1].max_multi, TILE_PIX, 0, st, g, ws.s[1].multis,
                 ws.s[1].scalars + 2, ws.s[1].partials, cube->slab, px_area);
    }
    MTN_LAUNCH_CHECK();
  }
  if (g_count_exec) {
    unsigned long long c[2][3] = {{0, 0, 0}, {0, 0, 0}};
    if (plan->n_pairs > 0)
      MTN_CUDA(cudaMemcpyAsync(c[0], ws.s[0].scalars + 8, sizeof(c[0]), cudaMemcpyDeviceToHost, st));
    if (stream2) MTN_CUDA(cudaMemcpyAsync(c[1], ws.s[1].scalars + 8, sizeof(c[1]), cudaMemcpyDeviceToHost, st));
    MTN_CUDA(cudaStreamSynchronize(st));
    for (int k = 0; k < 3; ++k) g_exec_counts[k] = c[0][k] + c[1][k];
  }
  mark(7, st);
  g_ev_valid = g_timing != 0;
  return MTN_OK;
}

static int fill_route_args(RouteArgs* a, int64_t n, const double* px, const double* sm_range, int32_t world,
                           const int32_t* bounds) {
  if (n < 0 || world < 1 || world > ROUTE_MAX_WORLD || !bounds || (n > 0 && (!px || !sm_range)))
    return fail(MTN_ERR_INVALID, "route: bad arguments (world <= 16)%s", "");
  a->n = n;
  a->px = px;
  a->sm_range = sm_range;
  a->world = world;
  for (int d = 0; d <= world; ++d) a->bounds[d] = bounds[d];
  return MTN_OK;
}

size_t mtn_route_scratch_bytes(int64_t n, int32_t world) {
  const int64_t nblk = std::max<int64_t>(1, (n + ROUTE_THREADS - 1) / ROUTE_THREADS);
  return align_up((size_t)world * nblk * 4) + scan_temp_bytes(nblk, 4) + align_up((size_t)world * 4 * 2);
}

int mtn_route_count(int64_t n, const double* px, const double* sm_range, int32_t world, const int32_t* bounds,
                    void* scratch, size_t scratch_bytes, int64_t* totals_out, void* stream) {
  RouteArgs a;
  if (int rc = fill_route_args(&a, n, px, sm_range, world, bounds)) return rc;
  if (!scratch || scratch_bytes < mtn_route_scratch_bytes(n, world) || !totals_out)
    return fail(MTN_ERR_WORKSPACE, "route: scratch too small%s", "");
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t nblk = std::max<int64_t>(1, (n + ROUTE_THREADS - 1) / ROUTE_THREADS);
  uint32_t* blk = (uint32_t*)scratch;
  void* scan_temp = (char*)scratch + align_up((size_t)world * nblk * 4);
  uint32_t* tot32 = (uint32_t*)((char*)scan_temp + scan_temp_bytes(nblk, 4));
  MTN_LAUNCH(route_count_kernel, (unsigned)nblk, ROUTE_THREADS, 0, st, a, nblk, blk);
  MTN_LAUNCH_CHECK();
  for (int d = 0; d < world; ++d)  // block counts -> exclusive block offsets, totals beside them
    if (int rc = exclusive_scan<uint32_t, uint32_t>(blk + d * nblk, blk + d * nblk, nblk, scan_temp, tot32 + d, st))
      return rc;
  MTN_LAUNCH(widen_totals_kernel, 1, 32, 0, st, tot32, world, totals_out);
  MTN_LAUNCH_CHECK();
  return MTN_OK;
}

int mtn_route_scatter(int64_t n, const double* px, const double* sm_range, int32_t world, const int32_t* bounds,
                      int32_t n_fields, const double* const* fields, double* const* inboxes, int64_t capacity,
                      const int64_t* src_offsets, const void* scratch, void* stream) {
  RouteArgs a;
  if (int rc = fill_route_args(&a, n, px, sm_range, world, bounds)) return rc;
  if (n_fields < 1 || n_fields > ROUTE_MAX_FIELDS || !fields || !inboxes || !src_offsets || !scratch || capacity < 0)
    return fail(MTN_ERR_INVALID, "route: bad scatter arguments (at most 12 fields)%s", "");
  if (n == 0) return MTN_OK;
  const int64_t nblk = std::max<int64_t>(1, (n + ROUTE_THREADS - 1) / ROUTE_THREADS);
  ScatterArgs s;
  s.n_fields = n_fields;
  for (int f = 0; f < ROUTE_MAX_FIELDS; ++f) s.src[f] = f < n_fields ? fields[f] : nullptr;
  for (int d = 0; d < ROUTE_MAX_WORLD; ++d) s.dst[d] = d < world ? inboxes[d] : nullptr;
  s.capacity = capacity;
  s.blk_off = (const uint32_t*)scratch;
  s.src_off = src_offsets;
  MTN_LAUNCH(route_scatter_kernel, (unsigned)nblk, ROUTE_THREADS, 0, (cudaStream_t)stream, a, s, nblk);
  MTN_LAUNCH_CHECK();
  return MTN_OK;
}

int mtn_set_timing(int enable) {
  g_timing = enable ? 1 : 0;
  return MTN_OK;
}

int mtn_set_count_exec(int enable) {
  g_count_exec = enable ? 1 : 0;
  return MTN_OK;
}

int mtn_last_exec_counts(int64_t* out3) {
  if (!out3) return fail(MTN_ERR_INVALID, "exec_counts: null output%s", "");
  for (int i = 0; i < 3; ++i) out3[i] = (int64_t)g_exec_counts[i];
  return MTN_OK;
}

int mtn_last_timing(float* ms_out, int n) {
  if (!ms_out || n < N_STAGES) return fail(MTN_ERR_INVALID, "timing: need room for 7 stages%s", "");
  if (!g_ev_valid) return fail(MTN_ERR_INVALID, "timing: enable with mtn_set_timing first%s", "");
  MTN_CUDA(cudaEventSynchronize(g_ev[N_STAGES]));
  for (int k = 0; k < N_STAGES; ++k) MTN_CUDA(cudaEventElapsedTime(ms_out + k, g_ev[k], g_ev[k + 1]));
  return MTN_OK;
}

int mtn_fp64_peak(double* tflops_out, double* ms_out, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  double* d = nullptr;
  MTN_CUDA(cudaMalloc(&d, 8));
  const int iters = 4096, blocks = sm_count() * 16;
  cudaEvent_t e0, e1;
  MTN_CUDA(cudaEventCreate(&e0));
  MTN_CUDA(cudaEventCreate(&e1));
  float best = 1e30f;
  for (int rep = 0; rep < 4; ++rep) {
    MTN_CUDA(cudaEventRecord(e0, st));
    MTN_LAUNCH(fp64_peak_kernel, blocks, 256, 0, st, d, iters, 0.999999, 1e-7);
    MTN_CUDA(cudaEventRecord(e1, st));
    MTN_CUDA(cudaEventSynchronize(e1));
    float ms = 0;
    MTN_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    if (rep > 0 && ms < best) best = ms;
  }
  MTN_CUDA(cudaGetLastError());
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(d);
  const double flops = 2.0 * 64.0 * iters * 256.0 * blocks;
  if (tflops_out) *tflops_out = flops / (best * 1e-3) / 1e12;
  if (ms_out) *ms_out = best;
  return MTN_OK;
}

int mtn_convolve_beam(const double* cube_in, double* cube_out, int32_t nx, int32_t ny, int32_t nc,
                      const double* kernel, int32_t ka, int32_t kb, double scale, void* stream) {
  if (!cube_in || !cube_out || !kernel || cube_in == cube_out)
    return fail(MTN_ERR_INVALID, "convolve_beam: bad pointers (in-place is not supported)%s", "");
  if (nx <= 0 || ny <= 0 || nc <= 0 || ka <= 0 || kb <= 0 || !(ka & 1) || !(kb & 1))
    return fail(MTN_ERR_INVALID, "convolve_beam: bad shape (the beam image must be odd x odd)%s", "");
  if ((int64_t)ka * kb > CONV_MAX_TAPS)
    return fail(MTN_ERR_LIMIT, "convolve_beam: beam image of %s%lld taps exceeds the 28000 (e.g. 167 x 167) that fit "
                "in shared memory", "", (long long)ka * kb);
  const size_t smem = (size_t)ka * kb * sizeof(double);
  static bool attr_set[MAX_DEVICES] = {false};
  {
    std::lock_guard<std::mutex> lock(g_once_mutex);
    const int dev = current_device();
    if (!attr_set[dev]) {
      MTN_CUDA(cudaFuncSetAttribute(convolve_beam_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)(CONV_MAX_TAPS * sizeof(double))));
      attr_set[dev] = true;
    }
  }
  const dim3 grid((nc + 31) / 32, (ny + CONV_WARPS * CONV_TY - 1) / (CONV_WARPS * CONV_TY), (nx + CONV_TX - 1) / CONV_TX);
  MTN_LAUNCH(convolve_beam_kernel, grid, CONV_WARPS * 32, smem, (cudaStream_t)stream, cube_in, cube_out, nx, ny, nc,
             kernel, ka, kb, scale);
  MTN_LAUNCH_CHECK();
  return MTN_OK;
}

int mtn_sky_to_pix(const MtnFrontEnd* fe, int64_t n, const double* xyz, const double* vxyz, const double* hsm,
                   double hsm_scalar, double* px, double* py, double* pz, double* v, double* D, double* sm_length,
                   void* stream) {
  if (!fe || n < 0 || (n > 0 && (!xyz || !vxyz || !px || !py || !pz || !v || !D || !sm_length)))
    return fail(MTN_ERR_INVALID, "sky_to_pix: bad arguments%s", "");
  if (!(fe->px_size_arcsec > 0.0) || !(fe->channel_width > 0.0))
    return fail(MTN_ERR_INVALID, "sky_to_pix: px_size and channel_width must be > 0%s", "");
  if (n == 0) return MTN_OK;
  FrontEndArgs a;
  a.n = n;
  a.xyz = xyz;
  a.vxyz = vxyz;
  a.hsm = hsm;
  a.hsm_scalar = hsm_scalar;
  for (int k = 0; k < 9; ++k) a.R[k] = fe->rotation[k];
  for (int k = 0; k < 3; ++k) a.unit[k] = fe->direction[k];
  a.distance_kpc = fe->distance_mpc * 1.0e3;
  a.vpeculiar = fe->vpeculiar;
  a.hubble = fe->hubble;
  a.sin_d0 = sin(fe->dec0_rad);
  a.cos_d0 = cos(fe->dec0_rad);
  a.a0 = fe->ra0_rad;
  a.rad_to_px = 1.0 / (fe->px_size_arcsec * (3.14159265358979323846 / 648000.0));
  a.crpix_x = fe->crpix[0];
  a.crpix_y = fe->crpix[1];
  a.crpix_z = fe->crpix[2];
  a.freq_mode = fe->freq_mode;
  a.spectral_centre = fe->spectral_centre;
  a.channel_width = fe->channel_width;
  a.px = px;
  a.py = py;
  a.pz = pz;
  a.v = v;
  a.D = D;
  a.sm_length = sm_length;
  MTN_LAUNCH(sky_to_pix_kernel, (unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream, a);
  MTN_LAUNCH_CHECK();
  return MTN_OK;
}

int mtn_table_error(int32_t kind, double* err_out) {
  if (kind < 0 || kind >= WT_KINDS || !err_out) return fail(MTN_ERR_INVALID, "table_error: bad kind%s", "");
  if (int rc = ensure_tables()) return rc;
  *err_out = g_table_err[kind];
  return MTN_OK;
}

int mtn_probe_kernel_integral(const MtnKernelEntry* entry, int32_t closed_form, int64_t n,
                              const double* dx, const double* dy, const double* h, double* w_out,
                              void* stream) {
  if (!entry || n < 0) return fail(MTN_ERR_INVALID, "probe: bad arguments%s", "");
  if (int rc = ensure_tables()) return rc;
  if (n == 0) return MTN_OK;
  MTN_LAUNCH(probe_kernel_integral_kernel, (unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream,
             entry->kind, entry->truncate, entry->norm, closed_form, n, dx, dy, h, w_out);
  MTN_LAUNCH_CHECK();
  return MTN_OK;
}

int mtn_probe_spectra(int32_t spectrum, int64_t n, const double* v, const double* sigma,
                      double sigma_scalar, const double* amp, int32_t n_channels,
                      const double* edges, double* s_out, void* stream) {
  if (n < 0 || n_channels <= 0) return fail(MTN_ERR_INVALID, "probe: bad arguments%s", "");
  if (n == 0) return MTN_OK;
  if (int rc = ensure_tables()) return rc;
  const int64_t tot = n * n_channels;
  MTN_LAUNCH(probe_spectra_kernel, (unsigned)((tot + 255) / 256), 256, 0, (cudaStream_t)stream, spectrum, n,
             v, sigma, sigma_scalar, amp, n_channels, edges, s_out);
  MTN_LAUNCH_CHECK();
  return MTN_OK;
}

}  // extern "C"
