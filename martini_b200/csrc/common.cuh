// Shared definitions for the martini_b200 CUDA library (sm_100a).
#pragma once

#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

#include <cstdio>
#include <cstring>

#include "../../include/martini_b200.h"

namespace mtn {

// ---------------------------------------------------------------------------------------
// Brick geometry.  A brick is TILE_X x TILE_Y pixels x CB channels; it is the unit of
// binning (sort key) and the unit of register accumulation in the projection kernel.
// ---------------------------------------------------------------------------------------
constexpr int TILE_X = 8;             // pixels along x (slowest cube axis)
constexpr int TILE_Y = 8;             // pixels along y
constexpr int TILE_PIX = TILE_X * TILE_Y;
constexpr int CB = 64;                // channels per brick (unit of binning): each lane owns 2 adjacent ones
constexpr int SUB_X = 4;              // a warp owns a SUB_X x SUB_Y pixel sub-block of the tile
constexpr int SUB_Y = 4;
constexpr int SUB_PIX = SUB_X * SUB_Y;  // = accumulator pairs per thread
constexpr int SUBS_Y = TILE_Y / SUB_Y;  // sub-blocks per tile row
constexpr int N_SUB = TILE_PIX / SUB_PIX;      // sub-blocks per tile
constexpr int PROJ_WARPS = N_SUB;              // warp = sub-block
constexpr int PROJ_THREADS = PROJ_WARPS * 32;
// resident CTAs per SM: register-limited (16 warps per SM at 128 registers)
#ifndef MTN_PROJ_CTAS
#define MTN_PROJ_CTAS (512 / PROJ_THREADS)
#endif
constexpr int PROJ_CTAS_PER_SM = MTN_PROJ_CTAS;
constexpr int PBATCH = 32;            // particle records staged per batch (= one per lane)
static_assert(PBATCH == 32, "the batch is indexed by lane in several places");

// erf table: Taylor coefficients of erf at the centres of ERF_NINT intervals of width
// 1/ERF_INV_W covering [0, ERF_SAT]; degree ERF_DEG.  Truncation error < 5e-18.
#ifndef MTN_ERF_INV_W
#define MTN_ERF_INV_W 16
#endif
#ifndef MTN_ERF_DEG
#define MTN_ERF_DEG 9
#endif
constexpr int ERF_INV_W = MTN_ERF_INV_W;
constexpr int ERF_DEG = MTN_ERF_DEG;
static_assert(ERF_DEG % 2 == 1, "rows are read as pairs of coefficients");
constexpr int ERF_NCOEF = ERF_DEG + 1;  // 10 doubles = 80 B per interval (16-B aligned rows)
constexpr int ERF_NINT = 6 * ERF_INV_W + 1;
// The compact erf table of the column / splat kernels, which keep it in shared memory and are
// bound by the shared-memory data pipe (every lane reads its own row): 1/64-wide intervals,
// and per interval only erf(x0) and A = 2/sqrt(pi) exp(-x0^2) at its centre x0 -- ONE 16-byte
// load per evaluation.  The Taylor coefficients of degree 2..5 follow from x0 and A in registers
// (erf^(n)(x) = (-1)^(n-1) A H_(n-1)(x), H = physicists' Hermite polynomials): nine FP64
// instructions on a pipe that was a quarter busy, against two more 16-byte loads per lane on
// the pipe that was 96 % busy (rounds 1-2 stored all six coefficients: 48-byte rows, 36 of the
// column kernel's 57 shared-memory wavefronts per particle).  Truncation as before:
// (1/128)^6 |erf^(6)| / 720 < 2e-14, five orders below what the 1e-9 flux tolerance needs.
constexpr int ERFC_INV_W = 64;
constexpr int ERFC_NCOEF = 2;  // {erf(x0), 2/sqrt(pi) exp(-x0^2)}
constexpr int ERFC_NINT = 6 * ERFC_INV_W + 1;
constexpr int ERFC_DOUBLES = ERFC_NCOEF * ERFC_NINT;
// One staged particle record: 80 bytes (common.cuh: Record), 16-B aligned so a single
// cp.async.bulk moves it.
constexpr int REC_DOUBLES = 10;
constexpr int REC_BYTES = REC_DOUBLES * 8;

// erf(x) == 1.0 exactly for x >= ~5.93 (1 - erf(x) < 2^-54); channels whose edges are
// both beyond ERF_SAT on the same side of the line centre contribute exactly 0 in the
// reference (scipy) and here, so skipping them is bit-safe.
constexpr double ERF_SAT = 6.0;

// Where a particle's contributions are computed (plan.cuh: footprint).  The general case is
// the brick kernel (project.cuh).  Two degenerate shapes get kernels of their own
// (streams.cuh), each with the lanes of a warp laid along the one axis that is left:
//   COLUMN  DiracDelta SPH kernel (sph_kernels.py:1136-1165): the particle reaches exactly one
//           pixel, so it is a line profile added to one voxel column -- key = (pixel, channel
//           superblock), lanes = channels;
//   SPLAT   DiracDelta spectrum (spectral_models.py:510-570): the particle reaches one channel
//           (two if it sits on a channel edge), so it is a kernel image added to one channel
//           map -- key = (tile, channel), lanes = pixels.
enum Route { ROUTE_BRICK = 0, ROUTE_COLUMN = 1, ROUTE_SPLAT = 2 };
constexpr int CSB = 1024;  // channels per superblock of the column kernel (8 KB of shared memory per warp)

struct Geo {
  int nx, ny, C;        // full cube
  int x_lo, x_hi;       // slab rows
  int ntx, nty, ncb;    // bricks along x (slab), y, channel (ncb = ceil(C/CB) + 1, see phase)
  int n_bricks;
  const int* phase;     // per-tile channel phase in [0, CB): block k = [ph + (k-1) CB, ph + k CB)
  int spectrum;         // MTN_SPECTRUM_*
  int edges_increasing; // 1 if edges[c+1] > edges[c]
  int nsb;              // channel superblocks of the column kernel: ceil(C / CSB)
  int route2;           // ROUTE_COLUMN / ROUTE_SPLAT: the second stream of this insertion; 0: none
  int64_t n_keys2;      // keys of the second stream
  int kind[MTN_MAX_KERNELS];  // MTN_KERNEL_* of each kernel-table entry
  // radius of the kernel's support in units of h_eff: the pixel integral is exactly zero from
  // there on (R >= 1 for the Wendland / spline kernels, sph_kernels.py:436, 683, 852, 1561;
  // d / h / sigma >= truncate for the Gaussian, :1030); +inf: no culling
  double support[MTN_MAX_KERNELS];
};

// One staged particle: everything the projection kernel needs.  The footprint (candidate box
// of martini.py:272-274 clipped to the slab, live channel window) is computed once, by the plan
// kernels that need it anyway, with the exact predicates; the projection kernel only clips it
// to the brick with integer arithmetic.
struct __align__(16) Record {
  double px, py;     // pixel coordinates
  double h;          // h_eff = sm_length * rescale
  double inv_h2;     // 1 / (h*h)
  double v;          // line centre [km/s]
  double inv_s;      // 1 / (sqrt(2) * sigma)   (Gaussian spectrum)
  double amp;        // mHI * D^-2 / 2.36e5  [Jy km/s]
  int32_t i0, i1, j0, j1;     // candidate box clipped to the slab, inclusive (plan.cuh: Foot)
  uint16_t c_first, c_last;   // live channel window, inclusive (plan.cuh: channel_window)
  uint8_t kid;                // kernel table index
  uint8_t pad[3];
};
static_assert(sizeof(Record) == REC_BYTES, "record size");
static_assert(REC_BYTES % 16 == 0, "cp.async.bulk moves multiples of 16 bytes");

// A unit of work for the projection kernel: a contiguous run of one brick's sorted pairs.
struct __align__(16) Item {
  uint32_t begin, end;  // range in the sorted pair array
  uint32_t brick;       // brick key
  int32_t slot;         // >= 0: write partial sums to partials[slot]; -1: write the cube
};

struct KernelTableDev {
  int n;
  int adaptive;
  int kind[MTN_MAX_KERNELS];
  int valid_is_max[MTN_MAX_KERNELS];
  double rescale[MTN_MAX_KERNELS];
  double size_in_fwhm[MTN_MAX_KERNELS];
  double valid_size[MTN_MAX_KERNELS];
  double truncate[MTN_MAX_KERNELS];
  double norm[MTN_MAX_KERNELS];
};

// ---------------------------------------------------------------------------------------
// error plumbing
// ---------------------------------------------------------------------------------------
extern thread_local char g_err[512];
extern thread_local int g_launches;

inline int fail(int code, const char* fmt, const char* a = "", long long b = 0) {
  snprintf(g_err, sizeof(g_err), fmt, a, b);
  return code;
}

#define MTN_CUDA(call)                                                                   \
  do {                                                                                   \
    cudaError_t e_ = (call);                                                             \
    if (e_ != cudaSuccess) {                                                             \
      snprintf(mtn::g_err, sizeof(mtn::g_err), "%s:%d: %s: %s", __FILE__, __LINE__, #call, \
               cudaGetErrorString(e_));                                                  \
      return MTN_ERR_CUDA;                                                               \
    }                                                                                    \
  } while (0)

#define MTN_LAUNCH_CHECK()                \
  do {                                    \
    ++mtn::g_launches;                    \
    MTN_CUDA(cudaGetLastError());         \
  } while (0)

// Kernel launch and dynamic shared memory.  MTN_HOST_EMU is defined only by the CPU test
// suite's SIMT emulator (tests/emu/, never built into libmartini_b200.so), which runs these
// same kernels as host code to check their logic against the oracle without a GPU.
#ifdef MTN_HOST_EMU
#define MTN_LAUNCH MTN_EMU_LAUNCH
#define MTN_DYN_SMEM(type, name) type* name = reinterpret_cast<type*>(mtn_emu::dyn_smem())
#else
#define MTN_LAUNCH(kernel, grid, block, smem, stream, ...) \
  kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#define MTN_DYN_SMEM(type, name) extern __shared__ __align__(128) type name[]
#endif

inline size_t align_up(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }

}  // namespace mtn
