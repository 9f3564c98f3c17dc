// Function tables used by the projection kernel, and their device-side evaluators.
//
// Two families, both filled once per device by the host in x87 extended precision
// (csrc/tables_host.hpp) and accurate to a few 1e-16 -- the level of libm itself:
//
//  * erf: Taylor coefficients about the centres of 1/16-wide intervals on [0, 6);
//  * pixel-integrated SPH kernels W(R^2) for the kernels whose closed form costs logs and
//    square roots (Wendland C2, Wendland C6, cubic spline,
//    quartic spline): piecewise degree-9 polynomials in R^2 on
//    intervals refined towards the points where the closed form is not analytic (below).
//    ~35 straight-line instructions instead of ~145 branchy ones, and two evaluations
//    interleave (the projection kernel evaluates two pixels per lane).
#pragma once

#include "common.cuh"

namespace mtn {

// ------------------------------------------------------------------------------------ erf
__device__ double g_erf_table[ERF_NINT * ERF_NCOEF];

// erf(t): exactly +-1 for |t| >= ERF_SAT (where erf rounds to 1 in float64 anyway), otherwise a
// Taylor polynomial about the centre of the table interval holding |t|.  Branch-free.  Two tables:
// g_erf_table (1/16-wide intervals, degree 9, truncation < 5e-18; read through L1 by the brick
// kernel, where few rows matter more than few loads) and the compact one (common.cuh) that the
// column / splat kernels copy into shared memory: every lane reads its own row there, and the
// L1 / shared-memory data pipe -- 128 B per cycle -- is what bounds them (a shared-memory row
// costs 4 wavefronts per 16-byte load of a warp, a global one 5.4 and three times the latency),
// so the compact table holds one 16-byte row per interval and the evaluator derives the rest.
__device__ double g_erf_table_compact[ERFC_DOUBLES];

// Exactly +-1 for |t| >= ERF_SAT without a test: the clamped argument lands in the last row,
// whose erf(x0) is 1.0 in float64 and whose A u is below 2^-58.  Never above 1: the truncated
// series stays below erf's distance from 1 everywhere (its remainder carries exp(-x0^2) too).
// The interval index comes from the float64 rounding trick (add 2^52 + 2^51: the sum's low
// mantissa word is the integer, the sum minus the constant its float64 value) instead of a
// float64 -> int32 -> float64 round trip through the conversion unit; any index whose centre is
// within 1/128 of |t| serves, so round-to-nearest of 64 |t| - 1/2 is as good as the floor.
__device__ __forceinline__ double erf_tab_compact(const double* __restrict__ table, double t) {
  constexpr double MAGIC = 6755399441055744.0;  // 2^52 + 2^51
  double a = fabs(t);
  a = a < ERF_SAT ? a : ERF_SAT;  // (NaN -> ERF_SAT, as fmin did)
  const double r_ = fma(a, (double)ERFC_INV_W, -0.5) + MAGIC;  // = MAGIC + round(64 a - 1/2)
  const int i = __double2loint(r_);  // in [0, 6 ERFC_INV_W]
  const double x0 = fma(r_ - MAGIC, 1.0 / ERFC_INV_W, 0.5 / ERFC_INV_W);
  const double u = a - x0;
  const double2 fa = reinterpret_cast<const double2*>(table)[i];  // {erf(x0), 2/sqrt(pi) exp(-x0^2)}
  // erf(x0 + u) = erf(x0) + A u (1 + u (k2 + u (k3 + u (k4 + u k5)))):
  //   k2 = -x0, k3 = (2 x0^2 - 1) / 3, k4 = -x0 (2 x0^2 - 3) / 6, k5 = (4 x0^4 - 12 x0^2 + 3) / 30
  const double q = x0 * x0;
  const double k3 = fma(q, 2.0 / 3.0, -1.0 / 3.0);
  const double k4 = x0 * fma(q, -1.0 / 3.0, 0.5);
  const double k5 = fma(fma(q, 4.0 / 30.0, -12.0 / 30.0), q, 3.0 / 30.0);
  double p = fma(k5, u, k4);
  p = fma(p, u, k3);
  p = fma(p, u, -x0);
  p = fma(p, u, 1.0);
  const double r = fma(fa.y * u, p, fa.x);
  return copysign(r, t);
}

__device__ __forceinline__ double erf_tab(double t) {
  const double a = fmin(fabs(t), ERF_SAT);
  const int i = (int)(a * ERF_INV_W);  // a = ERF_SAT lands in the last (saturated) interval
  const double u = a - ((double)i + 0.5) * (1.0 / ERF_INV_W);
  const double2* row = reinterpret_cast<const double2*>(g_erf_table + i * ERF_NCOEF);
  double2 c[ERF_NCOEF / 2];
#pragma unroll
  for (int k = ERF_NCOEF / 2 - 1; k >= 0; --k) c[k] = __ldg(row + k);
  double r = fma(c[ERF_NCOEF / 2 - 1].y, u, c[ERF_NCOEF / 2 - 1].x);
#pragma unroll
  for (int k = ERF_NCOEF / 2 - 2; k >= 0; --k) {
    r = fma(r, u, c[k].y);
    r = fma(r, u, c[k].x);
  }
  r = fmin(r, 1.0);
  r = fabs(t) >= ERF_SAT ? 1.0 : r;
  return copysign(r, t);
}

// ------------------------------------------------------------------- kernel-integral tables
// W(s), s = R^2 * scale, is tabulated directly in s -- no square root on the device.  The
// closed forms are analytic in s except at a few points: s = 0 (an s^2 log s term) and the
// ends of the kernel's pieces (half-integer powers of a^2 - s).  The support is cut into
// zones, each with one such anchor point at one end, and a zone is covered by intervals
// that shrink geometrically towards its anchor: with u = |s - anchor|, octave [2^e, 2^(e+1))
// of u is split into 8 equal intervals, so (interval width) / (distance to the anchor) <=
// 1/8 everywhere and a degree-9 interpolant converges like 33^-10; u < 2^-kmin is one "core"
// interval, where the non-analytic term itself is below 1e-16 of W(0).  The interval index
// is read off the exponent and top 3 mantissa bits of u, and so is the interval's centre;
// the polynomial's variable is u - centre (u = |s - anchor| is exact in float64 for every
// zone below), so only the ten coefficients are loaded.
constexpr int WT_DEG = 9;
constexpr int WT_ROW = 12;          // c0..c9 (in t = u - centre), interval centre, unused
// Wendland C6 and the quartic spline are tabulated too: their closed forms cost ~350
// instructions (40-term polynomials, a log or up to three asinh, up to four square roots);
// measured on the config-2 geometry with the C6 kernel: 5.45 ms (table) against 8.03 ms.
constexpr int WT_MAX_ZONES = 6;
constexpr int WT_KINDS = 6;         // indexed by MTN_KERNEL_*
constexpr int WT_MAX_ROWS = 4096;
// zones of a kind (compile-time kinds fold the zone search to the compares they need)
__host__ __device__ constexpr int wtab_zones_of(int kind) {
  return kind == MTN_KERNEL_WENDLANDC2 || kind == MTN_KERNEL_WENDLANDC6 ? 2
         : kind == MTN_KERNEL_CUBICSPLINE                                ? 4
                                                                         : WT_MAX_ZONES;
}
constexpr int WT_SUB_BITS = 3;      // 8 intervals per octave

struct WZone {
  double s_lo;    // the zone holds s_lo <= s < (next zone).s_lo; +inf for unused zones
  double anchor;  // u = |s - anchor|
  double core_centre;  // centre of the core interval: 2^-(kmin+1)
  int off;        // row = clamp((hi32(u) >> 17) + off, row0, last)
  int row0;       // the zone's core interval
  int last;       // the zone's last row
  int pad;
};

__constant__ WZone c_wzone[WT_KINDS][WT_MAX_ZONES];
__constant__ int c_wnz[WT_KINDS];        // 0: no table for this kind (closed form is used)
__constant__ double c_wscale[WT_KINDS];  // s = R^2 * scale (cubic spline: 4, its dij *= 2)
__constant__ double c_wend[WT_KINDS];    // W = 0 for s >= c_wend (edge of the support)
__device__ __align__(32) double g_wtab_rows[WT_MAX_ROWS * WT_ROW];

// Four consecutive doubles through the read-only path as ONE 256-bit load (sm_100: LDG.E.256).
// Every lane reads its own table row, so what a load instruction costs is L1 tag look-ups -- one
// per distinct 128-byte line among the lanes -- and three 32-byte loads per 96-byte row cost 40 %
// fewer of them than five 16-byte ones.
__device__ __forceinline__ void ldg256(const double* __restrict__ p, double& a, double& b, double& c, double& d) {
#ifdef MTN_HOST_EMU
  a = p[0]; b = p[1]; c = p[2]; d = p[3];
#else
  asm("ld.global.nc.v4.f64 {%0, %1, %2, %3}, [%4];" : "=d"(a), "=d"(b), "=d"(c), "=d"(d) : "l"(p));
#endif
}

__device__ __forceinline__ bool wtab_has(int kind) { return c_wnz[kind] > 0; }

// Table value of the pixel-integrated kernel (normalisation included, 1/h^2 not) at
// R2 = |d|^2 / h^2.  Straight-line code; with a compile-time `kind` the zone bounds fold
// into constant-bank operands.
__device__ __forceinline__ double wtab_eval(int kind, double R2) {
  const double s = R2 * c_wscale[kind];
  int z = 0;
#pragma unroll
  for (int k = 1; k < WT_MAX_ZONES; ++k)
    if (k < wtab_zones_of(kind)) z += s >= c_wzone[kind][k].s_lo ? 1 : 0;
  const WZone& zn = c_wzone[kind][z];
  const double u = fabs(s - zn.anchor);
  const int idx = min(max((__double2hiint(u) >> (20 - WT_SUB_BITS)) + zn.off, zn.row0), zn.last);
  const double* row = g_wtab_rows + idx * WT_ROW;  // (96-byte rows: 32-byte aligned)
  double2 c01, c23, c45, c67, c89;
  double unused0, unused1;
  ldg256(row + 8, c89.x, c89.y, unused0, unused1);
  ldg256(row + 4, c45.x, c45.y, c67.x, c67.y);
  ldg256(row, c01.x, c01.y, c23.x, c23.y);
  // the interval's centre (also stored in the row, for the host replica) from its own index:
  // start of the interval = key << 17 as a high word, plus half its width = the next mantissa bit
  const int key = idx - zn.off;
  const double centre = idx == zn.row0
                            ? zn.core_centre
                            : __hiloint2double((key << (20 - WT_SUB_BITS)) | (1 << (19 - WT_SUB_BITS)), 0);
  const double t = u - centre;
  double v = fma(c89.y, t, c89.x);
  v = fma(v, t, c67.y);
  v = fma(v, t, c67.x);
  v = fma(v, t, c45.y);
  v = fma(v, t, c45.x);
  v = fma(v, t, c23.y);
  v = fma(v, t, c23.x);
  v = fma(v, t, c01.y);
  v = fma(v, t, c01.x);
  return s >= c_wend[kind] ? 0.0 : v;  // the closed forms are exactly 0 from the edge on
}

}  // namespace mtn
