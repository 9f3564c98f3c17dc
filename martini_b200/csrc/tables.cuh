// Function tables used by the projection kernel, and their device-side evaluators.
//
// Two families, both filled once per device by the host in x87 extended precision
// (csrc/tables_host.hpp) and accurate to a few 1e-16 -- the level of libm itself:
//
//  * erf: Taylor coefficients about the centres of 1/16-wide intervals on [0, 6);
//  * pixel-integrated SPH kernels W(R^2) for the kernels whose closed form costs a log and
//    two square roots (Wendland C2, cubic spline): piecewise degree-9 polynomials in a
//    variable chosen per region so that the function is analytic there --
//       x = sqrt(s)        near the centre, on dyadic intervals [2^-k-1, 2^-k) split in 4
//                          (the s^2 log s term of the closed forms is only C^3 at s = 0),
//       x = sqrt(a^2 - s)  towards a knot or the edge of the support (the closed forms are
//                          odd analytic functions of sqrt(1 - R^2) resp. sqrt(4 - R^2) there).
//    ~40 straight-line instructions instead of ~145 branchy ones, and two evaluations
//    interleave (the projection kernel evaluates two pixels per lane).
#pragma once

#include "common.cuh"

namespace mtn {

// ------------------------------------------------------------------------------------ erf
__device__ double g_erf_table[ERF_NINT * ERF_NCOEF];

// erf(t) to ~1 ulp: exactly +-1 for |t| >= ERF_SAT (where erf rounds to 1 in float64 anyway),
// otherwise a degree-9 Taylor polynomial about the centre of the 1/16-wide interval holding
// |t| (|u| <= 1/32: truncation < 5e-18).  Branch-free.
__device__ __forceinline__ double erf_tab(double t) {
  const double a = fmin(fabs(t), ERF_SAT);
  const int i = (int)(a * ERF_INV_W);  // a = ERF_SAT lands in the last (saturated) interval
  const double u = a - ((double)i + 0.5) * (1.0 / ERF_INV_W);
  const double2* row = reinterpret_cast<const double2*>(g_erf_table + i * ERF_NCOEF);
  const double2 c89 = __ldg(row + 4), c67 = __ldg(row + 3), c45 = __ldg(row + 2),
                c23 = __ldg(row + 1), c01 = __ldg(row);
  double r = fma(c89.y, u, c89.x);
  r = fma(r, u, c67.y);
  r = fma(r, u, c67.x);
  r = fma(r, u, c45.y);
  r = fma(r, u, c45.x);
  r = fma(r, u, c23.y);
  r = fma(r, u, c23.x);
  r = fma(r, u, c01.y);
  r = fmin(fma(r, u, c01.x), 1.0);
  r = fabs(t) >= ERF_SAT ? 1.0 : r;
  return copysign(r, t);
}

// ------------------------------------------------------------------- kernel-integral tables
constexpr int WT_DEG = 9;
constexpr int WT_ROW = 12;          // c0..c9, interval centre, 1 / half-width
constexpr int WT_MAX_REGIONS = 4;
constexpr int WT_KINDS = 6;         // indexed by MTN_KERNEL_*
constexpr int WT_MAX_ROWS = 256;
constexpr int WT_DYADIC_KMIN = 8;   // first dyadic interval is [0, 2^-8)
constexpr int WT_DYADIC_SUB = 4;    // sub-intervals per octave

struct WRegion {
  double s_max;   // region holds s <= s_max (regions ascending; s = R^2 * scale)
  double a2;      // x = sqrt(sgn * s + a2): (sgn, a2) = (+1, 0) or (-1, a^2)
  double sgn;
  double x_lo;    // uniform regions: first interval starts here
  double inv_w;   // uniform regions: 1 / interval width
  int row0;       // first table row of the region
  int n_int;      // number of intervals (rows)
  int dyadic;     // 1: dyadic intervals in x, 0: uniform
  int pad;
};

__constant__ WRegion c_wreg[WT_KINDS][WT_MAX_REGIONS];
__constant__ int c_wnreg[WT_KINDS];      // 0: no table for this kind (closed form is used)
__constant__ double c_wscale[WT_KINDS];  // s = R^2 * scale (cubic spline: 4, its dij *= 2)
__device__ double g_wtab_rows[WT_MAX_ROWS * WT_ROW];

__device__ __forceinline__ bool wtab_has(int kind) { return c_wnreg[kind] > 0; }

// Table value of the pixel-integrated kernel (normalisation included, 1/h^2 not) at
// R2 = |d|^2 / h^2.  Straight-line code: selects, no divergent branches.
__device__ __forceinline__ double wtab_eval(int kind, double R2) {
  const double s = R2 * c_wscale[kind];
  const int nreg = c_wnreg[kind];
  int r = 0;
#pragma unroll
  for (int k = 0; k < WT_MAX_REGIONS - 1; ++k) r += (k + 1 < nreg && s > c_wreg[kind][k].s_max) ? 1 : 0;
  const WRegion& reg = c_wreg[kind][r];
  const double arg = fmax(fma(reg.sgn, s, reg.a2), 0.0);
  const double x = arg > 0.0 ? arg * rsqrt(arg) : 0.0;
  const int hi = __double2hiint(x);
  const int idx_dy = x < 1.0 / (1 << WT_DYADIC_KMIN)
                         ? 0
                         : (((hi >> 20) - 1023 + WT_DYADIC_KMIN) * WT_DYADIC_SUB + ((hi >> 18) & 3) + 1);
  const int idx_un = (int)((x - reg.x_lo) * reg.inv_w);
  const int idx = min(max(reg.dyadic ? idx_dy : idx_un, 0), reg.n_int - 1);
  const double2* row = reinterpret_cast<const double2*>(g_wtab_rows + (reg.row0 + idx) * WT_ROW);
  const double2 cw = __ldg(row + 5), c89 = __ldg(row + 4), c67 = __ldg(row + 3), c45 = __ldg(row + 2),
                c23 = __ldg(row + 1), c01 = __ldg(row);
  const double t = (x - cw.x) * cw.y;
  double v = fma(c89.y, t, c89.x);
  v = fma(v, t, c67.y);
  v = fma(v, t, c67.x);
  v = fma(v, t, c45.y);
  v = fma(v, t, c45.x);
  v = fma(v, t, c23.y);
  v = fma(v, t, c23.x);
  v = fma(v, t, c01.y);
  v = fma(v, t, c01.x);
  // at and beyond the end of the last region: outside the kernel's support (the closed forms
  // are exactly 0 at the edge itself)
  return (r == nreg - 1 && s >= reg.s_max) ? 0.0 : v;
}

}  // namespace mtn
