// Device functions: pixel-integrated SPH kernel weights and per-channel line spectra.
//
// Each function follows the closed-form approximation of the reference's
// `_kernel_integral` (martini/sph_kernels.py, lines cited per function) in float64.
// Geometric predicates (which branch a pixel falls in) are evaluated without FMA
// contraction so that branch selection matches numpy away from 1-ulp coincidences; inside
// a branch the arithmetic may be re-associated (Horner forms, reciprocals), which moves
// results by a few ulp -- ten orders below the 1e-6 x peak parity tolerance.
#pragma once

#include "common.cuh"
#include "tables.cuh"

namespace mtn {

// dx^2 + dy^2 exactly as np.power(dij, 2).sum(axis=0): two rounded squares, one add.
__device__ __forceinline__ double sq_dist(double dx, double dy) {
  return __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
}

// _WendlandC2Kernel._kernel_integral, sph_kernels.py:430-441.
__device__ __forceinline__ double w_wendland_c2(double dr2, double inv_h2) {
  const double R2 = dr2 * inv_h2;
  double val;
  if (R2 == 0.0) {
    val = 2.0 / 3.0;
  } else if (R2 < 1.0) {
    const double A = sqrt(1.0 - R2);
    // log((1 + A) / sqrt(R2))
    const double lg = log((1.0 + A) * rsqrt(R2));
    val = 5.0 * R2 * R2 * (0.5 * R2 + 3.0) * lg +
          A * (-27.0 / 2.0 * R2 * R2 - 14.0 / 3.0 * R2 + 2.0 / 3.0);
  } else {
    return 0.0;
  }
  return val * (21.0 / 2.0 / CUDART_PI) * inv_h2;
}

// Wendland C6 antiderivative, sph_kernels.py:592-674, regrouped as
//   P(z; R^2) + q * Q(z; R^2) + L * (7.21875 R^12 + 173.25 R^10 + 288.75 R^8)
// with q = sqrt(R^2 + z^2), L = log(q + z); P and Q are odd polynomials in z (Horner).
__device__ __forceinline__ double c6_indef(double R2, double z, double q, double L) {
  const double R4 = R2 * R2, R6 = R4 * R2, R8 = R4 * R4, R10 = R8 * R2, R12 = R8 * R4;
  const double z2 = z * z;
  const double p1 = 1.0 - 11.0 * R2 + 66.0 * R4 - 462.0 * R6 - 1155.0 * R8 - 231.0 * R10;
  const double p3 = -11.0 / 3.0 + 44.0 * R2 - 462.0 * R4 - 1540.0 * R6 - 385.0 * R8;
  const double p5 = 13.2 - 277.2 * R2 - 1386.0 * R4 - 462.0 * R6;
  const double p7 = -66.0 - 660.0 * R2 - 330.0 * R4;
  const double p9 = -(128.0 + 1.0 / 3.0) * (1.0 + R2);
  const double p11 = -21.0;
  const double P = z * (p1 + z2 * (p3 + z2 * (p5 + z2 * (p7 + z2 * (p9 + z2 * p11)))));
  const double q1 = 24.7813 * R10 + 530.75 * R8 + 767.25 * R6;
  const double q3 = 47.4792 * R8 + 819.5 * R6 + 896.5 * R4;
  const double q5 = 58.0167 * R6 + 752.4 * R4 + 550.0 * R2;
  const double q7 = 41.7 * R4 + 360.8 * R2 + 132.0;
  const double q9 = (16.0 + 4.0 / 15.0) * R2 + 70.4;
  const double q11 = 8.0 / 3.0;
  const double Q = z * (q1 + z2 * (q3 + z2 * (q5 + z2 * (q7 + z2 * (q9 + z2 * q11)))));
  return P + q * Q + L * (7.21875 * R12 + 173.25 * R10 + 288.75 * R8);
}

// _WendlandC6Kernel._kernel_integral, sph_kernels.py:676-685.
__device__ __forceinline__ double w_wendland_c6(double dr2, double h, double inv_h2) {
  const double R = sqrt(dr2) / h;
  const double norm = 1365.0 / 64.0 / CUDART_PI;
  double val;
  if (R == 0.0) {
    val = norm * 2.0 * (4.0 / 15.0);
  } else if (R < 1.0) {
    const double R2 = R * R;
    const double zmax = sqrt(1.0 - R2);
    const double q = sqrt(R2 + zmax * zmax);
    const double up = c6_indef(R2, zmax, q, log(q + zmax));
    // indef(R, 0): only the log terms survive, q = R, L = log(R)
    const double R4 = R2 * R2, R8 = R4 * R4;
    const double lo = log(R) * (7.21875 * R8 * R4 + 173.25 * R8 * R2 + 288.75 * R8);
    val = norm * 2.0 * (up - lo);
  } else {
    return 0.0;
  }
  return val * inv_h2;
}

// _CubicSplineKernel._kernel_integral, sph_kernels.py:821-858.
__device__ __forceinline__ double w_cubic_spline(double dx, double dy, double inv_h2) {
  // dij *= 2 (:821) is exact in binary, so R2 = ((2dx)^2 + (2dy)^2) / h^2
  const double dr2 = sq_dist(2.0 * dx, 2.0 * dy);
  const double R2 = dr2 * inv_h2;
  double val;
  if (R2 == 0.0) {
    val = 11.0 / 16.0 + 0.25 * 0.25;
  } else if (R2 <= 1.0) {
    const double A = sqrt(1.0 - R2);
    const double B = sqrt(4.0 - R2);
    const double lgA = log(1.0 + A);
    const double lgR = 0.5 * log(R2);  // log(sqrt(R2))
    const double I1 = A - 0.5 * A * A * A - 1.5 * R2 * A + 3.0 / 32.0 * A * (3.0 * R2 + 2.0) +
                      9.0 / 32.0 * R2 * R2 * (lgA - lgR);
    const double I3 = -B * (3.0 * R2 + 56.0) / 4.0 + A * (4.0 * R2 + 50.0) / 8.0 -
                      3.0 / 8.0 * R2 * (R2 + 16.0) * (log(2.0 + B) - lgA) +
                      2.0 * (3.0 * R2 + 4.0) * (B - A) + 2.0 * (B * B * B - A * A * A);
    val = I1 + 0.25 * I3;
  } else if (R2 <= 4.0) {
    const double B = sqrt(4.0 - R2);
    const double I2 = -B * (3.0 * R2 + 56.0) / 4.0 -
                      3.0 / 8.0 * R2 * (R2 + 16.0) * log((2.0 + B) * rsqrt(R2)) +
                      2.0 * (3.0 * R2 + 4.0) * B + 2.0 * B * B * B;
    val = 0.25 * I2;
  } else {
    return 0.0;
  }
  return val / 1.59689476201133 * inv_h2 * 4.0;
}

// _GaussianKernel._kernel_integral, sph_kernels.py:1024-1044.  The two truncation
// predicates keep the reference's division order (they switch a discontinuity).
__device__ __forceinline__ double w_gaussian(double dx, double dy, double h, double truncate,
                                             double norm) {
  const double sig = 0.42466090014400953;  // 1 / (2 sqrt(2 ln 2))
  const double dr = sqrt(sq_dist(dx, dy));
  if (__ddiv_rn(__ddiv_rn(__dsub_rn(dr, 0.70710678118654757), h), sig) > truncate) return 0.0;
  const double u = __ddiv_rn(__ddiv_rn(dr, h), sig);
  double ez = 0.0;
  if (truncate > u) {
    const double zmax = sqrt(__dsub_rn(__dmul_rn(truncate, truncate), __dmul_rn(u, u)));
    ez = erf_tab(zmax / 1.4142135623730951);
  }
  const double c = 1.0 / (h * 1.4142135623730951 * sig);
  const double ex = erf_tab((dx + 0.5) * c) - erf_tab((dx - 0.5) * c);
  const double ey = erf_tab((dy + 0.5) * c) - erf_tab((dy - 0.5) * c);
  return 0.25 * ez * ex * ey / norm;
}

// DiracDeltaKernel._kernel_integral, sph_kernels.py:1165 (strict on both axes).
__device__ __forceinline__ double w_dirac_delta(double dx, double dy) {
  return (fabs(dx) < 0.5 && fabs(dy) < 0.5) ? 1.0 : 0.0;
}

// IA(R, z, A) of _QuarticSplineKernel._kernel_integral, sph_kernels.py:1542-1553.
__device__ __forceinline__ double quartic_IA(double R, double R2, double A) {
  const double z = sqrt(A * A - R2);
  const double z2 = z * z;
  const double q = sqrt(z2 + R2);
  const double A2 = A * A;
  return A2 * A2 * z - 2.0 * A2 * A * z * q + 2.0 * A2 * z * (3.0 * R2 + z2) -
         A * R2 * (4.0 * A2 + 3.0 * R2) * asinh(z / R) / 2.0 -
         A * z * q * (5.0 * R2 + 2.0 * z2) / 2.0 + R2 * R2 * z + 2.0 * R2 * z2 * z / 3.0 +
         z2 * z2 * z / 5.0;
}

// _QuarticSplineKernel._kernel_integral, sph_kernels.py:1516-1564.
__device__ __forceinline__ double w_quartic_spline(double dr2, double h, double inv_h2) {
  const double R = sqrt(dr2) / h;
  double val;
  if (R == 0.0) {
    val = 384.0 / 3125.0;
  } else if (R < 1.0) {
    const double R2 = R * R;
    val = 0.0;
    if (R < 0.2) val += 10.0 * quartic_IA(R, R2, 0.2);
    if (R < 0.6) val -= 5.0 * quartic_IA(R, R2, 0.6);
    val += quartic_IA(R, R2, 1.0);
  } else {
    return 0.0;
  }
  return val * (2.0 * 15625.0 / 512.0 / CUDART_PI) * inv_h2;
}

// Closed forms, dispatched on the primitive kernel kind.  dx, dy = particle - pixel centre,
// as in martini.py:275-277.
__device__ __forceinline__ double kernel_weight_closed(int kind, double dx, double dy, double h,
                                                       double inv_h2, double truncate, double norm) {
  switch (kind) {
    case MTN_KERNEL_WENDLANDC2:
      return w_wendland_c2(sq_dist(dx, dy), inv_h2);
    case MTN_KERNEL_WENDLANDC6:
      return w_wendland_c6(sq_dist(dx, dy), h, inv_h2);
    case MTN_KERNEL_CUBICSPLINE:
      return w_cubic_spline(dx, dy, inv_h2);
    case MTN_KERNEL_GAUSSIAN:
      return w_gaussian(dx, dy, h, truncate, norm);
    case MTN_KERNEL_DIRACDELTA:
      return w_dirac_delta(dx, dy);
    case MTN_KERNEL_QUARTICSPLINE:
      return w_quartic_spline(sq_dist(dx, dy), h, inv_h2);
    default:
      return 0.0;
  }
}

// The weight the projection kernel uses: the tabulated form where one exists (Wendland C2,
// cubic spline; tables.cuh), the closed form otherwise.
__device__ __forceinline__ double kernel_weight(int kind, double dx, double dy, double h,
                                                double inv_h2, double truncate, double norm) {
  if (wtab_has(kind)) return wtab_eval(kind, sq_dist(dx, dy) * inv_h2) * inv_h2;
  return kernel_weight_closed(kind, dx, dy, h, inv_h2, truncate, norm);
}

// ---------------------------------------------------------------------------------------
// Spectra.  For channel c with edges (lo, hi) = (min, max) of edges[c], edges[c+1]:
//   Gaussian  : 0.5 * [erf((hi - v) / sqrt2 / sigma) - erf((lo - v) / sqrt2 / sigma)]
//               spectral_models.py:387-425
//   DiracDelta: heaviside(v - lo, 1) * heaviside(hi - v, 1)   spectral_models.py:544-570
// then * mHI * D^-2 / |hi - lo| / 2.36e5 (spectral_models.py:119-145).
// ---------------------------------------------------------------------------------------

// erf of a channel edge seen from a particle.
__device__ __forceinline__ double edge_erf(double edge, double v, double inv_s) {
  return erf_tab((edge - v) * inv_s);
}

__device__ __forceinline__ double dirac_channel(double lo, double hi, double v) {
  // np.heaviside(x, 1.0): 1 for x >= 0, 0 for x < 0, NaN for NaN
  return (v - lo >= 0.0 && hi - v >= 0.0) ? 1.0 : 0.0;
}

}  // namespace mtn
