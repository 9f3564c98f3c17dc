// Host-side generation of the function tables of tables.cuh, in x87 extended precision.
//
// Every table entry is derived from the reference's own closed forms -- erf's derivatives,
// and the `_kernel_integral` expressions of martini/sph_kernels.py (WendlandC2 :430-441,
// CubicSpline :821-858) evaluated in long double -- by Chebyshev interpolation on each
// interval of tables.cuh's zone layout; the fit is then checked against the closed form on
// a dense sample and the worst error kept (mtn_table_error), so a bad table cannot go
// unnoticed.
#pragma once

#include <cmath>
#include <cstring>
#include <vector>

#include "tables.cuh"

namespace mtn {

typedef long double ld;

// ----------------------------------------------------------------------------- closed forms
// Every closed form takes s as (anchor, ds) with s = anchor + ds, so that a^2 - s is formed as
// (a^2 - anchor) - ds: near a zone anchor the distance to it keeps its full relative
// precision (the cubic spline's sqrt(1 - s) term needs that down to 1 - s = 2^-53).
static const ld PI_L = 3.14159265358979323846264338327950288L;

// _WendlandC2Kernel._kernel_integral * h^2, s = R^2
static ld F_wendland_c2(ld anchor, ld ds) {
  const ld s = anchor + ds, om = (1.0L - anchor) - ds;  // 1 - s
  if (om <= 0.0L) return 0.0L;
  const ld norm = 21.0L / 2.0L / PI_L;
  if (s <= 0.0L) return norm * 2.0L / 3.0L;
  const ld A = sqrtl(om);
  return norm * (5.0L * s * s * (0.5L * s + 3.0L) * logl((1.0L + A) / sqrtl(s)) +
                 A * (-27.0L / 2.0L * s * s - 14.0L / 3.0L * s + 2.0L / 3.0L));
}

// _CubicSplineKernel._kernel_integral * h^2, s = R2 of the reference (= 4 |d|^2 / h^2)
static ld F_cubic_spline(ld anchor, ld ds) {
  const ld s = anchor + ds, om = (1.0L - anchor) - ds, fm = (4.0L - anchor) - ds;  // 1 - s, 4 - s
  const ld scale = 4.0L / 1.59689476201133L;
  if (fm < 0.0L) return 0.0L;
  if (s <= 0.0L) return scale * (11.0L / 16.0L + 0.25L * 0.25L);
  if (om >= 0.0L) {
    const ld A = sqrtl(om), B = sqrtl(fm);
    const ld I1 = A - 0.5L * A * A * A - 1.5L * s * A + 3.0L / 32.0L * A * (3.0L * s + 2.0L) +
                  9.0L / 32.0L * s * s * (logl(1.0L + A) - logl(sqrtl(s)));
    const ld I3 = -B * (3.0L * s + 56.0L) / 4.0L + A * (4.0L * s + 50.0L) / 8.0L -
                  3.0L / 8.0L * s * (s + 16.0L) * logl((2.0L + B) / (1.0L + A)) +
                  2.0L * (3.0L * s + 4.0L) * (B - A) + 2.0L * (B * B * B - A * A * A);
    return scale * (I1 + 0.25L * I3);
  }
  const ld B = sqrtl(fm);
  const ld I2 = -B * (3.0L * s + 56.0L) / 4.0L -
                3.0L / 8.0L * s * (s + 16.0L) * logl((2.0L + B) / sqrtl(s)) +
                2.0L * (3.0L * s + 4.0L) * B + 2.0L * B * B * B;
  return scale * 0.25L * I2;
}

// _WendlandC6Kernel._kernel_integral * h^2 (sph_kernels.py:575-685), s = R^2.  The reference's
// antiderivative regrouped as P(z) + q Q(z) + L (7.21875 R^12 + 173.25 R^10 + 288.75 R^8), P and
// Q odd polynomials in z with polynomial coefficients in R^2 (the same regrouping as the
// device's c6_indef, kernel_integrals.cuh).  Its decimal coefficients are the reference's
// float64 literals (24.7813, 128 + 1/3, ...: rounded in float64 first, as Python does).
static ld F_wendland_c6(ld anchor, ld ds) {
  const ld s = anchor + ds, om = (1.0L - anchor) - ds;  // R^2, 1 - R^2
  if (om <= 0.0L) return 0.0L;
  const ld norm = (ld)(1365.0 / 64.0) / PI_L;
  if (s <= 0.0L) return norm * 2.0L * (ld)(4.0 / 15.0);
  const ld z = sqrtl(om), z2 = z * z, q = sqrtl(s + z2), L = logl(q + z);
  const ld R2 = s, R4 = R2 * R2, R6 = R4 * R2, R8 = R4 * R4, R10 = R8 * R2, R12 = R8 * R4;
  const ld p1 = 1.0L - 11.0L * R2 + 66.0L * R4 - 462.0L * R6 - 1155.0L * R8 - 231.0L * R10;
  const ld p3 = -(ld)(11.0 / 3.0) + 44.0L * R2 - 462.0L * R4 - 1540.0L * R6 - 385.0L * R8;
  const ld p5 = (ld)13.2 - (ld)277.2 * R2 - 1386.0L * R4 - 462.0L * R6;
  const ld p7 = -66.0L - 660.0L * R2 - 330.0L * R4;
  const ld p9 = -(ld)(128.0 + 1.0 / 3.0) * (1.0L + R2);
  const ld p11 = -21.0L;
  const ld P = z * (p1 + z2 * (p3 + z2 * (p5 + z2 * (p7 + z2 * (p9 + z2 * p11)))));
  const ld q1 = (ld)24.7813 * R10 + (ld)530.75 * R8 + (ld)767.25 * R6;
  const ld q3 = (ld)47.4792 * R8 + (ld)819.5 * R6 + (ld)896.5 * R4;
  const ld q5 = (ld)58.0167 * R6 + (ld)752.4 * R4 + 550.0L * R2;
  const ld q7 = (ld)41.7 * R4 + (ld)360.8 * R2 + 132.0L;
  const ld q9 = (ld)(16.0 + 4.0 / 15.0) * R2 + (ld)70.4;
  const ld q11 = (ld)(8.0 / 3.0);
  const ld Q = z * (q1 + z2 * (q3 + z2 * (q5 + z2 * (q7 + z2 * (q9 + z2 * q11)))));
  const ld logs = (ld)7.21875 * R12 + (ld)173.25 * R10 + (ld)288.75 * R8;
  const ld up = P + q * Q + L * logs;      // indef(R, zmax)
  const ld lo = 0.5L * logl(s) * logs;     // indef(R, 0): q = R, L = log R, every z power gone
  return norm * 2.0L * (up - lo);
}

// _QuarticSplineKernel._kernel_integral * h^2 (sph_kernels.py:1516-1564), s = R^2.  The three
// IA pieces end at R = A = 0.2, 0.6, 1 (float64 values, as are the reference's A**2 under
// the square roots); `piece` = how many of them are switched on is fixed per zone by the
// caller through the zone's bounds, here it follows from the sign of A^2 - s.
static ld F_quartic_spline(ld anchor, ld ds) {
  const ld s = anchor + ds;
  if ((1.0L - anchor) - ds <= 0.0L) return 0.0L;
  const ld pref = 2.0L * (ld)(15625.0 / 512.0) / PI_L;
  if (s <= 0.0L) return pref * (ld)(384.0 / 3125.0);
  const ld R = sqrtl(s);
  auto IA = [&](double Ad, double A2d) -> ld {
    const ld A = (ld)Ad, z2 = ((ld)A2d - anchor) - ds;  // A^2 - R^2 with the zone's precision
    if (z2 <= 0.0L) return 0.0L;                      // piece switched off (R >= A)
    const ld z = sqrtl(z2), q = sqrtl(z2 + s), A2 = A * A;
    return A2 * A2 * z - 2.0L * A2 * A * z * q + 2.0L * A2 * z * (3.0L * s + z2) -
           A * s * (4.0L * A2 + 3.0L * s) * asinhl(z / R) / 2.0L -
           A * z * q * (5.0L * s + 2.0L * z2) / 2.0L + s * s * z + 2.0L * s * z2 * z / 3.0L +
           z2 * z2 * z / 5.0L;
  };
  return pref * (10.0L * IA(0.2, 0.2 * 0.2) - 5.0L * IA(0.6, 0.6 * 0.6) + IA(1.0, 1.0));
}

// --------------------------------------------------------------------------------- fitting
// Degree-WT_DEG Chebyshev interpolant of g(u) = f(anchor, dir * u) on u in [a, b], returned as
// monomial coefficients in t = u - centre (unscaled: the half-widths are powers of two, so
// dividing the coefficients of the scaled variable by hw^k is exact up to the final rounding
// to double; the centre is a dyadic number of the interval's own magnitude, hence exact too).
template <typename F>
static void fit_interval(F f, ld anchor, ld dir, ld a, ld b, double* row) {
  const int n = WT_DEG + 1;
  ld fx[n], cheb[n];
  const ld c = 0.5L * (a + b), hw = 0.5L * (b - a);
  for (int k = 0; k < n; ++k) fx[k] = f(anchor, dir * (c + hw * cosl(PI_L * (k + 0.5L) / n)));
  for (int j = 0; j < n; ++j) {
    ld sum = 0.0L;
    for (int k = 0; k < n; ++k) sum += fx[k] * cosl(PI_L * j * (k + 0.5L) / n);
    cheb[j] = sum * 2.0L / n;
  }
  cheb[0] *= 0.5L;
  // Chebyshev -> monomial through T_(j+1) = 2 t T_j - T_(j-1)
  ld mono[n] = {0}, Tprev[n] = {0}, Tcur[n] = {0}, Tnext[n];
  Tprev[0] = 1.0L;  // T_0
  Tcur[1] = 1.0L;   // T_1
  mono[0] += cheb[0];
  for (int i = 0; i < n; ++i) mono[i] += cheb[1] * Tcur[i];
  for (int j = 2; j < n; ++j) {
    for (int i = 0; i < n; ++i) Tnext[i] = (i > 0 ? 2.0L * Tcur[i - 1] : 0.0L) - Tprev[i];
    for (int i = 0; i < n; ++i) {
      mono[i] += cheb[j] * Tnext[i];
      Tprev[i] = Tcur[i];
      Tcur[i] = Tnext[i];
    }
  }
  ld scale = 1.0L;
  for (int i = 0; i < n; ++i) {
    row[i] = (double)(mono[i] * scale);
    scale /= hw;
  }
  row[WT_DEG + 1] = (double)c;
  row[WT_DEG + 2] = 0.0;
}

// One zone of the support: s_lo <= s < s_hi, intervals refined towards `anchor` (one of the
// two ends); u < 2^-kmin is the core interval.
struct ZoneSpec {
  ld s_lo, s_hi, anchor;
  int kmin;
};

// core-interval exponents of the Wendland C6 / quartic spline zones (u < 2^-k is one interval)
#ifndef WTAB_C6_K0
#define WTAB_C6_K0 20
#define WTAB_C6_K1 54
#define WTAB_Q_K0 10
#define WTAB_Q_KNOT 12
#define WTAB_Q_UP 8
#endif

struct HostTables {
  WZone zone[WT_KINDS][WT_MAX_ZONES];
  int nz[WT_KINDS];
  double scale[WT_KINDS];
  double end[WT_KINDS];
  std::vector<double> rows;
  double max_err[WT_KINDS];  // worst |table - closed form| / F(0) on a dense sample
};

static int hi32(double x) {
  long long b;
  memcpy(&b, &x, sizeof(b));
  return (int)(b >> 32);
}

// double-precision replica of the device evaluator (tables.cuh: wtab_eval)
static double host_wtab_eval(const HostTables& T, int kind, double R2) {
  const double s = R2 * T.scale[kind];
  int z = 0;
  for (int k = 1; k < WT_MAX_ZONES; ++k) z += s >= T.zone[kind][k].s_lo ? 1 : 0;
  const WZone& zn = T.zone[kind][z];
  if (s >= T.end[kind]) return 0.0;
  const double u = std::fabs(s - zn.anchor);
  const int idx = std::min(std::max((hi32(u) >> (20 - WT_SUB_BITS)) + zn.off, zn.row0), zn.last);
  const double* row = T.rows.data() + (size_t)idx * WT_ROW;
  // the centre from the interval's own index, as on the device; row[10] holds the same number
  const int key = idx - zn.off;
  const long long cbits = (long long)((key << (20 - WT_SUB_BITS)) | (1 << (19 - WT_SUB_BITS))) << 32;
  double centre;
  memcpy(&centre, &cbits, sizeof(centre));
  if (idx == zn.row0) centre = zn.core_centre;
  if (centre != row[10]) return NAN;  // a layout bug would poison max_err
  const double t = u - centre;
  double v = row[9];
  for (int k = 8; k >= 0; --k) v = std::fma(v, t, row[k]);
  return v;
}

template <typename F>
static void build_kind(HostTables& T, int kind, double scale, F f, const std::vector<ZoneSpec>& specs) {
  const int nsub = 1 << WT_SUB_BITS;
  T.scale[kind] = scale;
  T.nz[kind] = (int)specs.size();
  T.end[kind] = (double)specs.back().s_hi;
  for (size_t z = 0; z < specs.size(); ++z) {
    const ZoneSpec& sp = specs[z];
    WZone& zn = T.zone[kind][z];
    const bool up = sp.anchor <= sp.s_lo;  // u = s - anchor grows with s
    const ld u_max = up ? sp.s_hi - sp.anchor : sp.anchor - sp.s_lo;
    zn.s_lo = (double)sp.s_lo;
    zn.anchor = (double)sp.anchor;
    zn.core_centre = (double)ldexpl(1.0L, -sp.kmin - 1);
    zn.row0 = (int)(T.rows.size() / WT_ROW);
    zn.off = zn.row0 + 1 - ((1023 - sp.kmin) << WT_SUB_BITS);
    zn.pad = 0;
    auto add = [&](ld u0, ld u1) {  // interval [u0, u1) of u
      T.rows.resize(T.rows.size() + WT_ROW);
      fit_interval(f, sp.anchor, up ? 1.0L : -1.0L, u0, u1, T.rows.data() + T.rows.size() - WT_ROW);
    };
    add(0.0L, ldexpl(1.0L, -sp.kmin));
    for (int e = -sp.kmin;; ++e) {
      const ld lo = ldexpl(1.0L, e), w = lo / nsub;
      bool more = true;
      for (int j = 0; j < nsub && more; ++j) {
        if (lo + j * w > u_max) more = false;  // u == u_max itself still needs its interval
        else add(lo + j * w, lo + (j + 1) * w);
      }
      if (!more) break;
    }
    zn.last = (int)(T.rows.size() / WT_ROW) - 1;
  }
  for (size_t z = specs.size(); z < (size_t)WT_MAX_ZONES; ++z)
    T.zone[kind][z] = WZone{HUGE_VAL, 0.0, 0.0, 0, 0, 0, 0};
  // verify on a dense sample of s: uniform over the support, and geometrically approaching
  // every zone anchor from inside the zone
  const ld f0 = f(0.0L, 0.0L);
  double worst = 0.0;
  auto check = [&](ld anchor, ld s) {  // the truth is taken at the double the device sees
    const double R2 = (double)(s / scale);
    const double sd = R2 * scale;
    const double got = host_wtab_eval(T, kind, R2);
    const double err = (double)(fabsl((ld)got - f(anchor, (ld)sd - anchor)) / f0);
    if (!(err <= worst)) worst = err;  // NaN (a layout bug flagged by the replica) sticks
  };
  const ld s_end = specs.back().s_hi;
  for (int i = 0; i <= 40000; ++i) check(0.0L, (ld)i / 40000.0L * s_end * (1.0L - 1e-12L));
  for (const ZoneSpec& sp : specs) {
    const bool up = sp.anchor <= sp.s_lo;
    const ld u_max = up ? sp.s_hi - sp.anchor : sp.anchor - sp.s_lo;
    for (int i = 0; i < 6000; ++i) {
      const ld u = u_max * powl(2.0L, -(ld)i / 100.0L);  // 60 octaves, 100 samples each
      check(sp.anchor, up ? sp.anchor + u : sp.anchor - u);
    }
  }
  T.max_err[kind] = worst;
}

static HostTables build_kernel_tables() {
  HostTables T;
  for (int k = 0; k < WT_KINDS; ++k) {
    T.nz[k] = 0;
    T.scale[k] = 1.0;
    T.end[k] = 0.0;
    T.max_err[k] = 0.0;
    for (int z = 0; z < WT_MAX_ZONES; ++z) T.zone[k][z] = WZone{HUGE_VAL, 0.0, 0.0, 0, 0, 0, 0};
  }
  // Wendland C2: s^2 log s at 0, (1 - s)^(9/2) at the edge
  build_kind(T, MTN_KERNEL_WENDLANDC2, 1.0, F_wendland_c2,
             {{0.0L, 0.5L, 0.0L, 28}, {0.5L, 1.0L, 1.0L, 12}});
  // cubic spline (s = 4 |d|^2 / h^2): s^2 log s at 0; below the knot s = 1 the reference's
  // expression carries a small (~2e-3 W(0)) sqrt(1 - s) term, so that zone is refined down to
  // u = 2^-54, below the smallest non-zero 1 - s a double s can give (rows that are
  // practically never touched); analytic above the knot;
  // (4 - s)^(7/2) at the edge
  build_kind(T, MTN_KERNEL_CUBICSPLINE, 4.0, F_cubic_spline,
             {{0.0L, 0.5L, 0.0L, 28}, {0.5L, 1.0L, 1.0L, 54}, {1.0L, 2.5L, 1.0L, 6}, {2.5L, 4.0L, 4.0L, 16}});
  // Wendland C6: s^4 log s at 0, (1 - s)^(17/2) at the edge
  build_kind(T, MTN_KERNEL_WENDLANDC6, 1.0, F_wendland_c6,
             {{0.0L, 0.5L, 0.0L, WTAB_C6_K0}, {0.5L, 1.0L, 1.0L, WTAB_C6_K1}});
  // quartic spline: analytic at 0 (the s log s terms of its three pieces cancel) and between
  // its knots; (A^2 - s)^(9/2) below each of s = A^2 = 0.04, 0.36, 1, nothing of that piece
  // above -- one zone on either side of the two inner knots
  {
    const ld k1 = (ld)(0.2 * 0.2), k2 = (ld)(0.6 * 0.6);
    build_kind(T, MTN_KERNEL_QUARTICSPLINE, 1.0, F_quartic_spline,
               {{0.0L, 0.02L, 0.0L, WTAB_Q_K0}, {0.02L, k1, k1, WTAB_Q_KNOT}, {k1, 0.2L, k1, WTAB_Q_UP},
                {0.2L, k2, k2, WTAB_Q_KNOT}, {k2, 0.68L, k2, WTAB_Q_UP}, {0.68L, 1.0L, 1.0L, WTAB_Q_KNOT}});
  }
  return T;
}

// erf: Taylor coefficients about the interval centres,
// erf^(k)(x) = (2/sqrt(pi)) (-1)^(k-1) H_(k-1)(x) exp(-x^2), H = physicists' Hermite.
static void build_erf_table_generic(double* tab, int inv_w, int deg, int nint) {
  const ld two_over_sqrt_pi = 1.1283791670955125738961589031215452L;
  const int ncoef = deg + 1;
  for (int i = 0; i < nint; ++i) {
    const ld c = ((ld)i + 0.5L) / inv_w;
    const ld ex = expl(-c * c);
    ld h_prev = 0.0L, h = 1.0L, fact = 1.0L;  // H_(-1) := 0, H_0 = 1
    tab[i * ncoef + 0] = (double)erfl(c);
    for (int k = 1; k <= deg; ++k) {
      fact *= k;
      const ld sign = ((k - 1) & 1) ? -1.0L : 1.0L;
      tab[i * ncoef + k] = (double)(two_over_sqrt_pi * sign * h * ex / fact);
      const ld h_next = 2.0L * c * h - 2.0L * (k - 1) * h_prev;
      h_prev = h;
      h = h_next;
    }
  }
}
static void build_erf_table(double* tab) { build_erf_table_generic(tab, ERF_INV_W, ERF_DEG, ERF_NINT); }
// the compact table of the column / splat kernels (tables.cuh: erf_tab_compact): value and
// first derivative at the interval centres
static void build_erf_table_compact(double* tab) {
  const ld two_over_sqrt_pi = 1.1283791670955125738961589031215452L;
  for (int r = 0; r < ERFC_NINT; ++r) {
    const ld c = ((ld)r + 0.5L) / ERFC_INV_W;
    tab[ERFC_NCOEF * r + 0] = (double)erfl(c);
    tab[ERFC_NCOEF * r + 1] = (double)(two_over_sqrt_pi * expl(-c * c));
  }
}

}  // namespace mtn
