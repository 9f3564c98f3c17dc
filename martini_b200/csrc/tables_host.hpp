// Host-side generation of the function tables of tables.cuh, in x87 extended precision.
//
// Every table entry is derived from the reference's own closed forms -- erf's derivatives,
// and the `_kernel_integral` expressions of martini/sph_kernels.py (WendlandC2 :430-441,
// CubicSpline :821-858) evaluated in long double -- by Chebyshev interpolation on each
// interval; the fit is then checked against the closed form on a dense sample and the
// worst error kept (mtn_table_error), so a bad table cannot go unnoticed.
#pragma once

#include <cmath>
#include <vector>

#include "tables.cuh"

namespace mtn {

typedef long double ld;

// ----------------------------------------------------------------------------- closed forms
static const ld PI_L = 3.14159265358979323846264338327950288L;

// _WendlandC2Kernel._kernel_integral * h^2, s = R^2
static ld F_wendland_c2(ld s) {
  if (s >= 1.0L) return 0.0L;
  const ld norm = 21.0L / 2.0L / PI_L;
  if (s <= 0.0L) return norm * 2.0L / 3.0L;
  const ld A = sqrtl(1.0L - s);
  return norm * (5.0L * s * s * (0.5L * s + 3.0L) * logl((1.0L + A) / sqrtl(s)) +
                 A * (-27.0L / 2.0L * s * s - 14.0L / 3.0L * s + 2.0L / 3.0L));
}

// _CubicSplineKernel._kernel_integral * h^2, s = R2 of the reference (= 4 |d|^2 / h^2)
static ld F_cubic_spline(ld s) {
  const ld scale = 4.0L / 1.59689476201133L;
  if (s > 4.0L) return 0.0L;
  if (s <= 0.0L) return scale * (11.0L / 16.0L + 0.25L * 0.25L);
  if (s <= 1.0L) {
    const ld A = sqrtl(1.0L - s), B = sqrtl(4.0L - s);
    const ld I1 = A - 0.5L * A * A * A - 1.5L * s * A + 3.0L / 32.0L * A * (3.0L * s + 2.0L) +
                  9.0L / 32.0L * s * s * (logl(1.0L + A) - logl(sqrtl(s)));
    const ld I3 = -B * (3.0L * s + 56.0L) / 4.0L + A * (4.0L * s + 50.0L) / 8.0L -
                  3.0L / 8.0L * s * (s + 16.0L) * logl((2.0L + B) / (1.0L + A)) +
                  2.0L * (3.0L * s + 4.0L) * (B - A) + 2.0L * (B * B * B - A * A * A);
    return scale * (I1 + 0.25L * I3);
  }
  const ld B = sqrtl(4.0L - s);
  const ld I2 = -B * (3.0L * s + 56.0L) / 4.0L -
                3.0L / 8.0L * s * (s + 16.0L) * logl((2.0L + B) / sqrtl(s)) +
                2.0L * (3.0L * s + 4.0L) * B + 2.0L * B * B * B;
  return scale * 0.25L * I2;
}

// --------------------------------------------------------------------------------- fitting
// Degree-WT_DEG Chebyshev interpolant of f on [a, b], returned as monomial coefficients in
// t = (x - centre) / halfwidth.
template <typename F>
static void fit_interval(F f, ld a, ld b, double* row) {
  const int n = WT_DEG + 1;
  ld fx[n], cheb[n];
  const ld c = 0.5L * (a + b), hw = 0.5L * (b - a);
  for (int k = 0; k < n; ++k) fx[k] = f(c + hw * cosl(PI_L * (k + 0.5L) / n));
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
  for (int i = 0; i < n; ++i) row[i] = (double)mono[i];
  row[WT_DEG + 1] = (double)c;
  row[WT_DEG + 2] = (double)(1.0L / hw);
}

struct RegionSpec {
  ld s_max;     // region holds s <= s_max
  int var;      // 0: x = sqrt(s), 1: x = sqrt(a2 - s)
  ld a2;
  int dyadic;   // 1: dyadic intervals on [0, x_hi), 0: n uniform intervals on [x_lo, x_hi]
  ld x_lo, x_hi;
  int n;
};

struct HostTables {
  WRegion reg[WT_KINDS][WT_MAX_REGIONS];
  int nreg[WT_KINDS];
  double scale[WT_KINDS];
  std::vector<double> rows;
  double max_err[WT_KINDS];  // worst |table - closed form| / F(0) on a dense sample
};

// double-precision replica of the device evaluator (tables.cuh: wtab_eval)
static double host_wtab_eval(const HostTables& T, int kind, double R2) {
  const double s = R2 * T.scale[kind];
  const int nreg = T.nreg[kind];
  int r = 0;
  for (int k = 0; k < WT_MAX_REGIONS - 1; ++k) r += (k + 1 < nreg && s > T.reg[kind][k].s_max) ? 1 : 0;
  const WRegion& reg = T.reg[kind][r];
  if (r == nreg - 1 && s >= reg.s_max) return 0.0;
  const double arg = std::fmax(std::fma(reg.sgn, s, reg.a2), 0.0);
  const double x = std::sqrt(arg);
  int idx;
  if (reg.dyadic) {
    int e;
    const double m = std::frexp(x, &e);  // x = m 2^e, m in [0.5, 1)
    idx = x < 1.0 / (1 << WT_DYADIC_KMIN)
              ? 0
              : ((e - 1 + WT_DYADIC_KMIN) * WT_DYADIC_SUB + ((int)(m * 8.0) & 3) + 1);
  } else {
    idx = (int)((x - reg.x_lo) * reg.inv_w);
  }
  idx = std::min(std::max(idx, 0), reg.n_int - 1);
  const double* row = T.rows.data() + (size_t)(reg.row0 + idx) * WT_ROW;
  const double t = (x - row[10]) * row[11];
  double v = row[9];
  for (int k = 8; k >= 0; --k) v = std::fma(v, t, row[k]);
  return v;
}

template <typename F>
static void build_kind(HostTables& T, int kind, double scale, F f, const std::vector<RegionSpec>& specs) {
  T.scale[kind] = scale;
  T.nreg[kind] = (int)specs.size();
  for (size_t r = 0; r < specs.size(); ++r) {
    const RegionSpec& sp = specs[r];
    WRegion& reg = T.reg[kind][r];
    reg.s_max = (double)sp.s_max;
    reg.a2 = (double)sp.a2;
    reg.sgn = sp.var == 0 ? 1.0 : -1.0;
    reg.dyadic = sp.dyadic;
    reg.row0 = (int)(T.rows.size() / WT_ROW);
    reg.pad = 0;
    auto g = [&](ld x) { return f(sp.var == 0 ? x * x : sp.a2 - x * x); };
    std::vector<std::pair<ld, ld>> ivals;
    if (sp.dyadic) {
      ivals.push_back({0.0L, ldexpl(1.0L, -WT_DYADIC_KMIN)});
      for (int k = WT_DYADIC_KMIN; k >= 1; --k) {
        const ld lo = ldexpl(1.0L, -k), w = lo / WT_DYADIC_SUB;
        for (int j = 0; j < WT_DYADIC_SUB; ++j)
          if (lo + j * w < sp.x_hi) ivals.push_back({lo + j * w, lo + (j + 1) * w});
      }
      reg.x_lo = 0.0;
      reg.inv_w = 0.0;
    } else {
      const ld w = (sp.x_hi - sp.x_lo) / sp.n;
      for (int j = 0; j < sp.n; ++j) ivals.push_back({sp.x_lo + j * w, sp.x_lo + (j + 1) * w});
      reg.x_lo = (double)sp.x_lo;
      reg.inv_w = (double)(1.0L / w);
    }
    reg.n_int = (int)ivals.size();
    for (auto& iv : ivals) {
      T.rows.resize(T.rows.size() + WT_ROW);
      fit_interval(g, iv.first, iv.second, T.rows.data() + T.rows.size() - WT_ROW);
    }
  }
  // verify on a dense sample of s over the whole support
  const ld f0 = f(0.0L);
  const ld s_end = specs.back().s_max;
  double worst = 0.0;
  for (int i = 0; i <= 20000; ++i) {
    const ld u = (ld)i / 20000.0L;
    const ld s = (i % 2 ? u : u * u * u) * s_end * (1.0L - 1e-12L);  // dense near 0 too
    const double got = host_wtab_eval(T, kind, (double)(s / scale));
    const double err = (double)(fabsl((ld)got - f((ld)(double)(s / scale) * scale)) / f0);
    if (err > worst) worst = err;
  }
  T.max_err[kind] = worst;
}

static HostTables build_kernel_tables() {
  HostTables T;
  for (int k = 0; k < WT_KINDS; ++k) {
    T.nreg[k] = 0;
    T.scale[k] = 1.0;
    T.max_err[k] = 0.0;
    for (int r = 0; r < WT_MAX_REGIONS; ++r) T.reg[k][r] = WRegion{0, 0, 1, 0, 0, 0, 1, 0, 0};
  }
  const ld x_in = 0.625L;  // inner (dyadic) regions reach x = 0.625
  build_kind(T, MTN_KERNEL_WENDLANDC2, 1.0, F_wendland_c2,
             {{x_in * x_in, 0, 0.0L, 1, 0.0L, x_in, 0},
              {1.0L, 1, 1.0L, 0, 0.0L, sqrtl(1.0L - x_in * x_in), 24}});
  build_kind(T, MTN_KERNEL_CUBICSPLINE, 4.0, F_cubic_spline,
             {{x_in * x_in, 0, 0.0L, 1, 0.0L, x_in, 0},
              {1.0L, 1, 1.0L, 0, 0.0L, sqrtl(1.0L - x_in * x_in), 24},
              {2.56L, 0, 0.0L, 0, 1.0L, 1.6L, 32},
              {4.0L, 1, 4.0L, 0, 0.0L, 1.2L, 48}});
  return T;
}

// erf: Taylor coefficients about the interval centres,
// erf^(k)(x) = (2/sqrt(pi)) (-1)^(k-1) H_(k-1)(x) exp(-x^2), H = physicists' Hermite.
static void build_erf_table(double* tab) {
  const ld two_over_sqrt_pi = 1.1283791670955125738961589031215452L;
  for (int i = 0; i < ERF_NINT; ++i) {
    const ld c = ((ld)i + 0.5L) / ERF_INV_W;
    const ld ex = expl(-c * c);
    ld h_prev = 0.0L, h = 1.0L, fact = 1.0L;  // H_(-1) := 0, H_0 = 1
    tab[i * ERF_NCOEF + 0] = (double)erfl(c);
    for (int k = 1; k <= ERF_DEG; ++k) {
      fact *= k;
      const ld sign = ((k - 1) & 1) ? -1.0L : 1.0L;
      tab[i * ERF_NCOEF + k] = (double)(two_over_sqrt_pi * sign * h * ex / fact);
      const ld h_next = 2.0L * c * h - 2.0L * (k - 1) * h_prev;
      h_prev = h;
      h = h_next;
    }
  }
}

}  // namespace mtn
