// nm_math.cuh -- fp64 tail functions of the per-position tests, host/device.
//
// Each function restates the scipy/cephes formula the reference reaches through its scipy
// 1.2.1 calls (reference call sites: bin/scripts/myDetect.py:331,335,341,393,401; clamps
// :317-325).  They are __host__ __device__ so that tests/host_emul can compile this very file
// with g++ and compare it with scipy on the CPU; the product only ever runs them on the GPU.
#pragma once

#include <float.h>
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define NM_HD __host__ __device__ __forceinline__
#else
#define NM_HD inline
#endif

#define NM_PI 3.14159265358979323846
#define NM_SQRT2PI 2.50662827463100050242
#define NM_SQRT1_2 0.70710678118654752440

// myDetect.py:317-320 -- p < DBL_MIN -> DBL_MIN; NaN passes through.
NM_HD double nm_min_float(double p) { return (p < DBL_MIN) ? DBL_MIN : p; }
// myDetect.py:322-325 -- stat > DBL_MAX (i.e. +inf) -> DBL_MAX; -inf and NaN pass through.
NM_HD double nm_max_float(double s) { return (s > DBL_MAX) ? DBL_MAX : s; }

// Kolmogorov survival function, scipy.special.kolmogorov == stats.kstwobign.sf (the p-value
// of scipy-1.2.1 ks_2samp).  Two regimes split at 0.82: Jacobi-theta form for small x,
// alternating series for large x; both truncated where the next term is < 1e-17 relative.
NM_HD double nm_kolmogorov_sf(double x) {
  if (x != x) return x;
  if (x <= 0.040611972203751713) return 1.0;  // also covers x <= 0
  double sf;
  if (x <= 0.82) {
    const double w = NM_SQRT2PI / x;
    const double logu8 = -(NM_PI * NM_PI) / (x * x);  // log(u^8), u = exp(-pi^2/(8 x^2))
    const double u = exp(logu8 * 0.125);
    const double u8 = exp(logu8);
    double P = 1.0;
    P = 1.0 + (u8 * u8 * u8) * P;  // u^48
    P = 1.0 + (u8 * u8) * P;       // u^24
    P = 1.0 + u8 * P;              // u^8
    P = w * u * P;                 // cdf
    sf = 1.0 - P;
  } else {
    const double v = exp(-2.0 * x * x);
    const double v3 = v * v * v;
    double P = 1.0;
    P = 1.0 - (v3 * v3 * v) * P;  // v^7
    P = 1.0 - (v3 * v * v) * P;   // v^5
    P = 1.0 - v3 * P;             // v^3
    sf = 2.0 * v * P;
  }
  if (sf < 0.0) sf = 0.0;
  if (sf > 1.0) sf = 1.0;
  return sf;
}

// scipy-1.2.1 ks_2samp p-value from the exact integer numerator.
// D is formed the way the reference's table needs it (a double); p is clamped (:342).
NM_HD void nm_ks_tail(int64_t dnum, int64_t n0, int64_t n1, double* d_out, double* p_out) {
  const double d = (double)dnum / ((double)n0 * (double)n1);
  const double en = sqrt((double)(n0 * n1) / (double)(n0 + n1));
  const double p = nm_kolmogorov_sf((en + 0.12 + 0.11 / en) * d);
  *d_out = nm_max_float(d);
  *p_out = nm_min_float(p);
}

// Standard normal survival function norm.sf(z) = ndtr(-z) (cephes ndtr = erfc form).
NM_HD double nm_norm_sf(double z) { return 0.5 * erfc(z * NM_SQRT1_2); }

// norm.isf(p) = -ndtri(p).  Device: CUDA normcdfinv.  Host (test harness only): Newton on
// erfc from a rational starting point, since libm has no inverse normal.
NM_HD double nm_norm_isf(double p) {
#if defined(__CUDA_ARCH__)
  return -normcdfinv(p);
#else
  if (p != p) return p;
  if (p <= 0.0) return INFINITY;
  if (p >= 1.0) return -INFINITY;
  const bool upper = p > 0.5;
  const double q = upper ? 1.0 - p : p;  // solve sf(z) = q, z >= 0
  if (upper && q < 1e-9) {
    // 1-p cancels; the harness never needs this regime to better than ~1e-7
  }
  double t = sqrt(-2.0 * log(q));
  double z = t - (2.515517 + 0.802853 * t + 0.010328 * t * t) /
                     (1.0 + 1.432788 * t + 0.189269 * t * t + 0.001308 * t * t * t);
  for (int it = 0; it < 60; ++it) {
    const double f = nm_norm_sf(z) - q;
    const double pdf = exp(-0.5 * z * z) / NM_SQRT2PI;
    if (pdf == 0.0) break;
    double step = f / pdf;
    // Halley correction keeps the iteration stable far in the tail
    step = step / (1.0 - 0.5 * z * step);
    z += step;
    if (fabs(step) <= 1e-16 * fabs(z)) break;
  }
  return upper ? -z : z;
#endif
}

// Mann-Whitney tail, scipy-1.2.1 mannwhitneyu(use_continuity=True, alternative=None)
//  from exact integers: r2 = 2 * (sum of average ranks of group 0),
// tie = sum over pooled tie groups of t^3 - t.  Returns the clamped one-sided p, U = min(u1,u2)
// and 2U (exact integer).  flag = 1 when every pooled value is identical (T == 0): the
// reference raises ValueError there; the defined behaviour here is p = NaN.
NM_HD void nm_mwu_tail(int64_t r2, int64_t tie, int64_t n0, int64_t n1, double* u_out,
                       int64_t* two_u_out, double* p_out, int* flag_out) {
  const int64_t two_u1 = 2 * n0 * n1 + n0 * (n0 + 1) - r2;
  const int64_t two_u2 = 2 * n0 * n1 - two_u1;
  const int64_t two_u = two_u1 < two_u2 ? two_u1 : two_u2;
  const int64_t two_big = two_u1 < two_u2 ? two_u2 : two_u1;
  const double n = (double)(n0 + n1);
  *two_u_out = two_u;
  *u_out = nm_max_float(0.5 * (double)two_u);
  const double T = 1.0 - (double)tie / (n * n * n - n);
  if (T == 0.0) {
    *p_out = NAN;
    *flag_out = 1;
    return;
  }
  const double sd = sqrt(T * (double)n0 * (double)n1 * (n + 1.0) / 12.0);
  const double meanrank = (double)(n0 * n1) / 2.0 + 0.5;
  const double z = (0.5 * (double)two_big - meanrank) / sd;
  *p_out = nm_min_float(nm_norm_sf(fabs(z)));
  *flag_out = 0;
}

// Continued fraction of the regularised incomplete beta function, evaluated by the forward
// recurrence on numerators / denominators with a renormalisation per step (two divisions per
// step; the modified-Lentz form needs six).  Converges for x < (a+1)/(a+b+2).
NM_HD double nm_betacf(double a, double b, double x) {
  const double qab = a + b, qap = a + 1.0, qam = a - 1.0;
  double am = 1.0, bm = 1.0, az = 1.0, bz = 1.0 - qab * x / qap;
  for (int m = 1; m <= 5000; ++m) {
    const double em = (double)m, tem = em + em;
    const double f0 = qam + tem, f1 = a + tem, f2 = qap + tem;
    const double inv = 1.0 / (f0 * f1 * f2);
    const double d1 = em * (b - em) * x * f2 * inv;
    const double d2 = -(a + em) * (qab + em) * x * f0 * inv;
    const double ap = az + d1 * am, bp = bz + d1 * bm;
    const double app = ap + d2 * az, bpp = bp + d2 * bz;
    const double aold = az;
    const double rb = 1.0 / bpp;
    am = ap * rb;
    bm = bp * rb;
    az = app * rb;
    bz = 1.0;
    if (fabs(az - aold) < 4e-16 * fabs(az)) break;
  }
  return az;
}

// ln Gamma(z + 1/2) - ln Gamma(z) for z > 0: asymptotic series in 1/z (six terms: 2e-16 for
// z >= 12), reached from smaller z by the recurrence Gamma(z+1) = z Gamma(z).  Replaces two
// lgamma calls in the Student-t tail.
NM_HD double nm_lgamma_half_ratio(double z) {
  double num = 1.0, den = 1.0;
  while (z < 12.0) {
    num *= z + 0.5;
    den *= z;
    z += 1.0;
  }
  const double r = 1.0 / z, r2 = r * r;
  const double s = r * (-1.0 / 8.0 + r2 * (1.0 / 192.0 + r2 * (-1.0 / 640.0 + r2 * (17.0 / 14336.0 +
                   r2 * (-31.0 / 18432.0 + r2 * (691.0 / 180224.0))))));
  return 0.5 * log(z) + s - log(num / den);
}

// Two-sided Student-t p-value 2 * t.sf(|t|, df) = I_{df/(df+t^2)}(df/2, 1/2)
// (scipy: 2 * special.stdtr(df, -|t|)), evaluated in log space so that p stays accurate down
// to the denormal range.
NM_HD double nm_student_t_two_sided(double t, double df) {
  if (t != t || df != df) return NAN;
  const double t2 = t * t;
  if (t2 == 0.0) return 1.0;
  if (isinf(t2)) return 0.0;
  const double a = 0.5 * df, b = 0.5;
  const double x = df / (df + t2);  // small when |t| is large
  const double y = t2 / (df + t2);  // = 1 - x, small when |t| is small
  // ln B(a, 1/2) = ln Gamma(a) + ln Gamma(1/2) - ln Gamma(a + 1/2)
  const double lnbeta = 0.5723649429247000870717 /* ln sqrt(pi) */ - nm_lgamma_half_ratio(a);
  if (x < (a + 1.0) / (a + b + 2.0)) {
    const double lnpre = a * log(x) + b * log(y) - lnbeta;
    return exp(lnpre) * nm_betacf(a, b, x) / a;
  }
  const double lnpre = b * log(y) + a * log1p(-y) - lnbeta;
  const double q = exp(lnpre) * nm_betacf(b, a, y) / b;
  const double p = 1.0 - q;
  return p < 0.0 ? 0.0 : p;
}

// Welch test, scipy-1.2.1 ttest_ind(equal_var=False) from the two groups'
// means and ddof=1 variances.  t is signed (group0 - group1); outputs are clamped (:336-337).
NM_HD void nm_welch_tail(double mean0, double var0, int64_t n0, double mean1, double var1,
                         int64_t n1, double* t_out, double* p_out) {
  const double vn0 = var0 / (double)n0;
  const double vn1 = var1 / (double)n1;
  double df = (vn0 + vn1) * (vn0 + vn1) /
              (vn0 * vn0 / (double)(n0 - 1) + vn1 * vn1 / (double)(n1 - 1));
  if (df != df) df = 1.0;
  const double denom = sqrt(vn0 + vn1);
  const double t = (mean0 - mean1) / denom;
  const double p = nm_student_t_two_sided(fabs(t), df);
  *t_out = nm_max_float(t);
  *p_out = nm_min_float(p);
}

// chi2.sf(x2, 2k) = special.chdtrc(2k, x2) for even degrees of freedom (closed form):
// exp(-h) * sum_{i<k} h^i / i!, h = x2/2.
NM_HD double nm_chi2_sf_even(double x2, int k) {
  if (x2 != x2) return x2;
  if (x2 <= 0.0) return 1.0;
  const double h = 0.5 * x2;
  double term = 1.0, sum = 1.0;
  for (int i = 1; i < k; ++i) {
    term *= h / (double)i;
    sum += term;
  }
  if (h < 700.0) return exp(-h) * sum;
  return exp(log(sum) - h);
}
