// nm_lane.cuh -- the per-position work of ONE GPU lane (thread), host/device.
//
// A lane owns one genomic position: it sorts each group's <= 128 event means with a
// register-resident compare-exchange network, then merge-walks the two sorted groups once to
// get the exact integer KS numerator max|c0*n1 - c1*n0| (scipy-1.2.1 ks_2samp, call site
// bin/scripts/myDetect.py:341), and -- from the same walk -- the average-rank sum and tie term
// of mannwhitneyu (:331).  Welch moments (:335) come from the unsorted registers.
// The code is __host__ __device__ so that tests/host_emul compiles it with g++ and checks it
// against the oracle on the CPU; the product runs it only inside nm_lane_kernel.
#pragma once

#include "nm_math.cuh"

#define NM_LANE_MAX_N 128  // largest per-group coverage handled by the lane tier
#define NM_LANE_STEP 8     // network sizes are multiples of this

template <int N>
struct nm_sortnet;

#define NM_CE(i, j)                    \
  {                                    \
    const T lo_ = nm_min(x[i], x[j]);  \
    const T hi_ = nm_max(x[i], x[j]);  \
    x[i] = lo_;                        \
    x[j] = hi_;                        \
  }
NM_HD float nm_min(float a, float b) { return fminf(a, b); }
NM_HD float nm_max(float a, float b) { return fmaxf(a, b); }
#include "nm_sortnet.inc"
#undef NM_CE

// Per-position integer/moment results before the fp64 tails.
struct nm_lane_acc {
  int dnum;      // max over pooled points of |c0*n1 - c1*n0|
  int r2;        // 2 * (sum of average ranks of group 0)
  int tie;       // sum over pooled tie groups of t^3 - t
  double mean0, var0, mean1, var1;  // ddof=1 variances (numpy two-pass form)
};

// Two-pass mean / ddof=1 variance of x[0..n) (np.mean, np.var(ddof=1)); fp64 accumulation.
NM_HD void nm_moments(const float* x, int n, double* mean_out, double* var_out) {
  double s = 0.0;
#pragma unroll 4
  for (int k = 0; k < n; ++k) s += (double)x[k];
  const double mean = s / (double)n;
  double ss = 0.0;
#pragma unroll 4
  for (int k = 0; k < n; ++k) {
    const double dlt = (double)x[k] - mean;
    ss += dlt * dlt;
  }
  *mean_out = mean;
  *var_out = ss / (double)(n - 1);
}

// One merge-walk over the two sorted groups.  A(i)/B(j) return sorted element i/j, and +inf
// for i >= n0 / j >= n1 (network padding + one sentinel slot).  The ECDF difference is
// evaluated only where the next pooled value is strictly larger, i.e. after a whole tie
// group has been consumed from BOTH samples -- this is searchsorted(side='right') of the
// reference.  tmax >= n0+n1 is the (warp-uniform) trip count.
template <bool WANT_U, class AccA, class AccB>
NM_HD void nm_merge_walk(int n0, int n1, int tmax, const AccA& A, const AccB& B, nm_lane_acc* acc) {
  int i = 0, j = 0;
  float va = A(0), vb = B(0);
  float v = fminf(va, vb);
  int dmax = 0;
  int g = 0, ig = 0, r2 = 0, tie = 0;
  const int T = n0 + n1;
  for (int s = 0; s < tmax; ++s) {
    const bool act = s < T;
    const bool le = va <= vb;
    i += (act && le) ? 1 : 0;
    j += (act && !le) ? 1 : 0;
    va = A(i);
    vb = B(j);
    const float vn = fminf(va, vb);
    const bool endg = act && (vn > v);
    v = vn;
    int d = i * n1 - j * n0;
    d = d < 0 ? -d : d;
    if (endg) {
      dmax = d > dmax ? d : dmax;
      if (WANT_U) {
        const int tc = i + j;
        const int t = tc - g;
        r2 += (i - ig) * (g + tc + 1);
        tie += t * (t * t - 1);
        g = tc;
        ig = i;
      }
    }
  }
  acc->dnum = dmax;
  acc->r2 = r2;
  acc->tie = tie;
}

// Per-row outputs of the test stage ([(U,pU),(t,pt),(D,pks)], clamped).
struct nm_row_out {
  int dnum;
  double ks_d, ks_p;
  long long two_u;
  double u_stat, u_p;
  double t_stat, t_p;
  int flags;  // bit0: all pooled values identical (mannwhitneyu would raise in the reference)
};

NM_HD void nm_lane_finish(const nm_lane_acc& acc, int n0, int n1, bool want_u, bool want_t,
                          nm_row_out* o) {
  o->dnum = acc.dnum;
  nm_ks_tail(acc.dnum, n0, n1, &o->ks_d, &o->ks_p);
  o->flags = 0;
  o->two_u = 0;
  o->u_stat = o->u_p = o->t_stat = o->t_p = 0.0;
  if (want_u) {
    int64_t two_u;
    int flag;
    nm_mwu_tail(acc.r2, acc.tie, n0, n1, &o->u_stat, &two_u, &o->u_p, &flag);
    o->two_u = two_u;
    o->flags |= flag;
  }
  if (want_t) nm_welch_tail(acc.mean0, acc.var0, n0, acc.mean1, acc.var1, n1, &o->t_stat, &o->t_p);
}

// Sliding-window combination of KS p-values for one row (SURVEY 8a A8;
// bin/scripts/myDetect.py:379-404).  W(k, &z, &lnp) yields norm.isf(p) and ln(p) of window
// slot k (k = -nb..nb), where p is the clamped KS p of that row, or p = 1.0 exactly -- i.e.
// z = -inf, lnp = 0 -- where pos_check (:366-371) fails.  w[|k|] are the Stouffer weights
// 100 / WeightsDif^|k| (:396-400), wnorm = ||w||_2 over the whole window.
template <class WAcc>
NM_HD void nm_combine_row(int nb, const double* w, double wnorm, const WAcc& W, bool want_fisher,
                          bool want_stouffer, double* f_stat, double* f_p, double* s_stat,
                          double* s_p) {
  double lnsum = 0.0, zsum = 0.0;
  for (int k = -nb; k <= nb; ++k) {
    double z, lnp;
    W(k, &z, &lnp);
    lnsum += lnp;
    zsum += w[k < 0 ? -k : k] * z;
  }
  if (want_fisher) {
    const double x2 = -2.0 * lnsum;
    *f_stat = nm_max_float(x2);
    *f_p = nm_min_float(nm_chi2_sf_even(x2, 2 * nb + 1));
  }
  if (want_stouffer) {
    const double z = zsum / wnorm;
    *s_stat = nm_max_float(z);
    *s_p = nm_min_float(nm_norm_sf(z));
  }
}

// Stouffer weights exactly as the reference builds them (repeated division, :396-400).
NM_HD double nm_build_weights(int nb, double weights_dif, double* w /* [nb+1] */) {
  w[0] = 100.0;
  for (int k = 1; k <= nb; ++k) w[k] = w[k - 1] / weights_dif;
  double s = w[0] * w[0];
  for (int k = 1; k <= nb; ++k) s += 2.0 * w[k] * w[k];
  return sqrt(s);
}

// Dispatch a runtime network size (multiple of NM_LANE_STEP, <= NM_LANE_MAX_N) to a template.
#define NM_DISPATCH_N(nsel, CALL) \
  switch (nsel) {                 \
    case 8: { CALL(8); } break;     \
    case 16: { CALL(16); } break;   \
    case 24: { CALL(24); } break;   \
    case 32: { CALL(32); } break;   \
    case 40: { CALL(40); } break;   \
    case 48: { CALL(48); } break;   \
    case 56: { CALL(56); } break;   \
    case 64: { CALL(64); } break;   \
    case 72: { CALL(72); } break;   \
    case 80: { CALL(80); } break;   \
    case 88: { CALL(88); } break;   \
    case 96: { CALL(96); } break;   \
    case 104: { CALL(104); } break; \
    case 112: { CALL(112); } break; \
    case 120: { CALL(120); } break; \
    default: { CALL(128); } break;  \
  }
