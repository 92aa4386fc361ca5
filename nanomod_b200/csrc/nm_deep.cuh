// nm_deep.cuh -- per-element formulation of the same statistics for deep pileups, host/device.
//
// The block-per-position kernel sorts both groups in shared memory and then lets every
// thread take pooled elements x and count, by binary search in both sorted groups,
//   ua = #group0 <= x, ub = #group1 <= x   (searchsorted(side='right'), scipy-1.2.1 ks_2samp,
//                                            reference call site bin/scripts/myDetect.py:341)
//   la = #group0 <  x, lb = #group1 <  x   (only for the rank statistics, :331)
// The KS numerator is max_x |ua*n1 - ub*n0| by definition; the pooled tie group of x spans
// ranks la+lb+1 .. ua+ub, so 2*avgrank = la+lb+ua+ub+1 and each member contributes t^2-1 to
// sum(t^3 - t).  No serial walk, no state crossing thread boundaries.
#pragma once

#include "nm_lane.cuh"

struct nm_deep_acc {
  long long dnum;
  long long r2;
  long long tie;
};

NM_HD void nm_deep_acc_init(nm_deep_acc* a) {
  a->dnum = 0;
  a->r2 = 0;
  a->tie = 0;
}

NM_HD void nm_deep_acc_merge(nm_deep_acc* a, const nm_deep_acc& b) {
  a->dnum = a->dnum > b.dnum ? a->dnum : b.dnum;
  a->r2 += b.r2;
  a->tie += b.tie;
}

// Views of a sorted array: plain, or skewed by one word per 32 (element i at i + i/32) -- the deep
// kernel stores its sorted groups skewed so that threads whose indices are a multiple of 8 apart
// (each holds 8 consecutive elements; each walks a piece of 16) hit 32 different banks.
struct nm_view_plain {
  const float* p;
  NM_HD float operator[](int i) const { return p[i]; }
};
struct nm_view_skew {
  const float* p;
  NM_HD float operator[](int i) const { return p[i + (i >> 5)]; }
};

// 16-bit grid keys (nm_lane.cuh): group 0 in the low halves, group 1 in the high halves of ONE sorted
// array of key pairs (skewed like nm_view_skew); a view picks a half.  0xffff is the +inf padding.
struct nm_view_skew16 {
  const unsigned* p;
  int sh;  // 0: group 0, 16: group 1
  NM_HD int operator[](int i) const { return (int)((p[i + (i >> 5)] >> sh) & 0xffffu); }
};

// number of elements of sorted s[0..n) that are <= x
template <class V, class X>
NM_HD int nm_count_le(const V& s, int n, X x) {
  int lo = 0, hi = n;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (s[mid] <= x) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// number of elements of sorted s[0..n) that are < x
template <class V, class X>
NM_HD int nm_count_lt(const V& s, int n, X x) {
  int lo = 0, hi = n;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (s[mid] < x) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// Contribution of pooled element e (e < n0: group 0, else group 1).
template <class V>
NM_HD void nm_deep_element(const V& sa, int n0, const V& sb, int n1, int e, bool want_u, nm_deep_acc* one) {
  const bool is_a = e < n0;
  const auto x = is_a ? sa[e] : sb[e - n0];
  const long long ua = nm_count_le(sa, n0, x);
  const long long ub = nm_count_le(sb, n1, x);
  long long d = ua * (long long)n1 - ub * (long long)n0;
  one->dnum = d < 0 ? -d : d;
  if (want_u) {
    const long long la = nm_count_lt(sa, n0, x);
    const long long lb = nm_count_lt(sb, n1, x);
    const long long lo = la + lb, hi = ua + ub, t = hi - lo;
    one->tie = t * t - 1;
    one->r2 = is_a ? (lo + hi + 1) : 0;
  }
}
NM_HD void nm_deep_element(const float* sa, int n0, const float* sb, int n1, int e, bool want_u, nm_deep_acc* one) {
  nm_deep_element(nm_view_plain{sa}, n0, nm_view_plain{sb}, n1, e, want_u, one);
}

NM_HD void nm_deep_finish(const nm_deep_acc& acc, int n0, int n1, bool want_u, bool want_t,
                          double mean0, double var0, double mean1, double var1, nm_row_out* o) {
  o->dnum = (int)acc.dnum;
  nm_ks_tail(acc.dnum, n0, n1, &o->ks_d, &o->ks_p);
  o->flags = 0;
  if (want_u) {
    int64_t two_u;
    int flag;
    nm_mwu_tail(acc.r2, acc.tie, n0, n1, &o->u_stat, &two_u, &o->u_p, &flag);
    o->two_u = two_u;
    o->flags |= flag;
  }
  if (want_t) nm_welch_tail(mean0, var0, n0, mean1, var1, n1, &o->t_stat, &o->t_p);
}
