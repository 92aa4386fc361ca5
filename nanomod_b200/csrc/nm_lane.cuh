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

#include <string.h>

#include "nm_math.cuh"

#define NM_LANE_MAX_N 128  // largest per-group coverage handled by the lane tier
#define NM_LANE_STEP 4     // network sizes are multiples of this up to NM_LANE_FINE_MAX, of 8 beyond
#define NM_LANE_FINE_MAX 104

template <int N>
struct nm_sortnet;

// Sort keys.  Default: the float32 values themselves (FMNMX compare-exchange).  With
// -DNM_INT_KEYS the values are mapped to order-preserving int32 keys so that two thirds of the
// compare-exchanges can run as {min, a+b-min} with the additions issued as IMADs on the FMA
// pipe, which is otherwise idle (the ALU pipe is the bottleneck of the sort: tools/ubench.cu).
// Non-negative floats are their own bit pattern; a negative float of magnitude m maps to -m
// (0x80000000 - bits), so -0.0 and 0.0 share key 0 and stay tied as they are for scipy.  One
// compare + one predicated subtract per value.
#ifdef NM_INT_KEYS
typedef int nm_key;
NM_HD nm_key nm_make_key(float x) {
#if defined(__CUDA_ARCH__)
  const int k = __float_as_int(x);
#else
  int k;
  memcpy(&k, &x, sizeof(k));
#endif
  return k < 0 ? (int)(0x80000000u - (unsigned)k) : k;
}
// same value; the subtraction written as a multiply-add by a runtime -1 (FMA pipe, not ALU)
NM_HD nm_key nm_make_key(float x, int mone) {
#if defined(__CUDA_ARCH__)
  const int k = __float_as_int(x);
#else
  int k;
  memcpy(&k, &x, sizeof(k));
#endif
  return k < 0 ? k * mone + (int)((unsigned)mone << 31) : k;  // -k + INT_MIN as one IMAD
}
#define NM_KEY_PINF 0x7f800000
#define NM_KEY_NINF ((int)0x80800000)  // key of -inf
NM_HD int nm_min(int a, int b) { return a < b ? a : b; }
NM_HD int nm_max(int a, int b) { return a > b ? a : b; }
#define NM_CEB(i, j) nm_ceb(x[i], x[j], one, mone);
#else
typedef float nm_key;
NM_HD nm_key nm_make_key(float x) { return x; }
NM_HD nm_key nm_make_key(float x, int) { return x; }
#define NM_KEY_PINF INFINITY
#define NM_KEY_NINF (-INFINITY)
#ifdef NM_FLOAT_IMAD
#define NM_CEB(i, j) nm_ceb(x[i], x[j], one, mone);
#else
#define NM_CEB(i, j) NM_CE(i, j)
#endif
#endif
NM_HD float nm_min(float a, float b) { return fminf(a, b); }
NM_HD float nm_max(float a, float b) { return fmaxf(a, b); }
#ifndef NM_INT_KEYS
NM_HD int nm_min(int a, int b) { return a < b ? a : b; }
NM_HD int nm_max(int a, int b) { return a > b ? a : b; }
#endif

// ------------------------------------------------------------------------------------------
// Grid keys.  The reference's event means are decimals with three places (norm_mean =
// round(x, 3), bin/scripts/myRefBaseSignalAnnotation.py:1108), cast to float32 by the packer.
// For such data k = round(1000 x) is an exact 16-bit image of the value: order and ties of the
// keys are order and ties of the floats.  Two keys -- the e-th element of group 0 in the low
// half, of group 1 in the high half -- share one register, and one compare-exchange network
// pass made of packed 16-bit min/max (VIMNMX.U16x2) sorts BOTH groups: half the instructions
// of the two float sorts.
//
// Nothing is assumed about the input.  Every value is checked while its key is made:
//   t = fma(x, 1000, M)       M = 1.5*2^23 + bias: the low 16 bits of t's pattern are the key
//   c = t - M                 = k, exactly
//   r = fma(x, -1000, c)      = k - 1000 x, exactly (a few bits, far below 24)
//   q = fma(r, 0.001f, x)     = RN(x + (k/1000 - x)): equals x iff x is the float nearest to k/1000
// so x passes iff x == fl32(k/1000) -- then x is a function of its key and the key map is
// injective on the values that pass -- and |x| <= 32.766 keeps k inside the 16 bits.  A tile with a
// value that fails (off-grid data, a huge value, NaN) is sorted by the float path instead.
// tests/host_emul checks this statement against ALL float32 values on the CPU (same source),
// tests/test_gpu_grid.py on the device.
// Key u = k + 32768 in [2, 65534]; 0 is the -inf sentinel, 0xffff the +inf padding.
// ------------------------------------------------------------------------------------------
#define NM_GRID_SCALE 1000.0f
#define NM_GRID_RCP 0.001f
#define NM_GRID_LIM 32.766f
#define NM_GRID_MA 12615680.0f  // 1.5*2^23 + 32768: pattern of t = 0x4B408000 + k
#define NM_GRID_MB 12596416.0f  // NM_GRID_MA - 0x4B40: (pattern of t) * 65536 + 0x4B40xxxx has u in the high half
#define NM_GRID_PAD_A 0x4B40FFFFu
#define NM_GRID_PAD_B 0x0000B4BFu
#define NM_GRID_TRIES 2        // a warp whose first tiles all failed the check, this many of them, raises the give-up flag
#define NM_GRID_SKIP_CALLS 15  // calls for which the grid-key launch is skipped after one that gave up
#define NM_GRID_NINF 0u
#define NM_GRID_PINF 0xffffffffu

NM_HD unsigned nm_f2u(float x) {
#if defined(__CUDA_ARCH__)
  return __float_as_uint(x);
#else
  unsigned k;
  memcpy(&k, &x, sizeof(k));
  return k;
#endif
}

// Key-carrying pattern of x for the low (M = NM_GRID_MA) or high (NM_GRID_MB) half; *bad is raised
// when x is not the float32 image of a grid point (the range test is the caller's: max |x|).
// nm_grid_flag accumulates the outcome over the values of a row: a predicate (FSETP chain, ALU
// pipe), or with -DNM_GRID_FSUM the sum of squares of q - x (FADD + FFMA, FMA pipes), which is
// zero iff every q == x (-0.0 passes either way: it is 0 to every comparison of the path; NaN fails).
#ifdef NM_GRID_FSUM
typedef float nm_grid_flag;
#define NM_GRID_FLAG0 0.0f
NM_HD bool nm_grid_failed(float f) { return !(f == 0.0f); }
#else
typedef bool nm_grid_flag;
#define NM_GRID_FLAG0 false
NM_HD bool nm_grid_failed(bool f) { return f; }
#endif
NM_HD unsigned nm_grid_bits(float x, float M, nm_grid_flag* bad) {
  const float t = fmaf(x, NM_GRID_SCALE, M);
  const float c = t - M;
  const float r = fmaf(x, -NM_GRID_SCALE, c);
  const float q = fmaf(r, NM_GRID_RCP, x);
#ifdef NM_GRID_FSUM
  const float e = q - x;
  *bad = fmaf(e, e, *bad);
#else
  *bad = *bad || (q != x);
#endif
  return nm_f2u(t);
}

// Two 16-bit keys in one register, compared half by half.
struct nm_p16 {
  unsigned v;
};
NM_HD nm_p16 nm_min(nm_p16 a, nm_p16 b) {
  nm_p16 r;
#if defined(__CUDA_ARCH__)
  asm("min.u16x2 %0, %1, %2;" : "=r"(r.v) : "r"(a.v), "r"(b.v));
#else
  const unsigned al = a.v & 0xffffu, bl = b.v & 0xffffu, ah = a.v >> 16, bh = b.v >> 16;
  r.v = (al < bl ? al : bl) | ((ah < bh ? ah : bh) << 16);
#endif
  return r;
}
NM_HD nm_p16 nm_max(nm_p16 a, nm_p16 b) {
  nm_p16 r;
#if defined(__CUDA_ARCH__)
  asm("max.u16x2 %0, %1, %2;" : "=r"(r.v) : "r"(a.v), "r"(b.v));
#else
  const unsigned al = a.v & 0xffffu, bl = b.v & 0xffffu, ah = a.v >> 16, bh = b.v >> 16;
  r.v = (al > bl ? al : bl) | ((ah > bh ? ah : bh) << 16);
#endif
  return r;
}
// {min, a + b - min}: the sum of the two halves' maxima is a + b - min as ONE 32-bit integer
// expression (per half min + max = a + b, so no borrow crosses the halves); two IMADs (FMA pipe)
NM_HD void nm_ceb(nm_p16& a, nm_p16& b, int one, int mone) {
  const nm_p16 lo = nm_min(a, b);
  const unsigned t = a.v * (unsigned)one + b.v;
  b.v = lo.v * (unsigned)mone + t;
  a = lo;
}
// "B" flavour of the compare-exchange: for integers {min, a + b - min} with the two additions
// written as multiply-adds by runtime +-1 (IMAD, FMA pipe); floats always use {min, max}
NM_HD void nm_ceb(int& a, int& b, int one, int mone) {
  const int lo = a < b ? a : b;
  const int t = a * one + b;
  b = lo * mone + t;
  a = lo;
}
#ifdef NM_FLOAT_IMAD
// Float keys, no key conversion: lo = fminf (FMNMX, ALU pipe); hi = a + b - lo on the raw bit
// patterns as two IMADs (FMA pipe) -- exact because fminf returns one of its operands bit for
// bit (inputs are finite or the +inf padding; -0.0 / +0.0 may come out in either order, which
// the walk's float compares treat as the tie it is).
NM_HD void nm_ceb(float& a, float& b, int one, int mone) {
  const float lo = fminf(a, b);
#if defined(__CUDA_ARCH__)
  const int ia = __float_as_int(a), ib = __float_as_int(b), il = __float_as_int(lo);
  const int t = ia * one + ib;
  b = __int_as_float(il * mone + t);
#else
  int ia, ib, il;
  memcpy(&ia, &a, 4);
  memcpy(&ib, &b, 4);
  memcpy(&il, &lo, 4);
  const int hi = (int)((unsigned)il * (unsigned)mone + ((unsigned)ia * (unsigned)one + (unsigned)ib));
  memcpy(&b, &hi, 4);
#endif
  a = lo;
}
#else
NM_HD void nm_ceb(float& a, float& b, int, int) {
  const float lo = fminf(a, b), hi = fmaxf(a, b);
  a = lo;
  b = hi;
}
#endif

#define NM_CE(i, j)                    \
  {                                    \
    const T lo_ = nm_min(x[i], x[j]);  \
    const T hi_ = nm_max(x[i], x[j]);  \
    x[i] = lo_;                        \
    x[j] = hi_;                        \
  }
// Flavour of comparator p (index mod 12): one in NM_CE_MIX uses {min,max} on the ALU pipe, the
// others {min, a+b-min} with IMADs.  1:2 keeps both pipes equally busy in isolation
// (tools/ubench.cu); the rest of the kernel is ALU-heavy, see profiles/round1_variants.md.
#ifndef NM_CE_MIX
#define NM_CE_MIX 3
#endif
// packed 16-bit pairs: their own mix; 0 = every comparator in the IMAD flavour.  Measured on the bench
// workload (lane kernel, tools/gpu_r2_u.sh): 1 -> 1.531 ms, 2 -> 1.508, 3 -> 1.521, 4 -> 1.568, 0 -> 2.09 (spills):
// one packed pass instead of two float32 sorts leaves the ALU pipe room for every second comparator.
#ifndef NM_CE_MIX_P16
#define NM_CE_MIX_P16 2
#endif
template <class T>
NM_HD constexpr bool nm_ce_is_a(int p) { return p % NM_CE_MIX == 0; }
template <>
NM_HD constexpr bool nm_ce_is_a<nm_p16>(int p) {
  return NM_CE_MIX_P16 > 0 && p % (NM_CE_MIX_P16 > 0 ? NM_CE_MIX_P16 : 1) == 0;
}
#define NM_CEP(p, i, j) NM_CEP_SEL(nm_ce_is_a<T>(p), i, j)
#define NM_CEP_SEL(isA, i, j)        \
  if constexpr (isA) NM_CE(i, j) else { NM_CEB(i, j) }
#include "nm_sortnet.inc"
#undef NM_CEP
#undef NM_CEP_SEL

// Looped, small-footprint sorts for the large size classes (tools/gen_sortloop.py).
template <int N>
struct nm_sortloop;
#define NM_NO_UNROLL _Pragma("unroll 1")
// rotate the register array left by Q: x[k] <- x[k+Q] (register moves; indices stay static)
#define NM_ROTATE(x, N, Q)                                       \
  {                                                              \
    T t_[Q];                                                     \
    _Pragma("unroll") for (int k_ = 0; k_ < (Q); ++k_) t_[k_] = x[k_];                  \
    _Pragma("unroll") for (int k_ = 0; k_ < (N) - (Q); ++k_) x[k_] = x[k_ + (Q)];        \
    _Pragma("unroll") for (int k_ = 0; k_ < (Q); ++k_) x[(N) - (Q) + k_] = t_[k_];       \
  }
#include "nm_sortloop.inc"
#undef NM_ROTATE

#undef NM_CE
#undef NM_CEB

// The sorter of a size class: flat network for N < 72 (code <= ~20 KB), looped sort above.
// run() leaves the k-th smallest key in x[order(k)].
// Looped sorts are used for the largest size classes only: their flat networks (>= 50 KB of
// code with int keys) are instruction-fetch bound, while for N <= 104 the register moves of the
// looped form cost more than the fetch stalls they remove (profiles/round1_variants.md).
#ifndef NM_LOOPED_MIN_N
#define NM_LOOPED_MIN_N 112
#endif
template <int N, bool LOOPED = (N >= NM_LOOPED_MIN_N)>
struct nm_sorter;
template <int N>
struct nm_sorter<N, false> {
  template <class T>
  NM_HD static void run(T (&x)[N], int one, int mone) { nm_sortnet<N>::run(x, one, mone); }
  NM_HD static constexpr int order(int k) { return k; }
};
template <int N>
struct nm_sorter<N, true> {
  template <class T>
  NM_HD static void run(T (&x)[N], int one, int mone) { nm_sortloop<N>::run(x, one, mone); }
  NM_HD static int order(int k) { return nm_sortloop<N>::order(k); }
};

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

// Merge walk over the two sorted groups, split into two independent dependency chains that
// meet in the middle (instruction-level parallelism: each chain is a serial pointer chase).
//
// Column layout (stride S floats between consecutive elements; S = 32 on the device, where a
// lane's sorted group is stored transposed so that data-dependent indexing never bank-conflicts):
//   col[0] = -inf | col[(k+1)*S] = k-th smallest key, k < n | +inf from row n+1 .. N+1.
// Forward chain: takes the pooled elements 0 .. T/2-1 in ascending order (ties: group 0 first).
// Backward chain: takes T-1 .. T/2 in descending order (ties: group 1 first) -- the mirror rule,
// so both chains describe the same pooled order and stop at the same split (i*, j*).
// The ECDF difference c0*n1 - c1*n0 is evaluated only at tie-group boundaries, i.e. where the
// next pooled value differs -- searchsorted(side='right') of scipy-1.2.1 ks_2samp.  For the rank
// statistics each chain closes the tie groups that lie entirely on its side; the group that
// straddles the split (if any) is closed once, after the loop.  iters >= ceil((n0+n1)/2) is the
// warp-uniform trip count.
// Column element type K -> the type its values are compared in (16-bit grid keys widen to int).
template <class K>
struct nm_walk_val {
  typedef K type;
};
template <>
struct nm_walk_val<unsigned short> {
  typedef int type;
};

// Load of a column element.  A 16-bit key is widened to int behind an empty asm: told that the
// values fit 16 bits, the compiler otherwise computes min / max with packed 16-bit operations and
// re-masks every result (two extra instructions per step).
template <class K>
NM_HD typename nm_walk_val<K>::type nm_walk_load(const K* p) {
  return *p;
}
template <>
NM_HD int nm_walk_load<unsigned short>(const unsigned short* p) {
  int v = *p;
#if defined(__CUDA_ARCH__)
  asm("" : "+r"(v));
#endif
  return v;
}

template <bool WANT_U, int S, class K>
NM_HD void nm_merge_walk(const K* colA, const K* colB, int n0, int n1, int iters,
                         nm_lane_acc* acc) {
  typedef typename nm_walk_val<K>::type V;
  const int T = n0 + n1, T1 = T >> 1, T2 = T - T1;
  const K* fa = colA + S;
  const K* fb = colB + S;
  V va = nm_walk_load(fa), vb = nm_walk_load(fb), v = nm_min(va, vb);
  const K* ba = colA + n0 * S;
  const K* bb = colB + n1 * S;
  V ea = nm_walk_load(ba), eb = nm_walk_load(bb), w = nm_max(ea, eb);
  int df = 0, db = 0, dmax = 0;
  int fi = 0, g = 0, ig = 0;    // forward: group-0 count, open group start (count, group-0 count)
  int bi = n0, h = T, ih = n0;  // backward: group-0 count below, open group end (count, group-0 count)
  int r2 = 0, tie = 0;
  for (int s = 0; s < iters; ++s) {
    {
      const bool act = s < T1;
      const bool le = va <= vb;
      const bool ta = act && le, tb = act && !le;
      fa += ta ? S : 0;
      fb += tb ? S : 0;
      df += ta ? n1 : 0;
      df -= tb ? n0 : 0;
      va = nm_walk_load(fa);
      vb = nm_walk_load(fb);
      const V vn = nm_min(va, vb);
      const bool q = act && (vn > v);
      v = vn;
      const int ad = df < 0 ? -df : df;
      dmax = (q && ad > dmax) ? ad : dmax;
      if (WANT_U) {
        fi += ta ? 1 : 0;
        const int tc = s + 1;
        const int t = q ? tc - g : 0;
        const int ca = q ? fi - ig : 0;
        r2 += ca * (g + tc + 1);
        tie += t * (t * t - 1);
        g = q ? tc : g;
        ig = q ? fi : ig;
      }
    }
    {
      const bool act = s < T2;
      const bool ge = eb >= ea;
      const bool tb = act && ge, ta = act && !ge;
      bb -= tb ? S : 0;
      ba -= ta ? S : 0;
      db += tb ? n0 : 0;
      db -= ta ? n1 : 0;
      ea = nm_walk_load(ba);
      eb = nm_walk_load(bb);
      const V wn = nm_max(ea, eb);
      const bool q = act && (wn < w);
      w = wn;
      const int ad = db < 0 ? -db : db;
      dmax = (q && ad > dmax) ? ad : dmax;
      if (WANT_U) {
        bi -= ta ? 1 : 0;
        const int tc = T - (s + 1);
        const int t = q ? h - tc : 0;
        const int ca = q ? ih - bi : 0;
        r2 += ca * (tc + h + 1);
        tie += t * (t * t - 1);
        h = q ? tc : h;
        ih = q ? bi : ih;
      }
    }
  }
  if (WANT_U) {  // the tie group straddling the split: [g, h); empty (t = 0) when both closed
    const int t = h - g;
    r2 += (ih - ig) * (g + h + 1);
    tie += t * (t * t - 1);
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
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
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

// ------------------------------------------------------------------------------------------
// Plain-C++ statement of the KS-only fast walks of nm_lane_kernel.cu (nm_walk_ks_fast and
// nm_walk_ks_fast4 are spelled in inline PTX there).  Same chains, same start states, same
// evaluation points, same requirements on the trip count -- kept here so that the algorithm is
// checked against the oracle on the CPU (tests/test_host_emul.py); the kernel does not call it.
// Columns with stride S: row 0 = -inf, rows 1..n = sorted keys, rows n+1.. = +inf.
//   forward step : take the smaller head (ties: group 0), evaluate |i*T - m*n0| (i = group-0
//                  elements taken, m = all elements taken) when the next pooled value is larger
//   backward step: take the larger tail (ties: group 1), evaluate |i*T - r*n0| (i, r = group-0 /
//                  all elements remaining) when the next pooled value below is smaller
// ------------------------------------------------------------------------------------------
template <class V>
struct nm_chain {
  int ia, ib;  // forward: heads (rows ia+1 / ib+1); backward: tails (rows ia / ib)
  V va, vb;    // their values
  V v;         // value of the element taken last (start: see nm_walk_ks4)
  int m;       // forward: elements taken so far; backward: elements remaining
};

template <int S, class K, class V>
NM_HD void nm_chain_fwd(nm_chain<V>& c, const K* colA, const K* colB, int n0, int T, int* dmax) {
  const bool p = c.va <= c.vb;
  if (p) { ++c.ia; c.va = colA[(c.ia + 1) * S]; } else { ++c.ib; c.vb = colB[(c.ib + 1) * S]; }
  ++c.m;
  const V vn = nm_min(c.va, c.vb);
  const bool q = vn > c.v;
  c.v = vn;
  int d = c.ia * T - c.m * n0;
  d = d < 0 ? -d : d;
  if (q && d > *dmax) *dmax = d;
}

template <int S, class K, class V>
NM_HD void nm_chain_bwd(nm_chain<V>& c, const K* colA, const K* colB, int n0, int T, int* dmax) {
  const bool p = c.vb >= c.va;
  if (p) { --c.ib; c.vb = colB[c.ib * S]; } else { --c.ia; c.va = colA[c.ia * S]; }
  --c.m;
  const V wn = nm_max(c.va, c.vb);
  const bool q = wn < c.v;
  c.v = wn;
  int d = c.ia * T - c.m * n0;
  d = d < 0 ? -d : d;
  if (q && d > *dmax) *dmax = d;
}

// two chains meeting in the middle; iters >= ceil(T/2) and iters <= T
template <int S, class K>
NM_HD int nm_walk_ks2(const K* colA, const K* colB, int n0, int n1, int iters) {
  typedef typename nm_walk_val<K>::type V;
  const int T = n0 + n1;
  nm_chain<V> f = {0, 0, colA[S], colB[S], nm_min((V)colA[S], (V)colB[S]), 0};
  nm_chain<V> b = {n0, n1, colA[n0 * S], colB[n1 * S], nm_max((V)colA[n0 * S], (V)colB[n1 * S]), T};
  int dmax = 0;
  for (int s = 0; s < iters; ++s) {
    nm_chain_fwd<S>(f, colA, colB, n0, T, &dmax);
    nm_chain_bwd<S>(b, colA, colB, n0, T, &dmax);
  }
  return dmax;
}

// four chains: merge-path split of the pooled order at h = T/2 (ties: group 0 first), a forward
// and a backward chain per half; 2*it >= ceil(T/2), it <= T/2, 2^search_iters > max(n0, n1)
template <int S, class K>
NM_HD int nm_walk_ks4(const K* colA, const K* colB, int n0, int n1, int it, int search_iters) {
  typedef typename nm_walk_val<K>::type V;
  const int T = n0 + n1, h = T >> 1;
  int lo = h - n1 > 0 ? h - n1 : 0, hi = h < n0 ? h : n0;
  for (int k = 0; k < search_iters; ++k) {
    const int mid = (lo + hi) >> 1;
    const bool P = colB[(h - mid) * S] < colA[(mid + 1) * S];
    hi = P ? mid : hi;
    lo = P ? lo : mid + 1;
  }
  const int is = lo, js = h - lo;
  nm_chain<V> f1 = {0, 0, colA[S], colB[S], nm_min((V)colA[S], (V)colB[S]), 0};
  nm_chain<V> b1 = {is, js, colA[is * S], colB[js * S], nm_max((V)colA[is * S], (V)colB[js * S]), h};
  nm_chain<V> f2 = {is, js, colA[(is + 1) * S], colB[(js + 1) * S], nm_min((V)colA[(is + 1) * S], (V)colB[(js + 1) * S]), h};
  nm_chain<V> b2 = {n0, n1, colA[n0 * S], colB[n1 * S], nm_max((V)colA[n0 * S], (V)colB[n1 * S]), T};
  int dmax = 0;
  {  // the boundary between pooled elements h-1 and h belongs to neither half's chains
    int dj = is * T - h * n0;
    dj = dj < 0 ? -dj : dj;
    if (b1.v < f2.v) dmax = dj;
  }
  for (int s = 0; s < it; ++s) {
    nm_chain_fwd<S>(f1, colA, colB, n0, T, &dmax);
    nm_chain_bwd<S>(b1, colA, colB, n0, T, &dmax);
    nm_chain_fwd<S>(f2, colA, colB, n0, T, &dmax);
    nm_chain_bwd<S>(b2, colA, colB, n0, T, &dmax);
  }
  return dmax;
}

// Network size (class) for a longest row of n values: 8, 12, ..., 104, 112, 120, 128.
NM_HD constexpr int nm_lane_class(int n) {
  return n <= 8 ? 8 : n <= NM_LANE_FINE_MAX ? (n + NM_LANE_STEP - 1) / NM_LANE_STEP * NM_LANE_STEP : (n + 7) / 8 * 8;
}

// Size group of a longest row of n values: the lane tier launches once per group when a call
// mixes them (<= 64: 12+ warps/SM; <= 104: straight-line networks; <= 128: looped networks).
NM_HD constexpr int nm_lane_group(int n) {
  return nm_lane_class(n) <= 64 ? 0 : nm_lane_class(n) <= NM_LANE_FINE_MAX ? 1 : 2;
}

// Dispatch a runtime network size (a value of nm_lane_class) to a template.
#ifdef NM_ONLY_N  // analysis builds: a single network size (SASS inspection with cuobjdump)
#define NM_DISPATCH_N(nsel, CALL) { CALL(NM_ONLY_N); }
#else
#define NM_DISPATCH_N(nsel, CALL) \
  switch (nsel) {                 \
    case 8: { CALL(8); } break;   \
    case 12: { CALL(12); } break; \
    case 16: { CALL(16); } break; \
    case 20: { CALL(20); } break; \
    case 24: { CALL(24); } break; \
    case 28: { CALL(28); } break; \
    case 32: { CALL(32); } break; \
    case 36: { CALL(36); } break; \
    case 40: { CALL(40); } break; \
    case 44: { CALL(44); } break; \
    case 48: { CALL(48); } break; \
    case 52: { CALL(52); } break; \
    case 56: { CALL(56); } break; \
    case 60: { CALL(60); } break; \
    case 64: { CALL(64); } break; \
    case 68: { CALL(68); } break; \
    case 72: { CALL(72); } break; \
    case 76: { CALL(76); } break; \
    case 80: { CALL(80); } break; \
    case 84: { CALL(84); } break; \
    case 88: { CALL(88); } break; \
    case 92: { CALL(92); } break; \
    case 96: { CALL(96); } break; \
    case 100: { CALL(100); } break;\
    case 104: { CALL(104); } break;\
    case 112: { CALL(112); } break;\
    case 120: { CALL(120); } break;\
    default: { CALL(128); } break;  \
  }
#endif
