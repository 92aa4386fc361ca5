// TEST HARNESS ONLY -- compiles the product's host/device headers (nm_math.cuh, nm_lane.cuh,
// nm_deep.cuh) with g++ so that the per-lane logic and the fp64 tails can be checked against
// the oracle on a machine without a GPU.  Nothing in nanomod_b200/ links or loads this file;
// the product path is CUDA-only.
#include <math.h>
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <thread>
#include <vector>

#include "../../nanomod_b200/csrc/nm_lane.cuh"
#include "../../nanomod_b200/csrc/nm_deep.cuh"

namespace {
volatile int g_one = 1, g_mone = -1;  // runtime constants, as on the device

template <int N>
void lane_position(const float* a, int n0, const float* b, int n1, bool want_u, bool want_t,
                   nm_row_out* out) {
  nm_lane_acc acc;
  memset(&acc, 0, sizeof(acc));
  std::vector<nm_key> sa(N + 2), sb(N + 2);
  {
    nm_key x[N];
    for (int k = 0; k < N; ++k) x[k] = k < n0 ? nm_make_key(a[k]) : (nm_key)NM_KEY_PINF;
    if (want_t) nm_moments(a, n0, &acc.mean0, &acc.var0);
    nm_sorter<N>::run(x, g_one, g_mone);
    sa[0] = NM_KEY_NINF;
    for (int k = 0; k < N; ++k) sa[k + 1] = x[nm_sorter<N>::order(k)];
    sa[N + 1] = NM_KEY_PINF;
  }
  {
    nm_key x[N];
    for (int k = 0; k < N; ++k) x[k] = k < n1 ? nm_make_key(b[k]) : (nm_key)NM_KEY_PINF;
    if (want_t) nm_moments(b, n1, &acc.mean1, &acc.var1);
    nm_sorter<N>::run(x, g_one, g_mone);
    sb[0] = NM_KEY_NINF;
    for (int k = 0; k < N; ++k) sb[k + 1] = x[nm_sorter<N>::order(k)];
    sb[N + 1] = NM_KEY_PINF;
  }
  // iters deliberately larger than needed: lanes of a warp share the longest trip count
  const int iters = (n0 + n1 + 1) / 2 + 3;
  if (want_u)
    nm_merge_walk<true, 1>(sa.data(), sb.data(), n0, n1, iters, &acc);
  else
    nm_merge_walk<false, 1>(sa.data(), sb.data(), n0, n1, iters, &acc);
  nm_lane_finish(acc, n0, n1, want_u, want_t, out);
}
}  // namespace

extern "C" {

// KS numerator by the two-chain / four-chain fast walks (plain-C++ statement of the kernel's PTX
// walks) on already sorted groups.  tmax >= n0 + n1 plays the longest row of the warp: the
// trip counts are derived from it exactly as nm_lane_kernel does.  Returns -1 if the walk's
// precondition (the kernel's `fast` test) does not hold for this row.
int emul_fast_walk(const float* a_sorted, int n0, const float* b_sorted, int n1, int tmax, int chains) {
  const int nmax = n0 > n1 ? n0 : n1;
  std::vector<nm_key> sa(nmax + 2 + 4, (nm_key)NM_KEY_PINF), sb(nmax + 2 + 4, (nm_key)NM_KEY_PINF);
  sa[0] = sb[0] = NM_KEY_NINF;
  for (int k = 0; k < n0; ++k) sa[k + 1] = nm_make_key(a_sorted[k]);
  for (int k = 0; k < n1; ++k) sb[k + 1] = nm_make_key(b_sorted[k]);
  const int T = n0 + n1;
  if (chains == 2) {
    const int iters = (tmax + 1) >> 1;
    if (T < iters) return -1;
    return nm_walk_ks2<1>(sa.data(), sb.data(), n0, n1, iters);
  }
  const int it4 = (tmax + 3) >> 2;
  if ((T >> 1) < it4) return -1;
  int bits = 0;
  while ((1 << bits) <= nmax) ++bits;  // 32 - clz(nmax)
  return nm_walk_ks4<1>(sa.data(), sb.data(), n0, n1, it4, bits);
}

int emul_lane_position(const float* a, int n0, const float* b, int n1, int want_u, int want_t,
                       nm_row_out* out) {
  const int nmax = n0 > n1 ? n0 : n1;
  if (nmax > NM_LANE_MAX_N || n0 < 2 || n1 < 2) return 1;
  const int nsel = nm_lane_class(nmax);
  memset(out, 0, sizeof(*out));
#define CALL(NN) lane_position<NN>(a, n0, b, n1, want_u != 0, want_t != 0, out)
  NM_DISPATCH_N(nsel, CALL)
#undef CALL
  return 0;
}

// Deep tier: sorted copies + per-element rank counts, reduced exactly as the block kernel does.
int emul_deep_position(const float* a, int n0, const float* b, int n1, int want_u, int want_t,
                       nm_row_out* out) {
  std::vector<float> sa(a, a + n0), sb(b, b + n1);
  std::sort(sa.begin(), sa.end());
  std::sort(sb.begin(), sb.end());
  nm_deep_acc acc;
  nm_deep_acc_init(&acc);
  for (int e = 0; e < n0 + n1; ++e) {
    nm_deep_acc one;
    nm_deep_acc_init(&one);
    nm_deep_element(sa.data(), n0, sb.data(), n1, e, want_u != 0, &one);
    nm_deep_acc_merge(&acc, one);
  }
  double m0 = 0, v0 = 0, m1 = 0, v1 = 0;
  if (want_t) {
    nm_moments(a, n0, &m0, &v0);
    nm_moments(b, n1, &m1, &v1);
  }
  memset(out, 0, sizeof(*out));
  nm_deep_finish(acc, n0, n1, want_u != 0, want_t != 0, m0, v0, m1, v1, out);
  return 0;
}

void emul_combine(const double* ks_p, const int32_t* pos, const int32_t* seg, int64_t n, int nb,
                  double weights_dif, int want_fisher, int want_stouffer, double* f_stat,
                  double* f_p, double* s_stat, double* s_p) {
  std::vector<double> w(nb + 1);
  const double wnorm = nm_build_weights(nb, weights_dif, w.data());
  for (int64_t i = 0; i < n; ++i) {
    auto W = [&](int k, double* z, double* lnp) {
      const int64_t j = i + k;
      bool ok = j >= 0 && j < n;
      if (ok && k != 0) ok = seg[j] == seg[i] && (int64_t)pos[j] - (int64_t)pos[i] == (int64_t)k;
      const double p = ok ? ks_p[j] : 1.0;
      *z = nm_norm_isf(p);
      *lnp = log(p);
    };
    double fs = 0, fp = 0, ss = 0, sp = 0;
    nm_combine_row(nb, w.data(), wnorm, W, want_fisher != 0, want_stouffer != 0, &fs, &fp, &ss, &sp);
    if (want_fisher) { f_stat[i] = fs; f_p[i] = fp; }
    if (want_stouffer) { s_stat[i] = ss; s_p[i] = sp; }
  }
}

// The grid-key statement of nm_lane.cuh against EVERY float32 pattern: a value passes
// nm_grid_bits + the range test iff it is fl32(fl64(k / 1000)) for an integer |k| <= 32766 (what
// numpy's cast of the reference's round(x, 3) gives), and then both key patterns carry k + 32768.
// Returns the number of violations; *n_pass = patterns that pass (65 533 grid points and -0.0).
long long emul_grid_exhaustive(long long* n_pass, int n_threads) {
  std::atomic<long long> viol(0), pass(0);
  std::vector<std::thread> th;
  for (int t = 0; t < n_threads; ++t)
    th.emplace_back([&, t]() {
      long long v = 0, p = 0;
      const unsigned long long lo = (1ull << 32) * t / n_threads, hi = (1ull << 32) * (t + 1) / n_threads;
      for (unsigned long long b = lo; b < hi; ++b) {
        const unsigned bits = (unsigned)b;
        float x;
        memcpy(&x, &bits, 4);
        nm_grid_flag bad = NM_GRID_FLAG0, bad2 = NM_GRID_FLAG0;
        const unsigned ta = nm_grid_bits(x, NM_GRID_MA, &bad);
        const unsigned tb = nm_grid_bits(x, NM_GRID_MB, &bad2);
        const bool ok = !nm_grid_failed(bad) && fabsf(x) <= NM_GRID_LIM, ok2 = !nm_grid_failed(bad2) && fabsf(x) <= NM_GRID_LIM;
        if (ok != ok2) ++v;
        if (!ok) continue;
        ++p;
        const int k = (int)(ta & 0xffffu) - 32768;
        const float canon = (float)((double)k / 1000.0);
        if (!(canon == x) || k < -32766 || k > 32766) ++v;
        if ((ta >> 16) != 0x4B40u) ++v;
        const unsigned packed = tb * 65536u + ta;
        if ((packed >> 16) != (unsigned)(k + 32768) || (packed & 0xffffu) != (unsigned)(k + 32768)) ++v;
      }
      viol += v;
      pass += p;
    });
  for (auto& x : th) x.join();
  *n_pass = pass.load();
  return viol.load();
}

// Lane tier through the packed grid-key path (both groups sorted by one network pass over 16-bit
// key pairs, walks on the 16-bit columns).  Returns 2 when a value fails the grid test (the kernel
// then takes the float path), else fills *out like emul_lane_position.
}  // extern "C"
namespace {
template <int N>
int lane_position_grid(const float* a, int n0, const float* b, int n1, bool want_u, bool want_t, int walk,
                       nm_row_out* out) {
  nm_lane_acc acc;
  memset(&acc, 0, sizeof(acc));
  nm_p16 w[N];
  nm_grid_flag bad = NM_GRID_FLAG0;
  float m = 0.0f;
  for (int k = 0; k < N; ++k) {
    const unsigned ta = k < n0 ? nm_grid_bits(a[k], NM_GRID_MA, &bad) : NM_GRID_PAD_A;
    const unsigned tb = k < n1 ? nm_grid_bits(b[k], NM_GRID_MB, &bad) : NM_GRID_PAD_B;
    if (k < n0) m = fmaxf(m, fabsf(a[k]));
    if (k < n1) m = fmaxf(m, fabsf(b[k]));
    w[k].v = tb * 65536u + ta;
  }
  if (nm_grid_failed(bad) || !(m <= NM_GRID_LIM)) return 2;
  if (want_t) {
    nm_moments(a, n0, &acc.mean0, &acc.var0);
    nm_moments(b, n1, &acc.mean1, &acc.var1);
  }
  nm_sorter<N>::run(w, g_one, g_mone);
  std::vector<unsigned> col(N + 2 + 4, NM_GRID_PINF);
  col[0] = NM_GRID_NINF;
  for (int k = 0; k < N; ++k) col[k + 1] = w[nm_sorter<N>::order(k)].v;
  const unsigned short* ca = reinterpret_cast<const unsigned short*>(col.data());  // little endian: low half first
  const unsigned short* cb = ca + 1;
  const int T = n0 + n1, nmax = n0 > n1 ? n0 : n1;
  if (want_u || walk == 0) {
    const int iters = (T + 1) / 2 + 3;
    if (want_u)
      nm_merge_walk<true, 2>(ca, cb, n0, n1, iters, &acc);
    else
      nm_merge_walk<false, 2>(ca, cb, n0, n1, iters, &acc);
  } else if (walk == 2) {
    acc.dnum = nm_walk_ks2<2>(ca, cb, n0, n1, (T + 1) >> 1);
  } else {
    int bits = 0;
    while ((1 << bits) <= nmax) ++bits;
    if ((T >> 1) < ((T + 3) >> 2)) return 3;
    acc.dnum = nm_walk_ks4<2>(ca, cb, n0, n1, (T + 3) >> 2, bits);
  }
  nm_lane_finish(acc, n0, n1, want_u, want_t, out);
  return 0;
}
}  // namespace
extern "C" {
int emul_lane_position_grid(const float* a, int n0, const float* b, int n1, int want_u, int want_t, int walk,
                            nm_row_out* out) {
  const int nmax = n0 > n1 ? n0 : n1;
  if (nmax > NM_LANE_MAX_N || n0 < 2 || n1 < 2) return 1;
  const int nsel = nm_lane_class(nmax);
  memset(out, 0, sizeof(*out));
  int rc = 0;
#define CALL(NN) rc = lane_position_grid<NN>(a, n0, b, n1, want_u != 0, want_t != 0, walk, out)
  NM_DISPATCH_N(nsel, CALL)
#undef CALL
  return rc;
}

double emul_kolmogorov_sf(double x) { return nm_kolmogorov_sf(x); }
double emul_student_t_two_sided(double t, double df) { return nm_student_t_two_sided(t, df); }
double emul_chi2_sf_even(double x2, int k) { return nm_chi2_sf_even(x2, k); }
double emul_norm_sf(double z) { return nm_norm_sf(z); }
double emul_norm_isf(double p) { return nm_norm_isf(p); }
int emul_sizeof_row_out(void) { return (int)sizeof(nm_row_out); }
}
