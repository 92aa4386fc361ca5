// nm_api.cu -- plan / deep / combine kernels and the C ABI of include/nanomod_b200.h.
//
// Stage map (reference: bin/scripts/myDetect.py):
//   nm_plan_*          mfilter_coverage (:301-314) + the "present in both groups" rule of
//                      mtest2 (:428,431): ordered compaction of candidate positions into rows.
//   nm_lane_kernel     (nm_lane_kernel.cu) getKStest (:327-343), rows with <= 128 reads/group.
//   nm_deep_kernel     the same statistics for deeper rows: one CTA per position, TMA bulk
//                      load of the contiguous pileup slice, shared-memory bitonic sort, rank
//                      counts by binary search.
//   nm_combine_kernel  combin_pvalues / get_combin_pvalue / pos_check (:366-414).
// There is no CPU path: every entry point fails with NM_ERR_NO_DEVICE / NM_ERR_CUDA when the
// GPU is not usable.
#include <stdarg.h>

#include <algorithm>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "nm_device.cuh"
#include "nm_downsample.cuh"
#include "nm_huge.cuh"
#include "nm_rank.cuh"

// ------------------------------------------------------------------------------------------
// plan: coverage filter + ordered compaction (myDetect.py:301-314, :428-431)
// ------------------------------------------------------------------------------------------
#define NM_PLAN_THREADS 256
#define NM_PLAN_PER_THREAD 4
#define NM_PLAN_PER_BLOCK (NM_PLAN_THREADS * NM_PLAN_PER_THREAD)

// exclusive scan of one int per thread over a 256-thread block; returns block total in *total
__device__ __forceinline__ int nm_block_excl_scan(int v, int* total) {
  __shared__ int warp_tot[NM_PLAN_THREADS / 32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  int inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) warp_tot[wid] = inc;
  __syncthreads();
  int base = 0, tot = 0;
#pragma unroll
  for (int w = 0; w < NM_PLAN_THREADS / 32; ++w) {
    const int t = warp_tot[w];
    if (w < wid) base += t;
    tot += t;
  }
  __syncthreads();
  *total = tot;
  return base + inc - v;
}

__global__ void __launch_bounds__(NM_PLAN_THREADS)
nm_plan_count(const int64_t* __restrict__ off0, const int64_t* __restrict__ off1, int64_t n_pos,
              int mincov, const int32_t* __restrict__ seg, int n_seg, const int32_t* __restrict__ seg_cov,
              int* __restrict__ block_count, nm_summary* __restrict__ sum) {
  const int64_t p0 = (int64_t)blockIdx.x * NM_PLAN_PER_BLOCK + (int64_t)threadIdx.x * NM_PLAN_PER_THREAD;
  int cnt = 0, max_lane = 0, max_slack = 0, n_deep = 0, max_deep = 0, n_cand = 0, max_deep_t = 0, n_le64 = 0, n_le104 = 0;
  bool bad = false, ds_deep = false, ds_huge = false;
#pragma unroll
  for (int k = 0; k < NM_PLAN_PER_THREAD; ++k) {
    const int64_t p = p0 + k;
    if (p < n_pos) {
      const int64_t n0 = off0[p + 1] - off0[p];
      const int64_t n1 = off1[p + 1] - off1[p];
      ++n_cand;
      bad |= (n0 < 0) | (n1 < 0);
      if (n_seg > 0) bad |= (unsigned)seg[p] >= (unsigned)n_seg;  // per-segment arrays are indexed with it
      if (n0 >= mincov && n1 >= mincov) {
        ++cnt;
        if (seg_cov && n_seg > 0 && (unsigned)seg[p] < (unsigned)n_seg) {
          // a row the down-sampling branch will take, but too long for it: known before any test runs
          const int cov = seg_cov[seg[p]];
          if (cov > 0 && (n0 > cov || n1 > cov) && (n0 > NM_DS_MAX_READS || n1 > NM_DS_MAX_READS)) {
            ds_deep = true;
            if (cov > NM_DS_DEEP_MAX_COV || (n0 > cov && n0 > NM_DS_DEEP_MAX_READS) || (n1 > cov && n1 > NM_DS_DEEP_MAX_READS))
              ds_huge = true;
          }
        }
        const int64_t m = n0 > n1 ? n0 : n1;
        if (m <= NM_LANE_TIER_MAX) {
          const int cls = nm_lane_class((int)m);
          n_le64 += cls <= 64 ? 1 : 0;
          n_le104 += cls <= NM_LANE_FINE_MAX ? 1 : 0;
          max_lane = max_lane > (int)m ? max_lane : (int)m;
          const int slack = NM_LANE_TIER_MAX - (int)m;
          max_slack = max_slack > slack ? max_slack : slack;
        } else {
          ++n_deep;
          const int64_t tt = n0 + n1;
          max_deep_t = max_deep_t > (int)(tt < 0x7fffffff ? tt : 0x7fffffff) ? max_deep_t : (int)(tt < 0x7fffffff ? tt : 0x7fffffff);
          const int64_t cap = 1 << 24;
          const int p2 = nm_deep_p2((int)(n0 < cap ? n0 : cap)) + nm_deep_p2((int)(n1 < cap ? n1 : cap));
          if (p2 > NM_DEEP_TIER_MAX_POOLED) {  // too long for shared memory: the global-memory path (nm_huge.cu)
            atomicAdd(&sum->n_huge, 1);
            atomicAdd(&sum->huge_v0, (unsigned long long)n0);
            atomicAdd(&sum->huge_v1, (unsigned long long)n1);
          } else {
            max_deep = max_deep > p2 ? max_deep : p2;
          }
        }
      }
    }
  }
  // block totals: warp reductions, one shared-memory hop, ONE set of atomics per block -- and only
  // those that would change the summary (every block hitting the same addresses costs tens of
  // microseconds at 4.6 M candidates)
  unsigned flags = (bad ? 1u : 0u) | (ds_deep ? 2u : 0u) | (ds_huge ? 4u : 0u);
  flags = __reduce_or_sync(0xffffffffu, flags);
  cnt = __reduce_add_sync(0xffffffffu, cnt);
  n_cand = __reduce_add_sync(0xffffffffu, n_cand);
  n_le64 = __reduce_add_sync(0xffffffffu, n_le64);
  n_le104 = __reduce_add_sync(0xffffffffu, n_le104);
  n_deep = __reduce_add_sync(0xffffffffu, n_deep);
  max_lane = __reduce_max_sync(0xffffffffu, max_lane);
  max_slack = __reduce_max_sync(0xffffffffu, max_slack);
  max_deep = __reduce_max_sync(0xffffffffu, max_deep);
  max_deep_t = __reduce_max_sync(0xffffffffu, max_deep_t);
  __shared__ int red[10][NM_PLAN_THREADS / 32];
  if ((threadIdx.x & 31) == 0) {
    const int w = threadIdx.x >> 5;
    red[0][w] = max_lane; red[1][w] = max_slack; red[2][w] = max_deep; red[3][w] = n_deep; red[4][w] = max_deep_t;
    red[5][w] = cnt; red[6][w] = n_cand; red[7][w] = n_le64; red[8][w] = n_le104; red[9][w] = (int)flags;
  }
  __syncthreads();
  int total = 0;
  if (threadIdx.x == 0) {
    int ml = 0, ms = 0, md = 0, nd = 0, mt = 0, nc = 0, l64 = 0, l104 = 0, fl = 0;
    for (int w = 0; w < NM_PLAN_THREADS / 32; ++w) {
      ml = ml > red[0][w] ? ml : red[0][w];
      ms = ms > red[1][w] ? ms : red[1][w];
      md = md > red[2][w] ? md : red[2][w];
      nd += red[3][w];
      mt = mt > red[4][w] ? mt : red[4][w];
      total += red[5][w];
      nc += red[6][w];
      l64 += red[7][w];
      l104 += red[8][w];
      fl |= red[9][w];
    }
    if (fl & 1) sum->bad_input = 1;
    if (fl & 6) atomicOr(&sum->ds_too_deep, (int)((fl >> 1) & 3));
    if (nc != total) atomicAdd(&sum->n_filtered, nc - total);
    int* sp = sum->spread[blockIdx.x & (NM_SPREAD - 1)];
    if (total) atomicAdd(&sp[0], total);
    if (l64) atomicAdd(&sp[1], l64);
    if (l104) atomicAdd(&sp[2], l104);
    if (ml > *(volatile int*)&sum->max_lane_n) atomicMax(&sum->max_lane_n, ml);
    if (ms > *(volatile int*)&sum->max_lane_slack) atomicMax(&sum->max_lane_slack, ms);
    if (nd) {
      atomicAdd(&sum->n_deep, nd);
      atomicMax(&sum->max_deep_p2, md);
      atomicMax(&sum->max_deep_t, mt);
    }
  }
  if (threadIdx.x == 0) block_count[blockIdx.x] = total;
}

// single-block exclusive scan of the per-block counts (in place) + total row count
__global__ void __launch_bounds__(NM_PLAN_THREADS)
nm_plan_scan(int* __restrict__ block_count, int nblk, nm_summary* __restrict__ sum) {
  __shared__ unsigned long long carry_s;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  for (int base = 0; base < nblk; base += NM_PLAN_THREADS) {
    const int idx = base + threadIdx.x;
    const int v = idx < nblk ? block_count[idx] : 0;
    int total;
    const int ex = nm_block_excl_scan(v, &total);
    const unsigned long long carry = carry_s;
    if (idx < nblk) block_count[idx] = (int)(carry + (unsigned long long)ex);
    __syncthreads();
    if (threadIdx.x == 0) carry_s = carry + (unsigned long long)total;
    __syncthreads();
  }
  if (threadIdx.x == 0) sum->n_rows = carry_s;
}

__global__ void __launch_bounds__(NM_PLAN_THREADS)
nm_plan_scatter(const int64_t* __restrict__ off0, const int64_t* __restrict__ off1, int64_t n_pos,
                int mincov, const int* __restrict__ block_offset, int32_t* __restrict__ row_pos_index,
                int32_t* __restrict__ row_n0, int32_t* __restrict__ row_n1,
                int32_t* __restrict__ deep_rows, nm_summary* __restrict__ sum) {
  const int64_t p0 = (int64_t)blockIdx.x * NM_PLAN_PER_BLOCK + (int64_t)threadIdx.x * NM_PLAN_PER_THREAD;
  int n0v[NM_PLAN_PER_THREAD], n1v[NM_PLAN_PER_THREAD];
  bool keep[NM_PLAN_PER_THREAD];
  int cnt = 0;
#pragma unroll
  for (int k = 0; k < NM_PLAN_PER_THREAD; ++k) {
    const int64_t p = p0 + k;
    keep[k] = false;
    n0v[k] = n1v[k] = 0;
    if (p < n_pos) {
      const int64_t n0 = off0[p + 1] - off0[p];
      const int64_t n1 = off1[p + 1] - off1[p];
      keep[k] = (n0 >= mincov && n1 >= mincov);
      n0v[k] = (int)(n0 < 0x7fffffff ? n0 : 0x7fffffff);
      n1v[k] = (int)(n1 < 0x7fffffff ? n1 : 0x7fffffff);
      cnt += keep[k] ? 1 : 0;
    }
  }
  int total;
  int r = block_offset[blockIdx.x] + nm_block_excl_scan(cnt, &total);
#pragma unroll
  for (int k = 0; k < NM_PLAN_PER_THREAD; ++k) {
    if (keep[k]) {
      row_pos_index[r] = (int32_t)(p0 + k);
      row_n0[r] = n0v[k];
      row_n1[r] = n1v[k];
      if (n0v[k] > NM_LANE_TIER_MAX || n1v[k] > NM_LANE_TIER_MAX) {
        const int slot = atomicAdd(&sum->deep_cursor, 1);
        deep_rows[slot] = r;
      }
      ++r;
    }
  }
}

// ------------------------------------------------------------------------------------------
// tails: fp64 tails of the rank-sum and Welch tests for lane-tier rows (mannwhitneyu and
// ttest_ind tails, bin/scripts/myDetect.py:331-337), from the integers / moments the sort kernels
// left in scratch.  Deep rows are finished by nm_deep_kernel itself.
// ------------------------------------------------------------------------------------------
struct nm_tails_args {
  const int32_t* row_n0;
  const int32_t* row_n1;
  const int* acc_r2;
  const int* acc_tie;
  const double* acc_mom;
  int64_t n_rows;
  int want_u, want_t;
  int64_t* two_u;
  double* u_stat;
  double* u_p;
  double* t_stat;
  double* t_p;
  uint8_t* flags;
};

__global__ void __launch_bounds__(256) nm_tails_kernel(const nm_tails_args a) {
  const int64_t r = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (r >= a.n_rows) return;
  const int n0 = a.row_n0[r], n1 = a.row_n1[r];
  if (n0 > NM_LANE_TIER_MAX || n1 > NM_LANE_TIER_MAX) return;
  int flag = 0;
  if (a.want_u) {
    double us, up;
    int64_t two_u;
    nm_mwu_tail(a.acc_r2[r], a.acc_tie[r], n0, n1, &us, &two_u, &up, &flag);
    a.two_u[r] = two_u;
    if (a.u_stat) a.u_stat[r] = us;
    a.u_p[r] = up;
  }
  if (a.want_t) {
    const double4 m = reinterpret_cast<const double4*>(a.acc_mom)[r];
    double ts, tp;
    nm_welch_tail(m.x, m.y, n0, m.z, m.w, n1, &ts, &tp);
    a.t_stat[r] = ts;
    a.t_p[r] = tp;
  }
  if (a.flags) a.flags[r] = (uint8_t)flag;
}

// ------------------------------------------------------------------------------------------
// combine: sliding-window Fisher / weighted Stouffer over the KS p-values (myDetect.py:366-414)
// ------------------------------------------------------------------------------------------
#define NM_COMB_THREADS 256

struct nm_comb_args {
  const double* ks_p;
  const int32_t* ks_dnum;
  const int32_t* row_pos_index;
  const int32_t* row_n0;
  const int32_t* row_n1;
  const int32_t* pos;
  const int32_t* seg;
  const double* z_pre;     // norm.isf(ks_p) left by the dense lane kernel, or NULL
  const double* ln_pre;    // ln(ks_p) left by the dense lane kernel, or NULL
  const double* ks_d;      // the table's D column when present (nb == 0 returns the KS tuple)
  const int32_t* seg_cov;  // down-sampling thresholds per segment (nb == 0, D of resampled rows), or NULL
  int64_t n_rows;
  int nb;
  int want_fisher, want_stouffer;
  double wnorm;
  double w[NM_MAX_NB + 1];
  double* f_stat;
  double* f_p;
  double* s_stat;
  double* s_p;
  // candidate list for an armed head selection (nm_rank.cuh): rows [head_lo, head_hi) whose combined p
  // (head_col 0: Stouffer, 1: Fisher) has a key image in an exponent bin <= head_thr_bin are appended
  int2* head_cand;      // {row - head_lo, exponent bin}
  int* head_cursor;
  int64_t head_lo, head_hi;
  unsigned head_thr_bin;
  int head_cap, head_col;
};

struct nm_comb_win {
  const double* z;
  const double* lnp;
  const int* pos;
  const int* seg;
  int c;  // index of the centre row inside the shared tile
  __device__ __forceinline__ void operator()(int k, double* zo, double* lo) const {
    const int j = c + k;
    // a halo slot outside [0, n_rows) carries seg = -1 and never matches
    const bool okk = (k == 0) || (seg[j] == seg[c] && pos[j] - pos[c] == k);
    *zo = okk ? z[j] : -INFINITY;
    *lo = okk ? lnp[j] : 0.0;
  }
};

// NB > 0: neighborPvalues known at compile time (window loops unrolled, weights indexed statically);
// NB == 0: any value, read from the arguments.  Same operations in the same order either way.
template <int NB>
__global__ void __launch_bounds__(NM_COMB_THREADS) nm_combine_kernel(const nm_comb_args a) {
  __shared__ double z_s[NM_COMB_THREADS + 2 * NM_MAX_NB];
  __shared__ double l_s[NM_COMB_THREADS + 2 * NM_MAX_NB];
  __shared__ int pos_s[NM_COMB_THREADS + 2 * NM_MAX_NB];
  __shared__ int seg_s[NM_COMB_THREADS + 2 * NM_MAX_NB];
  const int nb = NB > 0 ? NB : a.nb;
  const int64_t tile0 = (int64_t)blockIdx.x * NM_COMB_THREADS;
  for (int t = threadIdx.x; t < NM_COMB_THREADS + 2 * nb; t += NM_COMB_THREADS) {
    const int64_t r = tile0 - nb + t;
    double z = -INFINITY, l = 0.0;
    int ps = 0, sg = -1;
    if (r >= 0 && r < a.n_rows) {
      const int64_t src = a.row_pos_index ? (int64_t)a.row_pos_index[r] : r;  // dense path: row == candidate
      ps = a.pos[src];
      sg = a.seg[src];
      if (a.want_stouffer) z = a.z_pre ? a.z_pre[r] : nm_norm_isf(a.ks_p[r]);
      if (a.want_fisher) l = a.ln_pre ? a.ln_pre[r] : log(a.ks_p[r]);
    }
    z_s[t] = z;
    l_s[t] = l;
    pos_s[t] = ps;
    seg_s[t] = sg;
  }
  __syncthreads();
  const int64_t r = tile0 + threadIdx.x;
  if (r >= a.n_rows) return;
  if (nb == 0) {
    // get_combin_pvalue returns the KS tuple itself when neighborPvalues == 0 (:413)
    // D as the table has it; without that column it is rebuilt from the numerator, which for a
    // down-sampled row is relative to the resampled sizes min(n, cov) (nm_downsample.cu)
    double d;
    if (a.ks_d) {
      d = a.ks_d[r];
    } else {
      int m0 = a.row_n0[r], m1 = a.row_n1[r];
      if (a.seg_cov) {
        const int64_t src = a.row_pos_index ? (int64_t)a.row_pos_index[r] : r;
        const int cov = a.seg_cov[a.seg[src]];
        if (cov > 0 && (m0 > cov || m1 > cov)) {
          m0 = m0 < cov ? m0 : cov;
          m1 = m1 < cov ? m1 : cov;
        }
      }
      d = nm_max_float((double)a.ks_dnum[r] / ((double)m0 * (double)m1));
    }
    const double p = a.ks_p[r];
    if (a.want_fisher) { a.f_stat[r] = d; a.f_p[r] = p; }
    if (a.want_stouffer) { a.s_stat[r] = d; a.s_p[r] = p; }
    return;
  }
  const nm_comb_win W{z_s, l_s, pos_s, seg_s, (int)threadIdx.x + nb};
  double fs = 0.0, fp = 0.0, ss = 0.0, sp = 0.0;
  nm_combine_row(nb, a.w, a.wnorm, W, a.want_fisher != 0, a.want_stouffer != 0, &fs, &fp, &ss, &sp);
  if (a.want_fisher) { a.f_stat[r] = fs; a.f_p[r] = fp; }
  if (a.want_stouffer) { a.s_stat[r] = ss; a.s_p[r] = sp; }
  if (a.head_cand && r >= a.head_lo && r < a.head_hi) {
    const unsigned bin = (unsigned)(nm_key_image(a.head_col ? fp : sp) >> 52);
    if (bin <= a.head_thr_bin) {
      const int slot = atomicAdd(a.head_cursor, 1);
      if (slot < a.head_cap) a.head_cand[slot] = make_int2((int)(r - a.head_lo), (int)bin);
    }
  }
}

// ------------------------------------------------------------------------------------------
// host side: handle, scratch, C ABI
// ------------------------------------------------------------------------------------------
struct nm_buf {
  void* p;
  size_t cap;
};

struct nm_handle {
  int device;
  int sm_count;
  cudaStream_t own_stream;
  nm_summary* d_sum;
  nm_summary* h_sum;  // pinned
  void* h_head;       // pinned landing area of nm_rank_head_device
  nm_buf d_block_count, d_deep_rows, d_acc_r2, d_acc_tie, d_acc_mom;
  // staging for nm_detect_host
  nm_buf d_vals0, d_vals1, d_off0, d_off1, d_pos, d_seg, d_seg_cov;
  nm_buf d_out[17];
  nm_buf d_rank, d_rank_keys[3], d_rank_order;
  nm_buf d_perm[2], d_class_scratch;  // class binning of the lane tier (nm_class_sort_run)  // nm_rank_*: scratch, staged key columns, result
  int64_t launches;
  int sm_limit;        // SMs the persistent lane kernel may occupy (0 = all)
  int no_class_sort;   // NANOMOD_B200_NO_CLASS_SORT=1: never bin rows by network class (A/B experiments)
  // nm_arm_head_select: a head selection to be launched by the next nm_detect_device call behind its own
  // kernels, BEFORE the call's host wait (a sharded step then has no idle gap between the tests and the exchange)
  struct {
    int armed, fired;
    const double* key[3];
    int64_t n_rows, want, cap;
    int reverse, have_geo;
    nm_head_geo geo;
    nm_head_record* records;
    nm_head_peers_dev peers;
    int use_cands;       // this call's combine kernel lists the candidates (nm_head_from_cands_run selects)
    unsigned thr_bin;    // ... rows whose key image lies in an exponent bin <= this
  } head;
  nm_buf d_head_cand;
  nm_buf d_ds_scratch;   // sorted groups of nm_downsample_deep_kernel, per CTA
  int head_cut_hint;     // bin at which the last candidate-list selection reached `want` (0: none / it failed)
  int64_t head_hint_n, head_hint_want;  // ... for this many rows and this `want`
  nm_head_peers_dev next_peers;  // nm_head_set_peers: taken by the next arming / selection (one shot)
  int grid_skip;       // calls left for which the grid-key launch is skipped (the last attempt found off-grid data)
  nm_buf d_retry;      // retry list of the grid-key launch
  int no_head_cands;   // NANOMOD_B200_NO_HEAD_CANDS=1: armed head selections always take the three-pass form (A/B experiments)
  int grid_u;          // NANOMOD_B200_GRID_U=1: take the grid-key kernel also when U is wanted (tests of that walk)
  int no_grid;         // NANOMOD_B200_NO_GRID=1: never try the 16-bit grid-key sort of the lane tier (A/B experiments, tests)
  int no_dense;        // NANOMOD_B200_NO_DENSE=1: never take the dense path (A/B experiments, tests of the general path)
  int dense_class;     // network class of the previous call when it had the dense shape (else 0): the next
                       // call is launched on that assumption without waiting for its plan summary
  int last_grid_tiles; // lane-tier tiles of the last call that went through the grid-key sort
  int last_path;       // 0 general, 1 dense, 2 dense launched speculatively, 3 / 4 speculative launch refused and
                       // the call re-run dense with the right network class / on the general path
  nm_buf d_comb_z, d_comb_ln, d_deep_fallback, d_exp0, d_exp1;
  // pipelined nm_detect_host: two slots of staging (inputs + outputs), copy streams and events
  nm_buf p_in[2][6];    // vals0, vals1, off0, off1, pos, seg of a slab
  nm_buf p_i16[2][2];   // the slab's int16 values (16-bit transport format), expanded into p_in[.][0..1]
  nm_buf p_out[2][17];  // the slab's table
  cudaStream_t s_in, s_out;
  cudaEvent_t ev_in[2], ev_cmp[2], ev_out[2];
  int64_t slab;         // candidates per slab (NANOMOD_B200_SLAB; 0 = never pipeline)
  cudaEvent_t* ev;     // plan start | tests start | deep start | combine start | end: the set of the call being issued
  cudaEvent_t ev_sets[3][5];  // set 0: synchronous calls; sets 1, 2: the two slots of nm_detect_device_async
  nm_summary* h_sums[3];      // pinned; h_sum points at the current call's (same numbering)
  cudaEvent_t ev_done[2];     // end of an asynchronous call's work on its stream
  // nm_detect_device_async: up to two calls in flight.  A call launched on the previous call's shape is
  // validated (and re-run if the device refused it) by nm_detect_finish.
  struct nm_pending {
    int active, done;      // issued and not yet finished | completed synchronously when it was issued
    int64_t n_rows;
    nm_pileup pl;
    nm_params prm;
    nm_table tb;
    void* stream;
    int try_grid, n_launched, head_fired;
  } pend[2];
  int async_next;
  double last_ms[4];   // plan, lane tier, deep tier, combine of the most recent call
  char err[512];
};

static char g_err[512] = "";

static int nm_fail(nm_handle* h, int code, const char* fmt, ...) {
  char* dst = h ? h->err : g_err;
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(dst, 512, fmt, ap);
  va_end(ap);
  return code;
}

#define NM_CUDA(h, call)                                                                  \
  do {                                                                                    \
    cudaError_t e_ = (call);                                                              \
    if (e_ != cudaSuccess)                                                                \
      return nm_fail(h, e_ == cudaErrorMemoryAllocation ? NM_ERR_OOM : NM_ERR_CUDA,       \
                     "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__,    \
                     __LINE__);                                                           \
  } while (0)

static int nm_reserve(nm_handle* h, nm_buf* b, size_t bytes) {
  if (bytes <= b->cap) return NM_OK;
  if (b->p) {
    NM_CUDA(h, cudaFree(b->p));
    b->p = nullptr;
    b->cap = 0;
  }
  const size_t want = bytes + bytes / 8 + 256;
  NM_CUDA(h, cudaMalloc(&b->p, want));
  b->cap = want;
  return NM_OK;
}

extern "C" int nm_version(void) { return NM_VERSION; }

extern "C" int64_t nm_padded_len(int64_t nvals) { return (nvals + 3) / 4 * 4 + 4; }

extern "C" const char* nm_last_error(const nm_handle* h) { return h ? h->err : g_err; }

extern "C" int64_t nm_launch_count(const nm_handle* h) { return h ? h->launches : 0; }

extern "C" int nm_set_sm_limit(nm_handle* h, int n_sms) {
  if (!h) return NM_ERR_BAD_ARG;
  if (n_sms < 0 || n_sms > h->sm_count) return nm_fail(h, NM_ERR_BAD_ARG, "sm limit %d out of [0,%d]", n_sms, h->sm_count);
  h->sm_limit = n_sms;
  return NM_OK;
}

extern "C" int nm_sm_count(const nm_handle* h) { return h ? h->sm_count : 0; }

extern "C" int nm_last_timings(const nm_handle* h, double* ms4) {
  if (!h || !ms4) return NM_ERR_BAD_ARG;
  for (int k = 0; k < 4; ++k) ms4[k] = h->last_ms[k];
  return NM_OK;
}

extern "C" int nm_last_path(const nm_handle* h) { return h ? h->last_path : -1; }
extern "C" int64_t nm_last_grid_tiles(const nm_handle* h) { return h ? h->last_grid_tiles : -1; }

extern "C" int nm_create(int device, nm_handle** out) {
  if (!out) return nm_fail(nullptr, NM_ERR_BAD_ARG, "nm_create: out is NULL");
  *out = nullptr;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return nm_fail(nullptr, NM_ERR_NO_DEVICE, "nm_create: no CUDA device (%s)",
                   e == cudaSuccess ? "count is 0" : cudaGetErrorString(e));
  if (device < 0 || device >= ndev)
    return nm_fail(nullptr, NM_ERR_BAD_ARG, "nm_create: device %d out of range [0,%d)", device, ndev);
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess)
    return nm_fail(nullptr, NM_ERR_CUDA, "nm_create: cudaGetDeviceProperties failed");
  if (prop.major != 10)
    return nm_fail(nullptr, NM_ERR_NO_DEVICE,
                   "nm_create: device %d is sm_%d%d; this library is built for sm_100a only",
                   device, prop.major, prop.minor);
  nm_handle* h = (nm_handle*)calloc(1, sizeof(nm_handle));
  if (!h) return nm_fail(nullptr, NM_ERR_OOM, "nm_create: out of host memory");
  h->device = device;
  h->sm_count = prop.multiProcessorCount;
  {
    const char* g = getenv("NANOMOD_B200_NO_CLASS_SORT");
    h->no_class_sort = (g && g[0] == '1') ? 1 : 0;
    const char* sl = getenv("NANOMOD_B200_SLAB");
    h->slab = sl ? atoll(sl) : 262144;
    const char* d = getenv("NANOMOD_B200_NO_DENSE");
    h->no_dense = (d && d[0] == '1') ? 1 : 0;
    const char* gk = getenv("NANOMOD_B200_NO_GRID");
    h->no_grid = (gk && gk[0] == '1') ? 1 : 0;
    const char* hc = getenv("NANOMOD_B200_NO_HEAD_CANDS");
    h->no_head_cands = (hc && hc[0] == '1') ? 1 : 0;
    const char* gu = getenv("NANOMOD_B200_GRID_U");
    h->grid_u = (gu && gu[0] == '1') ? 1 : 0;
  }
  int rc = NM_OK;
  do {
    if (cudaSetDevice(device) != cudaSuccess) { rc = NM_ERR_CUDA; break; }
    if (cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking) != cudaSuccess) { rc = NM_ERR_CUDA; break; }
    if (cudaMalloc(&h->d_sum, sizeof(nm_summary)) != cudaSuccess) { rc = NM_ERR_OOM; break; }
    for (int s = 0; s < 3 && rc == NM_OK; ++s) {
      if (cudaMallocHost(&h->h_sums[s], sizeof(nm_summary)) != cudaSuccess) { rc = NM_ERR_OOM; break; }
      for (int k = 0; k < 5 && rc == NM_OK; ++k)
        if (cudaEventCreate(&h->ev_sets[s][k]) != cudaSuccess) rc = NM_ERR_CUDA;
    }
    if (rc != NM_OK) break;
    h->h_sum = h->h_sums[0];
    h->ev = h->ev_sets[0];
    for (int s = 0; s < 2 && rc == NM_OK; ++s)
      if (cudaEventCreateWithFlags(&h->ev_done[s], cudaEventDisableTiming) != cudaSuccess) rc = NM_ERR_CUDA;
    if (cudaStreamCreateWithFlags(&h->s_in, cudaStreamNonBlocking) != cudaSuccess) rc = NM_ERR_CUDA;
    if (cudaStreamCreateWithFlags(&h->s_out, cudaStreamNonBlocking) != cudaSuccess) rc = NM_ERR_CUDA;
    for (int k = 0; k < 2 && rc == NM_OK; ++k)
      if (cudaEventCreateWithFlags(&h->ev_in[k], cudaEventDisableTiming) != cudaSuccess ||
          cudaEventCreateWithFlags(&h->ev_cmp[k], cudaEventDisableTiming) != cudaSuccess ||
          cudaEventCreateWithFlags(&h->ev_out[k], cudaEventDisableTiming) != cudaSuccess) rc = NM_ERR_CUDA;
  } while (0);
  if (rc != NM_OK) {
    nm_fail(nullptr, rc, "nm_create: CUDA set-up failed: %s", cudaGetErrorString(cudaGetLastError()));
    nm_destroy(h);
    return rc;
  }
  *out = h;
  return NM_OK;
}

extern "C" void nm_destroy(nm_handle* h) {
  if (!h) return;
  cudaSetDevice(h->device);
  nm_buf* bufs[] = {&h->d_block_count, &h->d_deep_rows, &h->d_acc_r2, &h->d_acc_tie, &h->d_acc_mom, &h->d_vals0, &h->d_vals1,
                    &h->d_off0,        &h->d_off1,      &h->d_pos,   &h->d_seg, &h->d_rank, &h->d_seg_cov, &h->d_perm[0], &h->d_perm[1], &h->d_class_scratch, &h->d_rank_keys[0],
                    &h->d_rank_keys[1], &h->d_rank_keys[2], &h->d_rank_order, &h->d_comb_z, &h->d_comb_ln, &h->d_deep_fallback, &h->d_exp0, &h->d_exp1, &h->d_retry, &h->d_head_cand, &h->d_ds_scratch};
  for (nm_buf* b : bufs)
    if (b->p) cudaFree(b->p);
  for (nm_buf& b : h->d_out)
    if (b.p) cudaFree(b.p);
  for (int s = 0; s < 2; ++s) {
    for (nm_buf& b : h->p_in[s])
      if (b.p) cudaFree(b.p);
    for (nm_buf& b : h->p_out[s])
      if (b.p) cudaFree(b.p);
    for (nm_buf& b : h->p_i16[s])
      if (b.p) cudaFree(b.p);
    if (h->ev_in[s]) cudaEventDestroy(h->ev_in[s]);
    if (h->ev_cmp[s]) cudaEventDestroy(h->ev_cmp[s]);
    if (h->ev_out[s]) cudaEventDestroy(h->ev_out[s]);
  }
  if (h->s_in) cudaStreamDestroy(h->s_in);
  if (h->s_out) cudaStreamDestroy(h->s_out);
  if (h->d_sum) cudaFree(h->d_sum);
  for (int s = 0; s < 3; ++s) {
    if (h->h_sums[s]) cudaFreeHost(h->h_sums[s]);
    for (int k = 0; k < 5; ++k)
      if (h->ev_sets[s][k]) cudaEventDestroy(h->ev_sets[s][k]);
  }
  for (int s = 0; s < 2; ++s)
    if (h->ev_done[s]) cudaEventDestroy(h->ev_done[s]);
  if (h->h_head) cudaFreeHost(h->h_head);
  if (h->own_stream) cudaStreamDestroy(h->own_stream);
  free(h);
}

static int nm_check_params(nm_handle* h, const nm_params* p, nm_params* eff) {
  if (!p) return nm_fail(h, NM_ERR_BAD_ARG, "params is NULL");
  *eff = *p;
  if (p->min_coverage < 3)  // NanoMod.py:66
    return nm_fail(h, NM_ERR_BAD_PARAM, "The coverage (%d) is too small", p->min_coverage);
  if (p->nb < 0)  // NanoMod.py:74
    return nm_fail(h, NM_ERR_BAD_PARAM, "The neighborPvalues (%d) cannot be smaller than 0", p->nb);
  if (p->nb > NM_MAX_NB)
    return nm_fail(h, NM_ERR_BAD_PARAM, "neighborPvalues (%d) exceeds NM_MAX_NB (%d)", p->nb, NM_MAX_NB);
  if (p->combine & ~(NM_COMBINE_FISHER | NM_COMBINE_STOUFFER))
    return nm_fail(h, NM_ERR_BAD_PARAM, "unknown combine mask %d", p->combine);
  if (!(p->weights_dif >= 1.0)) eff->weights_dif = 1.0;  // NanoMod.py:77-78
  if (p->ds_times < 0 || p->ds_times > NM_DS_MAX_TIMES)
    return nm_fail(h, NM_ERR_BAD_PARAM, "downsampling (%d) must be in [0, %d]", p->ds_times, NM_DS_MAX_TIMES);
  if (p->ds_times > 0 && (p->ds_index < 0 || p->ds_index >= p->ds_times))
    return nm_fail(h, NM_ERR_BAD_PARAM, "downsampling index (%d) outside [0, %d)", p->ds_index, p->ds_times);
  return NM_OK;
}

// Lane-tier launches.  perm == NULL: one launch over all rows (deep rows are skipped inside).
// Otherwise the rows are partitioned by size group and there is one launch per group --
// <= 64 (the 12-warps/SM instantiation), <= 104 (straight-line networks, 8 warps/SM), <= 128
// (looped networks, 6 warps/SM) -- each with shared memory sized for its own largest class.
static int nm_launch_tiers(nm_handle* h, const nm_kargs& ka, bool want_u, bool want_t, bool want_m, const nm_summary& sum,
                           const int32_t* perm, int deep_smem, int64_t n_rows, int64_t n_pos, cudaStream_t st) {
  const int n_deep = sum.n_deep, max_lane_n = sum.max_lane_n, max_deep_p2 = sum.max_deep_p2;
  NM_CUDA(h, cudaEventRecord(h->ev[1], st));
  if (n_rows > n_deep) {
    const int sms = h->sm_limit > 0 ? h->sm_limit : h->sm_count;
    nm_kargs kl = ka;
    kl.perm = perm;
    kl.gaps = (perm != nullptr || n_deep > 0 || n_rows != n_pos) ? 1 : 0;
    if (!perm) {
      kl.row_lo = 0;
      kl.row_hi = n_rows;
      kl.tile_cursor = &h->d_sum->tile_cursor[0];
      const cudaError_t e = (cudaError_t)nm_launch_lane(kl, want_u, want_m, max_lane_n, sms, st);
      if (e != cudaSuccess)
        return nm_fail(h, NM_ERR_CUDA, "nm_lane_kernel launch failed: %s", cudaGetErrorString(e));
      h->launches++;
    } else {
      const int64_t n_lane = n_rows - n_deep;
      const int64_t cut[4] = {0, sum.n_le64, sum.n_le104, n_lane};
      const int cap[3] = {64, NM_LANE_FINE_MAX, NM_LANE_TIER_MAX};
      for (int g = 0; g < 3; ++g) {
        if (cut[g + 1] <= cut[g]) continue;
        kl.row_lo = cut[g];
        kl.row_hi = cut[g + 1];
        kl.tile_cursor = &h->d_sum->tile_cursor[g];
        const int max_n = max_lane_n < cap[g] ? max_lane_n : cap[g];
        const cudaError_t e = (cudaError_t)nm_launch_lane(kl, want_u, want_m, max_n, sms, st);
        if (e != cudaSuccess) return nm_fail(h, NM_ERR_CUDA, "nm_lane_kernel launch failed: %s", cudaGetErrorString(e));
        h->launches++;
      }
    }
  }
  NM_CUDA(h, cudaEventRecord(h->ev[2], st));
  if (n_deep > 0) {
    nm_kargs kd = ka;
    if (!h->no_grid) {  // positions whose values are three-place decimals sort 16-bit key pairs; the others are listed
      const int rc_r = nm_reserve(h, &h->d_retry, sizeof(int32_t) * (size_t)n_deep);
      if (rc_r != NM_OK) return rc_r;
      kd.deep_retry_rows = (int32_t*)h->d_retry.p;
      kd.deep_retry_count = &h->d_sum->deep_fallback_count;
    }
    if (n_deep > sum.n_huge) {
      const cudaError_t e = (cudaError_t)nm_launch_deep(kd, want_u, want_t, want_m, n_deep, max_deep_p2, deep_smem, h->sm_count, st);
      if (e != cudaSuccess)
        return nm_fail(h, NM_ERR_CUDA, "nm_deep_kernel launch failed: %s", cudaGetErrorString(e));
      h->launches++;
    }
    if (sum.n_huge > 0) {  // rows too long for the deep tier's shared memory (the reference has no depth limit)
      int rc = nm_reserve(h, &h->d_deep_fallback, nm_huge_scratch_bytes(sum.n_huge, (long long)sum.huge_v0, (long long)sum.huge_v1));
      if (rc != NM_OK) return rc;
      int launches = 0;
      const cudaError_t e = (cudaError_t)nm_huge_run(ka, want_u, want_t, want_m, sum.n_huge, (long long)sum.huge_v0,
                                                     (long long)sum.huge_v1, h->d_deep_fallback.p, &launches, st);
      h->launches += launches;
      if (e != cudaSuccess) return nm_fail(h, NM_ERR_CUDA, "huge-row path failed: %s", cudaGetErrorString(e));
    }
  }
  if ((want_u || want_t) && n_rows > n_deep) {
    nm_tails_args ta;
    memset(&ta, 0, sizeof(ta));
    ta.row_n0 = ka.row_n0; ta.row_n1 = ka.row_n1;
    ta.acc_r2 = ka.acc_r2; ta.acc_tie = ka.acc_tie; ta.acc_mom = ka.acc_mom;
    ta.n_rows = n_rows; ta.want_u = want_u; ta.want_t = want_t;
    ta.two_u = ka.two_u; ta.u_stat = ka.u_stat; ta.u_p = ka.u_p; ta.t_stat = ka.t_stat; ta.t_p = ka.t_p;
    ta.flags = ka.flags;
    nm_tails_kernel<<<(unsigned)((n_rows + 255) / 256), 256, 0, st>>>(ta);
    NM_CUDA(h, cudaGetLastError());
    h->launches++;
  }
  NM_CUDA(h, cudaEventRecord(h->ev[3], st));
  return NM_OK;
}

// 16-bit transport format -> the float32 values every kernel works on: (float)((double)k * unit),
// the float32 nearest to the decimal k * unit (a float64 product is within 1e-16 of it, and a decimal
// with three places is never that close to a float32 rounding boundary)
__global__ void __launch_bounds__(256) nm_expand_i16(const int16_t* __restrict__ in, float* __restrict__ out, int64_t n,
                                                     double unit) {
  const int64_t i = ((int64_t)blockIdx.x * 256 + threadIdx.x) * 8;
  if (i + 8 <= n) {
    const int4 w = *reinterpret_cast<const int4*>(in + i);
    const int v[4] = {w.x, w.y, w.z, w.w};
    float o[8];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      o[2 * k] = (float)((double)(short)(v[k] & 0xffff) * unit);
      o[2 * k + 1] = (float)((double)(short)(v[k] >> 16) * unit);
    }
    *reinterpret_cast<float4*>(out + i) = make_float4(o[0], o[1], o[2], o[3]);
    *reinterpret_cast<float4*>(out + i + 4) = make_float4(o[4], o[5], o[6], o[7]);
  } else {
    for (int64_t k = i; k < n; ++k) out[k] = (float)((double)in[k] * unit);
  }
}

// nm_grid_selftest: the grid-key statement of nm_lane.cuh on every float32 pattern, on the device
// (the arithmetic the lane kernel runs: FFMA / FADD / FSETP as compiled for sm_100a).
__global__ void __launch_bounds__(256) nm_grid_selftest_kernel(unsigned long long* out /* [2]: violations, passes */) {
  unsigned long long viol = 0, pass = 0;
  const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
  for (unsigned long long b = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; b < (1ull << 32); b += stride) {
    const float x = __uint_as_float((unsigned)b);
    nm_grid_flag bad = NM_GRID_FLAG0, bad2 = NM_GRID_FLAG0;
    const unsigned ta = nm_grid_bits(x, NM_GRID_MA, &bad);
    const unsigned tb = nm_grid_bits(x, NM_GRID_MB, &bad2);
    const bool ok = !nm_grid_failed(bad) && fabsf(x) <= NM_GRID_LIM, ok2 = !nm_grid_failed(bad2) && fabsf(x) <= NM_GRID_LIM;
    if (ok != ok2) ++viol;
    if (!ok) continue;
    ++pass;
    const int k = (int)(ta & 0xffffu) - 32768;
    const float canon = (float)((double)k / 1000.0);
    if (!(canon == x) || k < -32766 || k > 32766) ++viol;
    if ((ta >> 16) != 0x4B40u) ++viol;
    const unsigned packed = tb * 65536u + ta;
    if ((packed >> 16) != (unsigned)(k + 32768) || (packed & 0xffffu) != (unsigned)(k + 32768)) ++viol;
  }
  for (int o = 16; o > 0; o >>= 1) {
    viol += __shfl_xor_sync(0xffffffffu, viol, o);
    pass += __shfl_xor_sync(0xffffffffu, pass, o);
  }
  if ((threadIdx.x & 31) == 0 && (viol | pass)) {
    atomicAdd(out, viol);
    atomicAdd(out + 1, pass);
  }
}

extern "C" int nm_grid_selftest(nm_handle* h, int64_t* violations, int64_t* passes) {
  if (!h || !violations || !passes) return NM_ERR_BAD_ARG;
  unsigned long long* d = nullptr;
  NM_CUDA(h, cudaMalloc(&d, 2 * sizeof(unsigned long long)));
  cudaMemsetAsync(d, 0, 2 * sizeof(unsigned long long), h->own_stream);
  nm_grid_selftest_kernel<<<h->sm_count * 8, 256, 0, h->own_stream>>>(d);
  unsigned long long r[2] = {0, 0};
  cudaError_t e = cudaMemcpyAsync(r, d, sizeof(r), cudaMemcpyDeviceToHost, h->own_stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(h->own_stream);
  cudaFree(d);
  NM_CUDA(h, e);
  *violations = (int64_t)r[0];
  *passes = (int64_t)r[1];
  return NM_OK;
}

static int nm_expand_i16_run(nm_handle* h, const int16_t* in, float* out, int64_t n, double unit, cudaStream_t st) {
  if (n <= 0) return NM_OK;
  nm_expand_i16<<<(unsigned)((n + 2047) / 2048), 256, 0, st>>>(in, out, n, unit);
  NM_CUDA(h, cudaGetLastError());
  h->launches++;
  return NM_OK;
}

// the shape the dense lane kernel assumes (it re-checks the same conditions on the device)
static bool nm_dense_shape(const nm_summary& s) { return nm_dense_shape_ok(s); }

static void nm_launch_combine(const nm_comb_args& ca, int64_t n_rows, cudaStream_t st) {
  const unsigned grid = (unsigned)((n_rows + NM_COMB_THREADS - 1) / NM_COMB_THREADS);
  switch (ca.nb) {
    case 1: nm_combine_kernel<1><<<grid, NM_COMB_THREADS, 0, st>>>(ca); break;
    case 2: nm_combine_kernel<2><<<grid, NM_COMB_THREADS, 0, st>>>(ca); break;
    case 3: nm_combine_kernel<3><<<grid, NM_COMB_THREADS, 0, st>>>(ca); break;
    case 4: nm_combine_kernel<4><<<grid, NM_COMB_THREADS, 0, st>>>(ca); break;
    case 5: nm_combine_kernel<5><<<grid, NM_COMB_THREADS, 0, st>>>(ca); break;
    default: nm_combine_kernel<0><<<grid, NM_COMB_THREADS, 0, st>>>(ca); break;
  }
}

static void nm_fill_comb_args(nm_comb_args* ca, const nm_pileup* pl, const nm_params& prm, const nm_table* tb,
                              int64_t n_rows, bool ds_on) {
  memset(ca, 0, sizeof(*ca));
  ca->ks_p = tb->ks_p; ca->ks_dnum = tb->ks_dnum; ca->row_pos_index = tb->row_pos_index;
  ca->row_n0 = tb->n0; ca->row_n1 = tb->n1; ca->pos = pl->pos; ca->seg = pl->seg;
  ca->ks_d = tb->ks_d; ca->seg_cov = ds_on ? pl->seg_cov : nullptr;
  ca->n_rows = n_rows; ca->nb = prm.nb;
  ca->want_fisher = (prm.combine & NM_COMBINE_FISHER) != 0; ca->want_stouffer = (prm.combine & NM_COMBINE_STOUFFER) != 0;
  ca->wnorm = nm_build_weights(prm.nb, prm.weights_dif, ca->w);
  ca->f_stat = tb->fisher_stat; ca->f_p = tb->fisher_p; ca->s_stat = tb->stouffer_stat; ca->s_p = tb->stouffer_p;
}

// The armed head selection (nm_arm_head_select), launched behind the call's last kernel.  It was armed for a
// call whose rows are its candidates; any other outcome leaves it unfired and the caller selects afterwards.
#define NM_HEAD_CAND_MIN 65536
static int nm_head_cand_cap(int64_t want) { return (int)(32 * want > NM_HEAD_CAND_MIN ? 32 * want : NM_HEAD_CAND_MIN); }

// Lets the combine kernel of this call list the candidates of the armed head selection, when the selection's
// primary key is the combined p-value column the kernel writes (rankUse = 'pv') over a row range of this table.
static int nm_head_cands_setup(nm_handle* h, nm_comb_args* ca, const nm_table* tb, int64_t n_rows) {
  h->head.use_cands = 0;
  if (h->no_head_cands || !h->head.armed || h->head.reverse || ca->nb <= 0 || !h->head.key[0] || h->head.want > (1 << 20)) return NM_OK;
  int col = -1;
  const double* base = nullptr;
  if (ca->want_stouffer && h->head.key[0] >= tb->stouffer_p && h->head.key[0] < tb->stouffer_p + n_rows) { col = 0; base = tb->stouffer_p; }
  else if (ca->want_fisher && h->head.key[0] >= tb->fisher_p && h->head.key[0] < tb->fisher_p + n_rows) { col = 1; base = tb->fisher_p; }
  if (col < 0) return NM_OK;
  const int64_t lo = h->head.key[0] - base;
  if (lo + h->head.n_rows > n_rows) return NM_OK;
  unsigned thr = nm_head_thr_bin(h->head.n_rows, h->head.want);  // what a null table needs
  if (!thr) return NM_OK;
  // a table with many significant rows holds far more rows below that than the head needs: after a selection
  // of the same size that succeeded, list only up to two bins above its cut
  if (h->head_cut_hint > 0 && h->head_hint_n == h->head.n_rows && h->head_hint_want == h->head.want &&
      (unsigned)h->head_cut_hint + 2u < thr)
    thr = (unsigned)h->head_cut_hint + 2u;
  const int cap = nm_head_cand_cap(h->head.want);
  const int rc = nm_reserve(h, &h->d_head_cand, sizeof(int2) * (size_t)cap);
  if (rc != NM_OK) return rc;
  ca->head_cand = (int2*)h->d_head_cand.p;
  ca->head_cursor = &h->d_sum->head_cursor;  // zeroed with the summary at the start of the call
  ca->head_lo = lo; ca->head_hi = lo + h->head.n_rows;
  ca->head_thr_bin = thr; ca->head_cap = cap; ca->head_col = col;
  h->head.use_cands = 1;
  h->head.thr_bin = thr;
  return NM_OK;
}

// The armed head selection (nm_arm_head_select), launched behind the call's last kernel.  It was armed for a
// call whose rows are its candidates; any other outcome leaves it unfired and the caller selects afterwards.
static int nm_fire_armed_head(nm_handle* h, int64_t n_rows, int64_t n_pos, cudaStream_t st) {
  h->head.fired = 0;
  const int use_cands = h->head.use_cands;
  h->head.use_cands = 0;
  if (!h->head.armed || n_rows != n_pos) return NM_OK;
  int rc = nm_reserve(h, &h->d_rank, nm_head_scratch_bytes(1));
  if (rc != NM_OK) return rc;
  int launches = 0;
  const nm_head_peers_dev* peers = h->head.peers.n > 0 ? &h->head.peers : nullptr;
  cudaError_t e;
  if (use_cands) {
    nm_head_peers_dev pr = h->head.peers;  // the refusal flag is wanted with or without peers
    e = (cudaError_t)nm_head_from_cands_run(h->head.key[0], h->head.key[1], h->head.key[2], h->head.n_rows, h->head.want,
                                            h->head.cap, h->head.geo, h->d_head_cand.p, &h->d_sum->head_cursor,
                                            nm_head_cand_cap(h->head.want), h->head.thr_bin,
                                            h->head.records, &h->d_sum->head_fail, &h->d_sum->head_cut, &launches, st, &pr);
    h->head_hint_n = h->head.n_rows;  // what the hint of this call's summary will be about
    h->head_hint_want = h->head.want;
  } else {
    e = (cudaError_t)nm_head_run(h->head.key[0], h->head.key[1], h->head.key[2], h->head.n_rows, h->head.reverse,
                                 h->head.want, h->head.cap, h->head.geo, h->d_rank.p, h->head.records,
                                 h->sm_count, &launches, st, peers);
  }
  h->launches += launches;
  if (e != cudaSuccess) return nm_fail(h, NM_ERR_CUDA, "armed head selection failed: %s", cudaGetErrorString(e));
  h->head.fired = 1;
  return NM_OK;
}

// Dense path: plan_count has been launched; lane kernel (+ U/t tails) + combine stencil, one sync
// at the end.  *refused is set when the kernel found another shape than `class_n` was sized for
// (only possible for a speculative launch); nothing was computed then.
// Split in two for nm_detect_device_async: nm_dense_enqueue issues everything up to the read-back of the
// summary into h->h_sum, nm_dense_complete looks at it once the stream has got there.
static int nm_dense_enqueue(nm_handle* h, const nm_pileup* pl, const nm_params& prm, const nm_table* tb, int class_n,
                            cudaStream_t st, int* try_grid_out, int* n_launched_out) {
  const bool want_u = prm.want_u != 0, want_t = prm.want_t != 0;
  const bool want_f = (prm.combine & NM_COMBINE_FISHER) != 0, want_s = (prm.combine & NM_COMBINE_STOUFFER) != 0;
  const int64_t n = pl->n_pos;
  int rc;
  nm_kargs ka;
  memset(&ka, 0, sizeof(ka));
  ka.vals0 = pl->vals0; ka.vals1 = pl->vals1; ka.off0 = pl->off0; ka.off1 = pl->off1;
  ka.row_pos_index = tb->row_pos_index; ka.row_n0 = tb->n0; ka.row_n1 = tb->n1;
  ka.w_row_pos_index = tb->row_pos_index; ka.w_n0 = tb->n0; ka.w_n1 = tb->n1;
  ka.n_rows = n; ka.n_pos = n; ka.one = 1; ka.mone = -1;
  ka.grid_tries = NM_GRID_TRIES;
  ka.ks_dnum = tb->ks_dnum; ka.ks_d = tb->ks_d; ka.ks_p = tb->ks_p;
  ka.two_u = tb->two_u; ka.u_stat = tb->u_stat; ka.u_p = tb->u_p;
  ka.t_stat = tb->t_stat; ka.t_p = tb->t_p; ka.flags = tb->flags;
  ka.sum = h->d_sum;
  ka.tile_cursor = &h->d_sum->dense_tile_cursor;
  if (want_u) {
    if ((rc = nm_reserve(h, &h->d_acc_r2, sizeof(int) * (size_t)n)) != NM_OK) return rc;
    if ((rc = nm_reserve(h, &h->d_acc_tie, sizeof(int) * (size_t)n)) != NM_OK) return rc;
    ka.acc_r2 = (int*)h->d_acc_r2.p;
    ka.acc_tie = (int*)h->d_acc_tie.p;
  }
  const bool want_m = want_t || tb->moments != nullptr;
  if (tb->moments) {
    ka.acc_mom = tb->moments;
  } else if (want_t) {
    if ((rc = nm_reserve(h, &h->d_acc_mom, sizeof(double) * 4 * (size_t)n)) != NM_OK) return rc;
    ka.acc_mom = (double*)h->d_acc_mom.p;
  }
  if (prm.nb > 0 && want_s) {
    if ((rc = nm_reserve(h, &h->d_comb_z, sizeof(double) * (size_t)n)) != NM_OK) return rc;
    ka.comb_z = (double*)h->d_comb_z.p;
  }
  if (prm.nb > 0 && want_f) {
    if ((rc = nm_reserve(h, &h->d_comb_ln, sizeof(double) * (size_t)n)) != NM_OK) return rc;
    ka.comb_ln = (double*)h->d_comb_ln.p;
  }
  const int sms = h->sm_limit > 0 ? h->sm_limit : h->sm_count;
  // Grid keys first (16-bit sort of three-place decimals, checked value by value on the device), then the
  // float32 kernel over whatever that launch listed or left unclaimed -- normally nothing.  After a call
  // whose data were not on the grid the attempt is skipped for a while.
  // not for short rows (the check costs what the packed sort saves) and not with the rank statistics (their walk over
  // 16-bit columns is slower than the float32 one by more than the sort gains: cfg3 2.65 -> 2.79 ms)
  const bool try_grid = !h->no_grid && h->grid_skip == 0 && class_n > 64 && (!want_u || h->grid_u);
  cudaError_t e;
  if (try_grid) {
    if ((rc = nm_reserve(h, &h->d_retry, sizeof(int32_t) * (size_t)((n + 31) / 32))) != NM_OK) return rc;
    ka.retry_tiles = (int32_t*)h->d_retry.p;
    e = (cudaError_t)nm_launch_lane_dense(ka, want_u, want_m, class_n, sms, true, st);
    if (e != cudaSuccess) return nm_fail(h, NM_ERR_CUDA, "nm_lane_dense_kernel (grid keys) launch failed: %s", cudaGetErrorString(e));
    h->launches++;
    ka.retry_mode = 1;
    ka.tile_cursor = &h->d_sum->retry_cursor;
  }
  e = (cudaError_t)nm_launch_lane_dense(ka, want_u, want_m, class_n, sms, false, st);
  if (e != cudaSuccess) return nm_fail(h, NM_ERR_CUDA, "nm_lane_dense_kernel launch failed: %s", cudaGetErrorString(e));
  h->launches++;
  NM_CUDA(h, cudaEventRecord(h->ev[2], st));
  if (want_u || want_t) {
    nm_tails_args ta;
    memset(&ta, 0, sizeof(ta));
    ta.row_n0 = tb->n0; ta.row_n1 = tb->n1;
    ta.acc_r2 = ka.acc_r2; ta.acc_tie = ka.acc_tie; ta.acc_mom = ka.acc_mom;
    ta.n_rows = n; ta.want_u = want_u; ta.want_t = want_t;
    ta.two_u = tb->two_u; ta.u_stat = tb->u_stat; ta.u_p = tb->u_p; ta.t_stat = tb->t_stat; ta.t_p = tb->t_p;
    ta.flags = tb->flags;
    nm_tails_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(ta);
    NM_CUDA(h, cudaGetLastError());
    h->launches++;
  }
  NM_CUDA(h, cudaEventRecord(h->ev[3], st));
  if (want_f || want_s) {
    nm_comb_args ca;
    nm_fill_comb_args(&ca, pl, prm, tb, n, false);
    ca.row_pos_index = nullptr;  // row == candidate
    ca.z_pre = ka.comb_z;
    ca.ln_pre = ka.comb_ln;
    if ((rc = nm_head_cands_setup(h, &ca, tb, n)) != NM_OK) return rc;
    nm_launch_combine(ca, n, st);
    NM_CUDA(h, cudaGetLastError());
    h->launches++;
  }
  NM_CUDA(h, cudaEventRecord(h->ev[4], st));
  if ((rc = nm_fire_armed_head(h, n, n, st)) != NM_OK) return rc;
  NM_CUDA(h, cudaMemcpyAsync(h->h_sum, h->d_sum, sizeof(nm_summary), cudaMemcpyDeviceToHost, st));
  *try_grid_out = try_grid ? 1 : 0;
  *n_launched_out = (try_grid ? 2 : 1) + ((want_u || want_t) ? 1 : 0) + ((want_f || want_s) ? 1 : 0);
  return NM_OK;
}

static int nm_dense_complete(nm_handle* h, int try_grid, int n_launched, nm_summary* sum_out, bool* refused) {
  *sum_out = *h->h_sum;
  h->last_grid_tiles = sum_out->grid_tiles;
  *refused = sum_out->dense_retry != 0;
  if (sum_out->head_fail) h->head.fired = 0;  // the candidate list could not give the head: the caller selects
  if (!*refused && (sum_out->head_cut > 0 || sum_out->head_fail)) h->head_cut_hint = sum_out->head_fail ? 0 : sum_out->head_cut;
  if (try_grid && !*refused)
    h->grid_skip = sum_out->grid_giveup ? NM_GRID_SKIP_CALLS : 0;
  else if (h->grid_skip > 0 && !*refused)
    --h->grid_skip;
  if (*refused) {
    h->launches -= n_launched;  // they did not compute
    // the refused kernels ran on an unvalidated shape: the tails / combine launches read rows the
    // lane kernel never wrote, which is harmless (every column is rewritten by the general path)
    return NM_OK;
  }
  float ms = 0.f;
  if (cudaEventElapsedTime(&ms, h->ev[0], h->ev[1]) == cudaSuccess) h->last_ms[0] = ms;
  if (cudaEventElapsedTime(&ms, h->ev[1], h->ev[2]) == cudaSuccess) h->last_ms[1] = ms;
  if (cudaEventElapsedTime(&ms, h->ev[2], h->ev[3]) == cudaSuccess) h->last_ms[2] = ms;
  if (cudaEventElapsedTime(&ms, h->ev[3], h->ev[4]) == cudaSuccess) h->last_ms[3] = ms;
  return NM_OK;
}

static int nm_run_dense(nm_handle* h, const nm_pileup* pl, const nm_params& prm, const nm_table* tb, int class_n,
                        cudaStream_t st, nm_summary* sum_out, bool* refused) {
  int try_grid = 0, n_launched = 0;
  const int rc = nm_dense_enqueue(h, pl, prm, tb, class_n, st, &try_grid, &n_launched);
  if (rc != NM_OK) return rc;
  NM_CUDA(h, cudaStreamSynchronize(st));
  return nm_dense_complete(h, try_grid, n_launched, sum_out, refused);
}

// argument checks of a device call on float32 values
static int nm_check_call(nm_handle* h, const nm_pileup* pl, const nm_params& prm, const nm_table* tb) {
  const bool want_u = prm.want_u != 0, want_t = prm.want_t != 0;
  const bool want_f = (prm.combine & NM_COMBINE_FISHER) != 0, want_s = (prm.combine & NM_COMBINE_STOUFFER) != 0;
  if (!pl->vals0 || !pl->vals1 || !pl->off0 || !pl->off1 || !pl->pos || !pl->seg)
    return nm_fail(h, NM_ERR_BAD_ARG, "pileup has a NULL array");
  if ((((uintptr_t)pl->vals0) | ((uintptr_t)pl->vals1)) & 15)
    return nm_fail(h, NM_ERR_BAD_ARG, "vals0/vals1 must be 16-byte aligned");
  if (!tb->row_pos_index || !tb->n0 || !tb->n1 || !tb->ks_dnum || !tb->ks_p)
    return nm_fail(h, NM_ERR_BAD_ARG, "table lacks a mandatory output (row_pos_index,n0,n1,ks_dnum,ks_p)");
  if (want_u && (!tb->two_u || !tb->u_p)) return nm_fail(h, NM_ERR_BAD_ARG, "want_u needs two_u and u_p");
  if (want_t && (!tb->t_stat || !tb->t_p)) return nm_fail(h, NM_ERR_BAD_ARG, "want_t needs t_stat and t_p");
  if (want_f && (!tb->fisher_stat || !tb->fisher_p)) return nm_fail(h, NM_ERR_BAD_ARG, "fisher outputs missing");
  if (want_s && (!tb->stouffer_stat || !tb->stouffer_p)) return nm_fail(h, NM_ERR_BAD_ARG, "stouffer outputs missing");
  return NM_OK;
}

static int nm_detect_device_impl(nm_handle* h, const nm_pileup* pl, const nm_params* params, const nm_table* tb,
                                 int64_t* n_rows_out, void* cuda_stream);
extern "C" int nm_detect_device(nm_handle* h, const nm_pileup* pl, const nm_params* params,
                                const nm_table* tb, int64_t* n_rows_out, void* cuda_stream) {
  if (!h) return nm_fail(nullptr, NM_ERR_BAD_ARG, "handle is NULL");
  h->head.fired = 0;
  h->head.use_cands = 0;  // set again by this call's own combine launch, if it lists candidates
  const int rc = nm_detect_device_impl(h, pl, params, tb, n_rows_out, cuda_stream);
  h->head.armed = 0;  // one shot (nm_arm_head_select)
  if (rc != NM_OK) h->head.fired = 0;
  return rc;
}
static int nm_detect_device_impl(nm_handle* h, const nm_pileup* pl, const nm_params* params, const nm_table* tb,
                                 int64_t* n_rows_out, void* cuda_stream) {
  if (!h) return nm_fail(nullptr, NM_ERR_BAD_ARG, "handle is NULL");
  if (!pl || !tb || !n_rows_out) return nm_fail(h, NM_ERR_BAD_ARG, "pileup/table/n_rows is NULL");
  nm_params prm;
  int rc = nm_check_params(h, params, &prm);
  if (rc != NM_OK) return rc;
  const bool want_u = prm.want_u != 0, want_t = prm.want_t != 0;
  const bool want_f = (prm.combine & NM_COMBINE_FISHER) != 0, want_s = (prm.combine & NM_COMBINE_STOUFFER) != 0;
  if (pl->n_pos < 0 || pl->n_pos > 0x7fffffffLL - NM_PLAN_PER_BLOCK)
    return nm_fail(h, NM_ERR_BAD_ARG, "n_pos (%lld) out of range", (long long)pl->n_pos);
  *n_rows_out = 0;
  if (pl->n_pos == 0) return NM_OK;
  nm_pileup pl_f32;  // int16 transport format: expanded into the handle's float32 scratch first
  if (!pl->vals0 && !pl->vals1 && pl->vals0_i16 && pl->vals1_i16 && pl->off0 && pl->off1) {
    if (!(pl->i16_unit > 0.0)) return nm_fail(h, NM_ERR_BAD_ARG, "i16_unit must be positive");
    if ((((uintptr_t)pl->vals0_i16) | ((uintptr_t)pl->vals1_i16)) & 15)
      return nm_fail(h, NM_ERR_BAD_ARG, "vals0_i16/vals1_i16 must be 16-byte aligned");
    if (pl->i16_total0 < 0 || pl->i16_total1 < 0)
      return nm_fail(h, NM_ERR_BAD_ARG, "int16 pileups on the device need i16_total0 / i16_total1 (values per group)");
    NM_CUDA(h, cudaSetDevice(h->device));
    if ((rc = nm_reserve(h, &h->d_exp0, sizeof(float) * (size_t)nm_padded_len(pl->i16_total0))) != NM_OK) return rc;
    if ((rc = nm_reserve(h, &h->d_exp1, sizeof(float) * (size_t)nm_padded_len(pl->i16_total1))) != NM_OK) return rc;
    cudaStream_t st0 = (cudaStream_t)cuda_stream;
    if ((rc = nm_expand_i16_run(h, pl->vals0_i16, (float*)h->d_exp0.p, pl->i16_total0, pl->i16_unit, st0)) != NM_OK) return rc;
    if ((rc = nm_expand_i16_run(h, pl->vals1_i16, (float*)h->d_exp1.p, pl->i16_total1, pl->i16_unit, st0)) != NM_OK) return rc;
    pl_f32 = *pl;
    pl_f32.vals0 = (const float*)h->d_exp0.p;
    pl_f32.vals1 = (const float*)h->d_exp1.p;
    pl_f32.vals0_i16 = pl_f32.vals1_i16 = nullptr;
    pl = &pl_f32;
  }
  if ((rc = nm_check_call(h, pl, prm, tb)) != NM_OK) return rc;

  const bool ds_on = pl->seg_cov != nullptr && pl->n_seg > 0 && prm.ds_times > 0;
  NM_CUDA(h, cudaSetDevice(h->device));
  cudaStream_t st = (cudaStream_t)cuda_stream;
  const int64_t n_pos = pl->n_pos;
  const int nblk = (int)((n_pos + NM_PLAN_PER_BLOCK - 1) / NM_PLAN_PER_BLOCK);
  if ((rc = nm_reserve(h, &h->d_block_count, sizeof(int) * (size_t)nblk)) != NM_OK) return rc;
  if ((rc = nm_reserve(h, &h->d_deep_rows, sizeof(int32_t) * (size_t)n_pos)) != NM_OK) return rc;

  // ---- plan, first pass: coverage filter counts and the shape summary
  for (int k = 0; k < 4; ++k) h->last_ms[k] = 0.0;
  NM_CUDA(h, cudaEventRecord(h->ev[0], st));
  NM_CUDA(h, cudaMemsetAsync(h->d_sum, 0, sizeof(nm_summary), st));
  nm_plan_count<<<nblk, NM_PLAN_THREADS, 0, st>>>(pl->off0, pl->off1, n_pos, prm.min_coverage, pl->seg,
                                                  pl->n_seg > 0 ? pl->n_seg : 0, ds_on ? pl->seg_cov : nullptr,
                                                  (int*)h->d_block_count.p, h->d_sum);
  NM_CUDA(h, cudaGetLastError());
  h->launches += 1;
  NM_CUDA(h, cudaEventRecord(h->ev[1], st));

  // ---- dense path: nothing filtered, nothing deep, one size group.  After a call of that shape
  // the next one is launched on the same assumption WITHOUT waiting for its summary (the kernel
  // validates it on the device and refuses to compute if it does not hold).
  const bool dense_allowed = !h->no_dense && !ds_on;
  nm_summary sum;
  bool have_sum = false;
  int dense_class = 0;
  h->last_path = 0;
  if (dense_allowed && h->dense_class > 0) {
    dense_class = h->dense_class;
    h->last_path = 2;
  } else {
    NM_CUDA(h, cudaMemcpyAsync(h->h_sum, h->d_sum, sizeof(nm_summary), cudaMemcpyDeviceToHost, st));
    NM_CUDA(h, cudaStreamSynchronize(st));
    sum = *h->h_sum;
    have_sum = true;
    if (dense_allowed && nm_dense_shape(sum)) {
      dense_class = nm_lane_class(sum.max_lane_n);
      h->last_path = 1;
    }
  }
  if (dense_class > 0) {
    bool refused = false;
    rc = nm_run_dense(h, pl, prm, tb, dense_class, st, &sum, &refused);
    if (rc != NM_OK) return rc;
    if (!refused) {
      h->dense_class = nm_lane_class(sum.max_lane_n);
      *n_rows_out = n_pos;
      return NM_OK;
    }
    have_sum = true;  // nm_run_dense read the summary back
    h->last_path = 3;
    if (nm_dense_shape(sum)) {  // dense after all, only another network class: launch it again, sized right
      NM_CUDA(h, cudaMemsetAsync(&h->d_sum->dense_retry, 0, 7 * sizeof(int), st));  // + the cursors and grid-key counters
      NM_CUDA(h, cudaMemsetAsync(&h->d_sum->head_cursor, 0, 2 * sizeof(int), st));   // + the head candidates of the refused pass
      rc = nm_run_dense(h, pl, prm, tb, nm_lane_class(sum.max_lane_n), st, &sum, &refused);
      if (rc != NM_OK) return rc;
      if (refused) return nm_fail(h, NM_ERR_CUDA, "internal: dense launch refused after validation");
      h->dense_class = nm_lane_class(sum.max_lane_n);
      *n_rows_out = n_pos;
      return NM_OK;
    }
    h->last_path = 4;
  }
  h->dense_class = 0;
  if (!have_sum) return nm_fail(h, NM_ERR_CUDA, "internal: plan summary missing");
  if (sum.bad_input)
    return nm_fail(h, NM_ERR_BAD_ARG, "offsets are not non-decreasing, or a segment id is outside [0, n_seg)");
  if (ds_on && (sum.ds_too_deep & 2))  // known from the plan pass: refuse before any test has run
    return nm_fail(h, NM_ERR_TOO_DEEP,
                   "--coverages: a position to be down-sampled has a group of more than %d reads, or more than %d reads with a "
                   "threshold above %d; the down-sampling branch does not take it (run without --coverages, or thin the pileup first)",
                   NM_DS_DEEP_MAX_READS, NM_DS_MAX_READS, NM_DS_DEEP_MAX_COV);

  // ---- plan, second pass: ordered compaction of the kept candidates into rows
  {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, h->ev[0], h->ev[1]) == cudaSuccess) h->last_ms[0] = ms;
  }
  const int64_t n_rows = n_pos - (int64_t)sum.n_filtered;
  *n_rows_out = n_rows;
  if (n_rows == 0) return NM_OK;
  nm_plan_scan<<<1, NM_PLAN_THREADS, 0, st>>>((int*)h->d_block_count.p, nblk, h->d_sum);
  nm_plan_scatter<<<nblk, NM_PLAN_THREADS, 0, st>>>(pl->off0, pl->off1, n_pos, prm.min_coverage,
                                                    (const int*)h->d_block_count.p, tb->row_pos_index,
                                                    tb->n0, tb->n1, (int32_t*)h->d_deep_rows.p, h->d_sum);
  NM_CUDA(h, cudaGetLastError());
  h->launches += 2;

  // ---- per-position tests
  nm_kargs ka;
  memset(&ka, 0, sizeof(ka));
  ka.vals0 = pl->vals0; ka.vals1 = pl->vals1; ka.off0 = pl->off0; ka.off1 = pl->off1;
  ka.row_pos_index = tb->row_pos_index; ka.row_n0 = tb->n0; ka.row_n1 = tb->n1;
  ka.n_rows = n_rows;
  ka.one = 1;
  ka.mone = -1;
  ka.sum = h->d_sum;
  ka.ks_dnum = tb->ks_dnum; ka.ks_d = tb->ks_d; ka.ks_p = tb->ks_p;
  ka.two_u = tb->two_u; ka.u_stat = tb->u_stat; ka.u_p = tb->u_p;
  ka.t_stat = tb->t_stat; ka.t_p = tb->t_p; ka.flags = tb->flags;
  ka.deep_rows = (const int32_t*)h->d_deep_rows.p; ka.n_deep = sum.n_deep;
  if (want_u) {
    if ((rc = nm_reserve(h, &h->d_acc_r2, sizeof(int) * (size_t)n_rows)) != NM_OK) return rc;
    if ((rc = nm_reserve(h, &h->d_acc_tie, sizeof(int) * (size_t)n_rows)) != NM_OK) return rc;
    ka.acc_r2 = (int*)h->d_acc_r2.p;
    ka.acc_tie = (int*)h->d_acc_tie.p;
  }
  // group moments: straight into the caller's array when asked for (--mstd), else scratch for t
  const bool want_m = want_t || tb->moments != nullptr;
  if (tb->moments) {
    ka.acc_mom = tb->moments;
  } else if (want_t) {
    if ((rc = nm_reserve(h, &h->d_acc_mom, sizeof(double) * 4 * (size_t)n_rows)) != NM_OK) return rc;
    ka.acc_mom = (double*)h->d_acc_mom.p;
  }
  // two groups, each stored skewed by one word per 32 after its sort (nm_deep_kernel.cu)
  const int deep_smem = 16 + (sum.max_deep_p2 + (sum.max_deep_p2 >> 5) + 32) * (int)sizeof(float);
  // Mixed coverage.  The lane kernel runs ONE network size per launch (several sizes in flight
  // thrash the instruction cache), normally that of the call's longest row.  If the rows span
  // several size groups and one group holds nearly all of them, the others are outliers: split
  // the call into one launch per group, so that 1 % of deep positions do not make the other 99 %
  // pay for their network.  (With a broad coverage distribution the groups interleave densely,
  // tiles stop being contiguous, and staging them row by row costs more than the smaller
  // networks save: profiles/round1_variants.md.)
  const int32_t* perm = nullptr;
  nm_summary sum2 = sum;
  const int64_t n_lane = n_rows - sum.n_deep;
  if (!h->no_class_sort && n_lane > 0) {
    const int cmin = nm_lane_class(NM_LANE_TIER_MAX - sum.max_lane_slack), cmax = nm_lane_class(sum.max_lane_n);
    const int gmin = cmin <= 64 ? 0 : cmin <= NM_LANE_FINE_MAX ? 1 : 2, gmax = cmax <= 64 ? 0 : cmax <= NM_LANE_FINE_MAX ? 1 : 2;
    if (gmin != gmax) {
      if ((rc = nm_reserve(h, &h->d_perm[0], sizeof(int32_t) * (size_t)n_rows)) != NM_OK) return rc;
      if ((rc = nm_reserve(h, &h->d_perm[1], sizeof(int32_t) * (size_t)n_rows)) != NM_OK) return rc;
      if ((rc = nm_reserve(h, &h->d_class_scratch, nm_group_sort_scratch_bytes(n_rows))) != NM_OK) return rc;
      int launches = 0;
      cudaError_t e = (cudaError_t)nm_group_keys_run(tb->n0, tb->n1, n_rows, (int32_t*)h->d_perm[0].p, h->d_class_scratch.p,
                                                     h->d_sum, &launches, st);
      if (e != cudaSuccess) return nm_fail(h, NM_ERR_CUDA, "group binning failed: %s", cudaGetErrorString(e));
      NM_CUDA(h, cudaMemcpyAsync(h->h_sum, h->d_sum, sizeof(nm_summary), cudaMemcpyDeviceToHost, st));
      NM_CUDA(h, cudaStreamSynchronize(st));
      sum2 = *h->h_sum;
      const int64_t g0 = sum2.n_le64, g1 = (int64_t)sum2.n_le104 - sum2.n_le64, g2 = n_lane - sum2.n_le104;
      const int64_t biggest = g0 > g1 ? (g0 > g2 ? g0 : g2) : (g1 > g2 ? g1 : g2);
      if (biggest * 8 >= n_lane * 7 && biggest != (gmax == 2 ? g2 : gmax == 1 ? g1 : g0)) {
        e = (cudaError_t)nm_group_sort_run(n_rows, (int32_t*)h->d_perm[0].p, (int32_t*)h->d_perm[1].p, &perm,
                                           h->d_class_scratch.p, &launches, st);
        if (e != cudaSuccess) return nm_fail(h, NM_ERR_CUDA, "group binning failed: %s", cudaGetErrorString(e));
      }
      h->launches += launches;
    }
  }
  rc = nm_launch_tiers(h, ka, want_u, want_t, want_m, sum2, perm, deep_smem, n_rows, n_pos, st);
  if (rc != NM_OK) return rc;

  // ---- down-sampling branch (myDetect.py:345-361): replaces the KS result of deep positions
  if (ds_on) {
    NM_CUDA(h, cudaMemsetAsync(&h->d_sum->ds_cursor, 0, 2 * sizeof(int), st));  // cursor + too_deep
    const cudaError_t e = (cudaError_t)nm_launch_downsample(ka, pl->pos, pl->seg, pl->seg_cov, prm.ds_times, prm.ds_index,
                                                           prm.ds_seed, &h->d_sum->ds_cursor, &h->d_sum->ds_too_deep,
                                                           h->sm_count, st);
    if (e != cudaSuccess) return nm_fail(h, NM_ERR_CUDA, "nm_downsample_kernel launch failed: %s", cudaGetErrorString(e));
    h->launches++;
    if (sum.ds_too_deep & 1) {  // rows beyond the warp kernel's shared-memory script: one CTA per row
      if ((rc = nm_reserve(h, &h->d_ds_scratch, nm_downsample_deep_scratch_bytes(h->sm_count))) != NM_OK) return rc;
      NM_CUDA(h, cudaMemsetAsync(&h->d_sum->ds_deep_cursor, 0, sizeof(int), st));
      const cudaError_t e2 = (cudaError_t)nm_launch_downsample_deep(ka, pl->pos, pl->seg, pl->seg_cov, prm.ds_times, prm.ds_index,
                                                                  prm.ds_seed, &h->d_sum->ds_deep_cursor, &h->d_sum->ds_too_deep,
                                                                  (float*)h->d_ds_scratch.p, h->sm_count, st);
      if (e2 != cudaSuccess) return nm_fail(h, NM_ERR_CUDA, "nm_downsample_deep_kernel launch failed: %s", cudaGetErrorString(e2));
      h->launches++;
    }
    NM_CUDA(h, cudaMemcpyAsync(h->h_sum, h->d_sum, sizeof(nm_summary), cudaMemcpyDeviceToHost, st));
  }

  // ---- neighbour combination
  if (want_f || want_s) {
    nm_comb_args ca;
    nm_fill_comb_args(&ca, pl, prm, tb, n_rows, ds_on);
    nm_launch_combine(ca, n_rows, st);
    NM_CUDA(h, cudaGetLastError());
    h->launches++;
  }
  NM_CUDA(h, cudaEventRecord(h->ev[4], st));
  if ((rc = nm_fire_armed_head(h, n_rows, n_pos, st)) != NM_OK) return rc;
  NM_CUDA(h, cudaMemcpyAsync(h->h_sum, h->d_sum, sizeof(nm_summary), cudaMemcpyDeviceToHost, st));
  NM_CUDA(h, cudaStreamSynchronize(st));
  h->last_grid_tiles = h->h_sum->grid_tiles;
  if (ds_on && h->h_sum->ds_too_deep)
    return nm_fail(h, NM_ERR_TOO_DEEP, "down-sampling supports at most %d reads per group", NM_DS_DEEP_MAX_READS);
  {
    float ms = 0.f;
    // ev[1] is re-recorded after the plan readback, so [0] covers only the plan kernels' span
    // up to the first record; tiers and combine are bracketed exactly
    if (cudaEventElapsedTime(&ms, h->ev[1], h->ev[2]) == cudaSuccess) h->last_ms[1] = ms;
    if (cudaEventElapsedTime(&ms, h->ev[2], h->ev[3]) == cudaSuccess) h->last_ms[2] = ms;
    if (cudaEventElapsedTime(&ms, h->ev[3], h->ev[4]) == cudaSuccess) h->last_ms[3] = ms;
  }
  return NM_OK;
}

// ------------------------------------------------------------------------------------------
// Asynchronous device entry: a caller that runs detection call after call on the same shape (genome shards,
// slabs, the steps of a benchmark) keeps up to two calls in flight, so that the device never waits for the host
// between them.  A call is issued without a host wait when the handle's previous call had the dense shape
// (nothing filtered, nothing deep, one network class): plan + lane + combine (+ an armed head selection) are
// launched on that assumption, the device validates it, and nm_detect_finish re-runs the call the ordinary way
// if the device refused.  Any other call is simply run to completion here.
// ------------------------------------------------------------------------------------------
extern "C" int nm_detect_device_async(nm_handle* h, const nm_pileup* pl, const nm_params* params, const nm_table* tb,
                                      void* cuda_stream, int* ticket_out) {
  if (!h) return nm_fail(nullptr, NM_ERR_BAD_ARG, "handle is NULL");
  if (!pl || !tb || !ticket_out) return nm_fail(h, NM_ERR_BAD_ARG, "pileup/table/ticket is NULL");
  *ticket_out = -1;
  const int s = h->async_next & 1;
  nm_handle::nm_pending* pd = &h->pend[s];
  if (pd->active) {
    h->head.armed = 0;
    return nm_fail(h, NM_ERR_BAD_ARG, "two asynchronous calls are in flight: nm_detect_finish the older one first");
  }
  const nm_handle::nm_pending* other = &h->pend[s ^ 1];
  if (other->active && !other->done && other->tb.ks_dnum == tb->ks_dnum) {
    h->head.armed = 0;
    return nm_fail(h, NM_ERR_BAD_ARG, "the two calls in flight must write different tables");
  }
  nm_params prm;
  int rc = nm_check_params(h, params, &prm);
  if (rc != NM_OK) { h->head.armed = 0; return rc; }
  const bool ds_on = pl->seg_cov != nullptr && pl->n_seg > 0 && prm.ds_times > 0;
  const bool fast = h->dense_class > 0 && !h->no_dense && !ds_on && pl->vals0 && pl->vals1 && pl->n_pos > 0 &&
                    pl->n_pos <= 0x7fffffffLL - NM_PLAN_PER_BLOCK;
  memset(pd, 0, sizeof(*pd));
  if (!fast) {
    int64_t n_rows = 0;
    rc = nm_detect_device(h, pl, params, tb, &n_rows, cuda_stream);
    if (rc != NM_OK) return rc;
    pd->active = pd->done = 1;
    pd->n_rows = n_rows;
    pd->head_fired = h->head.fired;
    *ticket_out = s;
    h->async_next++;
    return NM_OK;
  }
  h->head.fired = 0;
  h->head.use_cands = 0;
  if ((rc = nm_check_call(h, pl, prm, tb)) != NM_OK) { h->head.armed = 0; return rc; }
  NM_CUDA(h, cudaSetDevice(h->device));
  cudaStream_t st = (cudaStream_t)cuda_stream;
  const int64_t n_pos = pl->n_pos;
  const int nblk = (int)((n_pos + NM_PLAN_PER_BLOCK - 1) / NM_PLAN_PER_BLOCK);
  if ((rc = nm_reserve(h, &h->d_block_count, sizeof(int) * (size_t)nblk)) != NM_OK) return rc;
  if ((rc = nm_reserve(h, &h->d_deep_rows, sizeof(int32_t) * (size_t)n_pos)) != NM_OK) return rc;
  h->ev = h->ev_sets[1 + s];
  h->h_sum = h->h_sums[1 + s];
  h->err[0] = 0;
  do {
    cudaError_t e;
    rc = NM_ERR_CUDA;
    if ((e = cudaEventRecord(h->ev[0], st)) != cudaSuccess) break;
    if ((e = cudaMemsetAsync(h->d_sum, 0, sizeof(nm_summary), st)) != cudaSuccess) break;
    nm_plan_count<<<nblk, NM_PLAN_THREADS, 0, st>>>(pl->off0, pl->off1, n_pos, prm.min_coverage, pl->seg,
                                                    pl->n_seg > 0 ? pl->n_seg : 0, nullptr, (int*)h->d_block_count.p, h->d_sum);
    if ((e = cudaGetLastError()) != cudaSuccess) break;
    h->launches += 1;
    if ((e = cudaEventRecord(h->ev[1], st)) != cudaSuccess) break;
    if ((rc = nm_dense_enqueue(h, pl, prm, tb, h->dense_class, st, &pd->try_grid, &pd->n_launched)) != NM_OK) break;
    rc = NM_ERR_CUDA;
    if ((e = cudaEventRecord(h->ev_done[s], st)) != cudaSuccess) break;
    rc = NM_OK;
  } while (0);
  h->ev = h->ev_sets[0];
  h->h_sum = h->h_sums[0];
  pd->head_fired = h->head.fired;
  h->head.armed = 0;
  h->head.fired = 0;
  if (rc != NM_OK) {
    if (rc == NM_ERR_CUDA && !h->err[0]) nm_fail(h, rc, "asynchronous launch failed: %s", cudaGetErrorString(cudaGetLastError()));
    return rc;
  }
  pd->active = 1;
  pd->pl = *pl; pd->prm = *params; pd->tb = *tb; pd->stream = cuda_stream;
  *ticket_out = s;
  h->async_next++;
  return NM_OK;
}

extern "C" int nm_detect_finish(nm_handle* h, int ticket, int64_t* n_rows_out, int* head_fired_out) {
  if (!h) return nm_fail(nullptr, NM_ERR_BAD_ARG, "handle is NULL");
  if (ticket < 0 || ticket > 1 || !h->pend[ticket].active || !n_rows_out)
    return nm_fail(h, NM_ERR_BAD_ARG, "no asynchronous call behind ticket %d (or n_rows is NULL)", ticket);
  nm_handle::nm_pending* pd = &h->pend[ticket];
  pd->active = 0;
  if (head_fired_out) *head_fired_out = 0;
  if (pd->done) {
    *n_rows_out = pd->n_rows;
    if (head_fired_out) *head_fired_out = pd->head_fired;
    return NM_OK;
  }
  NM_CUDA(h, cudaSetDevice(h->device));
  NM_CUDA(h, cudaEventSynchronize(h->ev_done[ticket]));
  nm_summary sum;
  bool refused = false;
  h->ev = h->ev_sets[1 + ticket];
  h->h_sum = h->h_sums[1 + ticket];
  const int rc = nm_dense_complete(h, pd->try_grid, pd->n_launched, &sum, &refused);
  h->ev = h->ev_sets[0];
  h->h_sum = h->h_sums[0];
  if (rc != NM_OK) return rc;
  if (!refused) {
    h->dense_class = nm_lane_class(sum.max_lane_n);
    h->last_path = 2;
    *n_rows_out = pd->pl.n_pos;
    if (head_fired_out) *head_fired_out = pd->head_fired && !sum.head_fail;
    return NM_OK;
  }
  // the shape was not the previous call's: the ordinary call, from its plan pass (behind whatever else is in
  // flight on the stream); calls issued meanwhile on the same assumption are refused and re-run in their turn
  h->dense_class = 0;
  return nm_detect_device(h, &pd->pl, &pd->prm, &pd->tb, n_rows_out, pd->stream);
}

// ------------------------------------------------------------------------------------------
// Pipelined host entry.  The pileup is cut into slabs of `slab` candidates (+ a halo of nb
// candidates per side, recomputed: the same argument as for genome shards); while slab k is
// tested, slab k+1 is on its way in (copy stream) and the rows of slab k-1 are on their way out
// (second copy stream; PCIe is full duplex).  Host -> device traffic is 96 % of the bytes, so
// the call runs at the H2D line rate with kernels and the D2H hidden behind it; device memory
// holds two slabs instead of the whole pileup.
// ------------------------------------------------------------------------------------------
__global__ void nm_rebase_offsets(int64_t* off, int64_t n, int64_t base) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) off[i] -= base;
}

__global__ void nm_rebase_rows(int32_t* row_pos_index, int64_t lo, int64_t hi, int32_t add) {
  const int64_t i = lo + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < hi) row_pos_index[i] += add;
}

static int nm_detect_host_pipelined(nm_handle* h, const nm_pileup* pl, const nm_params& prm, const nm_table* tb,
                                    int64_t* n_rows_out) {
  const int64_t n = pl->n_pos, S = h->slab;
  const bool i16 = !pl->vals0 && pl->vals0_i16 != nullptr;
  const int nb = (prm.combine != 0) ? prm.nb : 0;
  const int64_t n_slabs = (n + S - 1) / S;
  void* const host_ptrs[17] = {tb->row_pos_index, tb->n0, tb->n1, tb->ks_dnum, tb->ks_d, tb->ks_p,
                               tb->two_u, tb->u_stat, tb->u_p, tb->t_stat, tb->t_p, tb->fisher_stat,
                               tb->fisher_p, tb->stouffer_stat, tb->stouffer_p, tb->flags, tb->moments};
  const size_t elem[17] = {4, 4, 4, 4, 8, 8, 8, 8, 8, 8, 8, 8, 8, 8, 8, 1, 32};
  const bool live[17] = {true, true, true, true, true, true,
                         prm.want_u != 0, prm.want_u != 0, prm.want_u != 0, prm.want_t != 0, prm.want_t != 0,
                         (prm.combine & NM_COMBINE_FISHER) != 0, (prm.combine & NM_COMBINE_FISHER) != 0,
                         (prm.combine & NM_COMBINE_STOUFFER) != 0, (prm.combine & NM_COMBINE_STOUFFER) != 0, true, true};
  int rc;
  // slab geometry and staging capacity
  int64_t max_v0 = 0, max_v1 = 0;
  for (int64_t k = 0; k < n_slabs; ++k) {
    const int64_t lo = k * S, hi = lo + S < n ? lo + S : n;
    const int64_t hlo = lo - nb > 0 ? lo - nb : 0, hhi = hi + nb < n ? hi + nb : n;
    const int64_t v0 = pl->off0[hhi] - pl->off0[hlo], v1 = pl->off1[hhi] - pl->off1[hlo];
    if (v0 < 0 || v1 < 0) return nm_fail(h, NM_ERR_BAD_ARG, "offsets must be non-decreasing");
    max_v0 = v0 > max_v0 ? v0 : max_v0;
    max_v1 = v1 > max_v1 ? v1 : max_v1;
  }
  const int64_t cap = S + 2 * nb;
  for (int s = 0; s < 2; ++s) {
    const size_t need[6] = {sizeof(float) * (size_t)nm_padded_len(max_v0), sizeof(float) * (size_t)nm_padded_len(max_v1),
                            sizeof(int64_t) * (size_t)(cap + 1), sizeof(int64_t) * (size_t)(cap + 1),
                            sizeof(int32_t) * (size_t)cap, sizeof(int32_t) * (size_t)cap};
    for (int b = 0; b < 6; ++b)
      if ((rc = nm_reserve(h, &h->p_in[s][b], need[b])) != NM_OK) return rc;
    if (i16) {
      if ((rc = nm_reserve(h, &h->p_i16[s][0], sizeof(int16_t) * (size_t)(max_v0 + 8))) != NM_OK) return rc;
      if ((rc = nm_reserve(h, &h->p_i16[s][1], sizeof(int16_t) * (size_t)(max_v1 + 8))) != NM_OK) return rc;
    }
    for (int c = 0; c < 17; ++c)
      if (host_ptrs[c] && (rc = nm_reserve(h, &h->p_out[s][c], elem[c] * (size_t)cap)) != NM_OK) return rc;
  }
  const int32_t* d_seg_cov = nullptr;
  if (pl->seg_cov && pl->n_seg > 0) {
    if ((rc = nm_reserve(h, &h->d_seg_cov, sizeof(int32_t) * (size_t)pl->n_seg)) != NM_OK) return rc;
    NM_CUDA(h, cudaMemcpyAsync(h->d_seg_cov.p, pl->seg_cov, sizeof(int32_t) * (size_t)pl->n_seg, cudaMemcpyHostToDevice, h->s_in));
    d_seg_cov = (const int32_t*)h->d_seg_cov.p;
  }
  cudaStream_t st = h->own_stream;

  auto stage_in = [&](int64_t k) -> int {
    const int s = (int)(k & 1);
    const int64_t lo = k * S, hi = lo + S < n ? lo + S : n;
    const int64_t hlo = lo - nb > 0 ? lo - nb : 0, hhi = hi + nb < n ? hi + nb : n;
    const int64_t m = hhi - hlo;
    const int64_t b0 = pl->off0[hlo], b1 = pl->off1[hlo];
    const int64_t v0 = pl->off0[hhi] - b0, v1 = pl->off1[hhi] - b1;
    // the slot's inputs were last read by slab k-2's kernels, which have completed (the compute call is synchronous)
    if (i16) {
      NM_CUDA(h, cudaMemcpyAsync(h->p_i16[s][0].p, pl->vals0_i16 + b0, sizeof(int16_t) * (size_t)v0, cudaMemcpyHostToDevice, h->s_in));
      NM_CUDA(h, cudaMemcpyAsync(h->p_i16[s][1].p, pl->vals1_i16 + b1, sizeof(int16_t) * (size_t)v1, cudaMemcpyHostToDevice, h->s_in));
      int rc2 = nm_expand_i16_run(h, (const int16_t*)h->p_i16[s][0].p, (float*)h->p_in[s][0].p, v0, pl->i16_unit, h->s_in);
      if (rc2 != NM_OK) return rc2;
      rc2 = nm_expand_i16_run(h, (const int16_t*)h->p_i16[s][1].p, (float*)h->p_in[s][1].p, v1, pl->i16_unit, h->s_in);
      if (rc2 != NM_OK) return rc2;
    } else {
      NM_CUDA(h, cudaMemcpyAsync(h->p_in[s][0].p, pl->vals0 + b0, sizeof(float) * (size_t)v0, cudaMemcpyHostToDevice, h->s_in));
      NM_CUDA(h, cudaMemcpyAsync(h->p_in[s][1].p, pl->vals1 + b1, sizeof(float) * (size_t)v1, cudaMemcpyHostToDevice, h->s_in));
    }
    NM_CUDA(h, cudaMemcpyAsync(h->p_in[s][2].p, pl->off0 + hlo, sizeof(int64_t) * (size_t)(m + 1), cudaMemcpyHostToDevice, h->s_in));
    NM_CUDA(h, cudaMemcpyAsync(h->p_in[s][3].p, pl->off1 + hlo, sizeof(int64_t) * (size_t)(m + 1), cudaMemcpyHostToDevice, h->s_in));
    NM_CUDA(h, cudaMemcpyAsync(h->p_in[s][4].p, pl->pos + hlo, sizeof(int32_t) * (size_t)m, cudaMemcpyHostToDevice, h->s_in));
    NM_CUDA(h, cudaMemcpyAsync(h->p_in[s][5].p, pl->seg + hlo, sizeof(int32_t) * (size_t)m, cudaMemcpyHostToDevice, h->s_in));
    const unsigned g = (unsigned)((m + 1 + 255) / 256);
    nm_rebase_offsets<<<g, 256, 0, h->s_in>>>((int64_t*)h->p_in[s][2].p, m + 1, b0);
    nm_rebase_offsets<<<g, 256, 0, h->s_in>>>((int64_t*)h->p_in[s][3].p, m + 1, b1);
    NM_CUDA(h, cudaGetLastError());
    h->launches += 2;
    NM_CUDA(h, cudaEventRecord(h->ev_in[s], h->s_in));
    return NM_OK;
  };

  int64_t row_base = 0;
  double ms_acc[4] = {0, 0, 0, 0};
  if ((rc = stage_in(0)) != NM_OK) return rc;
  for (int64_t k = 0; k < n_slabs; ++k) {
    const int s = (int)(k & 1);
    const int64_t lo = k * S, hi = lo + S < n ? lo + S : n;
    const int64_t hlo = lo - nb > 0 ? lo - nb : 0, hhi = hi + nb < n ? hi + nb : n;
    const int64_t m = hhi - hlo;
    if (k + 1 < n_slabs && (rc = stage_in(k + 1)) != NM_OK) return rc;  // on its way while slab k is tested
    NM_CUDA(h, cudaStreamWaitEvent(st, h->ev_in[s], 0));
    if (k >= 2) NM_CUDA(h, cudaStreamWaitEvent(st, h->ev_out[s], 0));  // the slot's table has left (slab k-2)
    nm_pileup dpl = {(const float*)h->p_in[s][0].p, (const int64_t*)h->p_in[s][2].p, (const float*)h->p_in[s][1].p,
                     (const int64_t*)h->p_in[s][3].p, (const int32_t*)h->p_in[s][4].p, (const int32_t*)h->p_in[s][5].p, m,
                     d_seg_cov, d_seg_cov ? pl->n_seg : 0};
    void* dp[17];
    for (int c = 0; c < 17; ++c) dp[c] = host_ptrs[c] ? h->p_out[s][c].p : nullptr;
    nm_table dtb = {(int32_t*)dp[0], (int32_t*)dp[1], (int32_t*)dp[2], (int32_t*)dp[3], (double*)dp[4], (double*)dp[5],
                    (int64_t*)dp[6], (double*)dp[7], (double*)dp[8], (double*)dp[9], (double*)dp[10], (double*)dp[11],
                    (double*)dp[12], (double*)dp[13], (double*)dp[14], (uint8_t*)dp[15], (double*)dp[16]};
    int64_t rows = 0;
    rc = nm_detect_device(h, &dpl, &prm, &dtb, &rows, (void*)st);
    if (rc != NM_OK) {
      cudaStreamSynchronize(h->s_in);
      cudaStreamSynchronize(h->s_out);
      return rc;
    }
    for (int q = 0; q < 4; ++q) ms_acc[q] += h->last_ms[q];
    // core rows of the slab: those whose candidate lies in [lo, hi)
    int64_t r_lo = lo - hlo, r_hi = hi - hlo;
    if (rows != m) {
      if (rows > 0) {
        int32_t* tmp = (int32_t*)malloc(sizeof(int32_t) * (size_t)rows);
        if (!tmp) return nm_fail(h, NM_ERR_OOM, "out of host memory");
        cudaError_t e = cudaMemcpyAsync(tmp, dp[0], sizeof(int32_t) * (size_t)rows, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) {
          free(tmp);
          return nm_fail(h, NM_ERR_CUDA, "row index readback failed: %s", cudaGetErrorString(e));
        }
        r_lo = std::lower_bound(tmp, tmp + rows, (int32_t)(lo - hlo)) - tmp;
        r_hi = std::lower_bound(tmp, tmp + rows, (int32_t)(hi - hlo)) - tmp;
        free(tmp);
      } else {
        r_lo = r_hi = 0;
      }
    }
    const int64_t core = r_hi - r_lo;
    if (core > 0) {
      if (hlo != 0) {  // row -> candidate index of the whole pileup
        nm_rebase_rows<<<(unsigned)((core + 255) / 256), 256, 0, st>>>((int32_t*)dp[0], r_lo, r_hi, (int32_t)hlo);
        NM_CUDA(h, cudaGetLastError());
        h->launches++;
      }
      NM_CUDA(h, cudaEventRecord(h->ev_cmp[s], st));
      NM_CUDA(h, cudaStreamWaitEvent(h->s_out, h->ev_cmp[s], 0));
      for (int c = 0; c < 17; ++c)
        if (host_ptrs[c] && live[c])
          NM_CUDA(h, cudaMemcpyAsync((unsigned char*)host_ptrs[c] + elem[c] * (size_t)row_base,
                                     (const unsigned char*)dp[c] + elem[c] * (size_t)r_lo, elem[c] * (size_t)core,
                                     cudaMemcpyDeviceToHost, h->s_out));
    }
    NM_CUDA(h, cudaEventRecord(h->ev_out[s], h->s_out));
    row_base += core;
  }
  NM_CUDA(h, cudaStreamSynchronize(h->s_out));
  NM_CUDA(h, cudaStreamSynchronize(h->s_in));
  for (int q = 0; q < 4; ++q) h->last_ms[q] = ms_acc[q];
  *n_rows_out = row_base;
  return NM_OK;
}

extern "C" int nm_detect_host(nm_handle* h, const nm_pileup* pl, const nm_params* params,
                              const nm_table* tb, int64_t* n_rows_out) {
  if (!h) return nm_fail(nullptr, NM_ERR_BAD_ARG, "handle is NULL");
  if (!pl || !tb || !n_rows_out) return nm_fail(h, NM_ERR_BAD_ARG, "pileup/table/n_rows is NULL");
  nm_params prm;
  int rc = nm_check_params(h, params, &prm);
  if (rc != NM_OK) return rc;
  *n_rows_out = 0;
  if (pl->n_pos < 0) return nm_fail(h, NM_ERR_BAD_ARG, "n_pos is negative");
  if (pl->n_pos == 0) return NM_OK;
  const bool i16 = !pl->vals0 && !pl->vals1 && pl->vals0_i16 && pl->vals1_i16;
  if (i16 && !(pl->i16_unit > 0.0)) return nm_fail(h, NM_ERR_BAD_ARG, "i16_unit must be positive");
  if ((!i16 && (!pl->vals0 || !pl->vals1)) || !pl->off0 || !pl->off1 || !pl->pos || !pl->seg)
    return nm_fail(h, NM_ERR_BAD_ARG, "pileup has a NULL array");
  NM_CUDA(h, cudaSetDevice(h->device));
  cudaStream_t st = h->own_stream;
  const int64_t n = pl->n_pos;
  const int64_t nv0 = pl->off0[n], nv1 = pl->off1[n];
  if (nv0 < 0 || nv1 < 0 || pl->off0[0] != 0 || pl->off1[0] != 0)
    return nm_fail(h, NM_ERR_BAD_ARG, "offsets must start at 0 and be non-decreasing");
  if (!tb->row_pos_index || !tb->n0 || !tb->n1 || !tb->ks_dnum || !tb->ks_p)
    return nm_fail(h, NM_ERR_BAD_ARG, "table lacks a mandatory output (row_pos_index,n0,n1,ks_dnum,ks_p)");
  if (h->slab > 0 && n >= 2 * h->slab) return nm_detect_host_pipelined(h, pl, prm, tb, n_rows_out);
  if ((rc = nm_reserve(h, &h->d_vals0, sizeof(float) * (size_t)nm_padded_len(nv0))) != NM_OK) return rc;
  if ((rc = nm_reserve(h, &h->d_vals1, sizeof(float) * (size_t)nm_padded_len(nv1))) != NM_OK) return rc;
  if ((rc = nm_reserve(h, &h->d_off0, sizeof(int64_t) * (size_t)(n + 1))) != NM_OK) return rc;
  if ((rc = nm_reserve(h, &h->d_off1, sizeof(int64_t) * (size_t)(n + 1))) != NM_OK) return rc;
  if ((rc = nm_reserve(h, &h->d_pos, sizeof(int32_t) * (size_t)n)) != NM_OK) return rc;
  if ((rc = nm_reserve(h, &h->d_seg, sizeof(int32_t) * (size_t)n)) != NM_OK) return rc;
  if (i16) {
    if ((rc = nm_reserve(h, &h->p_i16[0][0], sizeof(int16_t) * (size_t)(nv0 + 8))) != NM_OK) return rc;
    if ((rc = nm_reserve(h, &h->p_i16[0][1], sizeof(int16_t) * (size_t)(nv1 + 8))) != NM_OK) return rc;
    NM_CUDA(h, cudaMemcpyAsync(h->p_i16[0][0].p, pl->vals0_i16, sizeof(int16_t) * (size_t)nv0, cudaMemcpyHostToDevice, st));
    NM_CUDA(h, cudaMemcpyAsync(h->p_i16[0][1].p, pl->vals1_i16, sizeof(int16_t) * (size_t)nv1, cudaMemcpyHostToDevice, st));
    if ((rc = nm_expand_i16_run(h, (const int16_t*)h->p_i16[0][0].p, (float*)h->d_vals0.p, nv0, pl->i16_unit, st)) != NM_OK) return rc;
    if ((rc = nm_expand_i16_run(h, (const int16_t*)h->p_i16[0][1].p, (float*)h->d_vals1.p, nv1, pl->i16_unit, st)) != NM_OK) return rc;
  } else {
    NM_CUDA(h, cudaMemcpyAsync(h->d_vals0.p, pl->vals0, sizeof(float) * (size_t)nv0, cudaMemcpyHostToDevice, st));
    NM_CUDA(h, cudaMemcpyAsync(h->d_vals1.p, pl->vals1, sizeof(float) * (size_t)nv1, cudaMemcpyHostToDevice, st));
  }
  NM_CUDA(h, cudaMemcpyAsync(h->d_off0.p, pl->off0, sizeof(int64_t) * (size_t)(n + 1), cudaMemcpyHostToDevice, st));
  NM_CUDA(h, cudaMemcpyAsync(h->d_off1.p, pl->off1, sizeof(int64_t) * (size_t)(n + 1), cudaMemcpyHostToDevice, st));
  NM_CUDA(h, cudaMemcpyAsync(h->d_pos.p, pl->pos, sizeof(int32_t) * (size_t)n, cudaMemcpyHostToDevice, st));
  NM_CUDA(h, cudaMemcpyAsync(h->d_seg.p, pl->seg, sizeof(int32_t) * (size_t)n, cudaMemcpyHostToDevice, st));

  // device-side table mirrors exactly the outputs the caller asked for
  void* const host_ptrs[17] = {tb->row_pos_index, tb->n0, tb->n1, tb->ks_dnum, tb->ks_d, tb->ks_p,
                               tb->two_u, tb->u_stat, tb->u_p, tb->t_stat, tb->t_p, tb->fisher_stat,
                               tb->fisher_p, tb->stouffer_stat, tb->stouffer_p, tb->flags, tb->moments};
  const size_t elem[17] = {4, 4, 4, 4, 8, 8, 8, 8, 8, 8, 8, 8, 8, 8, 8, 1, 32};
  void* dev_ptrs[17];
  for (int k = 0; k < 17; ++k) {
    dev_ptrs[k] = nullptr;
    if (host_ptrs[k]) {
      if ((rc = nm_reserve(h, &h->d_out[k], elem[k] * (size_t)n)) != NM_OK) return rc;
      dev_ptrs[k] = h->d_out[k].p;
    }
  }
  const int32_t* d_seg_cov = nullptr;
  if (pl->seg_cov && pl->n_seg > 0) {
    if ((rc = nm_reserve(h, &h->d_seg_cov, sizeof(int32_t) * (size_t)pl->n_seg)) != NM_OK) return rc;
    NM_CUDA(h, cudaMemcpyAsync(h->d_seg_cov.p, pl->seg_cov, sizeof(int32_t) * (size_t)pl->n_seg, cudaMemcpyHostToDevice, st));
    d_seg_cov = (const int32_t*)h->d_seg_cov.p;
  }
  nm_pileup dpl = {(const float*)h->d_vals0.p, (const int64_t*)h->d_off0.p, (const float*)h->d_vals1.p,
                   (const int64_t*)h->d_off1.p, (const int32_t*)h->d_pos.p, (const int32_t*)h->d_seg.p, n,
                   d_seg_cov, d_seg_cov ? pl->n_seg : 0};
  nm_table dtb = {(int32_t*)dev_ptrs[0], (int32_t*)dev_ptrs[1], (int32_t*)dev_ptrs[2], (int32_t*)dev_ptrs[3],
                  (double*)dev_ptrs[4], (double*)dev_ptrs[5], (int64_t*)dev_ptrs[6], (double*)dev_ptrs[7],
                  (double*)dev_ptrs[8], (double*)dev_ptrs[9], (double*)dev_ptrs[10], (double*)dev_ptrs[11],
                  (double*)dev_ptrs[12], (double*)dev_ptrs[13], (double*)dev_ptrs[14], (uint8_t*)dev_ptrs[15],
                  (double*)dev_ptrs[16]};
  int64_t n_rows = 0;
  rc = nm_detect_device(h, &dpl, &prm, &dtb, &n_rows, (void*)st);
  if (rc != NM_OK) return rc;
  const bool live[17] = {true, true, true, true, true, true,
                         prm.want_u != 0, prm.want_u != 0, prm.want_u != 0, prm.want_t != 0, prm.want_t != 0,
                         (prm.combine & NM_COMBINE_FISHER) != 0, (prm.combine & NM_COMBINE_FISHER) != 0,
                         (prm.combine & NM_COMBINE_STOUFFER) != 0, (prm.combine & NM_COMBINE_STOUFFER) != 0, true, true};
  for (int k = 0; k < 17; ++k)
    if (host_ptrs[k] && live[k] && n_rows > 0)
      NM_CUDA(h, cudaMemcpyAsync(host_ptrs[k], dev_ptrs[k], elem[k] * (size_t)n_rows, cudaMemcpyDeviceToHost, st));
  NM_CUDA(h, cudaStreamSynchronize(st));
  *n_rows_out = n_rows;
  return NM_OK;
}


// ------------------------------------------------------------------------------------------
// ranking (myDetect.py:459-461) -- see nm_rank.cu
// ------------------------------------------------------------------------------------------
extern "C" int nm_rank_device(nm_handle* h, const double* key_comb, const double* key_ks, const double* key_u,
                              int64_t n_rows, int reverse, int32_t* order, void* cuda_stream) {
  if (!h) return nm_fail(nullptr, NM_ERR_BAD_ARG, "handle is NULL");
  if (n_rows < 0 || n_rows > 0x7fffffffLL) return nm_fail(h, NM_ERR_BAD_ARG, "n_rows (%lld) out of range", (long long)n_rows);
  if (n_rows == 0) return NM_OK;
  if (!key_ks || !order) return nm_fail(h, NM_ERR_BAD_ARG, "key_ks/order is NULL");
  NM_CUDA(h, cudaSetDevice(h->device));
  cudaStream_t st = (cudaStream_t)cuda_stream;
  int rc = nm_reserve(h, &h->d_rank, nm_rank_scratch_bytes(n_rows));
  if (rc != NM_OK) return rc;
  int launches = 0;
  const cudaError_t e = (cudaError_t)nm_rank_run(key_comb, key_ks, key_u, n_rows, reverse, order, h->d_rank.p, &launches, st);
  h->launches += launches;
  if (e != cudaSuccess) return nm_fail(h, NM_ERR_CUDA, "ranking failed: %s", cudaGetErrorString(e));
  NM_CUDA(h, cudaStreamSynchronize(st));
  return NM_OK;
}

extern "C" int nm_rank_host(nm_handle* h, const double* key_comb, const double* key_ks, const double* key_u,
                            int64_t n_rows, int reverse, int32_t* order) {
  if (!h) return nm_fail(nullptr, NM_ERR_BAD_ARG, "handle is NULL");
  if (n_rows < 0 || n_rows > 0x7fffffffLL) return nm_fail(h, NM_ERR_BAD_ARG, "n_rows (%lld) out of range", (long long)n_rows);
  if (n_rows == 0) return NM_OK;
  if (!key_ks || !order) return nm_fail(h, NM_ERR_BAD_ARG, "key_ks/order is NULL");
  NM_CUDA(h, cudaSetDevice(h->device));
  cudaStream_t st = h->own_stream;
  const double* host_keys[3] = {key_comb, key_ks, key_u};
  const double* dev_keys[3] = {nullptr, nullptr, nullptr};
  int rc;
  for (int k = 0; k < 3; ++k) {
    if (!host_keys[k]) continue;
    if ((rc = nm_reserve(h, &h->d_rank_keys[k], sizeof(double) * (size_t)n_rows)) != NM_OK) return rc;
    NM_CUDA(h, cudaMemcpyAsync(h->d_rank_keys[k].p, host_keys[k], sizeof(double) * (size_t)n_rows, cudaMemcpyHostToDevice, st));
    dev_keys[k] = (const double*)h->d_rank_keys[k].p;
  }
  if ((rc = nm_reserve(h, &h->d_rank_order, sizeof(int32_t) * (size_t)n_rows)) != NM_OK) return rc;
  rc = nm_rank_device(h, dev_keys[0], dev_keys[1], dev_keys[2], n_rows, reverse, (int32_t*)h->d_rank_order.p, (void*)st);
  if (rc != NM_OK) return rc;
  NM_CUDA(h, cudaMemcpyAsync(order, h->d_rank_order.p, sizeof(int32_t) * (size_t)n_rows, cudaMemcpyDeviceToHost, st));
  NM_CUDA(h, cudaStreamSynchronize(st));
  return NM_OK;
}


// Head of the ranking: the first rows of what nm_rank_device would return, found with three
// streaming passes over the primary key instead of a full sort (nm_rank.cu); the selected
// records (a few thousand) are ordered on the host.
extern "C" int nm_rank_head_device(nm_handle* h, const double* key_comb, const double* key_ks, const double* key_u,
                                   int64_t n_rows, int reverse, int64_t want, const nm_head_geometry* geometry,
                                   nm_head_row* rows_out, int64_t cap, int64_t* n_head, void* cuda_stream) {
  if (!h) return nm_fail(nullptr, NM_ERR_BAD_ARG, "handle is NULL");
  if (!n_head || !rows_out || cap <= 0) return nm_fail(h, NM_ERR_BAD_ARG, "rows_out/n_head is NULL or cap <= 0");
  *n_head = 0;
  if (n_rows < 0 || n_rows > 0x7fffffffLL) return nm_fail(h, NM_ERR_BAD_ARG, "n_rows (%lld) out of range", (long long)n_rows);
  if (n_rows == 0 || want <= 0) return NM_OK;
  if (!key_comb && !key_ks && !key_u) return nm_fail(h, NM_ERR_BAD_ARG, "no ranking key given");
  nm_head_geo geo;
  memset(&geo, 0, sizeof(geo));
  if (geometry) {
    if (!geometry->pos || !geometry->seg || geometry->row_offset < 0 || geometry->nearby < 0 ||
        geometry->row_offset + n_rows > geometry->n_rows_total)
      return nm_fail(h, NM_ERR_BAD_ARG, "head geometry: NULL pos/seg or a row range outside the row list");
    geo.row_pos_index = geometry->row_pos_index; geo.pos = geometry->pos; geo.seg = geometry->seg;
    geo.row_offset = geometry->row_offset; geo.n_rows_total = geometry->n_rows_total; geo.nearby = geometry->nearby;
  }
  NM_CUDA(h, cudaSetDevice(h->device));
  cudaStream_t st = (cudaStream_t)cuda_stream;
  int64_t dev_cap = cap < 16384 ? 16384 : cap;
  const size_t head_off = (sizeof(unsigned) * (4096 + 4) + 255) & ~(size_t)255;
  // pinned landing area: [4 info words][first records]; one stream sync covers both copies
  const int64_t first = 8192;
  if (!h->h_head) {
    NM_CUDA(h, cudaMallocHost(&h->h_head, 64 + sizeof(nm_head_record) * (size_t)first));
  }
  unsigned* info = (unsigned*)h->h_head;
  nm_head_record* first_recs = (nm_head_record*)((unsigned char*)h->h_head + 64);
  for (int attempt = 0; attempt < 2; ++attempt) {
    int rc = nm_reserve(h, &h->d_rank, nm_head_scratch_bytes(dev_cap));
    if (rc != NM_OK) return rc;
    int launches = 0;
    const cudaError_t e = (cudaError_t)nm_head_run(key_comb, key_ks, key_u, n_rows, reverse, want, dev_cap, geo, h->d_rank.p,
                                                   nullptr, h->sm_count, &launches, st);
    h->launches += launches;
    if (e != cudaSuccess) return nm_fail(h, NM_ERR_CUDA, "head selection failed: %s", cudaGetErrorString(e));
    NM_CUDA(h, cudaMemcpyAsync(info, (const unsigned*)h->d_rank.p + 4096, 4 * sizeof(unsigned), cudaMemcpyDeviceToHost, st));
    NM_CUDA(h, cudaMemcpyAsync(first_recs, (const unsigned char*)h->d_rank.p + head_off,
                               sizeof(nm_head_record) * (size_t)(first < dev_cap ? first : dev_cap), cudaMemcpyDeviceToHost, st));
    NM_CUDA(h, cudaStreamSynchronize(st));
    if ((int64_t)info[1] <= dev_cap) break;
    dev_cap = (int64_t)info[1];  // the cut bin holds more rows than fit: once more with room for all of them
  }
  const int64_t n_sel = (int64_t)info[1];
  if (n_sel > dev_cap) return nm_fail(h, NM_ERR_CUDA, "internal: head selection overflow");
  if (n_sel > cap)
    return nm_fail(h, NM_ERR_BAD_ARG, "the ranking head holds %lld rows (ties at the cut), rows_out has room for %lld",
                   (long long)n_sel, (long long)cap);
  nm_head_record* recs = first_recs;
  nm_head_record* heap = nullptr;
  if (n_sel > first) {
    heap = (nm_head_record*)malloc(sizeof(nm_head_record) * (size_t)n_sel);
    if (!heap) return nm_fail(h, NM_ERR_OOM, "out of host memory");
    cudaError_t e = cudaMemcpyAsync(heap, (const unsigned char*)h->d_rank.p + head_off, sizeof(nm_head_record) * (size_t)n_sel,
                                    cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) {
      free(heap);
      return nm_fail(h, NM_ERR_CUDA, "head readback failed: %s", cudaGetErrorString(e));
    }
    recs = heap;
  }
  struct Less {
    int reverse;
    bool operator()(const nm_head_record& a, const nm_head_record& b) const {
      for (int k = 0; k < 3; ++k)
        if (a.key[k] != b.key[k]) return a.key[k] < b.key[k];
      return reverse ? a.row > b.row : a.row < b.row;  // the stable order, reversed for rankUse='st'
    }
  };
  std::sort(recs, recs + n_sel, Less{reverse});
  for (int64_t i = 0; i < n_sel; ++i) {
    rows_out[i].row = recs[i].row;
    rows_out[i].seg = recs[i].seg;
    rows_out[i].pos = recs[i].pos;
    rows_out[i].full_nbhd = recs[i].full_nbhd;
    rows_out[i].reserved = 0;
    for (int k = 0; k < 3; ++k) rows_out[i].key[k] = recs[i].key[k];
  }
  if (heap) free(heap);
  *n_head = n_sel;
  return NM_OK;
}

// The same selection left ON THE DEVICE, unsorted, in the caller's record buffer: what a rank hands
// to the all-gather of a sharded run.  Nothing is waited for -- three streaming kernels and a header.
extern "C" int nm_rank_head_select_device(nm_handle* h, const double* key_comb, const double* key_ks, const double* key_u,
                                          int64_t n_rows, int reverse, int64_t want, const nm_head_geometry* geometry,
                                          nm_head_row* records_dev, int64_t cap, void* cuda_stream) {
  if (!h) return nm_fail(nullptr, NM_ERR_BAD_ARG, "handle is NULL");
  nm_head_peers_dev peers = h->next_peers;  // one shot, whatever becomes of this call
  peers.refused = nullptr;
  h->next_peers.n = 0;
  if (!records_dev || cap <= 0) return nm_fail(h, NM_ERR_BAD_ARG, "records_dev is NULL or cap <= 0");
  if (n_rows < 0 || n_rows > 0x7fffffffLL) return nm_fail(h, NM_ERR_BAD_ARG, "n_rows (%lld) out of range", (long long)n_rows);
  if (!key_comb && !key_ks && !key_u) return nm_fail(h, NM_ERR_BAD_ARG, "no ranking key given");
  nm_head_geo geo;
  memset(&geo, 0, sizeof(geo));
  if (geometry) {
    if (!geometry->pos || !geometry->seg || geometry->row_offset < 0 || geometry->nearby < 0 ||
        geometry->row_offset + n_rows > geometry->n_rows_total)
      return nm_fail(h, NM_ERR_BAD_ARG, "head geometry: NULL pos/seg or a row range outside the row list");
    geo.row_pos_index = geometry->row_pos_index; geo.pos = geometry->pos; geo.seg = geometry->seg;
    geo.row_offset = geometry->row_offset; geo.n_rows_total = geometry->n_rows_total; geo.nearby = geometry->nearby;
  }
  NM_CUDA(h, cudaSetDevice(h->device));
  cudaStream_t st = (cudaStream_t)cuda_stream;
  int rc = nm_reserve(h, &h->d_rank, nm_head_scratch_bytes(1));
  if (rc != NM_OK) return rc;
  if (n_rows == 0) {
    NM_CUDA(h, cudaMemsetAsync(records_dev, 0, sizeof(nm_head_row), st));
    return NM_OK;
  }
  int launches = 0;
  static_assert(sizeof(nm_head_record) == sizeof(nm_head_row), "device and ABI head records must have one layout");
  const cudaError_t e = (cudaError_t)nm_head_run(key_comb, key_ks, key_u, n_rows, reverse, want > 0 ? want : 1, cap, geo,
                                                 h->d_rank.p, (nm_head_record*)records_dev, h->sm_count, &launches, st,
                                                 peers.n > 0 ? &peers : nullptr);
  h->launches += launches;
  if (e != cudaSuccess) return nm_fail(h, NM_ERR_CUDA, "head selection failed: %s", cudaGetErrorString(e));
  return NM_OK;
}

// Arm the same selection for the NEXT nm_detect_device call on this handle: that call launches it on its own
// stream right behind its last kernel and before its host wait, provided the call's rows turn out to be its
// candidates (nothing filtered) -- the key columns and the geometry are given for that case.  One shot:
// nm_head_fired tells whether it ran; the arming is dropped when the call returns either way.
extern "C" int nm_arm_head_select(nm_handle* h, const double* key_comb, const double* key_ks, const double* key_u,
                                  int64_t n_rows, int reverse, int64_t want, const nm_head_geometry* geometry,
                                  nm_head_row* records_dev, int64_t cap) {
  if (!h) return nm_fail(nullptr, NM_ERR_BAD_ARG, "handle is NULL");
  h->head.armed = h->head.fired = 0;
  const nm_head_peers_dev peers_now = h->next_peers;  // one shot, whatever becomes of this call
  h->next_peers.n = 0;
  if (!records_dev || cap <= 0) return nm_fail(h, NM_ERR_BAD_ARG, "records_dev is NULL or cap <= 0");
  if (n_rows <= 0 || n_rows > 0x7fffffffLL) return nm_fail(h, NM_ERR_BAD_ARG, "n_rows (%lld) out of range", (long long)n_rows);
  if (!key_comb && !key_ks && !key_u) return nm_fail(h, NM_ERR_BAD_ARG, "no ranking key given");
  memset(&h->head.geo, 0, sizeof(h->head.geo));
  if (geometry) {
    if (!geometry->pos || !geometry->seg || geometry->row_offset < 0 || geometry->nearby < 0 ||
        geometry->row_offset + n_rows > geometry->n_rows_total)
      return nm_fail(h, NM_ERR_BAD_ARG, "head geometry: NULL pos/seg or a row range outside the row list");
    h->head.geo.row_pos_index = geometry->row_pos_index; h->head.geo.pos = geometry->pos; h->head.geo.seg = geometry->seg;
    h->head.geo.row_offset = geometry->row_offset; h->head.geo.n_rows_total = geometry->n_rows_total;
    h->head.geo.nearby = geometry->nearby;
  }
  h->head.key[0] = key_comb; h->head.key[1] = key_ks; h->head.key[2] = key_u;
  h->head.n_rows = n_rows; h->head.reverse = reverse; h->head.want = want > 0 ? want : 1; h->head.cap = cap;
  h->head.records = (nm_head_record*)records_dev;
  h->head.peers = peers_now;
  h->head.peers.refused = &h->d_sum->dense_retry;
  h->head.armed = 1;
  return NM_OK;
}

// Peers of the next head selection (armed or direct; one shot): the selection kernels store the header and the
// records into every base[p] as well -- the all-gather of a sharded run's heads, fused into their selection.
extern "C" int nm_head_set_peers(nm_handle* h, const nm_head_peers* peers) {
  if (!h) return nm_fail(nullptr, NM_ERR_BAD_ARG, "handle is NULL");
  memset(&h->next_peers, 0, sizeof(h->next_peers));
  if (!peers) return NM_OK;
  if (peers->n_peers < 0 || peers->n_peers > NM_MAX_PEERS)
    return nm_fail(h, NM_ERR_BAD_ARG, "n_peers (%d) outside [0, %d]", peers->n_peers, NM_MAX_PEERS);
  for (int p = 0; p < peers->n_peers; ++p) {
    if (!peers->base[p]) return nm_fail(h, NM_ERR_BAD_ARG, "peer %d has a NULL buffer", p);
    h->next_peers.base[p] = (nm_head_record*)peers->base[p];
  }
  h->next_peers.n = peers->n_peers;
  h->next_peers.epoch = peers->epoch;
  return NM_OK;
}

// Peer-visible device buffers (CUDA IPC): what nm_head_set_peers points at when the peers are other processes.
extern "C" int nm_peer_alloc(nm_handle* h, int64_t bytes, void** dev_ptr_out, unsigned char* ipc_handle_out) {
  if (!h) return nm_fail(nullptr, NM_ERR_BAD_ARG, "handle is NULL");
  if (bytes <= 0 || !dev_ptr_out || !ipc_handle_out) return nm_fail(h, NM_ERR_BAD_ARG, "nm_peer_alloc: bad argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == NM_IPC_HANDLE_BYTES, "IPC handle size");
  NM_CUDA(h, cudaSetDevice(h->device));
  void* p = nullptr;
  NM_CUDA(h, cudaMalloc(&p, (size_t)bytes));
  cudaError_t e = cudaMemset(p, 0, (size_t)bytes);
  cudaIpcMemHandle_t hd;
  if (e == cudaSuccess) e = cudaIpcGetMemHandle(&hd, p);
  if (e != cudaSuccess) {
    cudaFree(p);
    return nm_fail(h, NM_ERR_CUDA, "nm_peer_alloc: %s", cudaGetErrorString(e));
  }
  memcpy(ipc_handle_out, &hd, sizeof(hd));
  *dev_ptr_out = p;
  return NM_OK;
}
extern "C" int nm_peer_open(nm_handle* h, const unsigned char* ipc_handle, void** dev_ptr_out) {
  if (!h) return nm_fail(nullptr, NM_ERR_BAD_ARG, "handle is NULL");
  if (!ipc_handle || !dev_ptr_out) return nm_fail(h, NM_ERR_BAD_ARG, "nm_peer_open: bad argument");
  NM_CUDA(h, cudaSetDevice(h->device));
  cudaIpcMemHandle_t hd;
  memcpy(&hd, ipc_handle, sizeof(hd));
  void* p = nullptr;
  const cudaError_t e = cudaIpcOpenMemHandle(&p, hd, cudaIpcMemLazyEnablePeerAccess);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return nm_fail(h, NM_ERR_CUDA, "nm_peer_open: %s", cudaGetErrorString(e));
  }
  *dev_ptr_out = p;
  return NM_OK;
}
extern "C" int nm_peer_close(nm_handle* h, void* dev_ptr) {
  if (!h) return nm_fail(nullptr, NM_ERR_BAD_ARG, "handle is NULL");
  if (!dev_ptr) return NM_OK;
  NM_CUDA(h, cudaSetDevice(h->device));
  NM_CUDA(h, cudaIpcCloseMemHandle(dev_ptr));
  return NM_OK;
}
extern "C" int nm_peer_free(nm_handle* h, void* dev_ptr) {
  if (!h) return nm_fail(nullptr, NM_ERR_BAD_ARG, "handle is NULL");
  if (!dev_ptr) return NM_OK;
  NM_CUDA(h, cudaSetDevice(h->device));
  NM_CUDA(h, cudaFree(dev_ptr));
  return NM_OK;
}
extern "C" int nm_head_fired(const nm_handle* h) { return h ? h->head.fired : 0; }

// ------------------------------------------------------------------------------------------
// result records for the multi-GPU gather: 28 bytes per row, packed
// ------------------------------------------------------------------------------------------
#define NM_PACK_ROWS 256
__global__ void __launch_bounds__(NM_PACK_ROWS)
nm_pack_kernel(const int32_t* __restrict__ dnum, const double* __restrict__ ks_p, const double* __restrict__ c_stat,
               const double* __restrict__ c_p, int64_t row_lo, int64_t n, unsigned char* __restrict__ out) {
  __shared__ __align__(16) unsigned char rec[NM_PACK_ROWS * NM_RECORD_BYTES];
  const int64_t base = (int64_t)blockIdx.x * NM_PACK_ROWS;
  const int64_t i = base + threadIdx.x;
  if (i < n) {
    const int64_t r = row_lo + i;
    // 28-byte records are only 4-byte aligned: the doubles go in as two words each
    uint32_t* w = reinterpret_cast<uint32_t*>(rec + threadIdx.x * NM_RECORD_BYTES);
    const double v[3] = {ks_p[r], c_stat[r], c_p[r]};
    w[0] = (uint32_t)dnum[r];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const unsigned long long b = (unsigned long long)__double_as_longlong(v[k]);
      w[1 + 2 * k] = (uint32_t)b;
      w[2 + 2 * k] = (uint32_t)(b >> 32);
    }
  }
  __syncthreads();
  const int64_t rows_here = n - base < NM_PACK_ROWS ? n - base : NM_PACK_ROWS;
  const int words = (int)(rows_here * NM_RECORD_BYTES / 4);
  uint32_t* dst = reinterpret_cast<uint32_t*>(out + base * NM_RECORD_BYTES);
  const uint32_t* src = reinterpret_cast<const uint32_t*>(rec);
  for (int k = threadIdx.x; k < words; k += NM_PACK_ROWS) dst[k] = src[k];
}

extern "C" int nm_pack_records_device(nm_handle* h, const nm_table* tb, int64_t row_lo, int64_t n, int which_combine,
                                      void* records, void* cuda_stream) {
  if (!h) return nm_fail(nullptr, NM_ERR_BAD_ARG, "handle is NULL");
  if (!tb || !records || row_lo < 0 || n < 0) return nm_fail(h, NM_ERR_BAD_ARG, "table/records is NULL or a negative range");
  const double* cs = which_combine == NM_COMBINE_FISHER ? tb->fisher_stat : tb->stouffer_stat;
  const double* cp = which_combine == NM_COMBINE_FISHER ? tb->fisher_p : tb->stouffer_p;
  if (which_combine != NM_COMBINE_FISHER && which_combine != NM_COMBINE_STOUFFER)
    return nm_fail(h, NM_ERR_BAD_PARAM, "which_combine must be NM_COMBINE_FISHER or NM_COMBINE_STOUFFER");
  if (!tb->ks_dnum || !tb->ks_p || !cs || !cp) return nm_fail(h, NM_ERR_BAD_ARG, "table lacks a column of the record");
  if (n == 0) return NM_OK;
  NM_CUDA(h, cudaSetDevice(h->device));
  nm_pack_kernel<<<(unsigned)((n + NM_PACK_ROWS - 1) / NM_PACK_ROWS), NM_PACK_ROWS, 0, (cudaStream_t)cuda_stream>>>(
      tb->ks_dnum, tb->ks_p, cs, cp, row_lo, n, (unsigned char*)records);
  NM_CUDA(h, cudaGetLastError());
  h->launches++;
  return NM_OK;
}
