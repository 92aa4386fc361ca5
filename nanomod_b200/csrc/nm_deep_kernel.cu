// nm_deep_kernel.cu -- the deep tier: one CTA (256 threads) per position for rows with more than
// 128 reads in a group (getKStest, bin/scripts/myDetect.py:327-343, same statistics).
//
//   stage   the two contiguous pileup slices arrive by TMA bulk copy (cp.async.bulk + mbarrier)
//   sort    each group as a power-of-two array P = E*256 (>= 512), E elements per thread held in
//           REGISTERS: a per-thread network sorts the E-chunk, then a normalised bitonic sort
//           (all comparators ascending: first step of every merge compares i with i ^ (k-1))
//           runs its small strides in registers, strides that stay inside a warp through
//           shuffles, and only the few strides that cross warps through shared memory
//           (6 of 66 stages at P = 2048)
//   ranks   KS only: every thread takes a contiguous piece of the pooled order (merge-path
//           split by binary search) and walks it, evaluating c0*n1 - c1*n0 at tie-group ends;
//           with the rank statistics: per-element counts by binary search (nm_deep.cuh)
//   tails   thread 0, fp64
//   grid keys  when every value of the position is a three-place decimal (nm_lane.cuh "Grid keys": checked
//           value by value, block-wide) the two groups are packed into ONE array of 16-bit key pairs and
//           sorted by one pass of the same network with packed min / max: half the sort.  Any other
//           position takes the float32 sorts.
#include "nm_device.cuh"

template <class T>
__device__ __forceinline__ void nm_ce_up(T& a, T& b) {
  const T lo = nm_min(a, b), hi = nm_max(a, b);
  a = lo;
  b = hi;
}
__device__ __forceinline__ float nm_shfl_xor(float v, int m) { return __shfl_xor_sync(0xffffffffu, v, m); }
__device__ __forceinline__ nm_p16 nm_shfl_xor(nm_p16 v, int m) {
  nm_p16 r;
  r.v = __shfl_xor_sync(0xffffffffu, v.v, m);
  return r;
}
template <class T>
__device__ __forceinline__ T nm_deep_pinf();
template <>
__device__ __forceinline__ float nm_deep_pinf<float>() { return NM_INF; }
template <>
__device__ __forceinline__ nm_p16 nm_deep_pinf<nm_p16>() {
  nm_p16 r;
  r.v = NM_GRID_PINF;
  return r;
}

// exchange step of the normalised bitonic network across threads: partner thread tid ^ M,
// partner element i (same) or E-1-i (reversed: the "flip" first step of a merge)
template <int E, class T>
__device__ __forceinline__ void nm_deep_exchange(T (&x)[E], T* stage, int tid, int M, bool reversed) {
  int hb = M;  // highest set bit of M decides who keeps the minimum
  hb |= hb >> 1; hb |= hb >> 2; hb |= hb >> 4; hb |= hb >> 8;
  hb = (hb + 1) >> 1;
  const bool lower = (tid & hb) == 0;
  if (M < 32) {
    if (reversed) {
#pragma unroll
      for (int i = 0; i < (E + 1) / 2; ++i) {
        const int r = E - 1 - i;
        const T o1 = nm_shfl_xor(x[r], M);  // partner's counterpart of my x[i]
        const T o2 = nm_shfl_xor(x[i], M);  // partner's counterpart of my x[r]
        x[i] = lower ? nm_min(x[i], o1) : nm_max(x[i], o1);
        if (r != i) x[r] = lower ? nm_min(x[r], o2) : nm_max(x[r], o2);
      }
    } else {
#pragma unroll
      for (int i = 0; i < E; ++i) {
        const T o = nm_shfl_xor(x[i], M);
        x[i] = lower ? nm_min(x[i], o) : nm_max(x[i], o);
      }
    }
  } else {
    // through shared memory, transposed staging (element i of thread t at stage[i*256 + t]):
    // conflict-free for both the writes and the permuted reads
    __syncthreads();
#pragma unroll
    for (int i = 0; i < E; ++i) stage[i * NM_DEEP_THREADS + tid] = x[i];
    __syncthreads();
    const int pt = tid ^ M;
#pragma unroll
    for (int i = 0; i < E; ++i) {
      const T o = stage[(reversed ? E - 1 - i : i) * NM_DEEP_THREADS + pt];
      x[i] = lower ? nm_min(x[i], o) : nm_max(x[i], o);
    }
  }
}

// sort s[0 .. E*256) ascending in place (s is also used as the staging area)
template <int E, class T>
__device__ __forceinline__ void nm_deep_sort(T* s, int tid) {
  T x[E];
  // any E elements make a thread's initial run (the input order is arbitrary): take them strided,
  // which is bank-conflict free (tid * E + i would be an E-way conflict)
#pragma unroll
  for (int i = 0; i < E; ++i) x[i] = s[i * NM_DEEP_THREADS + tid];
  __syncthreads();  // everybody has its elements before anybody stages into s
  nm_sortnet<E>::run(x);
  constexpr int P = E * NM_DEEP_THREADS;
  // fully unrolled (8 merge levels, <= 8 cross-thread strides each): every stride, and with it the
  // choice shuffle / shared memory and the keep-min / keep-max mask, is a compile-time constant --
  // the rolled form spent 40 % of its instructions on that index arithmetic (profiles/round2_deep_experiments.md)
#pragma unroll
  for (int k = 2 * E; k <= P; k <<= 1) {
    nm_deep_exchange<E, T>(x, s, tid, k / E - 1, true);
#pragma unroll
    for (int j = k >> 2; j >= E; j >>= 1) nm_deep_exchange<E, T>(x, s, tid, j / E, false);
#pragma unroll
    for (int j = E >> 1; j > 0; j >>= 1) {
#pragma unroll
      for (int i = 0; i < E; ++i)
        if ((i & j) == 0) nm_ce_up(x[i], x[i | j]);
    }
  }
  __syncthreads();
  // sorted order, stored SKEWED (element p at p + p/32; nm_view_skew): tid * E + i alone is an E-way
  // bank conflict here and in every later access pattern with a stride of 8 or 16
#pragma unroll
  for (int i = 0; i < E; ++i) {
    const int p = tid * E + i;
    s[p + (p >> 5)] = x[i];
  }
  if (tid == 0) s[P + (P >> 5)] = nm_deep_pinf<T>();  // the sentinel one past the end
  __syncthreads();
}

template <int EMAX, class T>
__device__ __forceinline__ void nm_deep_sort_p(T* s, int P, int tid) {
  switch (P / NM_DEEP_THREADS) {
    case 2: nm_deep_sort<2, T>(s, tid); break;
    case 4: nm_deep_sort<4, T>(s, tid); break;
    case 8: nm_deep_sort<8, T>(s, tid); break;
    case 16: nm_deep_sort<16, T>(s, tid); break;
    case 32: if (EMAX >= 32) nm_deep_sort<(EMAX >= 32 ? 32 : 2), T>(s, tid); break;
    case 64: if (EMAX >= 64) nm_deep_sort<(EMAX >= 64 ? 64 : 2), T>(s, tid); break;
    default: if (EMAX >= 128) nm_deep_sort<(EMAX >= 128 ? 128 : 2), T>(s, tid); break;
  }
}

// KS numerator over this thread's piece [lo, hi) of the pooled order.  sa[n0] and sb[n1] are +inf.
template <class V>
__device__ __forceinline__ int nm_deep_walk(const V sa, int n0, const V sb, int n1, int lo, int hi) {
  // merge-path split of diagonal lo under the rule "ties: group 0 first"
  int il = lo - n1 > 0 ? lo - n1 : 0, ih = lo < n0 ? lo : n0;
  while (il < ih) {
    const int mid = (il + ih) >> 1;
    if (sa[mid] <= sb[lo - mid - 1]) il = mid + 1; else ih = mid;
  }
  int i = il, j = lo - il;
  auto va = sa[i], vb = sb[j];
  auto v = nm_min(va, vb);
  int dmax = 0;
  for (int s = lo; s < hi; ++s) {
    const bool le = va <= vb;
    i += le ? 1 : 0;
    j += le ? 0 : 1;
    va = sa[i];
    vb = sb[j];
    const auto vn = nm_min(va, vb);
    const bool q = vn > v;
    v = vn;
    int d = i * n1 - j * n0;
    d = d < 0 ? -d : d;
    dmax = (q && d > dmax) ? d : dmax;
  }
  return dmax;
}

// EMAX = largest per-thread chunk compiled in: 16 covers groups of up to 4096 reads at 3 CTAs/SM,
// 128 (groups up to 32768 reads) needs most of the register file for one CTA.
template <int EMAX>
__global__ void __launch_bounds__(NM_DEEP_THREADS, EMAX <= 16 ? 4 : 1)
nm_deep_kernel(const nm_kargs a, const int want_u, const int want_t, const int want_m) {
  extern __shared__ __align__(128) unsigned char nm_smem[];
  __shared__ double red_d[NM_DEEP_THREADS / 32];
  __shared__ long long red_l[3][NM_DEEP_THREADS / 32];
  __shared__ double bcast[2];
  uint64_t* bar = reinterpret_cast<uint64_t*>(nm_smem);
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;

  // rows blockIdx.x, blockIdx.x + gridDim.x, ... of the list (its length may live on the device: the retry list of
  // the key-pair kernel, usually empty -- a grid of one block per SM slot then costs a few microseconds)
  const int n_list = a.deep_count_ptr ? *a.deep_count_ptr : a.n_deep;
  if ((int)blockIdx.x >= n_list) return;
  if (tid == 0) nm_mbar_init(bar, 1);
  __syncthreads();
  unsigned parity = 0;
  for (int b = blockIdx.x; b < n_list; b += gridDim.x) {
  const int64_t r = a.deep_rows[b];
  const int32_t src = a.row_pos_index[r];
  const int n0 = a.row_n0[r], n1 = a.row_n1[r];
  const long long s0 = a.off0[src], s1 = a.off1[src];
  const int P0 = nm_deep_p2(n0), P1 = nm_deep_p2(n1);
  if (P0 + P1 > NM_DEEP_TIER_MAX_POOLED) continue;  // does not fit shared memory: nm_huge.cu takes the row
  const long long al0 = s0 & ~3LL, al1 = s1 & ~3LL;
  const int sh0 = (int)(s0 - al0), sh1 = (int)(s1 - al1);
  // [raw A: P0 + 8 floats][raw B: P1 + 8 floats]; the arrays start at the row's first value
  float* rawA = reinterpret_cast<float*>(nm_smem + 16);
  float* rawB = rawA + ((P0 + (P0 >> 5) + 8 + 3) & ~3);  // room for the skewed sorted form of group 0
  float* sa = rawA + sh0;
  float* sb = rawB + sh1;

  __syncthreads();  // the previous row's arrays are no longer read
  if (tid == 0) {
    const uint32_t b0 = (uint32_t)((sh0 + n0 + 3) & ~3) * 4u;
    const uint32_t b1 = (uint32_t)((sh1 + n1 + 3) & ~3) * 4u;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    nm_mbar_expect_tx(bar, b0 + b1);
    nm_bulk_g2s(rawA, a.vals0 + al0, b0, bar);
    nm_bulk_g2s(rawB, a.vals1 + al1, b1, bar);
  }
  __syncthreads();
  nm_mbar_wait(bar, parity);
  parity ^= 1u;
  __syncthreads();  // nobody pads before everyone has seen the copy complete
  for (int k = n0 + tid; k <= P0; k += NM_DEEP_THREADS) sa[k] = NM_INF;  // pads + one sentinel
  for (int k = n1 + tid; k <= P1; k += NM_DEEP_THREADS) sb[k] = NM_INF;

  double mean[2] = {0.0, 0.0}, var[2] = {0.0, 0.0};
  if (want_m) {  // Welch t and/or the --mstd output
    for (int g = 0; g < 2; ++g) {
      const float* s = g ? sb : sa;
      const int n = g ? n1 : n0;
      double part = 0.0;
      for (int k = tid; k < n; k += NM_DEEP_THREADS) part += (double)s[k];
      part = nm_warp_sum_d(part);
      if (lane == 0) red_d[wid] = part;
      __syncthreads();
      if (tid == 0) {
        double t = 0.0;
        for (int w = 0; w < NM_DEEP_THREADS / 32; ++w) t += red_d[w];
        bcast[0] = t / (double)n;
      }
      __syncthreads();
      const double m = bcast[0];
      part = 0.0;
      for (int k = tid; k < n; k += NM_DEEP_THREADS) {
        const double d = (double)s[k] - m;
        part += d * d;
      }
      part = nm_warp_sum_d(part);
      __syncthreads();
      if (lane == 0) red_d[wid] = part;
      __syncthreads();
      if (tid == 0) {
        double t = 0.0;
        for (int w = 0; w < NM_DEEP_THREADS / 32; ++w) t += red_d[w];
        bcast[1] = t / (double)(n - 1);
      }
      __syncthreads();
      mean[g] = m;
      var[g] = bcast[1];
      __syncthreads();
    }
  }
  __syncthreads();
  nm_deep_acc acc;
  nm_deep_acc_init(&acc);
  const int T = n0 + n1;
  nm_deep_sort_p<EMAX, float>(sa, P0, tid);
  nm_deep_sort_p<EMAX, float>(sb, P1, tid);
  if (want_u) {
    for (int e = tid; e < T; e += NM_DEEP_THREADS) {
      nm_deep_acc one;
      nm_deep_acc_init(&one);
      nm_deep_element(nm_view_skew{sa}, n0, nm_view_skew{sb}, n1, e, true, &one);
      nm_deep_acc_merge(&acc, one);
    }
  } else {
    const int per = (T + NM_DEEP_THREADS - 1) / NM_DEEP_THREADS;
    const int wlo = tid * per < T ? tid * per : T;
    const int whi = wlo + per < T ? wlo + per : T;
    acc.dnum = nm_deep_walk(nm_view_skew{sa}, n0, nm_view_skew{sb}, n1, wlo, whi);
  }
  acc.dnum = nm_warp_max_ll(acc.dnum);
  acc.r2 = nm_warp_sum_ll(acc.r2);
  acc.tie = nm_warp_sum_ll(acc.tie);
  if (lane == 0) {
    red_l[0][wid] = acc.dnum;
    red_l[1][wid] = acc.r2;
    red_l[2][wid] = acc.tie;
  }
  __syncthreads();
  if (tid == 0) {
    nm_deep_acc tot;
    nm_deep_acc_init(&tot);
    for (int w = 0; w < NM_DEEP_THREADS / 32; ++w) {
      nm_deep_acc one;
      one.dnum = red_l[0][w];
      one.r2 = red_l[1][w];
      one.tie = red_l[2][w];
      nm_deep_acc_merge(&tot, one);
    }
    nm_row_out o;
    o.two_u = 0;
    o.u_stat = o.u_p = o.t_stat = o.t_p = 0.0;
    nm_deep_finish(tot, n0, n1, want_u != 0, want_t != 0, mean[0], var[0], mean[1], var[1], &o);
    nm_store_row(a, r, o, want_u != 0, want_t != 0);
    if (want_m && a.acc_mom) reinterpret_cast<double4*>(a.acc_mom)[r] = make_double4(mean[0], var[0], mean[1], var[1]);
  }
  }  // rows of this block
}

// ------------------------------------------------------------------------------------------
// Deep tier, grid keys: ONE WARP per position, for positions whose values are three-place decimals
// (nm_lane.cuh "Grid keys") and whose groups have at most 2048 reads.  The two groups are packed
// into ONE array of P = 32 E key pairs (P = 512, 1024, 2048), E pairs per lane IN REGISTERS:
//   load    straight from global memory, coalesced (pair i of lane l is element l + 32 i of both
//           groups -- any assignment will do, the data are about to be sorted), every value checked;
//   sort    a register network sorts the lane's E pairs (packed 16-bit min / max: both groups at
//           once), then log2(32) merge levels of the normalised bitonic network: the cross-lane
//           strides by shuffle, the strides below E as register compare-exchanges.  51 of the 66
//           stage-steps of a 2048-sort never leave the registers (block kernel: 30) and nothing
//           goes through shared memory or a barrier: 7.2 k warp-instructions per position instead
//           of 19.5 k;
//   ranks   the sorted pairs go to the warp's shared-memory array (skewed by one word per lane chunk)
//           and every lane walks its piece of the pooled order (or counts by binary search for U).
// Positions it cannot take (a value off the grid, a longer group) go to a.deep_retry_rows, which
// nm_deep_kernel works afterwards.
// ------------------------------------------------------------------------------------------
#define NM_DEEPW_WARPS 4
#define NM_DEEPW_MAX_E 64

template <int LOG2E>
struct nm_view_chunk16 {  // sorted pair p at p + (p >> LOG2E); a view picks a half
  const unsigned* p;
  int sh;
  __device__ __forceinline__ int operator[](int i) const { return (int)((p[i + (i >> LOG2E)] >> sh) & 0xffffu); }
};

// keep the minimum (lower lane of the pair) or the maximum, half by half.  Written with the SIMD intrinsics, not
// nm_min / nm_max (inline PTX): the compiler predicates the two VIMNMX.U16x2 on `lower` instead of computing both
// and selecting (one instruction less per element and stage).
__device__ __forceinline__ nm_p16 nm_keep(nm_p16 a, nm_p16 b, bool lower) {
  nm_p16 r;
  r.v = lower ? __vminu2(a.v, b.v) : __vmaxu2(a.v, b.v);
  return r;
}

// exchange with lane ^ M: partner pair i, or E-1-i in the "flip" step that opens a merge level
template <int E, bool REVERSED>
__device__ __forceinline__ void nm_deepw_exchange(nm_p16 (&x)[E], int M, bool lower) {
  if (REVERSED) {
#pragma unroll
    for (int i = 0; i < E / 2; ++i) {
      const int r = E - 1 - i;
      const nm_p16 o1 = nm_shfl_xor(x[r], M), o2 = nm_shfl_xor(x[i], M);
      x[i] = nm_keep(x[i], o1, lower);
      x[r] = nm_keep(x[r], o2, lower);
    }
  } else {
#pragma unroll
    for (int i = 0; i < E; ++i) x[i] = nm_keep(x[i], nm_shfl_xor(x[i], M), lower);
  }
}

template <int E, int LOG2E>
__device__ __forceinline__ void nm_deepw_row(const nm_kargs& a, unsigned* sm, int64_t r, int n0, int n1, long long s0,
                                             long long s1, int lane, int want_u, int want_t, int want_m) {
  constexpr int P = 32 * E;
  nm_p16 x[E];
  nm_grid_flag bad = NM_GRID_FLAG0;
  float vmax = 0.0f;
  double sum0 = 0.0, sum1 = 0.0;
  const float* __restrict__ g0 = a.vals0 + s0;
  const float* __restrict__ g1 = a.vals1 + s1;
#pragma unroll
  for (int i0 = 0; i0 < E; i0 += 8) {
    float va[8], vb[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int k = lane + 32 * (i0 + j);
      va[j] = k < n0 ? __ldg(g0 + k) : 0.0f;
      vb[j] = k < n1 ? __ldg(g1 + k) : 0.0f;
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int k = lane + 32 * (i0 + j);
      const unsigned ta = nm_grid_bits(va[j], NM_GRID_MA, &bad), tb = nm_grid_bits(vb[j], NM_GRID_MB, &bad);
      vmax = fmaxf(vmax, fmaxf(fabsf(va[j]), fabsf(vb[j])));
      x[i0 + j].v = (k < n1 ? tb : NM_GRID_PAD_B) * 65536u + (k < n0 ? ta : NM_GRID_PAD_A);
      if (want_m) {
        sum0 += (double)va[j];
        sum1 += (double)vb[j];
      }
    }
  }
  if (__any_sync(0xffffffffu, nm_grid_failed(bad) || !(vmax <= NM_GRID_LIM))) {
    if (lane == 0) a.deep_retry_rows[atomicAdd(a.deep_retry_count, 1)] = (int32_t)r;
    return;
  }
  double mean0 = 0.0, var0 = 0.0, mean1 = 0.0, var1 = 0.0;
  if (want_m) {  // second pass over the values (L1 / L2 by now): numpy's two-pass variance
    mean0 = nm_warp_sum_d(sum0) / (double)n0;
    mean1 = nm_warp_sum_d(sum1) / (double)n1;
    double q0 = 0.0, q1 = 0.0;
    for (int k = lane; k < n0; k += 32) {
      const double d = (double)__ldg(g0 + k) - mean0;
      q0 += d * d;
    }
    for (int k = lane; k < n1; k += 32) {
      const double d = (double)__ldg(g1 + k) - mean1;
      q1 += d * d;
    }
    var0 = nm_warp_sum_d(q0) / (double)(n0 - 1);
    var1 = nm_warp_sum_d(q1) / (double)(n1 - 1);
  }

  const int one = a.one, mone = a.mone;  // runtime +-1: keeps the {min, a + b - min} flavour's additions IMADs (FMA pipe)
  nm_sortnet<E>::run(x, one, mone);
  // merge levels: the lane chunks are sorted runs of E; level lv merges runs of E << (lv-1) pairwise
  // (rolled: ptxas parks every predicated result in a temporary and moves it back, 5 instead of 3 instructions per
  // pair and stage -- but the unrolled form keeps the moves AND misses the instruction cache: 0.95 -> 1.09 ms)
#pragma unroll 1
  for (int lv = 1; lv <= 5; ++lv) {
    const int half = 1 << (lv - 1);
    nm_deepw_exchange<E, true>(x, 2 * half - 1, (lane & half) == 0);
#pragma unroll 1
    for (int m = half >> 1; m >= 1; m >>= 1) nm_deepw_exchange<E, false>(x, m, (lane & m) == 0);
#pragma unroll
    for (int j = E >> 1; j > 0; j >>= 1) {
#pragma unroll
      for (int i = 0; i < E; ++i)
        if ((i & j) == 0) {  // the kernel is ALU-bound: two of three comparators in the IMAD flavour
          if ((i + (i >> 3)) % 3 == 0) nm_ce_up(x[i], x[i | j]); else nm_ceb(x[i], x[i | j], one, mone);
        }
    }
  }
  __syncwarp();  // the previous position's walk is over on every lane
  {
    unsigned* dst = sm + lane * (E + 1);
#pragma unroll
    for (int i = 0; i < E; ++i) dst[i] = x[i].v;
    if (lane == 31) dst[E + 1] = NM_GRID_PINF;  // pair P: the sentinel one past the end (P + (P >> LOG2E))
  }
  __syncwarp();

  const nm_view_chunk16<LOG2E> ga{sm, 0}, gb{sm, 16};
  nm_deep_acc acc;
  nm_deep_acc_init(&acc);
  const int T = n0 + n1;
  if (want_u) {
    for (int e = lane; e < T; e += 32) {
      nm_deep_acc one;
      nm_deep_acc_init(&one);
      nm_deep_element(ga, n0, gb, n1, e, true, &one);
      nm_deep_acc_merge(&acc, one);
    }
  } else {
    const int per = (T + 31) >> 5;
    const int wlo = lane * per < T ? lane * per : T;
    const int whi = wlo + per < T ? wlo + per : T;
    acc.dnum = nm_deep_walk(ga, n0, gb, n1, wlo, whi);
  }
  acc.dnum = nm_warp_max_ll(acc.dnum);
  acc.r2 = nm_warp_sum_ll(acc.r2);
  acc.tie = nm_warp_sum_ll(acc.tie);
  if (lane == 0) {
    nm_row_out o;
    o.two_u = 0;
    o.u_stat = o.u_p = o.t_stat = o.t_p = 0.0;
    nm_deep_finish(acc, n0, n1, want_u != 0, want_t != 0, mean0, var0, mean1, var1, &o);
    nm_store_row(a, r, o, want_u != 0, want_t != 0);
    if (want_m && a.acc_mom) reinterpret_cast<double4*>(a.acc_mom)[r] = make_double4(mean0, var0, mean1, var1);
  }
  (void)P;
}

__global__ void __launch_bounds__(32 * NM_DEEPW_WARPS, 4)
nm_deepw_kernel(const nm_kargs a, const int want_u, const int want_t, const int want_m) {
  extern __shared__ __align__(128) unsigned char nm_smem[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  unsigned* sm = reinterpret_cast<unsigned*>(nm_smem) + (size_t)wib * (32 * (NM_DEEPW_MAX_E + 1) + 8);
  const int n_rows = a.deep_count_ptr ? *a.deep_count_ptr : a.n_deep;
  const int n_warps = gridDim.x * NM_DEEPW_WARPS;
  for (int w = blockIdx.x * NM_DEEPW_WARPS + wib; w < n_rows; w += n_warps) {
    const int64_t r = a.deep_rows[w];
    const int32_t src = a.row_pos_index[r];
    const int n0 = a.row_n0[r], n1 = a.row_n1[r];
    const int P0 = nm_deep_p2(n0), P1 = nm_deep_p2(n1);
    if (P0 + P1 > NM_DEEP_TIER_MAX_POOLED) continue;  // nm_huge.cu takes the row
    const int Pm = P0 > P1 ? P0 : P1;
    if (Pm > 32 * NM_DEEPW_MAX_E) {  // a group longer than 2048 reads: the block kernel
      if (lane == 0) a.deep_retry_rows[atomicAdd(a.deep_retry_count, 1)] = (int32_t)r;
      continue;
    }
    const long long s0 = a.off0[src], s1 = a.off1[src];
    if (Pm == 512)
      nm_deepw_row<16, 4>(a, sm, r, n0, n1, s0, s1, lane, want_u, want_t, want_m);
    else if (Pm == 1024)
      nm_deepw_row<32, 5>(a, sm, r, n0, n1, s0, s1, lane, want_u, want_t, want_m);
    else
      nm_deepw_row<64, 6>(a, sm, r, n0, n1, s0, s1, lane, want_u, want_t, want_m);
  }
}

template <int EMAX>
static int nm_launch_deep_t(const nm_kargs& ka, bool want_u, bool want_t, bool want_m, int n_deep, int smem_bytes, int sm_count,
                            cudaStream_t st) {
  cudaError_t e = cudaFuncSetAttribute(nm_deep_kernel<EMAX>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
  if (e != cudaSuccess) return (int)e;
  int blocks = 0;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks, nm_deep_kernel<EMAX>, NM_DEEP_THREADS, (size_t)smem_bytes);
  if (e != cudaSuccess) return (int)e;
  if (blocks < 1) return (int)cudaErrorInvalidConfiguration;
  // one block per row (the hardware overlaps a block's load with its neighbours' sorts); only the launch over the
  // key-pair kernel's retry list -- usually empty -- is sized to the machine and loops
  int grid = n_deep;
  if (ka.deep_count_ptr && grid > blocks * sm_count) grid = blocks * sm_count;
  nm_deep_kernel<EMAX><<<(unsigned)grid, NM_DEEP_THREADS, smem_bytes, st>>>(ka, want_u ? 1 : 0, want_t ? 1 : 0, want_m ? 1 : 0);
  return (int)cudaGetLastError();
}

// max_p2 = largest pow2(n0) + pow2(n1) among the deep rows (each >= NM_DEEP_MIN_P).  With ka.deep_retry_rows set the
// warp-per-position key-pair kernel runs first and the block kernel afterwards over the rows it listed (device-side
// count): positions with a value off the grid, or with more than 2048 reads in a group.
int nm_launch_deep(const nm_kargs& ka_in, bool want_u, bool want_t, bool want_m, int n_deep, int max_p2, int smem_bytes,
                   int sm_count, cudaStream_t st) {
  nm_kargs ka = ka_in;
  ka.n_deep = n_deep;
  if (ka.deep_retry_rows) {
    const int smem_w = NM_DEEPW_WARPS * (32 * (NM_DEEPW_MAX_E + 1) + 8) * (int)sizeof(unsigned);
    cudaError_t e = cudaFuncSetAttribute(nm_deepw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_w);
    if (e != cudaSuccess) return (int)e;
    int ctas = (n_deep + NM_DEEPW_WARPS - 1) / NM_DEEPW_WARPS;
    if (ctas > 4 * sm_count) ctas = 4 * sm_count;
    ka.n_deep = n_deep;
    nm_deepw_kernel<<<(unsigned)ctas, 32 * NM_DEEPW_WARPS, smem_w, st>>>(ka, want_u ? 1 : 0, want_t ? 1 : 0, want_m ? 1 : 0);
    e = cudaGetLastError();
    if (e != cudaSuccess) return (int)e;
    ka.deep_rows = ka.deep_retry_rows;
    ka.deep_count_ptr = ka.deep_retry_count;
  }
  // a group can be at most max_p2 - NM_DEEP_MIN_P long
  if (max_p2 - NM_DEEP_MIN_P <= 16 * NM_DEEP_THREADS) return nm_launch_deep_t<16>(ka, want_u, want_t, want_m, n_deep, smem_bytes, sm_count, st);
  return nm_launch_deep_t<128>(ka, want_u, want_t, want_m, n_deep, smem_bytes, sm_count, st);
}
