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
#include "nm_device.cuh"

__device__ __forceinline__ void nm_ce_up(float& a, float& b) {
  const float lo = fminf(a, b), hi = fmaxf(a, b);
  a = lo;
  b = hi;
}

// exchange step of the normalised bitonic network across threads: partner thread tid ^ M,
// partner element i (same) or E-1-i (reversed: the "flip" first step of a merge)
template <int E>
__device__ __forceinline__ void nm_deep_exchange(float (&x)[E], float* stage, int tid, int M, bool reversed) {
  int hb = M;  // highest set bit of M decides who keeps the minimum
  hb |= hb >> 1; hb |= hb >> 2; hb |= hb >> 4; hb |= hb >> 8;
  hb = (hb + 1) >> 1;
  const bool lower = (tid & hb) == 0;
  if (M < 32) {
    if (reversed) {
#pragma unroll
      for (int i = 0; i < (E + 1) / 2; ++i) {
        const int r = E - 1 - i;
        const float o1 = __shfl_xor_sync(0xffffffffu, x[r], M);  // partner's counterpart of my x[i]
        const float o2 = __shfl_xor_sync(0xffffffffu, x[i], M);  // partner's counterpart of my x[r]
        x[i] = lower ? fminf(x[i], o1) : fmaxf(x[i], o1);
        if (r != i) x[r] = lower ? fminf(x[r], o2) : fmaxf(x[r], o2);
      }
    } else {
#pragma unroll
      for (int i = 0; i < E; ++i) {
        const float o = __shfl_xor_sync(0xffffffffu, x[i], M);
        x[i] = lower ? fminf(x[i], o) : fmaxf(x[i], o);
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
      const float o = stage[(reversed ? E - 1 - i : i) * NM_DEEP_THREADS + pt];
      x[i] = lower ? fminf(x[i], o) : fmaxf(x[i], o);
    }
  }
}

// sort s[0 .. E*256) ascending in place (s is also used as the staging area)
template <int E>
__device__ __forceinline__ void nm_deep_sort(float* s, int tid) {
  float x[E];
#pragma unroll
  for (int i = 0; i < E; ++i) x[i] = s[tid * E + i];
  nm_sortnet<E>::run(x);
  constexpr int P = E * NM_DEEP_THREADS;
  for (int k = 2 * E; k <= P; k <<= 1) {
    nm_deep_exchange<E>(x, s, tid, k / E - 1, true);
    for (int j = k >> 2; j >= E; j >>= 1) nm_deep_exchange<E>(x, s, tid, j / E, false);
#pragma unroll
    for (int j = E >> 1; j > 0; j >>= 1) {
#pragma unroll
      for (int i = 0; i < E; ++i)
        if ((i & j) == 0) nm_ce_up(x[i], x[i | j]);
    }
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < E; ++i) s[tid * E + i] = x[i];
  __syncthreads();
}

template <int EMAX>
__device__ __forceinline__ void nm_deep_sort_p(float* s, int P, int tid) {
  switch (P / NM_DEEP_THREADS) {
    case 2: nm_deep_sort<2>(s, tid); break;
    case 4: nm_deep_sort<4>(s, tid); break;
    case 8: nm_deep_sort<8>(s, tid); break;
    case 16: nm_deep_sort<16>(s, tid); break;
    case 32: if (EMAX >= 32) nm_deep_sort<(EMAX >= 32 ? 32 : 2)>(s, tid); break;
    case 64: if (EMAX >= 64) nm_deep_sort<(EMAX >= 64 ? 64 : 2)>(s, tid); break;
    default: if (EMAX >= 128) nm_deep_sort<(EMAX >= 128 ? 128 : 2)>(s, tid); break;
  }
}

// KS numerator over this thread's piece [lo, hi) of the pooled order.  sa[n0] and sb[n1] are +inf.
__device__ __forceinline__ int nm_deep_walk(const float* sa, int n0, const float* sb, int n1, int lo, int hi) {
  // merge-path split of diagonal lo under the rule "ties: group 0 first"
  int il = lo - n1 > 0 ? lo - n1 : 0, ih = lo < n0 ? lo : n0;
  while (il < ih) {
    const int mid = (il + ih) >> 1;
    if (sa[mid] <= sb[lo - mid - 1]) il = mid + 1; else ih = mid;
  }
  int i = il, j = lo - il;
  float va = sa[i], vb = sb[j];
  float v = fminf(va, vb);
  int dmax = 0;
  for (int s = lo; s < hi; ++s) {
    const bool le = va <= vb;
    i += le ? 1 : 0;
    j += le ? 0 : 1;
    va = sa[i];
    vb = sb[j];
    const float vn = fminf(va, vb);
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
__global__ void __launch_bounds__(NM_DEEP_THREADS, EMAX <= 16 ? 3 : 1)
nm_deep_kernel(const nm_kargs a, const int want_u, const int want_t, const int want_m) {
  extern __shared__ __align__(128) unsigned char nm_smem[];
  __shared__ double red_d[NM_DEEP_THREADS / 32];
  __shared__ long long red_l[3][NM_DEEP_THREADS / 32];
  __shared__ double bcast[2];
  uint64_t* bar = reinterpret_cast<uint64_t*>(nm_smem);
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;

  if (a.deep_count_ptr && (int)blockIdx.x >= *a.deep_count_ptr) return;  // the list is shorter than the grid
  const int64_t r = a.deep_rows[blockIdx.x];
  const int32_t src = a.row_pos_index[r];
  const int n0 = a.row_n0[r], n1 = a.row_n1[r];
  const long long s0 = a.off0[src], s1 = a.off1[src];
  const int P0 = nm_deep_p2(n0), P1 = nm_deep_p2(n1);
  const long long al0 = s0 & ~3LL, al1 = s1 & ~3LL;
  const int sh0 = (int)(s0 - al0), sh1 = (int)(s1 - al1);
  // [raw A: P0 + 8 floats][raw B: P1 + 8 floats]; the arrays start at the row's first value
  float* rawA = reinterpret_cast<float*>(nm_smem + 16);
  float* rawB = rawA + P0 + 8;
  float* sa = rawA + sh0;
  float* sb = rawB + sh1;

  if (tid == 0) {
    nm_mbar_init(bar, 1);
    const uint32_t b0 = (uint32_t)((sh0 + n0 + 3) & ~3) * 4u;
    const uint32_t b1 = (uint32_t)((sh1 + n1 + 3) & ~3) * 4u;
    nm_mbar_expect_tx(bar, b0 + b1);
    nm_bulk_g2s(rawA, a.vals0 + al0, b0, bar);
    nm_bulk_g2s(rawB, a.vals1 + al1, b1, bar);
  }
  __syncthreads();
  nm_mbar_wait(bar, 0);
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
  nm_deep_sort_p<EMAX>(sa, P0, tid);
  nm_deep_sort_p<EMAX>(sb, P1, tid);

  nm_deep_acc acc;
  nm_deep_acc_init(&acc);
  if (want_u) {
    for (int e = tid; e < n0 + n1; e += NM_DEEP_THREADS) {
      nm_deep_acc one;
      nm_deep_acc_init(&one);
      nm_deep_element(sa, n0, sb, n1, e, true, &one);
      nm_deep_acc_merge(&acc, one);
    }
  } else {
    const int T = n0 + n1;
    const int per = (T + NM_DEEP_THREADS - 1) / NM_DEEP_THREADS;
    const int lo = tid * per < T ? tid * per : T;
    const int hi = lo + per < T ? lo + per : T;
    acc.dnum = nm_deep_walk(sa, n0, sb, n1, lo, hi);
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
}

// ------------------------------------------------------------------------------------------
// Binned variant (the default for deep rows of up to NM_DEEP2_MAX_T pooled reads).
//
// No sort at all.  The KS numerator is max over pooled x of |#{a <= x} * n1 - #{b <= x} * n0|
// (searchsorted(side='right') of scipy's ks_2samp), and the rank statistics need #{. < x} as
// well -- COUNTS, not an order.  So: map every value to one of NB value bins by a monotone
// function (clamped affine map of mean +- 4.5 sd), counting-sort both groups by bin in shared
// memory (count, scan, scatter), and let every element compare itself with the members of its OWN
// bin only; everything in lower bins is smaller, by monotonicity, and is covered by the bin's
// prefix count.  With NB ~ T/2 bins a bin holds a handful of values, so the work per element is
// a few dozen instructions where the bitonic network above spends several hundred, and ties,
// which a sort-and-walk has to treat specially, need nothing: tied values get equal counts.
// Rows whose values pile up in few bins (massive ties, all-identical pools) would make the
// in-bin comparisons quadratic; the scan phase measures that (sum of squared bin sizes) and such
// rows, like rows too long for shared memory, are passed on to nm_deep_kernel through a list.
// ------------------------------------------------------------------------------------------
#define NM_DEEP2_MAX_T 10240      // pooled reads per position that fit 2 CTAs/SM
#define NM_DEEP2_MAX_BINS 2048
#define NM_DEEP2_WORK_FACTOR 48   // accepted in-bin comparisons per element (average) before falling back

struct nm_deep2_args {
  nm_kargs k;
  int nbins;           // power of two <= NM_DEEP2_MAX_BINS
  int smem_floats;     // capacity of each of the raw / binned areas (floats)
  int32_t* fallback;   // rows handed to nm_deep_kernel
  int* fallback_count;
};

__device__ __forceinline__ int nm_deep2_bin(float x, float lo, float scale, int nbins) {
  // monotone non-decreasing in x: affine map, clamp, truncate
  float t = (x - lo) * scale;
  t = fminf(fmaxf(t, 0.0f), (float)(nbins - 1));
  return (int)t;
}

// order-preserving int image of a float (for the per-bin min / max)
__device__ __forceinline__ int nm_deep2_key(float x) {
  const int b = __float_as_int(x);
  return b >= 0 ? b : b ^ 0x7fffffff;
}

__device__ __forceinline__ long long nm_block_reduce_ll(long long v, long long* red, bool is_max) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  v = is_max ? nm_warp_max_ll(v) : nm_warp_sum_ll(v);
  __syncthreads();
  if (lane == 0) red[wid] = v;
  __syncthreads();
  long long t = red[0];
  for (int w = 1; w < NM_DEEP_THREADS / 32; ++w) t = is_max ? (red[w] > t ? red[w] : t) : t + red[w];
  return t;
}

__global__ void __launch_bounds__(NM_DEEP_THREADS, 2)
nm_deep2_kernel(const nm_deep2_args g, const int want_u, const int want_t, const int want_m) {
  extern __shared__ __align__(128) unsigned char nm_smem[];
  __shared__ double red_d[NM_DEEP_THREADS / 32];
  __shared__ long long red_l[NM_DEEP_THREADS / 32];
  __shared__ double bcast[2];
  __shared__ float bc_f[2];
  const nm_kargs& a = g.k;
  uint64_t* bar = reinterpret_cast<uint64_t*>(nm_smem);
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int NB = g.nbins;

  const int64_t r = a.deep_rows[blockIdx.x];
  const int32_t src = a.row_pos_index[r];
  const int n0 = a.row_n0[r], n1 = a.row_n1[r];
  const int T = n0 + n1;
  const long long s0 = a.off0[src], s1 = a.off1[src];
  const long long al0 = s0 & ~3LL, al1 = s1 & ~3LL;
  const int sh0 = (int)(s0 - al0), sh1 = (int)(s1 - al1);
  const int len0 = (sh0 + n0 + 3) & ~3, len1 = (sh1 + n1 + 3) & ~3;
  if (len0 + len1 > g.smem_floats) {  // does not fit this launch's shared memory: the sorting kernel takes it
    if (tid == 0) g.fallback[atomicAdd(g.fallback_count, 1)] = (int32_t)r;
    return;
  }
  // [bar 16 B][raw: smem_floats][binned: smem_floats][startA NB+1][startB NB+1][curA NB][curB NB][bmin NB][bmax NB]
  float* raw = reinterpret_cast<float*>(nm_smem + 16);
  float* binned = raw + g.smem_floats;
  int* startA = reinterpret_cast<int*>(binned + g.smem_floats);
  int* startB = startA + NB + 1;
  int* curA = startB + NB + 1;
  int* curB = curA + NB;
  int* bmin = curB + NB;  // smallest / largest value of a bin (order-preserving int images): a bin whose
  int* bmax = bmin + NB;  // values are all equal needs no comparisons at all
  const float* sa = raw + sh0;
  const float* sb = raw + len0 + sh1;
  float* ba = binned;
  float* bb = binned + n0;

  if (tid == 0) {
    nm_mbar_init(bar, 1);
    nm_mbar_expect_tx(bar, (uint32_t)(len0 + len1) * 4u);
    nm_bulk_g2s(raw, a.vals0 + al0, (uint32_t)len0 * 4u, bar);
    nm_bulk_g2s(raw + len0, a.vals1 + al1, (uint32_t)len1 * 4u, bar);
  }
  for (int b = tid; b < NB; b += NM_DEEP_THREADS) {
    curA[b] = 0;
    curB[b] = 0;
    bmin[b] = 0x7fffffff;
    bmax[b] = (int)0x80000000;
  }
  __syncthreads();
  nm_mbar_wait(bar, 0);

  // ---- exact fp64 moments (Welch t / --mstd), as in nm_deep_kernel; also give the binning range
  double mean[2] = {0.0, 0.0}, var[2] = {0.0, 0.0};
  for (int grp = 0; grp < 2; ++grp) {
    const float* s = grp ? sb : sa;
    const int n = grp ? n1 : n0;
    double part = 0.0;
    for (int k = tid; k < n; k += NM_DEEP_THREADS) part += (double)s[k];
    part = nm_warp_sum_d(part);
    __syncthreads();
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
    mean[grp] = m;
    var[grp] = bcast[1];
  }
  if (tid == 0) {
    // pooled range: the two group means +- 4.5 of the larger sd (values outside fall into the edge bins)
    const double sd = sqrt(var[0] > var[1] ? var[0] : var[1]);
    const double lo = (mean[0] < mean[1] ? mean[0] : mean[1]) - 4.5 * sd;
    const double hi = (mean[0] > mean[1] ? mean[0] : mean[1]) + 4.5 * sd;
    const double w = hi - lo;
    bc_f[0] = (float)lo;
    bc_f[1] = (w > 0.0 && w < 1e30) ? (float)((double)NB / w) : 0.0f;  // 0: everything in bin 0 -> falls back
  }
  __syncthreads();
  const float lo = bc_f[0], scale = bc_f[1];

  // ---- counting sort by bin: count, scan, scatter
  for (int k = tid; k < T; k += NM_DEEP_THREADS) {
    const float x = k < n0 ? sa[k] : sb[k - n0];
    const int b = nm_deep2_bin(x, lo, scale, NB);
    atomicAdd(k < n0 ? &curA[b] : &curB[b], 1);
    const int key = nm_deep2_key(x);
    if (key < bmin[b]) atomicMin(&bmin[b], key);
    if (key > bmax[b]) atomicMax(&bmax[b], key);
  }
  __syncthreads();
  {
    // exclusive scan of NB counts per group: NB / 256 consecutive bins per thread + block scan
    const int per = NB / NM_DEEP_THREADS > 0 ? NB / NM_DEEP_THREADS : 1;
    const int b0 = tid * per;
    int sumA = 0, sumB = 0;
    long long work = 0;
    if (b0 < NB) {
      for (int b = b0; b < b0 + per; ++b) {
        const int ca = curA[b], cb = curB[b];
        sumA += ca;
        sumB += cb;
        work += bmin[b] >= bmax[b] ? (long long)(ca + cb) : (long long)(ca + cb) * (ca + cb);
      }
    }
    // block exclusive scan of (sumA, sumB) packed in one 64-bit value
    long long packed = ((long long)sumA << 32) | (unsigned)sumB;
    long long inc = packed;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const long long t = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += t;
    }
    __shared__ long long wtot[NM_DEEP_THREADS / 32];
    if (lane == 31) wtot[wid] = inc;
    __syncthreads();
    long long base = 0;
    for (int w = 0; w < wid; ++w) base += wtot[w];
    const long long ex = base + inc - packed;
    int accA = (int)(ex >> 32), accB = (int)(ex & 0xffffffffLL);
    if (b0 < NB) {
      for (int b = b0; b < b0 + per; ++b) {
        const int ca = curA[b], cb = curB[b];
        startA[b] = accA;
        startB[b] = accB;
        curA[b] = accA;
        curB[b] = accB;
        accA += ca;
        accB += cb;
      }
    }
    if (tid == NM_DEEP_THREADS - 1) {
      startA[NB] = n0;
      startB[NB] = n1;
    }
    const long long total_work = nm_block_reduce_ll(work, red_l, false);
    if (total_work > (long long)NM_DEEP2_WORK_FACTOR * T) {  // values pile up in few bins: sort instead
      if (tid == 0) g.fallback[atomicAdd(g.fallback_count, 1)] = (int32_t)r;
      return;
    }
  }
  __syncthreads();
  for (int k = tid; k < n0; k += NM_DEEP_THREADS) {
    const float x = sa[k];
    ba[atomicAdd(&curA[nm_deep2_bin(x, lo, scale, NB)], 1)] = x;
  }
  for (int k = tid; k < n1; k += NM_DEEP_THREADS) {
    const float x = sb[k];
    bb[atomicAdd(&curB[nm_deep2_bin(x, lo, scale, NB)], 1)] = x;
  }
  __syncthreads();

  // ---- every element against its own bin (binned order: a warp's lanes share bins, the loads broadcast)
  long long dmax = 0, r2 = 0, tie = 0;
  for (int e = tid; e < T; e += NM_DEEP_THREADS) {
    const bool is_a = e < n0;
    const float x = is_a ? ba[e] : bb[e - n0];
    const int b = nm_deep2_bin(x, lo, scale, NB);
    const int a_lo = startA[b], a_hi = startA[b + 1], b_lo = startB[b], b_hi = startB[b + 1];
    int ua = a_hi, ub = b_hi, la = a_lo, lb = b_lo;  // a bin of equal values: everything in it ties with x
    if (bmin[b] < bmax[b]) {
      ua = a_lo;
      ub = b_lo;
      for (int k = a_lo; k < a_hi; ++k) {
        const float v = ba[k];
        ua += v <= x ? 1 : 0;
        la += v < x ? 1 : 0;
      }
      for (int k = b_lo; k < b_hi; ++k) {
        const float v = bb[k];
        ub += v <= x ? 1 : 0;
        lb += v < x ? 1 : 0;
      }
    }
    long long d = (long long)ua * n1 - (long long)ub * n0;
    d = d < 0 ? -d : d;
    dmax = d > dmax ? d : dmax;
    if (want_u) {
      const long long lo_r = la + lb, hi_r = ua + ub, t = hi_r - lo_r;
      tie += t * t - 1;
      r2 += is_a ? (lo_r + hi_r + 1) : 0;
    }
  }
  nm_deep_acc tot;
  tot.dnum = nm_block_reduce_ll(dmax, red_l, true);
  tot.r2 = want_u ? nm_block_reduce_ll(r2, red_l, false) : 0;
  tot.tie = want_u ? nm_block_reduce_ll(tie, red_l, false) : 0;
  if (tid == 0) {
    nm_row_out o;
    o.two_u = 0;
    o.u_stat = o.u_p = o.t_stat = o.t_p = 0.0;
    nm_deep_finish(tot, n0, n1, want_u != 0, want_t != 0, mean[0], var[0], mean[1], var[1], &o);
    nm_store_row(a, r, o, want_u != 0, want_t != 0);
    if (want_m && a.acc_mom) reinterpret_cast<double4*>(a.acc_mom)[r] = make_double4(mean[0], var[0], mean[1], var[1]);
  }
}

template <int EMAX>
static int nm_launch_deep_t(const nm_kargs& ka, bool want_u, bool want_t, bool want_m, int n_deep, int smem_bytes, cudaStream_t st) {
  cudaError_t e = cudaFuncSetAttribute(nm_deep_kernel<EMAX>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
  if (e != cudaSuccess) return (int)e;
  nm_deep_kernel<EMAX><<<(unsigned)n_deep, NM_DEEP_THREADS, smem_bytes, st>>>(ka, want_u ? 1 : 0, want_t ? 1 : 0, want_m ? 1 : 0);
  return (int)cudaGetLastError();
}

// max_p2 = largest pow2(n0) + pow2(n1) among the deep rows (each >= NM_DEEP_MIN_P)
int nm_launch_deep(const nm_kargs& ka, bool want_u, bool want_t, bool want_m, int n_deep, int max_p2, int smem_bytes,
                   cudaStream_t st) {
  // a group can be at most max_p2 - NM_DEEP_MIN_P long
  if (max_p2 - NM_DEEP_MIN_P <= 16 * NM_DEEP_THREADS) return nm_launch_deep_t<16>(ka, want_u, want_t, want_m, n_deep, smem_bytes, st);
  return nm_launch_deep_t<128>(ka, want_u, want_t, want_m, n_deep, smem_bytes, st);
}

int nm_launch_deep2(const nm_kargs& ka, bool want_u, bool want_t, bool want_m, int n_deep, int max_t, int32_t* fallback,
                    int* fallback_count, cudaStream_t st) {
  nm_deep2_args g;
  g.k = ka;
  const int t = max_t < NM_DEEP2_MAX_T ? max_t : NM_DEEP2_MAX_T;
  int nb = 256;
  while (nb < NM_DEEP2_MAX_BINS && nb < t / 4) nb <<= 1;  // ~4 values per bin
  g.nbins = nb;
  g.smem_floats = (t + 8 + 3) & ~3;  // two slices rounded out to 16-byte boundaries
  g.fallback = fallback;
  g.fallback_count = fallback_count;
  const int smem_bytes = 16 + 2 * g.smem_floats * (int)sizeof(float) + (6 * nb + 2) * (int)sizeof(int);
  cudaError_t e = cudaFuncSetAttribute(nm_deep2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
  if (e != cudaSuccess) return (int)e;
  nm_deep2_kernel<<<(unsigned)n_deep, NM_DEEP_THREADS, smem_bytes, st>>>(g, want_u ? 1 : 0, want_t ? 1 : 0, want_m ? 1 : 0);
  return (int)cudaGetLastError();
}
