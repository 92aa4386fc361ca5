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
// GRID: the 16-bit key-pair sort; a position with a value that is not a three-place decimal is put on
// a.deep_retry_rows instead of being computed, and the float32 instantiation (launched afterwards over that list)
// takes it.  Two instantiations rather than a branch: with both sorts in one kernel the float32 one ran 23 % slower.
template <int EMAX, bool GRID>
__global__ void __launch_bounds__(NM_DEEP_THREADS, EMAX <= 16 ? 4 : 1)
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
  if (P0 + P1 > NM_DEEP_TIER_MAX_POOLED) return;  // does not fit shared memory: nm_huge.cu takes the row
  const long long al0 = s0 & ~3LL, al1 = s1 & ~3LL;
  const int sh0 = (int)(s0 - al0), sh1 = (int)(s1 - al1);
  // [raw A: P0 + 8 floats][raw B: P1 + 8 floats]; the arrays start at the row's first value
  float* rawA = reinterpret_cast<float*>(nm_smem + 16);
  float* rawB = rawA + ((P0 + (P0 >> 5) + 8 + 3) & ~3);  // room for the skewed sorted form of group 0
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
  // ---- grid keys?  every value of both groups must be a three-place decimal within the 16-bit range
  const int Pm = P0 > P1 ? P0 : P1;
  if (GRID) {
    nm_grid_flag bad = NM_GRID_FLAG0;
    float vmax = 0.0f;
    for (int k = tid; k < Pm; k += NM_DEEP_THREADS) {
      const float xa = k < n0 ? sa[k] : 0.0f, xb = k < n1 ? sb[k] : 0.0f;
      (void)nm_grid_bits(xa, NM_GRID_MA, &bad);
      (void)nm_grid_bits(xb, NM_GRID_MB, &bad);
      vmax = fmaxf(vmax, fmaxf(fabsf(xa), fabsf(xb)));
    }
    if (__syncthreads_or(nm_grid_failed(bad) || !(vmax <= NM_GRID_LIM))) {
      if (tid == 0) a.deep_retry_rows[atomicAdd(a.deep_retry_count, 1)] = (int32_t)r;
      return;
    }
  }
  nm_deep_acc acc;
  nm_deep_acc_init(&acc);
  const int T = n0 + n1;
  const int per = (T + NM_DEEP_THREADS - 1) / NM_DEEP_THREADS;
  const int wlo = tid * per < T ? tid * per : T;
  const int whi = wlo + per < T ? wlo + per : T;
  if (GRID) {
    // key pairs in place over the longer group's array (index k is read and written by the same thread)
    nm_p16* pk = reinterpret_cast<nm_p16*>(P0 >= P1 ? sa : sb);
    for (int k = tid; k < Pm; k += NM_DEEP_THREADS) {
      const unsigned ta = k < n0 ? nm_f2u(fmaf(sa[k], NM_GRID_SCALE, NM_GRID_MA)) : NM_GRID_PAD_A;
      const unsigned tb = k < n1 ? nm_f2u(fmaf(sb[k], NM_GRID_SCALE, NM_GRID_MB)) : NM_GRID_PAD_B;
      pk[k].v = tb * 65536u + ta;
    }
    __syncthreads();
    nm_deep_sort_p<EMAX, nm_p16>(pk, Pm, tid);
    const nm_view_skew16 ga{reinterpret_cast<const unsigned*>(pk), 0}, gb{reinterpret_cast<const unsigned*>(pk), 16};
    if (want_u) {
      for (int e = tid; e < T; e += NM_DEEP_THREADS) {
        nm_deep_acc one;
        nm_deep_acc_init(&one);
        nm_deep_element(ga, n0, gb, n1, e, true, &one);
        nm_deep_acc_merge(&acc, one);
      }
    } else {
      acc.dnum = nm_deep_walk(ga, n0, gb, n1, wlo, whi);
    }
  } else {
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
      acc.dnum = nm_deep_walk(nm_view_skew{sa}, n0, nm_view_skew{sb}, n1, wlo, whi);
    }
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

template <int EMAX, bool GRID>
static int nm_launch_deep_t(const nm_kargs& ka, bool want_u, bool want_t, bool want_m, int n_deep, int smem_bytes, cudaStream_t st) {
  cudaError_t e = cudaFuncSetAttribute(nm_deep_kernel<EMAX, GRID>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
  if (e != cudaSuccess) return (int)e;
  nm_deep_kernel<EMAX, GRID><<<(unsigned)n_deep, NM_DEEP_THREADS, smem_bytes, st>>>(ka, want_u ? 1 : 0, want_t ? 1 : 0, want_m ? 1 : 0);
  return (int)cudaGetLastError();
}

// max_p2 = largest pow2(n0) + pow2(n1) among the deep rows (each >= NM_DEEP_MIN_P).  With ka.deep_retry_rows set the
// 16-bit key-pair kernel runs first and the float32 kernel afterwards over the rows it listed (device-side count).
int nm_launch_deep(const nm_kargs& ka_in, bool want_u, bool want_t, bool want_m, int n_deep, int max_p2, int smem_bytes,
                   cudaStream_t st) {
  nm_kargs ka = ka_in;
  // a group can be at most max_p2 - NM_DEEP_MIN_P long
  const bool small = max_p2 - NM_DEEP_MIN_P <= 16 * NM_DEEP_THREADS;
  if (ka.deep_retry_rows) {
    const int e = small ? nm_launch_deep_t<16, true>(ka, want_u, want_t, want_m, n_deep, smem_bytes, st)
                        : nm_launch_deep_t<128, true>(ka, want_u, want_t, want_m, n_deep, smem_bytes, st);
    if (e != (int)cudaSuccess) return e;
    ka.deep_rows = ka.deep_retry_rows;
    ka.deep_count_ptr = ka.deep_retry_count;
  }
  return small ? nm_launch_deep_t<16, false>(ka, want_u, want_t, want_m, n_deep, smem_bytes, st)
               : nm_launch_deep_t<128, false>(ka, want_u, want_t, want_m, n_deep, smem_bytes, st);
}
