// nm_lane_kernel.cu -- the lane tier: one lane per position, one warp (= one CTA) per tile of
// 32 consecutive rows.  getKStest (bin/scripts/myDetect.py:327-343) for rows with <= 128 reads
// per group.  Kept in its own translation unit because the 16 unrolled sorting networks take
// about a minute to compile.
//
// Shared memory of a CTA: [mbarrier 16 B][region A][region B], each region 32*(Ncls+1) floats.
//   1. the tile's two value slices (contiguous in the CSR arrays) arrive with one TMA bulk
//      copy each (cp.async.bulk + mbarrier); non-contiguous tiles use a cooperative gather;
//   2. every lane pulls its row into registers (LDS.128 when aligned), sorts it with the
//      size-Nsel network, and writes it back TRANSPOSED (element i of lane l at [i*32+l]), so
//      the data-dependent indexing of the merge walk is bank-conflict free;
//   3. one merge walk per lane gives the KS numerator (+ rank sums), then the fp64 tails.
#include "nm_device.cuh"

struct nm_smem_col {  // sorted group, transposed: element i of this lane lives at base[i*32]
  const float* base;
  __device__ __forceinline__ float operator()(int i) const { return base[i << 5]; }
};

// Load one group's row into registers (pad +inf), sort, write back transposed (+ sentinel).
template <int N>
__device__ __forceinline__ void nm_lane_sort_group(float* region, int base, int n, int lane,
                                                   bool want_t, double* mean, double* var) {
  float x[N];
  const float* raw = region + base;
  if (want_t) nm_moments(raw, n, mean, var);
  const bool vec = __all_sync(0xffffffffu, ((base & 3) == 0) && (n == N));
  if (vec) {
    const float4* raw4 = reinterpret_cast<const float4*>(raw);
#pragma unroll
    for (int q = 0; q < N / 4; ++q) {
      const float4 v = raw4[q];
      x[4 * q + 0] = v.x;
      x[4 * q + 1] = v.y;
      x[4 * q + 2] = v.z;
      x[4 * q + 3] = v.w;
    }
  } else {
#pragma unroll
    for (int k = 0; k < N; ++k) {
      const float v = raw[k];
      x[k] = (k < n) ? v : NM_INF;
    }
  }
  nm_sortnet<N>::run(x);
  __syncwarp();  // every lane has consumed its raw row; the region may now be overwritten
  float* col = region + lane;
#pragma unroll
  for (int k = 0; k < N; ++k) col[k << 5] = x[k];
  col[N << 5] = NM_INF;
}

template <int N>
__device__ __forceinline__ void nm_lane_tile(float* regA, float* regB, int base0, int base1,
                                             int n0, int n1, int lane, bool want_t,
                                             nm_lane_acc* acc) {
#pragma unroll 1
  for (int g = 0; g < 2; ++g) {
    double m = 0.0, v = 0.0;
    nm_lane_sort_group<N>(g ? regB : regA, g ? base1 : base0, g ? n1 : n0, lane, want_t, &m, &v);
    if (g) { acc->mean1 = m; acc->var1 = v; } else { acc->mean0 = m; acc->var0 = v; }
  }
  __syncwarp();
}

// cooperative copy of the tile's rows when they are not one contiguous slice of vals
__device__ __forceinline__ int nm_lane_gather(float* region, const float* __restrict__ vals,
                                              long long start, int n, int lane) {
  int incl = n;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  const int base = incl - n;
  for (int rlane = 0; rlane < 32; ++rlane) {
    const int rn = __shfl_sync(0xffffffffu, n, rlane);
    const int rbase = __shfl_sync(0xffffffffu, base, rlane);
    const long long rstart = __shfl_sync(0xffffffffu, start, rlane);
    for (int k = lane; k < rn; k += 32) region[rbase + k] = vals[rstart + k];
  }
  return base;
}

__global__ void __launch_bounds__(32, 8) nm_lane_kernel(const nm_kargs a, const int want_u,
                                                        const int want_t) {
  extern __shared__ __align__(128) unsigned char nm_smem[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(nm_smem);
  float* regA = reinterpret_cast<float*>(nm_smem + 16);
  float* regB = regA + a.region_floats;

  const int lane = threadIdx.x;
  const int64_t r = (int64_t)blockIdx.x * 32 + lane;
  int n0 = 0, n1 = 0;
  long long s0 = 0, s1 = 0;
  bool ok = false;
  if (r < a.n_rows) {
    const int nn0 = a.row_n0[r], nn1 = a.row_n1[r];
    if (nn0 <= NM_LANE_TIER_MAX && nn1 <= NM_LANE_TIER_MAX) {
      const int32_t src = a.row_pos_index[r];
      ok = true;
      n0 = nn0;
      n1 = nn1;
      s0 = a.off0[src];
      s1 = a.off1[src];
    }
  }
  if (!__any_sync(0xffffffffu, ok)) return;

  const long long big = 0x7fffffffffffffffLL;
  const long long first0 = nm_warp_min_ll(ok ? s0 : big), first1 = nm_warp_min_ll(ok ? s1 : big);
  const long long end0 = nm_warp_max_ll(ok ? s0 + n0 : -1), end1 = nm_warp_max_ll(ok ? s1 + n1 : -1);
  const int tot0 = __reduce_add_sync(0xffffffffu, n0), tot1 = __reduce_add_sync(0xffffffffu, n1);
  const bool contig0 = (end0 - first0) == (long long)tot0;
  const bool contig1 = (end1 - first1) == (long long)tot1;
  const int nmax = __reduce_max_sync(0xffffffffu, n0 > n1 ? n0 : n1);
  const int tmax = __reduce_max_sync(0xffffffffu, n0 + n1);
  const int nsel = (nmax + NM_LANE_STEP - 1) / NM_LANE_STEP * NM_LANE_STEP;

  // stage the tile: one TMA bulk copy per contiguous group slice, else a cooperative gather
  const long long al0 = first0 & ~3LL, al1 = first1 & ~3LL;
  int base0 = ok ? (int)(s0 - al0) : 0, base1 = ok ? (int)(s1 - al1) : 0;
  if (contig0 || contig1) {
    if (lane == 0) {
      nm_mbar_init(bar, 1);
      const uint32_t b0 = contig0 ? (uint32_t)(((first0 - al0) + tot0 + 3) & ~3LL) * 4u : 0u;
      const uint32_t b1 = contig1 ? (uint32_t)(((first1 - al1) + tot1 + 3) & ~3LL) * 4u : 0u;
      nm_mbar_expect_tx(bar, b0 + b1);
      if (contig0) nm_bulk_g2s(regA, a.vals0 + al0, b0, bar);
      if (contig1) nm_bulk_g2s(regB, a.vals1 + al1, b1, bar);
    }
    __syncwarp();
  }
  if (!contig0) base0 = nm_lane_gather(regA, a.vals0, s0, n0, lane);
  if (!contig1) base1 = nm_lane_gather(regB, a.vals1, s1, n1, lane);
  if (contig0 || contig1) nm_mbar_wait(bar, 0);
  __syncwarp();

  nm_lane_acc acc;
  acc.dnum = acc.r2 = acc.tie = 0;
  acc.mean0 = acc.var0 = acc.mean1 = acc.var1 = 0.0;
#define NM_CALL(NN) nm_lane_tile<NN>(regA, regB, base0, base1, n0, n1, lane, want_t != 0, &acc)
  NM_DISPATCH_N(nsel, NM_CALL)
#undef NM_CALL

  const nm_smem_col A{regA + lane}, B{regB + lane};
  if (want_u)
    nm_merge_walk<true>(n0, n1, tmax, A, B, &acc);
  else
    nm_merge_walk<false>(n0, n1, tmax, A, B, &acc);

  if (ok) {
    nm_row_out o;
    nm_lane_finish(acc, n0, n1, want_u != 0, want_t != 0, &o);
    nm_store_row(a, r, o, want_u != 0, want_t != 0);
  }
}

int nm_launch_lane(const nm_kargs& ka, bool want_u, bool want_t, int smem_bytes, cudaStream_t st) {
  cudaError_t e = cudaFuncSetAttribute(nm_lane_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       smem_bytes);
  if (e != cudaSuccess) return (int)e;
  const unsigned grid = (unsigned)((ka.n_rows + 31) / 32);
  nm_lane_kernel<<<grid, 32, smem_bytes, st>>>(ka, want_u ? 1 : 0, want_t ? 1 : 0);
  return (int)cudaGetLastError();
}
