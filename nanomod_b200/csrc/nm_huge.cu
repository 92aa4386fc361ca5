// nm_huge.cu -- positions too deep for the shared-memory deep tier (pow2(n0) + pow2(n1) >
// NM_DEEP_TIER_MAX_POOLED).  The reference has no depth limit (getKStest, bin/scripts/myDetect.py:
// 327-343 takes any two lists), so these rows must not fail the call; they are rare, so they take
// a plain global-memory route:
//   collect   the huge rows of the deep list, with the segment offsets of their values
//   copy      their values into two scratch arrays (group 0 / group 1), one segment per row
//   sort      CUB DeviceSegmentedSort (library code, like the ranking: this is not the hot path)
//   count     one CTA per row: every pooled element counts, by binary search in the two sorted
//             segments, #{. <= x} and #{. < x} (nm_deep.cuh) -> KS numerator, rank sum, tie term;
//             two-pass fp64 moments; the same fp64 tails as every other tier
#include <cub/device/device_segmented_sort.cuh>
#include <cuda_runtime.h>
#include <stdint.h>

#include "nm_device.cuh"
#include "nm_huge.cuh"

namespace {

inline size_t nm_al256(size_t x) { return (x + 255) & ~(size_t)255; }

// one block: pick the huge rows out of the deep list, prefix-sum their sizes
__global__ void __launch_bounds__(256) nm_huge_collect(const nm_kargs a, int32_t* __restrict__ rows, long long* __restrict__ offA,
                                                       long long* __restrict__ offB, int n_huge) {
  __shared__ int cursor;
  if (threadIdx.x == 0) cursor = 0;
  __syncthreads();
  for (int k = threadIdx.x; k < a.n_deep; k += 256) {
    const int32_t r = a.deep_rows[k];
    if (nm_deep_p2(a.row_n0[r]) + nm_deep_p2(a.row_n1[r]) > NM_DEEP_TIER_MAX_POOLED) {
      const int slot = atomicAdd(&cursor, 1);
      if (slot < n_huge) rows[slot] = r;
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    // order by row index (the atomic order is arbitrary; n_huge is small) and build the offsets
    for (int i = 1; i < n_huge; ++i) {
      const int32_t v = rows[i];
      int j = i - 1;
      while (j >= 0 && rows[j] > v) {
        rows[j + 1] = rows[j];
        --j;
      }
      rows[j + 1] = v;
    }
    long long sa = 0, sb = 0;
    for (int i = 0; i < n_huge; ++i) {
      offA[i] = sa;
      offB[i] = sb;
      sa += a.row_n0[rows[i]];
      sb += a.row_n1[rows[i]];
    }
    offA[n_huge] = sa;
    offB[n_huge] = sb;
  }
}

__global__ void __launch_bounds__(256) nm_huge_copy(const nm_kargs a, const int32_t* __restrict__ rows,
                                                    const long long* __restrict__ offA, const long long* __restrict__ offB,
                                                    float* __restrict__ va, float* __restrict__ vb) {
  const int32_t r = rows[blockIdx.y];
  const int32_t src = a.row_pos_index[r];
  const int n0 = a.row_n0[r], n1 = a.row_n1[r];
  const float* g0 = a.vals0 + a.off0[src];
  const float* g1 = a.vals1 + a.off1[src];
  for (long long k = (long long)blockIdx.x * 256 + threadIdx.x; k < n0; k += (long long)gridDim.x * 256)
    va[offA[blockIdx.y] + k] = g0[k] + 0.0f;  // -0.0 -> +0.0: the two must tie, whatever order the sort leaves them in
  for (long long k = (long long)blockIdx.x * 256 + threadIdx.x; k < n1; k += (long long)gridDim.x * 256)
    vb[offB[blockIdx.y] + k] = g1[k] + 0.0f;
}

__device__ __forceinline__ double nm_huge_block_sum(double v, double* red) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  v = nm_warp_sum_d(v);
  __syncthreads();
  if (lane == 0) red[wid] = v;
  __syncthreads();
  double t = 0.0;
  for (int w = 0; w < 8; ++w) t += red[w];
  return t;
}

__global__ void __launch_bounds__(256) nm_huge_count(const nm_kargs a, const int32_t* __restrict__ rows,
                                                     const long long* __restrict__ offA, const long long* __restrict__ offB,
                                                     const float* __restrict__ va, const float* __restrict__ vb, int want_u,
                                                     int want_t, int want_m) {
  __shared__ double red_d[8];
  __shared__ long long red_l[3][8];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int32_t r = rows[blockIdx.x];
  const int n0 = a.row_n0[r], n1 = a.row_n1[r];
  const float* sa = va + offA[blockIdx.x];
  const float* sb = vb + offB[blockIdx.x];
  double mean[2] = {0.0, 0.0}, var[2] = {0.0, 0.0};
  if (want_m) {
    for (int g = 0; g < 2; ++g) {
      const float* s = g ? sb : sa;
      const int n = g ? n1 : n0;
      double part = 0.0;
      for (int k = tid; k < n; k += 256) part += (double)s[k];
      const double m = nm_huge_block_sum(part, red_d) / (double)n;
      part = 0.0;
      for (int k = tid; k < n; k += 256) {
        const double d = (double)s[k] - m;
        part += d * d;
      }
      mean[g] = m;
      var[g] = nm_huge_block_sum(part, red_d) / (double)(n - 1);
    }
  }
  nm_deep_acc acc;
  nm_deep_acc_init(&acc);
  const long long T = (long long)n0 + n1;
  for (long long e = tid; e < T; e += 256) {
    nm_deep_acc one;
    nm_deep_acc_init(&one);
    nm_deep_element(sa, n0, sb, n1, (int)e, want_u != 0, &one);
    nm_deep_acc_merge(&acc, one);
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
    for (int w = 0; w < 8; ++w) {
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
    // the int32 numerator column cannot hold n0 * n1 beyond 2^31 - 1: it saturates; D itself is exact
    if (tot.dnum > 0x7fffffffLL) o.dnum = 0x7fffffff;
    nm_store_row(a, r, o, want_u != 0, want_t != 0);
    if (want_m && a.acc_mom) reinterpret_cast<double4*>(a.acc_mom)[r] = make_double4(mean[0], var[0], mean[1], var[1]);
  }
}

}  // namespace

size_t nm_huge_scratch_bytes(int n_huge, long long nv0, long long nv1) {
  size_t cub_a = 0, cub_b = 0;
  cub::DeviceSegmentedSort::SortKeys(nullptr, cub_a, (const float*)nullptr, (float*)nullptr, (long long)nv0, n_huge,
                                     (const long long*)nullptr, (const long long*)nullptr);
  cub::DeviceSegmentedSort::SortKeys(nullptr, cub_b, (const float*)nullptr, (float*)nullptr, (long long)nv1, n_huge,
                                     (const long long*)nullptr, (const long long*)nullptr);
  const size_t cub_bytes = cub_a > cub_b ? cub_a : cub_b;
  return nm_al256(sizeof(int32_t) * (size_t)n_huge) + 2 * nm_al256(sizeof(long long) * (size_t)(n_huge + 1)) +
         2 * nm_al256(sizeof(float) * (size_t)nv0) + 2 * nm_al256(sizeof(float) * (size_t)nv1) + nm_al256(cub_bytes) + 256;
}

int nm_huge_run(const nm_kargs& ka, bool want_u, bool want_t, bool want_m, int n_huge, long long nv0, long long nv1,
                void* scratch, int* launches, cudaStream_t st) {
  unsigned char* p = (unsigned char*)scratch;
  int32_t* rows = (int32_t*)p;
  p += nm_al256(sizeof(int32_t) * (size_t)n_huge);
  long long* offA = (long long*)p;
  p += nm_al256(sizeof(long long) * (size_t)(n_huge + 1));
  long long* offB = (long long*)p;
  p += nm_al256(sizeof(long long) * (size_t)(n_huge + 1));
  float* a_in = (float*)p;
  p += nm_al256(sizeof(float) * (size_t)nv0);
  float* a_out = (float*)p;
  p += nm_al256(sizeof(float) * (size_t)nv0);
  float* b_in = (float*)p;
  p += nm_al256(sizeof(float) * (size_t)nv1);
  float* b_out = (float*)p;
  p += nm_al256(sizeof(float) * (size_t)nv1);
  size_t cub_a = 0, cub_b = 0;
  cub::DeviceSegmentedSort::SortKeys(nullptr, cub_a, (const float*)nullptr, (float*)nullptr, nv0, n_huge,
                                     (const long long*)nullptr, (const long long*)nullptr);
  cub::DeviceSegmentedSort::SortKeys(nullptr, cub_b, (const float*)nullptr, (float*)nullptr, nv1, n_huge,
                                     (const long long*)nullptr, (const long long*)nullptr);
  nm_huge_collect<<<1, 256, 0, st>>>(ka, rows, offA, offB, n_huge);
  nm_huge_copy<<<dim3(64, (unsigned)n_huge), 256, 0, st>>>(ka, rows, offA, offB, a_in, b_in);
  cudaError_t e = cub::DeviceSegmentedSort::SortKeys((void*)p, cub_a, (const float*)a_in, a_out, nv0, n_huge, offA, offA + 1, st);
  if (e != cudaSuccess) return (int)e;
  e = cub::DeviceSegmentedSort::SortKeys((void*)p, cub_b, (const float*)b_in, b_out, nv1, n_huge, offB, offB + 1, st);
  if (e != cudaSuccess) return (int)e;
  nm_huge_count<<<(unsigned)n_huge, 256, 0, st>>>(ka, rows, offA, offB, a_out, b_out, want_u ? 1 : 0, want_t ? 1 : 0,
                                                   want_m ? 1 : 0);
  *launches += 5;
  return (int)cudaGetLastError();
}
