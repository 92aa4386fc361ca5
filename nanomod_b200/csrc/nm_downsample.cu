// nanomod_b200 -- the down-sampling branch of getKStest.
//
// Reference: bin/scripts/myDetect.py:339-361.  With cov = coverages[strand] > 0, a position
// where a group has more than cov reads gets its KS result from `downsampling` (100) resamples:
// each group with more than cov reads is replaced by np.random.choice(group, cov) (WITH
// replacement), ks_2samp is run on the pair, and the reported (D, p) is the pair at index
// int(downsampling * downsampling_quantile) of the p-values sorted ascending.  U and t are not
// affected.  The reference draws from numpy's unseeded global generator, so its numbers cannot
// be reproduced; this implementation defines its own counter-based stream (below), which
// the test suite's CPU checker restates bit for bit, and the procedure is validated statistically
// against the reference's (tests/test_oracle.py, tests/test_gpu_parity.py).
//
// Random stream.  Philox4x32-10, key = (seed_lo, seed_hi).  Draw t (0-based) of resample i of
// group g at the candidate (seg, pos) is word (t & 3) of the block with counter
// (t >> 2, i | g << 16, pos, seg); a 32-bit word r selects index (r * n) >> 32 of the group's
// values IN ASCENDING ORDER (equal values are interchangeable, so tie order does not matter).
//
// One warp per qualifying row, lanes = resamples (32 per round):
//   1. sort both groups (warp bitonic sort in shared memory), 2. merge them into a "script"
//   (pooled rank -> group, index, end-of-tie-group flag), 3. per round: every lane zeroes its
//   column of draw counters, draws, then walks the script accumulating its two resampled ECDF
//   numerators and max |c0*m1 - c1*m0| at the tie-group ends, 4. the reported numerator is the
//   (ds_index)-th largest of the resamples' (p is decreasing in D for fixed sizes).
#include <cuda_runtime.h>
#include <stdint.h>

#include "nm_device.cuh"
#include "nm_downsample.cuh"

namespace {

constexpr int kWarps = 4;  // rows in flight per CTA

__device__ __forceinline__ void nm_philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                                 uint32_t k1, uint32_t (&out)[4]) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// ascending bitonic sort of s[0..P) by one warp (P a power of two >= 32; pads are +inf)
__device__ __forceinline__ void nm_warp_bitonic(float* s, int P, int lane) {
  for (int k = 2; k <= P; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int t = lane; t < (P >> 1); t += 32) {
        const int lo = ((t & ~(j - 1)) << 1) | (t & (j - 1));  // index with bit j clear
        const int hi = lo | j;
        const bool up = (lo & k) == 0;
        const float a = s[lo], b = s[hi];
        if ((a > b) == up) {
          s[lo] = b;
          s[hi] = a;
        }
      }
      __syncwarp();
    }
  }
}

struct nm_ds_args {
  nm_kargs k;
  const int32_t* pos;
  const int32_t* seg;
  const int32_t* seg_cov;
  int times, index;
  uint32_t seed_lo, seed_hi;
  int* cursor;
  int* too_deep;
};

__global__ void __launch_bounds__(32 * kWarps) nm_downsample_kernel(const nm_ds_args a) {
  extern __shared__ __align__(16) unsigned char nm_smem[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  // per-warp shared memory: sa[256] sb[256] pv[512] (floats) | script[512] (u16) | dres[128]
  // (i32) | cnt[512][32] (u8)
  unsigned char* my = nm_smem + (size_t)wib * NM_DS_WARP_SMEM;
  float* sa = reinterpret_cast<float*>(my);
  float* sb = sa + NM_DS_MAX_N;
  float* pv = sb + NM_DS_MAX_N;
  uint16_t* script = reinterpret_cast<uint16_t*>(pv + 2 * NM_DS_MAX_N);
  int* dres = reinterpret_cast<int*>(script + 2 * NM_DS_MAX_N);
  uint8_t* cnt = reinterpret_cast<uint8_t*>(dres + NM_DS_MAX_TIMES);

  while (true) {
    // rows are handed out in chunks of 32: lane l of the warp tests row base + l
    long long base = 0;
    if (lane == 0) base = (long long)atomicAdd(a.cursor, 32);
    base = __shfl_sync(0xffffffffu, base, 0);
    if (base >= a.k.n_rows) break;
    const int64_t rr = base + lane;
    bool need = false;
    if (rr < a.k.n_rows) {
      const int32_t src = a.k.row_pos_index[rr];
      const int cov = a.seg_cov[a.seg[src]];
      need = cov > 0 && (a.k.row_n0[rr] > cov || a.k.row_n1[rr] > cov);
      if (need && (a.k.row_n0[rr] > NM_DS_MAX_N || a.k.row_n1[rr] > NM_DS_MAX_N)) need = false;  // nm_downsample_deep_kernel's
    }
    unsigned todo = __ballot_sync(0xffffffffu, need);
    while (todo) {
      const int l = __ffs(todo) - 1;
      todo &= todo - 1;
      const int64_t r = base + l;
      const int32_t src = a.k.row_pos_index[r];
      const int n0 = a.k.row_n0[r], n1 = a.k.row_n1[r];
      const int cov = a.seg_cov[a.seg[src]];
      const uint32_t cpos = (uint32_t)a.pos[src], cseg = (uint32_t)a.seg[src];
      const bool ds0 = n0 > cov, ds1 = n1 > cov;
      const int m0 = ds0 ? cov : n0, m1 = ds1 ? cov : n1;
      const int T = n0 + n1;

      // ---- 1. sort both groups
      int P0 = 32, P1 = 32;
      while (P0 < n0) P0 <<= 1;
      while (P1 < n1) P1 <<= 1;
      const float* g0 = a.k.vals0 + a.k.off0[src];
      const float* g1 = a.k.vals1 + a.k.off1[src];
      for (int t = lane; t < P0; t += 32) sa[t] = t < n0 ? g0[t] + 0.0f : INFINITY;
      for (int t = lane; t < P1; t += 32) sb[t] = t < n1 ? g1[t] + 0.0f : INFINITY;
      __syncwarp();
      nm_warp_bitonic(sa, P0, lane);
      nm_warp_bitonic(sb, P1, lane);

      // ---- 2. merge script: pooled rank of a[i] = i + #(b < a[i]); of b[j] = j + #(a <= b[j])
      for (int t = lane; t < T; t += 32) {
        const bool isb = t >= n0;
        const int idx = isb ? t - n0 : t;
        const float v = isb ? sb[idx] : sa[idx];
        const float* o = isb ? sa : sb;
        int lo = 0, hi = isb ? n0 : n1;
        while (lo < hi) {
          const int mid = (lo + hi) >> 1;
          const bool right = isb ? (o[mid] <= v) : (o[mid] < v);
          if (right) lo = mid + 1; else hi = mid;
        }
        const int rank = idx + lo;
        script[rank] = (uint16_t)(idx | (isb ? 0x8000 : 0));
        pv[rank] = v;
      }
      __syncwarp();
      for (int t = lane; t < T; t += 32)
        if (t == T - 1 || pv[t] < pv[t + 1]) script[t] |= 0x4000;
      __syncwarp();

      // ---- 3. resamples
      for (int round = 0; round * 32 < a.times; ++round) {
        const int i = round * 32 + lane;
        uint32_t* cw = reinterpret_cast<uint32_t*>(cnt);
        for (int t = lane; t < T * 8; t += 32) cw[t] = 0u;  // T counters x 32 lanes bytes
        __syncwarp();
        if (i < a.times) {
#pragma unroll 1
          for (int g = 0; g < 2; ++g) {
            if (!(g ? ds1 : ds0)) continue;
            const int n = g ? n1 : n0, m = g ? m1 : m0;
            uint8_t* col = cnt + (size_t)(g ? n0 : 0) * 32 + lane;
            for (int q = 0; 4 * q < m; ++q) {
              uint32_t w[4];
              nm_philox4x32_10((uint32_t)q, (uint32_t)i | ((uint32_t)g << 16), cpos, cseg, a.seed_lo, a.seed_hi, w);
#pragma unroll
              for (int e = 0; e < 4; ++e)
                if (4 * q + e < m) {
                  const int idx = (int)(((unsigned long long)w[e] * (unsigned long long)n) >> 32);
                  col[idx * 32] += 1;
                }
            }
          }
          int c0 = 0, c1 = 0, dmax = 0;
          for (int t = 0; t < T; ++t) {
            const int s = script[t];
            const int idx = s & 0x3fff;
            if (s & 0x8000)
              c1 += ds1 ? (int)cnt[(size_t)(n0 + idx) * 32 + lane] : 1;
            else
              c0 += ds0 ? (int)cnt[(size_t)idx * 32 + lane] : 1;
            if (s & 0x4000) {
              int d = c0 * m1 - c1 * m0;
              d = d < 0 ? -d : d;
              dmax = d > dmax ? d : dmax;
            }
          }
          dres[i] = dmax;
        }
        __syncwarp();
      }

      // ---- 4. the resample at sorted-p index `index` == the index-th largest numerator
      int sel = -1;
      for (int i = lane; i < a.times; i += 32) {
        const int v = dres[i];
        int greater = 0, equal = 0;
        for (int t = 0; t < a.times; ++t) {
          const int o = dres[t];
          greater += o > v;
          equal += o == v;
        }
        if (greater <= a.index && a.index < greater + equal) sel = v;
      }
      sel = __reduce_max_sync(0xffffffffu, sel);
      if (lane == 0) {
        double d, p;
        nm_ks_tail(sel, m0, m1, &d, &p);
        a.k.ks_dnum[r] = sel;
        if (a.k.ks_d) a.k.ks_d[r] = d;
        a.k.ks_p[r] = p;
      }
      __syncwarp();
    }
  }
}


// ------------------------------------------------------------------------------------------
// Rows with more than NM_DS_MAX_N reads in a group: one CTA per row.
// ------------------------------------------------------------------------------------------
constexpr int kDeepThreads = 256;
constexpr int kDeepWarps = kDeepThreads / 32;

// ascending bitonic sort of s[0..P) by the CTA (P a power of two >= 2 * kDeepThreads; pads are +inf)
__device__ __forceinline__ void nm_block_bitonic(float* s, int P) {
  for (int k = 2; k <= P; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int t = threadIdx.x; t < (P >> 1); t += kDeepThreads) {
        const int lo = ((t & ~(j - 1)) << 1) | (t & (j - 1));
        const int hi = lo | j;
        const bool up = (lo & k) == 0;
        const float x = s[lo], y = s[hi];
        if ((x > y) == up) {
          s[lo] = y;
          s[hi] = x;
        }
      }
      __syncthreads();
    }
  }
}

// number of elements <= v in the ascending array s[0..n)
__device__ __forceinline__ int nm_upper_bound(const float* s, int n, float v) {
  int lo = 0, hi = n;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (s[mid] <= v) lo = mid + 1; else hi = mid;
  }
  return lo;
}

__global__ void __launch_bounds__(kDeepThreads) nm_downsample_deep_kernel(const nm_ds_args a, float* __restrict__ scratch) {
  extern __shared__ __align__(16) unsigned char nm_smem[];
  float* sm = reinterpret_cast<float*>(nm_smem);  // the group being sorted | later: per-warp samples
  __shared__ int s_rows[kDeepThreads];
  __shared__ int s_n;
  __shared__ long long s_base;
  __shared__ int dres[NM_DS_MAX_TIMES];
  const int tid = threadIdx.x, lane = tid & 31, wib = tid >> 5;
  float* const srt0 = scratch + (size_t)blockIdx.x * 2 * NM_DS_DEEP_MAX_READS;
  float* const srt1 = srt0 + NM_DS_DEEP_MAX_READS;

  while (true) {
    __syncthreads();
    if (tid == 0) {
      s_base = (long long)atomicAdd(a.cursor, kDeepThreads);
      s_n = 0;
    }
    __syncthreads();
    const long long base = s_base;
    if (base >= a.k.n_rows) break;
    {
      const int64_t rr = base + tid;
      if (rr < a.k.n_rows) {
        const int32_t src = a.k.row_pos_index[rr];
        const int cov = a.seg_cov[a.seg[src]];
        const int n0 = a.k.row_n0[rr], n1 = a.k.row_n1[rr];
        if (cov > 0 && (n0 > cov || n1 > cov) && (n0 > NM_DS_MAX_N || n1 > NM_DS_MAX_N)) {
          if (cov > NM_DS_DEEP_MAX_COV || (n0 > cov && n0 > NM_DS_DEEP_MAX_READS) || (n1 > cov && n1 > NM_DS_DEEP_MAX_READS))
            atomicOr(a.too_deep, 2);  // reported by the host as NM_ERR_TOO_DEEP
          else
            s_rows[atomicAdd(&s_n, 1)] = tid;
        }
      }
    }
    __syncthreads();
    const int n_list = s_n;
    for (int q = 0; q < n_list; ++q) {
      const int64_t r = base + s_rows[q];
      const int32_t src = a.k.row_pos_index[r];
      const int n0 = a.k.row_n0[r], n1 = a.k.row_n1[r];
      const int cov = a.seg_cov[a.seg[src]];
      const uint32_t cpos = (uint32_t)a.pos[src], cseg = (uint32_t)a.seg[src];
      const bool ds0 = n0 > cov, ds1 = n1 > cov;
      const int m0 = ds0 ? cov : n0, m1 = ds1 ? cov : n1;

      // ---- 1. a down-sampled group is sorted (draws index its ascending values); the other is taken as it is
#pragma unroll 1
      for (int g = 0; g < 2; ++g) {
        const int n = g ? n1 : n0;
        const float* gv = g ? a.k.vals1 + a.k.off1[src] : a.k.vals0 + a.k.off0[src];
        float* dst = g ? srt1 : srt0;
        if (g ? ds1 : ds0) {
          int P = 2 * kDeepThreads;
          while (P < n) P <<= 1;
          __syncthreads();
          for (int t = tid; t < P; t += kDeepThreads) sm[t] = t < n ? gv[t] + 0.0f : INFINITY;
          __syncthreads();
          nm_block_bitonic(sm, P);
          for (int t = tid; t < n; t += kDeepThreads) dst[t] = sm[t];
        } else {
          for (int t = tid; t < n; t += kDeepThreads) dst[t] = gv[t] + 0.0f;
        }
      }
      __syncthreads();  // the sorted groups (global, written by this CTA) and the end of the use of sm as sort buffer

      // ---- 2. a warp per resample: gather the draws, sort the two samples, rank every sampled value in both
      float* S0 = sm + (size_t)wib * 2 * NM_DS_DEEP_MAX_COV;
      float* S1 = S0 + NM_DS_DEEP_MAX_COV;
      int P0 = 32, P1 = 32;
      while (P0 < m0) P0 <<= 1;
      while (P1 < m1) P1 <<= 1;
      for (int i = wib; i < a.times; i += kDeepWarps) {
#pragma unroll 1
        for (int g = 0; g < 2; ++g) {
          const int n = g ? n1 : n0, m = g ? m1 : m0, P = g ? P1 : P0;
          const float* sg = g ? srt1 : srt0;
          float* S = g ? S1 : S0;
          if (g ? ds1 : ds0) {
            for (int t = lane; t < P; t += 32) {
              float v = INFINITY;
              if (t < m) {
                uint32_t w[4];
                nm_philox4x32_10((uint32_t)(t >> 2), (uint32_t)i | ((uint32_t)g << 16), cpos, cseg, a.seed_lo, a.seed_hi, w);
                const uint32_t word = (t & 3) == 0 ? w[0] : (t & 3) == 1 ? w[1] : (t & 3) == 2 ? w[2] : w[3];
                v = sg[(int)(((unsigned long long)word * (unsigned long long)n) >> 32)];
              }
              S[t] = v;
            }
          } else {
            for (int t = lane; t < P; t += 32) S[t] = t < m ? sg[t] : INFINITY;
          }
        }
        __syncwarp();
        nm_warp_bitonic(S0, P0, lane);
        nm_warp_bitonic(S1, P1, lane);
        int dmax = 0;
        for (int k = lane; k < m0 + m1; k += 32) {
          const float v = k < m0 ? S0[k] : S1[k - m0];
          int d = nm_upper_bound(S0, m0, v) * m1 - nm_upper_bound(S1, m1, v) * m0;
          d = d < 0 ? -d : d;
          dmax = d > dmax ? d : dmax;
        }
        dmax = __reduce_max_sync(0xffffffffu, dmax);
        if (lane == 0) dres[i] = dmax;
        __syncwarp();
      }
      __syncthreads();

      // ---- 3. the resample at sorted-p index `index` == the index-th largest numerator
      if (wib == 0) {
        int sel = -1;
        for (int i = lane; i < a.times; i += 32) {
          const int v = dres[i];
          int greater = 0, equal = 0;
          for (int t = 0; t < a.times; ++t) {
            const int o = dres[t];
            greater += o > v;
            equal += o == v;
          }
          if (greater <= a.index && a.index < greater + equal) sel = v;
        }
        sel = __reduce_max_sync(0xffffffffu, sel);
        if (lane == 0) {
          double d, p;
          nm_ks_tail(sel, m0, m1, &d, &p);
          a.k.ks_dnum[r] = sel;
          if (a.k.ks_d) a.k.ks_d[r] = d;
          a.k.ks_p[r] = p;
        }
      }
      __syncthreads();
    }
  }
}

}  // namespace

int nm_launch_downsample(const nm_kargs& ka, const int32_t* pos, const int32_t* seg, const int32_t* seg_cov, int times,
                         int index, uint64_t seed, int* cursor, int* too_deep, int sm_count, cudaStream_t st) {
  nm_ds_args a;
  a.k = ka;
  a.pos = pos;
  a.seg = seg;
  a.seg_cov = seg_cov;
  a.times = times;
  a.index = index;
  a.seed_lo = (uint32_t)(seed & 0xffffffffu);
  a.seed_hi = (uint32_t)(seed >> 32);
  a.cursor = cursor;
  a.too_deep = too_deep;
  const int smem = kWarps * NM_DS_WARP_SMEM;
  cudaError_t e = cudaFuncSetAttribute(nm_downsample_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return (int)e;
  int blocks = 0;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks, nm_downsample_kernel, 32 * kWarps, (size_t)smem);
  if (e != cudaSuccess) return (int)e;
  if (blocks < 1) return (int)cudaErrorInvalidConfiguration;
  int64_t grid = (ka.n_rows + 32 * kWarps - 1) / (32 * kWarps);
  if (grid > (int64_t)blocks * sm_count) grid = (int64_t)blocks * sm_count;
  nm_downsample_kernel<<<(unsigned)grid, 32 * kWarps, (size_t)smem, st>>>(a);
  return (int)cudaGetLastError();
}

size_t nm_downsample_deep_scratch_bytes(int sm_count) {
  return sizeof(float) * 2 * (size_t)NM_DS_DEEP_MAX_READS * (size_t)(sm_count > 0 ? sm_count : 1);
}

int nm_launch_downsample_deep(const nm_kargs& ka, const int32_t* pos, const int32_t* seg, const int32_t* seg_cov, int times,
                              int index, uint64_t seed, int* cursor, int* too_deep, float* scratch, int sm_count,
                              cudaStream_t st) {
  nm_ds_args a;
  a.k = ka;
  a.pos = pos;
  a.seg = seg;
  a.seg_cov = seg_cov;
  a.times = times;
  a.index = index;
  a.seed_lo = (uint32_t)(seed & 0xffffffffu);
  a.seed_hi = (uint32_t)(seed >> 32);
  a.cursor = cursor;
  a.too_deep = too_deep;
  // the sort buffer (one group, padded to a power of two) and, afterwards, the warps' sample pairs
  const size_t smem = sizeof(float) * (size_t)(NM_DS_DEEP_MAX_READS > kDeepWarps * 2 * NM_DS_DEEP_MAX_COV ? NM_DS_DEEP_MAX_READS
                                                                                                         : kDeepWarps * 2 * NM_DS_DEEP_MAX_COV);
  cudaError_t e = cudaFuncSetAttribute(nm_downsample_deep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  int64_t grid = (ka.n_rows + kDeepThreads - 1) / kDeepThreads;
  if (grid > sm_count) grid = sm_count;
  nm_downsample_deep_kernel<<<(unsigned)grid, kDeepThreads, smem, st>>>(a, scratch);
  return (int)cudaGetLastError();
}
