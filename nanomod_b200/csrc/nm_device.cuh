// nm_device.cuh -- device-only helpers shared by the kernels: mbarrier / TMA bulk-copy PTX
// wrappers, warp reductions, the kernel argument block and the row store.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/nanomod_b200.h"
#include "nm_deep.cuh"
#include "nm_lane.cuh"

#define NM_INF __int_as_float(0x7f800000)

__device__ __forceinline__ uint32_t nm_smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void nm_mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(nm_smem_u32(bar)), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void nm_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(nm_smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
// raise the phase's pending byte count without arriving (a copy issued ahead of the phase's arrive)
__device__ __forceinline__ void nm_mbar_expect_tx_only(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], %1;" ::"r"(nm_smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void nm_mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done = 0;
  const uint32_t addr = nm_smem_u32(bar);
  while (!done) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
  }
}
// TMA bulk copy global -> shared (1-D): 16-byte aligned src/dst, size a multiple of 16.
__device__ __forceinline__ void nm_bulk_g2s(void* dst, const void* src, uint32_t bytes,
                                            uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
          "r"(nm_smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(nm_smem_u32(bar))
      : "memory");
}

// Per-thread asynchronous 16-byte copies global -> shared (LDGSTS): the scattered staging path
// of the lane tier (many small TMA bulk copies have a low issue rate; measured 2x slower).
__device__ __forceinline__ void nm_cp_async16(void* dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(nm_smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void nm_cp_async8(void* dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(nm_smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void nm_cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void nm_cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// L2 prefetch of a contiguous global range (no shared memory involved)
__device__ __forceinline__ void nm_prefetch_l2(const void* src, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}

__device__ __forceinline__ void nm_prefetch_l2_line(const void* src) {
  asm volatile("prefetch.global.L2 [%0];" ::"l"(src) : "memory");
}

__device__ __forceinline__ long long nm_warp_min_ll(long long v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const long long t = __shfl_xor_sync(0xffffffffu, v, o);
    v = t < v ? t : v;
  }
  return v;
}
__device__ __forceinline__ long long nm_warp_max_ll(long long v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const long long t = __shfl_xor_sync(0xffffffffu, v, o);
    v = t > v ? t : v;
  }
  return v;
}
__device__ __forceinline__ long long nm_warp_sum_ll(long long v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double nm_warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

#define NM_SPREAD 32
struct nm_summary {
  unsigned long long n_rows;
  int max_lane_n;   // max over lane-tier rows of max(n0,n1)
  int n_deep;       // rows with max(n0,n1) > NM_LANE_TIER_MAX
  int max_deep_p2;  // max over deep rows of pow2ceil(n0)+pow2ceil(n1)
  int deep_cursor;
  int tile_cursor[3];   // lane-tier work queues (tiles beyond the first wave), one per launch
  int ds_cursor;        // down-sampling work queue (rows, in chunks of 32)
  int ds_too_deep;      // plan pass: bit 0 a row to be down-sampled has more than NM_DS_MAX_READS reads (the block kernel takes it),
                        // bit 1 one is beyond that kernel too; the kernels raise it when they meet a row they cannot take
  int max_lane_slack;   // max over lane-tier rows of NM_LANE_TIER_MAX - max(n0,n1)  (-> shortest row)
  int n_le64, n_le104;  // class-binned calls: lane-tier rows whose network class is <= 64 / <= 104
  int n_filtered;       // candidates dropped by the coverage filter (0 => rows == candidates)
  int bad_input;        // a candidate with a negative read count (offsets not monotonic) / segment id out of range
  int dense_retry;      // set by nm_lane_dense_kernel: the call does not have the shape the launch assumed
  int dense_tile_cursor;
  // grid-key launch of the dense path (the seven ints from dense_retry on are zeroed together before a re-run)
  int retry_cursor;     // work queue of the float32 launch that follows it
  int retry_count;      // tiles on the retry list
  int grid_giveup;      // a warp found nothing but off-grid tiles: everybody stops claiming
  int grid_n_warps;     // warps of the grid-key launch: tiles >= grid_n_warps + dense_tile_cursor were never claimed
  int grid_tiles;       // tiles the grid-key launch computed (nm_last_grid_tiles)
  int max_deep_t;          // max over deep rows of n0 + n1
  int deep_fallback_count; // (reserved: a device-side count for nm_deep_kernel's row list)
  // {kept candidates, lane-tier candidates of network class <= 64, <= 104, -} counted by nm_plan_count;
  // every block adds to entry (blockIdx & 31): tens of thousands of atomics on ONE address serialise in L2
  int spread[32][4];
  int n_huge;                 // deep rows beyond the shared-memory deep tier (nm_huge.cu takes them)
  int ds_deep_cursor;         // work queue of nm_downsample_deep_kernel (rows, in chunks of 256)
  int head_cursor;            // candidates the combine kernel listed for an armed head selection (nm_rank.cuh)
  int head_fail;              // that list could not give the head (too few / too many candidates): select the ordinary way
  int head_cut;               // exponent bin at which that selection reached `want` rows (0: none)
  unsigned long long huge_v0, huge_v1;  // their values in group 0 / 1
};

// The shape the dense lane kernel takes: rows == candidates, nothing deep, and the general path
// would not split the call into one launch per size group (it does when the groups are not all the
// same and one that is not the longest holds >= 7/8 of the rows: outliers must not make everybody
// pay for their network).  Host (launch decision) and device (validation of a speculative launch).
__host__ __device__ __forceinline__ bool nm_dense_shape_ok(const nm_summary& s) {
  if (s.n_filtered != 0 || s.n_deep != 0 || s.bad_input != 0 || s.max_lane_n <= 0) return false;
  const int gmin = nm_lane_group(NM_LANE_TIER_MAX - s.max_lane_slack), gmax = nm_lane_group(s.max_lane_n);
  if (gmin == gmax) return true;
  long long total = 0, le64 = 0, le104 = 0;  // the counters are spread over 32 addresses (see nm_plan_count)
  for (int k = 0; k < NM_SPREAD; ++k) {
    total += s.spread[k][0];
    le64 += s.spread[k][1];
    le104 += s.spread[k][2];
  }
  const long long g0 = le64, g1 = le104 - le64, g2 = total - le104;
  const long long biggest = g0 > g1 ? (g0 > g2 ? g0 : g2) : (g1 > g2 ? g1 : g2);
  const long long of_gmax = gmax == 2 ? g2 : gmax == 1 ? g1 : g0;
  return !(biggest * 8 >= total * 7 && biggest != of_gmax);
}


__device__ __forceinline__ int nm_pow2ceil(int n) {
  int p = 1;
  while (p < n) p <<= 1;
  return p;
}


struct nm_kargs {
  const float* vals0;
  const float* vals1;
  const int64_t* off0;
  const int64_t* off1;
  const int32_t* row_pos_index;
  const int32_t* row_n0;
  const int32_t* row_n1;
  int64_t n_rows;
  // lane tier: the launch works on rows perm[row_lo .. row_hi) (perm == NULL: rows row_lo .. row_hi
  // themselves).  Calls with mixed coverage sort the rows by network class first (nm_class_sort)
  // and launch once per class group, so that every tile is homogeneous.
  const int32_t* perm;
  int64_t row_lo, row_hi;
  int gaps;           // the launch's rows are not all consecutive candidates (filtered / deep / other group rows between)
  int region_floats;  // floats per group region in shared memory (lane tier)
  int class_n;        // size class of the call's longest lane-tier row
  int one, mone;      // runtime 1 / -1: keeps the IMAD form of the integer compare-exchange
  // dense path, grid keys (nm_lane.cuh): the grid-key launch lists the tiles it could not take in retry_tiles
  // (count: sum->retry_count); the float32 launch after it (retry_mode = 1) works that list and the unclaimed tail
  int grid_tries;        // tiles a warp may see fail the check, with none passing, before it raises sum->grid_giveup
  int retry_mode;
  int32_t* retry_tiles;  // [number of tiles of the call]
  int* tile_cursor;   // device counter, zero at launch
  // rank sums / moments of the lane tier; their fp64 tails (normal and Student-t
  // tails: a lot of cold fp64 code) run afterwards in nm_tails_kernel, not inside the sort loop
  int* acc_r2;        // want_u: 2 * rank sum of group 0
  int* acc_tie;       // want_u: sum over tie groups of t^3 - t
  double* acc_mom;    // want_t: mean0, var0, mean1, var1 per row
  int32_t* ks_dnum;
  double* ks_d;
  double* ks_p;
  int64_t* two_u;
  double* u_stat;
  double* u_p;
  double* t_stat;
  double* t_p;
  uint8_t* flags;
  const int32_t* deep_rows;
  int n_deep;
  const int* deep_count_ptr;  // when set: only the first *deep_count_ptr entries of deep_rows are rows (device-side count)
  int32_t* deep_retry_rows;   // deep tier, grid keys: rows the key-pair kernel could not take (-> float32 kernel), or NULL = float32 only
  int* deep_retry_count;
  // dense path (nm_lane_dense_kernel): rows == candidates, no deep rows.  The kernel validates
  // that against the plan summary on the device, writes the row index / coverage columns itself
  // and leaves norm.isf(p) / ln p of the KS p-value for the combine stencil.
  nm_summary* sum;
  int64_t n_pos;
  int32_t* w_row_pos_index;
  int32_t* w_n0;
  int32_t* w_n1;
  double* comb_z;   // norm.isf(ks_p) per row (Stouffer), or NULL
  double* comb_ln;  // ln(ks_p) per row (Fisher), or NULL
};


__device__ __forceinline__ void nm_store_row(const nm_kargs& a, int64_t r, const nm_row_out& o,
                                             bool want_u, bool want_t) {
  a.ks_dnum[r] = o.dnum;
  if (a.ks_d) a.ks_d[r] = o.ks_d;
  a.ks_p[r] = o.ks_p;
  if (want_u) {
    a.two_u[r] = o.two_u;
    if (a.u_stat) a.u_stat[r] = o.u_stat;
    a.u_p[r] = o.u_p;
  }
  if (want_t) {
    a.t_stat[r] = o.t_stat;
    a.t_p[r] = o.t_p;
  }
  if (a.flags) a.flags[r] = (uint8_t)o.flags;
}


// sum of four accumulators that were filled by window slot (e & 3), combined in the order of the
// ROW index class ((e - shift) & 3): the result does not depend on the row's alignment
__device__ __forceinline__ double nm_sum4_by_row_class(const double (&s)[4], int shift) {
  const double c0 = shift == 0 ? s[0] : shift == 1 ? s[1] : shift == 2 ? s[2] : s[3];
  const double c1 = shift == 0 ? s[1] : shift == 1 ? s[2] : shift == 2 ? s[3] : s[0];
  const double c2 = shift == 0 ? s[2] : shift == 1 ? s[3] : shift == 2 ? s[0] : s[1];
  const double c3 = shift == 0 ? s[3] : shift == 1 ? s[0] : shift == 2 ? s[1] : s[2];
  return __dadd_rn(__dadd_rn(c0, c1), __dadd_rn(c2, c3));
}

// Welch moments of one row: one pass, fp64, four accumulators by row index mod 4, read through
// the same aligned 128-bit window as the sort.  A ROLLED loop with explicitly rounded
// operations: the instruction sequence applied to a row depends only on the row itself (not on
// its alignment, its tile or the tile's network size), so results are bit-identical however
// the genome is sharded -- and the loop body stays in the instruction cache.
__device__ __forceinline__ void nm_lane_moments(const float* region, int base, int n, double* mean,
                                                double* var) {
  const int shift = base & 3;
  const float4* raw4 = reinterpret_cast<const float4*>(region + (base - shift));
  const int nq = __reduce_max_sync(0xffffffffu, (shift + n + 3) >> 2);
  // One pass over the row with the data shifted by its first value K ("shifted data" variance):
  // sum d and sum d^2 of d = x - K, then mean = K + S/n and var = (SS - S^2/n)/(n-1).  K is a
  // sample of the row, so the final subtraction cancels at most a few bits; one float->double
  // conversion per value instead of two (F2F runs at a quarter of the fp64 rate and was the
  // bottleneck of the two-pass form).
  const double K = n > 0 ? (double)region[base] : 0.0;
  double s[4] = {0.0, 0.0, 0.0, 0.0}, ss[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll 2
  for (int q = 0; q < nq; ++q) {
    const float4 v4 = raw4[q];
    const float vv[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const bool valid = (unsigned)(4 * q + j - shift) < (unsigned)n;
      const double d = valid ? __dsub_rn((double)vv[j], K) : 0.0;
      s[j] = __dadd_rn(s[j], d);
      ss[j] = __fma_rn(d, d, ss[j]);
    }
  }
  const double S = nm_sum4_by_row_class(s, shift), SS = nm_sum4_by_row_class(ss, shift);
  const double dn = (double)n;
  const double sn = __ddiv_rn(S, dn);
  *mean = __dadd_rn(K, sn);
  const double num = __fma_rn(-S, sn, SS);
  *var = __ddiv_rn(num > 0.0 ? num : 0.0, (double)(n - 1));
}

// deep tier: each group is sorted as a power-of-two array of at least NM_DEEP_MIN_P elements
#define NM_DEEP_THREADS 256
#define NM_DEEP_MIN_P 512
__host__ __device__ __forceinline__ int nm_deep_p2(int n) {
  int p = NM_DEEP_MIN_P;
  while (p < n) p <<= 1;
  return p;
}
// host-side launcher of the deep tier (nm_deep_kernel.cu)
int nm_launch_deep(const nm_kargs& ka, bool want_u, bool want_t, bool want_m, int n_deep, int max_p2, int smem_bytes,
                   int sm_count, cudaStream_t st);


// host-side launcher of the lane tier (nm_lane_kernel.cu); returns a cudaError_t as int
int nm_launch_lane(const nm_kargs& ka, bool want_u, bool want_t, int max_n, int sm_count, cudaStream_t st);
// dense variant: rows == candidates (nothing filtered, nothing deep); class_n = network class of the launch
int nm_launch_lane_dense(const nm_kargs& ka, bool want_u, bool want_t, int class_n, int sm_count, bool grid_keys,
                         cudaStream_t st);
