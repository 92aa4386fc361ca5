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

// L2 prefetch of a contiguous global range (no shared memory involved)
__device__ __forceinline__ void nm_prefetch_l2(const void* src, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
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

struct nm_summary {
  unsigned long long n_rows;
  int max_lane_n;   // max over lane-tier rows of max(n0,n1)
  int n_deep;       // rows with max(n0,n1) > NM_LANE_TIER_MAX
  int max_deep_p2;  // max over deep rows of pow2ceil(n0)+pow2ceil(n1)
  int deep_cursor;
  int tile_cursor;  // lane-tier work queue (tiles beyond the first wave)
  int pad;
};


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
  int region_floats;  // floats per group region in shared memory (lane tier)
  int one, mone;      // runtime 1 / -1: keeps the IMAD form of the integer compare-exchange
  int* tile_cursor;   // device counter, zero at launch
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


// host-side launcher of the lane tier (nm_lane_kernel.cu); returns a cudaError_t as int
int nm_launch_lane(const nm_kargs& ka, bool want_u, bool want_t, int max_n, int sm_count,
                   cudaStream_t st);
