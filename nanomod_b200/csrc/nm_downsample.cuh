// nanomod_b200 -- internal interface of the down-sampling unit (nm_downsample.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "nm_device.cuh"

#define NM_DS_MAX_N NM_DS_MAX_READS  // reads per group a down-sampled position may have (public header)
// per-warp shared memory: two sorted groups, pooled values, merge script, per-resample results,
// draw counters [pooled element][lane]
#define NM_DS_WARP_SMEM (4 * NM_DS_MAX_N * 4 + 2 * NM_DS_MAX_N * 2 + NM_DS_MAX_TIMES * 4 + 2 * NM_DS_MAX_N * 32)

// Overwrites ks_dnum / ks_d / ks_p of every row whose segment has seg_cov > 0 and one of whose
// groups has more reads than that (myDetect.py:345-361).  cursor, too_deep: device ints, zero at
// launch; *too_deep becomes 1 if such a row has more than NM_DS_MAX_N reads in a group.
int nm_launch_downsample(const nm_kargs& ka, const int32_t* pos, const int32_t* seg, const int32_t* seg_cov, int times,
                         int index, uint64_t seed, int* cursor, int* too_deep, int sm_count, cudaStream_t st);

// Rows of that kind with MORE than NM_DS_MAX_N reads in a group (up to NM_DS_DEEP_MAX_READS; thresholds up to
// NM_DS_DEEP_MAX_COV): one CTA per row -- the groups are sorted in shared memory and parked in `scratch`
// (nm_downsample_deep_scratch_bytes), a warp per resample gathers its draws, sorts them and ranks them.  Same
// random stream, same result definition.  *too_deep |= 2 for a row beyond these limits.
size_t nm_downsample_deep_scratch_bytes(int sm_count);
int nm_launch_downsample_deep(const nm_kargs& ka, const int32_t* pos, const int32_t* seg, const int32_t* seg_cov, int times,
                              int index, uint64_t seed, int* cursor, int* too_deep, float* scratch, int sm_count,
                              cudaStream_t st);
