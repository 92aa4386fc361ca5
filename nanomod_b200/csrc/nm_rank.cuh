// nanomod_b200 -- internal interface of the ranking unit (nm_rank.cu), used by nm_api.cu.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

// bytes of device scratch nm_rank_run needs for n rows
size_t nm_rank_scratch_bytes(int64_t n);
// Writes to order_out (device) the row indices in ranked order; comb / u may be NULL (key
// absent).  All pointers are device pointers.  Returns a cudaError_t as int; adds the number of
// kernels launched to *launches.
int nm_rank_run(const double* comb, const double* ks, const double* u, int64_t n, int reverse, int32_t* order_out,
                void* scratch, int* launches, cudaStream_t st);
