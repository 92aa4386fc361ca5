// nanomod_b200 -- internal interface of the huge-row path (nm_huge.cu), used by nm_api.cu.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

struct nm_kargs;
// device scratch for n_huge rows holding nv0 / nv1 values in group 0 / 1
size_t nm_huge_scratch_bytes(int n_huge, long long nv0, long long nv1);
// tests of the rows of ka.deep_rows whose pow2(n0) + pow2(n1) exceeds NM_DEEP_TIER_MAX_POOLED
int nm_huge_run(const nm_kargs& ka, bool want_u, bool want_t, bool want_m, int n_huge, long long nv0, long long nv1,
                void* scratch, int* launches, cudaStream_t st);
