// ubench3.cu -- flat (nm_sortnet) vs looped (nm_sortloop) sort of N registers per thread:
// compare-exchanges per second per SM and effective sorts per second, for both key types
// (compile twice: default float keys, -DNM_INT_KEYS).  One warp per CTA as in the lane kernel.
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include "../nanomod_b200/csrc/nm_lane.cuh"

template <int N, bool LOOPED>
__global__ void k(nm_key* out, int iters, int seed, int one, int mone) {
  nm_key x[N];
#pragma unroll
  for (int i = 0; i < N; ++i) x[i] = (nm_key)((int)((threadIdx.x * 2654435761u + i * 40503u + seed) >> 9) - 4000000);
  for (int it = 0; it < iters; ++it) {
    if (LOOPED) nm_sortloop<N>::run(x, one, mone); else nm_sortnet<N>::run(x, one, mone);
#pragma unroll
    for (int i = 0; i < N; i += 2) x[i] = -x[i];
  }
  nm_key s = 0;
#pragma unroll
  for (int i = 0; i < N; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int N, bool LOOPED>
static void run(nm_key* out, int sms) {
  const int iters = 200;
  for (int wpsm = 4; wpsm <= 8; wpsm *= 2) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    const int blocks = sms * wpsm;
    k<N, LOOPED><<<blocks, 32>>>(out, 2, 1, 1, -1);
    cudaDeviceSynchronize();
    cudaEventRecord(a);
    k<N, LOOPED><<<blocks, 32>>>(out, iters, 1, 1, -1);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    const double nce = LOOPED ? nm_sortloop<N>::kComparators : nm_sortnet<N>::kComparators;
    const double sorts = (double)iters * 32.0 * wpsm;  // per SM
    printf("N=%3d %-6s warps/SM %d: %.3f ms  %.1f GCE/s/SM  %.1f Msorts/s/SM  err=%s\n", N, LOOPED ? "looped" : "flat",
           wpsm, ms, nce * sorts / (ms * 1e-3) / 1e9, sorts / (ms * 1e-3) / 1e6, cudaGetErrorString(cudaGetLastError()));
  }
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  nm_key* out; cudaMalloc(&out, p.multiProcessorCount * 16 * 32 * sizeof(nm_key));
#ifdef NM_INT_KEYS
  printf("int32 keys, mixed ALU/FMA compare-exchange\n");
#else
  printf("float keys, FMNMX compare-exchange\n");
#endif
  run<80, false>(out, p.multiProcessorCount);
  run<80, true>(out, p.multiProcessorCount);
  run<104, false>(out, p.multiProcessorCount);
  run<104, true>(out, p.multiProcessorCount);
  run<128, false>(out, p.multiProcessorCount);
  run<128, true>(out, p.multiProcessorCount);
  return 0;
}
