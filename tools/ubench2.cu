// ubench2.cu -- does straight-line sorting-network code run at ALU speed, or is it limited by
// instruction fetch?  Applies nm_sortnet<N> K times to N registers per thread (re-scrambling
// between applications so the data is not already sorted) for several N (= code footprints)
// and warps per SM, and prints compare-exchanges per second per SM.  ALU-bound reference:
// ~62.5 GCE/s/SM at 1965 MHz (tools/ubench.cu, two FMNMX per CE).
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include "../nanomod_b200/csrc/nm_lane.cuh"

template <int N>
__global__ void k(float* out, int iters, int seed) {
  float x[N];
#pragma unroll
  for (int i = 0; i < N; ++i) x[i] = (float)((threadIdx.x * 2654435761u + i * 40503u + seed) >> 8);
  for (int it = 0; it < iters; ++it) {
    nm_sortnet<N>::run(x);
    // cheap re-scramble: reverse halves via negation of alternating elements
#pragma unroll
    for (int i = 0; i < N; i += 2) x[i] = -x[i];
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < N; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int N>
static void run(float* out, int sms) {
  const int iters = 200;
  for (int wpsm = 4; wpsm <= 16; wpsm *= 2) {
    if (N > 64 && wpsm > 8) continue;  // register limit
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    const int blocks = sms * wpsm;  // one warp per CTA, as in the lane kernel
    k<N><<<blocks, 32>>>(out, 2, 1);
    cudaDeviceSynchronize();
    cudaEventRecord(a);
    k<N><<<blocks, 32>>>(out, iters, 1);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    const double ces = (double)nm_sortnet<N>::kComparators * iters * 32.0 * wpsm;  // per SM
    printf("N=%3d code~%5.1f KB  warps/SM %2d: %.3f ms  %.1f GCE/s/SM  err=%s\n", N,
           nm_sortnet<N>::kComparators * 32.0 / 1024.0, wpsm, ms, ces / (ms * 1e-3) / 1e9,
           cudaGetErrorString(cudaGetLastError()));
  }
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  float* out; cudaMalloc(&out, p.multiProcessorCount * 16 * 32 * sizeof(float));
  run<16>(out, p.multiProcessorCount);
  run<32>(out, p.multiProcessorCount);
  run<48>(out, p.multiProcessorCount);
  run<64>(out, p.multiProcessorCount);
  run<80>(out, p.multiProcessorCount);
  run<104>(out, p.multiProcessorCount);
  run<128>(out, p.multiProcessorCount);
  return 0;
}
