// ubench4.cu -- pair sort (two lanes per position: half sort, shuffle cross step, up-down merge)
// vs the flat one-lane sort of the same 2H elements.  Reports position-sorts per second per SM.
#include <cuda_runtime.h>
#include <stdio.h>
#include "../nanomod_b200/csrc/nm_lane.cuh"

template <int H>
__global__ void kpair(nm_key* out, int iters, int seed, int one, int mone) {
  nm_key x[H];
  const unsigned hmask = (threadIdx.x & 1) ? 0xffffffffu : 0u;
#pragma unroll
  for (int i = 0; i < H; ++i) x[i] = (nm_key)((int)((threadIdx.x * 2654435761u + i * 40503u + seed) >> 9) - 4000000);
  for (int it = 0; it < iters; ++it) {
    nm_halfsort<H>::run(x, one, mone);
#pragma unroll
    for (int k = 0; k < H; ++k) {  // cross step in each lane's own key space (odd lanes: negated)
      const int o = ~__shfl_xor_sync(0xffffffffu, (int)x[k], 1);
      x[k] = nm_min(x[k], (nm_key)o);
    }
    nm_updown<H>::run(x, one, mone);
#pragma unroll
    for (int i = 0; i < H; i += 2) x[i] = (nm_key)((int)x[i] ^ (int)(hmask | 0x55555u));
  }
  nm_key s = 0;
#pragma unroll
  for (int i = 0; i < H; ++i) s += x[nm_updown<H>::order(i)];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int N>
__global__ void kflat(nm_key* out, int iters, int seed, int one, int mone) {
  nm_key x[N];
#pragma unroll
  for (int i = 0; i < N; ++i) x[i] = (nm_key)((int)((threadIdx.x * 2654435761u + i * 40503u + seed) >> 9) - 4000000);
  for (int it = 0; it < iters; ++it) {
    nm_sortnet<N>::run(x, one, mone);
#pragma unroll
    for (int i = 0; i < N; i += 2) x[i] = (nm_key)((int)x[i] ^ 0x55555);
  }
  nm_key s = 0;
#pragma unroll
  for (int i = 0; i < N; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <class F>
static float timeit(F f) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  f(2); cudaDeviceSynchronize();
  cudaEventRecord(a); f(200); cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b); return ms;
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  const int sms = p.multiProcessorCount;
  nm_key* out; cudaMalloc(&out, sms * 32 * 32 * sizeof(nm_key));
#ifdef NM_INT_KEYS
  printf("int32 keys (mixed CE)\n");
#else
  printf("float keys\n");
#endif
  for (int w = 8; w <= 16; w += 4) {
    float ms = timeit([&](int it) { kpair<52><<<sms * w, 32>>>(out, it, 1, 1, -1); });
    printf("pair H=52  warps/SM %2d: %.3f ms  %.1f Mpos-sorts/s/SM  (%s)\n", w, ms, 200.0 * 16 * w / (ms * 1e-3) / 1e6, cudaGetErrorString(cudaGetLastError()));
    ms = timeit([&](int it) { kpair<64><<<sms * w, 32>>>(out, it, 1, 1, -1); });
    printf("pair H=64  warps/SM %2d: %.3f ms  %.1f Mpos-sorts/s/SM\n", w, ms, 200.0 * 16 * w / (ms * 1e-3) / 1e6);
  }
  for (int w = 4; w <= 8; w += 4) {
    float ms = timeit([&](int it) { kflat<104><<<sms * w, 32>>>(out, it, 1, 1, -1); });
    printf("flat N=104 warps/SM %2d: %.3f ms  %.1f Mpos-sorts/s/SM\n", w, ms, 200.0 * 32 * w / (ms * 1e-3) / 1e6);
    ms = timeit([&](int it) { kflat<128><<<sms * w, 32>>>(out, it, 1, 1, -1); });
    printf("flat N=128 warps/SM %2d: %.3f ms  %.1f Mpos-sorts/s/SM\n", w, ms, 200.0 * 32 * w / (ms * 1e-3) / 1e6);
  }
  return 0;
}
