// ubench.cu -- pipe-throughput microbenchmark for compare-exchange (CE) formulations on sm_100a.
// Not part of the product; used once to decide how the sorting networks should spell a CE.
//   A  float CE: FMNMX + FMNMX
//   B  int   CE: IMNMX + IMNMX (VIMNMX on sm_100)
//   C  int   CE: min + (a + b - lo) written in C (compiler's choice of IADD3 / IMAD)
//   D  int   CE: min + two mad.lo (forced onto the FMA pipe)
//   E  mix: two D-type CEs for every A-type CE
// Prints CE/clk/SM for each at 1, 2, 4, 8 warps per SM sub-partition.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>

#define R 32
#define ITERS 2000

template <int MODE> __device__ __forceinline__ void ce_f(float& a, float& b) {
  float lo = fminf(a, b), hi = fmaxf(a, b); a = lo; b = hi;
}
__device__ int g_one, g_mone;  // runtime 1 / -1 so that ptxas cannot fold the IMADs into IADD3
template <int MODE> __device__ __forceinline__ void ce_i(int& a, int& b, int one = 1, int mone = -1) {
  if (MODE == 1) { int lo = min(a, b), hi = max(a, b); a = lo; b = hi; }
  else if (MODE == 2) { int lo = min(a, b); int hi = a + b - lo; a = lo; b = hi; }
  else { int lo = min(a, b); int t, hi;
    asm volatile("mad.lo.s32 %0, %1, %3, %2;" : "=r"(t) : "r"(a), "r"(b), "r"(one));
    asm volatile("mad.lo.s32 %0, %1, %3, %2;" : "=r"(hi) : "r"(lo), "r"(t), "r"(mone));
    a = lo; b = hi; }
}

template <int MODE> __global__ void k(int* out, int seed) {
  int x[R];
  const int one = seed, mone = -seed;
#pragma unroll
  for (int i = 0; i < R; ++i) x[i] = (threadIdx.x * 2654435761u + i * 40503u + seed) >> 1;
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int d = 1; d < R; d <<= 1) {
#pragma unroll
      for (int i = 0; i < R; ++i) {
        if ((i & d) == 0) {
          const int j = i | d;
          if (MODE == 0) { float a = __int_as_float(x[i] & 0x7f7fffff), b = __int_as_float(x[j] & 0x7f7fffff); ce_f<0>(a, b); x[i] = __float_as_int(a); x[j] = __float_as_int(b); }
          else if (MODE == 4) {
            if (((i / 1) % 3) == 0) { float a = __int_as_float(x[i]), b = __int_as_float(x[j]); ce_f<0>(a, b); x[i] = __float_as_int(a); x[j] = __float_as_int(b); }
            else ce_i<3>(x[i], x[j], one, mone);
          } else ce_i<MODE>(x[i], x[j], one, mone);
        }
      }
    }
  }
  int s = 0;
#pragma unroll
  for (int i = 0; i < R; ++i) s ^= x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// pure float version without the masking ANDs (values kept as floats in registers)
__global__ void kf(float* out, int seed) {
  float x[R];
#pragma unroll
  for (int i = 0; i < R; ++i) x[i] = (float)((threadIdx.x * 2654435761u + i * 40503u + seed) >> 8);
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int d = 1; d < R; d <<= 1) {
#pragma unroll
      for (int i = 0; i < R; ++i) if ((i & d) == 0) ce_f<0>(x[i], x[i | d]);
    }
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < R; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <class F> static double time_ms(F f) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  f(); cudaDeviceSynchronize();
  cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b); return ms;
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  int* out; cudaMalloc(&out, 148 * 1024 * 4 * 8);
  int clk_khz; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  printf("device %s SMs %d clock attr %d kHz\n", p.name, p.multiProcessorCount, clk_khz);
  const double ces_per_thread = (double)ITERS * 5 * (R / 2);
  for (int wps = 1; wps <= 8; wps *= 2) {
    const int threads = 128 * wps, blocks = p.multiProcessorCount;
    double t[6];
    t[0] = time_ms([&] { kf<<<blocks, threads>>>((float*)out, 1); });
    t[1] = time_ms([&] { k<1><<<blocks, threads>>>(out, 1); });
    t[2] = time_ms([&] { k<2><<<blocks, threads>>>(out, 1); });
    t[3] = time_ms([&] { k<3><<<blocks, threads>>>(out, 1); });
    t[4] = time_ms([&] { k<4><<<blocks, threads>>>(out, 1); });
    printf("warps/SMSP %d:", wps);
    const char* nm[5] = {"A_fmnmx", "B_imnmx", "C_min+add", "D_min+2mad", "E_mix"};
    for (int m = 0; m < 5; ++m) {
      const double ce_per_s_per_sm = ces_per_thread * threads / (t[m] * 1e-3);
      printf("  %s %.3f ms (%.1f GCE/s/SM)", nm[m], t[m], ce_per_s_per_sm / 1e9);
    }
    printf("\n");
  }
  return 0;
}
