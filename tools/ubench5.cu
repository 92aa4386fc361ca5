// ubench5.cu -- issue rate of candidate compare-exchange building blocks on sm_100a, alone and
// interleaved (do two opcodes share a pipe?).  Not part of the product.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_build/ubench5 tools/ubench5.cu
// Every kernel runs 16 independent dependency chains per thread, ITERS times; the table prints
// warp-instructions per cycle per SM sub-partition at 1, 2, 4 warps per sub-partition.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define ITERS 4096
#define CH 16

#define OP2(name, asm_a, asm_b)                                                                     \
  __global__ void name(unsigned* out, unsigned seed, unsigned one) {                                 \
    unsigned x[CH], y[CH];                                                                           \
    _Pragma("unroll") for (int i = 0; i < CH; ++i) {                                                 \
      x[i] = (threadIdx.x * 2654435761u + i * 40503u + seed) & 0x3bff3bffu | 0x04000400u;           \
      y[i] = (threadIdx.x * 40503u + i * 2654435761u + seed) & 0x3bff3bffu | 0x04000400u;           \
    }                                                                                                \
    for (int it = 0; it < ITERS; ++it) {                                                             \
      _Pragma("unroll") for (int i = 0; i < CH; ++i) {                                               \
        asm volatile(asm_a : "+r"(x[i]) : "r"(y[i]), "r"(one));                                      \
        asm volatile(asm_b : "+r"(y[i]) : "r"(x[i]), "r"(one));                                      \
      }                                                                                              \
    }                                                                                                \
    unsigned s = 0;                                                                                  \
    _Pragma("unroll") for (int i = 0; i < CH; ++i) s ^= x[i] ^ y[i];                                 \
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;                                                  \
  }

OP2(k_imnmx, "min.s32 %0, %0, %1;", "max.s32 %0, %0, %1;")
OP2(k_u16x2, "min.u16x2 %0, %0, %1;", "max.u16x2 %0, %0, %1;")
OP2(k_hmnmx2, "min.f16x2 %0, %0, %1;", "max.f16x2 %0, %0, %1;")
OP2(k_bfmnmx2, "min.bf16x2 %0, %0, %1;", "max.bf16x2 %0, %0, %1;")
OP2(k_u16_h, "min.u16x2 %0, %0, %1;", "max.f16x2 %0, %0, %1;")
OP2(k_s32_h, "min.s32 %0, %0, %1;", "max.f16x2 %0, %0, %1;")
OP2(k_imad, "mad.lo.s32 %0, %0, %2, %1;", "mad.lo.s32 %0, %0, %2, %1;")
OP2(k_min_imad, "min.s32 %0, %0, %1;", "mad.lo.s32 %0, %0, %2, %1;")
OP2(k_h_imad, "min.f16x2 %0, %0, %1;", "mad.lo.s32 %0, %0, %2, %1;")
OP2(k_lop, "xor.b32 %0, %0, %1;", "and.b32 %0, %0, %1;")
OP2(k_lop_imad, "xor.b32 %0, %0, %1;", "mad.lo.s32 %0, %0, %2, %1;")
OP2(k_hadd2, "add.f16x2 %0, %0, %1;", "add.f16x2 %0, %0, %1;")
OP2(k_h_hadd2, "min.f16x2 %0, %0, %1;", "add.f16x2 %0, %0, %1;")
OP2(k_u16_hadd2, "min.u16x2 %0, %0, %1;", "add.f16x2 %0, %0, %1;")
OP2(k_prmt, "prmt.b32 %0, %0, %1, 0x5410;", "prmt.b32 %0, %0, %1, 0x7632;")
OP2(k_add_s16x2, "add.s16x2 %0, %0, %1;", "add.s16x2 %0, %0, %1;")
__global__ void k_min3(unsigned* out, unsigned seed, unsigned one) {
  int x[CH], y[CH], z[CH];
#pragma unroll
  for (int i = 0; i < CH; ++i) {
    x[i] = threadIdx.x * 2654435761u + i * 40503u + seed;
    y[i] = threadIdx.x * 40503u + i * 2654435761u + seed;
    z[i] = x[i] ^ (y[i] >> 3);
  }
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < CH; ++i) {
      x[i] = min(min(x[i], y[i]), z[i]) + (int)one;
      y[i] = max(max(x[i], y[i]), z[i]) - (int)one;
    }
  }
  unsigned s = 0;
#pragma unroll
  for (int i = 0; i < CH; ++i) s ^= x[i] ^ y[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// float ops need "f" registers
#define OPF(name, asm_a, asm_b)                                                                     \
  __global__ void name(unsigned* out, unsigned seed, unsigned one) {                                 \
    float x[CH], y[CH];                                                                              \
    const float c = __uint_as_float(0x3f800000u + (one - 1));                                        \
    _Pragma("unroll") for (int i = 0; i < CH; ++i) {                                                 \
      x[i] = (float)((threadIdx.x * 2654435761u + i * 40503u + seed) >> 12);                         \
      y[i] = (float)((threadIdx.x * 40503u + i * 2654435761u + seed) >> 12);                         \
    }                                                                                                \
    for (int it = 0; it < ITERS; ++it) {                                                             \
      _Pragma("unroll") for (int i = 0; i < CH; ++i) {                                               \
        asm volatile(asm_a : "+f"(x[i]) : "f"(y[i]), "f"(c));                                        \
        asm volatile(asm_b : "+f"(y[i]) : "f"(x[i]), "f"(c));                                        \
      }                                                                                              \
    }                                                                                                \
    float s = 0;                                                                                     \
    _Pragma("unroll") for (int i = 0; i < CH; ++i) s += x[i] + y[i];                                 \
    out[blockIdx.x * blockDim.x + threadIdx.x] = __float_as_uint(s);                                 \
  }

OPF(k_fmnmx, "min.f32 %0, %0, %1;", "max.f32 %0, %0, %1;")
OPF(k_fadd, "add.f32 %0, %0, %1;", "add.f32 %0, %0, %1;")
OPF(k_ffma, "fma.rn.f32 %0, %0, %2, %1;", "fma.rn.f32 %0, %0, %2, %1;")
OPF(k_ffma_imm, "fma.rn.f32 %0, %0, 0f3F800000, %1;", "fma.rn.f32 %0, %0, 0fBF800000, %1;")
OPF(k_fmnmx_fadd, "min.f32 %0, %0, %1;", "add.f32 %0, %0, %1;")
OPF(k_fmnmx_ffma_imm, "min.f32 %0, %0, %1;", "fma.rn.f32 %0, %0, 0f3F800000, %1;")
OPF(k_fmnmx3, "min.f32 %0, %0, %1, %2;", "max.f32 %0, %0, %1, %2;")

typedef void (*kern_t)(unsigned*, unsigned, unsigned);

static double time_ms(kern_t k, int blocks, int threads, unsigned* out) {
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  k<<<blocks, threads>>>(out, 1, 1);
  cudaDeviceSynchronize();
  cudaEventRecord(a);
  k<<<blocks, threads>>>(out, 1, 1);
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms;
  cudaEventElapsedTime(&ms, a, b);
  return ms;
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  int clk_khz;
  cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  unsigned* out;
  cudaMalloc(&out, 148 * 1024 * 4 * 4);
  printf("device %s SMs %d clock attr %d kHz\n", p.name, p.multiProcessorCount, clk_khz);
  struct { const char* name; kern_t k; } ks[] = {
      {"VIMNMX.s32 min+max", k_imnmx}, {"VIMNMX.U16x2 min+max", k_u16x2}, {"HMNMX2 f16x2 min+max", k_hmnmx2},
      {"HMNMX2 bf16x2 min+max", k_bfmnmx2}, {"VIMNMX.U16x2 min + HMNMX2 max", k_u16_h},
      {"VIMNMX.s32 min + HMNMX2 max", k_s32_h}, {"IMAD + IMAD", k_imad}, {"VIMNMX min + IMAD", k_min_imad},
      {"HMNMX2 min + IMAD", k_h_imad}, {"LOP3 + LOP3", k_lop}, {"LOP3 + IMAD", k_lop_imad}, {"HADD2 + HADD2", k_hadd2},
      {"HMNMX2 + HADD2", k_h_hadd2}, {"VIMNMX.U16x2 + HADD2", k_u16_hadd2}, {"PRMT + PRMT", k_prmt},
      {"add.s16x2 x2", k_add_s16x2}, {"VIMNMX3 s32 min3+max3", k_min3}, {"FMNMX min+max", k_fmnmx}, {"FADD + FADD", k_fadd},
      {"FFMA reg + FFMA reg", k_ffma}, {"FFMA imm + FFMA imm", k_ffma_imm}, {"FMNMX + FADD", k_fmnmx_fadd},
      {"FMNMX + FFMA imm", k_fmnmx_ffma_imm}, {"FMNMX3 min3+max3", k_fmnmx3}};
  const double inst_per_thread = (double)ITERS * CH * 2;
  for (auto& e : ks) {
    printf("%-34s", e.name);
    for (int wps = 1; wps <= 4; wps *= 2) {
      const int threads = 128 * wps, blocks = p.multiProcessorCount;
      const double ms = time_ms(e.k, blocks, threads, out);
      const double cycles = ms * 1e-3 * clk_khz * 1e3;
      printf("  w%d: %.3f ipc/smsp", wps, inst_per_thread * wps / cycles);
    }
    printf("  %s\n", cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
