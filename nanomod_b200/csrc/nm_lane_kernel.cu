// nm_lane_kernel.cu -- the lane tier: one lane per position, one warp (= one CTA) per tile of
// 32 consecutive rows.  getKStest (bin/scripts/myDetect.py:327-343) for rows with <= 128 reads
// per group.  Kept in its own translation unit because the 16 unrolled sorting networks take
// about a minute to compile.
//
// Shared memory of a warp: [mbarrier 16 B][region A][region B], each region 32*(Ncls+2) floats.
//   1. the tile's two value slices (contiguous in the CSR arrays) arrive with one TMA bulk
//      copy each (cp.async.bulk + mbarrier); non-contiguous tiles use a cooperative gather;
//   2. every lane pulls its row into registers (LDS.128 when aligned), sorts it with the
//      size-Nsel network, and writes it back TRANSPOSED (element i of lane l at [i*32+l]), so
//      the data-dependent indexing of the merge walk is bank-conflict free;
//   3. one merge walk per lane gives the KS numerator (+ rank sums), then the fp64 tails.
#include "nm_device.cuh"

#ifndef NM_LANE_WARPS
#define NM_LANE_WARPS 1  // warps (= independent 32-row tiles) per CTA
#endif

// Load one group's row into registers (pad +inf), sort, write back transposed with a -inf row
// in front and a +inf sentinel row behind (layout expected by nm_merge_walk).
template <int N>
__device__ __forceinline__ void nm_lane_sort_group(float* region, int base, int n, int lane,
                                                   bool want_t, int one, int mone, double* mean,
                                                   double* var) {
  nm_key x[N];
  const float* raw = region + base;
  if (want_t) nm_moments(raw, n, mean, var);
  const bool vec = __all_sync(0xffffffffu, ((base & 3) == 0) && (n == N));
  if (vec) {
    const float4* raw4 = reinterpret_cast<const float4*>(raw);
#pragma unroll
    for (int q = 0; q < N / 4; ++q) {
      const float4 v = raw4[q];
      x[4 * q + 0] = nm_make_key(v.x);
      x[4 * q + 1] = nm_make_key(v.y);
      x[4 * q + 2] = nm_make_key(v.z);
      x[4 * q + 3] = nm_make_key(v.w);
    }
  } else {
#pragma unroll
    for (int k = 0; k < N; ++k) {
      const float v = raw[k];
      x[k] = (k < n) ? nm_make_key(v) : (nm_key)NM_KEY_PINF;
    }
  }
  nm_sortnet<N>::run(x, one, mone);
  __syncwarp();  // every lane has consumed its raw row; the region may now be overwritten
  nm_key* col = reinterpret_cast<nm_key*>(region) + lane;
  col[0] = NM_KEY_NINF;
#pragma unroll
  for (int k = 0; k < N; ++k) col[(k + 1) << 5] = x[k];
  col[(N + 1) << 5] = NM_KEY_PINF;
}

template <int N>
__device__ __forceinline__ void nm_lane_tile(float* regA, float* regB, int base0, int base1,
                                             int n0, int n1, int lane, bool want_t, int one,
                                             int mone, nm_lane_acc* acc) {
#pragma unroll 1
  for (int g = 0; g < 2; ++g) {
    double m = 0.0, v = 0.0;
    nm_lane_sort_group<N>(g ? regB : regA, g ? base1 : base0, g ? n1 : n0, lane, want_t, one, mone, &m, &v);
    if (g) { acc->mean1 = m; acc->var1 = v; } else { acc->mean0 = m; acc->var0 = v; }
  }
  __syncwarp();
}

// ------------------------------------------------------------------------------------------
// Device fast path of nm_merge_walk (KS numerator only) for tiles whose 32 lanes all hold the
// same pooled count T: no activity predicates, pointer-chasing with predicated adds and loads,
// and the ECDF numerator carried on the FMA pipe.  Same algorithm, same tie rule, same
// evaluation points as nm_merge_walk (nm_lane.cuh), which the CPU tests cover; this spelling
// exists because the ALU pipe is the kernel's bottleneck and the generic code costs ~18
// instructions per pooled element against 11 here.
//   forward chain : fa/fb = shared addresses of the heads (rows i+1 / j+1 of the two columns)
//   backward chain: ba/bb = shared addresses of the tails (rows i2 / j2)
//   128*d = addr_a * T - c, with c advanced by 128*n0 per step  (row stride = 128 bytes)
// ------------------------------------------------------------------------------------------
#ifdef NM_INT_KEYS
#define NM_KT "s32"
#define NM_KREG "r"
#else
#define NM_KT "f32"
#define NM_KREG "f"
#endif

// one forward step: take the smaller head (ties: group 0), reload it, evaluate at a tie-group end
#define NM_FWD_STEP(fa, fb, va, vb, v, c, dmax)                                      \
  asm volatile(                                                                      \
      "{\n\t.reg .pred p, q;\n\t.reg .s32 d;\n\t.reg ." NM_KT " vn;\n\t"              \
      "setp.le." NM_KT " p, %2, %3;\n\t"                                             \
      "@p add.u32 %0, %0, 128;\n\t"                                                  \
      "@!p add.u32 %1, %1, 128;\n\t"                                                 \
      "@p ld.shared." NM_KT " %2, [%0];\n\t"                                         \
      "@!p ld.shared." NM_KT " %3, [%1];\n\t"                                        \
      "min." NM_KT " vn, %2, %3;\n\t"                                                \
      "setp.gt." NM_KT " q, vn, %4;\n\t"                                             \
      "mov." NM_KT " %4, vn;\n\t"                                                    \
      "mad.lo.s32 d, %0, %6, %7;\n\t"                                                \
      "abs.s32 d, d;\n\t"                                                            \
      "@q max.s32 %5, %5, d;\n\t}"                                                   \
      : "+r"(fa), "+r"(fb), "+" NM_KREG(va), "+" NM_KREG(vb), "+" NM_KREG(v), "+r"(dmax) \
      : "r"(T), "r"(c))

// one backward step: take the larger tail (ties: group 1), reload it, evaluate at a group start
#define NM_BWD_STEP(ba, bb, ea, eb, w, c, dmax)                                      \
  asm volatile(                                                                      \
      "{\n\t.reg .pred p, q;\n\t.reg .s32 d;\n\t.reg ." NM_KT " wn;\n\t"              \
      "setp.ge." NM_KT " p, %3, %2;\n\t"                                             \
      "@p sub.u32 %1, %1, 128;\n\t"                                                  \
      "@!p sub.u32 %0, %0, 128;\n\t"                                                 \
      "@p ld.shared." NM_KT " %3, [%1];\n\t"                                         \
      "@!p ld.shared." NM_KT " %2, [%0];\n\t"                                        \
      "max." NM_KT " wn, %2, %3;\n\t"                                                \
      "setp.lt." NM_KT " q, wn, %4;\n\t"                                             \
      "mov." NM_KT " %4, wn;\n\t"                                                    \
      "mad.lo.s32 d, %0, %6, %7;\n\t"                                                \
      "abs.s32 d, d;\n\t"                                                            \
      "@q max.s32 %5, %5, d;\n\t}"                                                   \
      : "+r"(ba), "+r"(bb), "+" NM_KREG(ea), "+" NM_KREG(eb), "+" NM_KREG(w), "+r"(dmax) \
      : "r"(T), "r"(c))

__device__ __forceinline__ int nm_walk_ks_uniform(const nm_key* colA, const nm_key* colB, int n0,
                                                  int n1) {
  const int T = n0 + n1, T1 = T >> 1;
  const unsigned A0 = nm_smem_u32(colA), B0 = nm_smem_u32(colB);
  unsigned fa = A0 + 128u, fb = B0 + 128u;
  unsigned ba = A0 + 128u * (unsigned)n0, bb = B0 + 128u * (unsigned)n1;
  nm_key va = colA[32], vb = colB[32], v = nm_min(va, vb);
  nm_key ea = colA[n0 << 5], eb = colB[n1 << 5], w = nm_max(ea, eb);
  const int k0 = 128 * n0;
  // forward: 128*d = (fa - A0 - 128)*T - 128*(s+1)*n0  ->  c = -((A0+128)*T) - k0*(s+1)
  int cf = -(int)((A0 + 128u) * (unsigned)T) - k0;
  // backward: 128*d = (ba - A0)*T - 128*(T-s-1)*n0     ->  c = -(A0*T) - k0*(T-s-1)
  int cb = -(int)(A0 * (unsigned)T) - k0 * (T - 1);
  int dmax = 0;
#pragma unroll 4
  for (int s = 0; s < T1; ++s) {
    NM_FWD_STEP(fa, fb, va, vb, v, cf, dmax);
    NM_BWD_STEP(ba, bb, ea, eb, w, cb, dmax);
    cf -= k0;
    cb += k0;
  }
  if (T & 1) NM_BWD_STEP(ba, bb, ea, eb, w, cb, dmax);  // the backward chain takes ceil(T/2)
  return dmax >> 7;
}

// cooperative copy of the tile's rows when they are not one contiguous slice of vals
__device__ __forceinline__ int nm_lane_gather(float* region, const float* __restrict__ vals,
                                              long long start, int n, int lane) {
  int incl = n;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  const int base = incl - n;
  for (int rlane = 0; rlane < 32; ++rlane) {
    const int rn = __shfl_sync(0xffffffffu, n, rlane);
    const int rbase = __shfl_sync(0xffffffffu, base, rlane);
    const long long rstart = __shfl_sync(0xffffffffu, start, rlane);
    for (int k = lane; k < rn; k += 32) region[rbase + k] = vals[rstart + k];
  }
  return base;
}

__global__ void __launch_bounds__(32 * NM_LANE_WARPS, 8 / NM_LANE_WARPS)
nm_lane_kernel(const nm_kargs a, const int want_u, const int want_t) {
  extern __shared__ __align__(128) unsigned char nm_smem[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  // per-warp slice of shared memory: [mbarrier 16 B][region A][region B]
  unsigned char* my = nm_smem + (size_t)wib * (16 + 2 * (size_t)a.region_floats * sizeof(float));
  uint64_t* bar = reinterpret_cast<uint64_t*>(my);
  float* regA = reinterpret_cast<float*>(my + 16);
  float* regB = regA + a.region_floats;

  const int64_t r = ((int64_t)blockIdx.x * NM_LANE_WARPS + wib) * 32 + lane;
  int n0 = 0, n1 = 0;
  long long s0 = 0, s1 = 0;
  bool ok = false;
  if (r < a.n_rows) {
    const int nn0 = a.row_n0[r], nn1 = a.row_n1[r];
    if (nn0 <= NM_LANE_TIER_MAX && nn1 <= NM_LANE_TIER_MAX) {
      const int32_t src = a.row_pos_index[r];
      ok = true;
      n0 = nn0;
      n1 = nn1;
      s0 = a.off0[src];
      s1 = a.off1[src];
    }
  }
  if (!__any_sync(0xffffffffu, ok)) return;

  const long long big = 0x7fffffffffffffffLL;
  const long long first0 = nm_warp_min_ll(ok ? s0 : big), first1 = nm_warp_min_ll(ok ? s1 : big);
  const long long end0 = nm_warp_max_ll(ok ? s0 + n0 : -1), end1 = nm_warp_max_ll(ok ? s1 + n1 : -1);
  const int tot0 = __reduce_add_sync(0xffffffffu, n0), tot1 = __reduce_add_sync(0xffffffffu, n1);
  const bool contig0 = (end0 - first0) == (long long)tot0;
  const bool contig1 = (end1 - first1) == (long long)tot1;
  const int nmax = __reduce_max_sync(0xffffffffu, n0 > n1 ? n0 : n1);
  const int tmax = __reduce_max_sync(0xffffffffu, n0 + n1);
  const int nsel = (nmax + NM_LANE_STEP - 1) / NM_LANE_STEP * NM_LANE_STEP;

  // stage the tile: one TMA bulk copy per contiguous group slice, else a cooperative gather
  const long long al0 = first0 & ~3LL, al1 = first1 & ~3LL;
  int base0 = ok ? (int)(s0 - al0) : 0, base1 = ok ? (int)(s1 - al1) : 0;
  if (contig0 || contig1) {
    if (lane == 0) {
      nm_mbar_init(bar, 1);
      const uint32_t b0 = contig0 ? (uint32_t)(((first0 - al0) + tot0 + 3) & ~3LL) * 4u : 0u;
      const uint32_t b1 = contig1 ? (uint32_t)(((first1 - al1) + tot1 + 3) & ~3LL) * 4u : 0u;
      nm_mbar_expect_tx(bar, b0 + b1);
      if (contig0) nm_bulk_g2s(regA, a.vals0 + al0, b0, bar);
      if (contig1) nm_bulk_g2s(regB, a.vals1 + al1, b1, bar);
    }
    __syncwarp();
  }
  if (!contig0) base0 = nm_lane_gather(regA, a.vals0, s0, n0, lane);
  if (!contig1) base1 = nm_lane_gather(regB, a.vals1, s1, n1, lane);
  if (contig0 || contig1) nm_mbar_wait(bar, 0);
  __syncwarp();

  nm_lane_acc acc;
  acc.dnum = acc.r2 = acc.tie = 0;
  acc.mean0 = acc.var0 = acc.mean1 = acc.var1 = 0.0;
#define NM_CALL(NN) \
  nm_lane_tile<NN>(regA, regB, base0, base1, n0, n1, lane, want_t != 0, a.one, a.mone, &acc)
  NM_DISPATCH_N(nsel, NM_CALL)
#undef NM_CALL

  const nm_key* colA = reinterpret_cast<const nm_key*>(regA) + lane;
  const nm_key* colB = reinterpret_cast<const nm_key*>(regB) + lane;
  const int iters = (tmax + 1) >> 1;
  const bool uniform = __all_sync(0xffffffffu, ok && (n0 + n1 == tmax));
  if (want_u)
    nm_merge_walk<true, 32>(colA, colB, n0, n1, iters, &acc);
  else if (uniform)
    acc.dnum = nm_walk_ks_uniform(colA, colB, n0, n1);
  else
    nm_merge_walk<false, 32>(colA, colB, n0, n1, iters, &acc);

  if (ok) {
    nm_row_out o;
    nm_lane_finish(acc, n0, n1, want_u != 0, want_t != 0, &o);
    nm_store_row(a, r, o, want_u != 0, want_t != 0);
  }
}

int nm_lane_smem_bytes(int region_floats) {
  return NM_LANE_WARPS * (16 + 2 * region_floats * (int)sizeof(float));
}

int nm_launch_lane(const nm_kargs& ka, bool want_u, bool want_t, cudaStream_t st) {
  const int smem_bytes = nm_lane_smem_bytes(ka.region_floats);
  cudaError_t e = cudaFuncSetAttribute(nm_lane_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       smem_bytes);
  if (e != cudaSuccess) return (int)e;
  const int64_t tiles = (ka.n_rows + 31) / 32;
  const unsigned grid = (unsigned)((tiles + NM_LANE_WARPS - 1) / NM_LANE_WARPS);
  nm_lane_kernel<<<grid, 32 * NM_LANE_WARPS, smem_bytes, st>>>(ka, want_u ? 1 : 0, want_t ? 1 : 0);
  return (int)cudaGetLastError();
}
