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

// Per-lane description of one row of a tile (tile = 32 consecutive rows, one per lane).
struct nm_tile_meta {
  int64_t r;        // row index
  long long s0, s1; // start of the row's slices in vals0 / vals1
  int n0, n1;       // coverage (0 when the lane has no lane-tier row)
  bool ok;
};

__device__ __forceinline__ nm_tile_meta nm_tile_fetch(const nm_kargs& a, int64_t tile, int lane) {
  nm_tile_meta m;
  m.r = tile * 32 + lane;
  m.n0 = m.n1 = 0;
  m.s0 = m.s1 = 0;
  m.ok = false;
  if (tile >= 0 && m.r < a.n_rows) {
    const int nn0 = a.row_n0[m.r], nn1 = a.row_n1[m.r];
    if (nn0 <= NM_LANE_TIER_MAX && nn1 <= NM_LANE_TIER_MAX) {
      const int32_t src = a.row_pos_index[m.r];
      m.ok = true;
      m.n0 = nn0;
      m.n1 = nn1;
      m.s0 = a.off0[src];
      m.s1 = a.off1[src];
    }
  }
  return m;
}

// Warp-level plan for staging a tile: where its two value slices start, whether each is one
// contiguous run of the CSR array (-> one TMA bulk copy) and each lane's offset in the region.
struct nm_tile_stage {
  long long al0, al1;      // 16-byte aligned starts of the bulk copies
  unsigned bytes0, bytes1; // bulk copy sizes (0: not contiguous, use the gather)
  int base0, base1;        // this lane's row offset inside region A / B (contiguous case)
  int nmax, tmax;
  bool any;
};

__device__ __forceinline__ nm_tile_stage nm_tile_plan(const nm_tile_meta& m) {
  nm_tile_stage st;
  const long long big = 0x7fffffffffffffffLL;
  st.any = __any_sync(0xffffffffu, m.ok);
  const long long first0 = nm_warp_min_ll(m.ok ? m.s0 : big), first1 = nm_warp_min_ll(m.ok ? m.s1 : big);
  const long long end0 = nm_warp_max_ll(m.ok ? m.s0 + m.n0 : -1), end1 = nm_warp_max_ll(m.ok ? m.s1 + m.n1 : -1);
  const int tot0 = __reduce_add_sync(0xffffffffu, m.n0), tot1 = __reduce_add_sync(0xffffffffu, m.n1);
  st.nmax = __reduce_max_sync(0xffffffffu, m.n0 > m.n1 ? m.n0 : m.n1);
  st.tmax = __reduce_max_sync(0xffffffffu, m.n0 + m.n1);
  st.al0 = first0 & ~3LL;
  st.al1 = first1 & ~3LL;
  const bool contig0 = st.any && (end0 - first0) == (long long)tot0;
  const bool contig1 = st.any && (end1 - first1) == (long long)tot1;
  st.bytes0 = contig0 ? (unsigned)(((first0 - st.al0) + tot0 + 3) & ~3LL) * 4u : 0u;
  st.bytes1 = contig1 ? (unsigned)(((first1 - st.al1) + tot1 + 3) & ~3LL) * 4u : 0u;
  st.base0 = m.ok ? (int)(m.s0 - st.al0) : 0;
  st.base1 = m.ok ? (int)(m.s1 - st.al1) : 0;
  return st;
}

// Persistent lane-tier kernel: every warp loops over tiles taken from a global cursor.  While
// a tile is being sorted, the next tile's metadata is already in registers and its value
// slices are being pulled into L2; its TMA copies are issued the moment the merge walk has
// released the two regions, so that they overlap the fp64 tails of the current tile.
__global__ void __launch_bounds__(32 * NM_LANE_WARPS, 8 / NM_LANE_WARPS)
nm_lane_kernel(const nm_kargs a, const int want_u, const int want_t) {
  extern __shared__ __align__(128) unsigned char nm_smem[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  // per-warp slice of shared memory: [mbarrier 16 B][region A][region B]
  unsigned char* my = nm_smem + (size_t)wib * (16 + 2 * (size_t)a.region_floats * sizeof(float));
  uint64_t* bar = reinterpret_cast<uint64_t*>(my);
  float* regA = reinterpret_cast<float*>(my + 16);
  float* regB = regA + a.region_floats;
  const int64_t n_tiles = (a.n_rows + 31) >> 5;
  const int64_t n_warps = (int64_t)gridDim.x * NM_LANE_WARPS;

  int64_t tile = (int64_t)blockIdx.x * NM_LANE_WARPS + wib;
  if (tile >= n_tiles) return;
  if (lane == 0) nm_mbar_init(bar, 1);
  __syncwarp();
  unsigned parity = 0;

  nm_tile_meta cur = nm_tile_fetch(a, tile, lane);
  nm_tile_stage cst = nm_tile_plan(cur);
  bool staged = false;  // TMA for `cur` already issued?

  while (true) {
    // ---- claim the next tile and start fetching its metadata (consumed after the walk)
    int64_t next = -1;
    {
      long long t = 0;
      if (lane == 0) t = (long long)atomicAdd(a.tile_cursor, 1) + n_warps;
      t = __shfl_sync(0xffffffffu, t, 0);
      next = t < n_tiles ? t : -1;
    }
    const nm_tile_meta nxt = nm_tile_fetch(a, next, lane);

    if (cst.any) {
      // ---- stage the current tile (unless its copies were issued at the end of the last one)
      if (!staged && (cst.bytes0 | cst.bytes1) && lane == 0) {
        nm_mbar_expect_tx(bar, cst.bytes0 + cst.bytes1);
        if (cst.bytes0) nm_bulk_g2s(regA, a.vals0 + cst.al0, cst.bytes0, bar);
        if (cst.bytes1) nm_bulk_g2s(regB, a.vals1 + cst.al1, cst.bytes1, bar);
      }
      int base0 = cst.base0, base1 = cst.base1;
      if (!cst.bytes0) base0 = nm_lane_gather(regA, a.vals0, cur.s0, cur.n0, lane);
      if (!cst.bytes1) base1 = nm_lane_gather(regB, a.vals1, cur.s1, cur.n1, lane);
      // plan the next tile now (its metadata loads have had the whole staging latency to land)
      // and pull its value slices into L2 while this tile is being sorted
      const nm_tile_stage nst = nm_tile_plan(nxt);
      if (lane == 0) {
        if (nst.bytes0) nm_prefetch_l2(a.vals0 + nst.al0, nst.bytes0);
        if (nst.bytes1) nm_prefetch_l2(a.vals1 + nst.al1, nst.bytes1);
      }
      if (cst.bytes0 | cst.bytes1) {
        nm_mbar_wait(bar, parity);
        parity ^= 1u;
      }
      __syncwarp();

      const int n0 = cur.n0, n1 = cur.n1;
      const int nsel = (cst.nmax + NM_LANE_STEP - 1) / NM_LANE_STEP * NM_LANE_STEP;
      nm_lane_acc acc;
      acc.dnum = acc.r2 = acc.tie = 0;
      acc.mean0 = acc.var0 = acc.mean1 = acc.var1 = 0.0;
#define NM_CALL(NN) \
  nm_lane_tile<NN>(regA, regB, base0, base1, n0, n1, lane, want_t != 0, a.one, a.mone, &acc)
      NM_DISPATCH_N(nsel, NM_CALL)
#undef NM_CALL

      const nm_key* colA = reinterpret_cast<const nm_key*>(regA) + lane;
      const nm_key* colB = reinterpret_cast<const nm_key*>(regB) + lane;
      const int iters = (cst.tmax + 1) >> 1;
      const bool uniform = __all_sync(0xffffffffu, cur.ok && (n0 + n1 == cst.tmax));
      if (want_u)
        nm_merge_walk<true, 32>(colA, colB, n0, n1, iters, &acc);
      else if (uniform)
        acc.dnum = nm_walk_ks_uniform(colA, colB, n0, n1);
      else
        nm_merge_walk<false, 32>(colA, colB, n0, n1, iters, &acc);

      // ---- the regions are free again: issue the next tile's copies before the fp64 tails
      __syncwarp();
      staged = false;
      if (nst.any && (nst.bytes0 | nst.bytes1)) {
        if (lane == 0) {
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          nm_mbar_expect_tx(bar, nst.bytes0 + nst.bytes1);
          if (nst.bytes0) nm_bulk_g2s(regA, a.vals0 + nst.al0, nst.bytes0, bar);
          if (nst.bytes1) nm_bulk_g2s(regB, a.vals1 + nst.al1, nst.bytes1, bar);
        }
        staged = true;
      }

      if (cur.ok) {
        nm_row_out o;
        nm_lane_finish(acc, n0, n1, want_u != 0, want_t != 0, &o);
        nm_store_row(a, cur.r, o, want_u != 0, want_t != 0);
      }
      cur = nxt;
      cst = nst;
    } else {
      cur = nxt;
      cst = nm_tile_plan(nxt);
      staged = false;
    }
    if (next < 0) break;
  }
}

int nm_lane_smem_bytes(int region_floats) {
  return NM_LANE_WARPS * (16 + 2 * region_floats * (int)sizeof(float));
}

int nm_launch_lane(const nm_kargs& ka, bool want_u, bool want_t, int sm_count, cudaStream_t st) {
  const int smem_bytes = nm_lane_smem_bytes(ka.region_floats);
  cudaError_t e = cudaFuncSetAttribute(nm_lane_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       smem_bytes);
  if (e != cudaSuccess) return (int)e;
  int per_sm = 0;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, nm_lane_kernel, 32 * NM_LANE_WARPS, smem_bytes);
  if (e != cudaSuccess) return (int)e;
  if (per_sm < 1) per_sm = 1;
  const int64_t tiles = (ka.n_rows + 31) / 32;
  int64_t grid = (tiles + NM_LANE_WARPS - 1) / NM_LANE_WARPS;
  const int64_t resident = (int64_t)per_sm * sm_count;
  if (grid > resident) grid = resident;
  nm_lane_kernel<<<(unsigned)grid, 32 * NM_LANE_WARPS, smem_bytes, st>>>(ka, want_u ? 1 : 0, want_t ? 1 : 0);
  return (int)cudaGetLastError();
}
