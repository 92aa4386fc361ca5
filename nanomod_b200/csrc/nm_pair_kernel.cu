// nm_pair_kernel.cu -- the pair tier: TWO lanes per position, one warp per tile of 16
// consecutive rows.  Same statistics as the lane tier (getKStest, bin/scripts/myDetect.py:327-343:
// the KS numerator, optionally the Welch moments) for pileups whose longest lane-tier row exceeds
// 64 reads.  Why a second tier: a one-lane sort of ~100 values is 35-60 KB of straight-line code,
// more than the 32 KB instruction cache, and needs 27 KB of shared memory per warp, so the lane
// kernel is instruction-fetch bound at 8 warps/SM (profiles/round1_variants.md).  Splitting a
// position over two lanes halves both: each lane sorts half of the row's window with a network
// of at most 68 inputs, the halves are merged with one shuffle step plus a small register
// network, and 16 warps/SM stay resident.
//
//   window     lane h of a pair owns window slots [h*H, (h+1)*H) of the row's 16-byte aligned
//              window (2H >= alignment shift + n); invalid slots are +inf.
//   key space  lane 0 works on the int32 keys u = key, lane 1 on u = ~key (order reversed), so
//              both lanes run identical code: each sorts ascending in its own space, i.e. lane 1
//              holds its half DEscending in true order.
//   cross step u[k] = min(u[k], ~shfl_xor(u[k], 1)): lane 0 keeps min(x0[k], k-th largest of
//              lane 1) -> the H smallest of the row, lane 1 the H largest; both results are
//              up-down sequences in the lane's own space.
//   merge      nm_updown<H> sorts the up-down sequence; lane 0 then holds ranks 0..H-1 ascending,
//              lane 1 ranks 2H-1..H (its slot k has true rank 2H-1-k).
//   columns    written transposed, 16 columns per region: key of rank r of row p at
//              region[(r+1)*16 + p], -inf in row 0, +inf from row n+1 to row 2H+1.
//   walk       lane 0 runs the forward chain of nm_merge_walk, lane 1 the backward chain, written
//              as ONE routine: the backward chain is the forward chain in the negated key space
//              with the two groups swapped (ties: group 1 first == the mirror rule).
#include "nm_device.cuh"

#ifdef NM_INT_KEYS

#define NM_PAIR_ROWS 16
#define NM_PAIR_MAX_WARPS 4
#define NM_PAIR_MAX_H 68

// one step of a lane's chain in its own frame (u-space, groups a'/b'):
//   take the smaller head (ties: a'), advance its pointer by `step` bytes, reload + re-map it,
//   evaluate |64*d| at a tie-group boundary.  64*d = fa*Tm + c (see nm_pair_walk).
#define NM_PAIR_STEP(fa, fb, ua, ub, v, c, dmax, step, Tm, m)                              \
  asm volatile(                                                                            \
      "{\n\t.reg .pred p, q;\n\t.reg .s32 d, vn;\n\t"                                      \
      "setp.le.s32 p, %2, %3;\n\t"                                                         \
      "@p add.s32 %0, %0, %6;\n\t"                                                         \
      "@!p add.s32 %1, %1, %6;\n\t"                                                        \
      "@p ld.shared.s32 %2, [%0];\n\t"                                                     \
      "@!p ld.shared.s32 %3, [%1];\n\t"                                                    \
      "@p xor.b32 %2, %2, %9;\n\t"                                                         \
      "@!p xor.b32 %3, %3, %9;\n\t"                                                        \
      "min.s32 vn, %2, %3;\n\t"                                                            \
      "setp.gt.s32 q, vn, %4;\n\t"                                                         \
      "mov.s32 %4, vn;\n\t"                                                                \
      "mad.lo.s32 d, %0, %7, %8;\n\t"                                                      \
      "abs.s32 d, d;\n\t"                                                                  \
      "@q max.s32 %5, %5, d;\n\t}"                                                         \
      : "+r"(fa), "+r"(fb), "+r"(ua), "+r"(ub), "+r"(v), "+r"(dmax)                        \
      : "r"(step), "r"(Tm), "r"(c), "r"(m))

// same with an activity predicate (lanes whose pooled count is smaller than the trip count)
#define NM_PAIR_STEP_ACT(fa, fb, ua, ub, v, c, dmax, step, Tm, m, s, T)                    \
  asm volatile(                                                                            \
      "{\n\t.reg .pred p, q, act, pa, pb;\n\t.reg .s32 d, vn;\n\t"                         \
      "setp.lt.s32 act, %10, %11;\n\t"                                                     \
      "setp.le.s32 p, %2, %3;\n\t"                                                         \
      "and.pred pa, p, act;\n\t"                                                           \
      "not.pred pb, p;\n\t"                                                                \
      "and.pred pb, pb, act;\n\t"                                                          \
      "@pa add.s32 %0, %0, %6;\n\t"                                                        \
      "@pb add.s32 %1, %1, %6;\n\t"                                                        \
      "@pa ld.shared.s32 %2, [%0];\n\t"                                                    \
      "@pb ld.shared.s32 %3, [%1];\n\t"                                                    \
      "@pa xor.b32 %2, %2, %9;\n\t"                                                        \
      "@pb xor.b32 %3, %3, %9;\n\t"                                                        \
      "min.s32 vn, %2, %3;\n\t"                                                            \
      "setp.gt.s32 q, vn, %4;\n\t"                                                         \
      "and.pred q, q, act;\n\t"                                                            \
      "mov.s32 %4, vn;\n\t"                                                                \
      "mad.lo.s32 d, %0, %7, %8;\n\t"                                                      \
      "abs.s32 d, d;\n\t"                                                                  \
      "@q max.s32 %5, %5, d;\n\t}"                                                         \
      : "+r"(fa), "+r"(fb), "+r"(ua), "+r"(ub), "+r"(v), "+r"(dmax)                        \
      : "r"(step), "r"(Tm), "r"(c), "r"(m), "r"(s), "r"(T))

// The KS numerator of one row, computed by its two lanes.  colA/colB: this row's columns
// (element of rank r at col[(r+1)*16]).  h = 0: forward chain, h = 1: backward chain.
__device__ __forceinline__ int nm_pair_walk(const int* colA, const int* colB, int n0, int n1, int h,
                                            int iters) {
  const int T = n0 + n1;
  // lane frame: a' = first group of the frame, b' = second; lane 1 swaps the groups
  const int* ca = h ? colB : colA;
  const int* cb = h ? colA : colB;
  const int na = h ? n1 : n0;           // n0' (size of a'); n1' = T - na
  const int start_row_a = h ? na : 1;   // forward: row 1 (rank 0); backward: row n (rank n-1)
  const int start_row_b = h ? (T - na) : 1;
  const int step = h ? -64 : 64;
  const int m = h ? -1 : 0;
  int fa = (int)nm_smem_u32(ca + start_row_a * NM_PAIR_ROWS);
  int fb = (int)nm_smem_u32(cb + start_row_b * NM_PAIR_ROWS);
  int ua = ca[start_row_a * NM_PAIR_ROWS] ^ m, ub = cb[start_row_b * NM_PAIR_ROWS] ^ m;
  int v = ua < ub ? ua : ub;
  // 64*i' = sgn*(fa - start_a);  64*d' = 64*(i'*T - (s+1)*na) = fa*(sgn*T) - sgn*start_a*T - 64*(s+1)*na
  const int Tm = h ? -T : T;
  int c = -fa * Tm - 64 * na;
  const int k0 = 64 * na;
  int dmax = 0;
  if (__all_sync(0xffffffffu, T >= iters)) {
#pragma unroll 4
    for (int s = 0; s < iters; ++s) {
      NM_PAIR_STEP(fa, fb, ua, ub, v, c, dmax, step, Tm, m);
      c -= k0;
    }
  } else {
#pragma unroll 2
    for (int s = 0; s < iters; ++s) {
      NM_PAIR_STEP_ACT(fa, fb, ua, ub, v, c, dmax, step, Tm, m, s, T);
      c -= k0;
    }
  }
  const int other = __shfl_xor_sync(0xffffffffu, dmax, 1);
  dmax = dmax > other ? dmax : other;
  return dmax >> 6;
}

// Load this lane's half of the row window (pad +inf), sort, cross-merge with the partner lane,
// write the column back transposed.
template <int H>
__device__ __forceinline__ void nm_pair_sort_group(float* region, int base, int n, int lane, int one,
                                                   int mone) {
  int x[H];
  const int h = lane & 1, p = lane >> 1;
  const int hmask = h ? -1 : 0;
  const int shift = base & 3;
  const float4* raw4 = reinterpret_cast<const float4*>(region + (base - shift)) + h * (H / 4);
  const int e0 = h * H - shift;  // window slot of x[0] minus the shift = row index of x[0]
  const bool full = __all_sync(0xffffffffu, (shift == 0) && (n == 2 * H));
  if (__builtin_expect(full, 1)) {
#pragma unroll
    for (int q = 0; q < H / 4; ++q) {
      const float4 v = raw4[q];
      x[4 * q + 0] = nm_make_key(v.x) ^ hmask;
      x[4 * q + 1] = nm_make_key(v.y) ^ hmask;
      x[4 * q + 2] = nm_make_key(v.z) ^ hmask;
      x[4 * q + 3] = nm_make_key(v.w) ^ hmask;
    }
  } else {
    const int pinf = NM_KEY_PINF ^ hmask;
#pragma unroll
    for (int q = 0; q < H / 4; ++q) {
      const float4 v4 = raw4[q];
      const float vv[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const bool valid = (unsigned)(e0 + 4 * q + j) < (unsigned)n;
        x[4 * q + j] = valid ? (nm_make_key(vv[j]) ^ hmask) : pinf;
      }
      if ((q & 3) == 3) asm volatile("" ::: "memory");
    }
  }
  nm_halfsort<H>::run(x, one, mone);
#pragma unroll
  for (int k = 0; k < H; ++k) {
    const int o = ~__shfl_xor_sync(0xffffffffu, x[k], 1);
    x[k] = x[k] < o ? x[k] : o;
  }
  nm_updown<H>::run(x, one, mone);
  __syncwarp();  // every lane has consumed its raw window; the region may now be overwritten
  // slot k (k-th smallest in the lane's space): lane 0 -> rank k, lane 1 -> rank 2H-1-k
  int* col = reinterpret_cast<int*>(region) + p;
  int* dst = col + (h ? 2 * H : 1) * NM_PAIR_ROWS;
  const int dstep = h ? -NM_PAIR_ROWS : NM_PAIR_ROWS;
#pragma unroll
  for (int k = 0; k < H; ++k) dst[k * dstep] = x[nm_updown<H>::order(k)] ^ hmask;
  if (h) col[(2 * H + 1) * NM_PAIR_ROWS] = NM_KEY_PINF; else col[0] = NM_KEY_NINF;
}

template <int H>
__device__ __forceinline__ void nm_pair_tile(float* regA, float* regB, int base0, int base1, int n0,
                                             int n1, int lane, int one, int mone) {
#pragma unroll 1
  for (int g = 0; g < 2; ++g)
    nm_pair_sort_group<H>(g ? regB : regA, g ? base1 : base0, g ? n1 : n0, lane, one, mone);
  __syncwarp();
}

#define NM_DISPATCH_H(hsel, CALL)   \
  switch (hsel) {                   \
    case 4: { CALL(4); } break;     \
    case 8: { CALL(8); } break;     \
    case 12: { CALL(12); } break;   \
    case 16: { CALL(16); } break;   \
    case 20: { CALL(20); } break;   \
    case 24: { CALL(24); } break;   \
    case 28: { CALL(28); } break;   \
    case 32: { CALL(32); } break;   \
    case 36: { CALL(36); } break;   \
    case 40: { CALL(40); } break;   \
    case 44: { CALL(44); } break;   \
    case 48: { CALL(48); } break;   \
    case 52: { CALL(52); } break;   \
    case 56: { CALL(56); } break;   \
    case 60: { CALL(60); } break;   \
    case 64: { CALL(64); } break;   \
    default: { CALL(68); } break;   \
  }

struct nm_pair_meta {
  int64_t r;
  long long s0, s1;
  int n0, n1;
  bool ok;
};

__device__ __forceinline__ nm_pair_meta nm_pair_fetch(const nm_kargs& a, int64_t tile, int lane) {
  nm_pair_meta m;
  m.r = tile * NM_PAIR_ROWS + (lane >> 1);
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

struct nm_pair_stage {
  long long al0, al1;
  unsigned bytes0, bytes1;
  int base0, base1;
  int need, tmax;  // need = max over rows and groups of (alignment shift + n)
  bool any;
};

__device__ __forceinline__ nm_pair_stage nm_pair_plan(const nm_pair_meta& m, int lane) {
  nm_pair_stage st;
  const long long big = 0x7fffffffffffffffLL;
  const bool even = (lane & 1) == 0;
  st.any = __any_sync(0xffffffffu, m.ok);
  const long long first0 = nm_warp_min_ll(m.ok ? m.s0 : big), first1 = nm_warp_min_ll(m.ok ? m.s1 : big);
  const long long end0 = nm_warp_max_ll(m.ok ? m.s0 + m.n0 : -1), end1 = nm_warp_max_ll(m.ok ? m.s1 + m.n1 : -1);
  const int tot0 = __reduce_add_sync(0xffffffffu, even ? m.n0 : 0), tot1 = __reduce_add_sync(0xffffffffu, even ? m.n1 : 0);
  st.tmax = __reduce_max_sync(0xffffffffu, m.n0 + m.n1);
  st.al0 = first0 & ~3LL;
  st.al1 = first1 & ~3LL;
  const bool contig0 = st.any && (end0 - first0) == (long long)tot0;
  const bool contig1 = st.any && (end1 - first1) == (long long)tot1;
  st.bytes0 = contig0 ? (unsigned)(((first0 - st.al0) + tot0 + 3) & ~3LL) * 4u : 0u;
  st.bytes1 = contig1 ? (unsigned)(((first1 - st.al1) + tot1 + 3) & ~3LL) * 4u : 0u;
  st.base0 = m.ok ? (int)(m.s0 - st.al0) : 0;
  st.base1 = m.ok ? (int)(m.s1 - st.al1) : 0;
  // gathered rows are placed at 4-float boundaries (shift 0), bulk-copied rows keep their alignment
  const int w0 = m.ok ? m.n0 + (st.bytes0 ? (st.base0 & 3) : 0) : 0;
  const int w1 = m.ok ? m.n1 + (st.bytes1 ? (st.base1 & 3) : 0) : 0;
  st.need = __reduce_max_sync(0xffffffffu, w0 > w1 ? w0 : w1);
  return st;
}

// cooperative copy of the tile's rows when they are not one contiguous slice: row p goes to a
// 4-float aligned offset (prefix of the padded lengths)
__device__ __forceinline__ int nm_pair_gather(float* region, const float* __restrict__ vals, long long start,
                                              int n, int lane) {
  const int padded = (lane & 1) ? 0 : ((n + 3) & ~3);
  int incl = padded;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  int base = incl - padded;
  base = __shfl_sync(0xffffffffu, base, lane & ~1);  // both lanes of the pair use the even lane's offset
  for (int rl = 0; rl < 32; rl += 2) {
    const int rn = __shfl_sync(0xffffffffu, n, rl);
    const int rbase = __shfl_sync(0xffffffffu, base, rl);
    const long long rstart = __shfl_sync(0xffffffffu, start, rl);
    for (int k = lane; k < rn; k += 32) region[rbase + k] = vals[rstart + k];
  }
  return base;
}

__global__ void __launch_bounds__(32 * NM_PAIR_MAX_WARPS, 4)
nm_pair_kernel(const nm_kargs a, const int want_t) {
  extern __shared__ __align__(128) unsigned char nm_smem[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int h = lane & 1, p = lane >> 1;
  unsigned char* my = nm_smem + (size_t)wib * (16 + 2 * (size_t)a.region_floats * sizeof(float));
  uint64_t* bar = reinterpret_cast<uint64_t*>(my);
  float* regA = reinterpret_cast<float*>(my + 16);
  float* regB = regA + a.region_floats;
  const int64_t n_tiles = (a.n_rows + NM_PAIR_ROWS - 1) / NM_PAIR_ROWS;
  const int warps_per_cta = blockDim.x >> 5;
  const int64_t n_warps = (int64_t)gridDim.x * warps_per_cta;

  int64_t tile = (int64_t)blockIdx.x * warps_per_cta + wib;
  if (tile >= n_tiles) return;
  if (lane == 0) nm_mbar_init(bar, 1);
  __syncwarp();
  unsigned parity = 0;

  nm_pair_meta cur = nm_pair_fetch(a, tile, lane);
  nm_pair_stage cst = nm_pair_plan(cur, lane);
  bool staged = false;

  while (true) {
    int64_t next = -1;
    {
      long long t = 0;
      if (lane == 0) t = (long long)atomicAdd(a.tile_cursor, 1) + n_warps;
      t = __shfl_sync(0xffffffffu, t, 0);
      next = t < n_tiles ? t : -1;
    }
    const nm_pair_meta nxt = nm_pair_fetch(a, next, lane);

    if (cst.any) {
      if (!staged && (cst.bytes0 | cst.bytes1) && lane == 0) {
        nm_mbar_expect_tx(bar, cst.bytes0 + cst.bytes1);
        if (cst.bytes0) nm_bulk_g2s(regA, a.vals0 + cst.al0, cst.bytes0, bar);
        if (cst.bytes1) nm_bulk_g2s(regB, a.vals1 + cst.al1, cst.bytes1, bar);
      }
      int base0 = cst.base0, base1 = cst.base1;
      if (!cst.bytes0) base0 = nm_pair_gather(regA, a.vals0, cur.s0, cur.n0, lane);
      if (!cst.bytes1) base1 = nm_pair_gather(regB, a.vals1, cur.s1, cur.n1, lane);
      const nm_pair_stage nst = nm_pair_plan(nxt, lane);
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
      double mean_g = 0.0, var_g = 0.0;
      if (want_t)  // lane 0 of a pair: group 0's moments, lane 1: group 1's (same routine as the lane tier)
        nm_lane_moments(h ? regB : regA, h ? base1 : base0, h ? n1 : n0, &mean_g, &var_g);
      int hsel = ((cst.need + 1) / 2 + 3) & ~3;
      if (hsel < 4) hsel = 4;
      if (2 * hsel > a.class_n) hsel = a.class_n;  // one hot code path per call (instruction cache)
#define NM_CALL(HH) nm_pair_tile<HH>(regA, regB, base0, base1, n0, n1, lane, a.one, a.mone)
      NM_DISPATCH_H(hsel, NM_CALL)
#undef NM_CALL

      const int* colA = reinterpret_cast<const int*>(regA) + p;
      const int* colB = reinterpret_cast<const int*>(regB) + p;
      const int iters = (cst.tmax + 1) >> 1;
      const int dnum = nm_pair_walk(colA, colB, n0, n1, h, iters);

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

      // lane 0 of the pair does the KS tail; the Welch tail runs in nm_tails_kernel
      if (cur.ok) {
        if (h == 0) {
          double d, pv;
          nm_ks_tail(dnum, n0, n1, &d, &pv);
          a.ks_dnum[cur.r] = dnum;
          if (a.ks_d) a.ks_d[cur.r] = d;
          a.ks_p[cur.r] = pv;
          if (a.flags) a.flags[cur.r] = 0;
        }
        if (want_t) {  // lane 0: group 0's moments, lane 1: group 1's; tails in nm_tails_kernel
          double2* mom = reinterpret_cast<double2*>(a.acc_mom) + 2 * cur.r + h;
          *mom = make_double2(mean_g, var_g);
        }
      }
      cur = nxt;
      cst = nst;
    } else {
      cur = nxt;
      cst = nm_pair_plan(nxt, lane);
      staged = false;
    }
    if (next < 0) break;
  }
}

bool nm_pair_tier_available() { return true; }

int nm_launch_pair(const nm_kargs& ka_in, bool want_t, int max_n, int sm_count, cudaStream_t st) {
  nm_kargs ka = ka_in;
  // per-lane half size of the longest row's window (row + up to 3 floats of alignment shift)
  int hcls = (((max_n + 3) + 1) / 2 + 3) & ~3;
  if (hcls > NM_PAIR_MAX_H) hcls = NM_PAIR_MAX_H;
  ka.region_floats = NM_PAIR_ROWS * (2 * hcls + 2);
  ka.class_n = hcls;
  const int per_warp = 16 + 2 * ka.region_floats * (int)sizeof(float);
  cudaError_t e = cudaFuncSetAttribute(nm_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       NM_PAIR_MAX_WARPS * per_warp);
  if (e != cudaSuccess) return (int)e;
  int best_w = 1, best_blocks = 0, best_warps = 0;
  for (int w = NM_PAIR_MAX_WARPS; w >= 1; --w) {
    int blocks = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks, nm_pair_kernel, 32 * w, (size_t)w * per_warp);
    if (e != cudaSuccess) return (int)e;
    if (blocks * w > best_warps) {
      best_warps = blocks * w;
      best_w = w;
      best_blocks = blocks;
    }
  }
  if (best_warps == 0) return (int)cudaErrorInvalidConfiguration;
  const int64_t tiles = (ka.n_rows + NM_PAIR_ROWS - 1) / NM_PAIR_ROWS;
  int64_t grid = (tiles + best_w - 1) / best_w;
  const int64_t resident = (int64_t)best_blocks * sm_count;
  if (grid > resident) grid = resident;
  nm_pair_kernel<<<(unsigned)grid, 32 * best_w, (size_t)best_w * per_warp, st>>>(ka, want_t ? 1 : 0);
  return (int)cudaGetLastError();
}

#else  // float keys: the pair tier relies on the int32 key space (u = ~key reverses the order)

bool nm_pair_tier_available() { return false; }
int nm_launch_pair(const nm_kargs&, bool, int, int, cudaStream_t) { return (int)cudaErrorNotSupported; }

#endif
