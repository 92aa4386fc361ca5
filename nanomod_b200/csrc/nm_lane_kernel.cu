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

#define NM_LANE_MAX_WARPS 4  // warps (= independent 32-row tiles) per CTA, chosen at launch

// Visit the row's values through 128-bit shared-memory loads from the 16-byte aligned address
// below the row (N/4 + 1 loads): window slot e holds row element e - shift and is valid iff
// 0 <= e - shift < n.  BODY sees (e, j_ = e & 3, valid, v).
#define NM_FOR_ROW(NQ, raw4, shift, n, BODY)                                     \
  _Pragma("unroll") for (int q_ = 0; q_ < (NQ); ++q_) {                          \
    const float4 v4_ = (raw4)[q_];                                               \
    const float vv_[4] = {v4_.x, v4_.y, v4_.z, v4_.w};                           \
    _Pragma("unroll") for (int j_ = 0; j_ < 4; ++j_) {                           \
      const int e = 4 * q_ + j_;                                                 \
      const bool valid = (unsigned)(e - (shift)) < (unsigned)(n);                \
      const float v = vv_[j_];                                                   \
      BODY                                                                       \
    }                                                                            \
    /* keep ptxas from hoisting every load above the first use (register pressure) */ \
    if ((q_ & 3) == 3) asm volatile("" ::: "memory");                            \
  }

// Load one group's row into registers (pad +inf), sort, write back transposed with a -inf row
// in front and a +inf sentinel row behind (layout expected by nm_merge_walk).
#ifndef NM_KEY_ALU  // default: the negation of negative values as an IMAD (FMA pipe), 1 % faster
#define NM_MK(v) nm_make_key(v, mone)
#else
#define NM_MK(v) nm_make_key(v)
#endif
template <int N>
__device__ __forceinline__ void nm_lane_sort_group(float* region, int base, int n, int lane,
                                                   int one, int mone) {
  nm_key x[N];
  const int shift = base & 3;
  const float4* raw4 = reinterpret_cast<const float4*>(region + (base - shift));
  const bool full = __all_sync(0xffffffffu, (shift == 0) && (n == N));
  if (__builtin_expect(full, 1)) {
#pragma unroll
    for (int q = 0; q < N / 4; ++q) {
      const float4 v = raw4[q];
      x[4 * q + 0] = NM_MK(v.x);
      x[4 * q + 1] = NM_MK(v.y);
      x[4 * q + 2] = NM_MK(v.z);
      x[4 * q + 3] = NM_MK(v.w);
    }
  } else {
    NM_FOR_ROW(N / 4, raw4, shift, n, { x[e] = valid ? NM_MK(v) : (nm_key)NM_KEY_PINF; })
    // the last (up to 3) elements of a long shifted row lie in window slots N..N+2; slots
    // 0..shift-1 are free exactly then (N + j valid  =>  j < shift), so they wrap around
    const float4 w4 = raw4[N / 4];
    const float ww[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const bool wrap = (unsigned)(N + j - shift) < (unsigned)n;
      x[j] = wrap ? NM_MK(ww[j]) : x[j];
    }
  }
  nm_sorter<N>::run(x, one, mone);
  __syncwarp();  // every lane has consumed its raw row; the region may now be overwritten
  nm_key* col = reinterpret_cast<nm_key*>(region) + lane;
  col[0] = NM_KEY_NINF;
#pragma unroll
  for (int k = 0; k < N; ++k) col[(k + 1) << 5] = x[nm_sorter<N>::order(k)];
  col[(N + 1) << 5] = NM_KEY_PINF;
}

template <int N>
__device__ __forceinline__ void nm_lane_tile(float* regA, float* regB, int base0, int base1,
                                             int n0, int n1, int lane, int one, int mone) {
#pragma unroll 1
  for (int g = 0; g < 2; ++g)
    nm_lane_sort_group<N>(g ? regB : regA, g ? base1 : base0, g ? n1 : n0, lane, one, mone);
  __syncwarp();
}

// ------------------------------------------------------------------------------------------
// Device fast path of nm_merge_walk (KS numerator only): no activity predicates,
// pointer-chasing with predicated adds and loads, and the ECDF numerator carried on the FMA
// pipe.  Same algorithm, same tie rule, same
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

// pointer bump of a chain: on the ALU pipe, or (NM_WALK_IMAD) as a multiply-add by a runtime 1
#ifdef NM_WALK_IMAD
#define NM_PTR_ADD(pred, reg, imm) pred " mad.lo.s32 " reg ", %8, " imm ", " reg ";\n\t"
#else
#define NM_PTR_ADD(pred, reg, imm) pred " add.s32 " reg ", " reg ", " imm ";\n\t"
#endif

// one forward step: take the smaller head (ties: group 0), reload it, evaluate at a tie-group end
#define NM_FWD_STEP(fa, fb, va, vb, v, c, dmax, one)                                      \
  asm volatile(                                                                      \
      "{\n\t.reg .pred p, q;\n\t.reg .s32 d;\n\t.reg ." NM_KT " vn;\n\t"              \
      "setp.le." NM_KT " p, %2, %3;\n\t"                                             \
      NM_PTR_ADD("@p", "%0", "128") NM_PTR_ADD("@!p", "%1", "128")                  \
      "@p ld.shared." NM_KT " %2, [%0];\n\t"                                         \
      "@!p ld.shared." NM_KT " %3, [%1];\n\t"                                        \
      "min." NM_KT " vn, %2, %3;\n\t"                                                \
      "setp.gt." NM_KT " q, vn, %4;\n\t"                                             \
      "mov." NM_KT " %4, vn;\n\t"                                                    \
      "mad.lo.s32 d, %0, %6, %7;\n\t"                                                \
      "abs.s32 d, d;\n\t"                                                            \
      "@q max.s32 %5, %5, d;\n\t}"                                                   \
      : "+r"(fa), "+r"(fb), "+" NM_KREG(va), "+" NM_KREG(vb), "+" NM_KREG(v), "+r"(dmax) \
      : "r"(T), "r"(c), "r"(one))

// one backward step: take the larger tail (ties: group 1), reload it, evaluate at a group start
#define NM_BWD_STEP(ba, bb, ea, eb, w, c, dmax, one)                                      \
  asm volatile(                                                                      \
      "{\n\t.reg .pred p, q;\n\t.reg .s32 d;\n\t.reg ." NM_KT " wn;\n\t"              \
      "setp.ge." NM_KT " p, %3, %2;\n\t"                                             \
      NM_PTR_ADD("@p", "%1", "-128") NM_PTR_ADD("@!p", "%0", "-128")                \
      "@p ld.shared." NM_KT " %3, [%1];\n\t"                                         \
      "@!p ld.shared." NM_KT " %2, [%0];\n\t"                                        \
      "max." NM_KT " wn, %2, %3;\n\t"                                                \
      "setp.lt." NM_KT " q, wn, %4;\n\t"                                             \
      "mov." NM_KT " %4, wn;\n\t"                                                    \
      "mad.lo.s32 d, %0, %6, %7;\n\t"                                                \
      "abs.s32 d, d;\n\t"                                                            \
      "@q max.s32 %5, %5, d;\n\t}"                                                   \
      : "+r"(ba), "+r"(bb), "+" NM_KREG(ea), "+" NM_KREG(eb), "+" NM_KREG(w), "+r"(dmax) \
      : "r"(T), "r"(c), "r"(one))

// Both chains run `iters` (warp-uniform, >= ceil(T/2) for every lane) steps, so on lanes whose
// T is smaller than the warp maximum they overlap in the middle; every evaluated point is still
// a genuine tie-group boundary, so the maximum is unchanged.  Requires iters <= T on every lane
// (a chain must not run off the end of the columns).
__device__ __forceinline__ int nm_walk_ks_fast(const nm_key* colA, const nm_key* colB, int n0,
                                               int n1, int iters, int one) {
  const int T = n0 + n1;
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
  for (int s = 0; s < iters; ++s) {
    NM_FWD_STEP(fa, fb, va, vb, v, cf, dmax, one);
    NM_BWD_STEP(ba, bb, ea, eb, w, cb, dmax, one);
    cf -= k0;
    cb += k0;
  }
  return dmax >> 7;
}

// Four chains: the pooled order is split at h = T/2 by a merge-path search (ties: group 0 first,
// the order both chain kinds assume), then a forward and a backward chain work on each half.
// The walk is latency bound (a step is compare -> pointer bump -> shared load -> compare), so
// halving the chain length matters more than the ~60 instructions of the search.  `it` is
// warp-uniform with 2*it >= ceil(T/2) and it <= T/2 on every lane.
__device__ __forceinline__ int nm_walk_ks_fast4(const nm_key* colA, const nm_key* colB, int n0,
                                                int n1, int it, int search_iters, int one) {
  const int T = n0 + n1, h = T >> 1;
  const unsigned A0 = nm_smem_u32(colA), B0 = nm_smem_u32(colB);
  // smallest i with b[h-i-1] < a[i]: i elements of group 0 among the first h pooled ones
  // (row k+1 of a column holds element k; row 0 is -inf and row n+1 is +inf)
  int lo = h - n1 > 0 ? h - n1 : 0, hi = h < n0 ? h : n0;
#pragma unroll 1
  for (int k = 0; k < search_iters; ++k) {
    const int mid = (lo + hi) >> 1;
    const bool P = colB[(h - mid) << 5] < colA[(mid + 1) << 5];
    hi = P ? mid : hi;
    lo = P ? lo : mid + 1;
  }
  const int is = lo, js = h - lo;
  unsigned f1a = A0 + 128u, f1b = B0 + 128u;
  unsigned b1a = A0 + 128u * (unsigned)is, b1b = B0 + 128u * (unsigned)js;
  unsigned f2a = b1a + 128u, f2b = b1b + 128u;
  unsigned b2a = A0 + 128u * (unsigned)n0, b2b = B0 + 128u * (unsigned)n1;
  nm_key v1a = colA[32], v1b = colB[32], v1 = nm_min(v1a, v1b);
  nm_key e1a = colA[is << 5], e1b = colB[js << 5], w1 = nm_max(e1a, e1b);
  nm_key v2a = colA[(is + 1) << 5], v2b = colB[(js + 1) << 5], v2 = nm_min(v2a, v2b);
  nm_key e2a = colA[n0 << 5], e2b = colB[n1 << 5], w2 = nm_max(e2a, e2b);
  const int k0 = 128 * n0;
  // forward, m elements consumed:  128*d = (fa - A0 - 128)*T - k0*m ; backward, r remaining:
  // 128*d = (ba - A0)*T - k0*r
  const int cF = -(int)((A0 + 128u) * (unsigned)T), cB = -(int)(A0 * (unsigned)T);
  int cf1 = cF - k0, cb1 = cB - k0 * (h - 1), cf2 = cF - k0 * (h + 1), cb2 = cB - k0 * (T - 1);
  // the boundary between pooled elements h-1 and h belongs to neither half's chains
  int dmax = 0;
  {
    int dj = 128 * (is * T - h * n0);
    dj = dj < 0 ? -dj : dj;
    if (w1 < v2) dmax = dj;
  }
#pragma unroll 2
  for (int s = 0; s < it; ++s) {
    NM_FWD_STEP(f1a, f1b, v1a, v1b, v1, cf1, dmax, one);
    NM_BWD_STEP(b1a, b1b, e1a, e1b, w1, cb1, dmax, one);
    NM_FWD_STEP(f2a, f2b, v2a, v2b, v2, cf2, dmax, one);
    NM_BWD_STEP(b2a, b2b, e2a, e2b, w2, cb2, dmax, one);
    cf1 -= k0;
    cb1 += k0;
    cf2 -= k0;
    cb2 += k0;
  }
  return dmax >> 7;
}

// ---- the same two walks over 16-bit grid-key columns: a column is one half of the 32-bit words
// of region A (row stride 128 bytes as before; group 1 lives 2 bytes above group 0), keys are
// zero-extended into 32-bit registers and compared as unsigned integers.
#ifdef NM_WALK16_IMAD  // pointer bump of a chain as a multiply-add by a runtime 1 (FMA pipe instead of ALU)
#define NM_PTR_ADD16(pred, reg, imm) pred " mad.lo.s32 " reg ", %8, " imm ", " reg ";\n\t"
#else
#define NM_PTR_ADD16(pred, reg, imm) pred " add.s32 " reg ", " reg ", " imm ";\n\t"
#endif
#define NM_FWD_STEP16(fa, fb, va, vb, v, c, dmax, one)                                    \
  asm volatile(                                                                      \
      "{\n\t.reg .pred p, q;\n\t.reg .s32 d;\n\t.reg .u32 vn;\n\t"                    \
      "setp.le.u32 p, %2, %3;\n\t"                                                   \
      NM_PTR_ADD16("@p", "%0", "128") NM_PTR_ADD16("@!p", "%1", "128")                  \
      "@p ld.shared.u16 %2, [%0];\n\t"                                               \
      "@!p ld.shared.u16 %3, [%1];\n\t"                                              \
      "min.u32 vn, %2, %3;\n\t"                                                      \
      "setp.gt.u32 q, vn, %4;\n\t"                                                   \
      "mov.u32 %4, vn;\n\t"                                                          \
      "mad.lo.s32 d, %0, %6, %7;\n\t"                                                \
      "abs.s32 d, d;\n\t"                                                            \
      "@q max.s32 %5, %5, d;\n\t}"                                                   \
      : "+r"(fa), "+r"(fb), "+r"(va), "+r"(vb), "+r"(v), "+r"(dmax)                  \
      : "r"(T), "r"(c), "r"(one))

#define NM_BWD_STEP16(ba, bb, ea, eb, w, c, dmax, one)                                    \
  asm volatile(                                                                      \
      "{\n\t.reg .pred p, q;\n\t.reg .s32 d;\n\t.reg .u32 wn;\n\t"                    \
      "setp.ge.u32 p, %3, %2;\n\t"                                                   \
      NM_PTR_ADD16("@p", "%1", "-128") NM_PTR_ADD16("@!p", "%0", "-128")                \
      "@p ld.shared.u16 %3, [%1];\n\t"                                               \
      "@!p ld.shared.u16 %2, [%0];\n\t"                                              \
      "max.u32 wn, %2, %3;\n\t"                                                      \
      "setp.lt.u32 q, wn, %4;\n\t"                                                   \
      "mov.u32 %4, wn;\n\t"                                                          \
      "mad.lo.s32 d, %0, %6, %7;\n\t"                                                \
      "abs.s32 d, d;\n\t"                                                            \
      "@q max.s32 %5, %5, d;\n\t}"                                                   \
      : "+r"(ba), "+r"(bb), "+r"(ea), "+r"(eb), "+r"(w), "+r"(dmax)                  \
      : "r"(T), "r"(c), "r"(one))

__device__ __forceinline__ unsigned nm_umin(unsigned a, unsigned b) { return a < b ? a : b; }
__device__ __forceinline__ unsigned nm_umax(unsigned a, unsigned b) { return a > b ? a : b; }

// colA / colB: the lane's two 16-bit columns (element stride 64)
__device__ __forceinline__ int nm_walk_ks_fast_g(const unsigned short* colA, const unsigned short* colB, int n0,
                                                 int n1, int iters, int one) {
  const int T = n0 + n1;
  const unsigned A0 = nm_smem_u32(colA), B0 = nm_smem_u32(colB);
  unsigned fa = A0 + 128u, fb = B0 + 128u;
  unsigned ba = A0 + 128u * (unsigned)n0, bb = B0 + 128u * (unsigned)n1;
  unsigned va = colA[64], vb = colB[64], v = nm_umin(va, vb);
  unsigned ea = colA[n0 << 6], eb = colB[n1 << 6], w = nm_umax(ea, eb);
  const int k0 = 128 * n0;
  int cf = -(int)((A0 + 128u) * (unsigned)T) - k0;
  int cb = -(int)(A0 * (unsigned)T) - k0 * (T - 1);
  int dmax = 0;
#pragma unroll 4
  for (int s = 0; s < iters; ++s) {
    NM_FWD_STEP16(fa, fb, va, vb, v, cf, dmax, one);
    NM_BWD_STEP16(ba, bb, ea, eb, w, cb, dmax, one);
    cf -= k0;
    cb += k0;
  }
  return dmax >> 7;
}

__device__ __forceinline__ int nm_walk_ks_fast4_g(const unsigned short* colA, const unsigned short* colB, int n0,
                                                  int n1, int it, int search_iters, int one) {
  const int T = n0 + n1, h = T >> 1;
  const unsigned A0 = nm_smem_u32(colA), B0 = nm_smem_u32(colB);
  int lo = h - n1 > 0 ? h - n1 : 0, hi = h < n0 ? h : n0;
#pragma unroll 1
  for (int k = 0; k < search_iters; ++k) {
    const int mid = (lo + hi) >> 1;
    const bool P = colB[(h - mid) << 6] < colA[(mid + 1) << 6];
    hi = P ? mid : hi;
    lo = P ? lo : mid + 1;
  }
  const int is = lo, js = h - lo;
  unsigned f1a = A0 + 128u, f1b = B0 + 128u;
  unsigned b1a = A0 + 128u * (unsigned)is, b1b = B0 + 128u * (unsigned)js;
  unsigned f2a = b1a + 128u, f2b = b1b + 128u;
  unsigned b2a = A0 + 128u * (unsigned)n0, b2b = B0 + 128u * (unsigned)n1;
  unsigned v1a = colA[64], v1b = colB[64], v1 = nm_umin(v1a, v1b);
  unsigned e1a = colA[is << 6], e1b = colB[js << 6], w1 = nm_umax(e1a, e1b);
  unsigned v2a = colA[(is + 1) << 6], v2b = colB[(js + 1) << 6], v2 = nm_umin(v2a, v2b);
  unsigned e2a = colA[n0 << 6], e2b = colB[n1 << 6], w2 = nm_umax(e2a, e2b);
  const int k0 = 128 * n0;
  const int cF = -(int)((A0 + 128u) * (unsigned)T), cB = -(int)(A0 * (unsigned)T);
  int cf1 = cF - k0, cb1 = cB - k0 * (h - 1), cf2 = cF - k0 * (h + 1), cb2 = cB - k0 * (T - 1);
  int dmax = 0;
  {
    int dj = 128 * (is * T - h * n0);
    dj = dj < 0 ? -dj : dj;
    if (w1 < v2) dmax = dj;
  }
#pragma unroll 2
  for (int s = 0; s < it; ++s) {
    NM_FWD_STEP16(f1a, f1b, v1a, v1b, v1, cf1, dmax, one);
    NM_BWD_STEP16(b1a, b1b, e1a, e1b, w1, cb1, dmax, one);
    NM_FWD_STEP16(f2a, f2b, v2a, v2b, v2, cf2, dmax, one);
    NM_BWD_STEP16(b2a, b2b, e2a, e2b, w2, cb2, dmax, one);
    cf1 -= k0;
    cb1 += k0;
    cf2 -= k0;
    cb2 += k0;
  }
  return dmax >> 7;
}

// Per-lane description of one row of a tile (tile = 32 rows of the launch's row list, one per lane).
struct nm_tile_meta {
  int64_t r;        // row index
  long long s0, s1; // start of the row's slices in vals0 / vals1
  int n0, n1;       // coverage (0 when the lane has no lane-tier row)
  bool ok;
};

__device__ __forceinline__ nm_tile_meta nm_tile_fetch(const nm_kargs& a, int64_t tile, int lane) {
  nm_tile_meta m;
  m.r = 0;
  m.n0 = m.n1 = 0;
  m.s0 = m.s1 = 0;
  m.ok = false;
  const int64_t idx = a.row_lo + tile * 32 + lane;
  if (tile >= 0 && idx < a.row_hi) {
    m.r = a.perm ? (int64_t)a.perm[idx] : idx;
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

// Two-step fetch for the short-row instantiation (registers to spare): the metadata of a tile is
// fetched one tile apart -- A: row index, n0, n1, candidate index (addresses depend on the tile id
// only); B: the two CSR offsets (addresses depend on A's candidate index) -- so that no load is
// consumed in the iteration that issues it.  A profile of the one-step version showed 7.5 % of
// the kernel's time waiting on these loads at the top of the loop.  (For long rows the extra
// live registers cost the N=100 sort more than the waits: 1.86 -> 2.02 ms, so they keep one step.)
struct nm_tile_head {
  int64_t r;
  int n0, n1;
  int32_t src;
  bool in_range;
};

__device__ __forceinline__ nm_tile_head nm_tile_fetch_a(const nm_kargs& a, int64_t tile, int lane) {
  nm_tile_head h;
  h.r = 0;
  h.n0 = h.n1 = 0;
  h.src = 0;
  const int64_t idx = a.row_lo + tile * 32 + lane;
  h.in_range = tile >= 0 && idx < a.row_hi;
  if (h.in_range) {
    h.r = a.perm ? (int64_t)a.perm[idx] : idx;
    h.n0 = a.row_n0[h.r];
    h.n1 = a.row_n1[h.r];
    h.src = a.row_pos_index[h.r];
  }
  return h;
}

__device__ __forceinline__ nm_tile_meta nm_tile_fetch_b(const nm_kargs& a, const nm_tile_head& h) {
  nm_tile_meta m;
  m.r = h.r;
  m.ok = h.in_range && h.n0 <= NM_LANE_TIER_MAX && h.n1 <= NM_LANE_TIER_MAX;
  m.n0 = m.ok ? h.n0 : 0;
  m.n1 = m.ok ? h.n1 : 0;
  m.s0 = m.s1 = 0;
  if (h.in_range) {
    m.s0 = a.off0[h.src];
    m.s1 = a.off1[h.src];
  }
  return m;
}

// Warp-level plan for staging a tile.  If a group's 32 slices form one contiguous run of the CSR
// array (the normal case when nothing was filtered and the rows are not class-binned) lane 0
// copies it in one piece with a TMA bulk copy; otherwise every lane copies its own row's slice
// (rounded out to 16-byte boundaries) into its own slot of the region with 16-byte cp.async.
struct nm_tile_stage {
  long long al0, al1;      // aligned source start: of the tile (contiguous) or of this lane's slice
  unsigned bytes0, bytes1; // copy size: of the tile (contiguous, lane 0 issues) or of this lane's slice
  int dst0, dst1;          // float offset of this lane's slot in region A / B (scattered mode; else 0)
  int base0, base1;        // this lane's row offset inside region A / B
  unsigned total;          // bytes the tile's mbarrier phase has to see (contiguous groups only)
  int nmax, tmax;
  bool any, contig0, contig1;
};

__device__ __forceinline__ int nm_warp_excl_sum(int v, int lane) {
  int inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  return inc - v;
}

__device__ __forceinline__ nm_tile_stage nm_tile_plan(const nm_tile_meta& m, int lane, int region_floats) {
  nm_tile_stage st;
  const long long big = 0x7fffffffffffffffLL;
  st.any = __any_sync(0xffffffffu, m.ok);
  const long long first0 = nm_warp_min_ll(m.ok ? m.s0 : big), first1 = nm_warp_min_ll(m.ok ? m.s1 : big);
  const long long end0 = nm_warp_max_ll(m.ok ? m.s0 + m.n0 : -1), end1 = nm_warp_max_ll(m.ok ? m.s1 + m.n1 : -1);
  st.tmax = __reduce_max_sync(0xffffffffu, m.n0 + m.n1);
  st.nmax = __reduce_max_sync(0xffffffffu, m.n0 > m.n1 ? m.n0 : m.n1);
  // "contiguous": the span from the first to the last slice fits the region (rows skipped in
  // between -- filtered, deep or binned elsewhere -- are copied along and ignored)
  st.contig0 = st.any && (end0 - (first0 & ~3LL)) + 3 <= (long long)region_floats;
  st.contig1 = st.any && (end1 - (first1 & ~3LL)) + 3 <= (long long)region_floats;
  st.dst0 = st.dst1 = 0;
  if (st.contig0) {
    st.al0 = first0 & ~3LL;
    st.bytes0 = (unsigned)((end0 - st.al0 + 3) & ~3LL) * 4u;
    st.base0 = m.ok ? (int)(m.s0 - st.al0) : 0;
  } else {
    st.al0 = m.s0 & ~3LL;
    const int len4 = m.ok ? (int)(((m.s0 - st.al0) + m.n0 + 3) & ~3LL) : 0;
    st.bytes0 = (unsigned)len4 * 4u;
    st.dst0 = nm_warp_excl_sum(len4, lane);
    st.base0 = st.dst0 + (int)(m.s0 - st.al0);
  }
  if (st.contig1) {
    st.al1 = first1 & ~3LL;
    st.bytes1 = (unsigned)((end1 - st.al1 + 3) & ~3LL) * 4u;
    st.base1 = m.ok ? (int)(m.s1 - st.al1) : 0;
  } else {
    st.al1 = m.s1 & ~3LL;
    const int len4 = m.ok ? (int)(((m.s1 - st.al1) + m.n1 + 3) & ~3LL) : 0;
    st.bytes1 = (unsigned)len4 * 4u;
    st.dst1 = nm_warp_excl_sum(len4, lane);
    st.base1 = st.dst1 + (int)(m.s1 - st.al1);
  }
  st.total = (st.contig0 ? st.bytes0 : 0u) + (st.contig1 ? st.bytes1 : 0u);
  return st;
}

// scattered rows: the warp copies one row's slice per step, 16 bytes per lane (coalesced; a
// per-lane loop over its own row would touch 32 different lines per instruction)
// (kept out of line: it is the rare path and must not lengthen the hot loop's code)
__device__ __noinline__ void nm_rows_copy_async(float* region, const float* __restrict__ vals, long long al,
                                                int dst, unsigned bytes, int lane) {
#pragma unroll 1
  for (int r = 0; r < 32; ++r) {
    const long long ral = __shfl_sync(0xffffffffu, al, r);
    const int rdst = __shfl_sync(0xffffffffu, dst, r);
    const unsigned rbytes = __shfl_sync(0xffffffffu, bytes, r);
    for (unsigned o = 16u * (unsigned)lane; o < rbytes; o += 512u)
      nm_cp_async16(reinterpret_cast<unsigned char*>(region + rdst) + o,
                    reinterpret_cast<const unsigned char*>(vals + ral) + o);
  }
}

// issue the tile's copies: bulk copies complete on `bar`, per-lane copies on the lane's
// cp.async group (nm_tile_wait)
__device__ __forceinline__ void nm_tile_issue(const nm_kargs& a, const nm_tile_stage& st, float* regA, float* regB,
                                              uint64_t* bar, int lane) {
  if (lane == 0 && st.total) {
    nm_mbar_expect_tx(bar, st.total);
    if (st.contig0 && st.bytes0) nm_bulk_g2s(regA, a.vals0 + st.al0, st.bytes0, bar);
    if (st.contig1 && st.bytes1) nm_bulk_g2s(regB, a.vals1 + st.al1, st.bytes1, bar);
  }
  if (!st.contig0) nm_rows_copy_async(regA, a.vals0, st.al0, st.dst0, st.bytes0, lane);
  if (!st.contig1) nm_rows_copy_async(regB, a.vals1, st.al1, st.dst1, st.bytes1, lane);
  if (!(st.contig0 && st.contig1)) nm_cp_async_commit();
}

__device__ __forceinline__ void nm_tile_wait(const nm_tile_stage& st, uint64_t* bar, unsigned& parity) {
  if (st.total) {
    nm_mbar_wait(bar, parity);
    parity ^= 1u;
  }
  if (!(st.contig0 && st.contig1)) nm_cp_async_wait_all();
  __syncwarp();
}

// pull the tile's value slices into L2 ahead of time
__device__ __forceinline__ void nm_tile_prefetch(const nm_kargs& a, const nm_tile_stage& st, int lane) {
  if (lane == 0) {
    if (st.contig0 && st.bytes0) nm_prefetch_l2(a.vals0 + st.al0, st.bytes0);
    if (st.contig1 && st.bytes1) nm_prefetch_l2(a.vals1 + st.al1, st.bytes1);
  }
}

// Persistent lane-tier kernel: every warp loops over tiles taken from a global cursor.  While
// a tile is being sorted, the next tile's metadata is already in registers and its value
// slices are being pulled into L2; its TMA copies are issued the moment the merge walk has
// released the two regions, so that they overlap the fp64 tails of the current tile.
template <int NMAX>
__global__ void __launch_bounds__(32 * NM_LANE_MAX_WARPS, NMAX <= 64 ? 4 : 2)
nm_lane_kernel(const nm_kargs a, const int want_u, const int want_t) {
  extern __shared__ __align__(128) unsigned char nm_smem[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  // per-warp slice of shared memory: [mbarrier 16 B][region A][region B]
  unsigned char* my = nm_smem + (size_t)wib * (16 + 2 * (size_t)a.region_floats * sizeof(float));
  uint64_t* bar = reinterpret_cast<uint64_t*>(my);
  float* regA = reinterpret_cast<float*>(my + 16);
  float* regB = regA + a.region_floats;
  const int64_t n_tiles = (a.row_hi - a.row_lo + 31) >> 5;
  const int warps_per_cta = blockDim.x >> 5;
  const int64_t n_warps = (int64_t)gridDim.x * warps_per_cta;

  const int64_t tile = (int64_t)blockIdx.x * warps_per_cta + wib;
  if (tile >= n_tiles) return;
  if (lane == 0) nm_mbar_init(bar, 1);
  __syncwarp();
  unsigned parity = 0;

  nm_tile_meta cur = nm_tile_fetch(a, tile, lane);
  nm_tile_stage cst = nm_tile_plan(cur, lane, a.region_floats);
  bool staged = false;  // TMA for `cur` already issued?
  // The tile after next is claimed one iteration early: the atomic's round trip and the two
  // dependent metadata loads of a tile each get a whole tile of work to land.
  long long claim = 0;
  if (lane == 0) claim = (long long)atomicAdd(a.tile_cursor, 1) + n_warps;
  constexpr bool kDeep = NMAX <= 64;  // two-step metadata pipeline (see nm_tile_fetch_a)
  nm_tile_head hn;
  int64_t tn = -1;  // kDeep: tile whose step A is in `hn`
  if constexpr (kDeep) {
    const long long t = __shfl_sync(0xffffffffu, claim, 0);
    tn = t < n_tiles ? t : -1;
    if (tn >= 0 && lane == 0) claim = (long long)atomicAdd(a.tile_cursor, 1) + n_warps;
    hn = nm_tile_fetch_a(a, tn, lane);
  }

  while (true) {
    // ---- the next tile: claimed earlier, metadata fetched now, consumed after the sorts
    nm_tile_meta nxt;
    nm_tile_head h2;
    int64_t t2 = -1;
    bool done;
    if constexpr (kDeep) {
      nxt = nm_tile_fetch_b(a, hn);  // its step A was issued a tile ago
      if (tn >= 0) {
        const long long t = __shfl_sync(0xffffffffu, claim, 0);
        t2 = t < n_tiles ? t : -1;
        if (t2 >= 0 && lane == 0) claim = (long long)atomicAdd(a.tile_cursor, 1) + n_warps;
      }
      h2 = nm_tile_fetch_a(a, t2, lane);
      done = tn < 0;
    } else {
      const long long t = __shfl_sync(0xffffffffu, claim, 0);
      const int64_t next = t < n_tiles ? t : -1;
      done = next < 0;
      if (!done && lane == 0) claim = (long long)atomicAdd(a.tile_cursor, 1) + n_warps;
      nxt = nm_tile_fetch(a, next, lane);
    }

    if (cst.any) {
      // ---- stage the current tile (unless its copies were issued at the end of the last one)
      if (!staged) nm_tile_issue(a, cst, regA, regB, bar, lane);
      const int base0 = cst.base0, base1 = cst.base1;
      nm_tile_wait(cst, bar, parity);

      const int n0 = cur.n0, n1 = cur.n1;
      // Network size of the tile.  Tiles whose longest row is more than half the launch's longest
      // row all use the launch's class: a handful of extra comparators costs far less than
      // keeping several 30-60 KB networks alive in the instruction cache.
      int nsel = nm_lane_class(cst.nmax);
      if (2 * cst.nmax > a.class_n) nsel = a.class_n;
      nm_lane_acc acc;
      acc.dnum = acc.r2 = acc.tie = 0;
      acc.mean0 = acc.var0 = acc.mean1 = acc.var1 = 0.0;
      if (want_t) {  // Welch moments from the raw rows, before the sort overwrites them
        nm_lane_moments(regA, base0, n0, &acc.mean0, &acc.var0);
        nm_lane_moments(regB, base1, n1, &acc.mean1, &acc.var1);
      }
#define NM_CALL(NN)                                                                          \
  if (NN <= NMAX)                                                                            \
    nm_lane_tile<(NN <= NMAX ? NN : NMAX)>(regA, regB, base0, base1, n0, n1, lane, a.one, a.mone)
      NM_DISPATCH_N(nsel, NM_CALL)
#undef NM_CALL

      // plan the next tile (its metadata loads have had both sorts to land) and pull its value
      // slices into L2 while this tile is walked and finished
      const nm_tile_stage nst = nm_tile_plan(nxt, lane, a.region_floats);
      nm_tile_prefetch(a, nst, lane);

      const nm_key* colA = reinterpret_cast<const nm_key*>(regA) + lane;
      const nm_key* colB = reinterpret_cast<const nm_key*>(regB) + lane;
      // KS-only fast walks: four chains where 8 warps/SM leave the walk latency bound (long
      // rows), two chains where 12 warps/SM hide it and the split search is pure overhead
      const int iters = (cst.tmax + 1) >> 1;
      const int it4 = (cst.tmax + 3) >> 2;
      constexpr bool kFour = NMAX > 64;
      const bool fast = kFour ? __all_sync(0xffffffffu, ((n0 + n1) >> 1) >= it4)
                              : __all_sync(0xffffffffu, n0 + n1 >= iters);
      if (want_u)
        nm_merge_walk<true, 32>(colA, colB, n0, n1, iters, &acc);
      else if (fast && kFour)
        acc.dnum = nm_walk_ks_fast4(colA, colB, n0, n1, it4, 32 - __clz(cst.nmax), a.one);
      else if (fast)
        acc.dnum = nm_walk_ks_fast(colA, colB, n0, n1, iters, a.one);
      else
        nm_merge_walk<false, 32>(colA, colB, n0, n1, iters, &acc);

      // ---- the regions are free again: issue the next tile's copies before the fp64 tails
      __syncwarp();
      staged = false;
      if (nst.any) {
        if (lane == 0 && nst.total) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        nm_tile_issue(a, nst, regA, regB, bar, lane);
        staged = true;
      }

      if (cur.ok) {
        // the KS tail runs here (it overlaps other warps' sorts); rank sums and moments are
        // handed to nm_tails_kernel
        double d, pv;
        nm_ks_tail(acc.dnum, n0, n1, &d, &pv);
        a.ks_dnum[cur.r] = acc.dnum;
        if (a.ks_d) a.ks_d[cur.r] = d;
        a.ks_p[cur.r] = pv;
        if (want_u) {
          a.acc_r2[cur.r] = acc.r2;
          a.acc_tie[cur.r] = acc.tie;
        }
        if (want_t) {
          double4* mom = reinterpret_cast<double4*>(a.acc_mom) + cur.r;
          *mom = make_double4(acc.mean0, acc.var0, acc.mean1, acc.var1);
        }
        if (a.flags && !want_u) a.flags[cur.r] = 0;
      }
      cur = nxt;
      cst = nst;
    } else {
      cur = nxt;
      cst = nm_tile_plan(nxt, lane, a.region_floats);
      staged = false;
    }
    if (done) break;
    if constexpr (kDeep) {
      tn = t2;
      hn = h2;
    }
  }
}

// ------------------------------------------------------------------------------------------
// Dense variant: rows == candidates (the coverage filter dropped nothing, no position is deeper
// than the lane tier).  That is the normal shape of a call, and it removes every indirection of
// the general kernel: row r IS candidate r, its slices are [off[r], off[r+1]), a tile's 32 rows
// are always one contiguous span.  Per warp, shared memory additionally holds the 33 + 33 CSR
// offsets of the current and of the next tile (double-buffered, fetched with cp.async one tile
// ahead: no register is tied up, no load is waited for), from which the TMA copies of the next
// tile are issued the moment the merge walk has released the regions.
// The launch ASSUMES the shape (it may be issued before the plan summary has reached the host);
// every warp validates the assumption against the device-side summary first and the kernel
// raises `dense_retry` instead of computing when it does not hold.
// Besides the KS columns the kernel writes row_pos_index / n0 / n1 (there is no scatter pass on
// this path) and norm.isf(p) / ln p for the combine stencil (the fp64 pipe is idle here).
// ------------------------------------------------------------------------------------------
#define NM_DENSE_META_I64 (2 * 2 * 33)  // two buffers x (off0[33] + off1[33])

__device__ __forceinline__ void nm_dense_meta_load(const nm_kargs& a, int64_t tile, long long* buf, int lane) {
  // entries k = 0..32 of both offset arrays, index clamped to n_pos (rows past the end get n = 0)
  const int64_t i0 = tile * 32 + lane;
  const int64_t i = i0 < a.n_pos ? i0 : a.n_pos;
  nm_cp_async8(buf + lane, a.off0 + i);
  nm_cp_async8(buf + 33 + lane, a.off1 + i);
  if (lane == 0) {
    const int64_t j0 = tile * 32 + 32;
    const int64_t j = j0 < a.n_pos ? j0 : a.n_pos;
    nm_cp_async8(buf + 32, a.off0 + j);
    nm_cp_async8(buf + 33 + 32, a.off1 + j);
  }
  nm_cp_async_commit();
}

// With a.retry_mode the kernel's work is what an nm_lane_grid_kernel launch before it left behind:
// the tiles on the retry list, then every tile that launch never claimed.
template <int NMAX>
__global__ void __launch_bounds__(32 * NM_LANE_MAX_WARPS, NMAX <= 64 ? 4 : 2)
nm_lane_dense_kernel(const nm_kargs a, const int want_u, const int want_t) {
  extern __shared__ __align__(128) unsigned char nm_smem[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  nm_summary* S = a.sum;
  {  // the shape this launch was sized for must be the shape the plan found
    const bool shape_ok = nm_dense_shape_ok(*S) && nm_lane_class(S->max_lane_n) == a.class_n;
    if (!shape_ok) {
      if (blockIdx.x == 0 && threadIdx.x == 0) S->dense_retry = 1;
      return;
    }
  }
  const size_t per_warp = 16 + NM_DENSE_META_I64 * sizeof(long long) + 2 * (size_t)a.region_floats * sizeof(float);
  unsigned char* my = nm_smem + (size_t)wib * per_warp;
  uint64_t* bar = reinterpret_cast<uint64_t*>(my);
  long long* meta = reinterpret_cast<long long*>(my + 16);
  float* regA = reinterpret_cast<float*>(my + 16 + NM_DENSE_META_I64 * sizeof(long long));
  float* regB = regA + a.region_floats;
  const int64_t n_rows = a.n_pos;
  const int64_t n_tiles = (n_rows + 31) >> 5;
  const int warps_per_cta = blockDim.x >> 5;
  const int64_t n_warps = (int64_t)gridDim.x * warps_per_cta;

  // work items: tiles 0 .. n_tiles-1, or (retry mode) the retry list followed by the unclaimed tail
  const bool retry = a.retry_mode != 0;
  int64_t n_items = n_tiles, n_listed = 0, tail0 = 0;
  if (retry) {
    n_listed = S->retry_count;
    tail0 = (int64_t)S->grid_n_warps + (int64_t)S->dense_tile_cursor;
    tail0 = tail0 < n_tiles ? tail0 : n_tiles;
    n_items = n_listed + (n_tiles - tail0);
  }
  auto item_tile = [&](int64_t i) -> int64_t {
    if (i >= n_items) return -1;
    if (!retry) return i;
    return i < n_listed ? (int64_t)a.retry_tiles[i] : tail0 + (i - n_listed);
  };

  int64_t tile = item_tile((int64_t)blockIdx.x * warps_per_cta + wib);
  if (tile < 0) return;
  if (lane == 0) nm_mbar_init(bar, 1);
  __syncwarp();
  unsigned parity = 0;
  int mb = 0;
  nm_dense_meta_load(a, tile, meta, lane);
  // lane 0 holds the next tile (claimed one tile ahead: the atomic and the list lookup get a whole tile of
  // work to land)
  long long claim = -1;
  if (lane == 0) claim = item_tile((int64_t)atomicAdd(a.tile_cursor, 1) + n_warps);
  bool staged = false;

  while (true) {
    const int64_t next = __shfl_sync(0xffffffffu, claim, 0);
    const bool done = next < 0;
    if (!done && lane == 0) claim = item_tile((int64_t)atomicAdd(a.tile_cursor, 1) + n_warps);

    // ---- this tile's offsets (landed a tile ago), then the next tile's go on their way
    nm_cp_async_wait_all();
    __syncwarp();
    const long long* m0 = meta + mb * 66;
    const long long* m1 = m0 + 33;
    const long long o0 = m0[lane], o1 = m1[lane];
    const int n0 = (int)(m0[lane + 1] - o0), n1 = (int)(m1[lane + 1] - o1);
    const long long al0 = m0[0] & ~3LL, al1 = m1[0] & ~3LL;
    const unsigned bytes0 = (unsigned)((m0[32] - al0 + 3) & ~3LL) * 4u, bytes1 = (unsigned)((m1[32] - al1 + 3) & ~3LL) * 4u;
    const int64_t r = tile * 32 + lane;
    const bool ok = r < n_rows;
    const int base0 = ok ? (int)(o0 - al0) : 0, base1 = ok ? (int)(o1 - al1) : 0;
    if (!done) nm_dense_meta_load(a, next, meta + (mb ^ 1) * 66, lane);

    if (!staged && lane == 0) {
      nm_mbar_expect_tx(bar, bytes0 + bytes1);
      nm_bulk_g2s(regA, a.vals0 + al0, bytes0, bar);
      nm_bulk_g2s(regB, a.vals1 + al1, bytes1, bar);
    }
    nm_mbar_wait(bar, parity);
    parity ^= 1u;
    __syncwarp();

    const int nmax = __reduce_max_sync(0xffffffffu, n0 > n1 ? n0 : n1);
    const int tmax = __reduce_max_sync(0xffffffffu, n0 + n1);
    int nsel = nm_lane_class(nmax);
    if (2 * nmax > a.class_n) nsel = a.class_n;
    nm_lane_acc acc;
    acc.dnum = acc.r2 = acc.tie = 0;
    acc.mean0 = acc.var0 = acc.mean1 = acc.var1 = 0.0;
    if (want_t) {  // Welch moments from the raw rows, before the sort overwrites them
      nm_lane_moments(regA, base0, n0, &acc.mean0, &acc.var0);
      nm_lane_moments(regB, base1, n1, &acc.mean1, &acc.var1);
    }
#define NM_CALL(NN)                                                                          \
  if (NN <= NMAX)                                                                            \
    nm_lane_tile<(NN <= NMAX ? NN : NMAX)>(regA, regB, base0, base1, n0, n1, lane, a.one, a.mone)
    NM_DISPATCH_N(nsel, NM_CALL)
#undef NM_CALL

    // ---- the next tile's span: known from its offsets (in shared memory by now); pull it into L2
    long long nal0 = 0, nal1 = 0;
    unsigned nbytes0 = 0, nbytes1 = 0;
    if (!done) {
      nm_cp_async_wait_all();
      __syncwarp();
      const long long* q0 = meta + (mb ^ 1) * 66;
      const long long* q1 = q0 + 33;
      nal0 = q0[0] & ~3LL;
      nal1 = q1[0] & ~3LL;
      nbytes0 = (unsigned)((q0[32] - nal0 + 3) & ~3LL) * 4u;
      nbytes1 = (unsigned)((q1[32] - nal1 + 3) & ~3LL) * 4u;
      if (lane == 0) {
        nm_prefetch_l2(a.vals0 + nal0, nbytes0);
        nm_prefetch_l2(a.vals1 + nal1, nbytes1);
      }
    }

    const nm_key* colA = reinterpret_cast<const nm_key*>(regA) + lane;
    const nm_key* colB = reinterpret_cast<const nm_key*>(regB) + lane;
    const int iters = (tmax + 1) >> 1;
    const int it4 = (tmax + 3) >> 2;
    constexpr bool kFour = NMAX > 64;
    const bool fast = kFour ? __all_sync(0xffffffffu, ((n0 + n1) >> 1) >= it4)
                            : __all_sync(0xffffffffu, n0 + n1 >= iters);
    if (want_u)
      nm_merge_walk<true, 32>(colA, colB, n0, n1, iters, &acc);
    else if (fast && kFour)
      acc.dnum = nm_walk_ks_fast4(colA, colB, n0, n1, it4, 32 - __clz(nmax), a.one);
    else if (fast)
      acc.dnum = nm_walk_ks_fast(colA, colB, n0, n1, iters, a.one);
    else
      nm_merge_walk<false, 32>(colA, colB, n0, n1, iters, &acc);

    // ---- the regions are free again: the next tile's copies overlap the fp64 tails
    __syncwarp();
    staged = false;
    if (!done) {
      if (lane == 0) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        nm_mbar_expect_tx(bar, nbytes0 + nbytes1);
        nm_bulk_g2s(regA, a.vals0 + nal0, nbytes0, bar);
        nm_bulk_g2s(regB, a.vals1 + nal1, nbytes1, bar);
      }
      staged = true;
    }

    if (ok) {
      double d, pv;
      nm_ks_tail(acc.dnum, n0, n1, &d, &pv);
      a.w_row_pos_index[r] = (int32_t)r;
      a.w_n0[r] = n0;
      a.w_n1[r] = n1;
      a.ks_dnum[r] = acc.dnum;
      if (a.ks_d) a.ks_d[r] = d;
      a.ks_p[r] = pv;
      if (a.comb_z) a.comb_z[r] = nm_norm_isf(pv);
      if (a.comb_ln) a.comb_ln[r] = log(pv);
      if (want_u) {
        a.acc_r2[r] = acc.r2;
        a.acc_tie[r] = acc.tie;
      }
      if (want_t) {
        double4* mom = reinterpret_cast<double4*>(a.acc_mom) + r;
        *mom = make_double4(acc.mean0, acc.var0, acc.mean1, acc.var1);
      }
      if (a.flags && !want_u) a.flags[r] = 0;
    }
    if (done) break;
    tile = next;
    mb ^= 1;
  }
}

// ------------------------------------------------------------------------------------------
// Grid-key variant of the dense kernel (nm_lane.cuh "Grid keys"): values that are three-place
// decimals -- what the reference stores -- are sorted as 16-bit keys, the two groups of a position
// packed into one register array and sorted by ONE network pass of VIMNMX.U16x2.  The sorted pairs
// go back transposed into region A alone (row k+1 = k-th smallest of group 0 | group 1 << 16;
// row 0 = 0 = -inf, row N+1 = all ones = +inf).
//  * a tile with a value that is not on the grid is put on the retry list instead of being
//    computed (nm_lane_dense_kernel with retry_mode runs afterwards); a warp whose first
//    a.grid_tries tiles all failed raises `grid_giveup`, on which everybody hands back what has
//    been claimed and stops: data that are not three-place decimals cost a few tiles' checks.
//  * NM_GRID_LOCKSTEP: tiles are claimed per CTA (4 consecutive tiles, one per warp) and the four
//    warps -- one per scheduler -- meet at a CTA barrier every round, so that they run through the
//    ~120 KB of straight-line code of a tile together and share its instruction-cache fills; the
//    two warps of a scheduler (different CTAs) still drift against each other.
// A kernel of its own rather than a branch of nm_lane_dense_kernel: with both sorts in one function
// both ran slower (float32 1.70 -> 2.25 ms; ncu stall no_instruction 0.45 -> 1.12 per issue).
// ------------------------------------------------------------------------------------------
#define NM_GRID_ELEM(v, valid, M, PAD, out)                         \
  {                                                                 \
    const float x_ = (valid) ? (v) : 0.0f;                          \
    const unsigned b_ = nm_grid_bits(x_, M, &bad);                  \
    vmax = fmaxf(vmax, fabsf(x_));                                  \
    out = (valid) ? b_ : (PAD);                                     \
  }

template <int N>
__device__ __forceinline__ bool nm_grid_tile(float* regA, float* regB, int base0, int base1, int n0, int n1,
                                             int lane, int one, int mone) {
  nm_p16 w[N];
  const unsigned sh16 = (unsigned)one << 16;  // runtime 65536: keeps the pack an IMAD (FMA pipe)
  const int shift0 = base0 & 3, shift1 = base1 & 3;
  const float4* rawA = reinterpret_cast<const float4*>(regA + (base0 - shift0));
  const float4* rawB = reinterpret_cast<const float4*>(regB + (base1 - shift1));
  nm_grid_flag bad = NM_GRID_FLAG0;
  float vmax = 0.0f;
  const bool full = __all_sync(0xffffffffu, ((shift0 | shift1) == 0) && (n0 == N) && (n1 == N));
  if (__builtin_expect(full, 1)) {
#pragma unroll
    for (int q = 0; q < N / 4; ++q) {
      const float4 a4 = rawA[q], b4 = rawB[q];
      const float av[4] = {a4.x, a4.y, a4.z, a4.w}, bv[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const unsigned ta = nm_grid_bits(av[j], NM_GRID_MA, &bad);
        const unsigned tb = nm_grid_bits(bv[j], NM_GRID_MB, &bad);
        vmax = fmaxf(vmax, fmaxf(fabsf(av[j]), fabsf(bv[j])));
        w[4 * q + j].v = tb * sh16 + ta;
      }
      if ((q & 3) == 3) asm volatile("" ::: "memory");
    }
  } else {
    // window slot e holds row element e - shift; the (up to 3) elements of a long shifted row that
    // lie in slots N..N+2 take the slots 0..shift-1, which are free exactly then
    const float4 wa4 = rawA[N / 4], wb4 = rawB[N / 4];
    const float wa[4] = {wa4.x, wa4.y, wa4.z, wa4.w}, wb[4] = {wb4.x, wb4.y, wb4.z, wb4.w};
#pragma unroll
    for (int q = 0; q < N / 4; ++q) {
      const float4 a4 = rawA[q], b4 = rawB[q];
      const float av[4] = {a4.x, a4.y, a4.z, a4.w}, bv[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int e = 4 * q + j;
        bool va = (unsigned)(e - shift0) < (unsigned)n0, vb = (unsigned)(e - shift1) < (unsigned)n1;
        float xa = av[j], xb = bv[j];
        if (e < 3) {
          const bool wra = (unsigned)(N + e - shift0) < (unsigned)n0, wrb = (unsigned)(N + e - shift1) < (unsigned)n1;
          xa = wra ? wa[e] : xa;
          xb = wrb ? wb[e] : xb;
          va = va || wra;
          vb = vb || wrb;
        }
        unsigned ta, tb;
        NM_GRID_ELEM(xa, va, NM_GRID_MA, NM_GRID_PAD_A, ta)
        NM_GRID_ELEM(xb, vb, NM_GRID_MB, NM_GRID_PAD_B, tb)
        w[e].v = tb * sh16 + ta;
      }
      if ((q & 3) == 3) asm volatile("" ::: "memory");
    }
  }
  if (__any_sync(0xffffffffu, nm_grid_failed(bad) || !(vmax <= NM_GRID_LIM))) return false;
  nm_sorter<N>::run(w, one, mone);
  __syncwarp();  // every lane has consumed its raw rows; region A may now be overwritten
  unsigned* col = reinterpret_cast<unsigned*>(regA) + lane;
  col[0] = NM_GRID_NINF;
#pragma unroll
  for (int k = 0; k < N; ++k) col[(k + 1) << 5] = w[nm_sorter<N>::order(k)].v;
  col[(N + 1) << 5] = NM_GRID_PINF;
  __syncwarp();
  return true;
}

template <int NMAX>
__global__ void __launch_bounds__(32 * NM_LANE_MAX_WARPS, 2)
nm_lane_grid_kernel(const nm_kargs a, const int want_u, const int want_t) {
  extern __shared__ __align__(128) unsigned char nm_smem[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  nm_summary* S = a.sum;
  {  // the shape this launch was sized for must be the shape the plan found
    const bool shape_ok = nm_dense_shape_ok(*S) && nm_lane_class(S->max_lane_n) == a.class_n;
    if (!shape_ok) {
      if (blockIdx.x == 0 && threadIdx.x == 0) S->dense_retry = 1;
      return;
    }
  }
  const size_t per_warp = 16 + NM_DENSE_META_I64 * sizeof(long long) + 2 * (size_t)a.region_floats * sizeof(float);
  unsigned char* my = nm_smem + (size_t)wib * per_warp;
  uint64_t* bar = reinterpret_cast<uint64_t*>(my);
  long long* meta = reinterpret_cast<long long*>(my + 16);
  float* regA = reinterpret_cast<float*>(my + 16 + NM_DENSE_META_I64 * sizeof(long long));
  float* regB = regA + a.region_floats;
  const int64_t n_rows = a.n_pos;
  const int64_t n_tiles = (n_rows + 31) >> 5;
  const int warps_per_cta = blockDim.x >> 5;
  const int64_t n_warps = (int64_t)gridDim.x * warps_per_cta;
  if (blockIdx.x == 0 && threadIdx.x == 0) S->grid_n_warps = (int)n_warps;
  // tiles >= n_tiles are empty rounds of a warp (offsets clamped: n0 = n1 = 0, nothing staged or written)
#ifdef NM_GRID_LOCKSTEP
  __shared__ long long s_claim[2];
  __shared__ int s_gu[2];
  const int claimer = threadIdx.x == 0;
  const int claim_n = warps_per_cta;
  if ((int64_t)blockIdx.x * warps_per_cta >= n_tiles) return;
#else
  const int claimer = lane == 0;
  const int claim_n = 1;
  if ((int64_t)blockIdx.x * warps_per_cta + wib >= n_tiles) return;
#endif
  int64_t tile = (int64_t)blockIdx.x * warps_per_cta + wib;
  if (lane == 0) nm_mbar_init(bar, 1);
  __syncwarp();
  unsigned parity = 0;
  int mb = 0;
  nm_dense_meta_load(a, tile < n_tiles ? tile : n_tiles, meta, lane);
  // the claimer holds the next claim and the give-up flag, both fetched one tile ahead
  long long claim = 0;
  int gu = 0;
  if (claimer) {
    claim = (long long)atomicAdd(a.tile_cursor, claim_n) + n_warps;
    gu = *(volatile int*)&S->grid_giveup;
  }
  bool staged = false;
  int grid_fail = 0, grid_done = 0, round = 0;

  while (true) {
#ifdef NM_GRID_LOCKSTEP
    if (claimer) {
      s_claim[round & 1] = claim;
      s_gu[round & 1] = gu;
    }
    __syncthreads();
    const long long t0 = s_claim[round & 1];
    const bool giveup = s_gu[round & 1] != 0;
    bool done = t0 >= n_tiles;  // CTA-uniform
    int64_t next = t0 + wib;
#else
    const long long t0 = __shfl_sync(0xffffffffu, claim, 0);
    const bool giveup = __shfl_sync(0xffffffffu, gu, 0) != 0 || (grid_done == 0 && grid_fail >= a.grid_tries);
    bool done = t0 >= n_tiles;
    int64_t next = t0;
#endif
    if (giveup && !done) {  // hand the claimed tile back, claim no more
      if (lane == 0 && next < n_tiles) a.retry_tiles[atomicAdd(&S->retry_count, 1)] = (int32_t)next;
      done = true;
    }
    if (!done && claimer) {
      claim = (long long)atomicAdd(a.tile_cursor, claim_n) + n_warps;
      gu = *(volatile int*)&S->grid_giveup;
    }
    ++round;
    const bool have = tile < n_tiles, have_next = !done && next < n_tiles;

    // ---- this tile's offsets (landed a tile ago), then the next tile's go on their way
    nm_cp_async_wait_all();
    __syncwarp();
    const long long* m0 = meta + mb * 66;
    const long long* m1 = m0 + 33;
    const long long o0 = m0[lane], o1 = m1[lane];
    const int n0 = (int)(m0[lane + 1] - o0), n1 = (int)(m1[lane + 1] - o1);
    const long long al0 = m0[0] & ~3LL, al1 = m1[0] & ~3LL;
    const unsigned bytes0 = (unsigned)((m0[32] - al0 + 3) & ~3LL) * 4u, bytes1 = (unsigned)((m1[32] - al1 + 3) & ~3LL) * 4u;
    const int64_t r = tile * 32 + lane;
    const bool ok = r < n_rows;
    const int base0 = ok ? (int)(o0 - al0) : 0, base1 = ok ? (int)(o1 - al1) : 0;
    if (have_next) nm_dense_meta_load(a, next, meta + (mb ^ 1) * 66, lane);

    bool computed = false;
    nm_lane_acc acc;
    acc.dnum = acc.r2 = acc.tie = 0;
    acc.mean0 = acc.var0 = acc.mean1 = acc.var1 = 0.0;
    int nmax = 0, tmax = 0;
    if (have) {
      if (!staged && lane == 0) {
        nm_mbar_expect_tx(bar, bytes0 + bytes1);
        nm_bulk_g2s(regA, a.vals0 + al0, bytes0, bar);
        nm_bulk_g2s(regB, a.vals1 + al1, bytes1, bar);
      }
      nm_mbar_wait(bar, parity);
      parity ^= 1u;
      __syncwarp();

      nmax = __reduce_max_sync(0xffffffffu, n0 > n1 ? n0 : n1);
      tmax = __reduce_max_sync(0xffffffffu, n0 + n1);
      int nsel = nm_lane_class(nmax);
      if (2 * nmax > a.class_n) nsel = a.class_n;
      if (want_t) {  // Welch moments from the raw rows, before the sort overwrites them
        nm_lane_moments(regA, base0, n0, &acc.mean0, &acc.var0);
        nm_lane_moments(regB, base1, n1, &acc.mean1, &acc.var1);
      }
#define NM_CALL(NN)                                                                          \
  if (NN <= NMAX)                                                                            \
    computed = nm_grid_tile<(NN <= NMAX ? NN : NMAX)>(regA, regB, base0, base1, n0, n1, lane, a.one, a.mone)
      NM_DISPATCH_N(nsel, NM_CALL)
#undef NM_CALL
      if (computed) {
        ++grid_done;
      } else {
        ++grid_fail;
        if (lane == 0) {
          a.retry_tiles[atomicAdd(&S->retry_count, 1)] = (int32_t)tile;
          if (grid_done == 0 && grid_fail >= a.grid_tries) S->grid_giveup = 1;
        }
      }
    }

    // ---- the next tile's span: known from its offsets (in shared memory by now); pull it into L2
    long long nal0 = 0, nal1 = 0;
    unsigned nbytes0 = 0, nbytes1 = 0;
    if (have_next) {
      nm_cp_async_wait_all();
      __syncwarp();
      const long long* q0 = meta + (mb ^ 1) * 66;
      const long long* q1 = q0 + 33;
      nal0 = q0[0] & ~3LL;
      nal1 = q1[0] & ~3LL;
      nbytes0 = (unsigned)((q0[32] - nal0 + 3) & ~3LL) * 4u;
      nbytes1 = (unsigned)((q1[32] - nal1 + 3) & ~3LL) * 4u;
      if (lane == 0) {
        // region B has been free since the key pairs were made (the sorted pairs live in region A): the next
        // tile's group-1 slice starts now, under the sort's shadow; group 0's follows when the walk is done
        nm_prefetch_l2(a.vals0 + nal0, nbytes0);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        nm_mbar_expect_tx_only(bar, nbytes1);
        nm_bulk_g2s(regB, a.vals1 + nal1, nbytes1, bar);
      }
    }

    if (computed) {
      const unsigned short* gA = reinterpret_cast<const unsigned short*>(regA) + 2 * lane;
      const unsigned short* gB = gA + 1;
      const int iters = (tmax + 1) >> 1;
      const int it4 = (tmax + 3) >> 2;
      constexpr bool kFour = NMAX > 64;
      const bool fast = kFour ? __all_sync(0xffffffffu, ((n0 + n1) >> 1) >= it4)
                              : __all_sync(0xffffffffu, n0 + n1 >= iters);
      if (want_u)
        nm_merge_walk<true, 64>(gA, gB, n0, n1, iters, &acc);
      else if (fast && kFour)
        acc.dnum = nm_walk_ks_fast4_g(gA, gB, n0, n1, it4, 32 - __clz(nmax), a.one);
      else if (fast)
        acc.dnum = nm_walk_ks_fast_g(gA, gB, n0, n1, iters, a.one);
      else
        nm_merge_walk<false, 64>(gA, gB, n0, n1, iters, &acc);
    }

    // ---- the regions are free again: the next tile's copies overlap the fp64 tails
    __syncwarp();
    staged = false;
    if (have_next) {
      if (lane == 0) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        nm_mbar_expect_tx(bar, nbytes0);
        nm_bulk_g2s(regA, a.vals0 + nal0, nbytes0, bar);
      }
      staged = true;
    }

    if (ok && computed) {
      double d, pv;
      nm_ks_tail(acc.dnum, n0, n1, &d, &pv);
      a.w_row_pos_index[r] = (int32_t)r;
      a.w_n0[r] = n0;
      a.w_n1[r] = n1;
      a.ks_dnum[r] = acc.dnum;
      if (a.ks_d) a.ks_d[r] = d;
      a.ks_p[r] = pv;
      if (a.comb_z) a.comb_z[r] = nm_norm_isf(pv);
      if (a.comb_ln) a.comb_ln[r] = log(pv);
      if (want_u) {
        a.acc_r2[r] = acc.r2;
        a.acc_tie[r] = acc.tie;
      }
      if (want_t) {
        double4* mom = reinterpret_cast<double4*>(a.acc_mom) + r;
        *mom = make_double4(acc.mean0, acc.var0, acc.mean1, acc.var1);
      }
      if (a.flags && !want_u) a.flags[r] = 0;
    }
    if (done) break;
    tile = next;
    mb ^= 1;
  }
  if (lane == 0 && grid_done) atomicAdd(&S->grid_tiles, grid_done);
}

static int nm_lane_warp_smem(int region_floats) { return 16 + 2 * region_floats * (int)sizeof(float); }

template <int NMAX>
static int nm_launch_lane_t(const nm_kargs& ka_in, bool want_u, bool want_t, int sm_count, cudaStream_t st) {
  nm_kargs ka = ka_in;
  int per_warp = nm_lane_warp_smem(ka.region_floats);
  cudaError_t e = cudaFuncSetAttribute(nm_lane_kernel<NMAX>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       NM_LANE_MAX_WARPS * per_warp);
  if (e != cudaSuccess) return (int)e;
  // CTA shape: as many resident warps per SM as shared memory and registers allow
  int best_w = 1, best_blocks = 0, best_warps = 0;
  for (int w = NM_LANE_MAX_WARPS; w >= 1; --w) {
    int blocks = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks, nm_lane_kernel<NMAX>, 32 * w, (size_t)w * per_warp);
    if (e != cudaSuccess) return (int)e;
    if (blocks * w > best_warps) {
      best_warps = blocks * w;
      best_w = w;
      best_blocks = blocks;
    }
  }
  if (best_warps == 0) return (int)cudaErrorInvalidConfiguration;
  // Grow the regions into the shared memory this occupancy leaves unused: a tile whose rows are
  // nearly contiguous (a few filtered / deep / differently binned rows in between) can then be
  // staged as one span with a single bulk copy.  Only when there are such rows: the extra
  // shared memory comes out of the L1 cache (2x50x: 0.90 -> 1.00 ms when grown needlessly).
  if (ka.gaps) {
    int dev = 0, smem_sm = 0, smem_blk = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&smem_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, dev);
    cudaDeviceGetAttribute(&smem_blk, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    int budget = smem_sm / best_blocks - 1024;  // 1 KB per resident CTA is reserved by the system
    if (budget > smem_blk) budget = smem_blk;
    const int grown = ((budget / best_w - 16) / 8) & ~3;  // floats per region
    if (grown > ka.region_floats) {
      const int pw = nm_lane_warp_smem(grown);
      int blocks = 0;
      if (cudaFuncSetAttribute(nm_lane_kernel<NMAX>, cudaFuncAttributeMaxDynamicSharedMemorySize, best_w * pw) == cudaSuccess &&
          cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks, nm_lane_kernel<NMAX>, 32 * best_w, (size_t)best_w * pw) == cudaSuccess &&
          blocks >= best_blocks) {
        ka.region_floats = grown;
        per_warp = pw;
      }
      (void)cudaGetLastError();
    }
  }
  const int64_t tiles = (ka.row_hi - ka.row_lo + 31) / 32;
  int64_t grid = (tiles + best_w - 1) / best_w;
  const int64_t resident = (int64_t)best_blocks * sm_count;
  if (grid > resident) grid = resident;
  nm_lane_kernel<NMAX><<<(unsigned)grid, 32 * best_w, (size_t)best_w * per_warp, st>>>(ka, want_u ? 1 : 0,
                                                                                       want_t ? 1 : 0);
  return (int)cudaGetLastError();
}

// One launch over rows [ka.row_lo, ka.row_hi) of the (optionally class-sorted) row list;
// max_n = longest row among them.  The regions hold the transposed columns (max class + 2 rows
// of 32) and, before that, the raw rows: 32 slices rounded out to 16-byte boundaries.
int nm_launch_lane(const nm_kargs& ka_in, bool want_u, bool want_t, int max_n, int sm_count, cudaStream_t st) {
  nm_kargs ka = ka_in;
  if (ka.row_hi <= ka.row_lo) return (int)cudaSuccess;
  const int ncls = nm_lane_class(max_n);
  // raw rows: consecutive candidates need 32*n + 6 floats; rows copied one by one are rounded
  // out to 16-byte boundaries each (up to 6 more floats per row)
  ka.region_floats = 32 * (ncls + (ka.gaps ? 6 : 2));
  ka.class_n = ncls;
  if (ncls <= 64) return nm_launch_lane_t<64>(ka, want_u, want_t, sm_count, st);
  return nm_launch_lane_t<128>(ka, want_u, want_t, sm_count, st);
}

template <class K>
static int nm_launch_dense_any(K kernel, const nm_kargs& ka, int per_warp, bool want_u, bool want_t, int sm_count,
                               cudaStream_t st) {
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, NM_LANE_MAX_WARPS * per_warp);
  if (e != cudaSuccess) return (int)e;
  int best_w = 1, best_blocks = 0, best_warps = 0;
  for (int w = NM_LANE_MAX_WARPS; w >= 1; --w) {
    int blocks = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks, kernel, 32 * w, (size_t)w * per_warp);
    if (e != cudaSuccess) return (int)e;
    if (blocks * w > best_warps) {
      best_warps = blocks * w;
      best_w = w;
      best_blocks = blocks;
    }
  }
  if (best_warps == 0) return (int)cudaErrorInvalidConfiguration;
  const int64_t tiles = (ka.n_pos + 31) / 32;
  int64_t grid = (tiles + best_w - 1) / best_w;
  const int64_t resident = (int64_t)best_blocks * sm_count;
  if (grid > resident) grid = resident;
  kernel<<<(unsigned)grid, 32 * best_w, (size_t)best_w * per_warp, st>>>(ka, want_u ? 1 : 0, want_t ? 1 : 0);
  return (int)cudaGetLastError();
}

// grid_keys: launch nm_lane_grid_kernel (ka.retry_tiles / ka.grid_tries set; classes > 64 only) instead of the
// float32 kernel
int nm_launch_lane_dense(const nm_kargs& ka_in, bool want_u, bool want_t, int class_n, int sm_count, bool grid_keys,
                         cudaStream_t st) {
  nm_kargs ka = ka_in;
  if (ka.n_pos <= 0) return (int)cudaSuccess;
  ka.region_floats = 32 * (class_n + 2);
  ka.class_n = class_n;
  const int meta = 16 + NM_DENSE_META_I64 * (int)sizeof(long long);
  const int per_warp = meta + 2 * ka.region_floats * (int)sizeof(float);
  if (grid_keys) return nm_launch_dense_any(nm_lane_grid_kernel<128>, ka, per_warp, want_u, want_t, sm_count, st);
  if (class_n <= 64) return nm_launch_dense_any(nm_lane_dense_kernel<64>, ka, per_warp, want_u, want_t, sm_count, st);
  return nm_launch_dense_any(nm_lane_dense_kernel<128>, ka, per_warp, want_u, want_t, sm_count, st);
}
