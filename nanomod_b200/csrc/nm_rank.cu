// nanomod_b200 -- ranking of the result rows on the device.
//
// Reference: mtest2, bin/scripts/myDetect.py:459-461
//   sorted(sign_test, key=lambda m: (m[1][sorted_ind][u], m[1][2][u], m[1][0][u]))   [::-1] for 'st'
// i.e. a STABLE ascending sort on the tuple (combined, KS, U) of p-values (rankUse='pv') or of
// statistics ('st', then the whole list reversed).  Here: least-significant-key-first passes of
// a stable radix sort (CUB DeviceRadixSort -- library code, this is not the hot path) over
// order-preserving 64-bit images of the doubles, carrying the row index.  NaN keys (the U
// p-value of an all-identical position) sort last, as numpy's lexsort puts them; -0.0 == 0.0.
#include <cub/device/device_radix_sort.cuh>
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include "nm_device.cuh"
#include "nm_rank.cuh"

namespace {

__device__ __forceinline__ unsigned long long nm_rank_key(double x) {
  if (x != x) return ~0ull;
  const unsigned long long b = (unsigned long long)__double_as_longlong(x + 0.0);
  return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}

__global__ void nm_rank_iota(int32_t* order, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) order[i] = (int32_t)i;
}

__global__ void nm_rank_gather_keys(const double* __restrict__ col, const int32_t* __restrict__ order,
                                    unsigned long long* __restrict__ keys, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) keys[i] = nm_rank_key(col[order[i]]);
}

__global__ void nm_rank_reverse(const int32_t* __restrict__ in, int32_t* __restrict__ out, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = in[n - 1 - i];
}

inline size_t nm_align256(size_t x) { return (x + 255) & ~(size_t)255; }

}  // namespace

// Layout of the scratch buffer: keys A | keys B | order A | order B | CUB temporary storage.
size_t nm_rank_scratch_bytes(int64_t n) {
  size_t cub_bytes = 0;
  cub::DoubleBuffer<unsigned long long> k(nullptr, nullptr);
  cub::DoubleBuffer<int32_t> v(nullptr, nullptr);
  cub::DeviceRadixSort::SortPairs(nullptr, cub_bytes, k, v, (int)n);
  return 2 * nm_align256(sizeof(unsigned long long) * (size_t)n) + 2 * nm_align256(sizeof(int32_t) * (size_t)n) +
         nm_align256(cub_bytes) + 256;
}

int nm_rank_run(const double* comb, const double* ks, const double* u, int64_t n, int reverse, int32_t* order_out,
                void* scratch, int* launches, cudaStream_t st) {
  if (n <= 0) return (int)cudaSuccess;
  unsigned char* p = (unsigned char*)scratch;
  unsigned long long* kA = (unsigned long long*)p;
  p += nm_align256(sizeof(unsigned long long) * (size_t)n);
  unsigned long long* kB = (unsigned long long*)p;
  p += nm_align256(sizeof(unsigned long long) * (size_t)n);
  int32_t* oA = (int32_t*)p;
  p += nm_align256(sizeof(int32_t) * (size_t)n);
  int32_t* oB = (int32_t*)p;
  p += nm_align256(sizeof(int32_t) * (size_t)n);
  size_t cub_bytes = 0;
  {
    cub::DoubleBuffer<unsigned long long> k(nullptr, nullptr);
    cub::DoubleBuffer<int32_t> v(nullptr, nullptr);
    cub::DeviceRadixSort::SortPairs(nullptr, cub_bytes, k, v, (int)n);
  }
  const unsigned grid = (unsigned)((n + 255) / 256);
  cub::DoubleBuffer<unsigned long long> keys(kA, kB);
  cub::DoubleBuffer<int32_t> ord(oA, oB);
  nm_rank_iota<<<grid, 256, 0, st>>>(ord.Current(), n);
  ++*launches;
  const double* cols[3] = {u, ks, comb};  // least significant key first
  for (int c = 0; c < 3; ++c) {
    if (!cols[c]) continue;
    nm_rank_gather_keys<<<grid, 256, 0, st>>>(cols[c], ord.Current(), keys.Current(), n);
    cudaError_t e = cub::DeviceRadixSort::SortPairs((void*)p, cub_bytes, keys, ord, (int)n, 0, 64, st);
    if (e != cudaSuccess) return (int)e;
    *launches += 2;
  }
  if (reverse) {
    nm_rank_reverse<<<grid, 256, 0, st>>>(ord.Current(), order_out, n);
    ++*launches;
  } else {
    cudaError_t e = cudaMemcpyAsync(order_out, ord.Current(), sizeof(int32_t) * (size_t)n, cudaMemcpyDeviceToDevice, st);
    if (e != cudaSuccess) return (int)e;
  }
  return (int)cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// Group binning of the lane tier's rows.  The lane kernel runs one sorting-network size per
// launch; a call whose rows span several size groups -- <= 64, <= 104, <= 128 reads -- can be
// split into one launch per group.  nm_group_keys counts the groups (the host decides whether
// splitting pays, nm_api.cu) and nm_group_sort_run makes the stable partition of the row indices
// by group (one 2-bit radix pass): rows keep genome order inside a group, so tiles stay nearly
// contiguous when the other groups are sparse.  Deep rows get key 3 and end up last.
// ------------------------------------------------------------------------------------------
namespace {

__global__ void __launch_bounds__(256)
nm_group_keys(const int32_t* __restrict__ row_n0, const int32_t* __restrict__ row_n1, int64_t n, uint8_t* __restrict__ keys,
              int32_t* __restrict__ rows, nm_summary* __restrict__ sum) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int le64 = 0, le104 = 0;
  if (r < n) {
    const int m = row_n0[r] > row_n1[r] ? row_n0[r] : row_n1[r];
    int key = 3;
    if (m <= NM_LANE_TIER_MAX) {
      const int cls = nm_lane_class(m);
      le64 = cls <= 64;
      le104 = cls <= NM_LANE_FINE_MAX;
      key = le64 ? 0 : le104 ? 1 : 2;
    }
    keys[r] = (uint8_t)key;
    rows[r] = (int32_t)r;
  }
  le64 = __reduce_add_sync(0xffffffffu, le64);
  le104 = __reduce_add_sync(0xffffffffu, le104);
  if ((threadIdx.x & 31) == 0) {
    if (le64) atomicAdd(&sum->n_le64, le64);
    if (le104) atomicAdd(&sum->n_le104, le104);
  }
}

}  // namespace

size_t nm_group_sort_scratch_bytes(int64_t n) {
  size_t cub_bytes = 0;
  cub::DoubleBuffer<uint8_t> k(nullptr, nullptr);
  cub::DoubleBuffer<int32_t> v(nullptr, nullptr);
  cub::DeviceRadixSort::SortPairs(nullptr, cub_bytes, k, v, (int)n, 0, 2);
  return 2 * nm_align256((size_t)n) + nm_align256(cub_bytes) + 256;
}

// keys + identity row list + group counts (added to *sum)
int nm_group_keys_run(const int32_t* row_n0, const int32_t* row_n1, int64_t n, int32_t* perm_a, void* scratch,
                      nm_summary* sum, int* launches, cudaStream_t st) {
  nm_group_keys<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(row_n0, row_n1, n, (uint8_t*)scratch, perm_a, sum);
  *launches += 1;
  return (int)cudaGetLastError();
}

// stable partition of perm_a by the keys nm_group_keys_run left in scratch
int nm_group_sort_run(int64_t n, int32_t* perm_a, int32_t* perm_b, const int32_t** perm_out, void* scratch,
                      int* launches, cudaStream_t st) {
  unsigned char* p = (unsigned char*)scratch;
  uint8_t* kA = p;
  p += nm_align256((size_t)n);
  uint8_t* kB = p;
  p += nm_align256((size_t)n);
  size_t cub_bytes = 0;
  {
    cub::DoubleBuffer<uint8_t> k(nullptr, nullptr);
    cub::DoubleBuffer<int32_t> v(nullptr, nullptr);
    cub::DeviceRadixSort::SortPairs(nullptr, cub_bytes, k, v, (int)n, 0, 2);
  }
  cub::DoubleBuffer<uint8_t> keys(kA, kB);
  cub::DoubleBuffer<int32_t> rows(perm_a, perm_b);
  const cudaError_t e = cub::DeviceRadixSort::SortPairs((void*)p, cub_bytes, keys, rows, (int)n, 0, 2, st);
  if (e != cudaSuccess) return (int)e;
  *launches += 2;
  *perm_out = rows.Current();
  return (int)cudaGetLastError();
}


// ------------------------------------------------------------------------------------------
// Head of the ranking without sorting every row.  The called-site rule (mboxplot / plot1,
// myDetect.py:279-297, :153-164) walks the ranked list from the top and stops after topN
// accepted sites, so only a short prefix of `sorted_sign_test` is ever looked at; on several
// GPUs only that prefix has to leave a shard.  Three passes over the primary key column:
//   nm_head_hist     4096-bin histogram of the top 12 bits (sign, exponent) of the key image
//   nm_head_cut      smallest bin b whose cumulative count reaches `want`
//   nm_head_compact  every row in a bin <= b becomes a 32-byte record (row, three key images)
// All rows of bins <= b precede every other row in the full ranking, so the records, sorted
// lexicographically (host side: a few thousand of them), ARE the first rows of the ranking.
// reverse (rankUse='st': the ascending list reversed) uses complemented images and prefers the
// higher row index, which is the reversed stable order.
// ------------------------------------------------------------------------------------------
namespace {

#define NM_HEAD_BINS 4096

__device__ __forceinline__ unsigned long long nm_head_image(const double* col, int64_t r, int reverse) {
  const unsigned long long k = col ? nm_rank_key(col[r]) : 0ull;
  return reverse ? ~k : k;
}

__global__ void __launch_bounds__(256) nm_head_hist(const double* __restrict__ k0, int64_t n, int reverse,
                                                    unsigned* __restrict__ hist) {
  __shared__ unsigned h[NM_HEAD_BINS];
  for (int b = threadIdx.x; b < NM_HEAD_BINS; b += 256) h[b] = 0;
  __syncthreads();
  // p-values crowd into a handful of exponent bins: the lanes of a warp that hit the same bin add once
  // (MATCH.ANY), otherwise the shared-memory atomics of a warp serialise on 3-4 addresses
  const int lane = threadIdx.x & 31;
  for (int64_t r0 = (int64_t)blockIdx.x * 256 + (threadIdx.x & ~31); r0 < n; r0 += (int64_t)gridDim.x * 256) {
    const int64_t r = r0 + lane;
    const unsigned bin = r < n ? (unsigned)(nm_head_image(k0, r, reverse) >> 52) : 0xffffffffu;
    const unsigned peers = __match_any_sync(0xffffffffu, bin);
    if (bin != 0xffffffffu && lane == __ffs(peers) - 1) atomicAdd(&h[bin], (unsigned)__popc(peers));
  }
  __syncthreads();
  for (int b = threadIdx.x; b < NM_HEAD_BINS; b += 256)
    if (h[b]) atomicAdd(&hist[b], h[b]);
}

// largest bin b with cum(b) <= limit (-1 if none), smallest bin b with cum(b) >= want; by one thread
// over the 256 partial sums and then inside one 16-bin chunk
__device__ void nm_head_search(const unsigned* hist, const unsigned* part, unsigned want, unsigned limit, int* cut_want,
                               unsigned* cum_want, int* cut_fit, unsigned* cum_fit) {
  constexpr int per = NM_HEAD_BINS / 256;
  unsigned cum = 0;
  *cut_want = NM_HEAD_BINS - 1;
  *cut_fit = -1;
  *cum_fit = 0;
  bool have_want = false;
  for (int t = 0; t < 256; ++t) {
    const unsigned nxt = cum + part[t];
    const bool w_here = !have_want && nxt >= want;
    const bool f_here = cum <= limit && nxt > limit;
    if (w_here || f_here) {
      unsigned c = cum;
      for (int b = 0; b < per; ++b) {
        c += hist[t * per + b];
        if (!have_want && c >= want) {
          *cut_want = t * per + b;
          *cum_want = c;
          have_want = true;
        }
        if (c <= limit) {
          *cut_fit = t * per + b;
          *cum_fit = c;
        }
      }
    } else if (nxt <= limit) {
      *cut_fit = t * per + per - 1;
      *cum_fit = nxt;
    }
    cum = nxt;
  }
  if (!have_want) *cum_want = cum;
}

// hist[NM_HEAD_BINS] = cut bin, hist[NM_HEAD_BINS + 1] = rows in bins <= cut, [+2] = compaction cursor (0)
__global__ void __launch_bounds__(256) nm_head_cut(unsigned* __restrict__ hist, unsigned want, unsigned cap, int fit_cap,
                                                   nm_head_record* __restrict__ header, long long n,
                                                   const nm_head_peers_dev peers) {
  __shared__ unsigned part[256];
  unsigned s = 0;
  for (int b = 0; b < NM_HEAD_BINS / 256; ++b) s += hist[threadIdx.x * (NM_HEAD_BINS / 256) + b];
  part[threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    int cut_w, cut_f;
    unsigned cum_w, cum_f;
    nm_head_search(hist, part, want, cap, &cut_w, &cum_w, &cut_f, &cum_f);
    int cut = cut_w;
    unsigned cum = cum_w;
    if (fit_cap && cum_w > cap) {  // the caller's buffer is fixed: stop at the last bin that still fits
      cut = cut_f;
      cum = cum_f;
    }
    hist[NM_HEAD_BINS] = (unsigned)cut;  // 0xffffffff: not even the first occupied bin fits
    hist[NM_HEAD_BINS + 1] = cum;
    hist[NM_HEAD_BINS + 2] = 0;
    if (header) {  // entry 0 of the caller's record buffer
      nm_head_record hdr;
      hdr.row = cum <= cap ? cum : 0;
      hdr.seg = hdr.pos = hdr.full_nbhd = 0;
      hdr.pad = (peers.refused && *peers.refused) ? -1 : peers.epoch;
      hdr.key[0] = (unsigned long long)n;
      hdr.key[1] = (cum == (unsigned)n && cum <= cap) ? 1ull : 0ull;
      hdr.key[2] = (unsigned)cut;
      header[0] = hdr;
      for (int p = 0; p < peers.n; ++p) peers.base[p][0] = hdr;
    }
  }
}

// one record of the head: the row, its three key images, its position and plot1's neighbourhood test
__device__ __forceinline__ nm_head_record nm_head_make_record(const double* __restrict__ k1, const double* __restrict__ k2,
                                                              int64_t r, unsigned long long i0, int reverse,
                                                              const nm_head_geo& geo) {
  nm_head_record rec;
  rec.row = (long long)r;
  rec.key[0] = i0;
  rec.key[1] = k1 ? nm_head_image(k1, r, reverse) : 0ull;
  rec.key[2] = k2 ? nm_head_image(k2, r, reverse) : 0ull;
  rec.seg = rec.pos = -1;
  rec.full_nbhd = 0;
  rec.pad = 0;
  if (geo.pos) {
    // the row inside the caller's whole row list (the ranked range may be a slice of it)
    const int64_t g = geo.row_offset + r;
    const int64_t c = geo.row_pos_index ? (int64_t)geo.row_pos_index[g] : g;
    rec.seg = geo.seg[c];
    rec.pos = geo.pos[c];
    // plot1 (myDetect.py:153-164): rows g-nearby .. g+nearby must be one contiguous run; positions
    // increase strictly inside a segment, so it is one iff its ends are 2*nearby positions apart
    const int64_t lo = g - geo.nearby, hi = g + geo.nearby;
    if (lo >= 0 && hi < geo.n_rows_total) {
      const int64_t cl = geo.row_pos_index ? (int64_t)geo.row_pos_index[lo] : lo;
      const int64_t ch = geo.row_pos_index ? (int64_t)geo.row_pos_index[hi] : hi;
      rec.full_nbhd = (geo.seg[cl] == geo.seg[ch] && (long long)geo.pos[ch] - (long long)geo.pos[cl] == 2LL * geo.nearby) ? 1 : 0;
    }
  }
  return rec;
}

__global__ void __launch_bounds__(256)
nm_head_compact(const double* __restrict__ k0, const double* __restrict__ k1, const double* __restrict__ k2, int64_t n,
                int reverse, unsigned* __restrict__ hist, nm_head_record* __restrict__ out, unsigned cap,
                const nm_head_geo geo, const nm_head_peers_dev peers) {
  const unsigned cut = hist[NM_HEAD_BINS];
  if (cut == 0xffffffffu) return;
  for (int64_t r = (int64_t)blockIdx.x * 256 + threadIdx.x; r < n; r += (int64_t)gridDim.x * 256) {
    const unsigned long long i0 = nm_head_image(k0, r, reverse);
    if ((unsigned)(i0 >> 52) <= cut) {
      const unsigned slot = atomicAdd(&hist[NM_HEAD_BINS + 2], 1u);
      if (slot < cap) {
        const nm_head_record rec = nm_head_make_record(k1, k2, r, i0, reverse, geo);
        out[slot] = rec;
        for (int p = 0; p < peers.n; ++p) peers.base[p][1 + slot] = rec;  // peer memory: the exchange itself
      }
    }
  }
}

// histogram, cut and compaction over the candidate list the combine kernel left (nm_rank.cuh), by one block.
// An entry is {row, exponent bin of its key image}; bins are counted relative to the listing threshold
// (rel = thr_bin - bin >= 0: larger rel = smaller p), so the histogram is 2048 shared-memory words, the cumulative
// counts from the smallest p upwards are one block-wide suffix scan, and the cut is found by all threads at once.
#define NM_HEAD_REL_BINS 2048
__global__ void __launch_bounds__(1024)
nm_head_from_cands(const double* __restrict__ k0, const double* __restrict__ k1, const double* __restrict__ k2, long long n,
                   unsigned want, unsigned cap, const nm_head_geo geo, const int2* __restrict__ cands,
                   const int* __restrict__ cursor, int cand_cap, unsigned thr_bin, nm_head_record* __restrict__ records,
                   int* __restrict__ fail, int* __restrict__ cut_out, const nm_head_peers_dev peers) {
  __shared__ unsigned hist[NM_HEAD_REL_BINS];   // then: suffix sums
  __shared__ unsigned warp_tot[32];
  __shared__ int s_rel_w, s_rel_f;               // largest rel with suffix >= want | smallest rel with suffix <= cap
  __shared__ unsigned s_slot;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int listed = *cursor;
  const bool over = listed > cand_cap;
  const int nc = over ? 0 : listed;
  hist[tid] = 0;
  hist[tid + 1024] = 0;
  if (tid == 0) { s_rel_w = -1; s_rel_f = NM_HEAD_REL_BINS; s_slot = 0; }
  __syncthreads();
  constexpr int U = 8;  // entries in flight per thread: the loop is latency-bound (one block)
  for (int i0 = 0; i0 < nc; i0 += U * 1024) {
    unsigned rel[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int i = i0 + u * 1024 + tid;
      rel[u] = i < nc ? thr_bin - (unsigned)__ldg(&cands[i].y) : 0xffffffffu;
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      // lanes of a warp that hit the same bin add once: the candidates crowd into the few bins next to the threshold
      const unsigned peers_m = __match_any_sync(0xffffffffu, rel[u]);
      if (rel[u] < NM_HEAD_REL_BINS && lane == __ffs(peers_m) - 1) atomicAdd(&hist[rel[u]], (unsigned)__popc(peers_m));
    }
  }
  __syncthreads();
  // suffix sums: thread t owns rel = 2047 - 2t and 2046 - 2t (descending rel = ascending p)
  const int r_hi = NM_HEAD_REL_BINS - 1 - 2 * tid, r_lo = r_hi - 1;
  const unsigned c_hi = hist[r_hi], c_lo = hist[r_lo];
  unsigned inc = c_hi + c_lo;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) warp_tot[wid] = inc;
  __syncthreads();
  unsigned base = 0;
  for (int w = 0; w < wid; ++w) base += warp_tot[w];
  const unsigned suf_lo = base + inc, suf_hi = suf_lo - c_lo;  // rows with rel >= r_lo / >= r_hi
  __syncthreads();
  hist[r_hi] = suf_hi;
  hist[r_lo] = suf_lo;
  if (suf_hi >= want) atomicMax(&s_rel_w, r_hi); else if (suf_lo >= want) atomicMax(&s_rel_w, r_lo);
  if (suf_lo <= cap) atomicMin(&s_rel_f, r_lo); else if (suf_hi <= cap) atomicMin(&s_rel_f, r_hi);
  __syncthreads();
  // every bin up to the listing threshold is counted in full, so the cut is the full histogram's cut as soon as
  // the list holds `want` rows
  const bool ok = !over && s_rel_w >= 0;
  int rel_cut = ok ? s_rel_w : 0;
  unsigned cum = ok ? hist[rel_cut] : 0;
  if (ok && cum > cap) {
    // the caller's buffer is fixed: stop at the last bin whose cumulative count still fits -- occupied or not, as the
    // three-pass form names it: the bin just below the first one that does not fit (rel 2047 is always empty)
    rel_cut = s_rel_f;
    cum = hist[rel_cut];
  }
  if (tid == 0) {
    const bool refused = peers.refused && *peers.refused;
    if (!ok) *fail = 1;
    *cut_out = ok ? (int)(thr_bin - (unsigned)s_rel_w) : 0;  // the bin that reaches `want`: the next call lists up to just above it
    nm_head_record hdr;
    hdr.row = (ok && cum <= cap) ? cum : 0;
    hdr.seg = hdr.pos = hdr.full_nbhd = 0;
    hdr.pad = refused ? -1 : ok ? peers.epoch : -2;
    hdr.key[0] = (unsigned long long)n;
    hdr.key[1] = (ok && cum == (unsigned)n && cum <= cap) ? 1ull : 0ull;
    hdr.key[2] = (unsigned long long)(thr_bin - (unsigned)rel_cut);
    records[0] = hdr;
    for (int p = 0; p < peers.n; ++p) peers.base[p][0] = hdr;
  }
  if (!ok) return;
  for (int i0 = 0; i0 < nc; i0 += U * 1024) {
    int2 c[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int i = i0 + u * 1024 + tid;
      c[u] = i < nc ? __ldg(&cands[i]) : make_int2(0, -1);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (c[u].y >= 0 && thr_bin - (unsigned)c[u].y >= (unsigned)rel_cut && (unsigned)c[u].y <= thr_bin) {
        const unsigned slot = atomicAdd(&s_slot, 1u);
        if (slot < cap) {
          const int64_t r = c[u].x;
          const nm_head_record rec = nm_head_make_record(k1, k2, r, nm_head_image(k0, r, 0), 0, geo);
          records[1 + slot] = rec;
          for (int p = 0; p < peers.n; ++p) peers.base[p][1 + slot] = rec;
        }
      }
    }
  }
}

}  // namespace

unsigned nm_head_thr_bin(int64_t n, int64_t want) {
  if (want <= 0 || n < 8 * want) return 0;
  int j = 0;
  while (j < 60 && (n >> (j + 1)) >= 4 * want) ++j;  // largest j with n / 2^j >= 4 * want
  if (j < 1) return 0;
  return 0x800u + 1023u - (unsigned)j - 1u;  // image >> 52 of the positive doubles below 2^-j
}

int nm_head_from_cands_run(const double* comb, const double* ks, const double* u, int64_t n, int64_t want, int64_t cap,
                           const nm_head_geo& geo, const void* cands, const int* cursor, int cand_cap, unsigned thr_bin,
                           nm_head_record* records, int* fail, int* cut_out, int* launches, cudaStream_t st,
                           const nm_head_peers_dev* peers_in) {
  nm_head_peers_dev peers;
  memset(&peers, 0, sizeof(peers));
  if (peers_in) peers = *peers_in;
  const double* cols[3] = {comb, ks, u};
  const double* k[3] = {nullptr, nullptr, nullptr};
  int m = 0;
  for (int c = 0; c < 3; ++c)
    if (cols[c]) k[m++] = cols[c];
  nm_head_from_cands<<<1, 1024, 0, st>>>(k[0], k[1], k[2], (long long)n, (unsigned)(want < n ? want : n), (unsigned)cap, geo,
                                         (const int2*)cands, cursor, cand_cap, thr_bin, records, fail, cut_out, peers);
  *launches += 1;
  return (int)cudaGetLastError();
}

size_t nm_head_scratch_bytes(int64_t cap) {
  return nm_align256(sizeof(unsigned) * (NM_HEAD_BINS + 4)) + sizeof(nm_head_record) * (size_t)cap;
}

int nm_head_run(const double* comb, const double* ks, const double* u, int64_t n, int reverse, int64_t want, int64_t cap,
                const nm_head_geo& geo, void* scratch, nm_head_record* records, int sm_count, int* launches, cudaStream_t st,
                const nm_head_peers_dev* peers_in) {
  nm_head_peers_dev peers;
  memset(&peers, 0, sizeof(peers));
  if (peers_in && records) peers = *peers_in;
  unsigned* hist = (unsigned*)scratch;
  nm_head_record* recs = records ? records + 1
                                 : (nm_head_record*)((unsigned char*)scratch + nm_align256(sizeof(unsigned) * (NM_HEAD_BINS + 4)));
  // primary key = the first present column, as in nm_rank_run (absent columns do not order)
  const double* cols[3] = {comb, ks, u};
  const double* k[3] = {nullptr, nullptr, nullptr};
  int m = 0;
  for (int c = 0; c < 3; ++c)
    if (cols[c]) k[m++] = cols[c];
  cudaError_t e = cudaMemsetAsync(hist, 0, sizeof(unsigned) * (NM_HEAD_BINS + 4), st);
  if (e != cudaSuccess) return (int)e;
  int64_t blocks = (n + 255) / 256;
  if (blocks > 8 * (int64_t)sm_count) blocks = 8 * (int64_t)sm_count;
  nm_head_hist<<<(unsigned)blocks, 256, 0, st>>>(k[0], n, reverse, hist);
  nm_head_cut<<<1, 256, 0, st>>>(hist, (unsigned)(want < n ? want : n), (unsigned)cap, records ? 1 : 0, records, (long long)n, peers);
  nm_head_compact<<<(unsigned)blocks, 256, 0, st>>>(k[0], k[1], k[2], n, reverse, hist, recs, (unsigned)cap, geo, peers);
  *launches += 3;
  return (int)cudaGetLastError();
}
