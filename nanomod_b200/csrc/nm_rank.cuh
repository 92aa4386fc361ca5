// nanomod_b200 -- internal interface of the ranking unit (nm_rank.cu), used by nm_api.cu.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

// bytes of device scratch nm_rank_run needs for n rows
size_t nm_rank_scratch_bytes(int64_t n);
// Writes to order_out (device) the row indices in ranked order; comb / u may be NULL (key
// absent).  All pointers are device pointers.  Returns a cudaError_t as int; adds the number of
// kernels launched to *launches.
int nm_rank_run(const double* comb, const double* ks, const double* u, int64_t n, int reverse, int32_t* order_out,
                void* scratch, int* launches, cudaStream_t st);

// Group binning of the lane tier (nm_rank.cu).  nm_group_keys_run writes one group key per row
// into scratch and the identity row list into perm_a, and adds the group counts n_le64 / n_le104
// to *sum (device); nm_group_sort_run then partitions the rows by group (stable; deep rows last)
// into one of perm_a / perm_b and returns which in *perm_out.
struct nm_summary;
size_t nm_group_sort_scratch_bytes(int64_t n);
int nm_group_keys_run(const int32_t* row_n0, const int32_t* row_n1, int64_t n, int32_t* perm_a, void* scratch,
                      nm_summary* sum, int* launches, cudaStream_t st);
int nm_group_sort_run(int64_t n, int32_t* perm_a, int32_t* perm_b, const int32_t** perm_out, void* scratch,
                      int* launches, cudaStream_t st);

// Head of the ranking (nm_rank.cu): records of every row whose primary key falls into the
// lowest exponent bins that together hold >= want rows.  scratch (device): a 4096 + 4 word
// histogram block followed by `cap` records; word [4096 + 1] = rows selected, [4096 + 2] = rows
// the compaction met (== selected; more than cap => the records are truncated).
struct nm_head_record {  // same layout as nm_head_row (include/nanomod_b200.h)
  long long row;
  int seg, pos;               // the row's segment id / position (-1 without geometry)
  int full_nbhd;              // plot1's neighbourhood test passed
  int pad;
  unsigned long long key[3];  // sort images of (combined, KS, U); 0 where absent
};
// optional geometry for the records: the ranked rows are rows [row_offset, row_offset + n) of a row
// list of n_rows_total rows (row -> candidate through row_pos_index, identity when NULL)
struct nm_head_geo {
  const int32_t* row_pos_index;
  const int32_t* pos;
  const int32_t* seg;
  int64_t row_offset;
  int64_t n_rows_total;
  int nearby;
};
// optional peers of the selection (a sharded run): the header and every record are ALSO stored, by the
// same kernels, into each peer's gathered-heads buffer (peer-mapped device memory over NVLink; base[p] is
// this rank's section there) -- the all-gather of the heads is fused into their selection.  The header's pad
// word carries `epoch` (which step's heads these are), or -1 when *refused is non-zero (the detect call the
// selection was armed for did not compute: nm_summary::dense_retry).
#define NM_MAX_PEERS 16
struct nm_head_peers_dev {
  int n, epoch;
  nm_head_record* base[NM_MAX_PEERS];
  const int* refused;
};
size_t nm_head_scratch_bytes(int64_t cap);
// records == NULL: the records follow the histogram block inside scratch.  Otherwise records[0] becomes
// a header (row = rows selected, key[0] = n, key[1] = 1 when the head holds every row, key[2] = cut bin)
// and records[1 .. cap] the selection; the cut is lowered to the bins that fit `cap` records.
int nm_head_run(const double* comb, const double* ks, const double* u, int64_t n, int reverse, int64_t want, int64_t cap,
                const nm_head_geo& geo, void* scratch, nm_head_record* records, int sm_count, int* launches, cudaStream_t st,
                const nm_head_peers_dev* peers = nullptr);

// Head selection from a candidate list.  When the selection is armed for a detect call whose primary ranking
// key is the combined p-value, the combine kernel itself lists every core row whose key image falls into an
// exponent bin <= thr_bin (p < 2^-j, j chosen so that a null table gives 4..8 x `want` such rows: a fraction of a
// percent of the rows, one atomic each); one block then does histogram, cut and compaction over that list
// instead of three passes over every row.  Same header, same records (in another order).  *fail is raised --
// and nothing usable written -- when the list overflowed or holds fewer than `want` rows.  *cut_out receives the
// bin that reached `want`: a table with many significant rows has far more rows below the null threshold than the
// head needs, so the next call on the handle lists only up to two bins above the last cut (nm_api.cu).
#if defined(__CUDACC__)
__device__ __forceinline__ unsigned long long nm_key_image(double x) {  // == nm_rank_key (nm_rank.cu)
  if (x != x) return ~0ull;
  const unsigned long long b = (unsigned long long)__double_as_longlong(x + 0.0);
  return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
#endif
unsigned nm_head_thr_bin(int64_t n, int64_t want);  // 0: the table is too short for the candidate path
int nm_head_from_cands_run(const double* comb, const double* ks, const double* u, int64_t n, int64_t want, int64_t cap,
                           const nm_head_geo& geo, const void* cands /* int2 {row, bin} */, const int* cursor, int cand_cap,
                           unsigned thr_bin, nm_head_record* records, int* fail, int* cut_out, int* launches, cudaStream_t st,
                           const nm_head_peers_dev* peers);
