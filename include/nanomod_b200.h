/* nanomod_b200.h -- C ABI of the B200-native NanoMod detection stage.
 *
 * One call = the reference's `mfilter_coverage(moptions); mtest2(moptions)` pair
 * (bin/scripts/myDetect.py:639-641; the functions at :301-314 and :416-457) up to, but not
 * including, ranking and text formatting, which stay on the host side of this boundary.
 * The reference has no FFI of its own: it calls scipy in-process.  The entry points below are
 * what a ctypes binding for that seam binds; INTEGRATION.md shows the stub.
 *
 * Plain C types only.  All buffers are caller-owned; the library owns only its scratch
 * memory and (for the *_host entry) its staging buffers, both inside the handle.
 */
#ifndef NANOMOD_B200_H
#define NANOMOD_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NM_VERSION 120 /* 0.2.0: dense path, ranking heads, int16 transport format (nm_pileup grew), no depth limit */

/* status codes (0 = ok).  nm_last_error(h) gives the text of the most recent failure. */
enum {
  NM_OK = 0,
  NM_ERR_BAD_ARG = 1,    /* null pointer, negative size, misaligned vals pointer ...          */
  NM_ERR_BAD_PARAM = 2,  /* option outside the reference's accepted range (NanoMod.py:66,74) */
  NM_ERR_CUDA = 3,       /* a CUDA runtime call failed                                       */
  NM_ERR_OOM = 4,        /* device or pinned-host allocation failed                          */
  NM_ERR_TOO_DEEP = 5,   /* a position to be down-sampled is beyond NM_DS_DEEP_MAX_READS / _COV  */
  NM_ERR_NO_DEVICE = 6   /* no CUDA device / not an sm_100 device                            */
};

/* combination methods: bit mask.  testMethod 'ks' == NM_COMBINE_NONE (myDetect.py:443),
 * 'fisher' (:392-393), 'stouffer' (:395-401).  Both bits may be set (BASELINE cfg 3). */
enum { NM_COMBINE_NONE = 0, NM_COMBINE_FISHER = 1, NM_COMBINE_STOUFFER = 2 };

#define NM_MAX_NB 32          /* largest neighborPvalues the combine kernel accepts */
#define NM_LANE_TIER_MAX 128  /* coverage per group handled by the lane-per-position tier */
#define NM_DEEP_TIER_MAX_POOLED 49152 /* pow2(n0)+pow2(n1) cap of the block-per-position tier */
#define NM_DS_MAX_READS 256    /* reads per group up to which a down-sampled position takes the warp kernel */
#define NM_DS_DEEP_MAX_READS 32768 /* ... and the block kernel beyond (a down-sampled group longer than this: NM_ERR_TOO_DEEP) */
#define NM_DS_DEEP_MAX_COV 1024    /* largest --coverages threshold for positions with more than NM_DS_MAX_READS reads */
#define NM_DS_MAX_TIMES 1024   /* largest `downsampling` */

/* Options that reach the device (subset of `moptions`, read at myDetect.py:301-414).
 * Names follow NanoMod.py:354-359. */
typedef struct nm_params {
  int32_t min_coverage;  /* MinCoverage      (default 5, must be >= 3: NanoMod.py:66)          */
  int32_t nb;            /* neighborPvalues  (default 2, >= 0: NanoMod.py:74)                  */
  double weights_dif;    /* WeightsDif       (default 2.0; values < 1 become 1.0: :77-78)      */
  int32_t combine;       /* NM_COMBINE_* mask (testMethod)                                     */
  int32_t want_u;        /* also compute mannwhitneyu (myDetect.py:331)                        */
  int32_t want_t;        /* also compute Welch ttest_ind (myDetect.py:335)                     */
  int32_t reserved;
  /* Down-sampling branch of getKStest (myDetect.py:339-361), active for positions of a segment
   * with nm_pileup.seg_cov > 0 where a group has more reads than that: */
  int32_t ds_times;      /* downsampling          (default 100; <= NM_DS_MAX_TIMES)            */
  int32_t ds_index;      /* int(downsampling * downsampling_quantile), 0 <= . < ds_times       */
  uint64_t ds_seed;      /* seed of the library's counter-based stream (nm_downsample.cu)      */
} nm_params;

/* CSR pileup over n_pos candidate positions (the reference's
 * moptions[ds]['norm_mean'][(chrom,strand)][pos] -> list, myDetect.py:569-572), positions in
 * the reference's iteration order: sorted (chrom,strand), ascending pos (:421,429).
 * valsG[offG[i] .. offG[i+1]) are the event means of group G (0 = wrkBase1, 1 = wrkBase2) at
 * candidate i.  A position missing from a group simply has an empty slice.
 * Requirements for the *_device entry: vals pointers 16-byte aligned and readable up to the
 * next multiple of 4 floats past offG[n_pos] (nm_padded_len). */
typedef struct nm_pileup {
  const float* vals0;
  const int64_t* off0; /* [n_pos + 1] */
  const float* vals1;
  const int64_t* off1; /* [n_pos + 1] */
  const int32_t* pos;  /* [n_pos] 0-based reference coordinate                     */
  const int32_t* seg;  /* [n_pos] id of the (chrom,strand) the position belongs to */
  int64_t n_pos;
  /* optional: per segment id, the down-sampling coverage (moptions['coverages'][0 if strand is
   * '+' else 1], myDetect.py:339); NULL or all <= 0 = no down-sampling */
  const int32_t* seg_cov; /* [n_seg] */
  int64_t n_seg;
  /* optional 16-bit transport format: when vals0 / vals1 are NULL, the event means are
   * vals*_i16[k] * i16_unit -- exactly the float32 nearest to that product, i.e. what casting the
   * reference's 0.001-grid float64 values (norm_mean = round(x, 3), myRefBaseSignalAnnotation.py:1108)
   * to float32 gives with i16_unit = 0.001.  Half the bytes over PCIe and in the staging copy; the
   * library expands them on the GPU and everything downstream is unchanged.  Same offsets, same
   * alignment / padding rule in elements (16-byte aligned, readable to the next multiple of 8). */
  const int16_t* vals0_i16;
  const int16_t* vals1_i16;
  double i16_unit;
  int64_t i16_total0; /* values per group (= off0[n_pos], off1[n_pos]); needed by the *_device entry, */
  int64_t i16_total1; /* where the offsets cannot be read by the host                                */
} nm_pileup;

/* Per-row outputs (SoA), each with capacity n_pos.  Row r is the r-th candidate that passes
 * the coverage filter in BOTH groups (myDetect.py:301-314 + :428,431), i.e. row order ==
 * order of moptions['sign_test'].  Pointers marked optional may be NULL. */
typedef struct nm_table {
  int32_t* row_pos_index; /* candidate index of the row                                      */
  int32_t* n0;            /* len(group 0)                                                    */
  int32_t* n1;            /* len(group 1)                                                    */
  int32_t* ks_dnum;       /* max_x |c0(x)*n1 - c1(x)*n0| : exact integer KS numerator        */
  double* ks_d;           /* optional: D = ks_dnum / (n0*n1)                                 */
  double* ks_p;           /* kstwobign.sf((en+0.12+0.11/en)*D), clamped to >= DBL_MIN        */
  int64_t* two_u;         /* want_u: 2*min(u1,u2), exact integer                             */
  double* u_stat;         /* optional: U                                                     */
  double* u_p;            /* want_u: one-sided normal-approximation p (scipy 1.2.1 default)  */
  double* t_stat;         /* want_t: Welch t (group0 - group1)                               */
  double* t_p;            /* want_t: two-sided p                                             */
  double* fisher_stat;    /* NM_COMBINE_FISHER: -2 sum ln p over the window                  */
  double* fisher_p;
  double* stouffer_stat;  /* NM_COMBINE_STOUFFER: weighted Z                                 */
  double* stouffer_p;
  uint8_t* flags;         /* optional: bit0 = all pooled values identical (U p is NaN)       */
  double* moments;        /* optional, 4 doubles per row: mean0, var0, mean1, var1 (ddof=1,  */
                          /* numpy two-pass form): the --mstd output (myDetect.py:437-438,    */
                          /* :540-545; std there is ddof=0 = sqrt(var*(n-1)/n))               */
} nm_table;

typedef struct nm_handle nm_handle;

int nm_version(void);
/* Number of floats a vals buffer must have allocated to hold nvals values. */
int64_t nm_padded_len(int64_t nvals);

/* One handle per GPU and per thread of use; calls on one handle must be serialised. */
int nm_create(int device, nm_handle** out);
void nm_destroy(nm_handle* h);
const char* nm_last_error(const nm_handle* h);

/* Device-resident entry: every pointer in `pileup` and `table` is a device pointer on the
 * handle's GPU.  Work is enqueued on `cuda_stream` (a cudaStream_t; NULL = default stream);
 * the call synchronises that stream once internally (to size its grids) and once more before
 * returning, so outputs are complete on return.  *n_rows receives the number of rows. */
int nm_detect_device(nm_handle* h, const nm_pileup* pileup, const nm_params* params,
                     const nm_table* table, int64_t* n_rows, void* cuda_stream);

/* Host entry: every pointer is host memory (pinned memory makes the copies asynchronous).
 * Copies the pileup to the GPU, runs nm_detect_device, copies the n_rows result rows back. */
int nm_detect_host(nm_handle* h, const nm_pileup* pileup, const nm_params* params,
                   const nm_table* table, int64_t* n_rows);

/* Ranking of the rows (mtest2, myDetect.py:459-461): order[k] = index of the k-th row of
 * moptions['sorted_sign_test'], i.e. a stable ascending sort on (key_comb, key_ks, key_u) --
 * the p-value columns for rankUse='pv', the statistic columns with reverse=1 for 'st'.
 * key_comb (testMethod 'ks') and key_u (U not computed) may be NULL.  NaN keys sort last.
 * _device: all pointers are device memory, work is enqueued on cuda_stream and complete on
 * return; _host: host memory. */
int nm_rank_device(nm_handle* h, const double* key_comb, const double* key_ks, const double* key_u,
                   int64_t n_rows, int reverse, int32_t* order, void* cuda_stream);
int nm_rank_host(nm_handle* h, const double* key_comb, const double* key_ks, const double* key_u,
                 int64_t n_rows, int reverse, int32_t* order);

/* The first rows of that ranking without sorting all of them -- what the called-site rule
 * (mboxplot / plot1, myDetect.py:279-297, :153-164) needs, and all that has to leave a GPU when
 * the genome is sharded.  Keys are DEVICE columns as for nm_rank_device, over the n_rows rows to
 * rank; rows_out is HOST memory with room for `cap` entries.  On return rows_out[0 .. *n_head)
 * are the leading rows of the ranking in order, *n_head >= min(want, n_rows): every row whose
 * primary key shares the cut's exponent bin is included, so rows not returned rank strictly
 * after every row returned.  With `geometry` (device arrays) each entry also carries the row's
 * segment / position and whether rows r-nearby .. r+nearby of the whole row list form one
 * contiguous run (plot1's requirement, :156-164); the ranked rows are rows
 * [row_offset, row_offset + n_rows) of that list (a shard's core rows inside core + halo). */
typedef struct nm_head_geometry {
  const int32_t* row_pos_index; /* row -> candidate, or NULL when rows are the candidates */
  const int32_t* pos;           /* per candidate */
  const int32_t* seg;
  int64_t row_offset;
  int64_t n_rows_total;
  int32_t nearby;
  int32_t reserved;
} nm_head_geometry;
typedef struct nm_head_row {
  int64_t row;       /* index into the ranked rows */
  int32_t seg, pos;  /* -1 without geometry */
  int32_t full_nbhd;
  int32_t reserved;
  uint64_t key[3];   /* sort images of the row's (combined, KS, U) keys, 0 where a key is absent: unsigned
                        lexicographic order of (key[0], key[1], key[2]), then row ascending (descending when
                        reverse), IS the ranking -- heads of several shards merge by them */
} nm_head_row;
int nm_rank_head_device(nm_handle* h, const double* key_comb, const double* key_ks, const double* key_u,
                        int64_t n_rows, int reverse, int64_t want, const nm_head_geometry* geometry,
                        nm_head_row* rows_out, int64_t cap, int64_t* n_head, void* cuda_stream);

/* The same selection left on the DEVICE, unsorted, without waiting for anything: records_dev has
 * room for cap + 1 entries; entry 0 is a header (row = entries that follow, key[0] = n_rows,
 * key[1] = 1 when they are ALL the rows, key[2] = exponent bin of the cut), entries 1.. are every row
 * of the bins up to the cut -- the cut is lowered, if need be, to the bins that fit `cap` entries
 * (header row = 0 and key[1] = 0: not even the first occupied bin fits).  This is what a rank of a
 * sharded run hands to the all-gather; sorting the gathered entries by (key, row) merges the heads. */
int nm_rank_head_select_device(nm_handle* h, const double* key_comb, const double* key_ks, const double* key_u,
                               int64_t n_rows, int reverse, int64_t want, const nm_head_geometry* geometry,
                               nm_head_row* records_dev, int64_t cap, void* cuda_stream);

/* Asynchronous form of nm_detect_device for callers that run call after call on the same shape (genome
 * shards, slabs, benchmark steps): up to TWO calls may be in flight on a handle, so the device never idles while
 * the host looks at a result.  nm_detect_device_async returns a ticket (0 or 1) as soon as the work is queued
 * on `cuda_stream` -- without any host wait when the handle's previous call had the dense shape (nothing
 * filtered, nothing deep, one network class): the kernels are launched on that assumption and validate it on
 * the device.  Any other call is run to completion before the function returns (the ticket is still valid).
 * nm_detect_finish(ticket) waits for that call only, re-runs it the ordinary way if the device refused the
 * assumed shape, and reports n_rows and whether an armed head selection (nm_arm_head_select) ran with it.
 * Inputs and outputs of a call must stay untouched until it is finished; two calls in flight must write
 * different tables; tickets are finished in the order they were issued.  nm_last_timings / nm_last_path /
 * nm_last_grid_tiles describe the call finished last.  (No counterpart in the reference, whose mtest2 is one
 * synchronous pass: myDetect.py:364-462.) */
int nm_detect_device_async(nm_handle* h, const nm_pileup* pileup, const nm_params* params, const nm_table* table,
                           void* cuda_stream, int* ticket_out);
int nm_detect_finish(nm_handle* h, int ticket, int64_t* n_rows_out, int* head_fired_out);

/* Arm that selection for the NEXT nm_detect_device call on this handle: the call launches it on its own
 * stream right behind its last kernel and BEFORE its host wait (a sharded step then has no idle gap between the
 * tests and the exchange of the heads), provided the call's rows turn out to be its candidates (nothing filtered)
 * -- the key columns and the geometry are given for that case.  One shot: nm_head_fired() tells whether it ran;
 * the arming is dropped when the call returns either way, and the caller then selects with
 * nm_rank_head_select_device as before.  (Replaces nothing in the reference: see nm_rank_head_device.) */
int nm_arm_head_select(nm_handle* h, const double* key_comb, const double* key_ks, const double* key_u,
                       int64_t n_rows, int reverse, int64_t want, const nm_head_geometry* geometry,
                       nm_head_row* records_dev, int64_t cap);
int nm_head_fired(const nm_handle* h);

/* The exchange of a sharded run's heads, fused into their selection.  nm_head_set_peers names, for the NEXT
 * nm_arm_head_select or nm_rank_head_select_device call on the handle (one shot), up to NM_MAX_PEERS record
 * buffers in peer-visible device memory: the selection kernels then store the header and every record they
 * select into each base[p] as well (plain stores over NVLink into the peers' HBM) -- base[p] is this rank's
 * section (cap + 1 records) of peer p's gathered-heads buffer, so when the kernels of all ranks have finished
 * every rank holds every head, with no collective kernel and no SM set aside for one.  The header's
 * `reserved` word carries `epoch` (the caller's step number: readers tell a section still holding an older
 * step from it), or -1 when the detect call the selection was armed for did not compute (a refused
 * speculative launch, nm_detect_device_async).  Completion is the caller's to establish (stream/device
 * synchronisation on every rank + a barrier) before the buffer is read.
 * nm_peer_alloc / nm_peer_open wrap cudaMalloc + cudaIpcGetMemHandle / cudaIpcOpenMemHandle for buffers shared
 * between the one-process-per-GPU ranks of a node; within one process any device pointer will do.
 * (Replaces nothing in the reference, which is single-process: see nm_rank_head_device.) */
#define NM_MAX_PEERS 16
#define NM_IPC_HANDLE_BYTES 64
typedef struct nm_head_peers {
  int32_t n_peers;
  int32_t epoch;
  nm_head_row* base[NM_MAX_PEERS];
} nm_head_peers;
int nm_head_set_peers(nm_handle* h, const nm_head_peers* peers);
int nm_peer_alloc(nm_handle* h, int64_t bytes, void** dev_ptr_out, unsigned char* ipc_handle_out /* 64 bytes */);
int nm_peer_open(nm_handle* h, const unsigned char* ipc_handle, void** dev_ptr_out);
int nm_peer_close(nm_handle* h, void* dev_ptr);
int nm_peer_free(nm_handle* h, void* dev_ptr);

/* Packs rows [row_lo, row_lo + n) of a device-resident table into fixed 28-byte records
 * { int32 ks_dnum | double ks_p | double comb_stat | double comb_p } (no padding) in `records`
 * (device, 28*n bytes): the unit a rank sends to rank 0 in multi-GPU runs.
 * which_combine: NM_COMBINE_FISHER or NM_COMBINE_STOUFFER selects the combined columns. */
#define NM_RECORD_BYTES 28
int nm_pack_records_device(nm_handle* h, const nm_table* table, int64_t row_lo, int64_t n, int which_combine,
                           void* records, void* cuda_stream);

/* Text of the per-position table (save_test, myDetect.py:522-538), formatted on the host by
 * n_threads threads: '%s %s %d %s %d %d %.3f %.3E %.3f %.3E %.3f %.3E[ %.3f %.3E]\n' per row with
 * pos + 1, exactly as the reference's Python '%' prints them.  All pointers are host memory;
 * u_*, t_* may be NULL (printed as 0); comb_* both NULL = no combined columns (testMethod 'ks' or
 * neighborPvalues 0).  Returns the bytes written to `out`, -(bytes needed) if out_cap is too
 * small, -1 on a bad argument.  nm_format_bound gives a safe per-row capacity. */
typedef struct nm_text_columns {
  const char* const* seg_chrom;  /* [n_seg] chromosome name of each segment id */
  const char* const* seg_strand; /* [n_seg] '+' or '-'                         */
  int32_t n_seg;
  int32_t reserved;
  int64_t n_rows;
  const int32_t* seg;  /* per row: segment id, 0-based position, base character, coverages */
  const int32_t* pos;
  const uint8_t* base;
  const int32_t* n0;
  const int32_t* n1;
  const double* u_stat;
  const double* u_p;
  const double* t_stat;
  const double* t_p;
  const double* ks_d;
  const double* ks_p;
  const double* comb_stat;
  const double* comb_p;
} nm_text_columns;
int64_t nm_format_bound(const nm_text_columns* columns);
int64_t nm_format_sign_test(const nm_text_columns* columns, int n_threads, char* out, int64_t out_cap);

/* Number of kernels this handle has launched so far (bench.py's gpu_launches). */
int64_t nm_launch_count(const nm_handle* h);

/* Restrict the persistent lane-tier kernel to n_sms SMs (0 = all).  Multi-GPU callers leave a
 * few SMs free so that NCCL's copy kernels can run while the next shard piece is computed. */
int nm_set_sm_limit(nm_handle* h, int n_sms);
int nm_sm_count(const nm_handle* h);

/* Device time (ms, CUDA events on the call's stream) of the most recent nm_detect_* call:
 * ms4[0] plan kernels, [1] lane-tier kernel, [2] deep-tier kernel + the U/t tails
 * kernel, [3] combine kernel. */
int nm_last_timings(const nm_handle* h, double* ms4);

/* Which code path the most recent nm_detect_* call took: 0 general (plan + compaction, lane /
 * deep tiers, combine), 1 dense (rows == candidates: no compaction pass, no indirection),
 * 2 dense launched on the previous call's shape without waiting for the plan summary,
 * 3 such a launch refused by the device-side check and the call re-run dense with the right network
 * class, 4 refused and re-run on the general path. */
int nm_last_path(const nm_handle* h);

/* How many 32-position tiles of the most recent nm_detect_* call the lane tier sorted through its
 * 16-bit grid-key path: values that are float32 images of decimals with three places (the
 * reference's norm_mean = round(x, 3), myRefBaseSignalAnnotation.py:1108; |x| <= 32.766) are
 * sorted as packed 16-bit keys, both groups in one network pass.  Every value is checked on the
 * device; a tile with any other value takes the float32 path.  Results are identical either way. */
int64_t nm_last_grid_tiles(const nm_handle* h);

/* Self-test of the grid-key check on the device: evaluates it on all 2^32 float32 patterns.
 * *passes = patterns accepted (65 534: the 65 533 grid points |k| <= 32766 and -0.0), *violations =
 * accepted patterns that are not fl32(k / 1000) or whose key halves are not k + 32768 (must be 0). */
int nm_grid_selftest(nm_handle* h, int64_t* violations, int64_t* passes);

#ifdef __cplusplus
}
#endif
#endif /* NANOMOD_B200_H */
