// nanomod_b200 -- native writer of the per-position table text (host code, no GPU work).
//
// Reference: save_test, bin/scripts/myDetect.py:522-538 -- one line per row of
// moptions['sign_test']:
//   '%s %s %d %s %d %d %.3f %.3E %.3f %.3E %.3f %.3E' % (chrom, strand, pos+1, base, n0, n1,
//                                                        U, pU, t, pt, D, pks)
// followed by ' %.3f %.3E' % (comb_stat, comb_p) when neighborPvalues > 0 and testMethod != 'ks',
// then '\n'.  The reference formats 4.6 M rows one Python '%' at a time and flushes after every
// line; here the rows are formatted by several host threads straight into one buffer.  C printf
// and Python's '%' agree digit for digit on finite doubles (both round correctly); the spellings
// of the specials are made to match Python's ('inf', '-inf', 'nan' / 'INF', '-INF', 'NAN').
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <thread>
#include <vector>

#include "../../include/nanomod_b200.h"

namespace {

inline int put_f(char* p, double v) {  // '%.3f'
  if (v != v) { memcpy(p, "nan", 3); return 3; }
  if (isinf(v)) { const int n = v < 0 ? 4 : 3; memcpy(p, v < 0 ? "-inf" : "inf", n); return n; }
  return snprintf(p, 400, "%.3f", v);
}
inline int put_e(char* p, double v) {  // '%.3E'
  if (v != v) { memcpy(p, "NAN", 3); return 3; }
  if (isinf(v)) { const int n = v < 0 ? 4 : 3; memcpy(p, v < 0 ? "-INF" : "INF", n); return n; }
  return snprintf(p, 64, "%.3E", v);
}

struct fmt_job {
  const nm_text_columns* c;
  int64_t lo, hi;
  char* out;       // this chunk's private area
  int64_t cap;     // its capacity
  int64_t used;    // bytes written, or -1 on overflow
  int64_t per_row; // nm_format_bound: room a row may need
};

void run_job(fmt_job* j) {
  const nm_text_columns& c = *j->c;
  char* p = j->out;
  char* const end = j->out + j->cap;
  for (int64_t r = j->lo; r < j->hi; ++r) {
    const char* chrom = c.seg_chrom[c.seg[r]];
    const char* strand = c.seg_strand[c.seg[r]];
    if ((int64_t)(end - p) < j->per_row) { j->used = -1; return; }
    p += sprintf(p, "%s %s %lld %c %d %d ", chrom, strand, (long long)c.pos[r] + 1, (char)c.base[r], (int)c.n0[r],
                 (int)c.n1[r]);
    const double vals[8] = {c.u_stat ? c.u_stat[r] : 0.0, c.u_p ? c.u_p[r] : 0.0, c.t_stat ? c.t_stat[r] : 0.0,
                            c.t_p ? c.t_p[r] : 0.0, c.ks_d[r], c.ks_p[r], c.comb_stat ? c.comb_stat[r] : 0.0,
                            c.comb_p ? c.comb_p[r] : 0.0};
    const int nv = (c.comb_stat && c.comb_p) ? 8 : 6;
    for (int k = 0; k < nv; ++k) {
      if (k) *p++ = ' ';
      p += (k & 1) ? put_e(p, vals[k]) : put_f(p, vals[k]);
    }
    *p++ = '\n';
  }
  j->used = p - j->out;
}

}  // namespace

// Worst-case bytes per row for the given table (segment names included).
extern "C" int64_t nm_format_bound(const nm_text_columns* c) {
  if (!c) return -1;
  size_t names = 0;
  for (int32_t s = 0; s < c->n_seg; ++s) {
    const size_t n = strlen(c->seg_chrom[s]) + strlen(c->seg_strand[s]);
    names = n > names ? n : names;
  }
  // 4 '%.3f' fields of up to 314 characters (|x| near DBL_MAX), 4 '%.3E' of 10, integers, blanks
  return (int64_t)(names + 64 + 4 * 316 + 4 * 12);
}

extern "C" int64_t nm_format_sign_test(const nm_text_columns* c, int n_threads, char* out, int64_t out_cap) {
  if (!c || !out || c->n_rows < 0) return -1;
  if (c->n_rows == 0) return 0;
  if (!c->seg || !c->pos || !c->base || !c->n0 || !c->n1 || !c->ks_d || !c->ks_p || !c->seg_chrom || !c->seg_strand) return -1;
  const int64_t per_row = nm_format_bound(c);
  if (n_threads < 1) n_threads = 1;
  if ((int64_t)n_threads > c->n_rows) n_threads = (int)c->n_rows;
  // every thread formats its rows into a private scratch area, the pieces are then packed
  std::vector<fmt_job> jobs((size_t)n_threads);
  std::vector<std::vector<char>> scratch((size_t)n_threads);
  std::vector<std::thread> th;
  const int64_t per = (c->n_rows + n_threads - 1) / n_threads;
  for (int t = 0; t < n_threads; ++t) {
    fmt_job& j = jobs[(size_t)t];
    j.c = c;
    j.lo = (int64_t)t * per;
    j.hi = j.lo + per < c->n_rows ? j.lo + per : c->n_rows;
    if (j.hi < j.lo) j.hi = j.lo;
    // typical rows are ~100 bytes; start there and retry a chunk with the true bound if needed
    scratch[(size_t)t].resize((size_t)((j.hi - j.lo) * 160 + per_row));
    j.out = scratch[(size_t)t].data();
    j.cap = (int64_t)scratch[(size_t)t].size();
    j.used = 0;
    j.per_row = per_row;
    th.emplace_back(run_job, &j);
  }
  for (auto& x : th) x.join();
  int64_t total = 0;
  for (int t = 0; t < n_threads; ++t) {
    fmt_job& j = jobs[(size_t)t];
    if (j.used < 0) {  // pathological values (hundreds of digits): redo this chunk with the bound
      scratch[(size_t)t].assign((size_t)((j.hi - j.lo) * per_row + per_row), 0);
      j.out = scratch[(size_t)t].data();
      j.cap = (int64_t)scratch[(size_t)t].size();
      run_job(&j);
      if (j.used < 0) return -1;
    }
    total += j.used;
  }
  if (total > out_cap) return -total;  // caller's buffer too small: -(bytes needed)
  char* p = out;
  for (int t = 0; t < n_threads; ++t) {
    memcpy(p, jobs[(size_t)t].out, (size_t)jobs[(size_t)t].used);
    p += jobs[(size_t)t].used;
  }
  return total;
}
