/*
 * CPU restatement (C, pthreads) of the reference's counting and am_clr loops.
 *
 * TEST / BASELINE INFRASTRUCTURE ONLY: used by tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs.  Never linked into or
 * called from the product package.
 *
 * oracle_count restates record_counter (autometa/common/kmers.py:161-206):
 *   max_length = len(seq) - k (:186); for i in range(max_length) (:193) — the last
 *   k-mer is never counted; a window counts iff kmer or revcomp is a key of the
 *   upper-case ACGT dictionary (:198-204) == all k bytes in {A,C,G,T};
 *   index = ref_kmers[canonical] (:200-204) comes in as `lut` (MSB-first code with
 *   A=0,T=1,C=2,G=3 -> column; built by oracle_np.kmer_columns from init_kmers, :56-89).
 * oracle_am_clr restates autometa_clr (:369-373) for a table with no all-NA column
 *   (the caller drops rows/columns first): y = (x+1)/sum(x); out = log(y / gmean(y)),
 *   gmean = exp(mean(log y)) as scipy.stats.gmean computes it.
 * Parity: pinned by tests/test_oracle.py against oracle_np (itself pinned to the
 * reference's golden outputs in tests/golden/).
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
  const uint8_t *seq;
  const int64_t *offsets;
  int64_t n;
  int k;
  const int32_t *lut;
  int ncols;
  int32_t *counts;
  int64_t lo, hi;
} count_job;

static int8_t code_of[256];
static int code_init = 0;

static void init_codes(void) {
  if (code_init) return;
  memset(code_of, -1, sizeof(code_of));
  code_of['A'] = 0; code_of['T'] = 1; code_of['C'] = 2; code_of['G'] = 3;
  code_init = 1;
}

static void *count_worker(void *arg) {
  count_job *j = (count_job *)arg;
  const int k = j->k;
  const uint32_t mask = (k == 16) ? 0xffffffffu : ((1u << (2 * k)) - 1u);
  for (int64_t c = j->lo; c < j->hi; ++c) {
    const uint8_t *s = j->seq + j->offsets[c];
    const int64_t L = j->offsets[c + 1] - j->offsets[c];
    int32_t *row = j->counts + c * (int64_t)j->ncols;
    memset(row, 0, sizeof(int32_t) * (size_t)j->ncols);
    const int64_t nwin = L - k; /* kmers.py:186 */
    if (nwin <= 0) continue;
    uint32_t code = 0;
    int run = 0; /* number of consecutive valid bases ending here */
    /* window i covers bases i..i+k-1; it is complete when base i+k-1 has been read */
    for (int64_t p = 0; p < nwin + k - 1; ++p) {
      const int8_t v = code_of[s[p]];
      if (v < 0) { run = 0; code = 0; }
      else { code = ((code << 2) | (uint32_t)v) & mask; ++run; }
      if (p >= k - 1 && run >= k) row[j->lut[code]]++;
    }
  }
  return NULL;
}

static void split_by_bases(const int64_t *offsets, int64_t n, int t, int64_t *cut) {
  const int64_t tot = offsets[n] - offsets[0];
  cut[0] = 0;
  int w = 1;
  for (int64_t i = 0; i < n && w < t; ++i)
    while (w < t && offsets[i] - offsets[0] >= tot * w / t) cut[w++] = i;
  while (w <= t) cut[w++] = n;
}

int oracle_count(const uint8_t *seq, const int64_t *offsets, int64_t n, int k, const int32_t *lut,
                 int ncols, int32_t *counts, int nthreads) {
  init_codes();
  if (k < 1 || k > 15 || n < 0) return -1;
  if (nthreads < 1) nthreads = 1;
  if (nthreads > 256) nthreads = 256;
  int64_t cut[258];
  split_by_bases(offsets, n, nthreads, cut);
  pthread_t th[256];
  count_job jobs[256];
  for (int t = 0; t < nthreads; ++t) {
    jobs[t] = (count_job){seq, offsets, n, k, lut, ncols, counts, cut[t], cut[t + 1]};
    if (nthreads == 1) count_worker(&jobs[t]);
    else pthread_create(&th[t], NULL, count_worker, &jobs[t]);
  }
  if (nthreads > 1)
    for (int t = 0; t < nthreads; ++t) pthread_join(th[t], NULL);
  return 0;
}

typedef struct {
  const int32_t *counts;
  int D;
  double *out;
  int64_t lo, hi;
} clr_job;

static void *clr_worker(void *arg) {
  clr_job *j = (clr_job *)arg;
  const int D = j->D;
  for (int64_t r = j->lo; r < j->hi; ++r) {
    const int32_t *x = j->counts + r * (int64_t)D;
    double *o = j->out + r * (int64_t)D;
    double S = 0.0;
    for (int c = 0; c < D; ++c) S += (double)x[c];
    double ml = 0.0;
    for (int c = 0; c < D; ++c) { o[c] = ((double)x[c] + 1.0) / S; ml += log(o[c]); }
    const double g = exp(ml / D);
    for (int c = 0; c < D; ++c) o[c] = log(o[c] / g);
  }
  return NULL;
}

int oracle_am_clr(const int32_t *counts, int64_t n, int D, double *out, int nthreads) {
  if (nthreads < 1) nthreads = 1;
  if (nthreads > 256) nthreads = 256;
  pthread_t th[256];
  clr_job jobs[256];
  for (int t = 0; t < nthreads; ++t) {
    jobs[t] = (clr_job){counts, D, out, n * t / nthreads, n * (t + 1) / nthreads};
    if (nthreads == 1) clr_worker(&jobs[t]);
    else pthread_create(&th[t], NULL, clr_worker, &jobs[t]);
  }
  if (nthreads > 1)
    for (int t = 0; t < nthreads; ++t) pthread_join(th[t], NULL);
  return 0;
}
