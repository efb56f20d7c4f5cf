// Host-side pieces of libautometa_b200: error state, k-mer column table, FASTA
// indexing, 2-bit packing, count work list.  No reference code is used; the
// behaviour restated here is cited per function in include/autometa_b200.h.
#include <algorithm>
#include <thread>
#include <vector>

#include "am_common.cuh"

namespace am {
thread_local char g_err[512] = "";
std::atomic<int64_t> g_launches{0};
std::atomic<int> g_reserved_sms{0};
}  // namespace am

namespace {
template <typename F>
void parallel_rows(int64_t n, int n_threads, F f) {
  n_threads = std::max(1, std::min(n_threads > 0 ? n_threads : (int)std::thread::hardware_concurrency(), 64));
  const int64_t block = 256;
  const int64_t n_blocks = (n + block - 1) / block;
  n_threads = (int)std::max<int64_t>(1, std::min<int64_t>(n_threads, n_blocks));
  std::atomic<int64_t> next{0};
  auto work = [&]() {
    for (;;) {
      const int64_t b = next.fetch_add(1);
      if (b >= n_blocks) break;
      f(b * block, std::min(n, (b + 1) * block));
    }
  };
  std::vector<std::thread> th;
  for (int t = 1; t < n_threads; ++t) th.emplace_back(work);
  work();
  for (auto &t : th) t.join();
}
}  // namespace

extern "C" {

int am_version(void) { return 100; }
const char *am_last_error(void) { return am::g_err; }
int64_t am_launch_count(void) { return am::g_launches.load(); }
int am_set_reserved_sms(int n) {
  if (n < 0 || n > 64) return am::fail(AM_EINVAL, "am_set_reserved_sms: n must be in 0..64");
  am::g_reserved_sms.store(n);
  return AM_OK;
}
int am_get_reserved_sms(void) { return am::g_reserved_sms.load(); }

// ---- count matrix <-> the per-column arrays of the frame `count` returns (kmers.py:205,325-326) --------------
// The reference's frame holds one Python object per cell; the drop-in's holds one (values, mask) array pair per
// k-mer column (pandas nullable integers).  Going between the row-major count matrix of the device and those
// 512 column arrays is a transpose of 0.4 GB at C2; numpy's strided copies take 0.8 s per direction, these
// blocked, threaded passes 0.1 s.

int am_counts_to_columns(const int32_t *h_counts, int64_t n_rows, int64_t n_cols, const uint8_t *h_countable,
                         int32_t *h_values_t, uint8_t *h_mask_t, int n_threads) {
  if (n_rows < 0 || n_cols < 0) return am::fail(AM_EINVAL, "am_counts_to_columns: bad shape");
  if (n_rows == 0 || n_cols == 0) return AM_OK;
  if (!h_counts || !h_values_t || !h_mask_t) return am::fail(AM_EINVAL, "am_counts_to_columns: null pointer");
  parallel_rows(n_rows, n_threads, [&](int64_t r0, int64_t r1) {
    for (int64_t c0 = 0; c0 < n_cols; c0 += 64) {
      const int64_t c1 = std::min(n_cols, c0 + 64);
      for (int64_t r = r0; r < r1; ++r) {
        const bool dead = h_countable && !h_countable[r];
        const int32_t *src = h_counts + r * n_cols;
        for (int64_t c = c0; c < c1; ++c) {
          const int32_t v = src[c];
          h_values_t[c * n_rows + r] = v;
          h_mask_t[c * n_rows + r] = (uint8_t)(dead || v == 0);
        }
      }
    }
  });
  return AM_OK;
}

int am_columns_to_counts(const void *const *h_col_values, const uint8_t *const *h_col_masks, int elem_size,
                         int64_t n_rows, int64_t n_cols, int32_t *h_counts, uint8_t *h_isna, int64_t *h_bad,
                         int n_threads) {
  if (n_rows < 0 || n_cols < 0 || (elem_size != 4 && elem_size != 8))
    return am::fail(AM_EINVAL, "am_columns_to_counts: bad shape or element size");
  if (h_bad) *h_bad = 0;
  if (n_rows == 0 || n_cols == 0) return AM_OK;
  if (!h_col_values || !h_col_masks || !h_counts || !h_isna) return am::fail(AM_EINVAL, "am_columns_to_counts: null pointer");
  std::atomic<int64_t> bad{0};
  parallel_rows(n_rows, n_threads, [&](int64_t r0, int64_t r1) {
    int64_t nb = 0;
    for (int64_t c0 = 0; c0 < n_cols; c0 += 64) {
      const int64_t c1 = std::min(n_cols, c0 + 64);
      for (int64_t c = c0; c < c1; ++c) {
        const uint8_t *m = h_col_masks[c];
        const int32_t *v4 = (const int32_t *)h_col_values[c];
        const int64_t *v8 = (const int64_t *)h_col_values[c];
        for (int64_t r = r0; r < r1; ++r) {
          const bool na = m[r] != 0;
          const int64_t v = na ? 0 : (elem_size == 4 ? (int64_t)v4[r] : v8[r]);
          if (v < 0 || v > 0x7fffffffLL) ++nb;
          h_counts[r * n_cols + c] = (int32_t)v;
          h_isna[r * n_cols + c] = (uint8_t)na;
        }
      }
    }
    if (nb) bad.fetch_add(nb);
  });
  if (h_bad) *h_bad = bad.load();
  return AM_OK;
}

// ---- k-mer column table (kmers.py:56-89) ------------------------------------
static int build_columns(int k, std::vector<uint32_t> &canon_sorted, std::vector<int32_t> &lut) {
  if (k < 1 || k > 7) return am::fail(AM_EINVAL, "k must be in 1..7");
  const uint32_t n = 1u << (2 * k);
  std::vector<uint32_t> canon(n);
  for (uint32_t c = 0; c < n; ++c) canon[c] = std::min(c, am::revcomp_code(c, k));
  canon_sorted = canon;
  std::sort(canon_sorted.begin(), canon_sorted.end());
  canon_sorted.erase(std::unique(canon_sorted.begin(), canon_sorted.end()), canon_sorted.end());
  lut.resize(n);
  for (uint32_t c = 0; c < n; ++c)
    lut[c] = (int32_t)(std::lower_bound(canon_sorted.begin(), canon_sorted.end(), canon[c]) -
                       canon_sorted.begin());
  return AM_OK;
}

int am_kmer_num_columns(int k) {
  std::vector<uint32_t> cs;
  std::vector<int32_t> lut;
  int rc = build_columns(k, cs, lut);
  if (rc) return rc;
  return (int)cs.size();
}

int am_kmer_column_lut(int k, int32_t *h_lut) {
  std::vector<uint32_t> cs;
  std::vector<int32_t> lut;
  int rc = build_columns(k, cs, lut);
  if (rc) return rc;
  memcpy(h_lut, lut.data(), lut.size() * sizeof(int32_t));
  return AM_OK;
}

int am_kmer_column_names(int k, char *h_names) {
  std::vector<uint32_t> cs;
  std::vector<int32_t> lut;
  int rc = build_columns(k, cs, lut);
  if (rc) return rc;
  static const char L[4] = {'A', 'T', 'C', 'G'};
  for (size_t i = 0; i < cs.size(); ++i) {
    char *s = h_names + i * (k + 1);
    for (int j = 0; j < k; ++j) s[j] = L[(cs[i] >> (2 * (k - 1 - j))) & 3u];
    s[k] = 0;
  }
  return AM_OK;
}

// ---- FASTA index (Bio.SeqIO.parse semantics used at kmers.py:143-148) ---------
// FASTA index of a byte range made of whole lines.  `count_only`: tally headers and sequence bytes (before the
// range's first header / after it); else write ids, offsets and sequence bytes from the given bases.
struct FastaRange {
  int64_t headers = 0;       // header lines in the range
  int64_t bytes_before = 0;  // sequence bytes on lines before the range's first header
  int64_t bytes_after = 0;   // sequence bytes after it
};

static inline bool is_blank_tail(uint8_t c) { return c == ' ' || c == '\t' || c == '\r' || c == '\v' || c == '\f'; }

static void fasta_range(const uint8_t *buf, int64_t p, int64_t end, bool count_only, FastaRange *st, bool in_record,
                        int64_t rec, int64_t out, uint8_t *seq_out, int64_t *offsets, int64_t *id_begin,
                        int64_t *id_end, int64_t cap) {
  while (p < end) {
    const uint8_t *nl = (const uint8_t *)memchr(buf + p, '\n', (size_t)(end - p));
    const int64_t e = nl ? (int64_t)(nl - buf) : end;  // line = [p, e)
    if (e > p && buf[p] == '>') {
      if (count_only) {
        ++st->headers;
      } else {
        if (rec < cap) {
          int64_t b = p + 1;
          while (b < e && (buf[b] == ' ' || buf[b] == '\t')) ++b;  // str.split() skips leading blanks
          int64_t t = b;
          while (t < e && buf[t] != ' ' && buf[t] != '\t' && buf[t] != '\r' && buf[t] != '\v' && buf[t] != '\f') ++t;
          id_begin[rec] = b;
          id_end[rec] = t;
          offsets[rec] = out;
        }
        ++rec;
      }
      in_record = true;
    } else {
      // Biopython: line.rstrip() per line, then remove ' ' and '\r' from the joined string
      int64_t e2 = e;
      while (e2 > p && is_blank_tail(buf[e2 - 1])) --e2;
      if (count_only) {
        int64_t nb = 0;
        for (int64_t q = p; q < e2; ++q) nb += (buf[q] != ' ' && buf[q] != '\r');
        if (st->headers) st->bytes_after += nb;
        else st->bytes_before += nb;
      } else if (in_record && rec <= cap) {
        for (int64_t q = p; q < e2; ++q) {
          const uint8_t c = buf[q];
          if (c == ' ' || c == '\r') continue;
          seq_out[out++] = c;
        }
      }
    }
    p = e + 1;
  }
}

// Threaded: the file is cut into ranges at line starts; pass 1 counts headers and sequence bytes per range, a
// prefix sum gives every range its first record number and output position, pass 2 writes.  (One thread parsed
// the 1 GB of C2 in 1.5 s, more than the GPU spends on everything else in `count`.)
int am_fasta_index_mt(const uint8_t *buf, int64_t len, uint8_t *seq_out, int64_t *offsets, int64_t *id_begin,
                      int64_t *id_end, int64_t cap, int64_t *n_records, int n_threads) {
  if (!buf || len < 0 || !n_records) return am::fail(AM_EINVAL, "am_fasta_index: bad args");
  if (n_threads <= 0) n_threads = (int)std::thread::hardware_concurrency();
  n_threads = std::max(1, std::min(n_threads, 64));
  const int64_t min_range = 1 << 20;
  int n_ranges = (int)std::max<int64_t>(1, std::min<int64_t>(n_threads, len / min_range));
  std::vector<int64_t> cut((size_t)n_ranges + 1, len);
  cut[0] = 0;
  for (int r = 1; r < n_ranges; ++r) {
    int64_t p = std::max(cut[(size_t)r - 1], len / n_ranges * r);
    const uint8_t *nl = p < len ? (const uint8_t *)memchr(buf + p, '\n', (size_t)(len - p)) : nullptr;
    cut[(size_t)r] = nl ? (int64_t)(nl - buf) + 1 : len;  // ranges start at line starts
  }
  std::vector<FastaRange> st((size_t)n_ranges);
  auto run = [&](auto fn) {
    std::vector<std::thread> th;
    for (int r = 1; r < n_ranges; ++r) th.emplace_back(fn, r);
    fn(0);
    for (auto &t : th) t.join();
  };
  run([&](int r) {
    fasta_range(buf, cut[(size_t)r], cut[(size_t)r + 1], true, &st[(size_t)r], false, 0, 0, nullptr, nullptr, nullptr,
                nullptr, 0);
  });
  std::vector<int64_t> rec0((size_t)n_ranges), out0((size_t)n_ranges);
  int64_t n = 0, out = 0;
  for (int r = 0; r < n_ranges; ++r) {
    rec0[(size_t)r] = n;
    out0[(size_t)r] = out;
    if (n > 0) out += st[(size_t)r].bytes_before;  // lines before the file's first header are not sequence
    out += st[(size_t)r].bytes_after;
    n += st[(size_t)r].headers;
  }
  *n_records = n;
  if (!seq_out && !offsets && !id_begin && !id_end) return AM_OK;  // count only: the caller sizes its arrays
  if (n > cap) return am::fail(AM_ECAPACITY, "am_fasta_index: record capacity too small");
  if (!seq_out || !offsets || !id_begin || !id_end) return am::fail(AM_EINVAL, "am_fasta_index: bad args");
  run([&](int r) {
    fasta_range(buf, cut[(size_t)r], cut[(size_t)r + 1], false, nullptr, rec0[(size_t)r] > 0, rec0[(size_t)r],
                out0[(size_t)r], seq_out, offsets, id_begin, id_end, cap);
  });
  offsets[n] = out;
  return AM_OK;
}

int am_fasta_index(const uint8_t *buf, int64_t len, uint8_t *seq_out, int64_t *offsets,
                   int64_t *id_begin, int64_t *id_end, int64_t cap, int64_t *n_records) {
  return am_fasta_index_mt(buf, len, seq_out, offsets, id_begin, id_end, cap, n_records, 0);
}

// ---- packing -----------------------------------------------------------------
static inline int64_t round_up(int64_t x, int64_t m) { return (x + m - 1) / m * m; }

int am_pack_plan(const int64_t *offsets, int64_t n, int64_t *starts, int32_t *lengths,
                 int64_t *total) {
  if (!offsets || n < 0 || !starts || !lengths || !total) return am::fail(AM_EINVAL, "am_pack_plan: bad args");
  int64_t pos = 0;
  for (int64_t i = 0; i < n; ++i) {
    int64_t L = offsets[i + 1] - offsets[i];
    if (L < 0 || L > 0x7fffffff) return am::fail(AM_EINVAL, "am_pack_plan: contig length out of range");
    starts[i] = pos;
    lengths[i] = (int32_t)L;
    pos += round_up(L, AM_GROUP_BASES);
  }
  // trailing pad: kernels may read one 16-byte vector plus one word past a contig's end
  *total = pos + 2 * AM_GROUP_BASES;
  return AM_OK;
}

int64_t am_pack_code_words(int64_t total_packed_bases) { return total_packed_bases / 16; }
int64_t am_pack_bitmap_words(int64_t total_packed_bases) {
  return (total_packed_bases / AM_GROUP_BASES + 31) / 32 + 1;
}

static uint8_t g_b2c[256];
static bool g_b2c_init = false;
static void init_b2c() {
  if (g_b2c_init) return;
  for (int i = 0; i < 256; ++i) g_b2c[i] = 255;
  g_b2c['A'] = 0; g_b2c['T'] = 1; g_b2c['C'] = 2; g_b2c['G'] = 3;
  g_b2c_init = true;
}

int am_pack_host(const uint8_t *seq, const int64_t *offsets, int64_t n, const int64_t *starts,
                 int64_t total, uint32_t *codes, uint32_t *bitmap, uint32_t *rank,
                 uint64_t *exc_mask, int64_t exc_capacity, int64_t *n_exc, int n_threads) {
  if (!seq || !offsets || n < 0 || (n > 0 && !starts) || !codes || !bitmap || !rank || !n_exc)
    return am::fail(AM_EINVAL, "am_pack_host: bad args");
  init_b2c();
  const int64_t n_groups = total / AM_GROUP_BASES;
  const int64_t n_bw = am_pack_bitmap_words(total);
  memset(codes, 0, (size_t)am_pack_code_words(total) * sizeof(uint32_t));
  memset(bitmap, 0, (size_t)n_bw * sizeof(uint32_t));
  // full-size temporary masks only for flagged groups: collect per thread
  if (n_threads < 1) n_threads = 1;
  n_threads = (int)std::min<int64_t>(n_threads, std::max<int64_t>(1, n));
  struct Exc { int64_t group; uint64_t mask; };
  std::vector<std::vector<Exc>> found(n_threads);
  // contiguous contig ranges balanced by bases
  std::vector<int64_t> cut(n_threads + 1, n);
  cut[0] = 0;
  {
    int64_t tot = offsets[n] - offsets[0];
    int t = 1;
    for (int64_t i = 0; i < n && t < n_threads; ++i)
      while (t < n_threads && offsets[i] - offsets[0] >= tot * t / n_threads) cut[t++] = i;
  }
  auto work = [&](int t) {
    auto &ex = found[t];
    for (int64_t i = cut[t]; i < cut[t + 1]; ++i) {
      const uint8_t *s = seq + offsets[i];
      const int64_t L = offsets[i + 1] - offsets[i];
      const int64_t g0 = starts[i] / AM_GROUP_BASES;
      for (int64_t b = 0; b < L; b += AM_GROUP_BASES) {
        const int64_t m = std::min<int64_t>(AM_GROUP_BASES, L - b);
        uint32_t w[4] = {0, 0, 0, 0};
        uint64_t bad = 0;
        for (int64_t j = 0; j < m; ++j) {
          uint8_t c = g_b2c[s[b + j]];
          if (c == 255) { bad |= 1ull << j; c = 0; }
          w[j >> 4] |= (uint32_t)c << (2 * (j & 15));
        }
        const int64_t g = g0 + b / AM_GROUP_BASES;
        uint32_t *dst = codes + g * 4;
        dst[0] = w[0]; dst[1] = w[1]; dst[2] = w[2]; dst[3] = w[3];
        if (bad) ex.push_back({g, bad});
      }
    }
  };
  if (n_threads == 1) {
    work(0);
  } else {
    std::vector<std::thread> th;
    for (int t = 0; t < n_threads; ++t) th.emplace_back(work, t);
    for (auto &x : th) x.join();
  }
  int64_t cnt = 0;
  for (auto &v : found) cnt += (int64_t)v.size();
  *n_exc = cnt;
  if (cnt > exc_capacity) return am::fail(AM_ECAPACITY, "am_pack_host: exception capacity too small");
  int64_t r = 0;
  for (auto &v : found)  // thread ranges are in contig order, groups ascending
    for (auto &e : v) {
      bitmap[e.group >> 5] |= 1u << (e.group & 31);
      exc_mask[r++] = e.mask;
    }
  uint32_t acc = 0;
  for (int64_t w = 0; w < n_bw; ++w) {
    rank[w] = acc;
    acc += (uint32_t)__builtin_popcount(bitmap[w]);
  }
  (void)n_groups;
  return AM_OK;
}

// ---- count work list -----------------------------------------------------------
int am_count_plan(const int32_t *lengths, int64_t n, int k, int32_t max_item_bases,
                  int32_t *items, int64_t item_cap, int64_t *n_items, int32_t *multi_rows,
                  int64_t multi_cap, int64_t *n_multi) {
  if (n < 0 || (n > 0 && !lengths) || !n_items || !n_multi || k < 1 || k > 7 || max_item_bases < AM_TILE_BASES ||
      max_item_bases % AM_TILE_BASES)
    return am::fail(AM_EINVAL, "am_count_plan: bad args");
  if (n > 0x7fffffff) return am::fail(AM_EINVAL, "am_count_plan: too many contigs");
  int64_t ni = 0, nm = 0;
  for (int64_t i = 0; i < n; ++i) {
    const int64_t nwin = (int64_t)lengths[i] - k;  // kmers.py:186 max_length = L - k
    if (nwin <= 0) {
      // still emit an empty item so the kernel writes the all-zero row
      if (ni < item_cap) { int32_t *it = items + 4 * ni; it[0] = (int32_t)i; it[1] = 0; it[2] = 0; it[3] = 0; }
      ++ni;
      continue;
    }
    const bool multi = nwin > max_item_bases;
    if (multi) { if (nm < multi_cap) multi_rows[nm] = (int32_t)i; ++nm; }
    for (int64_t w = 0; w < nwin; w += max_item_bases) {
      if (ni < item_cap) {
        int32_t *it = items + 4 * ni;
        it[0] = (int32_t)i;
        it[1] = (int32_t)w;
        it[2] = (int32_t)std::min<int64_t>(max_item_bases, nwin - w);
        it[3] = multi ? 1 : 0;
      }
      ++ni;
    }
  }
  *n_items = ni;
  *n_multi = nm;
  if (ni > item_cap || nm > multi_cap) return am::fail(AM_ECAPACITY, "am_count_plan: capacity too small");
  return AM_OK;
}

}  // extern "C"
