// Host-side TSV writer / reader for the k-mer tables: the wire format
// between count -> normalize -> embed (autometa/common/kmers.py:116,328,431,704).
// The reference goes through pandas (`df.to_csv(out, sep="\t", index=True, header=True)`,
// `pd.read_csv(fpath, sep="\t", index_col="contig")`); at 200k x 512 that is minutes per
// table.  These routines produce byte-identical files / bit-identical values on all host
// threads:
//   * floats are written as Python's repr(float) (shortest round-trip digits, fixed
//     notation for 1e-4 <= |x| < 1e16, else d.ddde+XX), NaN as an empty cell;
//   * counts are written as integers, 0 (= pd.NA in the reference's frame) as an empty cell;
//   * cells are parsed with the arithmetic of pandas' default C-engine converter
//     (precise_xstrtod of pandas/_libs/src/parser/tokenizer.c: at most 17 significant
//     digits accumulated in a double, one multiply/divide by an exact power of ten), so a
//     table read here equals the frame pandas would have built, not merely the
//     correctly-rounded value.
#include <algorithm>
#include <charconv>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "am_common.cuh"

namespace {

inline char *put_int(int64_t v, char *p) {
  auto r = std::to_chars(p, p + 24, v);
  return r.ptr;
}

// repr(float) for a finite or infinite double (NaN handled by the caller)
inline char *put_repr(double x, char *p) {
  if (std::isinf(x)) {
    if (x < 0) *p++ = '-';
    memcpy(p, "inf", 3);
    return p + 3;
  }
  if (x == 0.0) {
    if (std::signbit(x)) *p++ = '-';
    memcpy(p, "0.0", 3);
    return p + 3;
  }
  char tmp[48];
  auto r = std::to_chars(tmp, tmp + sizeof(tmp), x, std::chars_format::scientific);
  char *s = tmp;
  if (*s == '-') *p++ = *s++;
  char *e = s;
  while (*e != 'e') ++e;
  int exp10 = 0;
  for (char *q = e + 2; q < r.ptr; ++q) exp10 = exp10 * 10 + (*q - '0');  // to_chars does not terminate the text
  if (e[1] == '-') exp10 = -exp10;
  if (exp10 < -4 || exp10 >= 16) {  // scientific: to_chars' text is repr's
    const size_t n = (size_t)(r.ptr - s);
    memcpy(p, s, n);
    return p + n;
  }
  // digits without the point
  char d[24];
  int nd = 0;
  d[nd++] = s[0];
  for (char *q = s + 2; q < e; ++q) d[nd++] = *q;
  if (exp10 >= 0) {
    const int int_len = exp10 + 1;
    for (int i = 0; i < int_len; ++i) *p++ = i < nd ? d[i] : '0';
    *p++ = '.';
    if (nd > int_len) {
      memcpy(p, d + int_len, (size_t)(nd - int_len));
      p += nd - int_len;
    } else {
      *p++ = '0';
    }
  } else {
    *p++ = '0';
    *p++ = '.';
    for (int i = 0; i < -exp10 - 1; ++i) *p++ = '0';
    memcpy(p, d, (size_t)nd);
    p += nd;
  }
  return p;
}

template <typename Cell>
int write_table(const char *path, const uint8_t *header, int64_t header_len, const uint8_t *ids,
                const int64_t *id_offsets, int64_t n_rows, int64_t n_cols, int n_threads, Cell cell) {
  FILE *fh = fopen(path, "wb");
  if (!fh) return am::fail(AM_EINVAL, "am_tsv_write: cannot open output file");
  if (header_len && fwrite(header, 1, (size_t)header_len, fh) != (size_t)header_len) {
    fclose(fh);
    return am::fail(AM_EINVAL, "am_tsv_write: short write");
  }
  n_threads = std::max(1, std::min(n_threads, 256));
  const int64_t rows_per_job = 128;  // 1.7 MB of text per job at 512 columns: the buffers stay cache-resident
  const int64_t wave = rows_per_job * n_threads;
  // two sets of per-thread buffers: wave w is written to the file (one writer thread, in order) while wave w + 1 is
  // being formatted on all threads
  std::vector<std::string> bufs[2] = {std::vector<std::string>((size_t)n_threads),
                                      std::vector<std::string>((size_t)n_threads)};
  std::vector<size_t> used[2] = {std::vector<size_t>((size_t)n_threads, 0), std::vector<size_t>((size_t)n_threads, 0)};
  std::atomic<bool> ok{true};
  std::thread writer;
  int set = 0;
  for (int64_t r0 = 0; r0 < n_rows && ok.load(); r0 += wave, set ^= 1) {
    const int64_t r1 = std::min(n_rows, r0 + wave);
    std::vector<std::string> &cur = bufs[set];
    auto job = [&](int t) {
      const int64_t a = std::min(r1, r0 + t * rows_per_job), b = std::min(r1, a + rows_per_job);
      std::string &out = cur[(size_t)t];
      size_t need = 0;
      for (int64_t r = a; r < b; ++r) need += (size_t)(id_offsets[r + 1] - id_offsets[r]) + (size_t)n_cols * 26 + 2;
      if (out.size() < need) out.resize(need);  // grows once; never shrunk (no zero-fill per wave)
      char *p = out.data();
      for (int64_t r = a; r < b; ++r) {
        const size_t len = (size_t)(id_offsets[r + 1] - id_offsets[r]);
        memcpy(p, ids + id_offsets[r], len);
        p += len;
        for (int64_t c = 0; c < n_cols; ++c) {
          *p++ = '\t';
          p = cell(r, c, p);
        }
        *p++ = '\n';
      }
      used[set][(size_t)t] = (size_t)(p - out.data());
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < n_threads; ++t) pool.emplace_back(job, t);
    job(0);
    for (auto &th : pool) th.join();
    if (writer.joinable()) writer.join();  // the previous wave is on disk; its buffers are free for wave w + 1
    writer = std::thread([&, set]() {
      for (int t = 0; t < n_threads && ok.load(); ++t) {
        const size_t n = used[set][(size_t)t];
        if (n && fwrite(bufs[set][(size_t)t].data(), 1, n, fh) != n) ok.store(false);
        used[set][(size_t)t] = 0;
      }
    });
  }
  if (writer.joinable()) writer.join();
  if (fclose(fh) != 0) ok.store(false);
  return ok.load() ? AM_OK : am::fail(AM_EINVAL, "am_tsv_write: short write");
}

// ---- reader --------------------------------------------------------------------------
const double kPow10[] = {
    1e0,   1e1,   1e2,   1e3,   1e4,   1e5,   1e6,   1e7,   1e8,   1e9,   1e10,  1e11,  1e12,  1e13,  1e14,
    1e15,  1e16,  1e17,  1e18,  1e19,  1e20,  1e21,  1e22,  1e23,  1e24,  1e25,  1e26,  1e27,  1e28,  1e29,
    1e30,  1e31,  1e32,  1e33,  1e34,  1e35,  1e36,  1e37,  1e38,  1e39,  1e40,  1e41,  1e42,  1e43,  1e44,
    1e45,  1e46,  1e47,  1e48,  1e49,  1e50,  1e51,  1e52,  1e53,  1e54,  1e55,  1e56,  1e57,  1e58,  1e59,
    1e60,  1e61,  1e62,  1e63,  1e64,  1e65,  1e66,  1e67,  1e68,  1e69,  1e70,  1e71,  1e72,  1e73,  1e74,
    1e75,  1e76,  1e77,  1e78,  1e79,  1e80,  1e81,  1e82,  1e83,  1e84,  1e85,  1e86,  1e87,  1e88,  1e89,
    1e90,  1e91,  1e92,  1e93,  1e94,  1e95,  1e96,  1e97,  1e98,  1e99,  1e100, 1e101, 1e102, 1e103, 1e104,
    1e105, 1e106, 1e107, 1e108, 1e109, 1e110, 1e111, 1e112, 1e113, 1e114, 1e115, 1e116, 1e117, 1e118, 1e119,
    1e120, 1e121, 1e122, 1e123, 1e124, 1e125, 1e126, 1e127, 1e128, 1e129, 1e130, 1e131, 1e132, 1e133, 1e134,
    1e135, 1e136, 1e137, 1e138, 1e139, 1e140, 1e141, 1e142, 1e143, 1e144, 1e145, 1e146, 1e147, 1e148, 1e149,
    1e150, 1e151, 1e152, 1e153, 1e154, 1e155, 1e156, 1e157, 1e158, 1e159, 1e160, 1e161, 1e162, 1e163, 1e164,
    1e165, 1e166, 1e167, 1e168, 1e169, 1e170, 1e171, 1e172, 1e173, 1e174, 1e175, 1e176, 1e177, 1e178, 1e179,
    1e180, 1e181, 1e182, 1e183, 1e184, 1e185, 1e186, 1e187, 1e188, 1e189, 1e190, 1e191, 1e192, 1e193, 1e194,
    1e195, 1e196, 1e197, 1e198, 1e199, 1e200, 1e201, 1e202, 1e203, 1e204, 1e205, 1e206, 1e207, 1e208, 1e209,
    1e210, 1e211, 1e212, 1e213, 1e214, 1e215, 1e216, 1e217, 1e218, 1e219, 1e220, 1e221, 1e222, 1e223, 1e224,
    1e225, 1e226, 1e227, 1e228, 1e229, 1e230, 1e231, 1e232, 1e233, 1e234, 1e235, 1e236, 1e237, 1e238, 1e239,
    1e240, 1e241, 1e242, 1e243, 1e244, 1e245, 1e246, 1e247, 1e248, 1e249, 1e250, 1e251, 1e252, 1e253, 1e254,
    1e255, 1e256, 1e257, 1e258, 1e259, 1e260, 1e261, 1e262, 1e263, 1e264, 1e265, 1e266, 1e267, 1e268, 1e269,
    1e270, 1e271, 1e272, 1e273, 1e274, 1e275, 1e276, 1e277, 1e278, 1e279, 1e280, 1e281, 1e282, 1e283, 1e284,
    1e285, 1e286, 1e287, 1e288, 1e289, 1e290, 1e291, 1e292, 1e293, 1e294, 1e295, 1e296, 1e297, 1e298, 1e299,
    1e300, 1e301, 1e302, 1e303, 1e304, 1e305, 1e306, 1e307, 1e308};

inline bool is_digit(uint8_t c) { return (unsigned)(c - '0') < 10u; }

// One numeric cell [p, e) -> double with pandas' default converter arithmetic.  Returns false
// when the cell is not a plain number (the caller then reports the table as unsupported).
inline bool parse_cell(const uint8_t *p, const uint8_t *e, double *out, bool *plain_int) {
  *plain_int = false;
  while (p < e && (*p == ' ' || *p == '\t')) ++p;
  if (p == e) {
    *out = std::nan("");
    return true;
  }
  {  // [+-]digits, at most 18 of them: what pandas keeps as int64
    const uint8_t *q = p;
    if (*q == '-' || *q == '+') ++q;
    const uint8_t *d0 = q;
    while (q < e && is_digit(*q)) ++q;
    *plain_int = q == e && q > d0 && q - d0 <= 18;
  }
  const uint8_t *start = p;
  bool neg = false;
  if (*p == '-') { neg = true; ++p; } else if (*p == '+') { ++p; }
  double number = 0.0;
  int exponent = 0, num_digits = 0, num_decimals = 0;
  const int max_digits = 17;
  while (p < e && is_digit(*p)) {
    if (num_digits < max_digits) { number = number * 10.0 + (*p - '0'); ++num_digits; } else { ++exponent; }
    ++p;
  }
  if (p < e && *p == '.') {
    ++p;
    while (num_digits < max_digits && p < e && is_digit(*p)) {
      number = number * 10.0 + (*p - '0');
      ++p; ++num_digits; ++num_decimals;
    }
    if (num_digits >= max_digits) while (p < e && is_digit(*p)) ++p;
    exponent -= num_decimals;
  }
  if (num_digits == 0) {
    // the NA / infinity spellings pandas maps to floats by default
    const size_t len = (size_t)(e - start);
    auto is = [&](const char *s) { return len == strlen(s) && memcmp(start, s, len) == 0; };
    if (is("nan") || is("NaN") || is("NA") || is("N/A") || is("NULL") || is("null") || is("None") || is("n/a") ||
        is("<NA>") || is("#N/A") || is("#NA") || is("-NaN") || is("-nan") || is("#N/A N/A") || is("1.#IND") ||
        is("-1.#IND") || is("1.#QNAN") || is("-1.#QNAN")) {
      *out = std::nan("");
      return true;
    }
    if (is("inf") || is("Inf") || is("+inf") || is("+Inf") || is("Infinity") || is("infinity") || is("+Infinity")) {
      *out = HUGE_VAL;
      return true;
    }
    if (is("-inf") || is("-Inf") || is("-Infinity") || is("-infinity")) {
      *out = -HUGE_VAL;
      return true;
    }
    return false;
  }
  if (neg) number = -number;
  if (p < e && (*p == 'e' || *p == 'E')) {
    ++p;
    bool eneg = false;
    if (p < e && *p == '-') { eneg = true; ++p; } else if (p < e && *p == '+') { ++p; }
    int n = 0, nd = 0;
    while (nd < max_digits && p < e && is_digit(*p)) { n = n * 10 + (*p - '0'); ++nd; ++p; }
    if (nd == 0) return false;
    exponent += eneg ? -n : n;
  }
  while (p < e && *p == ' ') ++p;
  if (p != e) return false;
  if (exponent > 308) return false;
  if (exponent > 0) {
    number *= kPow10[exponent];
  } else if (exponent < -308) {
    if (exponent < -616) number = 0.0;
    else { number /= kPow10[-308 - exponent]; number /= kPow10[308]; }
  } else {
    number /= kPow10[-exponent];
  }
  if (number == HUGE_VAL || number == -HUGE_VAL) return false;
  *out = number;
  return true;
}

}  // namespace

extern "C" int am_tsv_write_f64(const char *path, const uint8_t *header, int64_t header_len, const uint8_t *ids,
                                const int64_t *id_offsets, int64_t n_rows, int64_t n_cols, const double *values,
                                int n_threads) {
  if (!path || n_rows < 0 || n_cols < 0 || (n_rows && (!ids || !id_offsets)) || (n_rows && n_cols && !values))
    return am::fail(AM_EINVAL, "am_tsv_write_f64: bad args");
  return write_table(path, header, header_len, ids, id_offsets, n_rows, n_cols, n_threads,
                     [=](int64_t r, int64_t c, char *p) {
                       const double x = values[r * n_cols + c];
                       return x != x ? p : put_repr(x, p);
                     });
}

extern "C" int am_tsv_write_i32(const char *path, const uint8_t *header, int64_t header_len, const uint8_t *ids,
                                const int64_t *id_offsets, int64_t n_rows, int64_t n_cols, const int32_t *counts,
                                int n_threads) {
  if (!path || n_rows < 0 || n_cols < 0 || (n_rows && (!ids || !id_offsets)) || (n_rows && n_cols && !counts))
    return am::fail(AM_EINVAL, "am_tsv_write_i32: bad args");
  return write_table(path, header, header_len, ids, id_offsets, n_rows, n_cols, n_threads,
                     [=](int64_t r, int64_t c, char *p) {
                       const int32_t x = counts[r * n_cols + c];
                       return x == 0 ? p : put_int(x, p);
                     });
}

// Shape of a TSV in memory: number of data rows (lines after the header, blank lines skipped)
// and the number of tab-separated fields of the header line.
extern "C" int am_tsv_shape(const uint8_t *buf, int64_t len, int64_t *n_rows, int64_t *n_fields) {
  if (!buf || len < 0 || !n_rows || !n_fields) return am::fail(AM_EINVAL, "am_tsv_shape: bad args");
  const uint8_t *nl = (const uint8_t *)memchr(buf, '\n', (size_t)len);
  const int64_t hend = nl ? (int64_t)(nl - buf) : len;
  int64_t f = 1;
  for (int64_t i = 0; i < hend; ++i) f += buf[i] == '\t';
  int64_t rows = 0, p = hend + 1;
  while (p < len) {
    const uint8_t *q = (const uint8_t *)memchr(buf + p, '\n', (size_t)(len - p));
    const int64_t e = q ? (int64_t)(q - buf) : len;
    int64_t e2 = e;
    if (e2 > p && buf[e2 - 1] == '\r') --e2;
    if (e2 > p) ++rows;
    p = e + 1;
  }
  *n_rows = rows;
  *n_fields = f;
  return AM_OK;
}

// Parse the data rows: field 0 of every row is the index label ([id_begin, id_end) into buf),
// fields 1..n_fields-1 go to values [n_rows x (n_fields-1)].  AM_EINVAL (with a message) for
// anything this fast path does not cover (quoted fields, ragged rows, non-numeric cells): the
// caller then uses its general reader.  col_all_int[j] = 1 when every cell of column j is a
// plain integer literal (pandas then makes the column int64 instead of float64).
extern "C" int am_tsv_parse_f64(const uint8_t *buf, int64_t len, int64_t n_rows, int64_t n_fields, int64_t *id_begin,
                                int64_t *id_end, double *values, uint8_t *col_all_int, int n_threads) {
  if (!buf || len < 0 || n_rows < 0 || n_fields < 1 || (n_rows && (!id_begin || !id_end)) ||
      (n_rows && n_fields > 1 && !values))
    return am::fail(AM_EINVAL, "am_tsv_parse_f64: bad args");
  // line starts of the data rows
  std::vector<int64_t> starts;
  starts.reserve((size_t)n_rows + 1);
  const uint8_t *nl = (const uint8_t *)memchr(buf, '\n', (size_t)len);
  int64_t p = nl ? (int64_t)(nl - buf) + 1 : len;
  while (p < len) {
    const uint8_t *q = (const uint8_t *)memchr(buf + p, '\n', (size_t)(len - p));
    const int64_t e = q ? (int64_t)(q - buf) : len;
    int64_t e2 = e;
    if (e2 > p && buf[e2 - 1] == '\r') --e2;
    if (e2 > p) {
      if ((int64_t)starts.size() / 2 >= n_rows) return am::fail(AM_EINVAL, "am_tsv_parse_f64: more rows than announced");
      starts.push_back(p);
      starts.push_back(e2);
    }
    p = e + 1;
  }
  if ((int64_t)starts.size() / 2 != n_rows) return am::fail(AM_EINVAL, "am_tsv_parse_f64: fewer rows than announced");
  n_threads = (int)std::max<int64_t>(1, std::min<int64_t>(std::min(n_threads, 256), n_rows));
  std::vector<int> bad((size_t)n_threads, 0);
  const int64_t ncol = n_fields - 1;
  std::vector<std::vector<uint8_t>> not_int((size_t)n_threads, std::vector<uint8_t>((size_t)ncol, 0));
  auto job = [&](int t) {
    const int64_t a = n_rows * t / n_threads, b = n_rows * (t + 1) / n_threads;
    for (int64_t r = a; r < b; ++r) {
      const uint8_t *s = buf + starts[(size_t)(2 * r)], *e = buf + starts[(size_t)(2 * r + 1)];
      const uint8_t *f = (const uint8_t *)memchr(s, '\t', (size_t)(e - s));
      const uint8_t *fe = f ? f : e;
      if (memchr(s, '"', (size_t)(fe - s))) { bad[(size_t)t] = 1; return; }
      id_begin[r] = (int64_t)(s - buf);
      id_end[r] = (int64_t)(fe - buf);
      const uint8_t *c = f ? f + 1 : e;
      for (int64_t j = 0; j < ncol; ++j) {
        if (!f) { bad[(size_t)t] = 2; return; }  // fewer fields than the header
        const uint8_t *g = (const uint8_t *)memchr(c, '\t', (size_t)(e - c));
        const uint8_t *ce = g ? g : e;
        bool plain_int;
        if (!parse_cell(c, ce, values + r * ncol + j, &plain_int)) { bad[(size_t)t] = 3; return; }
        if (!plain_int) not_int[(size_t)t][(size_t)j] = 1;
        if (j + 1 < ncol) { f = g; c = g ? g + 1 : e; }
        else if (g) { bad[(size_t)t] = 2; return; }  // more fields than the header
      }
    }
  };
  std::vector<std::thread> pool;
  for (int t = 1; t < n_threads; ++t) pool.emplace_back(job, t);
  job(0);
  for (auto &th : pool) th.join();
  for (int t = 0; t < n_threads; ++t) {
    if (bad[(size_t)t] == 1) return am::fail(AM_EINVAL, "am_tsv_parse_f64: quoted index label");
    if (bad[(size_t)t] == 2) return am::fail(AM_EINVAL, "am_tsv_parse_f64: ragged row");
    if (bad[(size_t)t] == 3) return am::fail(AM_EINVAL, "am_tsv_parse_f64: non-numeric cell");
  }
  if (col_all_int)
    for (int64_t j = 0; j < ncol; ++j) {
      uint8_t any = 0;
      for (int t = 0; t < n_threads; ++t) any |= not_int[(size_t)t][(size_t)j];
      col_all_int[j] = n_rows > 0 && !any;
    }
  return AM_OK;
}
