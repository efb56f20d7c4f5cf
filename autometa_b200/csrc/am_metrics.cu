// Per-eps cluster metrics + cutoff filter + median completeness on the device (sm_100a).
//
// Replaces, inside the eps sweep (autometa/binning/recursive_dbscan.py:180-192), the pandas
// join + groupby of add_metrics (autometa/binning/utilities.py:125-207) and the cutoff filter of
// apply_binning_metrics_filter (utilities.py:210-262):
//   per cluster: marker sums -> present (>= 1) and single-copy (== 1) marker counts (:178-182),
//   completeness = present / M * 100 (:187), purity = single / present * 100 (:189),
//   coverage / GC standard deviation with ddof = 1 (:191-192, NaN for single-contig clusters),
//   pass = completeness >= cc, purity >= pc, cov sigma <= cs, gc sigma <= gs (:257-262, NaN fails),
//   median completeness over the passing clusters (recursive_dbscan.py:186-192).
// Everything is a segmented reduction over the label vector with atomics; the contig x marker
// table is sparse (COO) and is accumulated into a dense [clusters x M] scratch that the counting
// pass zeroes again as it reads it (atomicExch), so no per-eps memset of the scratch is needed.
#include "am_common.cuh"

namespace {

struct MetricsBufs {
  int *cnt;          // [n] contigs per cluster
  double *s_cov;     // [n] sum -> mean
  double *s_gc;      // [n]
  double *ss_cov;    // [n] sum of squared deviations
  double *ss_gc;     // [n]
  int *present;      // [n]
  int *single;       // [n]
  int *hist;         // [M + 2] histogram of `present` over passing clusters
  int *table;        // [n * M] marker sums scratch (kept all-zero between calls)
};

inline int64_t al(int64_t x) { return (x + 255) & ~(int64_t)255; }

int64_t carve_metrics(void *base, int64_t n, int M, MetricsBufs *b) {
  int64_t off = 0;
  char *p = (char *)base;
  auto take = [&](int64_t bytes) {
    char *q = p ? p + off : nullptr;
    off += al(bytes);
    return q;
  };
  MetricsBufs t;
  // the first block (cnt .. hist) is zeroed on every call with one memset
  t.cnt = (int *)take(n * 4);
  t.s_cov = (double *)take(n * 8);
  t.s_gc = (double *)take(n * 8);
  t.ss_cov = (double *)take(n * 8);
  t.ss_gc = (double *)take(n * 8);
  t.present = (int *)take(n * 4);
  t.single = (int *)take(n * 4);
  t.hist = (int *)take((int64_t)(M + 2) * 4);
  t.table = (int *)take(n * (int64_t)M * 4);
  if (b) *b = t;
  return off;
}

// Warp-aggregated segmented sums: lanes holding the same cluster label combine their values with
// shuffles and only the group leader issues the atomics.  With few clusters (late in the sweep) this
// turns 200k same-address fp64 atomics per array into ~1/10 as many; with many clusters every group
// has one lane and the loop runs once.
__device__ __forceinline__ unsigned group_of(int key, bool active) {
  // inactive lanes get a key no active lane can have
  return __match_any_sync(0xffffffffu, active ? key : -1 - (int)(threadIdx.x & 31));
}
__device__ __forceinline__ double group_sum(unsigned mask, double v) {
  double s = 0.0;
  for (unsigned m = mask; m; m &= m - 1) s += __shfl_sync(mask, v, __ffs(m) - 1);
  return s;
}

__global__ void point_sums_kernel(const int64_t *__restrict__ labels, int n, const double *__restrict__ cov,
                                  const double *__restrict__ gc, int *cnt, double *s_cov, double *s_gc) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const bool active = i < n;
  const int c = active ? (int)labels[i] : 0;
  const unsigned mask = group_of(c, active);
  const double sc = group_sum(mask, active ? cov[i] : 0.0), sg = group_sum(mask, active ? gc[i] : 0.0);
  if (active && (int)(threadIdx.x & 31) == __ffs(mask) - 1) {
    atomicAdd(cnt + c, __popc(mask));
    atomicAdd(s_cov + c, sc);
    atomicAdd(s_gc + c, sg);
  }
}

__global__ void point_ss_kernel(const int64_t *__restrict__ labels, int n, const double *__restrict__ cov,
                                const double *__restrict__ gc, const int *__restrict__ cnt,
                                const double *__restrict__ s_cov, const double *__restrict__ s_gc,
                                double *ss_cov, double *ss_gc) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const bool active = i < n;
  const int c = active ? (int)labels[i] : 0;
  double dc = 0.0, dg = 0.0;
  if (active) {
    const double k = (double)cnt[c];
    dc = cov[i] - s_cov[c] / k;
    dg = gc[i] - s_gc[c] / k;
  }
  const unsigned mask = group_of(c, active);
  const double sc = group_sum(mask, dc * dc), sg = group_sum(mask, dg * dg);
  if (active && (int)(threadIdx.x & 31) == __ffs(mask) - 1) {
    atomicAdd(ss_cov + c, sc);
    atomicAdd(ss_gc + c, sg);
  }
}

__global__ void marker_add_kernel(const int64_t *__restrict__ labels, const int *__restrict__ rows,
                                  const int *__restrict__ cols, const int *__restrict__ vals, int nnz, int M,
                                  int *table) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= nnz) return;
  const int64_t c = labels[rows[k]];
  atomicAdd(table + c * M + cols[k], vals[k]);
}

__global__ void marker_count_kernel(const int64_t *__restrict__ labels, const int *__restrict__ rows,
                                    const int *__restrict__ cols, int nnz, int M, int *table, int *present,
                                    int *single) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= nnz) return;
  const int64_t c = labels[rows[k]];
  // the first entry of each (cluster, marker) pair to arrive takes the sum and leaves zero behind
  const int s = atomicExch(table + c * M + cols[k], 0);
  if (s >= 1) atomicAdd(present + (int)c, 1);
  if (s == 1) atomicAdd(single + (int)c, 1);
}

__global__ void finalize_kernel(const int *__restrict__ n_clusters, int M, const int *__restrict__ cnt,
                                const double *__restrict__ ss_cov, const double *__restrict__ ss_gc,
                                const int *__restrict__ present, const int *__restrict__ single, double cc,
                                double pc, double cs, double gs, double *__restrict__ completeness,
                                double *__restrict__ purity, double *__restrict__ cov_std,
                                double *__restrict__ gc_std, int *__restrict__ present_out,
                                int *__restrict__ single_out, int *hist) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= *n_clusters) return;
  present_out[c] = present[c];
  single_out[c] = single[c];
  const double comp = __dmul_rn(__ddiv_rn((double)present[c], (double)M), 100.0);
  const double pur = __dmul_rn(__ddiv_rn((double)single[c], (double)present[c]), 100.0);  // 0/0 -> NaN like pandas
  const double k1 = (double)cnt[c] - 1.0;
  const double sc = sqrt(ss_cov[c] / k1), sg = sqrt(ss_gc[c] / k1);    // x/0 -> NaN (0/0) for singletons
  completeness[c] = comp;
  purity[c] = pur;
  cov_std[c] = sc;
  gc_std[c] = sg;
  if (comp >= cc && pur >= pc && sc <= cs && sg <= gs) atomicAdd(hist + present[c], 1);
}

// median of completeness over the passing clusters = numpy.median of values present/M*100
__global__ void median_kernel(const int *__restrict__ hist, int M, const int *__restrict__ n_clusters,
                              double *__restrict__ summary) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  long long tot = 0;
  for (int v = 0; v <= M; ++v) tot += hist[v];
  double med = -INFINITY;
  if (tot > 0) {
    const long long lo_rank = (tot - 1) / 2, hi_rank = tot / 2;  // 0-based order statistics
    long long acc = 0;
    int vlo = -1, vhi = -1;
    for (int v = 0; v <= M; ++v) {
      const long long nxt = acc + hist[v];
      if (vlo < 0 && lo_rank < nxt) vlo = v;
      if (vhi < 0 && hi_rank < nxt) vhi = v;
      acc = nxt;
    }
    // numpy / pandas: mean of the two middle values; no FMA contraction so the bits match
    const double a = __dmul_rn(__ddiv_rn((double)vlo, (double)M), 100.0);
    const double b = __dmul_rn(__ddiv_rn((double)vhi, (double)M), 100.0);
    med = (vlo == vhi) ? a : __ddiv_rn(__dadd_rn(a, b), 2.0);
  }
  summary[0] = (double)*n_clusters;
  summary[1] = med;
  summary[2] = (double)tot;
}

}  // namespace

extern "C" int64_t am_metrics_work_bytes(int64_t n_points, int n_markers) {
  if (n_points < 0 || n_markers < 1) return 0;
  return carve_metrics(nullptr, std::max<int64_t>(n_points, 1), n_markers, nullptr);
}

// d_work must be zero-filled ONCE by the caller before the first call (the marker scratch is kept
// all-zero by the kernels themselves).
extern "C" int am_cluster_metrics(const int64_t *d_labels, int64_t n_points, const int32_t *d_n_clusters,
                                  const int32_t *d_marker_rows, const int32_t *d_marker_cols,
                                  const int32_t *d_marker_vals, int64_t nnz, int n_markers,
                                  const double *d_coverage, const double *d_gc, double completeness_cutoff,
                                  double purity_cutoff, double coverage_stddev_cutoff,
                                  double gc_content_stddev_cutoff, double *d_completeness, double *d_purity,
                                  double *d_cov_std, double *d_gc_std, int32_t *d_present, int32_t *d_single,
                                  double *d_summary, void *d_work, void *stream) {
  if (n_points < 0 || n_points > 0x7fffffff || n_markers < 1 || nnz < 0 || nnz > 0x7fffffff)
    return am::fail(AM_EINVAL, "am_cluster_metrics: bad shape");
  if (n_points == 0) return AM_OK;
  if (!d_labels || !d_n_clusters || !d_coverage || !d_gc || !d_completeness || !d_purity || !d_cov_std ||
      !d_gc_std || !d_present || !d_single || !d_summary || !d_work || (nnz > 0 && (!d_marker_rows || !d_marker_cols || !d_marker_vals)))
    return am::fail(AM_EINVAL, "am_cluster_metrics: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  const int n = (int)n_points, M = n_markers;
  MetricsBufs b;
  carve_metrics(d_work, n_points, M, &b);
  // no memset here: the caller zero-filled d_work once and every call leaves it zeroed (below)
  const int blocks = (n + 255) / 256;
  point_sums_kernel<<<blocks, 256, 0, st>>>(d_labels, n, d_coverage, d_gc, b.cnt, b.s_cov, b.s_gc);
  AM_LAUNCHED("point_sums_kernel");
  point_ss_kernel<<<blocks, 256, 0, st>>>(d_labels, n, d_coverage, d_gc, b.cnt, b.s_cov, b.s_gc, b.ss_cov, b.ss_gc);
  AM_LAUNCHED("point_ss_kernel");
  if (nnz > 0) {
    const int mb = (int)((nnz + 255) / 256);
    marker_add_kernel<<<mb, 256, 0, st>>>(d_labels, d_marker_rows, d_marker_cols, d_marker_vals, (int)nnz, M, b.table);
    AM_LAUNCHED("marker_add_kernel");
    marker_count_kernel<<<mb, 256, 0, st>>>(d_labels, d_marker_rows, d_marker_cols, (int)nnz, M, b.table, b.present, b.single);
    AM_LAUNCHED("marker_count_kernel");
  }
  finalize_kernel<<<blocks, 256, 0, st>>>(d_n_clusters, M, b.cnt, b.ss_cov, b.ss_gc, b.present, b.single,
                                          completeness_cutoff, purity_cutoff, coverage_stddev_cutoff,
                                          gc_content_stddev_cutoff, d_completeness, d_purity, d_cov_std, d_gc_std,
                                          d_present, d_single, b.hist);
  AM_LAUNCHED("finalize_kernel");
  median_kernel<<<1, 32, 0, st>>>(b.hist, M, d_n_clusters, d_summary);
  AM_LAUNCHED("median_kernel");
  // leave the WHOLE workspace zeroed (the marker table already is: marker_count_kernel clears what it reads), so
  // that one zero-filled buffer can serve sweeps of any (n, M) without a per-sweep n x M memset
  AM_CUDA(cudaMemsetAsync(b.cnt, 0, (size_t)((char *)b.table - (char *)b.cnt), st));
  return AM_OK;
}
