// Kernel 4 — DBSCAN(min_samples=1) eps step as connected components (sm_100a).
//
// Replaces DBSCAN(eps, min_samples=1, n_jobs).fit(X).labels_ at
// autometa/binning/recursive_dbscan.py:109.  With min_samples=1 every point is a
// core point: labels are the connected components of the eps-graph, numbered by
// ascending minimum member row (the order sklearn/cluster/_dbscan_inner.pyx visits).
// Pair predicate restates sklearn's KD-tree leaf test
// (sklearn/neighbors/_binary_tree.pxi.tp:1952-1957 + euclidean rdist): fp64,
// d = sum_j (a_j - b_j)^2 accumulated in column order WITHOUT fma contraction,
// joined iff d <= eps*eps.
//
// Per eps: points are binned into a hashed uniform grid of cell edge eps*(1+2^-20)
// (counting sort by bucket), then every point scans the 3^F neighbouring cells and
// unions itself (lock-free union-find, root = smallest row index) with every
// neighbour of smaller row index that passes the predicate.  The forest persists
// across the sweep's increasing eps values, so each step only adds merges.
#include "am_common.cuh"
#include "am_scan.cuh"

namespace {

constexpr int MAXF = 4;
constexpr unsigned FULL = 0xffffffffu;
// tables up to this many points take eps_union_warp_kernel (a warp per point), larger ones eps_union_kernel
// (a thread per point).  Measured on B200: the warp form wins at every size — 2000 points 103 -> ~10 us per
// step, 200k points (C4) 23.3 -> 13.1 ms for the 48-eps grid — so the thread form is only reachable through
// the measurement knob AM_EPS_WARP_MAX.
constexpr int WARP_PER_POINT_MAX = 0x7fffffff;

struct GridParams {
  double mn[MAXF];
  double inv_cell;
  int F;
  unsigned nbuckets_mask;
};

__device__ __forceinline__ unsigned hash_cell(const int *c, int F, unsigned mask) {
  unsigned h = 0x9E3779B9u;
  for (int d = 0; d < F; ++d) {
    h ^= (unsigned)c[d] + 0x9E3779B9u + (h << 6) + (h >> 2);
    h *= 0x85EBCA6Bu;
    h ^= h >> 13;
  }
  return h & mask;
}

__device__ __forceinline__ void cell_of(const double *x, const GridParams &gp, int *c) {
  for (int d = 0; d < gp.F; ++d) c[d] = (int)floor((x[d] - gp.mn[d]) * gp.inv_cell);
}

// `dyn` (device, may be null): {1 / cell edge, eps^2} of the step, read from memory instead of the launch
// parameters so that ONE captured launch sequence serves every eps of a sweep (am_sweep.cu)
__global__ void bucket_count_kernel(const double *__restrict__ X, int n, GridParams gp,
                                    const double *__restrict__ dyn,
                                    unsigned *__restrict__ bucket_of, unsigned *__restrict__ counts) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (dyn) gp.inv_cell = dyn[0];
  double x[MAXF];
  int c[MAXF];
  for (int d = 0; d < gp.F; ++d) x[d] = __ldg(X + (int64_t)i * gp.F + d);
  cell_of(x, gp, c);
  const unsigned b = hash_cell(c, gp.F, gp.nbuckets_mask);
  bucket_of[i] = b;
  atomicAdd(counts + b, 1u);
}

__global__ void bucket_scatter_kernel(const double *__restrict__ X, int n, int F,
                                      const unsigned *__restrict__ bucket_of,
                                      const unsigned *__restrict__ bucket_start,
                                      unsigned *__restrict__ fill, int *__restrict__ sorted_idx,
                                      double *__restrict__ sorted_x, const int *__restrict__ parent,
                                      int *__restrict__ sorted_root) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const unsigned b = bucket_of[i];
  const unsigned pos = bucket_start[b] + atomicAdd(fill + b, 1u);
  sorted_idx[pos] = i;
  sorted_root[pos] = parent[i];  // root at the start of the step (the forest is flat between steps)
  for (int d = 0; d < F; ++d) sorted_x[(int64_t)d * n + pos] = __ldg(X + (int64_t)i * F + d);
}

__device__ __forceinline__ int uf_find(int *parent, int x) {
  int p = *((volatile int *)(parent + x));
  while (p != x) {
    const int gp = *((volatile int *)(parent + p));
    if (gp != p) atomicMin(parent + x, gp);  // path halving (monotone: parents only decrease)
    x = p;
    p = gp;
  }
  return x;
}

__device__ __forceinline__ void uf_union(int *parent, int a, int b) {
  while (true) {
    a = uf_find(parent, a);
    b = uf_find(parent, b);
    if (a == b) return;
    if (a < b) { const int t = a; a = b; b = t; }  // a > b: hook the larger root under the smaller
    const int old = atomicCAS(parent + a, a, b);
    if (old == a) return;
  }
}

// per bucket: do all its points already belong to ONE component?  (roots at the start of the step:
// the forest is flattened after every step, so root(i) = parent[i])
__global__ void bucket_roots_kernel(int n, const unsigned *__restrict__ bucket_of, const int *__restrict__ parent,
                                    int *__restrict__ rmin_inv, int *__restrict__ rmax) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int r = parent[i];
  const unsigned b = bucket_of[i];
  atomicMax(rmin_inv + b, 0x7fffffff - r);
  atomicMax(rmax + b, r);
}

// one thread per point (in bucket-sorted order)
__global__ void __launch_bounds__(128)
eps_union_kernel(int n, GridParams gp, double eps2, const double *__restrict__ dyn,
                 const unsigned *__restrict__ bucket_start,
                 const int *__restrict__ sorted_idx, const double *__restrict__ sorted_x,
                 const int *__restrict__ rmin_inv, const int *__restrict__ rmax,
                 const int *__restrict__ sorted_root, int *__restrict__ parent) {
  const int pos = blockIdx.x * blockDim.x + threadIdx.x;
  if (pos >= n) return;
  if (dyn) {
    gp.inv_cell = dyn[0];
    eps2 = dyn[1];
  }
  const int F = gp.F;
  double x[MAXF];
  int c[MAXF];
  for (int d = 0; d < F; ++d) x[d] = sorted_x[(int64_t)d * n + pos];
  cell_of(x, gp, c);
  const int me = sorted_idx[pos];
  // root at the start of the step (other threads only ever hook roots under smaller roots, so a
  // bucket that was uniformly in my component then is still entirely connected to me)
  // (a late read may return a smaller ancestor instead: still a vertex I am connected to)
  const int myroot0 = *((volatile int *)(parent + me));
  int ncomb = 1;
  for (int d = 0; d < F; ++d) ncomb *= 3;
  unsigned seen[81];  // buckets already scanned by this thread (3^4 max)
  int nseen = 0;
  for (int comb = 0; comb < ncomb; ++comb) {
    int cc[MAXF];
    int t = comb;
    for (int d = 0; d < F; ++d) {
      cc[d] = c[d] + (t % 3) - 1;
      t /= 3;
    }
    const unsigned b = hash_cell(cc, F, gp.nbuckets_mask);
    bool dup = false;
    for (int s = 0; s < nseen; ++s) dup |= (seen[s] == b);
    if (dup) continue;
    seen[nseen++] = b;
    if (__ldg(rmax + b) == myroot0 && 0x7fffffff - __ldg(rmin_inv + b) == myroot0) continue;  // nothing to merge
    const unsigned beg = bucket_start[b], end = bucket_start[b + 1];
    for (unsigned j = beg; j < end; ++j) {
      // already in my component when the step began: nothing to merge, skip the distance test
      // (4 sequential bytes instead of 24 + the arithmetic; most candidates at large eps)
      if (__ldg(sorted_root + j) == myroot0) continue;
      const int other = sorted_idx[j];
      if (other >= me) continue;  // each unordered pair once, from its larger row
      double dist = 0.0;
      for (int d = 0; d < F; ++d) {
        const double tmp = __dsub_rn(x[d], sorted_x[(int64_t)d * n + j]);
        dist = __dadd_rn(dist, __dmul_rn(tmp, tmp));
      }
      if (dist <= eps2) uf_union(parent, me, other);
    }
  }
}

// One WARP per point (the production form).  With one thread per point every thread walks its 3^F cells one
// after the other — ~50 dependent global loads — so the pass is pure latency: 100 us for a 2000-point table
// (16 CTAs on an otherwise empty GPU), and at 200k points only ~10 warps per scheduler with nothing to overlap
// their stalls (ncu: issue active 29 %, long-scoreboard stalls on top).  Here the lanes look up the 3^F cells
// at once (hash, the uniform-root test, the bucket bounds), then the warp goes through the cells that are
// left, 32 candidates at a time.  Same tests, same unions, same forest as eps_union_kernel.
template <int F>
__global__ void __launch_bounds__(256)
eps_union_warp_kernel(int n, GridParams gp, double eps2, const double *__restrict__ dyn,
                      const unsigned *__restrict__ bucket_start, const int *__restrict__ sorted_idx,
                      const double *__restrict__ sorted_x, const int *__restrict__ rmin_inv,
                      const int *__restrict__ rmax, const int *__restrict__ sorted_root, int *__restrict__ parent) {
  const int lane = threadIdx.x & 31;
  const int pos = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  if (pos >= n) return;  // the whole warp leaves together
  if (dyn) {
    gp.inv_cell = dyn[0];
    eps2 = dyn[1];
  }
  // F is a template parameter: the per-feature loops unroll and x / c / cc live in registers (with a run-time
  // F they were indexed stack arrays — half of the kernel's instructions were local loads and stores)
  double x[F];
  int c[F];
#pragma unroll
  for (int d = 0; d < F; ++d) x[d] = sorted_x[(int64_t)d * n + pos];
#pragma unroll
  for (int d = 0; d < F; ++d) c[d] = (int)floor((x[d] - gp.mn[d]) * gp.inv_cell);
  const int me = sorted_idx[pos];
  const int myroot0 = *((volatile int *)(parent + me));
  constexpr int NCOMB = F == 1 ? 3 : F == 2 ? 9 : F == 3 ? 27 : 81;
  for (int base = 0; base < NCOMB; base += 32) {
    const int comb = base + lane;
    bool act = comb < NCOMB;
    unsigned b = 0xffffffffu, beg = 0, end = 0;
    if (act) {
      int t = comb;
      unsigned h = 0x9E3779B9u;  // hash_cell, unrolled
#pragma unroll
      for (int d = 0; d < F; ++d) {
        const int ccd = c[d] + (t % 3) - 1;
        t /= 3;
        h ^= (unsigned)ccd + 0x9E3779B9u + (h << 6) + (h >> 2);
        h *= 0x85EBCA6Bu;
        h ^= h >> 13;
      }
      b = h & gp.nbuckets_mask;
    }
    // a bucket reached through several cells (hash collisions) is scanned once per round of 32 cells; a
    // second scan in a later round (F = 4) would only repeat unions that are already done
    const unsigned same = __match_any_sync(FULL, b);
    if (act && (__ffs(same) - 1) != lane) act = false;
    if (act) {
      // all four words of the bucket in flight together (one latency, not two)
      const int hi = __ldg(rmax + b), lo_inv = __ldg(rmin_inv + b);
      beg = __ldg(bucket_start + b);
      end = __ldg(bucket_start + b + 1);
      act = end > beg && !(hi == myroot0 && 0x7fffffff - lo_inv == myroot0);  // else: nothing to merge
    }
    unsigned todo = __ballot_sync(FULL, act);
    while (todo) {
      const int src = __ffs(todo) - 1;
      todo &= todo - 1;
      const unsigned cb = __shfl_sync(FULL, beg, src), ce = __shfl_sync(FULL, end, src);
      for (unsigned j = cb + lane; j < ce; j += 32) {
        const int root_j = __ldg(sorted_root + j), other = __ldg(sorted_idx + j);
        if (root_j == myroot0) continue;  // already in my component when the step began
        if (other >= me) continue;        // each unordered pair once, from its larger row
        double dist = 0.0;
#pragma unroll
        for (int d = 0; d < F; ++d) {
          const double tmp = __dsub_rn(x[d], __ldg(sorted_x + (int64_t)d * n + j));
          dist = __dadd_rn(dist, __dmul_rn(tmp, tmp));
        }
        if (dist <= eps2) uf_union(parent, me, other);
      }
    }
  }
}

__global__ void compress_kernel(int *__restrict__ parent, int n, unsigned *__restrict__ is_root) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int r = i;
  while (true) {
    const int p = *((volatile int *)(parent + r));
    if (p == r) break;
    r = p;
  }
  is_root[i] = (r == i) ? 1u : 0u;
  // do not write parent[i] here: another thread may still be walking through i
}

__global__ void relabel_kernel(int *__restrict__ parent, int n, const unsigned *__restrict__ root_rank,
                               int64_t *__restrict__ labels, int *__restrict__ n_clusters) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0) n_clusters[0] = (int)root_rank[n];
  if (i >= n) return;
  int r = i;
  while (true) {
    const int p = __ldg(parent + r);
    if (p == r) break;
    r = p;
  }
  labels[i] = (int64_t)root_rank[r];
}

__global__ void flatten_kernel(int *__restrict__ parent, const int64_t *__restrict__ labels,
                               const unsigned *__restrict__ is_root, int n, int *__restrict__ root_of_label) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (is_root[i]) root_of_label[labels[i]] = i;
}

__global__ void flatten2_kernel(int *__restrict__ parent, const int64_t *__restrict__ labels, int n,
                                const int *__restrict__ root_of_label) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  parent[i] = root_of_label[labels[i]];
}

__global__ void iota_kernel(int *v, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    v[i] = (int)i;
}

unsigned pick_buckets(int64_t n) {
  unsigned b = 1024;
  while ((int64_t)b < 2 * n && b < (1u << 26)) b <<= 1;
  return b;
}

struct Work {
  unsigned *bucket_of;     // n
  unsigned *counts;        // nb
  unsigned *fill;          // nb
  unsigned *bucket_start;  // nb + 1
  int *sorted_idx;         // n
  double *sorted_x;        // MAXF * n
  unsigned *is_root;       // n
  unsigned *root_rank;     // n + 1
  int *root_of_label;      // n
  int *rmin_inv;           // nb
  int *rmax;               // nb
  int *sorted_root;        // n
  unsigned *scan_tmp;      // scan_tmp_words(max(n, nb))
};

inline int64_t align_up(int64_t x) { return (x + 255) & ~(int64_t)255; }

int64_t carve(void *base, int64_t n, Work *w) {
  const unsigned nb = pick_buckets(n);
  int64_t off = 0;
  char *p = (char *)base;
  auto take = [&](int64_t bytes) {
    char *q = p ? p + off : nullptr;
    off += align_up(bytes);
    return q;
  };
  Work t;
  t.bucket_of = (unsigned *)take(n * 4);
  t.counts = (unsigned *)take((int64_t)nb * 4);
  t.fill = (unsigned *)take((int64_t)nb * 4);
  t.bucket_start = (unsigned *)take(((int64_t)nb + 1) * 4);
  t.sorted_idx = (int *)take(n * 4);
  t.sorted_x = (double *)take((int64_t)MAXF * n * 8);
  t.is_root = (unsigned *)take(n * 4);
  t.root_rank = (unsigned *)take((n + 1) * 4);
  t.root_of_label = (int *)take(n * 4);
  t.rmin_inv = (int *)take((int64_t)nb * 4);
  t.rmax = (int *)take((int64_t)nb * 4);
  t.sorted_root = (int *)take(n * 4);
  t.scan_tmp = (unsigned *)take(am::scan_tmp_words(std::max<int64_t>(n + 1, nb + 1)) * 4);
  if (w) *w = t;
  return off;
}

}  // namespace

extern "C" int64_t am_dbscan_work_bytes(int64_t n_points) {
  if (n_points < 0) return 0;
  return carve(nullptr, n_points, nullptr);
}

extern "C" int am_iota_i32(int32_t *d_v, int64_t n, void *stream) {
  if (n == 0) return AM_OK;
  if (!d_v || n < 0) return am::fail(AM_EINVAL, "am_iota_i32: bad args");
  iota_kernel<<<(unsigned)std::min<int64_t>((n + 255) / 256, 4096), 256, 0, (cudaStream_t)stream>>>(d_v, n);
  AM_LAUNCHED("iota_kernel");
  return AM_OK;
}

// the launch sequence of one eps step; `dyn` null: cell edge / eps^2 from (inv_cell, eps2)
static int eps_union_launch(const double *d_X, int64_t n_points, int F, const double *h_min, double inv_cell,
                            double eps2, const double *d_dyn, int32_t *d_parent, void *d_work, cudaStream_t st) {
  const int n = (int)n_points;
  Work w;
  carve(d_work, n_points, &w);
  const unsigned nb = pick_buckets(n_points);
  GridParams gp;
  for (int d = 0; d < MAXF; ++d) gp.mn[d] = d < F ? h_min[d] : 0.0;
  gp.inv_cell = inv_cell;
  gp.F = F;
  gp.nbuckets_mask = nb - 1;
  // counts | fill are adjacent in the workspace (nb * 4 is a multiple of the 256-byte carving unit)
  AM_CUDA(cudaMemsetAsync(w.counts, 0, (size_t)((char *)(w.fill + nb) - (char *)w.counts), st));
  const int blocks = (n + 255) / 256;
  bucket_count_kernel<<<blocks, 256, 0, st>>>(d_X, n, gp, d_dyn, w.bucket_of, w.counts);
  AM_LAUNCHED("bucket_count_kernel");
  am::g_launches.fetch_add(am::exclusive_scan(w.counts, w.bucket_start, (int)nb, w.scan_tmp, st) - 1);
  AM_LAUNCHED("exclusive_scan");
  bucket_scatter_kernel<<<blocks, 256, 0, st>>>(d_X, n, F, w.bucket_of, w.bucket_start, w.fill,
                                                w.sorted_idx, w.sorted_x, d_parent, w.sorted_root);
  AM_LAUNCHED("bucket_scatter_kernel");
  AM_CUDA(cudaMemsetAsync(w.rmin_inv, 0, (size_t)nb * 4, st));
  AM_CUDA(cudaMemsetAsync(w.rmax, 0xFF, (size_t)nb * 4, st));
  bucket_roots_kernel<<<blocks, 256, 0, st>>>(n, w.bucket_of, d_parent, w.rmin_inv, w.rmax);
  AM_LAUNCHED("bucket_roots_kernel");
  static const int warp_max = [] {
    const char *v = getenv("AM_EPS_WARP_MAX");  // measurement knob
    return v ? atoi(v) : WARP_PER_POINT_MAX;
  }();
  if (n <= warp_max) {
    const unsigned wblocks = (unsigned)(((int64_t)n + 7) / 8);
#define AM_WARP_UNION(FF)                                                                                      \
  eps_union_warp_kernel<FF><<<wblocks, 256, 0, st>>>(n, gp, eps2, d_dyn, w.bucket_start, w.sorted_idx, w.sorted_x, \
                                                     w.rmin_inv, w.rmax, w.sorted_root, d_parent)
    switch (F) {
      case 1: AM_WARP_UNION(1); break;
      case 2: AM_WARP_UNION(2); break;
      case 3: AM_WARP_UNION(3); break;
      default: AM_WARP_UNION(4); break;
    }
#undef AM_WARP_UNION
    AM_LAUNCHED("eps_union_warp_kernel");
  } else {
    eps_union_kernel<<<(n + 127) / 128, 128, 0, st>>>(n, gp, eps2, d_dyn, w.bucket_start, w.sorted_idx,
                                                      w.sorted_x, w.rmin_inv, w.rmax, w.sorted_root, d_parent);
    AM_LAUNCHED("eps_union_kernel");
  }
  return AM_OK;
}

extern "C" int am_eps_grid(double eps, const double *h_min, const double *h_max, int F, double *out2) {
  if (F < 1 || F > MAXF) return am::fail(AM_EINVAL, "am_eps_grid: F must be in 1..4");
  if (!(eps > 0.0)) return am::fail(AM_EINVAL, "am_eps_grid: eps must be > 0");
  if (!h_min || !h_max || !out2) return am::fail(AM_EINVAL, "am_eps_grid: null pointer");
  double span[MAXF];
  for (int d = 0; d < F; ++d) {
    span[d] = h_max[d] - h_min[d];
    if (!(span[d] >= 0.0)) return am::fail(AM_EINVAL, "am_eps_grid: bad bounds (NaN?)");
  }
  am::eps_grid(eps, span, F, out2);
  return AM_OK;
}

extern "C" int am_eps_union(const double *d_X, int64_t n_points, int F, double eps,
                            const double *h_min, const double *h_max, int32_t *d_parent,
                            void *d_work, void *stream) {
  if (F < 1 || F > MAXF) return am::fail(AM_EINVAL, "am_eps_union: F must be in 1..4");
  if (n_points < 0 || n_points > 0x7fffffff) return am::fail(AM_EINVAL, "am_eps_union: bad n_points");
  if (!(eps > 0.0)) return am::fail(AM_EINVAL, "am_eps_union: eps must be > 0");
  if (n_points == 0) return AM_OK;
  if (!d_X || !h_min || !h_max || !d_parent || !d_work) return am::fail(AM_EINVAL, "am_eps_union: null pointer");
  double g[2];
  const int rc = am_eps_grid(eps, h_min, h_max, F, g);
  if (rc) return rc;
  return eps_union_launch(d_X, n_points, F, h_min, g[0], g[1], nullptr, d_parent, d_work, (cudaStream_t)stream);
}

extern "C" int am_eps_union_dyn(const double *d_X, int64_t n_points, int F, const double *h_min,
                                const double *d_dyn, int32_t *d_parent, void *d_work, void *stream) {
  if (F < 1 || F > MAXF) return am::fail(AM_EINVAL, "am_eps_union_dyn: F must be in 1..4");
  if (n_points < 0 || n_points > 0x7fffffff) return am::fail(AM_EINVAL, "am_eps_union_dyn: bad n_points");
  if (n_points == 0) return AM_OK;
  if (!d_X || !h_min || !d_dyn || !d_parent || !d_work) return am::fail(AM_EINVAL, "am_eps_union_dyn: null pointer");
  return eps_union_launch(d_X, n_points, F, h_min, 0.0, 0.0, d_dyn, d_parent, d_work, (cudaStream_t)stream);
}

extern "C" int am_labels_from_parent(int32_t *d_parent, int64_t n_points, int64_t *d_labels,
                                     int32_t *d_n_clusters, void *d_work, void *stream) {
  if (n_points < 0 || n_points > 0x7fffffff) return am::fail(AM_EINVAL, "am_labels_from_parent: bad n");
  if (n_points == 0) return AM_OK;
  if (!d_parent || !d_labels || !d_n_clusters || !d_work)
    return am::fail(AM_EINVAL, "am_labels_from_parent: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  const int n = (int)n_points;
  Work w;
  carve(d_work, n_points, &w);
  const int blocks = (n + 255) / 256;
  compress_kernel<<<blocks, 256, 0, st>>>(d_parent, n, w.is_root);
  AM_LAUNCHED("compress_kernel");
  am::g_launches.fetch_add(am::exclusive_scan(w.is_root, w.root_rank, n, w.scan_tmp, st) - 1);
  AM_LAUNCHED("exclusive_scan");
  relabel_kernel<<<blocks, 256, 0, st>>>(d_parent, n, w.root_rank, d_labels, d_n_clusters);
  AM_LAUNCHED("relabel_kernel");
  // flatten the forest (parent[i] = root) so the next, larger eps starts from depth-1 trees
  flatten_kernel<<<blocks, 256, 0, st>>>(d_parent, d_labels, w.is_root, n, w.root_of_label);
  AM_LAUNCHED("flatten_kernel");
  flatten2_kernel<<<blocks, 256, 0, st>>>(d_parent, d_labels, n, w.root_of_label);
  AM_LAUNCHED("flatten2_kernel");
  return AM_OK;
}
