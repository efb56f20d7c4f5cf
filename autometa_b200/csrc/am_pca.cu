// Kernel 3 (fp64 baseline path) — column sums, centred Gram, projection.
//
// Restates sklearn 1.9 PCA(svd_solver="covariance_eigh") as called from
// autometa/common/kmers.py:605: mean (_pca.py:560), covariance via the Gram
// matrix (_pca.py:604-610) and the transform X @ Vt.T - mean @ Vt.T
// (sklearn/decomposition/_base.py:151-159).  The Gram here is the fp64 CUDA-core
// version: exact-enough reference for the tensor-core (tcgen05) Gram and the
// fallback for shapes the tensor path does not cover.  All reductions are
// deterministic (fixed split order), so multi-GPU results do not depend on timing.
#include "am_common.cuh"

namespace {

// ---------------------------------------------------------------- column sums
constexpr int CS_ROWS = 8;  // row lanes per CTA (blockDim = (128, CS_ROWS))

__global__ void colsum_partial_kernel(const double *__restrict__ X, int64_t n, int D,
                                      double *__restrict__ partials) {
  // grid.x = column blocks of 128, grid.y = row splits
  __shared__ double red[CS_ROWS][128];
  const int c = blockIdx.x * 128 + threadIdx.x;
  double s = 0.0;
  if (c < D)
    for (int64_t r = (int64_t)blockIdx.y * CS_ROWS + threadIdx.y; r < n; r += (int64_t)gridDim.y * CS_ROWS)
      s += __ldg(X + r * D + c);
  red[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.y == 0 && c < D) {
    double t = 0.0;
#pragma unroll
    for (int i = 0; i < CS_ROWS; ++i) t += red[i][threadIdx.x];
    partials[(int64_t)blockIdx.y * D + c] = t;
  }
}

__global__ void colsum_reduce_kernel(const double *__restrict__ partials, int n_parts, int D,
                                     double *__restrict__ colsum) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= D) return;
  double s = 0.0;
  for (int p = 0; p < n_parts; ++p) s += partials[(int64_t)p * D + j];
  colsum[j] += s;
}

// column sums AND column max-abs in one pass (the max-abs sizes the fixed-point scale of the
// tensor-core Gram, am_gram_i8.cu)
__global__ void colstats_partial_kernel(const double *__restrict__ X, int64_t n, int D,
                                        double *__restrict__ psum, double *__restrict__ pmax) {
  __shared__ double red[CS_ROWS][128];
  __shared__ double redm[CS_ROWS][128];
  const int c = blockIdx.x * 128 + threadIdx.x;
  double s = 0.0, m = 0.0;
  if (c < D)
    for (int64_t r = (int64_t)blockIdx.y * CS_ROWS + threadIdx.y; r < n; r += (int64_t)gridDim.y * CS_ROWS) {
      const double v = __ldg(X + r * D + c);
      s += v;
      m = fmax(m, fabs(v));
    }
  red[threadIdx.y][threadIdx.x] = s;
  redm[threadIdx.y][threadIdx.x] = m;
  __syncthreads();
  if (threadIdx.y == 0 && c < D) {
    double t = 0.0, tm = 0.0;
#pragma unroll
    for (int i = 0; i < CS_ROWS; ++i) { t += red[i][threadIdx.x]; tm = fmax(tm, redm[i][threadIdx.x]); }
    psum[(int64_t)blockIdx.y * D + c] = t;
    pmax[(int64_t)blockIdx.y * D + c] = tm;
  }
}

__global__ void colstats_reduce_kernel(const double *__restrict__ psum, const double *__restrict__ pmax,
                                       int n_parts, int D, double *__restrict__ colsum,
                                       double *__restrict__ colmax) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= D) return;
  double s = 0.0, m = 0.0;
  for (int p = 0; p < n_parts; ++p) { s += psum[(int64_t)p * D + j]; m = fmax(m, pmax[(int64_t)p * D + j]); }
  colsum[j] += s;
  colmax[j] = fmax(colmax[j], m);
}

int colsum_splits(int64_t n) {
  int64_t s = (int64_t)am::num_sms() * 2;
  const int64_t need = (n + CS_ROWS - 1) / CS_ROWS;
  if (s > need) s = need;
  return (int)(s < 1 ? 1 : s);
}

// ---------------------------------------------------------------- Gram (fp64)
constexpr int GT = 128;  // output tile edge
constexpr int GK = 16;   // rows of X per pipeline stage

__global__ void __launch_bounds__(256, 1)
gram_kernel(const double *__restrict__ X, int64_t n, int D, const double *__restrict__ mu,
            double *__restrict__ partials, int64_t rows_per_split, int T) {
  extern __shared__ double sm[];
  double(*As)[GK][GT] = reinterpret_cast<double(*)[GK][GT]>(sm);
  double(*Bs)[GK][GT] = reinterpret_cast<double(*)[GK][GT]>(sm + 2 * GK * GT);
  const int ntiles = T * (T + 1) / 2;
  const int tile = blockIdx.x % ntiles;
  const int split = blockIdx.x / ntiles;
  int ti = 0, rem = tile;
  while (rem >= T - ti) { rem -= T - ti; ++ti; }
  const int tj = ti + rem;
  const int tid = threadIdx.x;
  const int ty = tid >> 4, tx = tid & 15;
  const int64_t row_begin = (int64_t)split * rows_per_split;
  const int64_t row_end = min(n, row_begin + rows_per_split);
  // loader mapping: element q = tid + 256*h -> (kr = q / 128, c = q % 128); c is fixed per thread
  const int lc = tid & 127;
  const int lk0 = tid >> 7;  // 0 or 1; kr = lk0 + 2*h
  const int colA = ti * GT + lc, colB = tj * GT + lc;
  const bool okA = colA < D, okB = colB < D;
  const double muA = okA ? __ldg(mu + colA) : 0.0;
  const double muB = okB ? __ldg(mu + colB) : 0.0;

  double acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.0;

  double ra[8], rb[8];
  auto load_regs = [&](int64_t r0) {
#pragma unroll
    for (int h = 0; h < 8; ++h) {
      const int64_t r = r0 + lk0 + 2 * h;
      const bool okr = r < row_end;
      ra[h] = (okr && okA) ? __ldg(X + r * D + colA) - muA : 0.0;
      rb[h] = (okr && okB) ? __ldg(X + r * D + colB) - muB : 0.0;
    }
  };
  auto store_regs = [&](int buf) {
#pragma unroll
    for (int h = 0; h < 8; ++h) {
      As[buf][lk0 + 2 * h][lc] = ra[h];
      Bs[buf][lk0 + 2 * h][lc] = rb[h];
    }
  };

  if (row_begin < row_end) {
    load_regs(row_begin);
    store_regs(0);
    __syncthreads();
    int buf = 0;
    for (int64_t r0 = row_begin; r0 < row_end; r0 += GK) {
      const bool more = r0 + GK < row_end;
      if (more) load_regs(r0 + GK);
#pragma unroll
      for (int k = 0; k < GK; ++k) {
        double a[8], b[8];
        const double2 *ap = reinterpret_cast<const double2 *>(&As[buf][k][ty * 8]);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const double2 v = ap[i];
          a[2 * i] = v.x;
          a[2 * i + 1] = v.y;
        }
#pragma unroll
        for (int m = 0; m < 4; ++m) {
          const double2 v = *reinterpret_cast<const double2 *>(&Bs[buf][k][m * 32 + tx * 2]);
          b[2 * m] = v.x;
          b[2 * m + 1] = v.y;
        }
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);
      }
      if (more) store_regs(buf ^ 1);
      __syncthreads();
      buf ^= 1;
    }
  }
  double *P = partials + ((int64_t)split * ntiles + tile) * (GT * GT);
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int m = 0; m < 4; ++m)
      *reinterpret_cast<double2 *>(P + (ty * 8 + i) * GT + m * 32 + tx * 2) =
          make_double2(acc[i][2 * m], acc[i][2 * m + 1]);
}

__global__ void gram_reduce_kernel(const double *__restrict__ partials, int n_splits, int T, int D,
                                   double *__restrict__ G) {
  const int ntiles = T * (T + 1) / 2;
  const int tile = blockIdx.x;
  int ti = 0, rem = tile;
  while (rem >= T - ti) { rem -= T - ti; ++ti; }
  const int tj = ti + rem;
  for (int e = threadIdx.x; e < GT * GT; e += blockDim.x) {
    const int r = e / GT, c = e % GT;
    const int gi = ti * GT + r, gj = tj * GT + c;
    if (gi >= D || gj >= D) continue;
    double s = 0.0;
    for (int sp = 0; sp < n_splits; ++sp) s += partials[((int64_t)sp * ntiles + tile) * (GT * GT) + e];
    G[(int64_t)gi * D + gj] += s;
    if (ti != tj) G[(int64_t)gj * D + gi] += s;
  }
}

void gram_shape(int D, int64_t n, int *T, int *S, int64_t *rows_per_split) {
  *T = (D + GT - 1) / GT;
  const int ntiles = *T * (*T + 1) / 2;
  int s = am::num_sms() / ntiles;
  if (s < 1) s = 1;
  int64_t rps = (n + s - 1) / s;
  rps = (rps + GK - 1) / GK * GK;
  if (rps < GK) rps = GK;
  *rows_per_split = rps;
  *S = s;
}

// ---------------------------------------------------------------- projection
constexpr int PM = 128;  // rows per CTA
constexpr int PN = 64;   // components per CTA pass
constexpr int PK = 16;
constexpr int PA_PITCH = PM + 2;
constexpr int PB_PITCH = PN + 2;

__global__ void __launch_bounds__(256)
project_kernel(const double *__restrict__ X, int64_t n, int D, const double *__restrict__ mu,
               const double *__restrict__ comps, int P, double *__restrict__ scores) {
  __shared__ double As[PK][PA_PITCH];
  __shared__ double Bs[PK][PB_PITCH];
  const int tid = threadIdx.x;
  const int ty = tid >> 4, tx = tid & 15;  // 16 x 16 threads; thread tile 8 rows x 4 comps
  const int64_t row0 = (int64_t)blockIdx.x * PM;
  const int p0 = blockIdx.y * PN;
  double acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;
  // loaders: A chunk 128 rows x 16 cols: q = tid + 256*h (h<8): r = q / 16, kk = q % 16
  //          B chunk 64 comps x 16 cols: q = tid + 256*h (h<4): p = q / 16, kk = q % 16
  const int lkk = tid & 15;
  for (int k0 = 0; k0 < D; k0 += PK) {
    const int col = k0 + lkk;
    const bool okc = col < D;
    const double m = okc ? __ldg(mu + col) : 0.0;
#pragma unroll
    for (int h = 0; h < 8; ++h) {
      const int r = (tid >> 4) + 16 * h;
      const int64_t gr = row0 + r;
      As[lkk][r] = (okc && gr < n) ? __ldg(X + gr * D + col) - m : 0.0;
    }
#pragma unroll
    for (int h = 0; h < 4; ++h) {
      const int p = (tid >> 4) + 16 * h;
      const int gp = p0 + p;
      Bs[lkk][p] = (okc && gp < P) ? __ldg(comps + (int64_t)gp * D + col) : 0.0;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < PK; ++k) {
      double a[8], b[4];
#pragma unroll
      for (int i = 0; i < 8; ++i) a[i] = As[k][ty * 8 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[k][tx + 16 * j];
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int64_t gr = row0 + ty * 8 + i;
    if (gr >= n) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int gp = p0 + tx + 16 * j;
      if (gp < P) scores[gr * P + gp] = acc[i][j];
    }
  }
}

}  // namespace

extern "C" int64_t am_colsum_work_bytes(int D) {
  return (int64_t)am::num_sms() * 2 * D * (int64_t)sizeof(double);
}

extern "C" int am_colsum(const double *d_X, int64_t n_rows, int D, double *d_colsum, void *d_work,
                         void *stream) {
  if (D < 1 || n_rows < 0) return am::fail(AM_EINVAL, "am_colsum: bad shape");
  if (n_rows == 0) return AM_OK;
  if (!d_X || !d_colsum || !d_work) return am::fail(AM_EINVAL, "am_colsum: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  const int S = colsum_splits(n_rows);
  dim3 grid((D + 127) / 128, S), block(128, CS_ROWS);
  colsum_partial_kernel<<<grid, block, 0, st>>>(d_X, n_rows, D, (double *)d_work);
  AM_LAUNCHED("colsum_partial_kernel");
  colsum_reduce_kernel<<<(D + 127) / 128, 128, 0, st>>>((const double *)d_work, S, D, d_colsum);
  AM_LAUNCHED("colsum_reduce_kernel");
  return AM_OK;
}

// mu = colsum / n;  C = (G - n (mu - centre)(mu - centre)^T) / (n - 1), in place on G.  n comes from
// device memory (it is an all-reduced tensor in the multi-GPU path).  One launch instead of ~8
// element-wise torch kernels between the Gram and the eigensolve.
__global__ void cov_finish_kernel(double *__restrict__ G, int D, const double *__restrict__ colsum,
                                  const double *__restrict__ center, const double *__restrict__ n_ptr,
                                  double *__restrict__ mu) {
  const double n = *n_ptr;
  const int64_t total = (int64_t)D * D;
  for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < total; q += (int64_t)gridDim.x * blockDim.x) {
    const int i = (int)(q / D), j = (int)(q % D);
    const double di = colsum[i] / n - (center ? center[i] : colsum[i] / n);
    const double dj = colsum[j] / n - (center ? center[j] : colsum[j] / n);
    G[q] = (G[q] - n * (di * dj)) / (n - 1.0);
    if (i == 0) mu[j] = colsum[j] / n;
  }
}

// flags[0] = min over columns of col_present, flags[1] = min over contigs of row_sum (am_clr's
// "column / row is NA everywhere" tests, kmers.py:369): one launch, read back once per step
__global__ void count_flags_kernel(const int *__restrict__ present, int ncols, const int *__restrict__ row_sum,
                                   int64_t n, int *__restrict__ flags) {
  __shared__ int s_min[32];
  int m = 0x7fffffff;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    m = min(m, row_sum[i]);
  for (int o = 16; o > 0; o >>= 1) m = min(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) s_min[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x < 32) {
    m = threadIdx.x < (blockDim.x >> 5) ? s_min[threadIdx.x] : 0x7fffffff;
    for (int o = 16; o > 0; o >>= 1) m = min(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (threadIdx.x == 0) atomicMin(flags + 1, m);
  }
  if (blockIdx.x == 0) {
    int p = 0x7fffffff;
    for (int c = threadIdx.x; c < ncols; c += blockDim.x) p = min(p, present[c]);
    for (int o = 16; o > 0; o >>= 1) p = min(p, __shfl_xor_sync(0xffffffffu, p, o));
    if ((threadIdx.x & 31) == 0) atomicMin(flags + 0, p);
  }
}

// SUM-reducible form of the two am_clr tests (kmers.py:369), so that they ride in the same all-reduce as
// the Gram: out[j] = 1.0 where column j occurs in this shard, out[ncols] = number of contigs without any
// counted window.  Integer-valued doubles: the atomic sum is exact in any order.
__global__ void stats_pack_kernel(const int *__restrict__ present, int ncols, const int *__restrict__ row_sum,
                                  int64_t n, double *__restrict__ out) {
  __shared__ int s_cnt[32];
  int z = 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    z += row_sum[i] <= 0;
  for (int o = 16; o > 0; o >>= 1) z += __shfl_xor_sync(0xffffffffu, z, o);
  if ((threadIdx.x & 31) == 0) s_cnt[threadIdx.x >> 5] = z;
  __syncthreads();
  if (threadIdx.x < 32) {
    z = threadIdx.x < (blockDim.x >> 5) ? s_cnt[threadIdx.x] : 0;
    for (int o = 16; o > 0; o >>= 1) z += __shfl_xor_sync(0xffffffffu, z, o);
    if (threadIdx.x == 0 && z) atomicAdd(out + ncols, (double)z);
  }
  if (blockIdx.x == 0)
    for (int c = threadIdx.x; c < ncols; c += blockDim.x) out[c] = present[c] > 0 ? 1.0 : 0.0;
}

// G += n [ (mu - c_new)(mu - c_new)^T - (mu - c_old)(mu - c_old)^T ],  mu = colsum / n:
// sum_r (x_r - a)(x_r - a)^T = sum_r (x_r - mu)(x_r - mu)^T + n (mu - a)(mu - a)^T for any a, so this moves a
// Gram taken about c_old to one about c_new.  The multi-GPU path uses it to turn each rank's Gram about
// its own provisional centre into one about a centre every rank knows without communicating
// (c_new = NULL: the origin, which is what sklearn's X^T X - n mu mu^T uses, _pca.py:604-610), so that
// one all-reduce of [G | colsum | n] is the only exchange.
__global__ void gram_recenter_kernel(double *__restrict__ G, int D, const double *__restrict__ colsum,
                                     const double *__restrict__ c_old, const double *__restrict__ c_new,
                                     const double *__restrict__ n_ptr) {
  const double n = *n_ptr;
  if (!(n > 0.0)) return;
  const int64_t total = (int64_t)D * D;
  for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < total; q += (int64_t)gridDim.x * blockDim.x) {
    const int i = (int)(q / D), j = (int)(q % D);
    const double mi = colsum[i] / n, mj = colsum[j] / n;
    const double oi = mi - (c_old ? c_old[i] : 0.0), oj = mj - (c_old ? c_old[j] : 0.0);
    const double ni = mi - (c_new ? c_new[i] : 0.0), nj = mj - (c_new ? c_new[j] : 0.0);
    G[q] += n * (ni * nj - oi * oj);
  }
}

extern "C" int64_t am_colstats_work_bytes(int D) { return 2 * am_colsum_work_bytes(D); }

extern "C" int am_colstats(const double *d_X, int64_t n_rows, int D, double *d_colsum,
                           double *d_colmaxabs, void *d_work, void *stream) {
  if (D < 1 || n_rows < 0) return am::fail(AM_EINVAL, "am_colstats: bad shape");
  if (n_rows == 0) return AM_OK;
  if (!d_X || !d_colsum || !d_colmaxabs || !d_work) return am::fail(AM_EINVAL, "am_colstats: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  const int S = colsum_splits(n_rows);
  double *psum = (double *)d_work;
  double *pmax = psum + (int64_t)am::num_sms() * 2 * D;
  dim3 grid((D + 127) / 128, S), block(128, CS_ROWS);
  colstats_partial_kernel<<<grid, block, 0, st>>>(d_X, n_rows, D, psum, pmax);
  AM_LAUNCHED("colstats_partial_kernel");
  colstats_reduce_kernel<<<(D + 127) / 128, 128, 0, st>>>(psum, pmax, S, D, d_colsum, d_colmaxabs);
  AM_LAUNCHED("colstats_reduce_kernel");
  return AM_OK;
}

extern "C" int am_cov_finish(double *d_G, int D, const double *d_colsum, const double *d_center,
                             const double *d_n, double *d_mu, void *stream) {
  if (D < 1) return am::fail(AM_EINVAL, "am_cov_finish: bad D");
  if (!d_G || !d_colsum || !d_n || !d_mu) return am::fail(AM_EINVAL, "am_cov_finish: null pointer");
  const int64_t total = (int64_t)D * D;
  cov_finish_kernel<<<(unsigned)std::min<int64_t>((total + 255) / 256, 1024), 256, 0, (cudaStream_t)stream>>>(
      d_G, D, d_colsum, d_center, d_n, d_mu);
  AM_LAUNCHED("cov_finish_kernel");
  return AM_OK;
}

extern "C" int am_count_flags(const int32_t *d_col_present, int ncols, const int32_t *d_row_sum, int64_t n_contigs,
                              int32_t *d_flags, void *stream) {
  if (!d_col_present || !d_flags || ncols < 1 || n_contigs < 0 || (n_contigs > 0 && !d_row_sum))
    return am::fail(AM_EINVAL, "am_count_flags: bad args");
  cudaStream_t st = (cudaStream_t)stream;
  AM_CUDA(cudaMemsetAsync(d_flags, 0x7f, 2 * sizeof(int32_t), st));
  const unsigned grid = (unsigned)std::max<int64_t>(1, std::min<int64_t>((n_contigs + 255) / 256, 296));
  count_flags_kernel<<<grid, 256, 0, st>>>(d_col_present, ncols, d_row_sum, n_contigs, d_flags);
  AM_LAUNCHED("count_flags_kernel");
  return AM_OK;
}

extern "C" int64_t am_gram_work_bytes(int D) {
  const int T = (D + GT - 1) / GT;
  const int ntiles = T * (T + 1) / 2;
  int s = am::num_sms() / ntiles;
  if (s < 1) s = 1;
  return (int64_t)s * ntiles * GT * GT * (int64_t)sizeof(double);
}

extern "C" int am_gram(const double *d_X, int64_t n_rows, int D, const double *d_mu, double *d_G,
                       void *d_work, void *stream) {
  if (D < 1 || n_rows < 0) return am::fail(AM_EINVAL, "am_gram: bad shape");
  if (n_rows == 0) return AM_OK;
  if (!d_X || !d_mu || !d_G || !d_work) return am::fail(AM_EINVAL, "am_gram: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  int T, S;
  int64_t rps;
  gram_shape(D, n_rows, &T, &S, &rps);
  const int ntiles = T * (T + 1) / 2;
  const size_t smem = (size_t)4 * GK * GT * sizeof(double);
  AM_CUDA(cudaFuncSetAttribute(gram_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  gram_kernel<<<ntiles * S, 256, smem, st>>>(d_X, n_rows, D, d_mu, (double *)d_work, rps, T);
  AM_LAUNCHED("gram_kernel");
  gram_reduce_kernel<<<ntiles, 256, 0, st>>>((const double *)d_work, S, T, D, d_G);
  AM_LAUNCHED("gram_reduce_kernel");
  return AM_OK;
}

extern "C" int am_project(const double *d_X, int64_t n_rows, int D, const double *d_mu,
                          const double *d_comps, int P, double *d_scores, void *stream) {
  if (D < 1 || P < 1 || n_rows < 0) return am::fail(AM_EINVAL, "am_project: bad shape");
  if (n_rows == 0) return AM_OK;
  if (!d_X || !d_mu || !d_comps || !d_scores) return am::fail(AM_EINVAL, "am_project: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  dim3 grid((unsigned)((n_rows + PM - 1) / PM), (P + PN - 1) / PN);
  project_kernel<<<grid, 256, 0, st>>>(d_X, n_rows, D, d_mu, d_comps, P, d_scores);
  AM_LAUNCHED("project_kernel");
  return AM_OK;
}

// Column presence and row sums of a ROW SUBSET of a count table (one partition of the large-data-mode driver,
// large_data_mode.py:454-470: counts.loc[counts.index.isin(rank_df.index)]): what am_count_kmers reports for a
// whole assembly, for tables that arrive as TSV.  One warp per row; presence is OR-ed per lane over the rows
// a warp visits and written once per warp.
__global__ void __launch_bounds__(256)
table_stats_kernel(const int32_t *__restrict__ counts, const int64_t *__restrict__ row_index, int64_t n, int D,
                   int32_t *__restrict__ present, int32_t *__restrict__ row_sum) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int c0 = 0; c0 < D; c0 += 32 * 32) {  // 32 columns per lane per sweep (bit j of `seen` <-> column c0 + 32 j + lane)
    uint32_t seen = 0u;
    for (int64_t r = warp; r < n; r += n_warps) {
      const int32_t *x = counts + (row_index ? row_index[r] : r) * (int64_t)D;
      int sum = 0;
#pragma unroll 4
      for (int j = 0; j < 32; ++j) {
        const int c = c0 + 32 * j + lane;
        const int v = c < D ? __ldg(x + c) : 0;
        sum += v;
        seen |= (uint32_t)(v > 0) << j;
      }
      for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
      if (lane == 0) {
        if (c0 == 0) row_sum[r] = sum;
        else row_sum[r] += sum;
      }
    }
    for (int j = 0; j < 32; ++j) {
      const int c = c0 + 32 * j + lane;
      if (c < D && ((seen >> j) & 1u)) present[c] = 1;
    }
  }
}

// Small fp64 vector helpers for the launch-lean pass (pipeline.FeatureStream queues a pass with ~15 C calls and
// no tensor-library kernel in between; at 25k contigs per GPU the host-side cost of queueing, not the GPU, bounds
// the stream): x[i] = x[i] * a + b, and the tail of the PCA fit — eigenvalues clamped at zero (_pca.py:630) and
// explained_variance_ratio_ = eigenvalue / trace (:645-646).
__global__ void axpb_kernel(double *__restrict__ x, int64_t n, double a, double b) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    x[i] = fma(x[i], a, b);
}

__global__ void pca_finish_kernel(double *__restrict__ evals, const double *__restrict__ total, int P,
                                  double *__restrict__ ratio) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < P) {
    const double v = fmax(evals[i], 0.0);
    evals[i] = v;
    ratio[i] = v / total[0];
  }
}

extern "C" int am_stats_pack(const int32_t *d_col_present, int ncols, const int32_t *d_row_sum, int64_t n_contigs,
                             double *d_out, void *stream) {
  if (!d_col_present || !d_out || ncols < 1 || n_contigs < 0 || (n_contigs > 0 && !d_row_sum))
    return am::fail(AM_EINVAL, "am_stats_pack: bad args");
  cudaStream_t st = (cudaStream_t)stream;
  AM_CUDA(cudaMemsetAsync(d_out + ncols, 0, sizeof(double), st));
  const unsigned grid = (unsigned)std::max<int64_t>(1, std::min<int64_t>((n_contigs + 255) / 256, 296));
  stats_pack_kernel<<<grid, 256, 0, st>>>(d_col_present, ncols, d_row_sum, n_contigs, d_out);
  AM_LAUNCHED("stats_pack_kernel");
  return AM_OK;
}

extern "C" int am_gram_recenter(double *d_G, int D, const double *d_colsum, const double *d_c_old,
                                const double *d_c_new, const double *d_n, void *stream) {
  if (D < 1) return am::fail(AM_EINVAL, "am_gram_recenter: bad D");
  if (!d_G || !d_colsum || !d_n) return am::fail(AM_EINVAL, "am_gram_recenter: null pointer");
  const int64_t total = (int64_t)D * D;
  gram_recenter_kernel<<<(unsigned)std::min<int64_t>((total + 255) / 256, 1024), 256, 0, (cudaStream_t)stream>>>(
      d_G, D, d_colsum, d_c_old, d_c_new, d_n);
  AM_LAUNCHED("gram_recenter_kernel");
  return AM_OK;
}

extern "C" int am_table_stats(const int32_t *d_counts, const int64_t *d_row_index, int64_t n_rows, int D,
                              int32_t *d_col_present, int32_t *d_row_sum, void *stream) {
  if (D < 1 || n_rows < 0) return am::fail(AM_EINVAL, "am_table_stats: bad shape");
  if (!d_col_present || (n_rows > 0 && (!d_counts || !d_row_sum))) return am::fail(AM_EINVAL, "am_table_stats: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  AM_CUDA(cudaMemsetAsync(d_col_present, 0, sizeof(int32_t) * (size_t)D, st));
  if (n_rows == 0) return AM_OK;
  const unsigned grid = (unsigned)std::max<int64_t>(1, std::min<int64_t>((n_rows + 7) / 8, (int64_t)am::num_sms() * 8));
  table_stats_kernel<<<grid, 256, 0, st>>>(d_counts, d_row_index, n_rows, D, d_col_present, d_row_sum);
  AM_LAUNCHED("table_stats_kernel");
  return AM_OK;
}

extern "C" int am_axpb_f64(double *d_x, int64_t n, double a, double b, void *stream) {
  if (n < 0 || (n > 0 && !d_x)) return am::fail(AM_EINVAL, "am_axpb_f64: bad args");
  if (n == 0) return AM_OK;
  const unsigned grid = (unsigned)std::min<int64_t>((n + 255) / 256, 1024);
  axpb_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(d_x, n, a, b);
  AM_LAUNCHED("axpb_kernel");
  return AM_OK;
}

extern "C" int am_zero_async(void *d_ptr, int64_t nbytes, void *stream) {
  if (nbytes < 0 || (nbytes > 0 && !d_ptr)) return am::fail(AM_EINVAL, "am_zero_async: bad args");
  if (nbytes) AM_CUDA(cudaMemsetAsync(d_ptr, 0, (size_t)nbytes, (cudaStream_t)stream));
  return AM_OK;
}

extern "C" int am_pca_finish(double *d_evals, const double *d_total, int P, double *d_ratio, void *stream) {
  if (P < 1 || !d_evals || !d_total || !d_ratio) return am::fail(AM_EINVAL, "am_pca_finish: bad args");
  pca_finish_kernel<<<(P + 127) / 128, 128, 0, (cudaStream_t)stream>>>(d_evals, d_total, P, d_ratio);
  AM_LAUNCHED("pca_finish_kernel");
  return AM_OK;
}
