// Symmetric eigensolver for the D x D covariance (D ~ 512) — one-sided (Hestenes)
// Jacobi in a single cooperative launch.
//
// Replaces scipy/numpy eigh inside sklearn PCA._fit_full(covariance_eigh)
// (sklearn/decomposition/_pca.py:611-624) plus svd_flip(u_based_decision=False)
// (sklearn/utils/extmath.py:973-981).  W starts as A, V as I; plane rotations are
// applied to row pairs (p,q) of W (and V) until all rows of W are mutually
// orthogonal: then row i of V is an eigenvector and lambda_i = w_i . v_i.
// Each round-robin step holds Dp/2 disjoint pairs, one warp per pair, so a step
// needs exactly one grid barrier.  Eigenvalues are returned in descending order.
#include "am_common.cuh"

namespace {

constexpr unsigned FULL = 0xffffffffu;
constexpr int EW = 2;  // warps per CTA

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
  return v;
}

__device__ __forceinline__ unsigned ld_acquire(const unsigned *p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

struct Ctl {
  unsigned barrier;      // monotonically increasing arrival counter
  unsigned pad0;
  unsigned long long offmax[2];  // per-sweep max |gamma|/sqrt(alpha beta), as ordered bits
  int sweeps_done;
  int converged;
};

__device__ __forceinline__ void grid_barrier(Ctl *ctl, unsigned nblocks, unsigned &epoch) {
  __syncthreads();
  if (threadIdx.x == 0) {
    ++epoch;
    __threadfence();
    atomicAdd(&ctl->barrier, 1u);
    const unsigned target = epoch * nblocks;
    while (ld_acquire(&ctl->barrier) < target) {
    }
    __threadfence();
  }
  __syncthreads();
}

__global__ void eigh_init_kernel(const double *__restrict__ A, int D, int Dp, double *__restrict__ W,
                                 double *__restrict__ V, Ctl *ctl, double pad_diag) {
  const int64_t n = (int64_t)Dp * Dp;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n;
       e += (int64_t)gridDim.x * blockDim.x) {
    const int i = (int)(e / Dp), j = (int)(e % Dp);
    double w = 0.0;
    if (i < D && j < D) w = 0.5 * (A[(int64_t)i * D + j] + A[(int64_t)j * D + i]);
    else if (i == j) w = pad_diag;
    W[e] = w;
    V[e] = (i == j) ? 1.0 : 0.0;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    ctl->barrier = 0;
    ctl->offmax[0] = 0ull;
    ctl->offmax[1] = 0ull;
    ctl->sweeps_done = 0;
    ctl->converged = 0;
  }
}

// One sweep = Dp-1 round-robin steps; pair i of step t (circle method):
//   i == 0 : (t, Dp-1);  i >= 1 : ((t+i) mod (Dp-1), (t-i) mod (Dp-1))
__global__ void __launch_bounds__(EW * 32)
jacobi_kernel(double *__restrict__ W, double *__restrict__ V, int Dp, int D, int max_sweeps,
              double tol, Ctl *ctl, double *__restrict__ evals, double *__restrict__ evecs,
              double *__restrict__ lam_tmp) {
  const int lane = threadIdx.x & 31;
  const int gw = blockIdx.x * EW + (threadIdx.x >> 5);
  const int nwarps = gridDim.x * EW;
  const int npairs = Dp / 2;
  const int m = Dp - 1;
  unsigned epoch = 0;
  int sweep = 0;
  for (; sweep < max_sweeps; ++sweep) {
    double local_max = 0.0;
    for (int t = 0; t < m; ++t) {
      for (int i = gw; i < npairs; i += nwarps) {
        int p, q;
        if (i == 0) { p = t; q = m; }
        else { p = (t + i) % m; q = (t - i + m) % m; }
        if (p > q) { const int x = p; p = q; q = x; }
        double *wp = W + (int64_t)p * Dp, *wq = W + (int64_t)q * Dp;
        double al = 0.0, be = 0.0, ga = 0.0;
        for (int j = lane; j < Dp; j += 32) {
          const double a = __ldcg(wp + j), b = __ldcg(wq + j);
          al = fma(a, a, al);
          be = fma(b, b, be);
          ga = fma(a, b, ga);
        }
        al = warp_sum(al); be = warp_sum(be); ga = warp_sum(ga);
        const double denom = sqrt(al) * sqrt(be);
        const double off = denom > 0.0 ? fabs(ga) / denom : 0.0;
        local_max = fmax(local_max, off);
        if (off > tol) {
          const double zeta = (be - al) / (2.0 * ga);
          const double tt = (zeta >= 0.0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
          const double c = 1.0 / sqrt(1.0 + tt * tt);
          const double s = c * tt;
          double *vp = V + (int64_t)p * Dp, *vq = V + (int64_t)q * Dp;
          for (int j = lane; j < Dp; j += 32) {
            const double a = __ldcg(wp + j), b = __ldcg(wq + j);
            __stcg(wp + j, c * a - s * b);
            __stcg(wq + j, s * a + c * b);
            const double x = __ldcg(vp + j), y = __ldcg(vq + j);
            __stcg(vp + j, c * x - s * y);
            __stcg(vq + j, s * x + c * y);
          }
        }
      }
      if (t == m - 1 && lane == 0)
        atomicMax(&ctl->offmax[sweep & 1], (unsigned long long)__double_as_longlong(local_max));
      grid_barrier(ctl, gridDim.x, epoch);
    }
    // all blocks read the same value after the barrier; reset the other slot for the next sweep
    const double mx = __longlong_as_double((long long)*((volatile unsigned long long *)&ctl->offmax[sweep & 1]));
    if (blockIdx.x == 0 && threadIdx.x == 0) ctl->offmax[(sweep + 1) & 1] = 0ull;
    grid_barrier(ctl, gridDim.x, epoch);
    if (mx <= tol) { ++sweep; break; }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    ctl->sweeps_done = sweep;
  }
  // eigenvalues: lambda_i = w_i . v_i
  for (int i = gw; i < Dp; i += nwarps) {
    const double *wi = W + (int64_t)i * Dp, *vi = V + (int64_t)i * Dp;
    double s = 0.0;
    for (int j = lane; j < Dp; j += 32) s = fma(__ldcg(wi + j), __ldcg(vi + j), s);
    s = warp_sum(s);
    if (lane == 0) lam_tmp[i] = s;
  }
  grid_barrier(ctl, gridDim.x, epoch);
  // rank by counting (descending, ties by index), emit the first D, with sklearn's sign rule
  for (int i = gw; i < Dp; i += nwarps) {
    const double li = __ldcg(lam_tmp + i);
    int rank = 0;
    for (int j = lane; j < Dp; j += 32) {
      const double lj = __ldcg(lam_tmp + j);
      rank += (lj > li) || (lj == li && j < i);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) rank += __shfl_xor_sync(FULL, rank, o);
    if (rank >= D) continue;  // padding rows sort last
    const double *vi = V + (int64_t)i * Dp;
    // argmax |v| (first occurrence), only over the real D columns
    double best = -1.0;
    int bidx = 0;
    for (int j = lane; j < D; j += 32) {
      const double a = fabs(__ldcg(vi + j));
      if (a > best) { best = a; bidx = j; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double ob = __shfl_xor_sync(FULL, best, o);
      const int oi = __shfl_xor_sync(FULL, bidx, o);
      if (ob > best || (ob == best && oi < bidx)) { best = ob; bidx = oi; }
    }
    const double sgn = __ldcg(vi + bidx) < 0.0 ? -1.0 : 1.0;
    if (lane == 0) evals[rank] = li;
    for (int j = lane; j < D; j += 32) evecs[(int64_t)rank * D + j] = sgn * __ldcg(vi + j);
  }
}

inline int padded_dim(int D) { return (D + 1) & ~1; }

}  // namespace

extern "C" int64_t am_eigh_work_bytes(int D) {
  const int64_t Dp = padded_dim(D);
  return (2 * Dp * Dp + Dp + 8) * (int64_t)sizeof(double) + 256;
}

extern "C" int am_eigh(const double *d_A, int D, double *d_evals, double *d_evecs, void *d_work,
                       int max_sweeps, double tol, void *stream) {
  if (D < 2 || D > 4096) return am::fail(AM_EINVAL, "am_eigh: D must be in 2..4096");
  if (!d_A || !d_evals || !d_evecs || !d_work) return am::fail(AM_EINVAL, "am_eigh: null pointer");
  if (max_sweeps <= 0) max_sweeps = 30;
  if (tol <= 0.0) tol = 1e-14;
  cudaStream_t st = (cudaStream_t)stream;
  const int Dp = padded_dim(D);
  double *W = (double *)d_work;
  double *V = W + (int64_t)Dp * Dp;
  double *lam = V + (int64_t)Dp * Dp;
  Ctl *ctl = (Ctl *)(lam + Dp + 8);
  // pad diagonal: below every real eigenvalue; the padding row never mixes (off-diagonals are 0)
  // -> use -(1 + sum |a_ii|) computed on device would need a sync; a fixed huge negative is safe.
  const double pad_diag = -1.0e150;
  eigh_init_kernel<<<am::num_sms(), 256, 0, st>>>(d_A, D, Dp, W, V, ctl, pad_diag);
  AM_LAUNCHED("eigh_init_kernel");
  int npairs = Dp / 2;
  int grid = (npairs + EW - 1) / EW;
  int occ = 0;
  AM_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, jacobi_kernel, EW * 32, 0));
  const int max_grid = occ * am::num_sms();
  if (grid > max_grid) grid = max_grid;
  void *args[] = {(void *)&W, (void *)&V, (void *)&Dp, (void *)&D, (void *)&max_sweeps, (void *)&tol,
                  (void *)&ctl, (void *)&d_evals, (void *)&d_evecs, (void *)&lam};
  AM_CUDA(cudaLaunchCooperativeKernel((void *)jacobi_kernel, dim3(grid), dim3(EW * 32), args, 0, st));
  am::g_launches.fetch_add(1);
  return AM_OK;
}
