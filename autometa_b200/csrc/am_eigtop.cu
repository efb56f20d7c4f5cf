// Top-P eigenpairs of the symmetric D x D covariance (D = 512 for k = 5) — sm_100a.
//
// Replaces scipy.linalg.eigh inside sklearn PCA._fit_full(covariance_eigh)
// (sklearn/decomposition/_pca.py:611-624) for the P = pca_dimensions leading pairs that
// kmers.embed keeps (autometa/common/kmers.py:601-605), with svd_flip(u_based_decision=False)
// (sklearn/utils/extmath.py:973-981) applied to the rows.  Same route as LAPACK's dsyevx:
//
//   1. tridiag_cluster_kernel  Householder tridiagonalisation Q^T A Q = T.  ONE thread-block
//      cluster of 16 CTAs holds the whole matrix in distributed shared memory (row i lives in CTA
//      i mod 16), so the 510 dependent steps never touch L2/HBM: per step every CTA does a fused
//      "apply previous rank-2 update + mat-vec with the new reflector" pass over its rows, the
//      p = tau*A*v slices are all-gathered with DSMEM stores, and ONE cluster barrier closes the
//      step; the O(D) reflector algebra is replicated in every CTA (bit-identical by construction).
//   2. bisect_kernel           P leading eigenvalues of T by multisection on the Sturm count
//      (one CTA per eigenvalue, 512 shifts per round).
//   3. invit_kernel            eigenvectors of T by inverse iteration (LU with partial pivoting).
//   4. cholqr_kernel           re-orthonormalisation (CholeskyQR2) — only matters for clusters.
//   5. backtransform_kernel    z = Q y (reflectors applied in reverse), sign rule, output.
#include <cooperative_groups.h>
#include <cuda.h>

#include "am_common.cuh"

namespace cg = cooperative_groups;

namespace {

constexpr unsigned FULL = 0xffffffffu;
constexpr int TC = 16;        // CTAs in the cluster
constexpr int TT = 512;       // threads per CTA
constexpr int TW = TT / 32;   // warps per CTA
constexpr int MAXRPW = 3;     // local rows per warp (RL <= 48)
constexpr int AW = 4;         // warps that run the replicated O(D) reflector algebra (one per SM sub-partition:
                              // with all 16 warps doing it redundantly the fp64 chains contend for issue slots)
constexpr int AT = AW * 32;
constexpr int NE = (576 + AT - 1) / AT;  // vector elements per algebra thread (rows tid + AT*m)
constexpr int SMEM_N_MAX = 576;
constexpr int PMAX = 128;

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
  return v;
}

// 1/q and 1/sqrt(q) to full double precision from the hardware approximations + Newton steps (q finite,
// positive for rsqrt, normal): a third of the latency of the IEEE sqrt / division sequences, which sit on
// the critical path of every Householder step
__device__ __forceinline__ double fast_rcp(double q) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(q));
  double e = fma(-q, r, 1.0);
  r = fma(r, e, r);
  e = fma(-q, r, 1.0);
  r = fma(r, e, r);
  return r;
}
__device__ __forceinline__ double fast_rsqrt(double q) {
  double r;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(q));
  // r <- r + r (1 - q r^2) / 2, three times (the approximation carries ~20 bits)
#pragma unroll
  for (int it = 0; it < 3; ++it) {
    const double h = 0.5 * r;
    const double e = fma(-q * r, r, 1.0);
    r = fma(h, e, r);
  }
  return r;
}

// all threads get the same, order-deterministic total; `red` holds >= 32 doubles
template <int NW>
__device__ __forceinline__ double block_sum(double v, double *red) {
  v = warp_sum(v);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  double s = (threadIdx.x & 31) < NW ? red[threadIdx.x & 31] : 0.0;
  s = warp_sum(s);
  __syncthreads();
  return s;
}

__device__ __forceinline__ uint32_t mapa_u32(uint32_t saddr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
  return r;
}
// 8-byte remote store that also counts 8 bytes on the receiver's mbarrier (data-carrying sync)
__device__ __forceinline__ void st_async_f64(uint32_t raddr, double v, uint32_t rbar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b64 [%0], %1, [%2];"
               ::"r"(raddr), "l"(__double_as_longlong(v)), "r"(rbar) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done) : "r"(bar), "r"(parity) : "memory");
  } while (!done);
}

struct TriOut {
  double *V;     // [n][np] reflectors (v_k in row k; zeros up to k, 1 at k+1)
  double *tau;   // [n]
  double *d;     // [n]
  double *e;     // [n]  (e[i] couples i and i+1)
  long long *dbg;  // optional [8] phase cycle counters (CTA 0, thread 0); may be NULL
  // optional "the cluster is resident" signal: CTA 0 stores started_value here as its first action, so that
  // another stream can hold its next (SM-filling) kernel back until this cluster HAS its 16 SMs
  // (am_stream_wait_geq); may be NULL
  unsigned int *started;
  unsigned int started_value;
};

__device__ __forceinline__ void signal_started(const TriOut &out) {
  if (out.started && blockIdx.x == 0 && threadIdx.x == 0) {
    *reinterpret_cast<volatile unsigned int *>(out.started) = out.started_value;
    __threadfence_system();
  }
}

// ------------------------------------------------------------------ 1. tridiagonalisation
// Step k (LAPACK dsytd2 with the rank-2 update deferred by one step).  x_k = current column k below
// the diagonal, alpha = x_k[k+1], sigma = |x_k[k+2:]|^2, beta = -sign(alpha) sqrt(alpha^2 + sigma),
// tau = (beta - alpha)/beta, v_k = (x_k - beta e_{k+1})/(alpha - beta).
//   B(k)  all warps: fused pass over the CTA's rows >= k — apply the pending update
//         A -= v_{k-1} w_{k-1}^T + w_{k-1} v_{k-1}^T and accumulate y = A x_k with the UNSCALED column
//         (A v_k = (y - beta A e_{k+1}) / (alpha - beta), and A e_{k+1} is the column that is gathered
//         anyway), so the mat-vec does not wait for sigma / sqrt / divisions.  Every row sends
//         {y_i, A[i][k+1]} to all 16 CTAs with one 16-byte st.async that also counts bytes on the
//         receiver's mbarrier (data-carrying synchronisation: no cluster barrier in the loop).
//   S(k)  while the inbox fills: sigma from the per-warp partials, beta, tau, 1/(alpha - beta);
//         the owner CTA writes v_k, e_k, tau_k for the back-transformation.
//   C(k)  p = tau (y - beta a)/(alpha - beta), gamma = p.v (one block reduction), w_k = p - tau gamma/2 v,
//         next column x_{k+1} = a - v w[k+1] - w, per-warp partials of its sigma.
// The O(D) algebra is replicated in every CTA from identical inputs (bit-identical results).
__device__ __forceinline__ void st_async_f64x2(uint32_t raddr, double a, double b, uint32_t rbar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.b64 [%0], {%1, %2}, [%3];"
               ::"r"(raddr), "l"(__double_as_longlong(a)), "l"(__double_as_longlong(b)), "r"(rbar) : "memory");
}

__global__ void __cluster_dims__(TC, 1, 1) __launch_bounds__(TT, 1)
tridiag_cluster_kernel(const double *__restrict__ A, int n, int np, int RL, TriOut out) {
  signal_started(out);
  cg::cluster_group cluster = cg::this_cluster();
  const int c = (int)cluster.block_rank();
  extern __shared__ double sm[];
  double *rows = sm;                          // RL x np
  double *vb = rows + (size_t)RL * np;        // 2 x np   scaled reflectors v_{k-1}, v_k
  double *wb = vb + 2 * np;                   // 2 x np   w_{k-1}, w_k
  double *xs = wb + 2 * np;                   // 2 x np   unscaled columns x_k (mat-vec operand)
  double2 *inbox = reinterpret_cast<double2 *>(xs + 2 * np);  // 2 x np  {y_i, A[i][k+1]}
  double *red = reinterpret_cast<double *>(inbox + 2 * np);   // 4 x TW rotating reduction slots
  double *sigp = red + 4 * TW;                // 2 x TW   per-warp partials of sigma
  __shared__ __align__(8) unsigned long long s_mbar[2];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t mbar0 = (uint32_t)__cvta_generic_to_shared(&s_mbar[0]);
  const uint32_t inbox_s = (uint32_t)__cvta_generic_to_shared(inbox);
  const int i0 = tid, i1 = tid + TT;  // the rows this thread owns in the replicated vector algebra

  for (int lr = 0; lr < RL; ++lr) {
    const int i = c + TC * lr;
    for (int j = tid; j < np; j += TT) {
      double a = 0.0;
      if (i < n && j < n) a = 0.5 * (__ldg(A + (size_t)i * n + j) + __ldg(A + (size_t)j * n + i));
      rows[(size_t)lr * np + j] = a;
    }
  }
  for (int j = tid; j < 2 * np; j += TT) { vb[j] = 0.0; wb[j] = 0.0; xs[j] = 0.0; inbox[j] = make_double2(0.0, 0.0); }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mbar0) : "memory");
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mbar0 + 8u) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  cluster.sync();  // every CTA's buffers and mbarriers exist before any remote store

  // prologue: all-gather column 0 (inbox slot 1, .y component), form x_0
  if (tid == 0)
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar0 + 8u), "r"((uint32_t)(n - 1) * 16u) : "memory");
  for (int lr = warp; lr < RL; lr += TW) {
    const int i = c + TC * lr;
    if (i < n && i >= 1 && lane < TC)
      st_async_f64x2(mapa_u32(inbox_s + (uint32_t)(np + i) * 16u, lane), 0.0, rows[(size_t)lr * np + 0],
                     mapa_u32(mbar0 + 8u, lane));
  }
  mbar_wait_cluster(mbar0 + 8u, 0u);
  double x[NE];
#pragma unroll
  for (int m = 0; m < NE; ++m) x[m] = 0.0;
  if (warp < AW) {
    double sp = 0.0;
#pragma unroll
    for (int m = 0; m < NE; ++m) {
      const int i = tid + AT * m;
      if (i >= 1 && i < n) x[m] = inbox[np + i].y;
      if (i < np) xs[i] = x[m];
      if (i >= 2) sp = fma(x[m], x[m], sp);
    }
    sp = warp_sum(sp);
    if (lane == 0) sigp[warp] = sp;
  }
  __syncthreads();

  int rslot = 0;
  long long tA = 0, tB = 0, tS = 0, tC = 0, t0 = clock64(), t1;
#define AM_TICK(acc) do { t1 = clock64(); acc += t1 - t0; t0 = t1; } while (0)
  for (int k = 0; k <= n - 3; ++k) {
    const int cur = k & 1, prv = (k + 1) & 1;
    double *vcur = vb + cur * np, *wcur = wb + cur * np;
    const double *vprev = vb + prv * np, *wprev = wb + prv * np;
    const double *xk = xs + cur * np;
    // arm this step's inbox: rows k+1..n-1 each deliver 16 bytes
    if (tid == 0)
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;"
                   ::"r"(mbar0 + 8u * cur), "r"((uint32_t)(n - k - 1) * 16u) : "memory");
    // ---- B. fused pass over own rows >= k: pending update (k-1), y = A x_k
    {
      int lrs[MAXRPW];
      double vpi[MAXRPW], wpi[MAXRPW], acc[MAXRPW];
      int nr = 0;
#pragma unroll
      for (int r = 0; r < MAXRPW; ++r) {
        const int lr = warp + TW * r;
        const int i = c + TC * lr;
        lrs[r] = -1; vpi[r] = 0.0; wpi[r] = 0.0; acc[r] = 0.0;
        if (lr < RL && i < n && i >= k) { lrs[r] = lr; vpi[r] = vprev[i]; wpi[r] = wprev[i]; nr = r + 1; }
      }
      if (nr > 0) {
#pragma unroll 2
        for (int j = (k & ~31) + lane; j < np; j += 32) {
          const double vpj = vprev[j], wpj = wprev[j], xj = xk[j];
#pragma unroll
          for (int r = 0; r < MAXRPW; ++r) {
            if (lrs[r] >= 0) {
              double *ap = rows + (size_t)lrs[r] * np + j;
              double a = *ap;
              a = fma(-vpi[r], wpj, a);
              a = fma(-wpi[r], vpj, a);
              *ap = a;
              acc[r] = fma(a, xj, acc[r]);
            }
          }
        }
        __syncwarp();
        const uint32_t rbar = mapa_u32(mbar0 + 8u * cur, (uint32_t)(lane & (TC - 1)));
#pragma unroll
        for (int r = 0; r < MAXRPW; ++r) {
          if (lrs[r] >= 0) {
            const int i = c + TC * lrs[r];
            const double y = warp_sum(acc[r]);
            const double a1 = rows[(size_t)lrs[r] * np + (k + 1)];  // post-update A[i][k+1]
            if (i >= k + 1 && lane < TC)
              st_async_f64x2(mapa_u32(inbox_s + (uint32_t)(cur * np + i) * 16u, lane), y, a1, rbar);
          }
        }
      }
    }
    AM_TICK(tB);
    // ---- S. reflector scalars while the inbox fills (algebra warps only)
    double beta = 0.0, tau = 0.0, scale = 0.0;
    double v[NE];
    if (warp < AW) {
      double sigma = 0.0;
#pragma unroll
      for (int w = 0; w < AW; ++w) sigma += sigp[cur * AW + w];
      const double alpha = xk[k + 1];
      if (!(sigma > 1e-290)) {
        beta = alpha; tau = 0.0; scale = 0.0;
      } else {
        beta = -copysign(sqrt(fma(alpha, alpha, sigma)), alpha);
        tau = (beta - alpha) / beta;
        scale = 1.0 / (alpha - beta);
      }
#pragma unroll
      for (int m = 0; m < NE; ++m) {
        const int i = tid + AT * m;
        v[m] = (i == k + 1) ? 1.0 : ((i > k + 1 && i < n) ? x[m] * scale : 0.0);
        if (c == (k % TC) && i < np) out.V[(size_t)k * np + i] = v[m];
      }
      if (c == (k % TC) && tid == 0) { out.e[k] = beta; out.tau[k] = tau; }
    }
    AM_TICK(tA);
    // slot 0 is used by steps 0, 2, ..; slot 1 by the prologue and steps 1, 3, ..
    if (warp < AW) mbar_wait_cluster(mbar0 + 8u * cur, (uint32_t)((cur ? ((k + 1) >> 1) : (k >> 1)) & 1));
    AM_TICK(tS);
    // ---- C. w_k and the next column (algebra warps; the other warps wait at the barriers)
    const double2 *ib = inbox + cur * np;
    const double ts = tau * scale;
    double pv[NE], av[NE];
    double *r = red + (rslot & 3) * TW;
    ++rslot;
    if (warp < AW) {
      double part = 0.0;
#pragma unroll
      for (int m = 0; m < NE; ++m) {
        const int i = tid + AT * m;
        const bool in = (i >= k + 1 && i < n);
        const double2 g = in ? ib[i] : make_double2(0.0, 0.0);
        pv[m] = in ? ts * fma(-beta, g.y, g.x) : 0.0;
        av[m] = g.y;
        part = fma(pv[m], v[m], part);
      }
      part = warp_sum(part);
      if (lane == 0) r[warp] = part;
    }
    __syncthreads();
    if (warp < AW) {
      double gam = 0.0;
#pragma unroll
      for (int w = 0; w < AW; ++w) gam += r[w];
      const double Kc = 0.5 * tau * gam;
      const double2 gk = ib[k + 1];
      const double wk1 = ts * fma(-beta, gk.y, gk.x) - Kc;  // w_k[k+1]  (v_k[k+1] = 1)
      double *xn = xs + prv * np;
      double sp = 0.0;
#pragma unroll
      for (int m = 0; m < NE; ++m) {
        const int i = tid + AT * m;
        const bool in = (i >= k + 1 && i < n);
        const double w = in ? fma(-Kc, v[m], pv[m]) : 0.0;
        if (i < np) { vcur[i] = v[m]; wcur[i] = w; }
        x[m] = (i >= k + 2 && i < n) ? av[m] - v[m] * wk1 - w : 0.0;
        if (i < np) xn[i] = x[m];
        if (i >= k + 3) sp = fma(x[m], x[m], sp);
      }
      sp = warp_sum(sp);
      if (lane == 0) sigp[prv * AW + warp] = sp;
    }
    __syncthreads();
    AM_TICK(tC);
  }
  if (out.dbg && c == 0 && tid == 0) { out.dbg[0] = tA; out.dbg[1] = tB; out.dbg[2] = tS; out.dbg[3] = tC; }

  // epilogue: pending update (n-3) on rows/cols >= n-2, then d and the last e
  {
    const int k = n - 2;
    const double *vprev = vb + ((k + 1) & 1) * np, *wprev = wb + ((k + 1) & 1) * np;
    for (int lr = warp; lr < RL; lr += TW) {
      const int i = c + TC * lr;
      if (i < n && i >= k) {
        const double vi = vprev[i], wi = wprev[i];
        for (int j = (k & ~31) + lane; j < np; j += 32) {
          double *ap = rows + (size_t)lr * np + j;
          double a = *ap;
          a = fma(-vi, wprev[j], a);
          a = fma(-wi, vprev[j], a);
          *ap = a;
        }
      }
    }
    __syncthreads();
    for (int lr = tid; lr < RL; lr += TT) {
      const int i = c + TC * lr;
      if (i < n) {
        out.d[i] = rows[(size_t)lr * np + i];
        if (i == n - 1) { out.e[n - 2] = rows[(size_t)lr * np + (n - 2)]; out.e[n - 1] = 0.0; out.tau[n - 2] = 0.0; out.tau[n - 1] = 0.0; }
      }
    }
  }
  cluster.sync();  // no CTA exits while a peer may still address its shared memory
}

// Register-resident variant for np <= 512 (the k = 5 table): the smem pass over the trailing matrix (16 B
// of shared-memory traffic per element and step) was the longest phase of a step; with 8 warps x 4 rows
// x 16 columns per lane the CTA's 32 x 512 block lives in 128 registers per thread and only the three
// operand vectors are read from shared memory.
constexpr int RT = 256;       // threads per CTA
constexpr int RW = RT / 32;   // warps
constexpr int RR = 4;         // local rows per warp (RL <= 32)
constexpr int NB = 16;        // 32-column blocks per row (np <= 512)
constexpr int JG = 4;         // column blocks per branch in the update pass
constexpr int RAW = RW;       // all 8 warps run the replicated reflector algebra: 2 vector entries per thread
constexpr int RAT = RAW * 32;
constexpr int RNE = (32 * NB + RAT - 1) / RAT;

// PROBE = true: per-phase cycle counters (scripts/dev_eig_once.py, AM_TRIDIAG_PROBE=1); off in production
// EXACT = true: n = np = 512 (the k = 5 table) are compile-time constants, which removes the bounds tests
// (a seventh of the loop body's instructions) from the 510-step loop.
template <bool PROBE, bool EXACT>
__global__ void __cluster_dims__(TC, 1, 1) __launch_bounds__(RT, 1)
tridiag_reg_kernel(const double *__restrict__ A, int n_, int np_, int RL_, TriOut out) {
  signal_started(out);
  const int n = EXACT ? 32 * NB : n_, np = EXACT ? 32 * NB : np_, RL = EXACT ? RR * RW : RL_;
  cg::cluster_group cluster = cg::this_cluster();
  const int c = (int)cluster.block_rank();
  extern __shared__ double sm[];
  double *rows = sm;                          // RL x np
  double *vb = rows + (size_t)RL * np;        // 2 x np   scaled reflectors v_{k-1}, v_k
  double *wb = vb + 2 * np;                   // 2 x np   w_{k-1}, w_k
  double *xs = wb + 2 * np;                   // 2 x np   unscaled columns x_k (mat-vec operand)
  double2 *inbox = reinterpret_cast<double2 *>(xs + 2 * np);  // 2 x np  {y_i, A[i][k+1]}
  double *red = reinterpret_cast<double *>(inbox + 2 * np);   // 4 x TW rotating reduction slots
  double *sigp = red + 4 * TW;                // 2 x TW   per-warp partials of sigma (layout shared with the other kernel)
  __shared__ __align__(8) unsigned long long s_mbar[2];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t mbar0 = (uint32_t)__cvta_generic_to_shared(&s_mbar[0]);
  const uint32_t inbox_s = (uint32_t)__cvta_generic_to_shared(inbox);
  
  for (int lr = 0; lr < RL; ++lr) {
    const int i = c + TC * lr;
    for (int j = tid; j < np; j += RT) {
      double a = 0.0;
      if (i < n && j < n) a = 0.5 * (__ldg(A + (size_t)i * n + j) + __ldg(A + (size_t)j * n + i));
      rows[(size_t)lr * np + j] = a;
    }
  }
  for (int j = tid; j < 2 * np; j += RT) { vb[j] = 0.0; wb[j] = 0.0; xs[j] = 0.0; inbox[j] = make_double2(0.0, 0.0); }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mbar0) : "memory");
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mbar0 + 8u) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  cluster.sync();  // every CTA's buffers and mbarriers exist before any remote store

  // prologue: all-gather column 0 (inbox slot 1, .y component), form x_0
  if (tid == 0)
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar0 + 8u), "r"((uint32_t)(n - 1) * 16u) : "memory");
  for (int lr = warp; lr < RL; lr += RW) {
    const int i = c + TC * lr;
    if (i < n && i >= 1 && lane < TC)
      st_async_f64x2(mapa_u32(inbox_s + (uint32_t)(np + i) * 16u, lane), 0.0, rows[(size_t)lr * np + 0],
                     mapa_u32(mbar0 + 8u, lane));
  }
  mbar_wait_cluster(mbar0 + 8u, 0u);
  double x[RNE];
#pragma unroll
  for (int m = 0; m < RNE; ++m) x[m] = 0.0;
  if (warp < RAW) {
    double sp = 0.0;
#pragma unroll
    for (int m = 0; m < RNE; ++m) {
      const int i = tid + RAT * m;
      if (i >= 1 && i < n) x[m] = inbox[np + i].y;
      if (i < np) xs[i] = x[m];
      if (i >= 2) sp = fma(x[m], x[m], sp);
    }
    sp = warp_sum(sp);
    if (lane == 0) sigp[warp] = sp;
  }
  __syncthreads();

  // the CTA's rows move into registers: warp w owns local rows w + 8 r, lane l the columns l + 32 jj
  double am[RR][NB];
#pragma unroll
  for (int r = 0; r < RR; ++r) {
    const int lr = warp + RW * r;
#pragma unroll
    for (int jj = 0; jj < NB; ++jj) {
      const int j = 32 * jj + lane;
      am[r][jj] = (lr < RL && j < np) ? rows[(size_t)lr * np + j] : 0.0;
    }
  }
  int rslot = 0;
  long long tA = 0, tB = 0, tS = 0, tC = 0, tB1 = 0, tB2 = 0, t0 = PROBE ? clock64() : 0, t1 = 0;
long long tprev = 0;
#define AM_TICK(acc) do { if (PROBE) { t1 = clock64(); acc += t1 - t0; tprev = t0; t0 = t1; } } while (0)
  for (int k = 0; k <= n - 3; ++k) {
    const int cur = k & 1, prv = (k + 1) & 1;
    double *vcur = vb + cur * np, *wcur = wb + cur * np;
    const double *vprev = vb + prv * np, *wprev = wb + prv * np;
    const double *xk = xs + cur * np;
    // arm this step's inbox: rows k+1..n-1 each deliver 16 bytes
    if (tid == 0)
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;"
                   ::"r"(mbar0 + 8u * cur), "r"((uint32_t)(n - k - 1) * 16u) : "memory");
    // ---- B. fused pass over own rows >= k, matrix in registers: pending update (k-1), y = A x_k
    {
      const int kb = k >> 5, jb1 = (k + 1) >> 5, l1 = (k + 1) & 31;
      bool act[RR];
      double vpi[RR], wpi[RR], acc[RR], a1v[RR];
      bool any = false;
#pragma unroll
      for (int r = 0; r < RR; ++r) {
        const int lr = warp + RW * r;
        const int i = c + TC * lr;
        act[r] = lr < RL && i < n && i >= k;
        vpi[r] = act[r] ? vprev[i] : 0.0;
        wpi[r] = act[r] ? wprev[i] : 0.0;
        acc[r] = 0.0;
        a1v[r] = 0.0;
        any |= act[r];
      }
      if (any) {
        // column blocks in groups of JG: one (warp-uniform) branch per group, so that the loads and the
        // 3-deep fp64 chains of JG x RR elements are scheduled together (a branch per block serialises
        // load latency + chain latency 16 times per step)
#pragma unroll
        for (int g = 0; g < NB / JG; ++g) {
          if (JG * g + JG - 1 >= kb && 32 * JG * g < np) {
            double vpj[JG], wpj[JG], xj[JG];
#pragma unroll
            for (int q = 0; q < JG; ++q) {
              const int j = 32 * (JG * g + q) + lane;
              const bool ok = j < np;
              vpj[q] = ok ? vprev[j] : 0.0;
              wpj[q] = ok ? wprev[j] : 0.0;
              xj[q] = ok ? xk[j] : 0.0;
            }
#pragma unroll
            for (int q = 0; q < JG; ++q) {
              const int jj = JG * g + q;
#pragma unroll
              for (int r = 0; r < RR; ++r) {
                if (act[r]) {
                  double a = am[r][jj];
                  a = fma(-vpi[r], wpj[q], a);
                  a = fma(-wpi[r], vpj[q], a);
                  am[r][jj] = a;
                  acc[r] = fma(a, xj[q], acc[r]);
                }
              }
            }
          }
        }
        // post-update A[i][k+1] lives in column block jb1 of lane l1 (picked up outside the update code:
        // a test per block inside it splits the block into branches)
#define AM_PICK(JJ) case JJ: a1v[0] = am[0][JJ]; a1v[1] = am[1][JJ]; a1v[2] = am[2][JJ]; a1v[3] = am[3][JJ]; break;
        switch (jb1) {
          AM_PICK(0) AM_PICK(1) AM_PICK(2) AM_PICK(3) AM_PICK(4) AM_PICK(5) AM_PICK(6) AM_PICK(7)
          AM_PICK(8) AM_PICK(9) AM_PICK(10) AM_PICK(11) AM_PICK(12) AM_PICK(13) AM_PICK(14) AM_PICK(15)
          default: break;
        }
#undef AM_PICK
        AM_TICK(tB1);
        if (PROBE && out.dbg && c == 0 && tid == 0 && (k & 63) == 0 && (k >> 6) < 8) out.dbg[8 + (k >> 6)] = t1 - tprev;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
          for (int r = 0; r < RR; ++r) acc[r] += __shfl_xor_sync(FULL, acc[r], o);
        }
        AM_TICK(tB2);
        const uint32_t rbar = mapa_u32(mbar0 + 8u * cur, (uint32_t)(lane & (TC - 1)));
#pragma unroll
        for (int r = 0; r < RR; ++r) {
          const double a1 = __shfl_sync(FULL, a1v[r], l1);  // post-update A[i][k+1]
          const int i = c + TC * (warp + RW * r);
          if (act[r] && i >= k + 1 && lane < TC)
            st_async_f64x2(mapa_u32(inbox_s + (uint32_t)(cur * np + i) * 16u, lane), acc[r], a1, rbar);
        }
      }
    }
    AM_TICK(tB);
    // ---- S. reflector scalars while the inbox fills (algebra warps only)
    double beta = 0.0, tau = 0.0, scale = 0.0;
    double v[RNE];
    if (warp < RAW) {
      double sigma = 0.0;
#pragma unroll
      for (int w = 0; w < RAW; ++w) sigma += sigp[cur * RAW + w];
      const double alpha = xk[k + 1];
      if (!(sigma > 1e-290)) {
        beta = alpha; tau = 0.0; scale = 0.0;
      } else {
        // |x| = q / sqrt(q) with one correction; tau = (beta - alpha) / beta = 1 + |alpha| / |x|;
        // 1 / (alpha - beta) = sign(alpha) / (|alpha| + |x|)
        const double q = fma(alpha, alpha, sigma);
        const double rs = fast_rsqrt(q);
        double nrm = q * rs;
        nrm = fma(0.5 * rs, fma(-nrm, nrm, q), nrm);
        beta = -copysign(nrm, alpha);
        tau = fma(fabs(alpha), rs, 1.0);
        scale = copysign(fast_rcp(fabs(alpha) + nrm), alpha);
      }
#pragma unroll
      for (int m = 0; m < RNE; ++m) {
        const int i = tid + RAT * m;
        v[m] = (i == k + 1) ? 1.0 : ((i > k + 1 && i < n) ? x[m] * scale : 0.0);
        if (c == (k % TC) && i < np) out.V[(size_t)k * np + i] = v[m];
      }
      if (c == (k % TC) && tid == 0) { out.e[k] = beta; out.tau[k] = tau; }
    }
    AM_TICK(tA);
    if (PROBE && out.dbg && c == 0 && tid == 0 && (k & 63) == 0 && (k >> 6) < 8) out.dbg[24 + (k >> 6)] = t1 - tprev;
    // slot 0 is used by steps 0, 2, ..; slot 1 by the prologue and steps 1, 3, ..
    if (warp < RAW) mbar_wait_cluster(mbar0 + 8u * cur, (uint32_t)((cur ? ((k + 1) >> 1) : (k >> 1)) & 1));
    AM_TICK(tS);
    if (PROBE && out.dbg && c == 0 && tid == 0 && (k & 63) == 0 && (k >> 6) < 8) out.dbg[16 + (k >> 6)] = t1 - tprev;
    // ---- C. w_k and the next column (algebra warps; the other warps wait at the barriers)
    const double2 *ib = inbox + cur * np;
    const double ts = tau * scale;
    double pv[RNE], av[RNE];
    double *r = red + (rslot & 3) * TW;
    ++rslot;
    if (warp < RAW) {
      double part = 0.0;
#pragma unroll
      for (int m = 0; m < RNE; ++m) {
        const int i = tid + RAT * m;
        const bool in = (i >= k + 1 && i < n);
        const double2 g = in ? ib[i] : make_double2(0.0, 0.0);
        pv[m] = in ? ts * fma(-beta, g.y, g.x) : 0.0;
        av[m] = g.y;
        part = fma(pv[m], v[m], part);
      }
      part = warp_sum(part);
      if (lane == 0) r[warp] = part;
    }
    __syncthreads();
    if (warp < RAW) {
      double gam = 0.0;
#pragma unroll
      for (int w = 0; w < RAW; ++w) gam += r[w];
      const double Kc = 0.5 * tau * gam;
      const double2 gk = ib[k + 1];
      const double wk1 = ts * fma(-beta, gk.y, gk.x) - Kc;  // w_k[k+1]  (v_k[k+1] = 1)
      double *xn = xs + prv * np;
      double sp = 0.0;
#pragma unroll
      for (int m = 0; m < RNE; ++m) {
        const int i = tid + RAT * m;
        const bool in = (i >= k + 1 && i < n);
        const double w = in ? fma(-Kc, v[m], pv[m]) : 0.0;
        if (i < np) { vcur[i] = v[m]; wcur[i] = w; }
        x[m] = (i >= k + 2 && i < n) ? av[m] - v[m] * wk1 - w : 0.0;
        if (i < np) xn[i] = x[m];
        if (i >= k + 3) sp = fma(x[m], x[m], sp);
      }
      sp = warp_sum(sp);
      if (lane == 0) sigp[prv * RAW + warp] = sp;
    }
    __syncthreads();
    AM_TICK(tC);
  }
  if (PROBE && out.dbg && c == 0 && tid == 0) { out.dbg[0] = tA; out.dbg[1] = tB; out.dbg[2] = tS; out.dbg[3] = tC; out.dbg[4] = tB1; out.dbg[5] = tB2; }

  // epilogue: pending update (n-3) on rows/cols >= n-2 (in registers), back to shared memory, then d and the last e
  {
    const int k = n - 2, kb = k >> 5;
    const double *vprev = vb + ((k + 1) & 1) * np, *wprev = wb + ((k + 1) & 1) * np;
#pragma unroll
    for (int r = 0; r < RR; ++r) {
      const int lr = warp + RW * r;
      const int i = c + TC * lr;
      if (lr < RL && i < n) {
        const bool upd = i >= k;
        const double vi = upd ? vprev[i] : 0.0, wi = upd ? wprev[i] : 0.0;
#pragma unroll
        for (int jj = 0; jj < NB; ++jj) {
          const int j = 32 * jj + lane;
          if (j < np) {
            double a = am[r][jj];
            if (upd && jj >= kb) {
              a = fma(-vi, wprev[j], a);
              a = fma(-wi, vprev[j], a);
            }
            rows[(size_t)lr * np + j] = a;
          }
        }
      }
    }
    __syncthreads();
    for (int lr = tid; lr < RL; lr += RT) {
      const int i = c + TC * lr;
      if (i < n) {
        out.d[i] = rows[(size_t)lr * np + i];
        if (i == n - 1) { out.e[n - 2] = rows[(size_t)lr * np + (n - 2)]; out.e[n - 1] = 0.0; out.tau[n - 2] = 0.0; out.tau[n - 1] = 0.0; }
      }
    }
  }
  cluster.sync();  // no CTA exits while a peer may still address its shared memory
}

// ------------------------------------------------------------------ 2. bisection
constexpr int BT = 256;  // shifts per round (8 bits of the bracket per round; the Sturm recurrence is bound by
                         // fp64 throughput at 512 shifts and by the dependent FMA chain at 256)

// Number of eigenvalues of T that are < x, from the sign changes of the leading principal minors
// u_j = (d_j - x) u_{j-1} - e_{j-1}^2 u_{j-2}  (Sturm sequence).  One FMA on the dependent chain per
// row instead of a division; both running minors are rescaled by a power of two every 4 rows.  An
// exactly zero minor takes the sign opposite to its predecessor (LAPACK dlaebz's pivmin rule).
__device__ __forceinline__ int sturm_count(const double2 *__restrict__ de, int n, double x) {
  // de[i] = {d_i, e_{i-1}^2}: one 16-byte broadcast load per row (the kernel is otherwise bound by
  // load-issue slots: 16 warps x 2 loads per row).  Only the FMA is on the dependent chain; signs are
  // read off the sign bits (an exactly zero minor has sign bit 0: a measure-zero event that moves the
  // bracket by at most one ulp), and both running minors are rescaled by a power of two every 16 rows
  // (growth per row <= 3 after the |T| <= 1 scaling, so 16 rows stay far inside the double range).
  double um = 1.0, u = de[0].x - x;
  int cnt = (__double2hiint(u) >> 31) & 1;
  for (int i0 = 1; i0 < n; i0 += 16) {
    const int i1 = min(n, i0 + 16);
    for (int i = i0; i < i1; ++i) {
      const double2 v = de[i];
      const double t = v.y * um;
      const double un = fma(v.x - x, u, -t);
      cnt += ((__double2hiint(un) ^ __double2hiint(u)) >> 31) & 1;
      um = u;
      u = un;
    }
    const int ex = max((__double2hiint(u) >> 20) & 0x7ff, (__double2hiint(um) >> 20) & 0x7ff);
    if (ex > 0 && ex < 2046) {
      const double f = __hiloint2double((2046 - ex) << 20, 0);  // 2^(1023 - ex)
      u *= f;
      um *= f;
    } else if (ex == 0) {
      u = (__double2hiint(u) < 0) ? -1.0 : 1.0;  // both minors vanished: restart from the sign only
      um = 0.0;
    }
  }
  return cnt;
}

__global__ void __launch_bounds__(BT)
bisect_kernel(const double *__restrict__ dg, const double *__restrict__ eg, int n, int P,
              double *__restrict__ lam, double *__restrict__ trace) {
  extern __shared__ double sm[];
  double *d = sm, *e2 = sm + n, *red = sm + 2 * n;  // red: 64
  __shared__ int first_warp[BT / 32];
  __shared__ double s_lo, s_hi;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int jtop = blockIdx.x;  // jtop-th largest
  double gl = 1e300, gu = -1e300, emax = 0.0, tr = 0.0;
  for (int i = tid; i < n; i += BT) {
    const double di = dg[i], el = i > 0 ? fabs(eg[i - 1]) : 0.0, er = i < n - 1 ? fabs(eg[i]) : 0.0;
    d[i] = di;
    e2[i] = i < n - 1 ? eg[i] * eg[i] : 0.0;
    gl = fmin(gl, di - el - er);
    gu = fmax(gu, di + el + er);
    emax = fmax(emax, er * er);
    tr += di;
  }
  // block min / max / sum
  for (int o = 16; o > 0; o >>= 1) {
    gl = fmin(gl, __shfl_xor_sync(FULL, gl, o));
    gu = fmax(gu, __shfl_xor_sync(FULL, gu, o));
    emax = fmax(emax, __shfl_xor_sync(FULL, emax, o));
  }
  if (lane == 0) { red[warp] = gl; red[16 + warp] = gu; red[32 + warp] = emax; }
  __syncthreads();
  gl = red[0]; gu = red[16]; emax = red[32];
  for (int w = 1; w < BT / 32; ++w) { gl = fmin(gl, red[w]); gu = fmax(gu, red[16 + w]); emax = fmax(emax, red[32 + w]); }
  __syncthreads();
  tr = block_sum<BT / 32>(tr, red);
  if (blockIdx.x == 0 && tid == 0 && trace) *trace = tr;
  const double tnorm = fmax(fabs(gl), fabs(gu));
  const double ulp = 2.220446049250313e-16;
  // work on T / 2^ex (|entries| <= 1): the minors then grow by at most 3x per row
  int ex = 0;
  if (tnorm > 0.0) frexp(tnorm, &ex);
  const double sdn = ldexp(1.0, -ex), sup = ldexp(1.0, ex);
  double2 *de = reinterpret_cast<double2 *>(sm + 2 * n + 64);  // packed {d_i, e_{i-1}^2}, scaled
  for (int i = tid; i < n; i += BT) de[i] = make_double2(d[i] * sdn, i > 0 ? e2[i - 1] * sdn * sdn : 0.0);
  __syncthreads();
  gl *= sdn; gu *= sdn;
  const double pivmin = 1e-290;
  double lo = gl - 2.0 * ulp * n - 2.0 * pivmin;
  double hi = gu + 2.0 * ulp * n + 2.0 * pivmin;
  const int target = n - jtop;  // want smallest x with count(x) >= target  (lambda index n-1-jtop ascending)
  for (int round = 0; round < 12; ++round) {
    const double width = hi - lo;
    if (!(width > 2.0 * ulp * fmax(fabs(lo), fabs(hi)) + 2.0 * pivmin)) break;
    const double xt = lo + width * ((double)(tid + 1) / (double)(BT + 1));
    const int cnt = sturm_count(de, n, xt);
    const unsigned ball = __ballot_sync(FULL, cnt >= target);
    if (lane == 0) first_warp[warp] = ball ? (warp * 32 + __ffs(ball) - 1) : -1;
    __syncthreads();
    if (tid == 0) {
      int first = -1;
      for (int w = 0; w < BT / 32; ++w)
        if (first_warp[w] >= 0) { first = first_warp[w]; break; }
      // points are lo + width*(t+1)/(BT+1)
      double nlo, nhi;
      if (first < 0) { nlo = lo + width * ((double)BT / (double)(BT + 1)); nhi = hi; }
      else { nhi = lo + width * ((double)(first + 1) / (double)(BT + 1)); nlo = first == 0 ? lo : lo + width * ((double)first / (double)(BT + 1)); }
      s_lo = nlo; s_hi = nhi;
    }
    __syncthreads();
    lo = s_lo; hi = s_hi;
    __syncthreads();
  }
  if (tid == 0) lam[jtop] = 0.5 * (lo + hi) * sup;
}


// ------------------------------------------------------------------ 3. inverse iteration
__global__ void __launch_bounds__(32)
invit_kernel(const double *__restrict__ dg, const double *__restrict__ eg, int n, int np, int P,
             const double *__restrict__ lam, double *__restrict__ Y) {
  extern __shared__ double sm[];
  double *dd = sm, *dl = sm + np, *du = sm + 2 * np, *du2 = sm + 3 * np, *rdd = sm + 4 * np, *b = sm + 5 * np;
  int *piv = (int *)(sm + 6 * np);
  const int lane = threadIdx.x;
  const int j = blockIdx.x;
  // shift: separate numerically coincident eigenvalues like dstein does
  double tnorm = 0.0;
  for (int i = lane; i < n; i += 32) tnorm = fmax(tnorm, fabs(dg[i]) + (i < n - 1 ? fabs(eg[i]) : 0.0) + (i > 0 ? fabs(eg[i - 1]) : 0.0));
  for (int o = 16; o > 0; o >>= 1) tnorm = fmax(tnorm, __shfl_xor_sync(FULL, tnorm, o));
  const double ulp = 2.220446049250313e-16;
  double shift = lam[j];
  {
    // count how many earlier (larger) eigenvalues sit within 10 ulp*|T| of this one
    int close = 0;
    for (int q = j - 1; q >= 0; --q) {
      if (fabs(lam[q] - lam[j]) <= 10.0 * ulp * tnorm) ++close; else break;
    }
    shift -= (double)close * 10.0 * ulp * tnorm;
  }
  const double tinyp = fmax(ulp * tnorm * 0.25, 1e-300);
  for (int i = lane; i < n; i += 32) {
    dd[i] = dg[i] - shift;
    du[i] = i < n - 1 ? eg[i] : 0.0;
    dl[i] = i < n - 1 ? eg[i] : 0.0;
    du2[i] = 0.0;
    // deterministic pseudo-random start in [-1, 1)
    unsigned h = (unsigned)i * 2654435761u ^ ((unsigned)j + 1u) * 40503u;
    h ^= h >> 15; h *= 2246822519u; h ^= h >> 13; h *= 3266489917u; h ^= h >> 16;
    b[i] = (double)(h >> 8) * (1.0 / 8388608.0) - 1.0;
  }
  __syncwarp();
  if (lane == 0) {
    // LU with partial pivoting of the tridiagonal (LAPACK dgttrf); the running diagonal / super-
    // diagonal entries are carried in registers so that only the division is on the dependent chain
    double dcur = dd[0], ducur = du[0];
    for (int i = 0; i < n - 1; ++i) {
      const double dli = dl[i], dnext = dd[i + 1], dunext = (i < n - 2) ? du[i + 1] : 0.0;
      if (fabs(dcur) >= fabs(dli)) {
        piv[i] = 0;
        if (fabs(dcur) < tinyp) dcur = copysign(tinyp, dcur);
        const double fact = dli * fast_rcp(dcur);
        dd[i] = dcur;
        du[i] = ducur;
        dl[i] = fact;
        du2[i] = 0.0;
        dcur = fma(-fact, ducur, dnext);
        ducur = dunext;
      } else {
        piv[i] = 1;
        const double fact = dcur * fast_rcp(dli);
        dd[i] = dli;
        dl[i] = fact;
        du[i] = dnext;
        du2[i] = dunext;
        dcur = fma(-fact, dnext, ducur);
        ducur = -fact * dunext;
      }
    }
    if (fabs(dcur) < tinyp) dcur = copysign(tinyp, dcur);
    dd[n - 1] = dcur;
  }
  __syncwarp();
  // back-substitution coefficients pre-multiplied by 1/d_i so that one FMA per row is on the chain:
  // x_i = b_i/d_i - (du_i/d_i) x_{i+1} - (du2_i/d_i) x_{i+2}
  for (int i = lane; i < n; i += 32) {
    const double r = 1.0 / dd[i];
    rdd[i] = r;
    du[i] *= r;
    du2[i] *= r;
  }
  __syncwarp();
  // dstein-style: start from a vector of norm ~ eps |T|, stop once the solve has grown it to O(1)
  // (the shift is an eigenvalue to working precision, so one or two solves suffice)
  for (int it = 0; it < 4; ++it) {
    if (lane == 0) {
      // forward substitution (dgtts2), running entry in a register
      double cur = b[0];
      for (int i = 0; i < n - 1; ++i) {
        const double nxt = b[i + 1], l = dl[i];
        if (!piv[i]) {
          b[i] = cur;
          cur = fma(-l, cur, nxt);
        } else {
          b[i] = nxt;
          cur = fma(-l, nxt, cur);
        }
      }
      // back substitution, two running entries in registers
      double x1 = cur * rdd[n - 1], x2 = 0.0;
      b[n - 1] = x1;
      for (int i = n - 2; i >= 0; --i) {
        const double t = fma(-du2[i], x2, b[i] * rdd[i]);  // x2 is one step old: off the chain
        const double xi = fma(-du[i], x1, t);
        b[i] = xi;
        x2 = x1;
        x1 = xi;
      }
    }
    __syncwarp();
    double mx = 0.0;
    for (int i = lane; i < n; i += 32) mx = fmax(mx, fabs(b[i]));
    for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(FULL, mx, o));
    const double inv = mx > 0.0 ? 1.0 / mx : 1.0;
    for (int i = lane; i < n; i += 32) b[i] *= inv;
    __syncwarp();
    // growth of a unit-max vector by ~1/(eps |T|) means the iterate is the eigenvector; the second
    // solve polishes it
    if (it >= 1 && mx * tnorm * ulp * (double)n >= 1e-3) break;
  }
  double ss = 0.0;
  for (int i = lane; i < n; i += 32) ss += b[i] * b[i];
  ss = warp_sum(ss);
  const double inv = 1.0 / sqrt(ss);
  for (int i = lane; i < np; i += 32) Y[(size_t)j * np + i] = i < n ? b[i] * inv : 0.0;
}

// ------------------------------------------------------------------ 4. CholeskyQR2 on the P rows of Y
constexpr int CT = 1024;
__device__ void orthonormalise_rows(double *Y, double *G, int n, int np, int P, bool have_m) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  __shared__ double s_emax[CT / 32];
  for (int pass = 0; pass < 2; ++pass) {
    const int npairs = P * (P + 1) / 2;
    for (int pr = warp; pr < npairs; pr += CT / 32) {
      // pr -> (a >= b)
      int a = (int)((sqrt(8.0 * pr + 1.0) - 1.0) * 0.5);
      while (a * (a + 1) / 2 > pr) --a;
      while ((a + 1) * (a + 2) / 2 <= pr) ++a;
      const int bb = pr - a * (a + 1) / 2;
      double s = 0.0;
      for (int i = lane; i < n; i += 32) s = fma(Y[(size_t)a * np + i], Y[(size_t)bb * np + i], s);
      s = warp_sum(s);
      if (lane == 0) { G[a * P + bb] = s; G[bb * P + a] = s; }
    }
    __syncthreads();
    // departure from orthonormality E = G - I
    double em = 0.0;
    for (int q = tid; q < P * P; q += CT) em = fmax(em, fabs(G[q] - ((q / P == q % P) ? 1.0 : 0.0)));
    for (int o = 16; o > 0; o >>= 1) em = fmax(em, __shfl_xor_sync(FULL, em, o));
    if (lane == 0) s_emax[warp] = em;
    __syncthreads();
    em = 0.0;
    for (int w = 0; w < CT / 32; ++w) em = fmax(em, s_emax[w]);
    __syncthreads();
    if (em <= 4e-16) return;  // already orthonormal to working precision
    if (em <= 1e-5 && have_m) {
      // well separated eigenvalues: symmetric (Loewdin) orthogonalisation, second order in E:
      // Y <- (I - E/2 + 3 E^2 / 8) Y, error O(E^3) <= 1e-15.  No sequential dependence.
      double *M = G + P * P;  // second P x P buffer
      for (int q = tid; q < P * P; q += CT) {
        const int a = q / P, b = q % P;
        double e2 = 0.0;
        for (int m = 0; m < P; ++m) {
          const double ea = G[a * P + m] - (a == m ? 1.0 : 0.0), eb = G[m * P + b] - (m == b ? 1.0 : 0.0);
          e2 = fma(ea, eb, e2);
        }
        const double e1 = G[q] - (a == b ? 1.0 : 0.0);
        M[q] = (a == b ? 1.0 : 0.0) - 0.5 * e1 + 0.375 * e2;
      }
      __syncthreads();
      // Y <- M Y with register accumulators: one work unit = 16 rows a of one element column i
      constexpr int AG = 16;
      const int ngroups = (P + AG - 1) / AG;
      const int nwork = ngroups * n;
      const int rounds = (nwork + CT - 1) / CT;
      for (int rd = 0; rd < rounds; ++rd) {
        const int work = rd * CT + tid;
        const int g = work / n, i = work % n;
        double acc[AG];
#pragma unroll
        for (int aa = 0; aa < AG; ++aa) acc[aa] = 0.0;
        if (work < nwork) {
          for (int b = 0; b < P; ++b) {
            const double y = Y[(size_t)b * np + i];
#pragma unroll
            for (int aa = 0; aa < AG; ++aa) {
              const int a = g * AG + aa;
              if (a < P) acc[aa] = fma(M[a * P + b], y, acc[aa]);
            }
          }
        }
        // rows of column i are read by the other groups of the same column in this round only
        __syncthreads();
        if (work < nwork) {
#pragma unroll
          for (int aa = 0; aa < AG; ++aa) {
            const int a = g * AG + aa;
            if (a < P) Y[(size_t)a * np + i] = acc[aa];
          }
        }
        __syncthreads();
      }
      return;
    }
    if (warp == 0) {
      for (int cidx = 0; cidx < P; ++cidx) {
        double dsum = G[cidx * P + cidx];
        for (int m = 0; m < cidx; ++m) dsum -= G[cidx * P + m] * G[cidx * P + m];
        const double lcc = sqrt(fmax(dsum, 1e-300));
        __syncwarp();
        for (int a = cidx + 1 + lane; a < P; a += 32) {
          double s = G[a * P + cidx];
          for (int m = 0; m < cidx; ++m) s -= G[a * P + m] * G[cidx * P + m];
          G[a * P + cidx] = s / lcc;
        }
        if (lane == 0) G[cidx * P + cidx] = lcc;
        __syncwarp();
      }
    }
    __syncthreads();
    // rows: q_a = (y_a - sum_{b<a} L[a][b] q_b) / L[a][a], independently per element i
    for (int i = tid; i < n; i += CT) {
      double q[PMAX];
      for (int a = 0; a < P; ++a) {
        double s = Y[(size_t)a * np + i];
        for (int bq = 0; bq < a; ++bq) s = fma(-G[a * P + bq], q[bq], s);
        q[a] = s / G[a * P + a];
      }
      for (int a = 0; a < P; ++a) Y[(size_t)a * np + i] = q[a];
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(CT)
cholqr_kernel(double *__restrict__ Yg, int n, int np, int P, int y_in_smem, int have_m,
              const int *__restrict__ mode) {
  if (mode && mode[0] != 2) return;  // the multi-CTA fast path already handled (or had nothing to do)
  extern __shared__ double sm[];
  double *G = sm;                 // P x P (+ a second P x P buffer when have_m)
  double *Ys = sm + (have_m ? 2 : 1) * P * P;    // P x np when it fits (the 50 x 512 case: 200 KB)
  if (y_in_smem) {
    for (int q = threadIdx.x; q < P * np; q += CT) Ys[q] = Yg[q];
    __syncthreads();
    orthonormalise_rows(Ys, G, n, np, P, have_m != 0);
    __syncthreads();
    for (int q = threadIdx.x; q < P * np; q += CT) Yg[q] = Ys[q];
  } else {
    orthonormalise_rows(Yg, G, n, np, P, have_m != 0);
  }
}

// ---- multi-CTA fast path of the re-orthonormalisation (the common case: eigenvalues separated
// well enough that the inverse-iteration vectors are orthogonal to 1e-5).  Three small launches
// spread over many SMs instead of one CTA: Gram of the P vectors, M = I - E/2 + 3E^2/8 (second-order
// Loewdin, error O(E^3)), Y <- M Y.  mode[0]: 0 = already orthonormal, 1 = M is valid, 2 = clustered
// eigenvalues -> the single-CTA CholeskyQR2 kernel takes over.
__global__ void __launch_bounds__(128)
ortho_gram_kernel(const double *__restrict__ Y, int n, int np, int P, double *__restrict__ G) {
  const int a = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int b = warp; b <= a; b += 4) {
    double s0 = 0.0, s1 = 0.0;
    for (int i = lane; i < n; i += 64) {
      s0 = fma(__ldg(Y + (size_t)a * np + i), __ldg(Y + (size_t)b * np + i), s0);
      if (i + 32 < n) s1 = fma(__ldg(Y + (size_t)a * np + i + 32), __ldg(Y + (size_t)b * np + i + 32), s1);
    }
    const double s = warp_sum(s0 + s1);
    if (lane == 0) { G[a * P + b] = s; G[b * P + a] = s; }
  }
}

__global__ void __launch_bounds__(1024)
ortho_m_kernel(const double *__restrict__ G, int P, double *__restrict__ M, int *__restrict__ mode) {
  extern __shared__ double sm[];
  double *E = sm;  // P x P
  __shared__ double s_em[32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  double em = 0.0;
  for (int q = tid; q < P * P; q += 1024) {
    const double e = G[q] - ((q / P == q % P) ? 1.0 : 0.0);
    E[q] = e;
    em = fmax(em, fabs(e));
  }
  for (int o = 16; o > 0; o >>= 1) em = fmax(em, __shfl_xor_sync(FULL, em, o));
  if (lane == 0) s_em[warp] = em;
  __syncthreads();
  em = 0.0;
  for (int w = 0; w < 32; ++w) em = fmax(em, s_em[w]);
  const int md = !(em > 4e-16) ? 0 : (em <= 1e-5 ? 1 : 2);
  if (tid == 0) mode[0] = md;
  if (md != 1) return;
  for (int q = tid; q < P * P; q += 1024) {
    const int a = q / P, b = q % P;
    double e2 = 0.0;
    for (int m = 0; m < P; ++m) e2 = fma(E[a * P + m], E[m * P + b], e2);
    M[q] = (a == b ? 1.0 : 0.0) - 0.5 * E[q] + 0.375 * e2;
  }
}

constexpr int OA_COLS = 64;  // element columns per CTA
__global__ void __launch_bounds__(256)
ortho_apply_kernel(double *__restrict__ Y, int n, int np, int P, const double *__restrict__ M,
                   const int *__restrict__ mode) {
  if (mode[0] != 1) return;
  extern __shared__ double sm[];
  double *Ms = sm;                 // P x P
  double *Yt = sm + P * P;         // P x OA_COLS tile
  const int i0 = blockIdx.x * OA_COLS, tid = threadIdx.x;
  for (int q = tid; q < P * P; q += 256) Ms[q] = M[q];
  for (int q = tid; q < P * OA_COLS; q += 256) {
    const int b = q / OA_COLS, i = i0 + q % OA_COLS;
    Yt[q] = (i < n) ? Y[(size_t)b * np + i] : 0.0;
  }
  __syncthreads();
  // thread = one column i and a strided set of rows a
  const int ci = tid % OA_COLS, a0 = tid / OA_COLS;  // 4 row groups
  for (int a = a0; a < P; a += 256 / OA_COLS) {
    double acc0 = 0.0, acc1 = 0.0;
    for (int b = 0; b + 1 < P; b += 2) {
      acc0 = fma(Ms[a * P + b], Yt[b * OA_COLS + ci], acc0);
      acc1 = fma(Ms[a * P + b + 1], Yt[(b + 1) * OA_COLS + ci], acc1);
    }
    if (P & 1) acc0 = fma(Ms[a * P + P - 1], Yt[(P - 1) * OA_COLS + ci], acc0);
    if (i0 + ci < n) Y[(size_t)a * np + i0 + ci] = acc0 + acc1;
  }
}

// ------------------------------------------------------------------ 5. back-transformation + sign + output
// One CTA of 4 warps per eigenvector, z in registers (thread owns rows tid + 128 m), one barrier per
// round, and TWO reflectors per reduction round: for H_b H_a z with a = k, b = k-1,
//   alpha = v_a.z,  beta = v_b.(z - tau_a alpha v_a) = v_b.z - tau_a alpha (v_b.v_a),
// so the three dot products v_a.z, v_b.z, v_b.v_a are reduced together (one shuffle tree + one
// shared-memory exchange), then
// z -= tau_a alpha v_a + tau_b beta v_b.  The reflector pair of the next round is prefetched.
constexpr int KT = 128;              // one CTA (4 warps) per eigenvector
constexpr int ZL = (SMEM_N_MAX + KT - 1) / KT;  // rows per thread
__global__ void __launch_bounds__(KT)
backtransform_kernel(const double *__restrict__ V, const double *__restrict__ tau, int n, int np, int P,
                     const double *__restrict__ Y, const double *__restrict__ lam,
                     double *__restrict__ evals, double *__restrict__ evecs) {
  __shared__ double red[4][3][KT / 32];  // rotating slots x {v_a.z, v_b.z, v_b.v_a} x warps
  __shared__ double s_best[KT / 32], s_val[KT / 32];
  __shared__ int s_idx[KT / 32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int j = blockIdx.x;
  double z[ZL], va[ZL], vb[ZL], na[ZL], nb[ZL];
#pragma unroll
  for (int m = 0; m < ZL; ++m) {
    const int i = tid + KT * m;
    z[m] = (i < np) ? Y[(size_t)j * np + i] : 0.0;
    na[m] = 0.0; nb[m] = 0.0;
  }
  auto load_pair = [&](int k, double (&xa)[ZL], double (&xb)[ZL]) {
#pragma unroll
    for (int m = 0; m < ZL; ++m) {
      const int i = tid + KT * m;
      xa[m] = (i < np && k >= 0) ? __ldg(V + (size_t)k * np + i) : 0.0;
      xb[m] = (i < np && k >= 1) ? __ldg(V + (size_t)(k - 1) * np + i) : 0.0;
    }
  };
  int k = n - 3, slot = 0;
  if (k >= 0) load_pair(k, na, nb);
  for (; k >= 0; k -= 2) {
#pragma unroll
    for (int m = 0; m < ZL; ++m) { va[m] = na[m]; vb[m] = nb[m]; }
    if (k - 2 >= 0) load_pair(k - 2, na, nb);
    const double ta = __ldg(tau + k), tb = (k >= 1) ? __ldg(tau + k - 1) : 0.0;
    double s1 = 0.0, s2 = 0.0, s3 = 0.0;
#pragma unroll
    for (int m = 0; m < ZL; ++m) {
      s1 = fma(va[m], z[m], s1);
      s2 = fma(vb[m], z[m], s2);
      s3 = fma(vb[m], va[m], s3);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      s1 += __shfl_xor_sync(FULL, s1, o);
      s2 += __shfl_xor_sync(FULL, s2, o);
      s3 += __shfl_xor_sync(FULL, s3, o);
    }
    double(*r)[KT / 32] = red[slot & 3];
    ++slot;
    if (lane == 0) { r[0][warp] = s1; r[1][warp] = s2; r[2][warp] = s3; }
    __syncthreads();
    s1 = (r[0][0] + r[0][1]) + (r[0][2] + r[0][3]);
    s2 = (r[1][0] + r[1][1]) + (r[1][2] + r[1][3]);
    s3 = (r[2][0] + r[2][1]) + (r[2][2] + r[2][3]);
    const double ca = ta * s1;
    const double cb = tb * fma(-ca, s3, s2);
#pragma unroll
    for (int m = 0; m < ZL; ++m) z[m] = fma(-cb, vb[m], fma(-ca, va[m], z[m]));
  }
  // sklearn svd_flip(u_based_decision=False): largest-|.| entry (first occurrence) made positive
  double best = -1.0, bval = 0.0;
  int bidx = 0;
#pragma unroll
  for (int m = 0; m < ZL; ++m) {
    const int i = tid + KT * m;
    if (i < n) {
      const double a = fabs(z[m]);
      if (a > best) { best = a; bidx = i; bval = z[m]; }
    }
  }
  for (int o = 16; o > 0; o >>= 1) {
    const double ob = __shfl_xor_sync(FULL, best, o);
    const int oi = __shfl_xor_sync(FULL, bidx, o);
    const double ov = __shfl_xor_sync(FULL, bval, o);
    if (ob > best || (ob == best && oi < bidx)) { best = ob; bidx = oi; bval = ov; }
  }
  if (lane == 0) { s_best[warp] = best; s_idx[warp] = bidx; s_val[warp] = bval; }
  __syncthreads();
  best = s_best[0]; bidx = s_idx[0]; bval = s_val[0];
  for (int w = 1; w < KT / 32; ++w)
    if (s_best[w] > best || (s_best[w] == best && s_idx[w] < bidx)) { best = s_best[w]; bidx = s_idx[w]; bval = s_val[w]; }
  const double sgn = bval < 0.0 ? -1.0 : 1.0;
#pragma unroll
  for (int m = 0; m < ZL; ++m) {
    const int i = tid + KT * m;
    if (i < n) evecs[(size_t)j * n + i] = sgn * z[m];
  }
  if (tid == 0) evals[j] = lam[j];
}

// ------------------------------------------------------------------ 5b. blocked back-transformation
// z = H_0 H_1 ... H_{n-3} y with the reflectors grouped WB at a time in compact WY form (LAPACK dlarft,
// forward / columnwise):  H_{k0} ... H_{k0+m-1} = I - V T V^T,  T upper triangular,
//   T[j][j] = tau_j,  T[0:j, j] = -tau_j T[0:j, 0:j] (V[:, 0:j]^T v_j).
// Applying a block costs 4 barriers instead of one per pair of reflectors (255 -> 128 barriers with a
// quarter of the dependent reductions per barrier); the T factors are shared by all P eigenvectors.
constexpr int WB = 16;    // reflectors per block
constexpr int WT = 256;   // threads
constexpr int WZ = (SMEM_N_MAX + WT - 1) / WT;  // vector entries per thread

__global__ void __launch_bounds__(WT)
wy_factor_kernel(const double *__restrict__ V, const double *__restrict__ tau, int n, int np,
                 double *__restrict__ Tf) {
  extern __shared__ double sm[];
  const int pitch = np + 1;  // rows of the tile start in different banks
  double *Vt = sm;                 // WB x pitch
  double *G = Vt + WB * pitch;     // WB x (WB + 1)
  double *T = G + WB * (WB + 1);   // WB x (WB + 1)
  const int tid = threadIdx.x, lane = tid & 31;
  const int k0 = blockIdx.x * WB;
  const int m = min(WB, (n - 2) - k0);  // reflectors k0 .. k0+m-1 (k <= n-3)
  for (int q = tid; q < WB * np; q += WT) {
    const int r = q / np, i = q - r * np;
    Vt[r * pitch + i] = (r < m) ? __ldg(V + (size_t)(k0 + r) * np + i) : 0.0;
  }
  for (int q = tid; q < WB * (WB + 1); q += WT) T[q] = 0.0;
  __syncthreads();
  {
    const int a = tid & (WB - 1), c = tid >> 4;  // WT = WB * WB pairs
    double s = 0.0;
    if (a <= c)
      for (int i = k0; i < np; ++i) s = fma(Vt[a * pitch + i], Vt[c * pitch + i], s);  // v_k is zero up to k
    G[a * (WB + 1) + c] = s;
  }
  __syncthreads();
  if (tid < 32) {
    const int a = lane;
    if (a < m) T[a * (WB + 1) + a] = __ldg(tau + k0 + a);
    __syncwarp();
    for (int j = 1; j < m; ++j) {
      const double tj = __ldg(tau + k0 + j);
      const double tmp = (a < j) ? -tj * G[a * (WB + 1) + j] : 0.0;
      double acc = 0.0;
      for (int c = 0; c < j; ++c) {
        const double tc = __shfl_sync(FULL, tmp, c);
        if (a < j && c >= a) acc = fma(T[a * (WB + 1) + c], tc, acc);
      }
      if (a < j) T[a * (WB + 1) + j] = acc;
      __syncwarp();
    }
  }
  __syncthreads();
  for (int q = tid; q < WB * WB; q += WT) Tf[(size_t)blockIdx.x * WB * WB + q] = T[(q / WB) * (WB + 1) + (q % WB)];
}

__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}

__global__ void __launch_bounds__(WT)
backtransform_wy_kernel(const double *__restrict__ V, const double *__restrict__ Tf, int n, int np, int P,
                        const double *__restrict__ Y, const double *__restrict__ lam,
                        double *__restrict__ evals, double *__restrict__ evecs) {
  extern __shared__ double sm[];
  double *Vt = sm;                    // 2 x WB x np   (double-buffered reflector tiles)
  double *Tt = Vt + 2 * WB * np;      // 2 x WB x WB   (their T factors)
  double *zs = Tt + 2 * WB * WB;      // np
  double *us = zs + np;               // WB
  double *ws = us + WB;               // WB
  __shared__ double s_best[WT / 32], s_val[WT / 32];
  __shared__ int s_idx[WT / 32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int j = blockIdx.x;
  double z[WZ];
#pragma unroll
  for (int q = 0; q < WZ; ++q) {
    const int i = tid + WT * q;
    z[q] = (i < np) ? Y[(size_t)j * np + i] : 0.0;
    if (i < np) zs[i] = z[q];
  }
  const int nblk = (n - 2 + WB - 1) / WB;
  const int per_row = np / 2;
  // tile b (rows k0 .. k0+WB-1 of V; rows past n-3 are zero) and T_b -> buffer b & 1, asynchronously
  auto prefetch = [&](int b) {
    const int k0 = b * WB;
    const int m = min(WB, (n - 2) - k0);
    const double2 *src = reinterpret_cast<const double2 *>(V + (size_t)k0 * np);
    double2 *dst = reinterpret_cast<double2 *>(Vt + (size_t)(b & 1) * WB * np);
    const int q_valid = m * per_row;  // chunks of the rows that exist
    for (int q = tid; q < WB * per_row; q += WT) {
      if (q < q_valid) cp_async16(dst + q, src + q);
      else dst[q] = make_double2(0.0, 0.0);
    }
    if (tid < WB * WB / 2)
      cp_async16(reinterpret_cast<double2 *>(Tt + (b & 1) * WB * WB) + tid,
                 reinterpret_cast<const double2 *>(Tf + (size_t)b * WB * WB) + tid);
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  prefetch(nblk - 1);
  for (int b = nblk - 1; b >= 0; --b) {
    if (b > 0) {
      prefetch(b - 1);
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();
    const double *Vb = Vt + (size_t)(b & 1) * WB * np;
    // u = V_b^T z: warp w takes reflectors w and w + 8
    {
      double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
      const double *v0 = Vb + warp * np, *v1 = Vb + (warp + 8) * np;
      for (int i = lane; i < np; i += 64) {
        const bool two = i + 32 < np;  // np is a multiple of 32, not necessarily of 64
        const double za = zs[i], zb = two ? zs[i + 32] : 0.0;
        s0 = fma(v0[i], za, s0);
        s1 = fma(v1[i], za, s1);
        s2 = fma(two ? v0[i + 32] : 0.0, zb, s2);
        s3 = fma(two ? v1[i + 32] : 0.0, zb, s3);
      }
      s0 += s2;
      s1 += s3;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        s0 += __shfl_xor_sync(FULL, s0, o);
        s1 += __shfl_xor_sync(FULL, s1, o);
      }
      if (lane == 0) { us[warp] = s0; us[warp + 8] = s1; }
    }
    __syncthreads();
    // w = T_b u
    if (tid < WB) {
      const double *Tb = Tt + (b & 1) * WB * WB + tid * WB;
      double a0 = 0.0, a1 = 0.0;
      for (int c = tid; c < WB; c += 2) {
        a0 = fma(Tb[c], us[c], a0);
        if (c + 1 < WB) a1 = fma(Tb[c + 1], us[c + 1], a1);
      }
      ws[tid] = a0 + a1;
    }
    __syncthreads();
    // z -= V_b w
#pragma unroll
    for (int q = 0; q < WZ; ++q) {
      const int i = tid + WT * q;
      if (i < np) {
        double a0 = 0.0, a1 = 0.0;
#pragma unroll
        for (int r = 0; r < WB; r += 2) {
          a0 = fma(Vb[r * np + i], ws[r], a0);
          a1 = fma(Vb[(r + 1) * np + i], ws[r + 1], a1);
        }
        z[q] -= a0 + a1;
        zs[i] = z[q];
      }
    }
    __syncthreads();
  }
  // sklearn svd_flip(u_based_decision=False): largest-|.| entry (first occurrence) made positive
  double best = -1.0, bval = 0.0;
  int bidx = 0;
#pragma unroll
  for (int q = 0; q < WZ; ++q) {
    const int i = tid + WT * q;
    if (i < n) {
      const double a = fabs(z[q]);
      if (a > best) { best = a; bidx = i; bval = z[q]; }
    }
  }
  for (int o = 16; o > 0; o >>= 1) {
    const double ob = __shfl_xor_sync(FULL, best, o);
    const int oi = __shfl_xor_sync(FULL, bidx, o);
    const double ov = __shfl_xor_sync(FULL, bval, o);
    if (ob > best || (ob == best && oi < bidx)) { best = ob; bidx = oi; bval = ov; }
  }
  if (lane == 0) { s_best[warp] = best; s_idx[warp] = bidx; s_val[warp] = bval; }
  __syncthreads();
  best = s_best[0]; bidx = s_idx[0]; bval = s_val[0];
  for (int w = 1; w < WT / 32; ++w)
    if (s_best[w] > best || (s_best[w] == best && s_idx[w] < bidx)) { best = s_best[w]; bidx = s_idx[w]; bval = s_val[w]; }
  const double sgn = bval < 0.0 ? -1.0 : 1.0;
#pragma unroll
  for (int q = 0; q < WZ; ++q) {
    const int i = tid + WT * q;
    if (i < n) evecs[(size_t)j * n + i] = sgn * z[q];
  }
  if (tid == 0) evals[j] = lam[j];
}

inline int round_up32(int x) { return (x + 31) & ~31; }

struct TopWork {
  double *V, *tau, *d, *e, *lam, *Y, *dbg, *orth, *Tf;
};

int64_t carve_top(void *base, int n, int P, TopWork *w) {
  const int np = round_up32(n);
  int64_t off = 0;
  char *p = (char *)base;
  auto take = [&](int64_t bytes) {
    char *q = p ? p + off : nullptr;
    off += (bytes + 255) & ~(int64_t)255;
    return (double *)q;
  };
  TopWork t;
  t.V = take((int64_t)n * np * 8);
  t.tau = take((int64_t)np * 8);
  t.d = take((int64_t)np * 8);
  t.e = take((int64_t)np * 8);
  t.lam = take((int64_t)PMAX * 8);
  t.Y = take((int64_t)P * np * 8);
  t.dbg = take(256);
  t.orth = take(((int64_t)2 * P * P + 8) * 8);
  t.Tf = take((int64_t)((n + WB - 1) / WB) * WB * WB * 8);
  if (w) *w = t;
  return off;
}

}  // namespace

extern "C" int64_t am_eigh_topk_work_bytes(int D, int P) {
  if (D < 1 || P < 1) return 0;
  return carve_top(nullptr, D, P, nullptr);
}

extern "C" int am_eigh_topk(const double *d_A, int D, int P, double *d_evals, double *d_evecs,
                            double *d_trace, void *d_work, void *stream) {
  return am_eigh_topk_signal(d_A, D, P, d_evals, d_evecs, d_trace, d_work, nullptr, 0u, stream);
}

extern "C" int am_stream_wait_geq(void *stream, const uint32_t *d_flag, uint32_t value) {
  typedef CUresult (*WaitFn)(CUstream, CUdeviceptr, cuuint32_t, unsigned int);
  static WaitFn fn = nullptr;
  if (!fn) {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuStreamWaitValue32", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = (WaitFn)p;
  }
  if (!fn) return am::fail(AM_ECUDA, "am_stream_wait_geq: cuStreamWaitValue32 entry point unavailable");
  if (!d_flag) return am::fail(AM_EINVAL, "am_stream_wait_geq: null pointer");
  const CUresult r = fn((CUstream)stream, (CUdeviceptr)(uintptr_t)d_flag, value, CU_STREAM_WAIT_VALUE_GEQ);
  if (r != CUDA_SUCCESS) return am::fail(AM_ECUDA, "am_stream_wait_geq: cuStreamWaitValue32 failed");
  return AM_OK;
}

extern "C" int am_eigh_topk_signal(const double *d_A, int D, int P, double *d_evals, double *d_evecs,
                                   double *d_trace, void *d_work, uint32_t *d_started, uint32_t started_value,
                                   void *stream) {
  if (D < 3 || D > SMEM_N_MAX) return am::fail(AM_EINVAL, "am_eigh_topk: D must be in 3..576 (use am_eigh otherwise)");
  if (P < 1 || P > D || P > PMAX) return am::fail(AM_EINVAL, "am_eigh_topk: P must be in 1..min(D,128)");
  if (!d_A || !d_evals || !d_evecs || !d_work) return am::fail(AM_EINVAL, "am_eigh_topk: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  const int n = D, np = round_up32(n);
  const int RL = (n + TC - 1) / TC;
  if (RL > MAXRPW * TW) return am::fail(AM_EINVAL, "am_eigh_topk: D too large for the cluster path");
  TopWork w;
  carve_top(d_work, n, P, &w);
  TriOut out{w.V, w.tau, w.d, w.e, (long long *)w.dbg, d_started, started_value};

  // 1. tridiagonalisation (one cluster of 16 CTAs)
  {
    const size_t smem = ((size_t)RL * np + 10 * (size_t)np + 6 * TW) * sizeof(double);
    if (smem > 227 * 1024) return am::fail(AM_EINVAL, "am_eigh_topk: matrix does not fit the cluster's shared memory");
    const char *tv = getenv("AM_TRIDIAG_SMEM");
    const bool reg = np <= 32 * NB && RL <= RR * RW && !(tv && tv[0] == '1');
    const bool probe = getenv("AM_TRIDIAG_PROBE") != nullptr;
    const bool exact = n == 32 * NB;
    auto kern = !reg ? tridiag_cluster_kernel
                     : probe ? (exact ? tridiag_reg_kernel<true, true> : tridiag_reg_kernel<true, false>)
                             : (exact ? tridiag_reg_kernel<false, true> : tridiag_reg_kernel<false, false>);
    AM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    AM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(TC);
    cfg.blockDim = dim3(reg ? RT : TT);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = TC;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    AM_CUDA(cudaLaunchKernelEx(&cfg, kern, d_A, n, np, RL, out));
    AM_LAUNCHED("tridiag_cluster_kernel");
  }
  // 2. bisection
  {
    const size_t smem = ((size_t)4 * n + 64 + 2) * sizeof(double);
    bisect_kernel<<<P, BT, smem, st>>>(w.d, w.e, n, P, w.lam, d_trace);
    AM_LAUNCHED("bisect_kernel");
  }
  // 3. inverse iteration
  {
    const size_t smem = ((size_t)7 * np) * sizeof(double);
    if (smem > 48 * 1024)
      AM_CUDA(cudaFuncSetAttribute(invit_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    invit_kernel<<<P, 32, smem, st>>>(w.d, w.e, n, np, P, w.lam, w.Y);
    AM_LAUNCHED("invit_kernel");
  }
  // 4. CholeskyQR2
  if (P > 1) {
    const int have_m = (P <= 64) ? 1 : 0;  // second P x P buffer for the Loewdin fast path
    size_t smem = (size_t)(have_m ? 2 : 1) * P * P * sizeof(double);
    int y_in_smem = 0;
    if (smem + (size_t)P * np * sizeof(double) <= 227 * 1024) { smem += (size_t)P * np * sizeof(double); y_in_smem = 1; }
    if (smem > 48 * 1024)
      AM_CUDA(cudaFuncSetAttribute(cholqr_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    // fast path over many SMs; the single-CTA CholeskyQR2 only runs for clustered eigenvalues
    double *Gg = w.orth, *Mg = w.orth + (size_t)P * P;
    int *mode = (int *)(w.orth + 2 * (size_t)P * P);
    ortho_gram_kernel<<<P, 128, 0, st>>>(w.Y, n, np, P, Gg);
    AM_LAUNCHED("ortho_gram_kernel");
    if ((size_t)P * P * sizeof(double) > 48 * 1024)
      AM_CUDA(cudaFuncSetAttribute(ortho_m_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((size_t)P * P * sizeof(double))));
    ortho_m_kernel<<<1, 1024, (size_t)P * P * sizeof(double), st>>>(Gg, P, Mg, mode);
    AM_LAUNCHED("ortho_m_kernel");
    {
      const size_t sm2 = ((size_t)P * P + (size_t)P * OA_COLS) * sizeof(double);
      if (sm2 > 48 * 1024)
        AM_CUDA(cudaFuncSetAttribute(ortho_apply_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm2));
      ortho_apply_kernel<<<(n + OA_COLS - 1) / OA_COLS, 256, sm2, st>>>(w.Y, n, np, P, Mg, mode);
      AM_LAUNCHED("ortho_apply_kernel");
    }
    cholqr_kernel<<<1, CT, smem, st>>>(w.Y, n, np, P, y_in_smem, have_m, mode);
    AM_LAUNCHED("cholqr_kernel");
  }
  // 5. back-transformation
  {
    const char *bs = getenv("AM_BACKTRANSFORM_SEQ");
    if (n < 3 + WB || (bs && bs[0] == '1')) {
      backtransform_kernel<<<P, KT, 0, st>>>(w.V, w.tau, n, np, P, w.Y, w.lam, d_evals, d_evecs);
      AM_LAUNCHED("backtransform_kernel");
    } else {
      const int nblk = (n - 2 + WB - 1) / WB;
      const size_t sm1 = ((size_t)WB * (np + 1) + 2 * WB * (WB + 1)) * sizeof(double);
      const size_t sm2 = ((size_t)2 * WB * np + 2 * WB * WB + np + 64 + 2 * WB) * sizeof(double);
      AM_CUDA(cudaFuncSetAttribute(wy_factor_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm1));
      AM_CUDA(cudaFuncSetAttribute(backtransform_wy_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm2));
      wy_factor_kernel<<<nblk, WT, sm1, st>>>(w.V, w.tau, n, np, w.Tf);
      AM_LAUNCHED("wy_factor_kernel");
      backtransform_wy_kernel<<<P, WT, sm2, st>>>(w.V, w.Tf, n, np, P, w.Y, w.lam, d_evals, d_evecs);
      AM_LAUNCHED("backtransform_wy_kernel");
    }
  }
  return AM_OK;
}
