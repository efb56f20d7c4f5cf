// PCA projection on the tensor cores — scores = (X - mu) @ components^T from the int8 digit planes
// (tcgen05.mma kind::i8, TMEM accumulators, TMA-staged K-major operands) — sm_100a.
//
// Replaces PCA.transform (sklearn/decomposition/_base.py:151-159: X @ components_.T - mean_ @
// components_.T), the second half of fit_transform at autometa/common/kmers.py:605.
//
// The digit planes written by the fused normalise pass (am_normalize_planes) hold
//   q_rj = round((x_rj - c_j) 2^(46 - e_j))  as six balanced base-256 digits a_s[r, j].
// The components are folded with the column scales and quantised the same way,
//   u_pj = round(V_pj 2^e_j 2^(46 - f_p)),  2^f_p > max_j |V_pj 2^e_j|,   digits b_t[p, j],
// so   sum_j (x_rj - c_j) V_pj = 2^(f_p - 12) sum_d 2^(-8d) P_d[r, p],   P_d = sum_{s+t=d} a_s b_t^T
// with every digit product exact in int32 (K = 512 columns).  The exact mean enters only through
// the per-component constant b_p = sum_j (mu_j - c_j) V_pj.  Tile = 128 contigs x 64 components
// (50 used), K pipelined in 64-column TMA stages; same warp roles as the Gram kernel.
#include "am_common.cuh"
#include "am_tc.cuh"

using namespace amtc;

namespace {

constexpr int PJ_M = 128;   // contigs per tile (TMEM lanes)
constexpr int PJ_N = 64;    // components per accumulator (padded)
constexpr int PJ_BK = 64;   // columns (bytes) per stage: two K = 32 MMA steps, 64B swizzle
constexpr int PJ_S = 6;     // digit planes stored
constexpr int PJ_W = 4;     // digit pairs multiplied: s + t <= PJ_W (15 pairs, 5 accumulators).  Scores are
                            // compared at 1e-8 of their range; the dropped pairs weigh <= 2^-40 of the
                            // leading one (measured error 1e-11 .. 1e-10 of max|score|), unlike the Gram,
                            // whose error is amplified by 1/gap in the eigenvectors and keeps every pair.
constexpr int PJ_NP = PJ_W + 1;  // planes actually staged (0 .. PJ_W)
constexpr int PJ_THREADS = 192;
constexpr int PJ_A_BYTES = PJ_M * PJ_BK;  // 8192
constexpr int PJ_B_BYTES = PJ_N * PJ_BK;  // 4096
constexpr int PJ_STAGES = 3;

struct ProjParams {
  int n_rows, D, P, n_tiles, n_ksteps;
  int n_acc, n_ops;
  uint8_t os[16], ot[16], on[16], od[16], ofirst[16];
};

__device__ __forceinline__ void digits6p(double v, int (&a)[6]) {
  const long long q = __double2ll_rn(v);
  int lo = (int)((unsigned)(q + 0x800000LL) & 0xFFFFFFu) - 0x800000;
  int hi = (int)((q - lo) >> 24);
  a[5] = ((lo + 128) & 255) - 128; lo = (lo - a[5]) >> 8;
  a[4] = ((lo + 128) & 255) - 128; lo = (lo - a[4]) >> 8;
  a[3] = ((lo + 128) & 255) - 128; hi += (lo - a[3]) >> 8;
  a[2] = ((hi + 128) & 255) - 128; hi = (hi - a[2]) >> 8;
  a[1] = ((hi + 128) & 255) - 128; hi = (hi - a[1]) >> 8;
  a[0] = hi;
}

// components -> digit planes uplanes[S][64][Dp], exponents f[64], constants b[64]
__global__ void __launch_bounds__(256)
prep_components_kernel(const double *__restrict__ comps, int P, int D, int Dp, const int *__restrict__ e,
                       const double *__restrict__ center, const double *__restrict__ mu,
                       int8_t *__restrict__ uplanes, int *__restrict__ f, double *__restrict__ b) {
  __shared__ double redm[8], reds[8];
  __shared__ int s_f;
  const int p = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  double mx = 0.0, bs = 0.0;
  if (p < P)
    for (int j = tid; j < D; j += 256) {
      const double v = comps[(size_t)p * D + j];
      mx = fmax(mx, fabs(ldexp(v, e[j])));
      bs = fma(mu[j] - center[j], v, bs);
    }
  for (int o = 16; o > 0; o >>= 1) {
    mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    bs += __shfl_xor_sync(0xffffffffu, bs, o);
  }
  if (lane == 0) { redm[warp] = mx; reds[warp] = bs; }
  __syncthreads();
  if (tid == 0) {
    double m = 0.0, s = 0.0;
    for (int w = 0; w < 8; ++w) { m = fmax(m, redm[w]); s += reds[w]; }
    int fp = 0;
    if (m > 0.0 && isfinite(m)) fp = ilogb(m) + 1;
    s_f = fp;
    f[p] = fp;
    b[p] = s;
  }
  __syncthreads();
  const int fp = s_f;
  for (int j = tid; j < Dp; j += 256) {
    int a[6] = {0, 0, 0, 0, 0, 0};
    if (p < P && j < D) digits6p(ldexp(comps[(size_t)p * D + j], e[j] + 46 - fp), a);
#pragma unroll
    for (int s = 0; s < 6; ++s) uplanes[((size_t)s * PJ_N + p) * Dp + j] = (int8_t)a[s];
  }
}

__global__ void __launch_bounds__(PJ_THREADS, 1)
project_i8_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                  ProjParams p, const int *__restrict__ f, const double *__restrict__ bconst,
                  double *__restrict__ scores) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  constexpr uint32_t stage_bytes = PJ_NP * (PJ_A_BYTES + PJ_B_BYTES);
  const uint32_t bars = base + PJ_STAGES * stage_bytes;
  const uint32_t bar_full = bars, bar_empty = bars + 8u * PJ_STAGES, bar_accf = bars + 16u * PJ_STAGES,
                 bar_acce = bar_accf + 8u;
  __shared__ uint32_t s_tmem;
  __shared__ int s_f[PJ_N];
  __shared__ double s_b[PJ_N];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x < PJ_N) { s_f[threadIdx.x] = f[threadIdx.x]; s_b[threadIdx.x] = bconst[threadIdx.x]; }
  if (threadIdx.x == 0) {
    for (int i = 0; i < PJ_STAGES; ++i) { mbar_init(bar_full + 8u * i, 1); mbar_init(bar_empty + 8u * i, 1); }
    mbar_init(bar_accf, 1);
    mbar_init(bar_acce, 4);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&s_tmem)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = s_tmem;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
        for (int ks = 0; ks < p.n_ksteps; ++ks) {
          mbar_wait(bar_empty + 8u * stage, phase ^ 1u);
          const uint32_t fb = bar_full + 8u * stage;
          mbar_expect_tx(fb, stage_bytes);
          const uint32_t sa = base + (uint32_t)stage * stage_bytes;
          const uint32_t sb = sa + PJ_NP * PJ_A_BYTES;
          // one box per operand covers all PJ_NP staged planes (third box dimension): 2 TMA issues per stage
          // instead of 10 from this one thread; the planes land back to back, [plane][row][64 B], as before
          // CTAs start on different column strips (the int32 sums are exact, so the order over K does not
          // matter): all 148 CTAs otherwise read the same 64-byte column offset of their rows at the same time
          const int kc = (ks + (int)blockIdx.x) % p.n_ksteps;
          tma_load_3d(sa, &tmA, kc * PJ_BK, tile * PJ_M, 0, fb);
          tma_load_3d(sb, &tmB, kc * PJ_BK, 0, 0, fb);
          if (++stage == PJ_STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // D = S32, A = B = signed int8, both K-major, M = 128; N per op
      const uint32_t idesc0 = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(PJ_M >> 4) << 24);
      int stage = 0;
      uint32_t phase = 0, acc_phase = 0;
      // per-op descriptor offsets, instruction descriptors and TMEM columns once, in registers: the index
      // arithmetic of 12 MMAs per stage in ONE issuing thread was the longest part of the pipeline
      constexpr int MAXOPS = 8;
      uint32_t a_off[MAXOPS], b_off[MAXOPS], idn[MAXOPS], tdc[MAXOPS];
#pragma unroll
      for (int i = 0; i < MAXOPS; ++i) {
        const bool on = i < p.n_ops;
        a_off[i] = on ? ((uint32_t)p.os[i] * PJ_A_BYTES) >> 4 : 0u;
        b_off[i] = on ? ((uint32_t)p.ot[i] * PJ_B_BYTES) >> 4 : 0u;
        idn[i] = idesc0 | ((uint32_t)((PJ_N * (on ? p.on[i] : 1)) >> 3) << 17);
        tdc[i] = tmem + (uint32_t)(on ? p.od[i] : 0) * PJ_N;
      }
      // K-major, 64B swizzle: rows of 64 B, 8-row groups 512 B apart; the second K step starts 32 B into the row
      const uint64_t dA0 = make_desc(base, 0, 512, 4), dB0 = make_desc(base + PJ_NP * PJ_A_BYTES, 0, 512, 4);
      for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
        mbar_wait(bar_acce, acc_phase ^ 1u);
        tc_fence_after();
        bool first = true;
        for (int ks = 0; ks < p.n_ksteps; ++ks) {
          mbar_wait(bar_full + 8u * stage, phase);
          tc_fence_after();
          const uint32_t sa = base + (uint32_t)stage * stage_bytes;
          const uint32_t sb = sa + PJ_NP * PJ_A_BYTES;
          const uint64_t so = (uint64_t)(((uint32_t)stage * stage_bytes) >> 4);
#pragma unroll
          for (int kk = 0; kk < PJ_BK / 32; ++kk) {
            if (first) {
              // first K step of a tile: the planes of an op are issued one by one with their own accumulate flag
              for (int i = 0; i < p.n_ops; ++i) {
                const uint64_t da = make_desc(sa + (uint32_t)p.os[i] * PJ_A_BYTES + 32u * kk, 0, 512, 4);
                const uint32_t td = tmem + (uint32_t)p.od[i] * PJ_N;
                if (p.ofirst[i]) {
                  for (int j = 0; j < p.on[i]; ++j) {
                    const uint64_t dbj = make_desc(sb + (uint32_t)(p.ot[i] + j) * PJ_B_BYTES + 32u * kk, 0, 512, 4);
                    umma_i8(td + (uint32_t)j * PJ_N, da, dbj, idesc0 | ((uint32_t)(PJ_N >> 3) << 17),
                            ((p.ofirst[i] >> j) & 1) ? 0u : 1u);
                  }
                } else {
                  const uint64_t db = make_desc(sb + (uint32_t)p.ot[i] * PJ_B_BYTES + 32u * kk, 0, 512, 4);
                  umma_i8(td, da, db, idesc0 | ((uint32_t)((PJ_N * p.on[i]) >> 3) << 17), 1u);
                }
              }
              first = false;
            } else {
              const uint64_t oa = so + (uint64_t)(2 * kk), ob = so + (uint64_t)(2 * kk);
#pragma unroll
              for (int i = 0; i < MAXOPS; ++i)
                if (i < p.n_ops)  // planes t .. t+on-1 are adjacent 64-row blocks: one N = 64 on operand
                  umma_i8(tdc[i], dA0 + oa + a_off[i], dB0 + ob + b_off[i], idn[i], 1u);
            }
          }
          umma_commit(bar_empty + 8u * stage);
          if (++stage == PJ_STAGES) { stage = 0; phase ^= 1u; }
        }
        umma_commit(bar_accf);
        acc_phase ^= 1u;
      }
    }
  } else {
    const int q = warp & 3;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
      mbar_wait(bar_accf, acc_phase);
      tc_fence_after();
      const int64_t row = (int64_t)tile * PJ_M + q * 32 + lane;
      // Phase A: the accumulators leave TMEM — all five weights of a 16-column block behind ONE wait, combined
      // into fp64 in registers (64 doubles per thread) — and TMEM is handed back at once, so the next tile's
      // MMAs start while phase B (scaling, stores) is still running.  (With scaling and stores before the
      // hand-back the whole epilogue sat between two tiles' MMAs: 5 accumulators = 320 of the 512 TMEM columns
      // leave no room for a second set.)
      double g[PJ_N];
#pragma unroll
      for (int c = 0; c < PJ_N / 16; ++c) {
        uint32_t v[PJ_NP][16];
#pragma unroll
        for (int d = 0; d < PJ_NP; ++d) {
          const uint32_t ta = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(d * PJ_N + c * 16);
          asm volatile(
              "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
              : "=r"(v[d][0]), "=r"(v[d][1]), "=r"(v[d][2]), "=r"(v[d][3]), "=r"(v[d][4]), "=r"(v[d][5]),
                "=r"(v[d][6]), "=r"(v[d][7]), "=r"(v[d][8]), "=r"(v[d][9]), "=r"(v[d][10]), "=r"(v[d][11]),
                "=r"(v[d][12]), "=r"(v[d][13]), "=r"(v[d][14]), "=r"(v[d][15])
              : "r"(ta));
        }
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          // same order of operations as before (d = 0 first, weights 256^-d): bit-identical scores
          double acc = 0.0, wd = 1.0;
#pragma unroll
          for (int d = 0; d < PJ_NP; ++d) {
            acc = fma((double)(int32_t)v[d][j], wd, acc);
            wd *= 1.0 / 256.0;
          }
          g[c * 16 + j] = acc;
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_acce);
      acc_phase ^= 1u;
      // Phase B: off the MMA warp's critical path
      if (row < p.n_rows) {
        double *o = scores + row * p.P;  // 16-byte aligned when P is even (row pitch 8 P bytes)
        const bool vec = (p.P & 1) == 0;
#pragma unroll
        for (int pc = 0; pc < PJ_N; pc += 2) {
          if (pc + 1 < p.P) {
            const double a = ldexp(g[pc], s_f[pc] - 12) - s_b[pc];
            const double b = ldexp(g[pc + 1], s_f[pc + 1] - 12) - s_b[pc + 1];
            if (vec) {
              *reinterpret_cast<double2 *>(o + pc) = make_double2(a, b);
            } else {
              o[pc] = a;
              o[pc + 1] = b;
            }
          } else if (pc < p.P) {
            o[pc] = ldexp(g[pc], s_f[pc] - 12) - s_b[pc];
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
  }
}

inline int64_t align256(int64_t x) { return (x + 255) & ~(int64_t)255; }
inline int plane_pitch(int D) { return (D + 15) & ~15; }

}  // namespace

extern "C" int64_t am_project_planes_work_bytes(int D, int S) {
  if (D < 1 || S != PJ_S) return 0;
  return align256((int64_t)PJ_S * PJ_N * plane_pitch(D) + 1024) + align256(PJ_N * 4) + align256(PJ_N * 8);
}

extern "C" int am_project_planes(const int8_t *d_planes, int64_t n_rows, int D, int S, const int32_t *d_e,
                                 const double *d_center, const double *d_mu, const double *d_comps, int P,
                                 double *d_scores, void *d_work, void *stream) {
  if (S != PJ_S) return am::fail(AM_EINVAL, "am_project_planes: S must be 6");
  if (D < 1 || P < 1 || P > PJ_N || n_rows < 0 || n_rows > 0x7fffffff)
    return am::fail(AM_EINVAL, "am_project_planes: bad shape (P <= 64)");
  if (n_rows == 0) return AM_OK;
  if (!d_planes || !d_e || !d_center || !d_mu || !d_comps || !d_scores || !d_work)
    return am::fail(AM_EINVAL, "am_project_planes: null pointer");
  EncodeTiledFn encode = get_encode();
  if (!encode) return am::fail(AM_ECUDA, "am_project_planes: cuTensorMapEncodeTiled entry point unavailable");
  cudaStream_t st = (cudaStream_t)stream;
  const int Dp = plane_pitch(D);
  char *wb = (char *)d_work;
  int8_t *uplanes = (int8_t *)wb;
  wb += align256((int64_t)PJ_S * PJ_N * Dp + 1024);
  int *f = (int *)wb;
  wb += align256(PJ_N * 4);
  double *bconst = (double *)wb;
  prep_components_kernel<<<PJ_N, 256, 0, st>>>(d_comps, P, D, Dp, d_e, d_center, d_mu, uplanes, f, bconst);
  AM_LAUNCHED("prep_components_kernel");

  ProjParams p;
  memset(&p, 0, sizeof(p));
  p.n_rows = (int)n_rows;
  p.D = D;
  p.P = P;
  p.n_tiles = (int)((n_rows + PJ_M - 1) / PJ_M);
  p.n_ksteps = (D + PJ_BK - 1) / PJ_BK;
  {
    const int dmax = PJ_W;  // weights 0 .. PJ_W
    int nops = 0;
    bool seen[8] = {false, false, false, false, false, false, false, false};
    auto add_op = [&](int s, int t0, int nt) {
      p.os[nops] = (uint8_t)s; p.ot[nops] = (uint8_t)t0; p.on[nops] = (uint8_t)nt; p.od[nops] = (uint8_t)(s + t0);
      uint8_t fl = 0;
      for (int j = 0; j < nt; ++j)
        if (!seen[s + t0 + j]) { fl |= (uint8_t)(1u << j); seen[s + t0 + j] = true; }
      p.ofirst[nops] = fl;
      ++nops;
    };
    for (int s = 0; s <= PJ_W; ++s) {
      int t0 = 0, left = PJ_W + 1 - s;
      while (left > 0) { const int nt = std::min(left, 4); add_op(s, t0, nt); t0 += nt; left -= nt; }
    }
    p.n_ops = nops;
    p.n_acc = dmax + 1;
  }
  CUtensorMap tmA, tmB;
  {
    cuuint64_t gdimA[3] = {(cuuint64_t)D, (cuuint64_t)n_rows, (cuuint64_t)PJ_S};
    cuuint64_t gstrA[2] = {(cuuint64_t)Dp, (cuuint64_t)Dp * (cuuint64_t)n_rows};
    cuuint64_t gdimB[3] = {(cuuint64_t)D, (cuuint64_t)PJ_N, (cuuint64_t)PJ_S};
    cuuint64_t gstrB[2] = {(cuuint64_t)Dp, (cuuint64_t)Dp * (cuuint64_t)PJ_N};
    cuuint32_t estr[3] = {1, 1, 1};
    cuuint32_t boxA[3] = {PJ_BK, PJ_M, PJ_NP};  // planes 0 .. PJ_W in one box
    cuuint32_t boxB[3] = {PJ_BK, PJ_N, PJ_NP};
    CUresult r1 = encode(&tmA, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, (void *)d_planes, gdimA, gstrA, boxA, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    CUresult r2 = encode(&tmB, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, (void *)uplanes, gdimB, gstrB, boxB, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r1 != CUDA_SUCCESS || r2 != CUDA_SUCCESS) return am::fail(AM_ECUDA, "am_project_planes: cuTensorMapEncodeTiled failed");
  }
  const size_t smem = (size_t)PJ_STAGES * PJ_NP * (PJ_A_BYTES + PJ_B_BYTES) + 16 * PJ_STAGES + 64 + 1024;
  AM_CUDA(cudaFuncSetAttribute(project_i8_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int grid = std::min(p.n_tiles, am::grid_sms());
  project_i8_kernel<<<grid, PJ_THREADS, smem, st>>>(tmA, tmB, p, f, bconst, d_scores);
  AM_LAUNCHED("project_i8_kernel");
  return AM_OK;
}
