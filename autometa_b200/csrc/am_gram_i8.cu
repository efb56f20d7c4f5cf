// Kernel 3 (tensor-core path) — centred Gram G = (X - mu)^T (X - mu) on the 5th-gen tensor
// cores (tcgen05.mma kind::i8, accumulators in TMEM, operands staged by TMA) — sm_100a.
//
// Replaces the BLAS product X.T @ X of sklearn PCA._fit_full(covariance_eigh)
// (sklearn/decomposition/_pca.py:604-610), reached from autometa/common/kmers.py:605.
//
// fp64 accuracy from integer tensor cores (error-free "sliced" product):
//   1. quantize_planes_kernel: every centred value is turned into a fixed-point integer
//        q = round((x - mu_j) * 2^(8S-2-e_j)),   2^e_j > max_i |x_ij - mu_j|,
//      written as S balanced base-256 digits a_0..a_{S-1} in [-128,127] ("planes" of int8,
//      row-major like X, so K = contig rows and the operands are MN-major for the MMA).
//   2. gram_i8_kernel: sum_r q_ri q_rj = sum_{s,t} 256^(2(S-1)-s-t) sum_r a_s[r,i] a_t[r,j];
//      the digit products are exact in int32 and each weight d = s+t gets its own TMEM
//      accumulator P_d (128 x 64 int32, adjacent TMEM columns).  Because the accumulators of
//      A_s x B_t, t = 0..S-1-s, are adjacent and the B planes are adjacent in shared memory, one
//      tcgen05.mma with N = 64 (S-s) (<= 256) covers a whole row of pairs: 9 instructions per K
//      step instead of 22.  Pairs with s+t >= S (relative weight <= 2^-8S) are dropped, except
//      the square term (S/2, S/2), the only one with a systematic contribution (diag of G).  One persistent CTA per SM: warp 0 = TMA producer (planes -> smem,
//      128B/64B swizzle), warp 1 = single-thread tcgen05.mma issuer, warps 2-5 = epilogue
//      (tcgen05.ld -> fp64 combine sum_d 2^-8d P_d -> partial tile in the workspace).
//   3. gram_i8_reduce_kernel: partial tiles are summed over row chunks in a FIXED order
//      (deterministic), scaled by 2^(e_i+e_j-12) and mirrored into the symmetric G.
// With S = 6 the quantisation step is 2^-46 of the column range; the integer arithmetic adds no
// rounding at all, so G is at least as accurate as an fp64 BLAS Gram.
#include <stdlib.h>

#include "am_common.cuh"
#include "am_tc.cuh"

namespace {

constexpr int GI_M = 128;       // G rows per tile  (TMEM lanes)
constexpr int GI_N = 64;        // G cols per tile  (TMEM columns per accumulator)
constexpr int GI_K = 32;        // contig rows per stage = K of one kind::i8 MMA
constexpr int GI_MAXS = 7;      // S * GI_N <= 512 TMEM columns
constexpr int GI_THREADS = 192;
constexpr int GI_A_BYTES = GI_M * GI_K;  // 4096 per plane per stage
constexpr int GI_B_BYTES = GI_N * GI_K;  // 2048 per plane per stage

using namespace amtc;

struct GramI8Params {
  int S;              // digit planes
  int n_rows;         // contig rows
  int rows_per_chunk; // multiple of GI_K
  int n_chunks;
  int TI;             // ceil(D / 128)
  int TJ;             // ceil(D / 64)
  int n_tiles;        // upper-triangular tiles: sum_I (TJ - 2I)
  int stages;
  int n_acc;          // TMEM accumulators (distinct weights d = s + t)
  int n_ops;          // MMA instructions per K step
  // op i: A plane os[i] (128 cols) x B planes [ot[i], ot[i] + on[i]) side by side (N = 64 * on[i])
  // -> accumulators od[i] .. od[i] + on[i] - 1 (adjacent TMEM columns); ofirst[i]: bit j set when
  // accumulator od[i] + j is written for the first time in the K step (overwrite on the first K step)
  uint8_t os[16], ot[16], on[16], od[16], ofirst[16];
  int dbg_mode;       // AM_GRAM_DBG: 0 normal; 1 = no MMA issue (TMA pipeline only); 2 = no TMA (MMAs on
                      // stale smem, garbage result) — the two halves of the pipeline timed in isolation
};

__device__ __forceinline__ void decode_tile(int tile, int TJ, int &I, int &Jb) {
  I = 0;
  int rem = tile;
  while (rem >= TJ - 2 * I) { rem -= TJ - 2 * I; ++I; }
  Jb = 2 * I + rem;
}

// ------------------------------------------------------------------ 2. the tensor-core Gram
__global__ void __launch_bounds__(GI_THREADS, 1)
gram_i8_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               GramI8Params p, double *__restrict__ part) {
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment for the 128B-swizzled tiles
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int S = p.S, NS = p.stages;
  const uint32_t stage_bytes = (uint32_t)S * (GI_A_BYTES + GI_B_BYTES);
  const uint32_t bars = base + (uint32_t)NS * stage_bytes;  // full[NS], empty[NS], acc_full, acc_empty
  const uint32_t bar_full = bars, bar_empty = bars + 8u * NS, bar_accf = bars + 16u * NS, bar_acce = bar_accf + 8u;
  __shared__ uint32_t s_tmem;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_items = p.n_chunks * p.n_tiles;

  if (threadIdx.x == 0) {
    for (int i = 0; i < NS; ++i) { mbar_init(bar_full + 8u * i, 1); mbar_init(bar_empty + 8u * i, 1); }
    mbar_init(bar_accf, 1);
    mbar_init(bar_acce, 4);  // one arrive per epilogue warp
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
    // ===== TMA producer =====
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int chunk = item / p.n_tiles, tile = item % p.n_tiles;
        int I, Jb;
        decode_tile(tile, p.TJ, I, Jb);
        const int row0 = chunk * p.rows_per_chunk;
        const int row1 = min(p.n_rows, row0 + p.rows_per_chunk);
        for (int r = row0; r < row1; r += GI_K) {
          mbar_wait(bar_empty + 8u * stage, phase ^ 1u);
          const uint32_t fb = bar_full + 8u * stage;
          if (p.dbg_mode == 2) { mbar_arrive(fb); if (++stage == NS) { stage = 0; phase ^= 1u; } continue; }
          mbar_expect_tx(fb, stage_bytes);
          const uint32_t sa = base + (uint32_t)stage * stage_bytes;
          const uint32_t sb = sa + (uint32_t)S * GI_A_BYTES;
          for (int s = 0; s < S; ++s) {
            tma_load_3d(sa + (uint32_t)s * GI_A_BYTES, &tmA, I * GI_M, r, s, fb);
            tma_load_3d(sb + (uint32_t)s * GI_B_BYTES, &tmB, Jb * GI_N, r, s, fb);
          }
          if (++stage == NS) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (one thread) =====
    if (lane == 0) {
      // instruction descriptor: D=S32, A=B=signed int8, both MN-major, M=128; N is set per op
      const uint32_t idesc0 = (2u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) |
                              ((uint32_t)(GI_M >> 4) << 24);
      int stage = 0;
      uint32_t phase = 0, acc_phase = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int chunk = item / p.n_tiles;
        const int row0 = chunk * p.rows_per_chunk;
        const int row1 = min(p.n_rows, row0 + p.rows_per_chunk);
        mbar_wait(bar_acce, acc_phase ^ 1u);  // epilogue has drained the accumulators
        tc_fence_after();
        bool first = true;
        for (int r = row0; r < row1; r += GI_K) {
          mbar_wait(bar_full + 8u * stage, phase);
          tc_fence_after();
          const uint32_t sa = base + (uint32_t)stage * stage_bytes;
          const uint32_t sb = sa + (uint32_t)S * GI_A_BYTES;
          for (int i = 0; i < (p.dbg_mode == 1 ? 0 : p.n_ops); ++i) {
            // A: 128 cols x 32 rows, 128B swizzle (atom = 128 B x 8 rows -> SBO 1024)
            const uint64_t da = make_desc(sa + (uint32_t)p.os[i] * GI_A_BYTES, GI_K * 128, 1024, 2);
            // B: on[i] planes side by side = on[i] atoms of 64 cols x 32 rows along N, 64B swizzle
            // (atom = 64 B x 8 rows -> SBO 512; next atom along N = next plane -> LBO 2048)
            const uint64_t db = make_desc(sb + (uint32_t)p.ot[i] * GI_B_BYTES, GI_B_BYTES, 512, 4);
            const uint32_t idesc = idesc0 | ((uint32_t)((GI_N * p.on[i]) >> 3) << 17);
            const uint32_t td = tmem + (uint32_t)p.od[i] * GI_N;
            if (first && p.ofirst[i]) {
              // some accumulators of this op are touched for the first time in this item: issue them
              // one plane at a time so each gets its own overwrite/accumulate flag
              for (int j = 0; j < p.on[i]; ++j) {
                const uint64_t dbj = make_desc(sb + (uint32_t)(p.ot[i] + j) * GI_B_BYTES, GI_B_BYTES, 512, 4);
                const uint32_t idj = idesc0 | ((uint32_t)(GI_N >> 3) << 17);
                umma_i8(td + (uint32_t)j * GI_N, da, dbj, idj, ((p.ofirst[i] >> j) & 1) ? 0u : 1u);
              }
            } else {
              umma_i8(td, da, db, idesc, 1u);
            }
          }
          first = false;
          umma_commit(bar_empty + 8u * stage);  // frees the smem stage when the MMAs have read it
          if (++stage == NS) { stage = 0; phase ^= 1u; }
        }
        umma_commit(bar_accf);  // accumulators complete
        acc_phase ^= 1u;
      }
    }
  } else {
    // ===== epilogue: TMEM -> fp64 partial tile =====
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    uint32_t acc_phase = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      mbar_wait(bar_accf, acc_phase);
      tc_fence_after();
      double *out = part + ((size_t)item * GI_M + (size_t)(q * 32 + lane)) * GI_N;
      for (int c = 0; c < GI_N / 16; ++c) {
        double g[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) g[j] = 0.0;
        double wd = 1.0;
        for (int d = 0; d < p.n_acc; ++d) {
          uint32_t v[16];
          const uint32_t ta = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(d * GI_N + c * 16);
          asm volatile(
              "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
              : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
              : "r"(ta));
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
          for (int j = 0; j < 16; ++j) g[j] = fma((double)(int32_t)v[j], wd, g[j]);
          wd *= 1.0 / 256.0;
        }
#pragma unroll
        for (int j = 0; j < 16; j += 2) *reinterpret_cast<double2 *>(out + c * 16 + j) = make_double2(g[j], g[j + 1]);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_acce);
      acc_phase ^= 1u;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
  }
}

// ------------------------------------------------------------------ 2b. CTA-pair version (cta_group::2)
// The 1-CTA kernel is bound by the tensor pipe's operand consumption from shared memory (profiles/:
// MMA-only time = full time).  tcgen05.mma.cta_group::2 lets two SMs work on one instruction: M = 256
// rows (128 per CTA) share ONE B operand of which each CTA supplies half the columns.  Here the two
// 128-row halves are not two row blocks of G but two DIGIT PLANES of the same 128 G rows: CTA0 holds
// the even planes (s = 0, 2, 4), CTA1 the odd ones (s + 1), so the pair computes
//     CTA0: A_s  x [B_t0 .. B_t0+nt-1]  -> weights s + t        (accumulator index = weight)
//     CTA1: A_s+1 x the same B planes    -> weights s + 1 + t    (accumulator index = weight - 1)
// with 4 MMAs per K step (N = 256, 128, 256, 128) instead of 10, and each CTA stages only 3 A planes
// and 4 B planes (20 KB per stage instead of 36 KB).  CTA1 also forms the weight-6 pairs (1,5), (3,3),
// (5,1): more of the exact product, not less.  Both CTAs write a partial tile; the reduce kernel sums.
constexpr int G2_KS = 2;          // K steps (GI_K = 32 rows each) per pipeline stage: one box load covers 64 rows
constexpr int G2_STAGES = 5;
constexpr int G2_THREADS = 256;  // warp 0: TMA (A planes), 1: MMA issuer, 2-5: epilogue, 6, 7: TMA (B planes)
constexpr uint32_t G2_A_BYTES = G2_KS * GI_A_BYTES, G2_B_BYTES = G2_KS * GI_B_BYTES;  // per plane per stage
constexpr uint32_t G2_STAGE_BYTES = 3u * G2_A_BYTES + 4u * G2_B_BYTES;  // 40960

__device__ __forceinline__ uint32_t cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t mapa_shared(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_load_3d_2sm(uint32_t dst, const CUtensorMap *tm, int c0, int c1, int c2,
                                                uint32_t leader_bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
      ::"r"(dst), "l"((uint64_t)tm), "r"(c0), "r"(c1), "r"(c2), "r"(leader_bar)
      : "memory");
}
__device__ __forceinline__ void umma_i8_2sm(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit_2sm(uint32_t bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
      ::"r"(bar), "h"((uint16_t)3)
      : "memory");
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}

// PROBE = true compiles in the AM_GRAM_DBG switch-off flags (timing probes); the production
// instance has none of those branches (they cost 9 % in the MMA-issuing thread).
template <int NS, bool PROBE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(G2_THREADS, 1)
gram_i8x2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 GramI8Params p, double *__restrict__ part) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bars = base + (uint32_t)NS * G2_STAGE_BYTES;
  const uint32_t bar_full = bars, bar_empty = bars + 8u * NS, bar_accf = bars + 16u * NS, bar_acce = bar_accf + 8u;
  __shared__ uint32_t s_tmem;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_rank();
  const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
  const int n_items = p.n_chunks * p.n_tiles;

  if (threadIdx.x == 0) {
    for (int i = 0; i < NS; ++i) { mbar_init(bar_full + 8u * i, 1); mbar_init(bar_empty + 8u * i, 1); }
    mbar_init(bar_accf, 1);
    mbar_init(bar_acce, 8);  // 4 epilogue warps x 2 CTAs, counted in the leader
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&s_tmem)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem = s_tmem;

  if (warp == 0 || warp >= 6) {
    // ===== TMA producers (both CTAs; every load signals the LEADER's full barrier).  Issuing the 7 box
    // loads of a stage from one thread costs more than its MMAs take, so three warps share them:
    // warp 0 arms the barrier and loads the 3 A planes, warps 6 and 7 load 2 B planes each. =====
    if (lane == 0) {
      const bool do_a = warp == 0;
      int stage = 0;
      uint32_t phase = 0;
      const int bplane[2][4] = {{0, 1, 4, 0}, {2, 3, 5, 1}};
      for (int item = pair; item < n_items; item += n_pairs) {
        const int chunk = item / p.n_tiles, tile = item % p.n_tiles;
        int I, Jb;
        decode_tile(tile, p.TJ, I, Jb);
        const int row0 = chunk * p.rows_per_chunk;
        const int row1 = min(p.n_rows, row0 + p.rows_per_chunk);
        for (int r = row0; r < row1; r += G2_KS * GI_K) {
          mbar_wait(bar_empty + 8u * stage, phase ^ 1u);
          const uint32_t lfb = mapa_shared(bar_full + 8u * stage, 0);
          if (PROBE && (p.dbg_mode & 2)) {  // timing probe: no loads, MMAs run on whatever is in shared memory
            if (rank == 0 && do_a) mbar_arrive(bar_full + 8u * stage);
            if (++stage == NS) { stage = 0; phase ^= 1u; }
            continue;
          }
          const uint32_t sa = base + (uint32_t)stage * G2_STAGE_BYTES;
          const uint32_t sb = sa + 3u * G2_A_BYTES;
          if (do_a) {
            if (rank == 0) mbar_expect_tx(bar_full + 8u * stage, 2u * G2_STAGE_BYTES);
#pragma unroll
            for (int k = 0; k < 3; ++k) tma_load_3d_2sm(sa + (uint32_t)k * G2_A_BYTES, &tmA, I * GI_M, r, 2 * k + (int)rank, lfb);
          } else {
            const int q0 = 2 * (warp - 6);
#pragma unroll
            for (int q = 0; q < 2; ++q)
              tma_load_3d_2sm(sb + (uint32_t)(q0 + q) * G2_B_BYTES, &tmB, Jb * GI_N, r, bplane[rank][q0 + q], lfb);
          }
          if (++stage == NS) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: one thread of the leader CTA drives both SMs =====
    if (lane == 0 && rank == 0) {
      // D = S32, A = B = signed int8, both MN-major, M = 256 (128 rows per CTA)
      const uint32_t idesc0 = (2u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(256 >> 4) << 24);
      const uint32_t id256 = idesc0 | ((uint32_t)(256 >> 3) << 17), id128 = idesc0 | ((uint32_t)(128 >> 3) << 17);
      int stage = 0;
      uint32_t phase = 0, acc_phase = 0;
      // A plane of a stage: 64 rows x 128 B (SW128 atoms of 8 rows); B plane: 64 rows x 64 B (SW64), planes
      // G2_B_BYTES apart (= the stride between the N atoms of one MMA); K step h starts 32 rows further in
      const uint32_t sb0 = base + 3u * G2_A_BYTES;
      const uint64_t dA0 = make_desc(base, GI_K * 128, 1024, 2), dA1 = make_desc(base + G2_A_BYTES, GI_K * 128, 1024, 2),
                     dA2 = make_desc(base + 2u * G2_A_BYTES, GI_K * 128, 1024, 2);
      const uint64_t dB0 = make_desc(sb0, G2_B_BYTES, 512, 4), dB2 = make_desc(sb0 + 2u * G2_B_BYTES, G2_B_BYTES, 512, 4),
                     dB3 = make_desc(sb0 + 3u * G2_B_BYTES, G2_B_BYTES, 512, 4);
      for (int item = pair; item < n_items; item += n_pairs) {
        const int chunk = item / p.n_tiles;
        const int row0 = chunk * p.rows_per_chunk;
        const int row1 = min(p.n_rows, row0 + p.rows_per_chunk);
        mbar_wait(bar_acce, acc_phase ^ 1u);
        tc_fence_after();
        uint32_t acc = 0u;
        for (int r = row0; r < row1; r += G2_KS * GI_K) {
          mbar_wait(bar_full + 8u * stage, phase);
          tc_fence_after();
          // descriptors of this stage = stage-0 descriptors + stage offset (address field, 16-byte units)
          const uint64_t so = (uint64_t)((uint32_t)stage * (G2_STAGE_BYTES >> 4));
#pragma unroll
          for (int h = 0; h < G2_KS; ++h) {
            if (r + h * GI_K >= row1) break;  // chunks end on multiples of 32 rows, not of 64
            const uint64_t oa = so + (uint64_t)(h * (GI_A_BYTES >> 4)), ob = so + (uint64_t)(h * (GI_B_BYTES >> 4));
            const uint64_t a0 = dA0 + oa, a1 = dA1 + oa, a2 = dA2 + oa, b0 = dB0 + ob, b2 = dB2 + ob, b3 = dB3 + ob;
            if (!(PROBE && (p.dbg_mode & 1))) {
              umma_i8_2sm(tmem + 0u * GI_N, a0, b0, id256, acc);   // planes (0|1) x B(0,1 | 2,3)   -> idx 0..3
              umma_i8_2sm(tmem + 4u * GI_N, a0, b2, id128, acc);   // planes (0|1) x B(4 | 5)       -> idx 4,5
              umma_i8_2sm(tmem + 2u * GI_N, a1, b0, id256, 1u);    // planes (2|3) x B(0,1 | 2,3)   -> idx 2..5
              umma_i8_2sm(tmem + 4u * GI_N, a2, b3, id128, 1u);    // planes (4|5) x B(0 | 1)       -> idx 4,5
            }
            acc = 1u;
          }
          umma_commit_2sm(bar_empty + 8u * stage);
          if (++stage == NS) { stage = 0; phase ^= 1u; }
        }
        umma_commit_2sm(bar_accf);
        acc_phase ^= 1u;
      }
    }
  } else {
    // ===== epilogue (both CTAs): accumulator index a holds weight a + rank =====
    const int q = warp & 3;
    uint32_t acc_phase = 0;
    const uint32_t leader_acce = mapa_shared(bar_acce, 0);
    for (int item = pair; item < n_items; item += n_pairs) {
      mbar_wait(bar_accf, acc_phase);
      tc_fence_after();
      const int chunk = item / p.n_tiles, tile = item % p.n_tiles;
      const size_t pidx = ((size_t)(chunk + (int)rank * p.n_chunks) * p.n_tiles + tile);
      double *out = part + (pidx * GI_M + (size_t)(q * 32 + lane)) * GI_N;
      for (int c = 0; c < ((PROBE && (p.dbg_mode & 4)) ? 0 : GI_N / 16); ++c) {
        double g[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) g[j] = 0.0;
        double wd = rank ? (1.0 / 256.0) : 1.0;
        for (int d = 0; d < 6; ++d) {
          uint32_t v[16];
          const uint32_t ta = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(d * GI_N + c * 16);
          asm volatile(
              "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
              : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
              : "r"(ta));
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
          for (int j = 0; j < 16; ++j) g[j] = fma((double)(int32_t)v[j], wd, g[j]);
          wd *= 1.0 / 256.0;
        }
#pragma unroll
        for (int j = 0; j < 16; j += 2) *reinterpret_cast<double2 *>(out + c * 16 + j) = make_double2(g[j], g[j + 1]);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_remote(leader_acce);
      acc_phase ^= 1u;
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
  }
}

// ------------------------------------------------------------------ 1. quantisation into digit planes
__global__ void prep_scales_kernel(const double *__restrict__ mu, const double *__restrict__ colmaxabs,
                                   double bound_all, int D, int S, int *__restrict__ e,
                                   double *__restrict__ sc) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= D) return;
  // |x - centre| <= max|x| + |centre|; max|x| per column when given, else the caller's global bound
  const double bound = (colmaxabs ? colmaxabs[j] : bound_all) + fabs(mu[j]);
  int ej = 0;
  if (bound > 0.0 && isfinite(bound)) ej = ilogb(bound) + 1;  // 2^ej > bound >= max|x - mu|
  e[j] = ej;
  sc[j] = ldexp(1.0, 8 * S - 2 - ej);
}

__global__ void __launch_bounds__(256)
quantize_planes_kernel(const double *__restrict__ X, int64_t n, int D, int Dp, const double *__restrict__ mu,
                       const double *__restrict__ sc, int S, int8_t *__restrict__ planes) {
  // one thread = 4 consecutive columns of one row: a warp reads 1 KB of the row and writes 128
  // contiguous bytes per plane
  const int groups = Dp / 4;
  const int64_t total = n * groups;
  const double lim = ldexp(1.0, 8 * S - 2);
  const bool vec_ok = (D % 2 == 0);  // rows are 16-byte aligned only for even D
  for (int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; w < total; w += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = w / groups;
    const int c0 = (int)(w % groups) * 4;
    double x[4] = {0.0, 0.0, 0.0, 0.0};
    const double *xr = X + r * D + c0;
    if (vec_ok && c0 + 3 < D) {
      const double2 a = __ldcs(reinterpret_cast<const double2 *>(xr));
      const double2 b2 = __ldcs(reinterpret_cast<const double2 *>(xr) + 1);
      x[0] = a.x; x[1] = a.y; x[2] = b2.x; x[3] = b2.y;
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (c0 + j < D) x[j] = __ldcs(xr + j);
    }
    long long q[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = c0 + j;
      double v = 0.0;
      if (c < D) v = (x[j] - __ldg(mu + c)) * __ldg(sc + c);
      v = fmin(fmax(v, -lim), lim);
      q[j] = __double2ll_rn(v);
    }
    for (int s = S - 1; s >= 0; --s) {
      uint32_t packed = 0u;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int a = (int)((q[j] + 128) & 255) - 128;  // balanced digit in [-128, 127]
        q[j] = (q[j] - a) >> 8;
        packed |= (uint32_t)(a & 255) << (8 * j);
      }
      *reinterpret_cast<uint32_t *>(planes + ((size_t)s * n + r) * Dp + c0) = packed;
    }
  }
}

// ------------------------------------------------------------------ 3. fixed-order reduction + scaling + mirror
__global__ void __launch_bounds__(256)
gram_i8_reduce_kernel(const double *__restrict__ part, GramI8Params p, int n_parts, int D,
                      const int *__restrict__ e, double *__restrict__ G) {
  const int tile = blockIdx.x;
  int I, Jb;
  decode_tile(tile, p.TJ, I, Jb);
  for (int el = blockIdx.y * blockDim.x + threadIdx.x; el < GI_M * GI_N; el += gridDim.y * blockDim.x) {
    const int i = I * GI_M + el / GI_N, j = Jb * GI_N + el % GI_N;
    if (i >= D || j >= D || j < i) continue;
    double s = 0.0;
    for (int ch = 0; ch < n_parts; ++ch) s += part[((size_t)ch * p.n_tiles + tile) * (GI_M * GI_N) + el];
    const double g = ldexp(s, e[i] + e[j] - 12);
    G[(size_t)i * D + j] += g;
    if (j != i) G[(size_t)j * D + i] += g;
  }
}

// ------------------------------------------------------------------ host side
inline int64_t align256(int64_t x) { return (x + 255) & ~(int64_t)255; }

// `sms`: the CTAs the launch will use, so that the work items are a whole number of waves.  More SMs never
// give fewer row chunks, so buffers sized with num_sms() fit a launch under any SM reservation.
void gram_i8_shape(int64_t n_rows, int D, int S, GramI8Params *p, int sms) {
  p->S = S;
  p->n_rows = (int)n_rows;
  p->TI = (D + GI_M - 1) / GI_M;
  p->TJ = (D + GI_N - 1) / GI_N;
  int nt = 0;
  for (int I = 0; I < p->TI; ++I) nt += std::max(0, p->TJ - 2 * I);
  p->n_tiles = nt;
  // row chunks: enough work items for ~5 waves of persistent CTAs, chunk <= 16384 rows (int32
  // headroom: S pairs x rows x 2^14 < 2^31), >= 512 rows
  int64_t want_items = (int64_t)sms * 5;
  int64_t chunks = std::max<int64_t>(1, want_items / nt);
  int64_t rpc = (n_rows + chunks - 1) / chunks;
  rpc = std::max<int64_t>(rpc, 512);
  rpc = std::min<int64_t>(rpc, 16384);
  rpc = (rpc + GI_K - 1) / GI_K * GI_K;
  p->rows_per_chunk = (int)rpc;
  p->n_chunks = (int)((n_rows + rpc - 1) / rpc);
  const int stage_bytes = S * (GI_A_BYTES + GI_B_BYTES);
  p->stages = std::max(2, std::min(8, (200 * 1024) / stage_bytes));
  // pairs kept: s + t <= S-1, plus the square term (S/2, S/2) of weight S when S is even — the only
  // dropped term with a systematic (non zero-mean) contribution, on the diagonal of G
  const int dmax = (S % 2 == 0 && S * 64 + 64 <= 512) ? S : S - 1;
  int nops = 0;
  bool seen[8] = {false, false, false, false, false, false, false, false};
  auto add_op = [&](int s, int t0, int nt) {
    p->os[nops] = (uint8_t)s; p->ot[nops] = (uint8_t)t0; p->on[nops] = (uint8_t)nt; p->od[nops] = (uint8_t)(s + t0);
    uint8_t f = 0;
    for (int j = 0; j < nt; ++j)
      if (!seen[s + t0 + j]) { f |= (uint8_t)(1u << j); seen[s + t0 + j] = true; }
    p->ofirst[nops] = f;
    ++nops;
  };
  for (int s = 0; s < S; ++s) {
    // A_s against B_0 .. B_{S-1-s}  (weights d = s .. S-1), at most 4 planes (N = 256) per MMA
    int t0 = 0, left = S - s;
    while (left > 0) { const int nt = std::min(left, 4); add_op(s, t0, nt); t0 += nt; left -= nt; }
  }
  if (dmax == S) add_op(S / 2, S / 2, 1);  // the square term of weight S
  p->n_ops = nops;
  p->n_acc = dmax + 1;
  const char *dm = getenv("AM_GRAM_DBG");
  p->dbg_mode = dm ? atoi(dm) : 0;
}

inline int plane_pitch(int D) { return (D + 15) & ~15; }

}  // namespace

extern "C" int64_t am_planes_bytes(int64_t n_rows, int D, int S) {
  if (n_rows < 0 || D < 1 || S < 2 || S > GI_MAXS) return 0;
  return align256((int64_t)S * n_rows * plane_pitch(D) + 1024);
}

extern "C" int am_plane_scales(const double *d_center, const double *d_colmaxabs, double bound, int D, int S,
                               int32_t *d_e, double *d_sc, void *stream) {
  if (D < 1 || S < 2 || S > GI_MAXS) return am::fail(AM_EINVAL, "am_plane_scales: bad D or S");
  if (!d_center || !d_e || !d_sc) return am::fail(AM_EINVAL, "am_plane_scales: null pointer");
  if (!d_colmaxabs && !(bound >= 0.0)) return am::fail(AM_EINVAL, "am_plane_scales: need colmaxabs or a bound");
  prep_scales_kernel<<<(D + 127) / 128, 128, 0, (cudaStream_t)stream>>>(d_center, d_colmaxabs, bound, D, S, d_e, d_sc);
  AM_LAUNCHED("prep_scales_kernel");
  return AM_OK;
}

extern "C" int am_quantize_planes(const double *d_X, int64_t n_rows, int D, const double *d_center,
                                  const double *d_sc, int S, int8_t *d_planes, void *stream) {
  if (D < 1 || n_rows < 0 || S < 2 || S > GI_MAXS) return am::fail(AM_EINVAL, "am_quantize_planes: bad shape");
  if (n_rows == 0) return AM_OK;
  if (!d_X || !d_center || !d_sc || !d_planes) return am::fail(AM_EINVAL, "am_quantize_planes: null pointer");
  const int Dp = plane_pitch(D);
  const int64_t total = n_rows * (Dp / 4);
  const unsigned grid = (unsigned)std::min<int64_t>((total + 255) / 256, (int64_t)am::num_sms() * 16);
  quantize_planes_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(d_X, n_rows, D, Dp, d_center, d_sc, S, d_planes);
  AM_LAUNCHED("quantize_planes_kernel");
  return AM_OK;
}

extern "C" int64_t am_gram_planes_work_bytes(int64_t n_rows, int D, int S) {
  if (n_rows < 1 || D < 1 || S < 2 || S > GI_MAXS) return 0;
  GramI8Params p;
  gram_i8_shape(n_rows, D, S, &p, am::num_sms());  // upper bound over every reservation
  return align256((int64_t)2 * p.n_chunks * p.n_tiles * GI_M * GI_N * 8);  // x2: the CTA-pair kernel writes two partials
}

extern "C" int am_gram_planes(const int8_t *d_planes, int64_t n_rows, int D, int S, const int32_t *d_e,
                              double *d_G, void *d_work, void *stream) {
  if (D < 1 || n_rows < 0 || n_rows > 0x7fffffff) return am::fail(AM_EINVAL, "am_gram_planes: bad shape");
  if (S < 2 || S > GI_MAXS) return am::fail(AM_EINVAL, "am_gram_planes: S (digit planes) must be in 2..7");
  if (n_rows == 0) return AM_OK;
  if (!d_planes || !d_e || !d_G || !d_work) return am::fail(AM_EINVAL, "am_gram_planes: null pointer");
  EncodeTiledFn encode = get_encode();
  if (!encode) return am::fail(AM_ECUDA, "am_gram_planes: cuTensorMapEncodeTiled entry point unavailable");
  cudaStream_t st = (cudaStream_t)stream;
  GramI8Params p;
  gram_i8_shape(n_rows, D, S, &p, am::grid_sms());
  const int Dp = plane_pitch(D);
  double *part = (double *)d_work;
  CUtensorMap tmA, tmB;
  {
    cuuint64_t gdim[3] = {(cuuint64_t)D, (cuuint64_t)n_rows, (cuuint64_t)S};
    cuuint64_t gstr[2] = {(cuuint64_t)Dp, (cuuint64_t)Dp * (cuuint64_t)n_rows};
    cuuint32_t estr[3] = {1, 1, 1};
    cuuint32_t boxA[3] = {GI_M, GI_K, 1};
    cuuint32_t boxB[3] = {GI_N, GI_K, 1};
    CUresult r1 = encode(&tmA, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, (void *)d_planes, gdim, gstr, boxA, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    CUresult r2 = encode(&tmB, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, (void *)d_planes, gdim, gstr, boxB, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r1 != CUDA_SUCCESS || r2 != CUDA_SUCCESS) return am::fail(AM_ECUDA, "am_gram_planes: cuTensorMapEncodeTiled failed");
  }
  int n_parts = p.n_chunks;
  const char *pm = getenv("AM_GRAM_PAIR");
  const bool use_pair = (S == 6) && !(pm && pm[0] == '0');
  CUtensorMap tmA2, tmB2;  // the pair kernel loads 64-row boxes
  if (use_pair) {
    // CTA pairs (cta_group::2): one cluster of 2 per work item
    {
      cuuint64_t gdim[3] = {(cuuint64_t)D, (cuuint64_t)n_rows, (cuuint64_t)S};
      cuuint64_t gstr[2] = {(cuuint64_t)Dp, (cuuint64_t)Dp * (cuuint64_t)n_rows};
      cuuint32_t estr[3] = {1, 1, 1};
      cuuint32_t boxA[3] = {GI_M, G2_KS * GI_K, 1};
      cuuint32_t boxB[3] = {GI_N, G2_KS * GI_K, 1};
      CUresult r1 = encode(&tmA2, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, (void *)d_planes, gdim, gstr, boxA, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                           CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      CUresult r2 = encode(&tmB2, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, (void *)d_planes, gdim, gstr, boxB, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B,
                           CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r1 != CUDA_SUCCESS || r2 != CUDA_SUCCESS) return am::fail(AM_ECUDA, "am_gram_planes: cuTensorMapEncodeTiled failed");
    }
    const char *sv = getenv("AM_GRAM_STAGES");
    const int ns = sv ? atoi(sv) : G2_STAGES;
    const size_t smem = (size_t)ns * G2_STAGE_BYTES + 16 * (size_t)ns + 64 + 1024;
    const int n_items = p.n_chunks * p.n_tiles;
    const int pairs = std::max(1, std::min(n_items, am::grid_sms() / 2));
    auto launch = [&](auto kern) -> cudaError_t {
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return e;
      kern<<<2 * pairs, G2_THREADS, smem, st>>>(tmA2, tmB2, p, part);
      return cudaGetLastError();
    };
    if (p.dbg_mode || sv) {  // probes: AM_GRAM_DBG bit flags, AM_GRAM_STAGES in {3, 5}
      if (ns == 3) AM_CUDA(launch(gram_i8x2_kernel<3, true>));
      else AM_CUDA(launch(gram_i8x2_kernel<G2_STAGES, true>));
    } else {
      AM_CUDA(launch(gram_i8x2_kernel<G2_STAGES, false>));
    }
    AM_LAUNCHED("gram_i8x2_kernel");
    n_parts = 2 * p.n_chunks;
  } else {
    const size_t smem = (size_t)p.stages * S * (GI_A_BYTES + GI_B_BYTES) + 16 * (size_t)p.stages + 64 + 1024;
    AM_CUDA(cudaFuncSetAttribute(gram_i8_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int n_items = p.n_chunks * p.n_tiles;
    const int grid = std::min(n_items, am::grid_sms());
    gram_i8_kernel<<<grid, GI_THREADS, smem, st>>>(tmA, tmB, p, part);
    AM_LAUNCHED("gram_i8_kernel");
  }
  gram_i8_reduce_kernel<<<dim3(p.n_tiles, GI_M * GI_N / 256), 256, 0, st>>>(part, p, n_parts, D, d_e, d_G);
  AM_LAUNCHED("gram_i8_reduce_kernel");
  return AM_OK;
}

// convenience: scales from (mu, colmaxabs) + quantise + Gram of the planes, all in d_work
extern "C" int64_t am_gram_i8_work_bytes(int64_t n_rows, int D, int S) {
  if (n_rows < 1 || D < 1 || S < 2 || S > GI_MAXS) return 0;
  return am_planes_bytes(n_rows, D, S) + am_gram_planes_work_bytes(n_rows, D, S) + align256((int64_t)D * 4) +
         align256((int64_t)D * 8);
}

extern "C" int am_gram_i8(const double *d_X, int64_t n_rows, int D, const double *d_mu,
                          const double *d_colmaxabs, int S, double *d_G, void *d_work, void *stream) {
  if (D < 1 || n_rows < 0 || n_rows > 0x7fffffff) return am::fail(AM_EINVAL, "am_gram_i8: bad shape");
  if (S < 2 || S > GI_MAXS) return am::fail(AM_EINVAL, "am_gram_i8: S (digit planes) must be in 2..7");
  if (n_rows == 0) return AM_OK;
  if (!d_X || !d_mu || !d_colmaxabs || !d_G || !d_work) return am::fail(AM_EINVAL, "am_gram_i8: null pointer");
  char *b = (char *)d_work;
  int8_t *planes = (int8_t *)b;
  b += am_planes_bytes(n_rows, D, S);
  void *gwork = b;
  b += am_gram_planes_work_bytes(n_rows, D, S);
  int32_t *e = (int32_t *)b;
  b += align256((int64_t)D * 4);
  double *sc = (double *)b;
  int rc = am_plane_scales(d_mu, d_colmaxabs, -1.0, D, S, e, sc, stream);
  if (rc) return rc;
  rc = am_quantize_planes(d_X, n_rows, D, d_mu, sc, S, planes, stream);
  if (rc) return rc;
  return am_gram_planes(planes, n_rows, D, S, e, d_G, gwork, stream);
}
