// Kernel 1 — canonical k-mer counting over 2-bit-packed contigs (sm_100a).
//
// Replaces record_counter / seq_counter (autometa/common/kmers.py:161-265).
// One warp owns one work item (a contig, or a <=max_item_bases slice of a long
// contig) and a private 4^k-bin histogram in shared memory.  Each lane loads one
// 16-byte vector = 64 bases (coalesced 512 B per warp), plus the first word of
// the next group, and extracts the 64 windows that START in its segment with one
// funnel shift each: with bases packed LSB-first the k-mer code of window j is
// simply bits [2j, 2j+2k) of the stream, so no rolling hash is needed.  The
// histogram is indexed by the raw (non-canonical) window value; canonical folding
// (kmer + reverse complement -> one column, kmers.py:84-88,198-204) happens once
// per contig when the histogram is flushed, through a per-k table.
// Invalid bases (N, IUPAC, lower case) are handled through the sparse exception
// structure written by the packers: one flag bit per 64-base group, a rank array
// and one 64-bit mask per flagged group; the common case touches only the flag.
#include <stdlib.h>

#include "am_common.cuh"

namespace {

constexpr unsigned FULL = 0xffffffffu;

struct FoldCache {
  uint32_t *d_fold[16][8];  // [device][k]
};
FoldCache g_fold = {};

// fold table entry for column c: low 16 bits = swizzled histogram slot of the
// canonical k-mer, high 16 bits = slot of its reverse complement (0xFFFF when
// the k-mer is its own reverse complement — even k only).
// histogram slot of a raw window value: the bank index (low 5 bits) is mixed with the next 5 bits so
// that skewed base composition does not pile windows onto a few banks (measured: without the mix the
// kernel is 10 % slower although it issues 40 % fewer instructions — it is bound by shared-memory
// atomic wavefronts, not by ALU)
__host__ __device__ inline uint32_t swz(uint32_t v) { return v ^ ((v >> 5) & 31u); }

int get_fold_table(int k, uint32_t **out, int *ncols_out) {
  int dev = 0;
  AM_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 16) return am::fail(AM_EINVAL, "device index out of range");
  const uint32_t n = 1u << (2 * k);
  std::vector<uint32_t> canon;  // sorted unique canonical codes
  {
    std::vector<uint32_t> c(n);
    for (uint32_t i = 0; i < n; ++i) c[i] = std::min(i, am::revcomp_code(i, k));
    std::sort(c.begin(), c.end());
    c.erase(std::unique(c.begin(), c.end()), c.end());
    canon.swap(c);
  }
  *ncols_out = (int)canon.size();
  if (g_fold.d_fold[dev][k]) {
    *out = g_fold.d_fold[dev][k];
    return AM_OK;
  }
  std::vector<uint32_t> tab(canon.size());
  for (size_t col = 0; col < canon.size(); ++col) {
    uint32_t code = canon[col];
    uint32_t rc = am::revcomp_code(code, k);
    // window values are LSB-first (first base = lowest digit)
    uint32_t va = am::digit_reverse(code, k);
    uint32_t vb = am::digit_reverse(rc, k);
    uint32_t a = swz(va), b = (rc == code) ? 0xFFFFu : swz(vb);
    tab[col] = a | (b << 16);
  }
  uint32_t *d = nullptr;
  AM_CUDA(cudaMalloc(&d, tab.size() * sizeof(uint32_t)));
  AM_CUDA(cudaMemcpy(d, tab.data(), tab.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
  g_fold.d_fold[dev][k] = d;
  *out = d;
  return AM_OK;
}

__global__ void zero_rows_kernel(int32_t *counts, int32_t *row_sum, const int32_t *rows,
                                 int64_t n_rows, int ncols) {
  for (int64_t r = blockIdx.x; r < n_rows; r += gridDim.x) {
    const int64_t row = rows[r];
    for (int c = threadIdx.x; c < ncols; c += blockDim.x) counts[row * ncols + c] = 0;
    if (row_sum && threadIdx.x == 0) row_sum[row] = 0;
  }
}

template <int K>
__device__ __forceinline__ uint32_t window_slot(const uint32_t (&w)[5], int j) {
  constexpr uint32_t MASK = (1u << (2 * K)) - 1u;
  const int wi = j >> 4, sh = (j & 15) * 2;
  const uint32_t v = __funnelshift_r(w[wi], w[wi + 1], sh) & MASK;
  return swz(v);
}
// byte offset (4 * slot) of window j inside the warp's histogram: one funnel shift + one mask
template <int K>
__device__ __forceinline__ uint32_t window_byte(const uint32_t (&w)[5], int j) {
  constexpr uint32_t MASK4 = ((1u << (2 * K)) - 1u) << 2;
  const int wi = j >> 4, sh = (j & 15) * 2;
  uint32_t v;
  if (sh >= 2) v = __funnelshift_r(w[wi], w[wi + 1], sh - 2);
  else v = w[wi] << 2;  // window 0 of a word: shift left instead (the bits above come from the same word)
  v &= MASK4;
  return v ^ ((v >> 5) & 0x7cu);   // same mix as swz(), on the pre-scaled value
}

// One warp tile: lane l takes the SEG bases that start at window t0 + SEG * l of the work item (SEG = 64: one
// 16-byte load per lane, the steady state; SEG = 32 / 16: the remainder of a contig spread over all 32
// lanes, so that the last tile of a contig costs what it holds instead of a full 2048-window tile — at
// 5 kbp per contig the half-empty last tile was most of the per-contig overhead).  A segment never
// straddles a 64-base group (SEG divides 64), so the exception mask of ONE group (+ K-1 bits of the
// next) covers all its windows.
// The tile is split in its LOAD (codes + validity of the lane's segment) and its COUNT (the shared-memory
// atomics) so that the kernel can issue the loads of the next tile before it counts the current one: the
// global-load latency (the top stall reason, 21 % of the samples in profiles/r2c) then overlaps the atomics.
template <int K, int SEG>
__device__ __forceinline__ void tile_load(const uint4 *__restrict__ codes4, const uint32_t *__restrict__ codes,
                                          const uint32_t *__restrict__ bitmap, const uint32_t *__restrict__ rank,
                                          const uint64_t *__restrict__ exc_mask, int64_t gbase, int t0, int nw,
                                          int lane, uint32_t (&wreg)[5], uint64_t &valid_out) {
  constexpr uint64_t SEGMASK = SEG == 64 ? ~0ull : ((1ull << SEG) - 1ull);
  const int seg = t0 + lane * SEG;
  int nv = nw - seg;
  nv = nv < 0 ? 0 : (nv > SEG ? SEG : nv);
  valid_out = 0ull;
#pragma unroll
  for (int q = 0; q < 5; ++q) wreg[q] = 0u;
  if (nv > 0) {
    const int64_t g = gbase + (seg >> 6);
    const int sub = seg & 63;
    if (SEG == 64) {
      const uint4 x = __ldg(codes4 + g);
      wreg[0] = x.x; wreg[1] = x.y; wreg[2] = x.z; wreg[3] = x.w;
      wreg[4] = __ldg(codes + 4 * g + 4);
    } else if (SEG == 32) {
      const uint2 x = __ldg(reinterpret_cast<const uint2 *>(codes + 4 * g + (sub >> 4)));
      wreg[0] = x.x; wreg[1] = x.y;
      wreg[2] = __ldg(codes + 4 * g + (sub >> 4) + 2);
    } else {
      wreg[0] = __ldg(codes + 4 * g + (sub >> 4));
      wreg[1] = __ldg(codes + 4 * g + (sub >> 4) + 1);
    }
    const int64_t g1 = g + 1;
    const uint32_t bw = __ldg(bitmap + (g >> 5));
    const uint32_t bw1 = ((g1 >> 5) == (g >> 5)) ? bw : __ldg(bitmap + (g1 >> 5));
    const uint32_t bit = (bw >> (g & 31)) & 1u;
    const uint32_t bit1 = (bw1 >> (g1 & 31)) & 1u;
    uint64_t valid = (nv == 64) ? ~0ull : ((1ull << nv) - 1ull);
    if (bit | bit1) {
      uint64_t m = 0, m1 = 0;
      if (bit) m = __ldg(exc_mask + __ldg(rank + (g >> 5)) + __popc(bw & ((1u << (g & 31)) - 1u)));
      if (bit1) m1 = __ldg(exc_mask + __ldg(rank + (g1 >> 5)) + __popc(bw1 & ((1u << (g1 & 31)) - 1u)));
      uint64_t bad = m;
      if (K > 1) {
        const uint64_t hi = m1 & ((1ull << (K - 1)) - 1ull);
#pragma unroll
        for (int s = 1; s < K; ++s) bad |= (m >> s) | (hi << (64 - s));
      }
      if (SEG != 64) bad >>= sub;  // windows of this segment
      valid &= ~bad;
    }
    valid_out = valid & SEGMASK;
  }
}

template <int K, int SEG>
__device__ __forceinline__ void tile_count(const uint32_t (&wreg)[5], uint64_t valid_out, uint32_t *hist) {
  constexpr int NB = 1 << (2 * K);
  constexpr uint64_t SEGMASK = SEG == 64 ? ~0ull : ((1ull << SEG) - 1ull);
  // warp-uniform choice: the common all-valid tile takes the 3-instruction path; a tile with
  // any partial / flagged lane sends its invalid windows to a dummy bin (no divergent branches)
  char *hb = reinterpret_cast<char *>(hist);
  if (__all_sync(FULL, valid_out == SEGMASK)) {
#pragma unroll
    for (int j = 0; j < SEG; ++j) atomicAdd(reinterpret_cast<uint32_t *>(hb + window_byte<K>(wreg, j)), 1u);
  } else {
    const uint32_t vlo = (uint32_t)valid_out, vhi = (uint32_t)(valid_out >> 32);
#pragma unroll
    for (int j = 0; j < SEG; ++j) {
      const uint32_t bit = ((j < 32 ? vlo : vhi) >> (j & 31)) & 1u;
      const uint32_t off = bit ? window_byte<K>(wreg, j) : (uint32_t)(NB * 4);
      atomicAdd(reinterpret_cast<uint32_t *>(hb + off), 1u);
    }
  }
}

template <int K, int WARPS, int MINB>
__global__ void __launch_bounds__(WARPS * 32, MINB)
count_kernel(const uint4 *__restrict__ codes4, const uint32_t *__restrict__ codes,
             const uint32_t *__restrict__ bitmap, const uint32_t *__restrict__ rank,
             const uint64_t *__restrict__ exc_mask, const int64_t *__restrict__ starts,
             const int4 *__restrict__ items, int n_items, const uint32_t *__restrict__ fold,
             int ncols, int32_t *__restrict__ counts, int32_t *__restrict__ col_present,
             int32_t *__restrict__ row_sum, int *work_counter) {
  extern __shared__ uint32_t smem[];
  constexpr int NB = 1 << (2 * K);
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  uint32_t *hist = smem + warp * (NB + 32);  // + dummy bin for invalid windows (padded to 32 words)
  for (int i = lane; i < NB; i += 32) hist[i] = 0;
  __syncwarp();

  uint32_t present = 0;  // bit s <-> column lane + 32*s (used when ncols <= 1024)
  const bool present_in_reg = ncols <= 1024;

  // work items are claimed one ahead: the atomic on the work counter, the item record and the contig's start
  // offset (three dependent global loads) are in flight while the current item is counted / flushed
  int item = 0;
  if (lane == 0) item = atomicAdd(work_counter, 1);
  item = __shfl_sync(FULL, item, 0);
  int4 it = make_int4(0, 0, 0, 0);
  int64_t start = 0;
  if (item < n_items) {
    it = __ldg(items + item);
    start = __ldg(starts + it.x);
  }
  while (item < n_items) {
    int next_item = 0;
    if (lane == 0) next_item = atomicAdd(work_counter, 1);
    const int row = it.x, w0 = it.y, nw = it.z, flags = it.w;
    const int64_t gbase = (start + (int64_t)w0) >> 6;

    // full-width tiles while more than half a tile remains, then the remainder at 32 or 16 bases per lane;
    // the loads of tile i+1 (or of the remainder) are issued before tile i is counted
    const int nfull = nw > AM_TILE_BASES / 2 ? (nw - AM_TILE_BASES / 2 + AM_TILE_BASES - 1) / AM_TILE_BASES : 0;
    const int rem_t0 = nfull * AM_TILE_BASES, rem = nw - rem_t0;
    uint32_t wa[5];
    uint64_t va;
    if (nfull > 0) tile_load<K, 64>(codes4, codes, bitmap, rank, exc_mask, gbase, 0, nw, lane, wa, va);
    for (int i = 0; i < nfull; ++i) {
      uint32_t wb[5];
      uint64_t vb;
      if (i + 1 < nfull) tile_load<K, 64>(codes4, codes, bitmap, rank, exc_mask, gbase, (i + 1) * AM_TILE_BASES, nw, lane, wb, vb);
      else if (rem > AM_TILE_BASES / 4) tile_load<K, 32>(codes4, codes, bitmap, rank, exc_mask, gbase, rem_t0, nw, lane, wb, vb);
      else tile_load<K, 16>(codes4, codes, bitmap, rank, exc_mask, gbase, rem_t0, nw, lane, wb, vb);
      tile_count<K, 64>(wa, va, hist);
#pragma unroll
      for (int q = 0; q < 5; ++q) wa[q] = wb[q];
      va = vb;
    }
    if (nfull == 0) {
      if (rem > AM_TILE_BASES / 4) tile_load<K, 32>(codes4, codes, bitmap, rank, exc_mask, gbase, rem_t0, nw, lane, wa, va);
      else tile_load<K, 16>(codes4, codes, bitmap, rank, exc_mask, gbase, rem_t0, nw, lane, wa, va);
    }
    // the next item's record and start offset: in flight during the remainder tile and the flush
    next_item = __shfl_sync(FULL, next_item, 0);
    int4 nit = make_int4(0, 0, 0, 0);
    int64_t nstart = 0;
    if (next_item < n_items) {
      nit = __ldg(items + next_item);
      nstart = __ldg(starts + nit.x);
    }
    if (rem > AM_TILE_BASES / 4) tile_count<K, 32>(wa, va, hist);
    else if (rem > 0) tile_count<K, 16>(wa, va, hist);
    __syncwarp();

    // flush: fold the raw histogram into canonical columns, store the row
    int32_t total = 0;
    int32_t *out = counts + (int64_t)row * ncols;
    for (int s = 0; s < ncols; s += 32) {
      const int col = s + lane;
      if (col < ncols) {
        const uint32_t ab = __ldg(fold + col);
        const uint32_t a = ab & 0xFFFFu, b = ab >> 16;
        uint32_t c = hist[a];
        if (b != 0xFFFFu) c += hist[b];
        if (flags & 1) {
          if (c) atomicAdd(out + col, (int32_t)c);
        } else {
          __stcs(out + col, (int32_t)c);
        }
        total += (int32_t)c;
        if (c) {
          if (present_in_reg) present |= 1u << (s >> 5);
          else col_present[col] = 1;
        }
      }
    }
    if (row_sum) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) total += __shfl_xor_sync(FULL, total, o);
      if (lane == 0) {
        if (flags & 1) atomicAdd(row_sum + row, total);
        else row_sum[row] = total;
      }
    }
    __syncwarp();
    for (int i = lane; i < NB; i += 32) hist[i] = 0;
    __syncwarp();
    item = next_item;
    it = nit;
    start = nstart;
  }
  if (present_in_reg) {
    for (int s = 0; s * 32 < ncols; ++s) {
      const int col = s * 32 + lane;
      if (col < ncols && ((present >> s) & 1u)) col_present[col] = 1;
    }
  }
}

template <int K, int WARPS, int MINB>
int launch_count(const uint32_t *d_codes, const uint32_t *d_bitmap, const uint32_t *d_rank,
                 const uint64_t *d_exc_mask, const int64_t *d_starts, const int32_t *d_items,
                 int64_t n_items, const uint32_t *fold, int ncols, int32_t *d_counts,
                 int32_t *d_col_present, int32_t *d_row_sum, int *d_counter, cudaStream_t st) {
  auto kern = count_kernel<K, WARPS, MINB>;
  const size_t smem = (size_t)WARPS * ((1u << (2 * K)) + 32) * sizeof(uint32_t);
  if (smem > 48 * 1024)
    AM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int occ = 0;
  AM_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, WARPS * 32, smem));
  if (occ < 1) return am::fail(AM_ECUDA, "count kernel does not fit on an SM");
  static const int occ_cap = [] {
    const char *v = getenv("AM_COUNT_OCC");  // measurement knob: resident CTAs per SM the grid is sized for
    return v ? atoi(v) : 0;
  }();
  if (occ_cap > 0 && occ_cap < occ) occ = occ_cap;
  int64_t grid = (int64_t)am::num_sms() * occ;
  const int64_t need = (n_items + WARPS - 1) / WARPS;
  if (grid > need) grid = need;
  if (grid < 1) grid = 1;
  kern<<<(unsigned)grid, WARPS * 32, smem, st>>>(
      reinterpret_cast<const uint4 *>(d_codes), d_codes, d_bitmap, d_rank, d_exc_mask, d_starts,
      reinterpret_cast<const int4 *>(d_items), (int)n_items, fold, ncols, d_counts, d_col_present,
      d_row_sum, d_counter);
  AM_LAUNCHED("count_kernel");
  return AM_OK;
}

}  // namespace

extern "C" int am_count_kmers(const uint32_t *d_codes, const uint32_t *d_exc_bitmap,
                              const uint32_t *d_exc_rank, const uint64_t *d_exc_mask,
                              const int64_t *d_starts, const int32_t *d_items, int64_t n_items,
                              const int32_t *d_multi_rows, int64_t n_multi, int64_t n_contigs,
                              int k, int32_t *d_counts, int32_t *d_col_present,
                              int32_t *d_row_sum, void *d_work, void *stream) {
  if (k < 1 || k > 7) return am::fail(AM_EINVAL, "am_count_kmers: k must be in 1..7");
  if (n_items < 0 || n_items > 0x7fffffff) return am::fail(AM_EINVAL, "am_count_kmers: bad n_items");
  if (n_items == 0 || n_contigs == 0) return AM_OK;
  if (!d_codes || !d_exc_bitmap || !d_exc_rank || !d_starts || !d_items || !d_counts ||
      !d_col_present || !d_work)
    return am::fail(AM_EINVAL, "am_count_kmers: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  uint32_t *fold = nullptr;
  int ncols = 0;
  int rc = get_fold_table(k, &fold, &ncols);
  if (rc) return rc;
  int *counter = (int *)d_work;
  AM_CUDA(cudaMemsetAsync(counter, 0, sizeof(int), st));
  if (n_multi > 0) {
    if (!d_multi_rows) return am::fail(AM_EINVAL, "am_count_kmers: multi rows missing");
    unsigned grid = (unsigned)std::min<int64_t>(n_multi, 4096);
    zero_rows_kernel<<<grid, 256, 0, st>>>(d_counts, d_row_sum, d_multi_rows, n_multi, ncols);
    AM_LAUNCHED("zero_rows_kernel");
  }
  // resident CTAs per SM the register allocation is sized for: 6 (40 registers, 12 bytes spilled) or 5 (48
  // registers); AM_COUNT_MINB=5 selects the latter (timing experiments)
  const char *mb = getenv("AM_COUNT_MINB");
  const bool five = mb && mb[0] == '5';
#define AM_COUNT_CASE(KK, WW)                                                                    \
  case KK:                                                                                       \
    return five ? launch_count<KK, WW, 5>(d_codes, d_exc_bitmap, d_exc_rank, d_exc_mask, d_starts, \
                                          d_items, n_items, fold, ncols, d_counts, d_col_present,  \
                                          d_row_sum, counter, st)                                  \
                : launch_count<KK, WW, 6>(d_codes, d_exc_bitmap, d_exc_rank, d_exc_mask, d_starts, \
                                          d_items, n_items, fold, ncols, d_counts, d_col_present,  \
                                          d_row_sum, counter, st);
  switch (k) {
    AM_COUNT_CASE(1, 8)
    AM_COUNT_CASE(2, 8)
    AM_COUNT_CASE(3, 8)
    AM_COUNT_CASE(4, 8)
    AM_COUNT_CASE(5, 8)
    AM_COUNT_CASE(6, 8)
    AM_COUNT_CASE(7, 3)
  }
  return am::fail(AM_EINVAL, "am_count_kmers: unsupported k");
}
