// Device packer: ASCII contig bytes already in HBM -> 2-bit codes + exception
// structure (layout documented in include/autometa_b200.h).  The k-mer kernels
// count upper-case A/C/G/T only (autometa/common/kmers.py:197-204: a window counts
// iff the k-mer string or its reverse complement is a key of the ACGT dictionary),
// so every other byte becomes code 0 + an exception bit.
// One thread per 64-base group; bytes are converted four at a time with byte-SIMD
// compares and multiply-gathers.
#include "am_common.cuh"
#include "am_scan.cuh"

namespace {

__device__ __forceinline__ int find_contig(const int64_t *__restrict__ starts, int64_t n, int64_t pos) {
  // last i with starts[i] <= pos
  int64_t lo = 0, hi = n - 1;
  while (lo < hi) {
    const int64_t mid = (lo + hi + 1) >> 1;
    if (__ldg(starts + mid) <= pos) lo = mid; else hi = mid - 1;
  }
  return (int)lo;
}

// convert 4 ASCII bytes: returns 8 bits of codes (byte i -> bits 2i..2i+1) and 4 invalid bits.
// Bits 1 and 2 of 'A' 'C' 'T' 'G' are the 2-bit values 0 1 2 3; the upper-case letter that has those bits is
// rebuilt arithmetically in every byte lane and compared with the input (exact zero-byte test), which costs
// half the instructions of four emulated byte-wise compares.
__device__ __forceinline__ void conv4(uint32_t w, uint32_t &code8, uint32_t &bad4) {
  const uint32_t x0 = (w >> 1) & 0x01010101u, x1 = (w >> 2) & 0x01010101u;
  // A 0x41, C 0x41+2, T 0x41+0x13, G 0x41+2+0x13-0x0f: no carries between byte lanes
  const uint32_t f = 0x41414141u + 2u * x0 + 0x13u * x1 - 0x0fu * (x0 & x1);
  const uint32_t v = w ^ f;
  const uint32_t inv = ((((v & 0x7f7f7f7fu) + 0x7f7f7f7fu) | v) >> 7) & 0x01010101u;  // 1 where the byte differs
  // swap the two bits -> A0 T1 C2 G3; invalid bytes are stored as code 0
  const uint32_t y = ((x0 << 1) | x1) & ((inv ^ 0x01010101u) * 3u);
  code8 = (y * 0x01041040u) >> 24;
  bad4 = ((inv * 0x01020408u) >> 24) & 0xFu;
}

// GC: also count the bases Biopython's gc_fraction(seq, "remove") counts, case-insensitively:
// gc = #{G,C,S}, at = #{A,T,W,U} (autometa/common/metagenome.py:303-320 -> Bio.SeqUtils.gc_fraction)
template <bool GC>
__device__ __forceinline__ void pack_group(const uint8_t *__restrict__ seq, int64_t src, int nbases,
                                           uint32_t (&codes)[4], uint64_t &bad, uint32_t &gc, uint32_t &at) {
  // read 17 aligned words covering [src, src+64)
  const uint32_t *base = reinterpret_cast<const uint32_t *>(seq + (src & ~(int64_t)3));
  const int sh = (int)(src & 3) * 8;
  uint32_t prev = __ldg(base);
  bad = 0;
  uint32_t gcb = 0, atb = 0;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    uint32_t c32 = 0;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int wi = q * 4 + r;
      const uint32_t next = __ldg(base + wi + 1);
      const uint32_t w = __funnelshift_r(prev, next, sh);
      prev = next;
      uint32_t c8, b4;
      conv4(w, c8, b4);
      c32 |= c8 << (8 * r);
      bad |= (uint64_t)b4 << (4 * wi);
      if (GC) {
        const uint32_t u = w & 0xDFDFDFDFu;  // fold case: only b and b^0x20 map to a letter
        // bit 7 of a byte lane is SET where the byte differs from the pattern (exact zero-byte test)
        auto differs = [u](uint32_t pat) {
          const uint32_t v = u ^ pat;
          return ((v & 0x7f7f7f7fu) + 0x7f7f7f7fu) | v;
        };
        const uint32_t ng = differs(0x47474747u) & differs(0x43434343u) & differs(0x53535353u);
        const uint32_t na = differs(0x41414141u) & differs(0x54545454u) & differs(0x57575757u) & differs(0x55555555u);
        const int nb = nbases - 4 * wi;
        const uint32_t keep = nb >= 4 ? 0x80808080u : (nb <= 0 ? 0u : (0x80808080u >> (8 * (4 - nb))));
        gcb += __popc(~ng & keep);
        atb += __popc(~na & keep);
      }
    }
    codes[q] = c32;
  }
  gc = gcb;
  at = atb;
  if (nbases < 64) {
    // bytes past the contig end belong to the next record: clear them
    const uint64_t keep = nbases <= 0 ? 0ull : ((1ull << nbases) - 1ull);
    bad &= keep;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int nb = nbases - 16 * q;
      const uint32_t m = nb >= 16 ? 0xFFFFFFFFu : (nb <= 0 ? 0u : ((1u << (2 * nb)) - 1u));
      codes[q] &= m;
    }
  }
}

template <bool GC>
__global__ void __launch_bounds__(256)
pack_kernel(const uint8_t *__restrict__ seq, const int64_t *__restrict__ offsets,
            const int64_t *__restrict__ starts, int64_t n_contigs, int64_t n_groups,
            uint4 *__restrict__ codes4, uint32_t *__restrict__ bitmap, uint32_t *__restrict__ pop,
            uint32_t *__restrict__ gc_at) {
  // n_groups is padded by the launcher to a multiple of 32 so every warp writes one bitmap word
  const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t c[4] = {0, 0, 0, 0};
  uint64_t bad = 0;
  uint32_t gc = 0, at = 0;
  int i = -1;
  if (g < n_groups && n_contigs > 0) {
    const int64_t pos = g * AM_GROUP_BASES;
    i = find_contig(starts, n_contigs, pos);
    const int64_t off = __ldg(offsets + i);
    const int64_t L = __ldg(offsets + i + 1) - off;
    const int64_t local = pos - __ldg(starts + i);
    if (local < L) {
      const int64_t rem = L - local;
      pack_group<GC>(seq, off + local, rem > 64 ? 64 : (int)rem, c, bad, gc, at);
    }
  }
  if (g < n_groups) codes4[g] = make_uint4(c[0], c[1], c[2], c[3]);
  if (GC) {
    // one atomic pair per (warp, contig): 32 consecutive groups mostly share a contig
    const unsigned peers = __match_any_sync(0xffffffffu, i);
    uint32_t sg = 0, sa = 0;
    for (unsigned m = peers; m; m &= m - 1) {
      const int src = __ffs(m) - 1;
      sg += __shfl_sync(peers, gc, src);
      sa += __shfl_sync(peers, at, src);
    }
    if (i >= 0 && (int)(threadIdx.x & 31) == __ffs(peers) - 1) {
      if (sg) atomicAdd(gc_at + 2 * (int64_t)i, sg);
      if (sa) atomicAdd(gc_at + 2 * (int64_t)i + 1, sa);
    }
  }
  const uint32_t ball = __ballot_sync(0xffffffffu, bad != 0);
  if ((threadIdx.x & 31) == 0) {
    const int64_t w = g >> 5;
    bitmap[w] = ball;
    pop[w] = __popc(ball);
  }
}

__global__ void __launch_bounds__(256)
pack_mask_kernel(const uint8_t *__restrict__ seq, const int64_t *__restrict__ offsets,
                 const int64_t *__restrict__ starts, int64_t n_contigs, int64_t n_groups,
                 const uint32_t *__restrict__ bitmap, const uint32_t *__restrict__ rank,
                 int64_t n_words, uint64_t *__restrict__ exc_mask, int64_t cap,
                 int64_t *__restrict__ n_exc) {
  const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g == 0) n_exc[0] = (int64_t)rank[n_words - 1];
  if (g >= n_groups) return;
  const uint32_t bw = __ldg(bitmap + (g >> 5));
  if (!((bw >> (g & 31)) & 1u)) return;
  const int64_t r = (int64_t)__ldg(rank + (g >> 5)) + __popc(bw & ((1u << (g & 31)) - 1u));
  if (r >= cap) return;
  const int64_t pos = g * AM_GROUP_BASES;
  const int i = find_contig(starts, n_contigs, pos);
  const int64_t off = __ldg(offsets + i);
  const int64_t L = __ldg(offsets + i + 1) - off;
  const int64_t local = pos - __ldg(starts + i);
  uint32_t c[4], gc, at;
  uint64_t bad = 0;
  const int64_t rem = L - local;
  pack_group<false>(seq, off + local, rem > 64 ? 64 : (int)rem, c, bad, gc, at);
  exc_mask[r] = bad;
}

__global__ void gc_finish_kernel(const uint32_t *__restrict__ gc_at, int64_t n, double *__restrict__ gc_content) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double g = (double)gc_at[2 * i], len = g + (double)gc_at[2 * i + 1];
  // gc_fraction(...) * 100 in Python floats; a record without countable bases gives 0
  gc_content[i] = len > 0.0 ? __dmul_rn(__ddiv_rn(g, len), 100.0) : 0.0;
}

}  // namespace

extern "C" int am_gc_finish(const uint32_t *d_gc_at, int64_t n_contigs, double *d_gc_content, void *stream) {
  if (n_contigs < 0 || (n_contigs && (!d_gc_at || !d_gc_content)))
    return am::fail(AM_EINVAL, "am_gc_finish: bad args");
  if (!n_contigs) return AM_OK;
  gc_finish_kernel<<<(unsigned)((n_contigs + 255) / 256), 256, 0, (cudaStream_t)stream>>>(d_gc_at, n_contigs,
                                                                                       d_gc_content);
  AM_LAUNCHED("gc_finish_kernel");
  return AM_OK;
}

extern "C" int64_t am_pack_work_words(int64_t total_packed_bases) {
  const int64_t n_words = am_pack_bitmap_words(total_packed_bases);
  return n_words + am::scan_tmp_words(n_words);
}

extern "C" int am_pack_device(const uint8_t *d_seq, const int64_t *d_offsets, const int64_t *d_starts,
                              int64_t n_contigs, int64_t total_packed_bases, uint32_t *d_codes,
                              uint32_t *d_exc_bitmap, uint32_t *d_exc_rank, uint64_t *d_exc_mask,
                              int64_t exc_capacity, int64_t *d_n_exc, void *d_work, void *stream) {
  return am_pack_device_gc(d_seq, d_offsets, d_starts, n_contigs, total_packed_bases, d_codes, d_exc_bitmap,
                           d_exc_rank, d_exc_mask, exc_capacity, d_n_exc, nullptr, d_work, stream);
}

extern "C" int am_pack_device_gc(const uint8_t *d_seq, const int64_t *d_offsets, const int64_t *d_starts,
                                 int64_t n_contigs, int64_t total_packed_bases, uint32_t *d_codes,
                                 uint32_t *d_exc_bitmap, uint32_t *d_exc_rank, uint64_t *d_exc_mask,
                                 int64_t exc_capacity, int64_t *d_n_exc, uint32_t *d_gc_at, void *d_work,
                                 void *stream) {
  if (n_contigs < 0 || total_packed_bases < 0 || total_packed_bases % AM_GROUP_BASES)
    return am::fail(AM_EINVAL, "am_pack_device: bad sizes");
  if (!d_seq || !d_offsets || !d_starts || !d_codes || !d_exc_bitmap || !d_exc_rank || !d_n_exc || !d_work)
    return am::fail(AM_EINVAL, "am_pack_device: null pointer");
  if (((uintptr_t)d_seq & 3) || ((uintptr_t)d_codes & 15))
    return am::fail(AM_EINVAL, "am_pack_device: d_seq must be 4-byte and d_codes 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t n_groups = total_packed_bases / AM_GROUP_BASES;
  const int64_t n_words = am_pack_bitmap_words(total_packed_bases);
  if (n_words > 0x7fffffff) return am::fail(AM_EINVAL, "am_pack_device: assembly too large for one call");
  // d_work: n_words uint32 popcounts
  uint32_t *pop = (uint32_t *)d_work;
  AM_CUDA(cudaMemsetAsync(d_exc_bitmap, 0, (size_t)n_words * 4, st));
  AM_CUDA(cudaMemsetAsync(pop, 0, (size_t)n_words * 4, st));
  const int64_t threads = (n_groups + 31) / 32 * 32;
  const unsigned blocks = (unsigned)((threads + 255) / 256);
  if (d_gc_at && n_contigs) AM_CUDA(cudaMemsetAsync(d_gc_at, 0, (size_t)n_contigs * 8, st));
  if (blocks) {
    if (d_gc_at)
      pack_kernel<true><<<blocks, 256, 0, st>>>(d_seq, d_offsets, d_starts, n_contigs, n_groups,
                                                reinterpret_cast<uint4 *>(d_codes), d_exc_bitmap, pop, d_gc_at);
    else
      pack_kernel<false><<<blocks, 256, 0, st>>>(d_seq, d_offsets, d_starts, n_contigs, n_groups,
                                                 reinterpret_cast<uint4 *>(d_codes), d_exc_bitmap, pop, nullptr);
    AM_LAUNCHED("pack_kernel");
  }
  // rank[w] = flagged groups before bitmap word w (multi-CTA scan; its tile totals live behind pop in d_work)
  const int n_scan = am::exclusive_scan(pop, d_exc_rank, (int)(n_words - 1), pop + n_words, st);
  for (int i = 0; i < n_scan; ++i) AM_LAUNCHED("exclusive_scan");
  if (blocks) {
    pack_mask_kernel<<<blocks, 256, 0, st>>>(d_seq, d_offsets, d_starts, n_contigs, n_groups,
                                             d_exc_bitmap, d_exc_rank, n_words, d_exc_mask,
                                             exc_capacity, d_n_exc);
    AM_LAUNCHED("pack_mask_kernel");
  }
  return AM_OK;
}
