// Kernel 2 — fused row-wise log-ratio normalisation (sm_100a).
//
// Replaces autometa_clr (autometa/common/kmers.py:333-373) and the scikit-bio
// multiplicative_replacement + clr / ilr calls (kmers.py:418-421).
// One warp per output row.  Counts are small integers, so log(n) comes from a
// shared-memory table of fp64 logs (built per CTA with the same log() the slow
// path uses, so both paths give identical bits); only counts beyond the table
// call log().  Row/column drops of am_clr (kmers.py:369) are fused as gathers.
// Column sums of the output (the PCA mean, sklearn _pca.py:560) are produced
// deterministically: per-CTA partials in d_work, reduced in fixed order.
#include "am_common.cuh"

namespace {

constexpr unsigned FULL = 0xffffffffu;
constexpr int LOGT = 2048;   // table holds log(1..LOGT)
constexpr int NWARPS = 8;
constexpr int MAXD = 8192;   // 4^7 / 2

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
  return v;
}

__device__ __forceinline__ double log_int(const double *logt, int n) {
  // n >= 1
  return n <= LOGT ? logt[n - 1] : log((double)n);
}

// METHOD 0: am_clr   out_j = log(x_j + 1) - mean_j log(x_j + 1)
// METHOD 1: clr      p = x/S; q = (p==0 ? delta : p (1 - z delta)); q /= sum q; log; centre
template <int METHOD>
__global__ void __launch_bounds__(NWARPS * 32)
normalize_kernel(const int32_t *__restrict__ counts, int64_t n_out, int D,
                 const int64_t *__restrict__ row_index, const int32_t *__restrict__ col_index,
                 int Dk, double *__restrict__ out, double *__restrict__ partials,
                 double *__restrict__ pmax) {
  extern __shared__ double sm[];
  double *logt = sm;               // LOGT
  double *csum = sm + LOGT;        // NWARPS * Dk (only when partials != nullptr)
  double *cmx = csum + NWARPS * Dk;  // NWARPS * Dk
  for (int i = threadIdx.x; i < LOGT; i += blockDim.x) logt[i] = log((double)(i + 1));
  if (partials)
    for (int i = threadIdx.x; i < 2 * NWARPS * Dk; i += blockDim.x) csum[i] = 0.0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double *mysum = csum + warp * Dk;
  double *mymax = cmx + warp * Dk;
  const int64_t wstride = (int64_t)gridDim.x * NWARPS;
  for (int64_t r = (int64_t)blockIdx.x * NWARPS + warp; r < n_out; r += wstride) {
    const int64_t src = row_index ? row_index[r] : r;
    const int32_t *x = counts + src * D;
    double *o = out + r * Dk;
    if (METHOD == 0) {
      double s = 0.0;
      for (int j = lane; j < Dk; j += 32) {
        const int c = col_index ? col_index[j] : j;
        s += log_int(logt, __ldg(x + c) + 1);
      }
      const double mean = warp_sum(s) / (double)Dk;
      for (int j = lane; j < Dk; j += 32) {
        const int c = col_index ? col_index[j] : j;
        const double v = log_int(logt, __ldg(x + c) + 1) - mean;
        __stcs(o + j, v);
        if (partials) { mysum[j] += v; mymax[j] = fmax(mymax[j], fabs(v)); }
      }
    } else {
      // closure + multiplicative replacement (delta = 1/D^2 over the retained columns)
      double S = 0.0;
      int z = 0;
      for (int j = lane; j < Dk; j += 32) {
        const int c = col_index ? col_index[j] : j;
        const int v = __ldg(x + c);
        S += (double)v;
        z += (v == 0);
      }
      S = warp_sum(S);
#pragma unroll
      for (int o2 = 16; o2 > 0; o2 >>= 1) z += __shfl_xor_sync(FULL, z, o2);
      const double delta = (1.0 / (double)Dk) * (1.0 / (double)Dk);
      const double scale = 1.0 - (double)z * delta;
      // second closure inside clr(): sum of q
      double Q = 0.0;
      for (int j = lane; j < Dk; j += 32) {
        const int c = col_index ? col_index[j] : j;
        const int v = __ldg(x + c);
        Q += v == 0 ? delta : scale * ((double)v / S);
      }
      Q = warp_sum(Q);
      const double lzero = log(delta / Q);
      const double lscale = log(scale / S / Q);
      double sl = 0.0;
      for (int j = lane; j < Dk; j += 32) {
        const int c = col_index ? col_index[j] : j;
        const int v = __ldg(x + c);
        sl += v == 0 ? lzero : (log_int(logt, v) + lscale);
      }
      const double mean = warp_sum(sl) / (double)Dk;
      for (int j = lane; j < Dk; j += 32) {
        const int c = col_index ? col_index[j] : j;
        const int v = __ldg(x + c);
        const double l = (v == 0 ? lzero : (log_int(logt, v) + lscale)) - mean;
        __stcs(o + j, l);
        if (partials) { mysum[j] += l; mymax[j] = fmax(mymax[j], fabs(l)); }
      }
    }
  }
  if (partials) {
    __syncthreads();
    for (int j = threadIdx.x; j < Dk; j += blockDim.x) {
      double s = 0.0;
#pragma unroll
      for (int w = 0; w < NWARPS; ++w) s += csum[w * Dk + j];
      partials[(int64_t)blockIdx.x * Dk + j] = s;
      double m = 0.0;
#pragma unroll
      for (int w = 0; w < NWARPS; ++w) m = fmax(m, cmx[w * Dk + j]);
      pmax[(int64_t)blockIdx.x * Dk + j] = m;
    }
  }
}


// Fast path for the k = 5 table (D = D_keep = 512, no column gather): the row lives in registers
// (lane owns columns 64 m + 2 lane + {0,1}, m = 0..7: 8-byte loads, 16-byte stores, every warp
// instruction covers one contiguous 256 B / 512 B run), column sums and column max|v| are kept in
// registers across all rows of the warp and reduced once per CTA in a fixed order.
template <int METHOD>
__global__ void __launch_bounds__(NWARPS * 32)
normalize512_kernel(const int32_t *__restrict__ counts, int64_t n_out, const int64_t *__restrict__ row_index,
                    double *__restrict__ out, double *__restrict__ psum, double *__restrict__ pmax) {
  constexpr int D = 512;
  extern __shared__ double sm[];
  double *logt = sm;          // LOGT
  double *red = sm + LOGT;    // NWARPS * D (only when psum != nullptr)
  for (int i = threadIdx.x; i < LOGT; i += blockDim.x) logt[i] = log((double)(i + 1));
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double csum[16], cmax[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) { csum[i] = 0.0; cmax[i] = 0.0; }
  const int64_t wstride = (int64_t)gridDim.x * NWARPS;
  for (int64_t r = (int64_t)blockIdx.x * NWARPS + warp; r < n_out; r += wstride) {
    const int64_t src = row_index ? row_index[r] : r;
    const int2 *x2 = reinterpret_cast<const int2 *>(counts + src * D) + lane;
    int xv[16];
#pragma unroll
    for (int m = 0; m < 8; ++m) {
      const int2 t = __ldcs(x2 + 32 * m);
      xv[2 * m] = t.x;
      xv[2 * m + 1] = t.y;
    }
    double l[16];
    if (METHOD == 0) {
      double s = 0.0;
#pragma unroll
      for (int i = 0; i < 16; ++i) { l[i] = log_int(logt, xv[i] + 1); s += l[i]; }
      const double mean = warp_sum(s) / (double)D;
#pragma unroll
      for (int i = 0; i < 16; ++i) l[i] -= mean;
    } else {
      double S = 0.0;
      int z = 0;
#pragma unroll
      for (int i = 0; i < 16; ++i) { S += (double)xv[i]; z += (xv[i] == 0); }
      S = warp_sum(S);
#pragma unroll
      for (int o2 = 16; o2 > 0; o2 >>= 1) z += __shfl_xor_sync(FULL, z, o2);
      const double delta = (1.0 / (double)D) * (1.0 / (double)D);
      const double scale = 1.0 - (double)z * delta;
      double Q = 0.0;
#pragma unroll
      for (int i = 0; i < 16; ++i) Q += xv[i] == 0 ? delta : scale * ((double)xv[i] / S);
      Q = warp_sum(Q);
      const double lzero = log(delta / Q);
      const double lscale = log(scale / S / Q);
      double sl = 0.0;
#pragma unroll
      for (int i = 0; i < 16; ++i) { l[i] = xv[i] == 0 ? lzero : (log_int(logt, xv[i]) + lscale); sl += l[i]; }
      const double mean = warp_sum(sl) / (double)D;
#pragma unroll
      for (int i = 0; i < 16; ++i) l[i] -= mean;
    }
    double2 *o2p = reinterpret_cast<double2 *>(out + r * D) + lane;
#pragma unroll
    for (int m = 0; m < 8; ++m) __stcs(o2p + 32 * m, make_double2(l[2 * m], l[2 * m + 1]));
    if (psum) {
#pragma unroll
      for (int i = 0; i < 16; ++i) { csum[i] += l[i]; cmax[i] = fmax(cmax[i], fabs(l[i])); }
    }
  }
  if (psum) {
    // cross-warp reduction in a fixed order: sums, then maxima, through the same buffer
#pragma unroll
    for (int m = 0; m < 8; ++m) {
      red[warp * D + 64 * m + 2 * lane] = csum[2 * m];
      red[warp * D + 64 * m + 2 * lane + 1] = csum[2 * m + 1];
    }
    __syncthreads();
    for (int j = threadIdx.x; j < D; j += blockDim.x) {
      double t = 0.0;
#pragma unroll
      for (int w = 0; w < NWARPS; ++w) t += red[w * D + j];
      psum[(int64_t)blockIdx.x * D + j] = t;
    }
    __syncthreads();
#pragma unroll
    for (int m = 0; m < 8; ++m) {
      red[warp * D + 64 * m + 2 * lane] = cmax[2 * m];
      red[warp * D + 64 * m + 2 * lane + 1] = cmax[2 * m + 1];
    }
    __syncthreads();
    for (int j = threadIdx.x; j < D; j += blockDim.x) {
      double t = 0.0;
#pragma unroll
      for (int w = 0; w < NWARPS; ++w) t = fmax(t, red[w * D + j]);
      pmax[(int64_t)blockIdx.x * D + j] = t;
    }
  }
}

// balanced base-256 digits (most significant first) of q = round(v), |v| <= 2^46
__device__ __forceinline__ void digits6(double v, int (&a)[6]) {
  const long long q = __double2ll_rn(v);
  int lo = (int)((unsigned)(q + 0x800000LL) & 0xFFFFFFu) - 0x800000;  // balanced low 24 bits
  int hi = (int)((q - lo) >> 24);
  a[5] = ((lo + 128) & 255) - 128; lo = (lo - a[5]) >> 8;
  a[4] = ((lo + 128) & 255) - 128; lo = (lo - a[4]) >> 8;
  a[3] = ((lo + 128) & 255) - 128; hi += (lo - a[3]) >> 8;  // carry into the high half
  a[2] = ((hi + 128) & 255) - 128; hi = (hi - a[2]) >> 8;
  a[1] = ((hi + 128) & 255) - 128; hi = (hi - a[1]) >> 8;
  a[0] = hi;
}

// Fused normalise + quantise (D = 512, S = 6): besides the fp64 row and the column sums, every
// value is written as the six int8 digit planes that the tensor-core Gram (am_gram_planes) and
// projection (am_project_planes) consume, so the normalised matrix is never re-read.  Values are
// taken about a provisional centre (the exact mean is only known after this pass; the Gram is
// corrected by n (mu - centre)(mu - centre)^T afterwards).  Lane owns columns 128 m + 4 lane + e:
// 16-byte count loads, 4-byte digit stores (128 contiguous bytes per warp instruction).
template <int METHOD>
__global__ void __launch_bounds__(NWARPS * 32, 2)  // two CTAs per SM: the pass is latency-bound at 8 warps
normalize512_planes_kernel(const int32_t *__restrict__ counts, int64_t n_out,
                           const int64_t *__restrict__ row_index, double *__restrict__ out,
                           double *__restrict__ psum, const double *__restrict__ center,
                           const double *__restrict__ sc, int8_t *__restrict__ planes) {
  constexpr int D = 512;
  extern __shared__ double sm[];
  double *logt = sm;              // LOGT
  double *cen = sm + LOGT;        // [16][32] permuted: value i of lane l at i*32 + l
  double *scl = cen + D;          // same layout
  double *red = scl + D;          // NWARPS * D
  double *coef = red + NWARPS * D;  // ilr only: sqrt(c / (c + 1)) and 1 / c, same permuted layout
  double *rcp = coef + D;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < LOGT; i += blockDim.x) logt[i] = log((double)(i + 1));
  for (int q = threadIdx.x; q < D; q += blockDim.x) {
    const int i = q >> 5, l = q & 31;
    const int col = 128 * (i >> 2) + 4 * l + (i & 3);
    // ilr has D - 1 output columns; the pad column is written as exact zeros
    const bool pad = METHOD == 2 && col == D - 1;
    cen[q] = (pad || !planes) ? 0.0 : center[col];
    scl[q] = (pad || !planes) ? 0.0 : sc[col];
    if (METHOD == 2) {
      coef[q] = col ? sqrt((double)col / (double)(col + 1)) : 0.0;
      rcp[q] = col ? 1.0 / (double)col : 0.0;
    }
  }
  __syncthreads();
  double csum[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) csum[i] = 0.0;
  const double lim = 70368744177664.0;  // 2^46
  const int64_t wstride = (int64_t)gridDim.x * NWARPS;
  for (int64_t r = (int64_t)blockIdx.x * NWARPS + warp; r < n_out; r += wstride) {
    const int64_t src = row_index ? row_index[r] : r;
    const int4 *x4 = reinterpret_cast<const int4 *>(counts + src * D) + lane;
    int xv[16];
#pragma unroll
    for (int m = 0; m < 4; ++m) {
      const int4 t = __ldcs(x4 + 32 * m);
      xv[4 * m] = t.x; xv[4 * m + 1] = t.y; xv[4 * m + 2] = t.z; xv[4 * m + 3] = t.w;
    }
    double l[16];
    if (METHOD == 0) {
      double s = 0.0;
#pragma unroll
      for (int i = 0; i < 16; ++i) { l[i] = log_int(logt, xv[i] + 1); s += l[i]; }
      const double mean = warp_sum(s) / (double)D;
#pragma unroll
      for (int i = 0; i < 16; ++i) l[i] -= mean;
    } else {
      double S = 0.0;
      int z = 0;
#pragma unroll
      for (int i = 0; i < 16; ++i) { S += (double)xv[i]; z += (xv[i] == 0); }
      S = warp_sum(S);
#pragma unroll
      for (int o2 = 16; o2 > 0; o2 >>= 1) z += __shfl_xor_sync(FULL, z, o2);
      const double delta = (1.0 / (double)D) * (1.0 / (double)D);
      const double scale = 1.0 - (double)z * delta;
      // closure of the replaced row: q = x / S * scale; one reciprocal per row instead of 512
      // divisions (differs from x / S in the last bit at most, 1e-16 of the 1e-12 budget)
      const double sS = scale / S;
      double Q = 0.0;
#pragma unroll
      for (int i = 0; i < 16; ++i) Q += xv[i] == 0 ? delta : sS * (double)xv[i];
      Q = warp_sum(Q);
      const double lzero = log(delta / Q);
      const double lscale = log(scale / S / Q);
      double sl = 0.0;
#pragma unroll
      for (int i = 0; i < 16; ++i) { l[i] = xv[i] == 0 ? lzero : (log_int(logt, xv[i]) + lscale); sl += l[i]; }
      if (METHOD == 1) {
        const double mean = warp_sum(sl) / (double)D;
#pragma unroll
        for (int i = 0; i < 16; ++i) l[i] -= mean;
      } else {
        // ilr: out_{c-1} = sqrt(c/(c+1)) (mean(l[0:c]) - l[c]) (centring cancels).  Exclusive prefix
        // sums in column order c = 128 m + 4 lane + e: lane-local, then a warp scan per 128-column block.
        double v[16];
        double carry = 0.0;
#pragma unroll
        for (int m = 0; m < 4; ++m) {
          const double s4 = (l[4 * m] + l[4 * m + 1]) + (l[4 * m + 2] + l[4 * m + 3]);
          double incl = s4;
#pragma unroll
          for (int o2 = 1; o2 < 32; o2 <<= 1) {
            const double t = __shfl_up_sync(FULL, incl, o2);
            if (lane >= o2) incl += t;
          }
          double pre = carry + (incl - s4);
          carry += __shfl_sync(FULL, incl, 31);
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int i = 4 * m + e;
            v[i] = coef[i * 32 + lane] * (pre * rcp[i * 32 + lane] - l[i]);
            pre += l[i];
          }
        }
        // shift by one column so that lane/slot (m, e) holds output column 128 m + 4 lane + e
#pragma unroll
        for (int m = 0; m < 4; ++m) {
          const double nxt = __shfl_sync(FULL, v[4 * m], (lane + 1) & 31);
          const double wrap = m < 3 ? __shfl_sync(FULL, v[m < 3 ? 4 * (m + 1) : 0], 0) : 0.0;
          l[4 * m] = v[4 * m + 1];
          l[4 * m + 1] = v[4 * m + 2];
          l[4 * m + 2] = v[4 * m + 3];
          l[4 * m + 3] = lane == 31 ? wrap : nxt;
        }
      }
    }
    // either output may be absent (block-uniform): the planes feed the Gram on the critical path, the
    // fp64 rows (the frame `normalize` returns) can be written by a second launch that overlaps the
    // eigensolve — both launches run this same code, so the values are bit-identical
    if (out) {
      double2 *o2p = reinterpret_cast<double2 *>(out + r * D) + 2 * lane;
#pragma unroll
      for (int m = 0; m < 4; ++m) {
        __stcs(o2p + 64 * m, make_double2(l[4 * m], l[4 * m + 1]));
        __stcs(o2p + 64 * m + 1, make_double2(l[4 * m + 2], l[4 * m + 3]));
      }
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) csum[i] += l[i];
    if (planes) {
      uint32_t *pl = reinterpret_cast<uint32_t *>(planes + (size_t)r * D) + lane;
      const size_t plane_words = (size_t)n_out * (D / 4);
#pragma unroll
      for (int m = 0; m < 4; ++m) {
        // balanced base-256 digits without a carry chain: with B = 0x80 in each of the six digit bytes, the
        // plain bytes b_j of q + B satisfy q = sum (b_j - 128) 256^j, and b_j - 128 as an int8 is b_j ^ 0x80.
        // (|q| <= 2^46 < B keeps q + B inside [0, 256^6).)  Same digits as digits6(), 3 instructions instead
        // of ~30; the 4 x 6 bytes of a lane's four columns are then transposed into the six plane words
        // with byte permutes.
        uint32_t lo[4], hi[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int i = 4 * m + e;
          double v = (l[i] - cen[i * 32 + lane]) * scl[i * 32 + lane];
          v = fmin(fmax(v, -lim), lim);
          const unsigned long long B = 0x0000808080808080ull;
          const unsigned long long x = ((unsigned long long)__double2ll_rn(v) + B) ^ B;
          lo[e] = (uint32_t)x;
          hi[e] = (uint32_t)(x >> 32);
        }
        const uint32_t t0 = __byte_perm(lo[0], lo[1], 0x5140), t1 = __byte_perm(lo[2], lo[3], 0x5140);
        const uint32_t t2 = __byte_perm(lo[0], lo[1], 0x7362), t3 = __byte_perm(lo[2], lo[3], 0x7362);
        const uint32_t t4 = __byte_perm(hi[0], hi[1], 0x5140), t5 = __byte_perm(hi[2], hi[3], 0x5140);
        uint32_t word[6];
        word[5] = __byte_perm(t0, t1, 0x5410);  // least significant digit
        word[4] = __byte_perm(t0, t1, 0x7632);
        word[3] = __byte_perm(t2, t3, 0x5410);
        word[2] = __byte_perm(t2, t3, 0x7632);
        word[1] = __byte_perm(t4, t5, 0x5410);
        word[0] = __byte_perm(t4, t5, 0x7632);  // most significant digit
#pragma unroll
        for (int s = 0; s < 6; ++s) __stcs(pl + (size_t)s * plane_words + 32 * m, word[s]);
      }
    }
  }
  if (!psum) return;
  // column sums: cross-warp reduction in a fixed order
#pragma unroll
  for (int i = 0; i < 16; ++i) red[warp * D + 128 * (i >> 2) + 4 * lane + (i & 3)] = csum[i];
  __syncthreads();
  for (int j = threadIdx.x; j < D; j += blockDim.x) {
    double t = 0.0;
#pragma unroll
    for (int w = 0; w < NWARPS; ++w) t += red[w * D + j];
    psum[(int64_t)blockIdx.x * D + j] = t;
  }
}

// ilr: l_j as in clr (before centring; centring cancels), then
//   out_{i-1} = sqrt(i/(i+1)) * (mean(l[0:i]) - l[i]),  i = 1..Dk-1
__global__ void __launch_bounds__(NWARPS * 32)
ilr_kernel(const int32_t *__restrict__ counts, int64_t n_out, int D,
           const int64_t *__restrict__ row_index, const int32_t *__restrict__ col_index, int Dk,
           double *__restrict__ out) {
  extern __shared__ double sm[];
  double *logt = sm;  // LOGT
  const int chunk = (Dk + 31) / 32;
  const int pitch = Dk + Dk / 16 + 2;  // padded: slot(j) = j + j/16
  double *lbuf = sm + LOGT + (threadIdx.x >> 5) * 2 * pitch;
  double *obuf = lbuf + pitch;
  for (int i = threadIdx.x; i < LOGT; i += blockDim.x) logt[i] = log((double)(i + 1));
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t wstride = (int64_t)gridDim.x * NWARPS;
  for (int64_t r = (int64_t)blockIdx.x * NWARPS + warp; r < n_out; r += wstride) {
    const int64_t src = row_index ? row_index[r] : r;
    const int32_t *x = counts + src * D;
    double S = 0.0;
    int z = 0;
    for (int j = lane; j < Dk; j += 32) {
      const int c = col_index ? col_index[j] : j;
      const int v = __ldg(x + c);
      S += (double)v;
      z += (v == 0);
    }
    S = warp_sum(S);
#pragma unroll
    for (int o2 = 16; o2 > 0; o2 >>= 1) z += __shfl_xor_sync(FULL, z, o2);
    const double delta = (1.0 / (double)Dk) * (1.0 / (double)Dk);
    const double scale = 1.0 - (double)z * delta;
    double Q = 0.0;
    for (int j = lane; j < Dk; j += 32) {
      const int c = col_index ? col_index[j] : j;
      const int v = __ldg(x + c);
      Q += v == 0 ? delta : scale * ((double)v / S);
    }
    Q = warp_sum(Q);
    const double lzero = log(delta / Q);
    const double lscale = log(scale / S / Q);
    for (int j = lane; j < Dk; j += 32) {
      const int c = col_index ? col_index[j] : j;
      const int v = __ldg(x + c);
      lbuf[j + (j >> 4)] = v == 0 ? lzero : (log_int(logt, v) + lscale);
    }
    __syncwarp();
    // contiguous chunk per lane, exclusive prefix of chunk sums across the warp
    const int j0 = lane * chunk;
    const int j1 = min(j0 + chunk, Dk);
    double cs = 0.0;
    for (int j = j0; j < j1; ++j) cs += lbuf[j + (j >> 4)];
    double incl = cs;
#pragma unroll
    for (int o2 = 1; o2 < 32; o2 <<= 1) {
      const double t = __shfl_up_sync(FULL, incl, o2);
      if (lane >= o2) incl += t;
    }
    double pre = incl - cs;  // sum of l[0:j0]
    for (int j = j0; j < j1; ++j) {
      const double l = lbuf[j + (j >> 4)];
      if (j >= 1) {
        const double i = (double)j;
        const int oj = j - 1;
        obuf[oj + (oj >> 4)] = sqrt(i / (i + 1.0)) * (pre / i - l);
      }
      pre += l;
    }
    __syncwarp();
    double *o = out + r * (int64_t)(Dk - 1);
    for (int j = lane; j < Dk - 1; j += 32) __stcs(o + j, obuf[j + (j >> 4)]);
    __syncwarp();
  }
}

__global__ void reduce_partials_kernel(const double *__restrict__ partials, const double *__restrict__ pmax,
                                       int n_parts, int D, double *__restrict__ colsum,
                                       double *__restrict__ colmax) {
  // blockDim = (32 columns, 32 part lanes): lane y sums parts y, y+32, ... (fixed order), then a
  // fixed-order tree over the 32 lanes -> deterministic
  __shared__ double rs[32][33], rm[32][33];
  const int j = blockIdx.x * 32 + threadIdx.x;
  double s = 0.0, m = 0.0;
  if (j < D)
    for (int p = threadIdx.y; p < n_parts; p += 32) {
      s += partials[(int64_t)p * D + j];
      if (pmax) m = fmax(m, pmax[(int64_t)p * D + j]);
    }
  rs[threadIdx.y][threadIdx.x] = s;
  rm[threadIdx.y][threadIdx.x] = m;
  __syncthreads();
  if (threadIdx.y == 0 && j < D) {
    double t = 0.0, tm = 0.0;
    for (int y = 0; y < 32; ++y) { t += rs[y][threadIdx.x]; tm = fmax(tm, rm[y][threadIdx.x]); }
    colsum[j] += t;
    if (colmax) colmax[j] = fmax(colmax[j], tm);
  }
}

int norm_grid(int64_t n_rows) {
  int64_t g = (int64_t)am::num_sms() * 4;
  const int64_t need = (n_rows + NWARPS - 1) / NWARPS;
  if (g > need) g = need;
  return (int)(g < 1 ? 1 : g);
}

}  // namespace

extern "C" int64_t am_normalize_work_bytes(int D) {
  return 2 * (int64_t)am::num_sms() * 4 * D * (int64_t)sizeof(double);
}

extern "C" int am_normalize(const int32_t *d_counts, int64_t n_rows_out, int D,
                            const int64_t *d_row_index, const int32_t *d_col_index, int D_keep,
                            int method, double *d_out, double *d_colsum, double *d_colmaxabs,
                            void *d_work, void *stream) {
  if (D < 1 || D > MAXD || D_keep < 1 || D_keep > D) return am::fail(AM_EINVAL, "am_normalize: bad D");
  if (method < 0 || method > 2) return am::fail(AM_EINVAL, "am_normalize: bad method");
  if (n_rows_out == 0) return AM_OK;
  if (!d_counts || !d_out || n_rows_out < 0) return am::fail(AM_EINVAL, "am_normalize: null pointer");
  if (d_colsum && !d_work) return am::fail(AM_EINVAL, "am_normalize: d_colsum needs d_work");
  if (d_colmaxabs && !d_colsum) return am::fail(AM_EINVAL, "am_normalize: d_colmaxabs needs d_colsum");
  if (method == AM_NORM_ILR && (d_colsum || D_keep < 2))
    return am::fail(AM_EINVAL, "am_normalize: ilr does not produce column sums / needs D>=2");
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = norm_grid(n_rows_out);
  if (method == AM_NORM_ILR) {
    const int pitch = D_keep + D_keep / 16 + 2;
    const size_t smem = (size_t)(LOGT + NWARPS * 2 * pitch) * sizeof(double);
    if (smem > 48 * 1024)
      AM_CUDA(cudaFuncSetAttribute(ilr_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (smem > 227 * 1024) return am::fail(AM_EINVAL, "am_normalize: ilr row too wide for shared memory");
    ilr_kernel<<<grid, NWARPS * 32, smem, st>>>(d_counts, n_rows_out, D, d_row_index, d_col_index,
                                                D_keep, d_out);
    AM_LAUNCHED("ilr_kernel");
    return AM_OK;
  }
  double *partials = d_colsum ? (double *)d_work : nullptr;
  double *pmax = partials ? partials + (int64_t)am::num_sms() * 4 * D : nullptr;
  if (D == 512 && D_keep == 512 && !d_col_index) {
    const size_t smem5 = (size_t)(LOGT + (partials ? NWARPS * 512 : 0)) * sizeof(double);
    if (method == AM_NORM_AM_CLR) {
      AM_CUDA(cudaFuncSetAttribute(normalize512_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem5));
      normalize512_kernel<0><<<grid, NWARPS * 32, smem5, st>>>(d_counts, n_rows_out, d_row_index, d_out, partials, pmax);
    } else {
      AM_CUDA(cudaFuncSetAttribute(normalize512_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem5));
      normalize512_kernel<1><<<grid, NWARPS * 32, smem5, st>>>(d_counts, n_rows_out, d_row_index, d_out, partials, pmax);
    }
    AM_LAUNCHED("normalize512_kernel");
    if (partials) {
      reduce_partials_kernel<<<(D_keep + 31) / 32, dim3(32, 32), 0, st>>>(partials, pmax, grid, D_keep, d_colsum, d_colmaxabs);
      AM_LAUNCHED("reduce_partials_kernel");
    }
    return AM_OK;
  }
  const size_t smem = (size_t)(LOGT + (partials ? 2 * NWARPS * D_keep : 0)) * sizeof(double);
  if (smem > 227 * 1024) return am::fail(AM_EINVAL, "am_normalize: row too wide for fused column sums");
  if (method == AM_NORM_AM_CLR) {
    if (smem > 48 * 1024)
      AM_CUDA(cudaFuncSetAttribute(normalize_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    normalize_kernel<0><<<grid, NWARPS * 32, smem, st>>>(d_counts, n_rows_out, D, d_row_index,
                                                         d_col_index, D_keep, d_out, partials, pmax);
  } else {
    if (smem > 48 * 1024)
      AM_CUDA(cudaFuncSetAttribute(normalize_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    normalize_kernel<1><<<grid, NWARPS * 32, smem, st>>>(d_counts, n_rows_out, D, d_row_index,
                                                         d_col_index, D_keep, d_out, partials, pmax);
  }
  AM_LAUNCHED("normalize_kernel");
  if (partials) {
    reduce_partials_kernel<<<(D_keep + 31) / 32, dim3(32, 32), 0, st>>>(partials, pmax, grid, D_keep, d_colsum, d_colmaxabs);
    AM_LAUNCHED("reduce_partials_kernel");
  }
  return AM_OK;
}

extern "C" int am_normalize_planes(const int32_t *d_counts, int64_t n_rows_out, int D,
                                   const int64_t *d_row_index, int method, double *d_out,
                                   double *d_colsum, const double *d_center, const double *d_sc, int S,
                                   int8_t *d_planes, void *d_work, void *stream) {
  if (D != 512 || S != 6) return am::fail(AM_EINVAL, "am_normalize_planes: only D = 512 (k = 5) and S = 6");
  if (method != AM_NORM_AM_CLR && method != AM_NORM_CLR && method != AM_NORM_ILR)
    return am::fail(AM_EINVAL, "am_normalize_planes: unknown method");
  if (n_rows_out == 0) return AM_OK;
  if (!d_counts || n_rows_out < 0 || (!d_out && !d_planes) || (d_planes && (!d_center || !d_sc)) ||
      (d_colsum && !d_work))
    return am::fail(AM_EINVAL, "am_normalize_planes: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = norm_grid(n_rows_out);
  double *partials = d_colsum ? (double *)d_work : nullptr;
  const size_t smem = (size_t)(LOGT + 4 * 512 + NWARPS * 512) * sizeof(double);
  if (method == AM_NORM_AM_CLR) {
    AM_CUDA(cudaFuncSetAttribute(normalize512_planes_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    normalize512_planes_kernel<0><<<grid, NWARPS * 32, smem, st>>>(d_counts, n_rows_out, d_row_index, d_out, partials,
                                                                   d_center, d_sc, d_planes);
  } else if (method == AM_NORM_ILR) {
    AM_CUDA(cudaFuncSetAttribute(normalize512_planes_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    normalize512_planes_kernel<2><<<grid, NWARPS * 32, smem, st>>>(d_counts, n_rows_out, d_row_index, d_out, partials,
                                                                   d_center, d_sc, d_planes);
  } else {
    AM_CUDA(cudaFuncSetAttribute(normalize512_planes_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    normalize512_planes_kernel<1><<<grid, NWARPS * 32, smem, st>>>(d_counts, n_rows_out, d_row_index, d_out, partials,
                                                                   d_center, d_sc, d_planes);
  }
  AM_LAUNCHED("normalize512_planes_kernel");
  if (partials) {
    reduce_partials_kernel<<<(512 + 31) / 32, dim3(32, 32), 0, st>>>(partials, nullptr, grid, 512, d_colsum, nullptr);
    AM_LAUNCHED("reduce_partials_kernel");
  }
  return AM_OK;
}
