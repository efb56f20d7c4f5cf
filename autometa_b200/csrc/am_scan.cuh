// Single-CTA exclusive scan shared by the packer and the DBSCAN step.
#pragma once
#include <cuda_runtime.h>

namespace am {

// single-CTA exclusive scan (n up to a few million): out[i] = sum_{j<i} in[j]; out[n] = total
static __global__ void __launch_bounds__(1024) exclusive_scan_kernel(const unsigned *__restrict__ in,
                                                              unsigned *__restrict__ out, int n) {
  __shared__ unsigned warp_tot[32];
  __shared__ unsigned carry_s;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) carry_s = 0;
  __syncthreads();
  for (int base = 0; base < n; base += 1024 * 4) {
    unsigned v[4];
    unsigned s = 0;
    const int i0 = base + tid * 4;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      v[k] = (i0 + k < n) ? in[i0 + k] : 0u;
      s += v[k];
    }
    unsigned incl = s;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    if (warp == 0) {
      unsigned w = warp_tot[lane];
      unsigned wi = w;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const unsigned t = __shfl_up_sync(0xffffffffu, wi, o);
        if (lane >= o) wi += t;
      }
      warp_tot[lane] = wi - w;  // exclusive
    }
    __syncthreads();
    const unsigned carry = carry_s;
    unsigned run = carry + warp_tot[warp] + (incl - s);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (i0 + k < n) out[i0 + k] = run;
      run += v[k];
    }
    __syncthreads();
    if (tid == 1023) carry_s = run;
    __syncthreads();
  }
  if (tid == 0) out[n] = carry_s;
}

// ---- multi-CTA exclusive scan (three launches): per-tile scan + tile totals, scan of the totals
// (single CTA: at most a few thousand tiles), add the tile offsets.  out[n] = total.
constexpr int SCAN_TILE = 4096;  // 1024 threads x 4 items

static __global__ void __launch_bounds__(1024) scan_tiles_kernel(const unsigned *__restrict__ in,
                                                                 unsigned *__restrict__ out, int n,
                                                                 unsigned *__restrict__ tile_tot) {
  __shared__ unsigned warp_tot[32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int i0 = blockIdx.x * SCAN_TILE + tid * 4;
  unsigned v[4], s = 0;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    v[k] = (i0 + k < n) ? in[i0 + k] : 0u;
    s += v[k];
  }
  unsigned incl = s;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) warp_tot[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    const unsigned w = warp_tot[lane];
    unsigned wi = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned t = __shfl_up_sync(0xffffffffu, wi, o);
      if (lane >= o) wi += t;
    }
    warp_tot[lane] = wi - w;
    if (lane == 31) tile_tot[blockIdx.x] = wi;
  }
  __syncthreads();
  unsigned run = warp_tot[warp] + (incl - s);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    if (i0 + k < n) out[i0 + k] = run;
    run += v[k];
  }
}

static __global__ void __launch_bounds__(1024) scan_add_kernel(unsigned *__restrict__ out, int n,
                                                               const unsigned *__restrict__ tile_off) {
  const int i0 = blockIdx.x * SCAN_TILE + threadIdx.x * 4;
  const unsigned off = tile_off[blockIdx.x];
#pragma unroll
  for (int k = 0; k < 4; ++k)
    if (i0 + k < n) out[i0 + k] += off;
  if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) out[n] = tile_off[gridDim.x];
}

inline int64_t scan_tmp_words(int64_t n) { return 2 * ((n + SCAN_TILE - 1) / SCAN_TILE + 2); }

// tmp: scan_tmp_words(n) unsigned.  Launches 3 kernels on `st`; returns the number launched.
static inline int exclusive_scan(const unsigned *in, unsigned *out, int n, unsigned *tmp, cudaStream_t st) {
  const int tiles = (n + SCAN_TILE - 1) / SCAN_TILE;
  if (tiles <= 1) {
    exclusive_scan_kernel<<<1, 1024, 0, st>>>(in, out, n);
    return 1;
  }
  unsigned *tot = tmp, *off = tmp + tiles + 1;
  scan_tiles_kernel<<<tiles, 1024, 0, st>>>(in, out, n, tot);
  exclusive_scan_kernel<<<1, 1024, 0, st>>>(tot, off, tiles);
  scan_add_kernel<<<tiles, 1024, 0, st>>>(out, n, off);
  return 3;
}

}  // namespace am
