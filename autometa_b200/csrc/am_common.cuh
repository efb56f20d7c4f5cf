// Shared helpers for libautometa_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <atomic>
#include <algorithm>
#include <vector>

#include "../../include/autometa_b200.h"

namespace am {

extern thread_local char g_err[512];
extern std::atomic<int64_t> g_launches;
extern std::atomic<int> g_reserved_sms;

inline int fail(int code, const char *fmt, const char *a = "", const char *b = "") {
  snprintf(g_err, sizeof(g_err), fmt, a, b);
  return code;
}

inline int cuda_fail(cudaError_t e, const char *where) {
  snprintf(g_err, sizeof(g_err), "%s: %s", where, cudaGetErrorString(e));
  return AM_ECUDA;
}

#define AM_CUDA(call)                                         \
  do {                                                        \
    cudaError_t _e = (call);                                  \
    if (_e != cudaSuccess) return am::cuda_fail(_e, #call);   \
  } while (0)

// after every kernel launch: count it and surface launch-configuration errors
#define AM_LAUNCHED(name)                                     \
  do {                                                        \
    am::g_launches.fetch_add(1, std::memory_order_relaxed);   \
    cudaError_t _e = cudaGetLastError();                      \
    if (_e != cudaSuccess) return am::cuda_fail(_e, name);    \
  } while (0)

inline int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
  }
  return n;
}

// SMs a persistent kernel with a STATIC work split may size its grid for: all of them, minus the ones the
// caller has set aside (am_set_reserved_sms) for a latency-bound kernel that runs next to it — the 16-CTA
// tridiagonalisation cluster of the previous pass in pipeline.FeatureStream.  A static split over 148 CTAs of
// which 16 cannot be placed finishes only after those 16 get an SM, i.e. after the other kernel: the two
// would run one after the other instead of side by side.
inline int grid_sms() {
  const int n = num_sms() - g_reserved_sms.load(std::memory_order_relaxed);
  return n < 2 ? 2 : n;
}

// IEEE double operations that the compiler may not contract into an FMA, on either side of the boundary:
// the eps schedule of the sweep (Python float arithmetic in the reference) has to come out bit-identical
// whether the host or a device kernel advances it.
__host__ __device__ __forceinline__ double mul_rn(double a, double b) {
#ifdef __CUDA_ARCH__
  return __dmul_rn(a, b);
#else
  volatile double r = a * b;
  return r;
#endif
}
__host__ __device__ __forceinline__ double add_rn(double a, double b) {
#ifdef __CUDA_ARCH__
  return __dadd_rn(a, b);
#else
  volatile double r = a + b;
  return r;
#endif
}
__host__ __device__ __forceinline__ double div_rn(double a, double b) {
#ifdef __CUDA_ARCH__
  return __ddiv_rn(a, b);
#else
  volatile double r = a / b;
  return r;
#endif
}

// Cell list of one eps step (am_dbscan.cu): cell edge slightly above eps so that |a-b| <= eps implies cell
// indices differ by <= 1 even after rounding; the cell count per axis is capped so indices stay well inside
// the int range.  out2 = {1 / cell edge, eps^2}.
__host__ __device__ inline void eps_grid(double eps, const double *span, int F, double *out2) {
  double cell = mul_rn(eps, 1.0 + 1.0 / 1048576.0);
  for (int d = 0; d < F; ++d)
    if (div_rn(span[d], cell) > 1.0e9) cell = div_rn(span[d], 1.0e9);
  out2[0] = div_rn(1.0, cell);
  out2[1] = mul_rn(eps, eps);
}

// reverse complement of an MSB-first base-4 code with A=0,T=1,C=2,G=3
inline uint32_t revcomp_code(uint32_t code, int k) {
  uint32_t rc = 0;
  for (int i = 0; i < k; ++i) {
    rc = (rc << 2) | ((code & 3u) ^ 1u);
    code >>= 2;
  }
  return rc;
}
// reverse the base-4 digits (MSB-first code <-> LSB-first window value)
inline uint32_t digit_reverse(uint32_t code, int k) {
  uint32_t r = 0;
  for (int i = 0; i < k; ++i) {
    r = (r << 2) | (code & 3u);
    code >>= 2;
  }
  return r;
}

}  // namespace am
