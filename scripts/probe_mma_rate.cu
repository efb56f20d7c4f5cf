// Raw issue rate of tcgen05.mma kind::i8 / kind::f16 on one SM for K-major vs MN-major operands.
// Operands are uninitialised shared memory (values are irrelevant for timing).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../autometa_b200/csrc/am_tc.cuh"
using namespace amtc;

__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}

// mode 0: i8 MN-major (as the Gram kernel), 1: i8 K-major SW128, 2: bf16 K-major SW128
__global__ void __launch_bounds__(128) probe(int mode, int reps, int N, long long *out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint32_t tmem_base;
  __shared__ __align__(8) uint64_t bar;
  const uint32_t sa = (smem_u32(smem) + 1023u) & ~1023u, sb = sa + 32768u;
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_base)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (threadIdx.x == 0) mbar_init(smem_u32(&bar), 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t td = tmem_base;
  if (threadIdx.x == 0) {
    uint32_t idesc;
    uint64_t da, db;
    if (mode == 0) {        // i8, both MN-major (bits 15,16), SW128 A / SW64 B like gram_i8_kernel
      idesc = (2u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(128 >> 4) << 24) | ((uint32_t)(N >> 3) << 17);
      da = make_desc(sa, 32 * 128, 1024, 2);
      db = make_desc(sb, 2048, 512, 4);
    } else if (mode == 1) { // i8 K-major, SW128 atoms (128 B = 128 k-values wide; one MMA uses 32 of them)
      idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(128 >> 4) << 24) | ((uint32_t)(N >> 3) << 17);
      da = make_desc(sa, 0, 1024, 2);
      db = make_desc(sb, 0, 1024, 2);
    } else {                // bf16 K-major: D=f32 (1<<4), A=B=bf16 (1<<7, 1<<10)
      idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(128 >> 4) << 24) | ((uint32_t)(N >> 3) << 17);
      da = make_desc(sa, 0, 1024, 2);
      db = make_desc(sb, 0, 1024, 2);
    }
    const long long t0 = clock64();
    if (mode == 3) {  // i8 MN-major, operands rotate through 4 x (A 4 KB + B 8 KB) like pipeline stages
      const uint32_t idm = (2u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(128 >> 4) << 24) | ((uint32_t)(N >> 3) << 17);
      for (int r = 0; r < reps; ++r) {
        const uint32_t o = (uint32_t)(r & 3) * 16384u;
        umma_i8(td + (uint32_t)((r & 1) * 256), make_desc(sa + o, 32 * 128, 1024, 2), make_desc(sa + o + 4096u, 2048, 512, 4), idm, 1u);
      }
    } else
    for (int r = 0; r < reps; ++r) {
      if (mode == 2) umma_f16(td + (uint32_t)((r & 1) * 256), da, db, idesc, 1u);
      else umma_i8(td + (uint32_t)((r & 1) * 256), da, db, idesc, 1u);
    }
    umma_commit(smem_u32(&bar));
    mbar_wait(smem_u32(&bar), 0);
    const long long t1 = clock64();
    out[blockIdx.x] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(td) : "memory");
}

__device__ __forceinline__ void umma_i8_2sm(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
// cta_group::2: M = 256 (128 rows per CTA), each CTA supplies N/2 columns of B
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128) probe2(int rotate, int reps, int N, long long *out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint32_t tmem_base;
  __shared__ __align__(8) uint64_t bar;
  __shared__ __align__(8) uint64_t bar2;
  uint32_t rank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  const uint32_t sa = (smem_u32(smem) + 1023u) & ~1023u;
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_base)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); mbar_init(smem_u32(&bar2), 1); }
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  tc_fence_after();
  const uint32_t td = tmem_base;
  if (threadIdx.x == 0 && rank == 0) {
    const uint32_t idm = (2u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(256 >> 4) << 24) | ((uint32_t)(N >> 3) << 17);
    const long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
      const uint32_t o = rotate ? (uint32_t)(r & 3) * 16384u : 0u;
      const uint32_t tcol = rotate == 2 ? 0u : (uint32_t)((r & 1) * 256);  // rotate 2: every MMA accumulates into the same columns
      umma_i8_2sm(td + tcol, make_desc(sa + o, 32 * 128, 1024, 2), make_desc(sa + o + 4096u, 2048, 512, 4), idm, 1u);
      // rotate 3 / 4: a tcgen05.commit (to a scratch barrier nobody waits on) after every 4th / 8th MMA
      if ((rotate == 3 && (r & 3) == 3) || (rotate == 4 && (r & 7) == 7))
        asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                     ::"r"(smem_u32(&bar2)), "h"((uint16_t)3) : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(&bar)), "h"((uint16_t)1) : "memory");
    mbar_wait(smem_u32(&bar), 0);
    out[blockIdx.x >> 1] = clock64() - t0;
  }
  tc_fence_before();
  __syncthreads();
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(td) : "memory");
}

int main() {
  long long *d;
  cudaMalloc(&d, 148 * sizeof(long long));
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  cudaFuncSetAttribute(probe2, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  const char *names[] = {"i8  MN-major", "i8  K-major ", "bf16 K-major", "i8  MN rotating operands"};
  for (int N : {256, 128})
    for (int rot = 1; rot < 5; ++rot) {
      const int reps = 4000;
      for (int it = 0; it < 2; ++it) probe2<<<148, 128, 100 * 1024>>>(rot, reps, N, d);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("pair: %s\n", cudaGetErrorString(e)); return 1; }
      long long h[74];
      cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
      double avg = 0; for (int i = 0; i < 74; ++i) avg += h[i]; avg /= 74;
      printf("i8 pair (M=256) N=%d rotate=%d: %.1f clk/MMA, %.0f MAC/clk/SM\n", N, rot, avg / reps, 128.0 * N * 32 / (avg / reps));
    }
  for (int N : {256, 128})
    for (int mode = 0; mode < 4; ++mode) {
      const int reps = 4000;
      for (int it = 0; it < 2; ++it) probe<<<148, 128, 100 * 1024>>>(mode, reps, N, d);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("%s: %s\n", names[mode], cudaGetErrorString(e)); return 1; }
      long long h[148];
      cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
      double avg = 0; for (int i = 0; i < 148; ++i) avg += h[i]; avg /= 148;
      const double kper = mode == 2 ? 16 : 32;  // K per instruction
      printf("%s N=%d: %.1f clk/MMA, %.0f MAC/clk/SM\n", names[mode], N, avg / reps, 128.0 * N * kper / (avg / reps));
    }
  return 0;
}
