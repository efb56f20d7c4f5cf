// fp64 FMA issue rate per SM: W warps per CTA (one CTA per SM), C independent chains per thread.
#include <cstdio>
#include <cuda_runtime.h>
template <int C>
__global__ void k(double *out, int reps, long long *clk) {
  double a[C], x = 1.0000001, y = 1e-9;
#pragma unroll
  for (int i = 0; i < C; ++i) a[i] = threadIdx.x * 1e-3 + i;
  __syncthreads();
  const long long t0 = clock64();
  for (int r = 0; r < reps; ++r) {
#pragma unroll
    for (int i = 0; i < C; ++i) a[i] = fma(a[i], x, y);
  }
  const long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < C; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}
template <int C>
void run(int warps) {
  double *o; long long *c;
  cudaMalloc(&o, 148 * 1024 * 8); cudaMalloc(&c, 148 * 8);
  const int reps = 2000;
  k<C><<<148, warps * 32>>>(o, reps, c);
  k<C><<<148, warps * 32>>>(o, reps, c);
  cudaDeviceSynchronize();
  long long h[148]; cudaMemcpy(h, c, sizeof(h), cudaMemcpyDeviceToHost);
  double avg = 0; for (int i = 0; i < 148; ++i) avg += h[i]; avg /= 148;
  printf("warps=%2d chains=%2d: %.2f clk per warp-DFMA per SM-wide slot, %.1f FMA/clk/SM\n", warps, C,
         avg / (reps * C), (double)warps * 32 * reps * C / avg);
  cudaFree(o); cudaFree(c);
}
int main() {
  for (int w : {4, 8, 16, 32}) { run<1>(w); run<4>(w); run<16>(w); }
  return 0;
}
