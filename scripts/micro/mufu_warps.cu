// MUFU.EX2 throughput vs warps per SM sub-partition, 8 independent ex2 in flight per warp.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(float* out, int iters) {
  float a[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = -0.001f * (threadIdx.x + i);
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = a[i] * -0.5f;   // 1 FMUL per MUFU, independent streams
  }
  long long t1 = clock64();
  float s = 0; for (int i = 0; i < 8; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0)
    printf("threads/SM=%d (warps/SMSP=%.1f): %.2f ex2 lanes/clk/SM\n", blockDim.x, blockDim.x / 128.0,
           (double)iters * 8 * blockDim.x / (double)(t1 - t0));
}
int main() {
  float* d; cudaMalloc(&d, 148 * 1024 * 4);
  for (int th : {128, 256, 512, 1024}) { k<<<148, th>>>(d, 4000); cudaDeviceSynchronize(); }
  return 0;
}
