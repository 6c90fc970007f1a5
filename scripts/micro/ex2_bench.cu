// Microbenchmark: MUFU ex2 throughput for f32 / f16x2 / bf16x2 on sm_100a.
#include <cstdio>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
template <int MODE>
__global__ void k(float* out, int iters) {
  float a0 = threadIdx.x * 1e-3f - 1.f, a1 = a0 - 0.1f, a2 = a0 - 0.2f, a3 = a0 - 0.3f;
  unsigned h0 = 0xB800B900u + threadIdx.x, h1 = h0 + 7, h2 = h0 + 11, h3 = h0 + 13;
  for (int i = 0; i < iters; ++i) {
    if (MODE == 0) {
      asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a0));
      asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a1));
      asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a2));
      asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a3));
      a0 -= 1.5f; a1 -= 1.5f; a2 -= 1.5f; a3 -= 1.5f;
    } else if (MODE == 1) {
      asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(h0));
      asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(h1));
      asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(h2));
      asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(h3));
      h0 ^= 0x80008000u; h1 ^= 0x80008000u; h2 ^= 0x80008000u; h3 ^= 0x80008000u;
    } else {
      asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(h0));
      asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(h1));
      asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(h2));
      asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(h3));
      h0 ^= 0x80008000u; h1 ^= 0x80008000u; h2 ^= 0x80008000u; h3 ^= 0x80008000u;
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + __uint_as_float(h0 ^ h1 ^ h2 ^ h3);
}
template <int MODE>
void run(const char* name) {
  float* d; cudaMalloc(&d, 148 * 8 * 512 * 4);
  int iters = 20000;
  k<MODE><<<148 * 4, 512>>>(d, 100);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k<MODE><<<148 * 4, 512>>>(d, iters);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double instr = 148.0 * 4 * 512 * iters * 4;  // thread-level ex2 instructions
  int mhz; cudaDeviceGetAttribute(&mhz, cudaDevAttrClockRate, 0);
  printf("%s: %.3f ms, %.1f thread-instr/clk/SM (at %d MHz nominal)\n", name, ms, instr / (ms * 1e-3) / 148 / (mhz * 1e3), mhz / 1000);
}
int main() { run<0>("ex2.f32"); run<1>("ex2.f16x2"); run<2>("ex2.bf16x2"); return 0; }
