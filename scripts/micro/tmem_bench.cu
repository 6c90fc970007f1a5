// Microbenchmark: tcgen05.ld (TMEM -> registers) bandwidth per SM on sm_100a.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
template <int NW>
__global__ void k(float* out, int iters) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = slot + (((uint32_t)(warp & 3) * 32) << 16);
  uint32_t acc = 0;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    uint32_t v[32];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
          "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
          : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
            "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
            "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
            "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
          : "r"(base + (uint32_t)(c * 32 + ((warp >> 2) * 128) % 512))
          : "memory");
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      acc ^= v[0] ^ v[31];
    }
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = (float)acc;
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    double bytes = (double)iters * 4 * 32 * 4 * 32 * NW;  // per SM (one CTA per SM)
    printf("warps=%d: %.1f clk/iter, %.1f B/clk/SM\n", NW, (double)(t1 - t0) / iters, bytes / (double)(t1 - t0));
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(slot), "r"(512u) : "memory");
}
int main() {
  float* d; cudaMalloc(&d, 148 * 1024 * 4);
  cudaFuncSetAttribute(k<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaFuncSetAttribute(k<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaFuncSetAttribute(k<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  k<4><<<148, 128, 200 * 1024>>>(d, 2000); cudaDeviceSynchronize();
  k<8><<<148, 256, 200 * 1024>>>(d, 2000); cudaDeviceSynchronize();
  k<16><<<148, 512, 200 * 1024>>>(d, 2000); cudaDeviceSynchronize();
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
