// Fixed cost of one kernel boundary inside a CUDA graph on B200, for the launch shapes the UNet program uses.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/launch_floor launch_floor.cu && /tmp/launch_floor
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)

__global__ void __launch_bounds__(320, 1) k_empty(int* sink) {
  extern __shared__ uint8_t sm[];
  if (sink && threadIdx.x == 0 && blockIdx.x == 1 << 30) sink[0] = sm[0];
}
__global__ void __launch_bounds__(320, 1) k_tmem(int* sink) {
  extern __shared__ uint8_t sm[];
  __shared__ uint32_t slot;
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"((uint32_t)__cvta_generic_to_shared(&slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  __syncthreads();
  const uint32_t base = slot;
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(base) : "memory");
  if (sink && threadIdx.x == 0 && blockIdx.x == 1 << 30) sink[0] = sm[0];
}
__global__ void k_small(int* sink) {
  if (sink && threadIdx.x == 0 && blockIdx.x == 1 << 30) sink[0] = 1;
}

template <typename F>
static float time_graph(F&& body, int n, cudaStream_t s) {
  cudaGraph_t g; cudaGraphExec_t ge;
  cudaStreamBeginCapture(s, cudaStreamCaptureModeGlobal);
  for (int i = 0; i < n; ++i) body(i);
  cudaStreamEndCapture(s, &g);
  cudaGraphInstantiate(&ge, g, 0);
  for (int i = 0; i < 3; ++i) cudaGraphLaunch(ge, s);
  cudaStreamSynchronize(s);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0, s);
  for (int i = 0; i < 10; ++i) cudaGraphLaunch(ge, s);
  cudaEventRecord(e1, s); cudaStreamSynchronize(s);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  cudaGraphExecDestroy(ge); cudaGraphDestroy(g);
  return ms * 1000.f / (10.f * n);
}

int main() {
  cudaStream_t s; CK(cudaStreamCreate(&s));
  const int big = 225 * 1024, half = 110 * 1024;
  CK(cudaFuncSetAttribute(k_empty, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
  CK(cudaFuncSetAttribute(k_tmem, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
  const int N = 200;
  printf("empty 148x320, 225 KB smem          : %.2f us / launch\n", time_graph([&](int) { k_empty<<<148, 320, big, s>>>(nullptr); }, N, s));
  printf("empty 148x320, 110 KB smem          : %.2f us / launch\n", time_graph([&](int) { k_empty<<<148, 320, half, s>>>(nullptr); }, N, s));
  printf("empty 148x320, 0 KB smem            : %.2f us / launch\n", time_graph([&](int) { k_empty<<<148, 320, 0, s>>>(nullptr); }, N, s));
  printf("empty 296x320, 110 KB smem          : %.2f us / launch\n", time_graph([&](int) { k_empty<<<296, 320, half, s>>>(nullptr); }, N, s));
  printf("tmem alloc 148x320, 225 KB smem     : %.2f us / launch\n", time_graph([&](int) { k_tmem<<<148, 320, big, s>>>(nullptr); }, N, s));
  printf("small 1184x256, no smem             : %.2f us / launch\n", time_graph([&](int) { k_small<<<1184, 256, 0, s>>>(nullptr); }, N, s));
  printf("alternate 225 KB smem / small kernel: %.2f us / launch\n", time_graph([&](int i) { if (i & 1) k_small<<<1184, 256, 0, s>>>(nullptr); else k_empty<<<148, 320, big, s>>>(nullptr); }, N, s));
  printf("alternate 225 KB / 110 KB smem      : %.2f us / launch\n", time_graph([&](int i) { if (i & 1) k_empty<<<148, 320, half, s>>>(nullptr); else k_empty<<<148, 320, big, s>>>(nullptr); }, N, s));
  CK(cudaFuncSetAttribute(k_small, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
  printf("alternate 225 KB / small (carveout 100 on the small one): %.2f us / launch\n", time_graph([&](int i) { if (i & 1) k_small<<<1184, 256, 0, s>>>(nullptr); else k_empty<<<148, 320, big, s>>>(nullptr); }, N, s));
  // programmatic dependent launch between empty kernels
  {
    cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[0].val.programmaticStreamSerializationAllowed = 1;
    auto pdl = [&](int smem) {
      return time_graph([&](int) {
        cudaLaunchConfig_t c = {}; c.gridDim = dim3(148); c.blockDim = dim3(320); c.dynamicSmemBytes = smem; c.stream = s; c.attrs = at; c.numAttrs = 1;
        int* np = nullptr; cudaLaunchKernelEx(&c, k_empty, np);
      }, N, s);
    };
    printf("empty 148x320, 225 KB smem, PDL attr: %.2f us / launch\n", pdl(big));
    printf("empty 148x320, 110 KB smem, PDL attr: %.2f us / launch\n", pdl(half));
  }
  return 0;
}
