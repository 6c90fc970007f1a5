// Elementwise / normalisation kernels of the Flux.1 DiT step (reference: src/BlackForest/Flux.py).
#include "common.h"
#include "ptx.cuh"

namespace ldn {

// timestep_embedding_flux(t, 256) (src/sample/sampling_util.py:78-104): cat(cos, sin)(1000 t * exp(-ln(1e4) i / 128))
__global__ void flux_temb_kernel(const float* __restrict__ t, float* __restrict__ out) {
  const int b = blockIdx.x, i = threadIdx.x;  // 128 threads
  const float f = expf(-9.210340371976184f * (float)i / 128.0f);
  const float a = 1000.0f * t[b] * f;
  out[b * 256 + i] = cosf(a);
  out[b * 256 + 128 + i] = sinf(a);
}
void launch_flux_temb(const float* t, int B, float* out, cudaStream_t stream) {
  flux_temb_kernel<<<B, 128, 0, stream>>>(t, out);
  LDN_CUDA(cudaGetLastError());
}

__global__ void vec_add3_kernel(const float* __restrict__ a, const float* __restrict__ b, const float* __restrict__ c, int n,
                                float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = a[i] + (b ? b[i] : 0.f) + (c ? c[i] : 0.f);
}
void launch_vec_add3(const float* a, const float* b, const float* c, int n, float* out, cudaStream_t stream) {
  vec_add3_kernel<<<(n + 255) / 256, 256, 0, stream>>>(a, b, c, n, out);
  LDN_CUDA(cudaGetLastError());
}

// Modulated LayerNorm (DoubleStreamBlock / SingleStreamBlock / LastLayer: (1 + scale) * LayerNorm(x) + shift, LayerNorm
// without affine, eps 1e-6; Flux.py:302-303, 334, 391, 468-469). One warp per token row, the row held in registers.
template <int MAXV>
__global__ void modln_kernel(const bf16* __restrict__ x, int rows, int C, const float* __restrict__ shift,
                             const float* __restrict__ scale, bf16* __restrict__ out) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= rows) return;
  const int nvec = C >> 3;
  const bf16* src = x + (size_t)warp * C;
  float v[MAXV][8];
  float sum = 0.f;
#pragma unroll
  for (int k = 0; k < MAXV; ++k) {
    const int vi = lane + k * 32;
    if (vi < nvec) {
      const uint4 u = *reinterpret_cast<const uint4*>(src + vi * 8);
      const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        v[k][2 * i] = bf16_lo(w[i]);
        v[k][2 * i + 1] = bf16_hi(w[i]);
        sum += v[k][2 * i] + v[k][2 * i + 1];
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const float mean = sum / (float)C;
  float sq = 0.f;
#pragma unroll
  for (int k = 0; k < MAXV; ++k) {
    const int vi = lane + k * 32;
    if (vi < nvec) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float d = v[k][i] - mean;
        sq += d * d;
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
  const float rstd = rsqrtf(sq / (float)C + 1e-6f);
  bf16* dst = out + (size_t)warp * C;
#pragma unroll
  for (int k = 0; k < MAXV; ++k) {
    const int vi = lane + k * 32;
    if (vi < nvec) {
      const int c = vi * 8;
      float y[8];
#pragma unroll
      for (int i = 0; i < 8; i += 4) {
        const float4 sc = *reinterpret_cast<const float4*>(scale + c + i);
        const float4 sh = *reinterpret_cast<const float4*>(shift + c + i);
        y[i] = fmaf((v[k][i] - mean) * rstd, 1.0f + sc.x, sh.x);
        y[i + 1] = fmaf((v[k][i + 1] - mean) * rstd, 1.0f + sc.y, sh.y);
        y[i + 2] = fmaf((v[k][i + 2] - mean) * rstd, 1.0f + sc.z, sh.z);
        y[i + 3] = fmaf((v[k][i + 3] - mean) * rstd, 1.0f + sc.w, sh.w);
      }
      *reinterpret_cast<uint4*>(dst + c) =
          make_uint4(pack_bf16x2(y[0], y[1]), pack_bf16x2(y[2], y[3]), pack_bf16x2(y[4], y[5]), pack_bf16x2(y[6], y[7]));
    }
  }
}
void launch_modln(const bf16* x, int rows, int C, const float* shift, const float* scale, bf16* out, cudaStream_t stream) {
  LDN_CHECK(C % 8 == 0 && C <= 8 * 32 * 12, "modln: C must be a multiple of 8 and <= 3072");
  const int threads = 256;
  const int blocks = (rows * 32 + threads - 1) / threads;
  if (C <= 8 * 32 * 2)
    modln_kernel<2><<<blocks, threads, 0, stream>>>(x, rows, C, shift, scale, out);
  else
    modln_kernel<12><<<blocks, threads, 0, stream>>>(x, rows, C, shift, scale, out);
  LDN_CUDA(cudaGetLastError());
}

// QKNorm + RoPE (Flux.py:173-200 RMSNorm over the 128-wide head with a learned scale, eps 1e-6; apply_rope :67-82), in place.
// One warp per (row, q-or-k head): a lane owns 4 consecutive dims = 2 rotation pairs.
__global__ void qk_norm_rope_kernel(bf16* __restrict__ qk, long long ld, int rows, int heads,
                                    const float* __restrict__ q_scale, const float* __restrict__ k_scale,
                                    const float* __restrict__ pe) {
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const long long total = (long long)rows * 2 * heads;
  if (warp >= total) return;
  const int hh = (int)(warp % (2 * heads));  // 0..heads-1: q heads, heads..2heads-1: k heads
  const long long row = warp / (2 * heads);
  bf16* p = qk + row * ld + (long long)hh * 128 + lane * 4;
  const uint2 u = *reinterpret_cast<const uint2*>(p);
  float x0 = bf16_lo(u.x), x1 = bf16_hi(u.x), x2 = bf16_lo(u.y), x3 = bf16_hi(u.y);
  float ss = x0 * x0 + x1 * x1 + x2 * x2 + x3 * x3;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  const float r = rsqrtf(ss * (1.0f / 128.0f) + 1e-6f);
  const float* sc = (hh < heads ? q_scale : k_scale) + lane * 4;
  x0 *= r * sc[0]; x1 *= r * sc[1]; x2 *= r * sc[2]; x3 *= r * sc[3];
  // the reference rounds the normalised q / k to the model dtype before the rotation (QKNorm returns q.to(v))
  x0 = __bfloat162float(__float2bfloat16(x0)); x1 = __bfloat162float(__float2bfloat16(x1));
  x2 = __bfloat162float(__float2bfloat16(x2)); x3 = __bfloat162float(__float2bfloat16(x3));
  const float4 cs = *reinterpret_cast<const float4*>(pe + row * 128 + lane * 4);  // (cos, sin) of pairs 2 lane, 2 lane + 1
  const float y0 = cs.x * x0 - cs.y * x1, y1 = cs.y * x0 + cs.x * x1;
  const float y2 = cs.z * x2 - cs.w * x3, y3 = cs.w * x2 + cs.z * x3;
  *reinterpret_cast<uint2*>(p) = make_uint2(pack_bf16x2(y0, y1), pack_bf16x2(y2, y3));
}
void launch_qk_norm_rope(bf16* qk, long long ld, int rows, int heads, const float* q_scale, const float* k_scale,
                         const float* pe, cudaStream_t stream) {
  const long long total = (long long)rows * 2 * heads;
  const int threads = 256;
  const long long blocks = (total * 32 + threads - 1) / threads;
  qk_norm_rope_kernel<<<(unsigned)blocks, threads, 0, stream>>>(qk, ld, rows, heads, q_scale, k_scale, pe);
  LDN_CUDA(cudaGetLastError());
}

}  // namespace ldn
