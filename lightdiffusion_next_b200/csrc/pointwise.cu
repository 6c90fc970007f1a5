// Small / HBM-bound kernels around the tensor-core path: time embedding, tiny-M linears, conv_in / conv_out,
// nearest 2x upsample, stride-2 patch gather, CFG + solver step, weight ingest.
#include "common.h"
#include "ptx.cuh"

namespace ldn {

__device__ __forceinline__ float silu_p(float x) { return x / (1.f + expf(-x)); }

__device__ __forceinline__ float load_as_f32(const void* src, int dtype, size_t i) {
  if (dtype == 0) return reinterpret_cast<const float*>(src)[i];
  if (dtype == 1) return __half2float(reinterpret_cast<const __half*>(src)[i]);
  return __bfloat162float(reinterpret_cast<const bf16*>(src)[i]);
}

// ------------------------------------------------------------------ timestep embedding
// Reference: ModelSamplingDiscrete.timestep (src/sample/sampling.py:309-320): t = argmin_k |log(sigma) - log_sigmas[k]|;
// timestep_embedding (src/sample/sampling_util.py:56-76): cat(cos(t*f), sin(t*f)), f_i = exp(-ln(1e4) * i / half).
// grid: B blocks of 256 threads.
__global__ void timestep_embed_kernel(const float* __restrict__ sigma, const float* __restrict__ log_sigmas,
                                      int n_sigmas, int dim, float* __restrict__ out, float* __restrict__ t_out) {
  const int b = blockIdx.x;
  __shared__ float s_val[256];
  __shared__ int s_idx[256];
  const float ls = logf(sigma[b]);
  float best = INFINITY;
  int besti = 0x7fffffff;
  for (int k = threadIdx.x; k < n_sigmas; k += blockDim.x) {
    const float d = fabsf(ls - log_sigmas[k]);
    if (d < best) {  // strided ascending k per thread: first minimum kept
      best = d;
      besti = k;
    }
  }
  s_val[threadIdx.x] = best;
  s_idx[threadIdx.x] = besti;
  __syncthreads();
  for (int o = blockDim.x / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) {
      const float v2 = s_val[threadIdx.x + o];
      const int i2 = s_idx[threadIdx.x + o];
      if (v2 < s_val[threadIdx.x] || (v2 == s_val[threadIdx.x] && i2 < s_idx[threadIdx.x])) {
        s_val[threadIdx.x] = v2;
        s_idx[threadIdx.x] = i2;
      }
    }
    __syncthreads();
  }
  const float t = (float)s_idx[0];
  if (threadIdx.x == 0 && t_out) t_out[b] = t;
  const int half = dim / 2;
  for (int i = threadIdx.x; i < half; i += blockDim.x) {
    const float f = expf(-9.210340371976184f * (float)i / (float)half);  // ln(10000)
    const float a = t * f;
    out[(size_t)b * dim + i] = cosf(a);
    out[(size_t)b * dim + half + i] = sinf(a);
  }
}

void launch_timestep_embed(const float* sigma, int Bn, const float* log_sigmas, int n_sigmas, int dim, float* out,
                           float* t_index_out, cudaStream_t stream) {
  timestep_embed_kernel<<<Bn, 256, 0, stream>>>(sigma, log_sigmas, n_sigmas, dim, out, t_index_out);
  LDN_CUDA(cudaGetLastError());
}

// ------------------------------------------------------------------ tiny-M linear (time-embedding MLP, the 22 emb_layers)
// HBM-bound on the weight matrix: the (optionally SiLU'd) activations are staged once per block in shared memory and every
// warp streams whole weight rows with all of a row's 16-byte loads in flight before the first use.
template <int KV, int NB>  // KV: 16-byte weight vectors per lane and row (K <= 256 * KV); NB: activation rows per pass
__global__ void small_linear_kernel(const float* __restrict__ x, int Bn, int K, const bf16* __restrict__ W,
                                    const float* __restrict__ bias, int N, int silu_in, int silu_out,
                                    float* __restrict__ out, long long ldw) {
  extern __shared__ float s_x[];  // [NB][K]
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int warps = blockDim.x >> 5;
  for (int b0 = 0; b0 < Bn; b0 += NB) {
    const int nb = min(NB, Bn - b0);
    __syncthreads();
    for (int i = threadIdx.x; i < NB * K; i += blockDim.x) {
      const int r = NB == 1 ? 0 : i / K;
      float v = 0.f;
      if (r < nb) {
        v = x[(size_t)(b0 + r) * K + (i - r * K)];
        if (silu_in) v = silu_p(v);
      }
      s_x[i] = v;
    }
    __syncthreads();
    for (int n = blockIdx.x * warps + warp; n < N; n += gridDim.x * warps) {
      const bf16* w = W + (size_t)n * ldw;
      uint4 u[KV];
#pragma unroll
      for (int j = 0; j < KV; ++j) {
        const int k = lane * 8 + j * 256;
        u[j] = (k < K) ? __ldg(reinterpret_cast<const uint4*>(w + k)) : make_uint4(0, 0, 0, 0);
      }
      float acc[NB];
#pragma unroll
      for (int r = 0; r < NB; ++r) acc[r] = 0.f;
#pragma unroll
      for (int j = 0; j < KV; ++j) {
        const int k = lane * 8 + j * 256;
        if (k < K) {
          const uint32_t ww[4] = {u[j].x, u[j].y, u[j].z, u[j].w};
          float wf[8];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            wf[2 * i] = bf16_lo(ww[i]);
            wf[2 * i + 1] = bf16_hi(ww[i]);
          }
#pragma unroll
          for (int r = 0; r < NB; ++r) {  // rows past nb hold zeros in shared memory: no branch needed
            const float4 x0 = *reinterpret_cast<const float4*>(s_x + r * K + k);
            const float4 x1 = *reinterpret_cast<const float4*>(s_x + r * K + k + 4);
            acc[r] = fmaf(x0.x, wf[0], acc[r]); acc[r] = fmaf(x0.y, wf[1], acc[r]);
            acc[r] = fmaf(x0.z, wf[2], acc[r]); acc[r] = fmaf(x0.w, wf[3], acc[r]);
            acc[r] = fmaf(x1.x, wf[4], acc[r]); acc[r] = fmaf(x1.y, wf[5], acc[r]);
            acc[r] = fmaf(x1.z, wf[6], acc[r]); acc[r] = fmaf(x1.w, wf[7], acc[r]);
          }
        }
      }
#pragma unroll
      for (int r = 0; r < NB; ++r) {
        float v = acc[r];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0 && r < nb) {
          v += bias ? bias[n] : 0.f;
          if (silu_out) v = silu_p(v);
          out[(size_t)(b0 + r) * N + n] = v;
        }
      }
    }
  }
}

template <int KV, int NB>
static void launch_small_linear_t(const float* x, int Bn, int K, const bf16* W, const float* bias, int N, bool silu_in,
                                  bool silu_out, float* out, cudaStream_t stream, long long ldw) {
  const int threads = 256, warps = threads / 32;
  const size_t smem = sizeof(float) * NB * K;
  static bool attr = false;
  if (!attr) {
    LDN_CUDA(cudaFuncSetAttribute(small_linear_kernel<KV, NB>, cudaFuncAttributeMaxDynamicSharedMemorySize, 136 * 1024));
    attr = true;
  }
  // enough warps in flight to cover HBM latency with one weight row (K * 2 bytes) per warp, a few rows per warp overall
  int blocks = (N + warps - 1) / warps;
  const int cap = 148 * (smem > 48 * 1024 ? 2 : 6);
  if (blocks > cap) blocks = cap;
  small_linear_kernel<KV, NB><<<blocks, threads, smem, stream>>>(x, Bn, K, W, bias, N, silu_in ? 1 : 0, silu_out ? 1 : 0, out,
                                                                ldw);
  LDN_CUDA(cudaGetLastError());
}

template <int KV>
static void launch_small_linear_kv(const float* x, int Bn, int K, const bf16* W, const float* bias, int N, bool silu_in,
                                   bool silu_out, float* out, cudaStream_t stream, long long ldw) {
  if (Bn == 1) launch_small_linear_t<KV, 1>(x, Bn, K, W, bias, N, silu_in, silu_out, out, stream, ldw);
  else if (Bn == 2) launch_small_linear_t<KV, 2>(x, Bn, K, W, bias, N, silu_in, silu_out, out, stream, ldw);
  else launch_small_linear_t<KV, 8>(x, Bn, K, W, bias, N, silu_in, silu_out, out, stream, ldw);
}

void launch_small_linear(const float* x, int Bn, int K, const bf16* W, const float* bias, int N, bool silu_in,
                         bool silu_out, float* out, cudaStream_t stream, long long ldw) {
  LDN_CHECK(K % 8 == 0 && K <= 4096, "small_linear: K must be a multiple of 8 and at most 4096");
  if (ldw == 0) ldw = K;
  LDN_CHECK(ldw % 8 == 0, "small_linear: weight row stride must be a multiple of 8");
  if (K <= 512) launch_small_linear_kv<2>(x, Bn, K, W, bias, N, silu_in, silu_out, out, stream, ldw);
  else if (K <= 1280) launch_small_linear_kv<5>(x, Bn, K, W, bias, N, silu_in, silu_out, out, stream, ldw);
  else if (K <= 2048) launch_small_linear_kv<8>(x, Bn, K, W, bias, N, silu_in, silu_out, out, stream, ldw);
  else launch_small_linear_kv<16>(x, Bn, K, W, bias, N, silu_in, silu_out, out, stream, ldw);
}

// ------------------------------------------------------------------ conv_in: 3x3, tiny Cin, NCHW fp32 -> NHWC bf16
// Fuses BaseModel.apply_model's input scaling x / sqrt(sigma^2 + 1) (src/sample/sampling.py:29-40).
// Wt: [Cout, 3, 3, Cin] bf16. thread = (4 consecutive pixels of a row, 8 output channels): every weight vector fetched
// from shared memory feeds 32 FMAs (the one-pixel version sat on the shared-memory bandwidth).
__global__ void conv_in_kernel(const float* __restrict__ x, const float* __restrict__ sigma,
                               const bf16* __restrict__ Wt, const float* __restrict__ bias, int B, int H, int W, int Cin,
                               int Cout, bf16* __restrict__ out, int flags, int ldo) {
  extern __shared__ float s_w[];  // transposed: [9*Cin][Cout] so the 8 output channels of a thread are contiguous
  const int K = 9 * Cin;
  for (int i = threadIdx.x; i < Cout * K; i += blockDim.x) {  // consecutive threads -> consecutive o: conflict-free
    const int k = i / Cout, o = i - k * Cout;
    s_w[i] = __bfloat162float(Wt[(size_t)o * K + k]);
  }
  __syncthreads();
  const int groups = Cout / 8;
  const int wq = (W + 3) / 4;  // pixel quads per row
  const size_t total = (size_t)B * H * wq * groups;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const int g = (int)(idx % groups);
    const size_t quad = idx / groups;
    const int x0 = (int)(quad % wq) * 4;
    const int y = (int)((quad / wq) % H);
    const int b = (int)(quad / ((size_t)wq * H));
    const float sg = sigma ? sigma[b] : 0.f;
    const float scale = sigma ? rsqrtf(sg * sg + 1.f) : 1.f;
    float acc[4][8];
#pragma unroll
    for (int px = 0; px < 4; ++px)
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[px][i] = bias ? bias[g * 8 + i] : 0.f;
    for (int ky = 0; ky < 3; ++ky) {
      const int yy = y + ky - 1;
      if (yy < 0 || yy >= H) continue;
      for (int c = 0; c < Cin; ++c) {
        const float* row = x + (((size_t)b * Cin + c) * H + yy) * W;
        float in[6];
#pragma unroll
        for (int i = 0; i < 6; ++i) {
          const int xx = x0 + i - 1;
          float v = (xx >= 0 && xx < W) ? row[xx] * scale : 0.f;
          if (flags & 1) v = tanhf(v * (1.0f / 3.0f)) * 3.0f;
          in[i] = v;
        }
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          const float4* wr = reinterpret_cast<const float4*>(s_w + ((ky * 3 + kx) * Cin + c) * Cout + g * 8);
          const float4 w0 = wr[0], w1 = wr[1];
#pragma unroll
          for (int px = 0; px < 4; ++px) {
            const float v = in[px + kx];
            acc[px][0] = fmaf(v, w0.x, acc[px][0]); acc[px][1] = fmaf(v, w0.y, acc[px][1]);
            acc[px][2] = fmaf(v, w0.z, acc[px][2]); acc[px][3] = fmaf(v, w0.w, acc[px][3]);
            acc[px][4] = fmaf(v, w1.x, acc[px][4]); acc[px][5] = fmaf(v, w1.y, acc[px][5]);
            acc[px][6] = fmaf(v, w1.z, acc[px][6]); acc[px][7] = fmaf(v, w1.w, acc[px][7]);
          }
        }
      }
    }
#pragma unroll
    for (int px = 0; px < 4; ++px) {
      if (x0 + px < W) {
        if (flags & 2) {
#pragma unroll
          for (int i = 0; i < 8; ++i) acc[px][i] = fmaxf(acc[px][i], 0.f);
        }
        uint4 ov;
        ov.x = pack_bf16x2(acc[px][0], acc[px][1]);
        ov.y = pack_bf16x2(acc[px][2], acc[px][3]);
        ov.z = pack_bf16x2(acc[px][4], acc[px][5]);
        ov.w = pack_bf16x2(acc[px][6], acc[px][7]);
        *reinterpret_cast<uint4*>(out + (((size_t)b * H + y) * W + x0 + px) * ldo + g * 8) = ov;
      }
    }
  }
}

void launch_conv_in(const float* x, const float* sigma, const bf16* Wt, const float* bias, int B, int H, int W,
                    int Cin, int Cout, bf16* out, cudaStream_t stream, int flags, int ldo) {
  LDN_CHECK(Cout % 8 == 0, "conv_in: Cout must be a multiple of 8");
  if (ldo == 0) ldo = Cout;
  const size_t smem = sizeof(float) * Cout * 9 * Cin;
  LDN_CHECK(smem <= 200 * 1024, "conv_in: weights do not fit in shared memory");
  static bool attr = false;
  if (!attr) {
    LDN_CUDA(cudaFuncSetAttribute(conv_in_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr = true;
  }
  const size_t total = (size_t)B * H * ((W + 3) / 4) * (Cout / 8);
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 2) blocks = 148 * 2;
  conv_in_kernel<<<blocks, 256, smem, stream>>>(x, sigma, Wt, bias, B, H, W, Cin, Cout, out, flags, ldo);
  LDN_CUDA(cudaGetLastError());
}

// ------------------------------------------------------------------ conv_out: 3x3, Cin -> tiny Cout, one warp per pixel
// Input is the already normalised+SiLU'd NHWC bf16 tensor. Fuses denoised = x - eps * sigma
// (EPS.calculate_denoised, src/sample/sampling.py:42-56) and writes NCHW fp32.
template <int COUT>
__global__ void conv_out_kernel(const bf16* __restrict__ h, const bf16* __restrict__ Wt, const float* __restrict__ bias,
                                const float* __restrict__ x, const float* __restrict__ sigma, int B, int H, int W,
                                int Cin, float* __restrict__ denoised, float* __restrict__ eps_out) {
  extern __shared__ float s_w[];  // [COUT][9*Cin]
  const int K = 9 * Cin;
  for (int i = threadIdx.x; i < COUT * K; i += blockDim.x) s_w[i] = __bfloat162float(Wt[i]);
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int warps_per_block = blockDim.x >> 5;
  const size_t npix = (size_t)B * H * W;
  const int nvec = Cin >> 3;
  for (size_t pix = (size_t)blockIdx.x * warps_per_block + (threadIdx.x >> 5); pix < npix;
       pix += (size_t)gridDim.x * warps_per_block) {
    const int xw = (int)(pix % W);
    const int y = (int)((pix / W) % H);
    const int b = (int)(pix / ((size_t)W * H));
    float acc[COUT];
#pragma unroll
    for (int o = 0; o < COUT; ++o) acc[o] = 0.f;
    for (int ky = 0; ky < 3; ++ky) {
      const int yy = y + ky - 1;
      if (yy < 0 || yy >= H) continue;
      for (int kx = 0; kx < 3; ++kx) {
        const int xx = xw + kx - 1;
        if (xx < 0 || xx >= W) continue;
        const bf16* src = h + (((size_t)b * H + yy) * W + xx) * Cin;
        const int kbase = (ky * 3 + kx) * Cin;
        for (int v = lane; v < nvec; v += 32) {
          const uint4 u = *reinterpret_cast<const uint4*>(src + v * 8);
          const uint32_t w[4] = {u.x, u.y, u.z, u.w};
          float f[8];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            f[2 * i] = bf16_lo(w[i]);
            f[2 * i + 1] = bf16_hi(w[i]);
          }
#pragma unroll
          for (int o = 0; o < COUT; ++o) {
            const float* wr = s_w + o * K + kbase + v * 8;
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[o] = fmaf(f[i], wr[i], acc[o]);
          }
        }
      }
    }
#pragma unroll
    for (int o = 0; o < COUT; ++o) {
#pragma unroll
      for (int s = 16; s > 0; s >>= 1) acc[o] += __shfl_xor_sync(0xffffffffu, acc[o], s);
    }
    if (lane == 0) {
#pragma unroll
      for (int o = 0; o < COUT; ++o) {
        const float e = acc[o] + (bias ? bias[o] : 0.f);
        const size_t oi = (((size_t)b * COUT + o) * H + y) * W + xw;
        if (eps_out) eps_out[oi] = e;
        if (denoised) denoised[oi] = x[oi] - e * sigma[b];
      }
    }
  }
}

void launch_conv_out(const bf16* h, const bf16* Wt, const float* bias, const float* x, const float* sigma, int B,
                     int H, int W, int Cin, int Cout, float* denoised, float* eps_out, cudaStream_t stream) {
  LDN_CHECK(Cin % 8 == 0, "conv_out: Cin must be a multiple of 8");
  const size_t smem = sizeof(float) * Cout * 9 * Cin;
  LDN_CHECK(smem <= 200 * 1024, "conv_out: weights do not fit in shared memory");
  const int blocks = 148 * 2;
  if (Cout == 4) {
    static bool attr = false;
    if (!attr) {
      LDN_CUDA(cudaFuncSetAttribute(conv_out_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      attr = true;
    }
    conv_out_kernel<4><<<blocks, 512, smem, stream>>>(h, Wt, bias, x, sigma, B, H, W, Cin, denoised, eps_out);
  } else if (Cout == 3) {
    static bool attr = false;
    if (!attr) {
      LDN_CUDA(cudaFuncSetAttribute(conv_out_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      attr = true;
    }
    conv_out_kernel<3><<<blocks, 512, smem, stream>>>(h, Wt, bias, x, sigma, B, H, W, Cin, denoised, eps_out);
  } else {
    LDN_CHECK(false, "conv_out: only Cout 3 or 4");
  }
  LDN_CUDA(cudaGetLastError());
}

// Second half of the UNet output head when the 3x3 conv itself runs on the tensor cores (implicit GEMM, Cout padded to 16
// columns, fp32 [pixels, 16] scratch): eps = acc + bias; denoised = x - eps * sigma (EPS.calculate_denoised,
// src/sample/sampling.py:42-56), written as NCHW fp32.
__global__ void conv_out_finish_kernel(const float* __restrict__ acc16, const float* __restrict__ bias,
                                       const float* __restrict__ x, const float* __restrict__ sigma, int B, int HW,
                                       int cout, float* __restrict__ denoised) {
  const size_t total = (size_t)B * cout * HW;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int pix = (int)(i % HW);
    const int o = (int)((i / HW) % cout);
    const int b = (int)(i / ((size_t)HW * cout));
    const float e = acc16[((size_t)b * HW + pix) * 16 + o] + (bias ? bias[o] : 0.f);
    denoised[i] = x[i] - e * sigma[b];
  }
}
void launch_conv_out_finish(const float* acc16, const float* bias, const float* x, const float* sigma, int B, int HW,
                            int cout, float* denoised, cudaStream_t stream) {
  const size_t total = (size_t)B * cout * HW;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  conv_out_finish_kernel<<<blocks, 256, 0, stream>>>(acc16, bias, x, sigma, B, HW, cout, denoised);
  LDN_CUDA(cudaGetLastError());
}

// VAE encoder tail: moments[b, o, pix] = bq[o] + sum_i Wq[o, i] * (acc16[pix, i] + bc[i])  -- the encoder's conv_out
// (512 -> 8, run as a 16-column tensor-core conv into fp32 scratch) followed by quant_conv 1x1 (8 -> 8), NCHW fp32 out
// (AutoencodingEngine.encode, src/AutoEncoders/VariationalAE.py:148-172).
__global__ void vae_moments_finish_kernel(const float* __restrict__ acc16, const float* __restrict__ bc,
                                          const float* __restrict__ Wq, const float* __restrict__ bq, int B, int HW, int zc2,
                                          float* __restrict__ out) {
  const size_t total = (size_t)B * zc2 * HW;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int pix = (int)(i % HW);
    const int o = (int)((i / HW) % zc2);
    const int b = (int)(i / ((size_t)HW * zc2));
    const float* a = acc16 + ((size_t)b * HW + pix) * 16;
    float acc = bq[o];
    for (int k = 0; k < zc2; ++k) acc = fmaf(Wq[o * zc2 + k], a[k] + bc[k], acc);
    out[i] = acc;
  }
}
void launch_vae_moments_finish(const float* acc16, const float* bc, const float* Wq, const float* bq, int B, int HW,
                               int zc2, float* out, cudaStream_t stream) {
  LDN_CHECK(zc2 <= 16, "vae_moments_finish: at most 16 channels");
  const size_t total = (size_t)B * zc2 * HW;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  if (blocks < 1) blocks = 1;
  vae_moments_finish_kernel<<<blocks, 256, 0, stream>>>(acc16, bc, Wq, bq, B, HW, zc2, out);
  LDN_CUDA(cudaGetLastError());
}

// VAE decoder tail: rgb[pix, o] = clamp((acc16[pix, o] + bias[o] + 1) / 2, 0, 1), o < 3 -- the decoder's conv_out (128 -> 3,
// run as a 16-column tensor-core conv into fp32 scratch) + process_output (VariationalAE.py:602-604), NHWC fp32 out.
__global__ void vae_rgb_finish_kernel(const float* __restrict__ acc16, const float* __restrict__ bias, size_t npix,
                                      int cout, float* __restrict__ out, int raw) {
  for (size_t pix = (size_t)blockIdx.x * blockDim.x + threadIdx.x; pix < npix; pix += (size_t)gridDim.x * blockDim.x) {
    const float4 a = *reinterpret_cast<const float4*>(acc16 + pix * 16);
    const float v[4] = {a.x, a.y, a.z, a.w};
    for (int o = 0; o < cout; ++o) {
      const float e = v[o] + bias[o];
      out[pix * cout + o] = raw ? e : fminf(fmaxf((e + 1.0f) * 0.5f, 0.f), 1.f);
    }
  }
}
void launch_vae_rgb_finish(const float* acc16, const float* bias, size_t npix, int cout, float* out, cudaStream_t stream,
                           int raw) {
  LDN_CHECK(cout <= 4, "vae_rgb_finish: at most 4 channels");
  int blocks = (int)((npix + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  vae_rgb_finish_kernel<<<blocks, 256, 0, stream>>>(acc16, bias, npix, cout, out, raw);
  LDN_CUDA(cudaGetLastError());
}

// ------------------------------------------------------------------ nearest 2x upsample (F.interpolate nearest, ResBlock.py:135)
__global__ void upsample2x_kernel(const bf16* __restrict__ x, int B, int H, int W, int C, bf16* __restrict__ out) {
  const int nvec = C >> 3;
  const size_t total = (size_t)B * (2 * H) * (2 * W) * nvec;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const int v = (int)(idx % nvec);
    size_t pix = idx / nvec;
    const int ox = (int)(pix % (2 * W));
    const int oy = (int)((pix / (2 * W)) % (2 * H));
    const int b = (int)(pix / ((size_t)4 * W * H));
    const uint4 u = *reinterpret_cast<const uint4*>(x + (((size_t)b * H + (oy >> 1)) * W + (ox >> 1)) * C + v * 8);
    *reinterpret_cast<uint4*>(out + pix * C + v * 8) = u;
  }
}
void launch_upsample2x(const bf16* x, int B, int H, int W, int C, bf16* out, cudaStream_t stream) {
  LDN_CHECK(C % 8 == 0, "upsample: C % 8");
  const size_t total = (size_t)B * 4 * H * W * (C / 8);
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  upsample2x_kernel<<<blocks, 256, 0, stream>>>(x, B, H, W, C, out);
  LDN_CUDA(cudaGetLastError());
}

// ------------------------------------------------------------------ stride-2 3x3 patch gather
// pad_before = 1: UNet Downsample1 (conv stride 2, padding 1; ResBlock.py:173-182), Ho = ceil(H/2).
// pad_before = 0: VAE encoder Downsample (F.pad (0,1,0,1) then conv stride 2, padding 0; VariationalAE.py:224-254), Ho = floor(H/2).
// out[(b,oy,ox), tap*C + c] = x[b, 2oy+ky-pad_before, 2ox+kx-pad_before, c]  (zero outside the image)
__global__ void im2col_s2_kernel(const bf16* __restrict__ x, int B, int H, int W, int C, int Ho, int Wo, int pad_before,
                                 bf16* __restrict__ out) {
  const int nvec = C >> 3;
  const size_t total = (size_t)B * Ho * Wo * 9 * nvec;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const int v = (int)(idx % nvec);
    size_t r = idx / nvec;
    const int tap = (int)(r % 9);
    r /= 9;
    const int ox = (int)(r % Wo);
    const int oy = (int)((r / Wo) % Ho);
    const int b = (int)(r / ((size_t)Wo * Ho));
    const int yy = 2 * oy + tap / 3 - pad_before;
    const int xx = 2 * ox + tap % 3 - pad_before;
    uint4 u = make_uint4(0, 0, 0, 0);
    if (yy >= 0 && yy < H && xx >= 0 && xx < W)
      u = *reinterpret_cast<const uint4*>(x + (((size_t)b * H + yy) * W + xx) * C + v * 8);
    *reinterpret_cast<uint4*>(out + (r * 9 + tap) * C + v * 8) = u;
  }
}
void launch_im2col_s2(const bf16* x, int B, int H, int W, int C, bf16* out, cudaStream_t stream, int pad_before) {
  LDN_CHECK(C % 8 == 0 && H >= 2 && W >= 2, "im2col_s2: bad shape");
  const int Ho = pad_before ? (H + 1) / 2 : H / 2, Wo = pad_before ? (W + 1) / 2 : W / 2;
  const size_t total = (size_t)B * Ho * Wo * 9 * (C / 8);
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  im2col_s2_kernel<<<blocks, 256, 0, stream>>>(x, B, H, W, C, Ho, Wo, pad_before, out);
  LDN_CUDA(cudaGetLastError());
}

// ------------------------------------------------------------------ fills / conversions (weight ingest)
__global__ void fill_bf16_kernel(bf16* p, size_t n, float v) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    p[i] = __float2bfloat16(v);
}
void launch_fill_bf16(bf16* p, size_t n, float v, cudaStream_t stream) {
  if (n == 0) return;
  int blocks = (int)((n + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  fill_bf16_kernel<<<blocks, 256, 0, stream>>>(p, n, v);
  LDN_CUDA(cudaGetLastError());
}

__global__ void to_bf16_kernel(const void* src, int dtype, size_t n, bf16* dst) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    dst[i] = __float2bfloat16(load_as_f32(src, dtype, i));
}
__global__ void to_f32_kernel(const void* src, int dtype, size_t n, float* dst) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    dst[i] = load_as_f32(src, dtype, i);
}
void launch_convert_to_bf16(const void* src, int src_dtype, size_t n, bf16* dst, cudaStream_t stream) {
  if (n == 0) return;
  int blocks = (int)((n + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  to_bf16_kernel<<<blocks, 256, 0, stream>>>(src, src_dtype, n, dst);
  LDN_CUDA(cudaGetLastError());
}
void launch_convert_to_f32(const void* src, int src_dtype, size_t n, float* dst, cudaStream_t stream) {
  if (n == 0) return;
  int blocks = (int)((n + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  to_f32_kernel<<<blocks, 256, 0, stream>>>(src, src_dtype, n, dst);
  LDN_CUDA(cudaGetLastError());
}

// OIHW -> O,kh,kw,I
__global__ void repack_conv_kernel(const void* src, int dtype, int O, int I, int kh, int kw, bf16* dst) {
  const size_t n = (size_t)O * I * kh * kw;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    // i indexes dst: ((o*kh + y)*kw + x)*I + c
    const int c = (int)(i % I);
    size_t r = i / I;
    const int xk = (int)(r % kw);
    r /= kw;
    const int yk = (int)(r % kh);
    const int o = (int)(r / kh);
    dst[i] = __float2bfloat16(load_as_f32(src, dtype, (((size_t)o * I + c) * kh + yk) * kw + xk));
  }
}
void launch_repack_conv_weight(const void* src, int src_dtype, int O, int I, int kh, int kw, bf16* dst,
                               cudaStream_t stream) {
  const size_t n = (size_t)O * I * kh * kw;
  int blocks = (int)((n + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  repack_conv_kernel<<<blocks, 256, 0, stream>>>(src, src_dtype, O, I, kh, kw, dst);
  LDN_CUDA(cudaGetLastError());
}

// ------------------------------------------------------------------ CFG combine + solver update
// denoised = uncond + (cond - uncond) * cfg                     (torch.lerp, src/sample/CFG.py:55-60)
// mode 0: x' = c0 * x - c1 * denoised                           (dpmpp_2m_cfgpp as executed: samplers.py:952-953,
//                                                                 c0 = sigma_next/sigma, c1 = expm1(-h))
// mode 1: d = (x - denoised) / c2 ; x' = x + d * c0 + noise*c1  (euler ancestral: samplers.py:728-732,
//                                                                 c2 = sigma, c0 = sigma_down - sigma, c1 = sigma_up)
// mode 2: denoised only
__global__ void cfg_step_kernel(const float* __restrict__ x, const float* __restrict__ du, const float* __restrict__ dc,
                                float cfg, int mode, float c0, float c1, float c2, const float* __restrict__ noise,
                                float* __restrict__ x_out, float* __restrict__ den_out, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float u = du[i], c = dc[i];
    // torch.lerp(start=u, end=c, weight=cfg): weight >= 0.5 uses end - (end-start)*(1-w)
    const float diff = c - u;
    const float den = (fabsf(cfg) < 0.5f) ? (u + cfg * diff) : (c - diff * (1.f - cfg));
    if (den_out) den_out[i] = den;
    if (mode == 0) {
      x_out[i] = c0 * x[i] - c1 * den;
    } else if (mode == 1) {
      const float xv = x[i];
      const float d = (xv - den) / c2;
      float r = xv + d * c0;
      if (noise) r += noise[i] * c1;
      x_out[i] = r;
    }
  }
}
void launch_cfg_step(const float* x, const float* den_uncond, const float* den_cond, float cfg, int mode, float c0,
                     float c1, float c2, const float* noise, float* x_out, float* denoised_out, size_t n,
                     cudaStream_t stream) {
  if (n == 0) return;
  int blocks = (int)((n + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  cfg_step_kernel<<<blocks, 256, 0, stream>>>(x, den_uncond, den_cond, cfg, mode, c0, c1, c2, noise, x_out,
                                              denoised_out, n);
  LDN_CUDA(cudaGetLastError());
}

// ------------------------------------------------------------------ bilinear resample (multiscale steps)
// F.interpolate(x, size=(oh, ow), mode="bilinear", align_corners=False) on fp32 NCHW planes, the call the reference makes
// around its half-resolution sampler steps (src/sample/samplers.py:821-835).  Same arithmetic as ATen's
// upsample_bilinear2d: source index = (in / out) * (dst + 0.5) - 0.5 clamped at 0, truncated; the second tap is clamped to
// the last row / column; value = h0 * (w0 * a + w1 * b) + h1 * (w0 * c + w1 * d).
__global__ void resample_bilinear_kernel(const float* __restrict__ src, float* __restrict__ dst, int planes, int h, int w,
                                         int oh, int ow, float rh, float rw) {
  const size_t total = (size_t)planes * oh * ow;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int ox = (int)(i % ow);
    const size_t t = i / ow;
    const int oy = (int)(t % oh);
    const size_t pl = t / oh;
    const float sy = fmaxf(rh * (oy + 0.5f) - 0.5f, 0.f);
    const float sx = fmaxf(rw * (ox + 0.5f) - 0.5f, 0.f);
    const int y0 = (int)sy, x0 = (int)sx;
    const int yp = y0 < h - 1 ? 1 : 0, xp = x0 < w - 1 ? 1 : 0;
    const float h1 = sy - y0, h0 = 1.f - h1, w1 = sx - x0, w0 = 1.f - w1;
    const float* s = src + pl * (size_t)h * w + (size_t)y0 * w + x0;
    dst[i] = h0 * (w0 * s[0] + w1 * s[xp]) + h1 * (w0 * s[(size_t)yp * w] + w1 * s[(size_t)yp * w + xp]);
  }
}
void launch_resample_bilinear(const float* src, float* dst, int planes, int h, int w, int oh, int ow,
                              cudaStream_t stream) {
  const size_t total = (size_t)planes * oh * ow;
  if (total == 0) return;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  resample_bilinear_kernel<<<blocks, 256, 0, stream>>>(src, dst, planes, h, w, oh, ow, (float)h / (float)oh,
                                                       (float)w / (float)ow);
  LDN_CUDA(cudaGetLastError());
}

// ------------------------------------------------------------------ bislerp (LatentUpscale of the HiresFix branch)
// One separable pass of the reference's `bislerp` (src/Utilities/upscale.py:5-128): along one axis the two taps of bilinear
// resampling (positions and ratio derived exactly as the reference derives them: bilinear interpolation of the index
// ramps, floor / truncation) are blended SPHERICALLY as C-vectors -- directions slerped, norms blended linearly; nearly
// parallel vectors take the first tap, nearly opposite ones a linear blend.  One thread per output position of a line.
// Strides in elements: element (line, c, i) lives at base(line) + c * stride_c + i * stride_i.
__global__ void bislerp_pass_kernel(const float* __restrict__ src, float* __restrict__ dst, int lines, int inner, int C,
                                    int L_in, int L_out, long long in_outer, long long in_inner, long long in_c, long long in_i,
                                    long long out_outer, long long out_inner, long long out_c, long long out_i) {
  const long long total = (long long)lines * L_out;
  const float scale = (float)L_in / (float)L_out;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int X = (int)(idx % L_out);
    const int line = (int)(idx / L_out);
    const int lo_ = line % inner, ou_ = line / inner;
    // taps: F.interpolate(arange(L_in), size=L_out, mode="bilinear") and the same for the ramp shifted by one (last entry clamped)
    float sp = scale * ((float)X + 0.5f) - 0.5f;
    sp = sp < 0.f ? 0.f : sp;
    const int i0 = (int)sp;
    const int ip = i0 < L_in - 1 ? 1 : 0;
    const float l1 = sp - (float)i0, l0 = 1.f - l1;
    const float v1 = __fadd_rn(__fmul_rn(l0, (float)i0), __fmul_rn(l1, (float)(i0 + ip)));
    const float r = v1 - floorf(v1);
    const int c1 = (int)v1;
    const float h0 = (float)min(i0 + 1, L_in - 1), h1 = (float)min(i0 + ip + 1, L_in - 1);
    const int c2 = (int)__fadd_rn(__fmul_rn(l0, h0), __fmul_rn(l1, h1));
    const float* a = src + ou_ * in_outer + lo_ * in_inner + (long long)c1 * in_i;
    const float* b = src + ou_ * in_outer + lo_ * in_inner + (long long)c2 * in_i;
    float na = 0.f, nb = 0.f;
    for (int c = 0; c < C; ++c) {
      const float x = a[c * in_c], y = b[c * in_c];
      na = fmaf(x, x, na);
      nb = fmaf(y, y, nb);
    }
    na = sqrtf(na);
    nb = sqrtf(nb);
    float dot = 0.f;
    for (int c = 0; c < C; ++c) {
      const float ua = na == 0.f ? 0.f : a[c * in_c] / na;
      const float ub = nb == 0.f ? 0.f : b[c * in_c] / nb;
      dot = fmaf(ua, ub, dot);
    }
    const float omega = acosf(dot);
    const float so = sinf(omega);
    const float wa = sinf((1.f - r) * omega) / so, wb = sinf(r * omega) / so;
    const float nrm = na * (1.f - r) + nb * r;
    float* o = dst + ou_ * out_outer + lo_ * out_inner + (long long)X * out_i;
    for (int c = 0; c < C; ++c) {
      const float x = a[c * in_c], y = b[c * in_c];
      const float ua = na == 0.f ? 0.f : x / na;
      const float ub = nb == 0.f ? 0.f : y / nb;
      float res = (wa * ua + wb * ub) * nrm;
      if (dot > 1.f - 1e-5f) res = x;
      if (dot < 1e-5f - 1.f) res = x * (1.f - r) + y * r;
      o[c * out_c] = res;
    }
  }
}
// src [n, c, h, w] fp32 -> dst [n, c, H, W] fp32; tmp holds n * c * h * W floats (width pass first, as the reference).
void launch_bislerp(const float* src, float* tmp, float* dst, int n, int c, int h, int w, int H, int W, cudaStream_t stream) {
  {
    const long long total = (long long)n * h * W;
    int blocks = (int)((total + 255) / 256);
    if (blocks > 148 * 8) blocks = 148 * 8;
    bislerp_pass_kernel<<<blocks, 256, 0, stream>>>(src, tmp, n * h, h, c, w, W, (long long)c * h * w, w, (long long)h * w, 1,
                                                    (long long)c * h * W, W, (long long)h * W, 1);
    LDN_CUDA(cudaGetLastError());
  }
  {
    const long long total = (long long)n * W * H;
    int blocks = (int)((total + 255) / 256);
    if (blocks > 148 * 8) blocks = 148 * 8;
    bislerp_pass_kernel<<<blocks, 256, 0, stream>>>(tmp, dst, n * W, W, c, h, H, (long long)c * h * W, 1, (long long)h * W, W,
                                                    (long long)c * H * W, 1, (long long)H * W, W);
    LDN_CUDA(cudaGetLastError());
  }
}

// ------------------------------------------------------------------ row softmax (VAE mid attention, materialised scores)
// out[r, :] = softmax(in[r, :] * scale); one block per row.
__global__ void softmax_rows_kernel(const bf16* __restrict__ in, long long ld_in, bf16* __restrict__ out,
                                    long long ld_out, int cols, float scale) {
  const bf16* src = in + (size_t)blockIdx.x * ld_in;
  bf16* dst = out + (size_t)blockIdx.x * ld_out;
  __shared__ float red[32];
  float mx = -INFINITY;
  for (int c = threadIdx.x; c < cols; c += blockDim.x) mx = fmaxf(mx, __bfloat162float(src[c]));
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
  __syncthreads();
  mx = red[0];
  for (int w = 1; w < (int)(blockDim.x >> 5); ++w) mx = fmaxf(mx, red[w]);
  __syncthreads();
  float sum = 0.f;
  for (int c = threadIdx.x; c < cols; c += blockDim.x) sum += __expf((__bfloat162float(src[c]) - mx) * scale);
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = sum;
  __syncthreads();
  sum = 0.f;
  for (int w = 0; w < (int)(blockDim.x >> 5); ++w) sum += red[w];
  const float inv = 1.f / sum;
  for (int c = threadIdx.x; c < cols; c += blockDim.x)
    dst[c] = __float2bfloat16(__expf((__bfloat162float(src[c]) - mx) * scale) * inv);
}
// Register-resident variant: one block per row, the row is read once with 16-byte loads (MAXV per thread), reduced, and
// written once -- 2 bytes read + 2 bytes written per element instead of three scalar passes.
template <int MAXV, typename TIn>
__global__ void __launch_bounds__(512) softmax_rows_vec_kernel(const TIn* __restrict__ in, long long ld_in,
                                                               bf16* __restrict__ out, long long ld_out, int cols,
                                                               float scale) {
  const TIn* src = in + (size_t)blockIdx.x * ld_in;
  bf16* dst = out + (size_t)blockIdx.x * ld_out;
  __shared__ float red[32];
  const int nvec = cols >> 3;
  const float sl = scale * 1.4426950408889634f;
  float v[MAXV][8];
  float mx = -INFINITY;
#pragma unroll
  for (int k = 0; k < MAXV; ++k) {
    const int vi = threadIdx.x + k * 512;
    if (vi < nvec) {
      if constexpr (sizeof(TIn) == 4) {  // fp32 logits (VAE mid-block attention): two 16-byte loads
        const float4 a = *reinterpret_cast<const float4*>(src + (size_t)vi * 8);
        const float4 b = *reinterpret_cast<const float4*>(src + (size_t)vi * 8 + 4);
        v[k][0] = a.x; v[k][1] = a.y; v[k][2] = a.z; v[k][3] = a.w;
        v[k][4] = b.x; v[k][5] = b.y; v[k][6] = b.z; v[k][7] = b.w;
      } else {
        const uint4 u = *reinterpret_cast<const uint4*>(src + (size_t)vi * 8);
        const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          v[k][2 * i] = bf16_lo(w[i]);
          v[k][2 * i + 1] = bf16_hi(w[i]);
        }
      }
#pragma unroll
      for (int i = 0; i < 8; i += 2) mx = fmaxf(mx, fmaxf(v[k][i], v[k][i + 1]));
    }
  }
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
  __syncthreads();
  mx = red[0];
  for (int w = 1; w < 16; ++w) mx = fmaxf(mx, red[w]);
  __syncthreads();
  const float moff = mx * sl;
  float sum = 0.f;
#pragma unroll
  for (int k = 0; k < MAXV; ++k) {
    const int vi = threadIdx.x + k * 512;
    if (vi < nvec) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        v[k][i] = exp2f(fmaf(v[k][i], sl, -moff));
        sum += v[k][i];
      }
    }
  }
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = sum;
  __syncthreads();
  sum = 0.f;
  for (int w = 0; w < 16; ++w) sum += red[w];
  const float inv = 1.f / sum;
#pragma unroll
  for (int k = 0; k < MAXV; ++k) {
    const int vi = threadIdx.x + k * 512;
    if (vi < nvec) {
      uint4 o;
      o.x = pack_bf16x2(v[k][0] * inv, v[k][1] * inv);
      o.y = pack_bf16x2(v[k][2] * inv, v[k][3] * inv);
      o.z = pack_bf16x2(v[k][4] * inv, v[k][5] * inv);
      o.w = pack_bf16x2(v[k][6] * inv, v[k][7] * inv);
      *reinterpret_cast<uint4*>(dst + (size_t)vi * 8) = o;
    }
  }
}
// Scalar fallback for fp32 logits (odd widths): three passes, one block per row.
__global__ void softmax_rows_f32_kernel(const float* __restrict__ in, long long ld_in, bf16* __restrict__ out,
                                        long long ld_out, int cols, float scale) {
  const float* src = in + (size_t)blockIdx.x * ld_in;
  bf16* dst = out + (size_t)blockIdx.x * ld_out;
  __shared__ float red[32];
  float mx = -INFINITY;
  for (int c = threadIdx.x; c < cols; c += blockDim.x) mx = fmaxf(mx, src[c]);
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
  __syncthreads();
  mx = red[0];
  for (int w = 1; w < (int)(blockDim.x >> 5); ++w) mx = fmaxf(mx, red[w]);
  __syncthreads();
  float sum = 0.f;
  for (int c = threadIdx.x; c < cols; c += blockDim.x) sum += __expf((src[c] - mx) * scale);
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = sum;
  __syncthreads();
  sum = 0.f;
  for (int w = 0; w < (int)(blockDim.x >> 5); ++w) sum += red[w];
  const float inv = 1.f / sum;
  for (int c = threadIdx.x; c < cols; c += blockDim.x) dst[c] = __float2bfloat16(__expf((src[c] - mx) * scale) * inv);
}
void launch_softmax_rows_f32(const float* in, long long ld_in, bf16* out, long long ld_out, int rows, int cols, float scale,
                             cudaStream_t stream) {
  const bool vec_ok = cols % 8 == 0 && ld_in % 8 == 0 && ld_out % 8 == 0 && (reinterpret_cast<uintptr_t>(in) & 15) == 0 &&
                      (reinterpret_cast<uintptr_t>(out) & 15) == 0;
  if (vec_ok && cols <= 512 * 8 * 4)
    softmax_rows_vec_kernel<4, float><<<rows, 512, 0, stream>>>(in, ld_in, out, ld_out, cols, scale);
  else if (vec_ok && cols <= 512 * 8 * 16)
    softmax_rows_vec_kernel<16, float><<<rows, 512, 0, stream>>>(in, ld_in, out, ld_out, cols, scale);
  else
    softmax_rows_f32_kernel<<<rows, 512, 0, stream>>>(in, ld_in, out, ld_out, cols, scale);
  LDN_CUDA(cudaGetLastError());
}
void launch_softmax_rows(const bf16* in, long long ld_in, bf16* out, long long ld_out, int rows, int cols, float scale,
                         cudaStream_t stream) {
  const bool vec_ok = cols % 8 == 0 && ld_in % 8 == 0 && ld_out % 8 == 0 && (reinterpret_cast<uintptr_t>(in) & 15) == 0 &&
                      (reinterpret_cast<uintptr_t>(out) & 15) == 0;
  if (vec_ok && cols <= 512 * 8 * 4)
    softmax_rows_vec_kernel<4, bf16><<<rows, 512, 0, stream>>>(in, ld_in, out, ld_out, cols, scale);
  else if (vec_ok && cols <= 512 * 8 * 16)
    softmax_rows_vec_kernel<16, bf16><<<rows, 512, 0, stream>>>(in, ld_in, out, ld_out, cols, scale);
  else
    softmax_rows_kernel<<<rows, 512, 0, stream>>>(in, ld_in, out, ld_out, cols, scale);
  LDN_CUDA(cudaGetLastError());
}

}  // namespace ldn

namespace ldn {

// ------------------------------------------------------------------ VAE conv_out (Cin -> 3) with the image mapping fused
__global__ void conv_out_rgb_kernel(const bf16* __restrict__ h, const bf16* __restrict__ Wt, const float* __restrict__ bias,
                                    int B, int H, int W, int Cin, float* __restrict__ rgb) {
  extern __shared__ float s_wrgb[];  // [3][9*Cin]
  const int K = 9 * Cin;
  for (int i = threadIdx.x; i < 3 * K; i += blockDim.x) s_wrgb[i] = __bfloat162float(Wt[i]);
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  const size_t npix = (size_t)B * H * W;
  const int nvec = Cin >> 3;
  for (size_t pix = (size_t)blockIdx.x * wpb + (threadIdx.x >> 5); pix < npix; pix += (size_t)gridDim.x * wpb) {
    const int xw = (int)(pix % W);
    const int y = (int)((pix / W) % H);
    const int b = (int)(pix / ((size_t)W * H));
    float acc[3] = {0.f, 0.f, 0.f};
    for (int ky = 0; ky < 3; ++ky) {
      const int yy = y + ky - 1;
      if (yy < 0 || yy >= H) continue;
      for (int kx = 0; kx < 3; ++kx) {
        const int xx = xw + kx - 1;
        if (xx < 0 || xx >= W) continue;
        const bf16* src = h + (((size_t)b * H + yy) * W + xx) * Cin;
        const int kbase = (ky * 3 + kx) * Cin;
        for (int v = lane; v < nvec; v += 32) {
          const uint4 u = *reinterpret_cast<const uint4*>(src + v * 8);
          const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float f0 = bf16_lo(w[i]), f1 = bf16_hi(w[i]);
#pragma unroll
            for (int o = 0; o < 3; ++o) {
              const float* wr = s_wrgb + o * K + kbase + v * 8 + 2 * i;
              acc[o] = fmaf(f0, wr[0], fmaf(f1, wr[1], acc[o]));
            }
          }
        }
      }
    }
#pragma unroll
    for (int o = 0; o < 3; ++o)
#pragma unroll
      for (int s = 16; s > 0; s >>= 1) acc[o] += __shfl_xor_sync(0xffffffffu, acc[o], s);
    if (lane == 0) {
#pragma unroll
      for (int o = 0; o < 3; ++o) {
        const float e = acc[o] + (bias ? bias[o] : 0.f);
        rgb[pix * 3 + o] = fminf(fmaxf((e + 1.0f) * 0.5f, 0.f), 1.f);
      }
    }
  }
}
void launch_conv_out_rgb(const bf16* h, const bf16* Wt, const float* bias, int B, int H, int W, int Cin, float* rgb,
                         cudaStream_t stream) {
  LDN_CHECK(Cin % 8 == 0, "conv_out_rgb: Cin must be a multiple of 8");
  const size_t smem = sizeof(float) * 3 * 9 * Cin;
  conv_out_rgb_kernel<<<148 * 4, 512, smem, stream>>>(h, Wt, bias, B, H, W, Cin, rgb);
  LDN_CUDA(cudaGetLastError());
}

__global__ void conv1x1_f32_kernel(const float* __restrict__ x, const float* __restrict__ W, const float* __restrict__ bias,
                                   int B, int Cin, int Cout, int HW, float* __restrict__ y) {
  const size_t total = (size_t)B * Cout * HW;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int p = (int)(i % HW);
    const int o = (int)((i / HW) % Cout);
    const int b = (int)(i / ((size_t)HW * Cout));
    float acc = bias ? bias[o] : 0.f;
    for (int c = 0; c < Cin; ++c) acc = fmaf(W[o * Cin + c], x[((size_t)b * Cin + c) * HW + p], acc);
    y[i] = acc;
  }
}
void launch_conv1x1_f32(const float* x, const float* W, const float* bias, int B, int Cin, int Cout, int HW, float* y,
                        cudaStream_t stream) {
  const size_t total = (size_t)B * Cout * HW;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  conv1x1_f32_kernel<<<blocks, 256, 0, stream>>>(x, W, bias, B, Cin, Cout, HW, y);
  LDN_CUDA(cudaGetLastError());
}

// Token ids in [0, vocab) read the checkpoint's table; ids vocab, vocab + 1, ... read the small textual-inversion table
// (SDClipModel.set_up_textual_embeddings appends them to a temporary Embedding, src/SD15/SDClip.py:247-259). The host
// validates ids before the call; an id that is still out of range reads the last vocabulary row instead of foreign memory.
__global__ void clip_embed_kernel(const long long* __restrict__ ids, const float* __restrict__ tok, int vocab,
                                  const float* __restrict__ extra, const int* __restrict__ extra_n,
                                  const float* __restrict__ pos, int rows, int T, int C, bf16* __restrict__ out) {
  const size_t total = (size_t)rows * C;
  const int n_extra = extra_n ? *extra_n : 0;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const int r = (int)(i / C);
    const long long id = ids[r];
    const float* src;
    if (id >= 0 && id < vocab) src = tok + (size_t)id * C;
    else if (id >= vocab && id - vocab < n_extra) src = extra + (size_t)(id - vocab) * C;
    else src = tok + (size_t)(vocab - 1) * C;
    out[i] = __float2bfloat16(src[c] + pos[(size_t)(r % T) * C + c]);
  }
}
void launch_clip_embed(const long long* ids, const float* tok_emb, int vocab, const float* extra, const int* extra_n,
                       const float* pos_emb, int rows, int T, int C, bf16* out, cudaStream_t stream) {
  const size_t total = (size_t)rows * C;
  int blocks = (int)((total + 255) / 256);
  clip_embed_kernel<<<blocks, 256, 0, stream>>>(ids, tok_emb, vocab, extra, extra_n, pos_emb, rows, T, C, out);
  LDN_CUDA(cudaGetLastError());
}

}  // namespace ldn
