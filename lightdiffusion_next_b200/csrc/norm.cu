// GroupNorm(+SiLU) and LayerNorm for token-major (NHWC) bf16 activations. HBM-bound kernels.
//
// Reference call sites replaced: F.group_norm via src/cond/cast.py:241 (ResBlock in/out layers
// src/AutoEncoders/ResBlock.py:252,280 eps 1e-5 + SiLU; SpatialTransformer.norm src/NeuralNetwork/transformer.py:286-293
// eps 1e-6; VAE Normalize eps 1e-6) and F.layer_norm via cast.py:281 (transformer.py:154-157).
//
// GroupNorm input may be the *virtual* channel concat [x0 | x1] of the UNet skip connection
// (torch.cat([h, hs.pop()], 1), src/NeuralNetwork/unet.py:750): groups may straddle the boundary (C=960, 1920),
// which is why statistics are accumulated per channel-vector and folded into groups afterwards.
#include "common.h"
#include "ptx.cuh"

namespace ldn {

// ------------------------------------------------------------------ GroupNorm statistics (deterministic)
// Kernel 1, grid (splits, B), block (C/8)*R threads: every thread owns 8 consecutive channels for a strided set of
// pixels, folds them into at most two (group, sum, sumsq) partials in shared memory, then warp g reduces the partials of
// group g in a fixed order and ADDS them to the (batch, group) accumulator of this GroupNorm instance as 64-bit FIXED-POINT
// integers (sum * 2^16, sum of squares * 2^12): integer addition is associative, so the totals do not depend on the order
// in which the blocks arrive -- bit-deterministic without the arrival counter + memory fence + serial last-block fold the
// first version ended with (a ~3 us tail on a ~6 us kernel).  The apply kernel converts the totals itself.  Accumulators
// live in per-instance slots of the workspace, zeroed once per program execution (one memset node at its start).
// Range: |sum| < 1.4e14, sum of squares < 2.2e15 per (batch, group) -- an RMS of ~67 000 over a 491 520-element group.
// (LDN_GN_SUM_SCALE / LDN_GN_SQ_SCALE: common.h -- the conv epilogues that take the statistics themselves use them too)

__global__ void __launch_bounds__(640, 2) gn_stats_kernel(const bf16* __restrict__ x0, int C0, const bf16* __restrict__ x1, int C1, int HW,
                                int cpg, int rows_per_block, int R, unsigned long long* __restrict__ acc) {
  const int C = C0 + C1;
  const int nvec = C >> 3;
  const bool active = (int)threadIdx.x < nvec * R;  // the block is padded to whole warps
  const int cv = threadIdx.x % nvec;
  const int prow = threadIdx.x / nvec;
  const int b = blockIdx.y;
  __shared__ float s_part[640][4];  // per thread: {sum_lo, sq_lo, sum_hi, sq_hi} for groups g_lo = c/cpg and g_lo+1
  const int c = cv * 8;
  if (active) {
    const bf16* src;
    int ld;
    if (c < C0) {
      src = x0 + (size_t)b * HW * C0 + c;
      ld = C0;
    } else {
      src = x1 + (size_t)b * HW * C1 + (c - C0);
      ld = C1;
    }
    float s[8], q[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) s[i] = q[i] = 0.f;
    const int p_begin = blockIdx.x * rows_per_block;
    const int p_end = min(HW, p_begin + rows_per_block);
    int pix = p_begin + prow;
    for (; pix + 3 * R < p_end; pix += 4 * R) {  // four independent 16-byte loads in flight per thread
      uint4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) v[u] = *reinterpret_cast<const uint4*>(src + (size_t)(pix + u * R) * ld);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const uint32_t w[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float a = bf16_lo(w[i]), bb = bf16_hi(w[i]);
          s[2 * i] += a;
          q[2 * i] += a * a;
          s[2 * i + 1] += bb;
          q[2 * i + 1] += bb * bb;
        }
      }
    }
    for (; pix < p_end; pix += R) {
      const uint4 v = *reinterpret_cast<const uint4*>(src + (size_t)pix * ld);
      const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float a = bf16_lo(w[i]), bb = bf16_hi(w[i]);
        s[2 * i] += a;
        q[2 * i] += a * a;
        s[2 * i + 1] += bb;
        q[2 * i + 1] += bb * bb;
      }
    }
    const int g_hi_begin = (c / cpg + 1) * cpg;  // first channel of the second group this thread touches
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (c + i < g_hi_begin) {
        a0 += s[i];
        a1 += q[i];
      } else {
        a2 += s[i];
        a3 += q[i];
      }
    }
    s_part[threadIdx.x][0] = a0;
    s_part[threadIdx.x][1] = a1;
    s_part[threadIdx.x][2] = a2;
    s_part[threadIdx.x][3] = a3;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  for (int g = warp; g < 32; g += nwarps) {
    // channel vectors that touch group g: cv in [first, last]
    const int first = (g * cpg) >> 3;
    const int last = ((g + 1) * cpg - 1) >> 3;
    const int ncv = last - first + 1;
    double sum = 0.0, sq = 0.0;
    for (int idx = lane; idx < ncv * R; idx += 32) {
      const int v = first + idx % ncv;
      const int pr = idx / ncv;
      const int t = pr * nvec + v;
      const int g_lo = (v * 8) / cpg;
      if (g_lo == g) {
        sum += (double)s_part[t][0];
        sq += (double)s_part[t][1];
      } else if (g_lo + 1 == g) {
        sum += (double)s_part[t][2];
        sq += (double)s_part[t][3];
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      sum += __shfl_xor_sync(0xffffffffu, sum, o);
      sq += __shfl_xor_sync(0xffffffffu, sq, o);
    }
    if (lane == 0) {
      atomicAdd(&acc[(b * 32 + g) * 2], (unsigned long long)__double2ll_rn(sum * LDN_GN_SUM_SCALE));
      atomicAdd(&acc[(b * 32 + g) * 2 + 1], (unsigned long long)__double2ll_rn(sq * LDN_GN_SQ_SCALE));
    }
  }
}

// SiLU with ONE MUFU: x * sigmoid(x) = x * (0.5 + 0.5 * tanh(x / 2)) (tanh.approx: relative error 2^-11, far below the bf16
// rounding of the output).  x / (1 + exp(-x)) costs two (ex2 + rcp): at 16 MUFU results per clock and SM that was 4.5 us of the
// ~11 us the level-0 apply kernel takes.
__device__ __forceinline__ float silu_f(float x) {
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(0.5f * x));
  const float hx = 0.5f * x;
  return fmaf(hx, t, hx);
}

// ------------------------------------------------------------------ GroupNorm apply (+SiLU)
// Same thread layout as the statistics kernel: grid (splits, B), block (C/8)*R threads; a thread owns 8 consecutive
// channels (scale / shift held in registers) and walks a strided set of pixels with four 16-byte loads in flight.
// (two blocks of <= 640 threads per SM: at 54 registers the kernel fitted once and a 256-block grid needed two waves)
__global__ void __launch_bounds__(640, 2) gn_apply_kernel(const bf16* __restrict__ x0, int C0, const bf16* __restrict__ x1, int C1, int HW,
                                int cpg, const float* __restrict__ gamma, const float* __restrict__ beta, int silu,
                                const unsigned long long* __restrict__ acc, float eps, bf16* __restrict__ out,
                                int rows_per_block, int R, double inv_sum, double inv_sq) {
  const int C = C0 + C1;
  const int nvec = C >> 3;
  if ((int)threadIdx.x >= nvec * R) return;  // the block is padded to whole warps
  const int cv = threadIdx.x % nvec;
  const int prow = threadIdx.x / nvec;
  const int b = blockIdx.y;
  const int c = cv * 8;
  // (mean, rstd) of the (at most two) groups this thread's 8 channels touch, from the fixed-point totals.  The totals are
  // converted and scaled in double (two multiplications by host-computed reciprocals and one FMA: E[x^2] - mean^2 needs the
  // width), the rest runs in fp32: rsqrt.approx + one Newton step (~1 ulp).  The first version divided and took the square root
  // in double in every thread -- ~110 FP64 instructions per thread in front of a ~10 us kernel.
  const int g_lo = c / cpg;
  const int g_hi_begin = (g_lo + 1) * cpg;  // first channel of the second group
  float mean2[2], rstd2[2];
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const int g = min(g_lo + k, 31);
    const double mean = (double)(long long)acc[(b * 32 + g) * 2] * inv_sum;
    const double ex2 = (double)(long long)acc[(b * 32 + g) * 2 + 1] * inv_sq;
    const float var = fmaxf((float)fma(-mean, mean, ex2), 0.f) + eps;
    float r = rsqrtf(var);
    r = r * fmaf(-0.5f * var, r * r, 1.5f);
    mean2[k] = (float)mean;
    rstd2[k] = r;
  }
  float a[8], sh[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int k = (c + i < g_hi_begin) ? 0 : 1;
    a[i] = rstd2[k] * gamma[c + i];
    sh[i] = beta[c + i] - mean2[k] * a[i];
  }
  const bf16* src;
  int ld;
  if (c < C0) {
    src = x0 + (size_t)b * HW * C0 + c;
    ld = C0;
  } else {
    src = x1 + (size_t)b * HW * C1 + (c - C0);
    ld = C1;
  }
  bf16* dst = out + (size_t)b * HW * C + c;
  const int p_begin = blockIdx.x * rows_per_block;
  const int p_end = min(HW, p_begin + rows_per_block);
  auto apply = [&](const uint4& v, int pix) {
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
    uint32_t o[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float y0 = fmaf(bf16_lo(w[i]), a[2 * i], sh[2 * i]);
      float y1 = fmaf(bf16_hi(w[i]), a[2 * i + 1], sh[2 * i + 1]);
      if (silu) {
        y0 = silu_f(y0);
        y1 = silu_f(y1);
      }
      o[i] = pack_bf16x2(y0, y1);
    }
    *reinterpret_cast<uint4*>(dst + (size_t)pix * C) = make_uint4(o[0], o[1], o[2], o[3]);
  };
  int pix = p_begin + prow;
  for (; pix + 3 * R < p_end; pix += 4 * R) {
    uint4 v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) v[u] = *reinterpret_cast<const uint4*>(src + (size_t)(pix + u * R) * ld);
#pragma unroll
    for (int u = 0; u < 4; ++u) apply(v[u], pix + u * R);
  }
  for (; pix < p_end; pix += R) apply(*reinterpret_cast<const uint4*>(src + (size_t)pix * ld), pix);
}

void launch_groupnorm(const bf16* x0, int C0, const bf16* x1, int C1, int B, int HW, int groups, float eps,
                      const float* gamma, const float* beta, bool silu, bf16* out, float* stats_ws, int slot,
                      cudaStream_t stream, bool have_stats) {
  const int C = C0 + C1;
  LDN_CHECK(groups == 32, "groupnorm: only 32 groups supported");
  LDN_CHECK(C % 32 == 0 && C % 8 == 0 && C0 % 8 == 0, "groupnorm: channel counts must be multiples of 8/32");
  LDN_CHECK(C / 8 <= 640, "groupnorm: too many channels");
  LDN_CHECK(slot >= 0 && slot < LDN_GN_SLOTS, "groupnorm: statistics slot out of range");
  const int cpg = C / groups;
  for (int c = 0; c < C; c += 8)  // a thread's 8 channels may touch two groups, not more (cpg >= 7, or 4)
    LDN_CHECK((c + 7) / cpg - c / cpg <= 1, "groupnorm: channels per group too small for the 8-channel vectors");
  // workspace: [slot][B][32] x (sum, sum of squares) as 64-bit fixed point; the slot must be zero when the statistics run
  unsigned long long* acc = groupnorm_slot(stats_ws, slot, B);
  const int nvec = C / 8;
  // R pixel rows per block pass; every thread should see >= 4 pixels (>= 8 when the tensor is large) so that its
  // 16-byte loads overlap, and the grid should still cover the 148 SMs where the tensor is big enough for that.
  int R = 640 / nvec;  // blocks of at most 640 threads (the kernels' launch bounds)
  if (R < 1) R = 1;
  if (R > 16) R = 16;
  while (R > 1 && (HW / (R * 4)) * B < 148) R >>= 1;
  const int threads = (nvec * R + 31) / 32 * 32;
  const int want_blocks = (148 * 2 + B - 1) / B;  // per batch row
  int rows_per_block = (HW + want_blocks - 1) / want_blocks;
  if (rows_per_block < 8 * R) rows_per_block = (HW >= 8 * R * want_blocks / 2) ? 8 * R : 4 * R;
  int splits = (HW + rows_per_block - 1) / rows_per_block;
  if (splits > 1024) {
    rows_per_block = (HW + 1023) / 1024;
    splits = (HW + rows_per_block - 1) / rows_per_block;
  }
  if (!have_stats) {
    gn_stats_kernel<<<dim3(splits, B), threads, 0, stream>>>(x0, C0, x1, C1, HW, cpg, rows_per_block, R, acc);
    LDN_CUDA(cudaGetLastError());
  }
  const double n_elems = (double)HW * cpg;  // per (batch, group)
  gn_apply_kernel<<<dim3(splits, B), threads, 0, stream>>>(x0, C0, x1, C1, HW, cpg, gamma, beta, silu ? 1 : 0, acc, eps, out,
                                                           rows_per_block, R, 1.0 / (LDN_GN_SUM_SCALE * n_elems),
                                                           1.0 / (LDN_GN_SQ_SCALE * n_elems));
  LDN_CUDA(cudaGetLastError());
}

// ------------------------------------------------------------------ LayerNorm: one warp per row, row held in registers
template <int MAXV>
__global__ void layernorm_kernel(const bf16* __restrict__ x, int rows, int C, float eps, const float* __restrict__ gamma,
                                 const float* __restrict__ beta, bf16* __restrict__ out, float* __restrict__ out_f32) {
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
  const float rstd = rsqrtf(sq / (float)C + eps);
  bf16* dst = out + (size_t)warp * C;
#pragma unroll
  for (int k = 0; k < MAXV; ++k) {
    const int vi = lane + k * 32;
    if (vi < nvec) {
      const int c = vi * 8;
      const float4 g0 = *reinterpret_cast<const float4*>(gamma + c);
      const float4 g1 = *reinterpret_cast<const float4*>(gamma + c + 4);
      const float4 b0 = *reinterpret_cast<const float4*>(beta + c);
      const float4 b1 = *reinterpret_cast<const float4*>(beta + c + 4);
      const float g[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
      const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
      if (out_f32) {
#pragma unroll
        for (int i = 0; i < 8; ++i) out_f32[(size_t)warp * C + c + i] = (v[k][i] - mean) * rstd * g[i] + bb[i];
      } else {
        uint32_t o[4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
          o[i] = pack_bf16x2((v[k][2 * i] - mean) * rstd * g[2 * i] + bb[2 * i],
                             (v[k][2 * i + 1] - mean) * rstd * g[2 * i + 1] + bb[2 * i + 1]);
        *reinterpret_cast<uint4*>(dst + c) = make_uint4(o[0], o[1], o[2], o[3]);
      }
    }
  }
}

void launch_layernorm(const bf16* x, int rows, int C, float eps, const float* gamma, const float* beta, bf16* out,
                      cudaStream_t stream, float* out_f32) {
  LDN_CHECK(C % 8 == 0 && C <= 8 * 32 * 6, "layernorm: C must be a multiple of 8 and <= 1536");
  const int threads = 256;
  const int blocks = (rows * 32 + threads - 1) / threads;
  const int nvec = C / 8;
  if (nvec <= 64)
    layernorm_kernel<2><<<blocks, threads, 0, stream>>>(x, rows, C, eps, gamma, beta, out, out_f32);
  else if (nvec <= 96)
    layernorm_kernel<3><<<blocks, threads, 0, stream>>>(x, rows, C, eps, gamma, beta, out, out_f32);
  else
    layernorm_kernel<6><<<blocks, threads, 0, stream>>>(x, rows, C, eps, gamma, beta, out, out_f32);
  LDN_CUDA(cudaGetLastError());
}

}  // namespace ldn
