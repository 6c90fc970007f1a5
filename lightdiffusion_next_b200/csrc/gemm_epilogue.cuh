// Shared pieces of the tcgen05 GEMM kernels (gemm.cu: one-CTA tiles; gemm_pair.cu: CTA-pair tiles): tile constants, GELU,
// and the fused epilogue that drains one 128 x BN fp32 accumulator tile from TMEM.
#pragma once
#include "common.h"
#include "ptx.cuh"

namespace ldn {

__device__ __forceinline__ float tanh_approx(float x) {
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(x));
  return t;
}

static constexpr int kBM = 128;
static constexpr int kBK = 64;
static constexpr int kGemmThreads = 320;  // warp 0 TMA, warp 1 MMA, warps 2..9 epilogue (two per TMEM lane quarter)

// GELU(x) = x * Phi(x) with the exact-erf definition the reference uses (F.gelu, src/cond/Activation.py:31).
// erf via Abramowitz-Stegun 7.1.26 (|abs err| <= 1.5e-7, far below bf16 output rounding): 2 MUFU + ~12 FMA/ALU per value
// instead of erff's ~40 instructions -- the GEGLU GEMMs (K = C only) are epilogue-bound, so this is their critical path.
__device__ __forceinline__ float gelu_erf(float x) {
  const float z = fabsf(x) * 0.70710678118654752440f;
  float t;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, z, 1.0f)));
  float p = fmaf(1.061405429f, t, -1.453152027f);
  p = fmaf(p, t, 1.421413741f);
  p = fmaf(p, t, -0.284496736f);
  p = fmaf(p, t, 0.254829592f);
  p *= t;
  const float e = __expf(-z * z);        // exp(-x^2 / 2)
  const float erf_abs = fmaf(-p, e, 1.0f);  // erf(|x| / sqrt(2))
  const float phi = 0.5f * (1.0f + copysignf(erf_abs, x));
  return x * phi;
}

// Two GELUs per instruction stream (FFMA2 / FMUL2): same A&S 7.1.26 arithmetic as gelu_erf, evaluated on a pair.
// Returns a * gelu(g) element-wise.
__device__ __forceinline__ float2 geglu2(float2 a, float2 g) {
  const float2 ax = make_float2(fabsf(g.x), fabsf(g.y));
  const float2 den = ffma2(ax, make_float2(0.3275911f * 0.70710678118654752440f, 0.3275911f * 0.70710678118654752440f),
                           make_float2(1.0f, 1.0f));
  float2 t;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t.x) : "f"(den.x));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t.y) : "f"(den.y));
  // negated polynomial: q = -(a1 t + a2 t^2 + ... + a5 t^5)
  float2 q = ffma2(make_float2(-1.061405429f, -1.061405429f), t, make_float2(1.453152027f, 1.453152027f));
  q = ffma2(q, t, make_float2(-1.421413741f, -1.421413741f));
  q = ffma2(q, t, make_float2(0.284496736f, 0.284496736f));
  q = ffma2(q, t, make_float2(-0.254829592f, -0.254829592f));
  q = fmul2(q, t);
  const float2 u = fmul2(fmul2(g, g), make_float2(-0.72134752044448170368f, -0.72134752044448170368f));  // -x^2/2 * log2(e)
  float2 e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e.x) : "f"(u.x));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e.y) : "f"(u.y));
  const float2 erf_abs = ffma2(q, e, make_float2(1.0f, 1.0f));  // erf(|x| / sqrt(2))
  const float2 phi = ffma2(make_float2(copysignf(erf_abs.x, g.x), copysignf(erf_abs.y, g.y)), make_float2(0.5f, 0.5f),
                           make_float2(0.5f, 0.5f));
  return fmul2(a, fmul2(g, phi));
}

// Lean epilogue: out = acc (+ bias) (+ rowbias) (+ residual) as bf16 with 256-bit accesses -- the residual projections,
// proj_in and the ResBlock convs.  The general epilogue below carries activations, column gates, head-slot scatter, fp32
// output, GEGLU and split-K as run-time branches; these kernels are bound by their epilogue's instruction stream
// (profiles/r2_experiments.md section 16: six more live registers and a few dead branches cost 0.38 ms per step), so the
// common case gets its own compile-time variant.  The residual and bias loads are issued before the TMEM wait.
__device__ __forceinline__ void gemm_epilogue_tile_lean(const GemmParams& p, const int BN, const int n0, const long long out_row,
                                                        const int batch, const uint32_t t_lane, const int ehalf) {
  const float* rb = p.rowbias ? p.rowbias + (long long)batch * p.ld_rowbias : nullptr;
  for (int c = ehalf * 16; c < BN; c += 32) {
    uint32_t v[16];
    tmem_ld16(t_lane + (uint32_t)c, v);
    const int n = n0 + c;
    const bool ok = out_row >= 0 && n < p.N;
    uint32_t w[8];
    float4 b4[4];
    if (ok) {
      if (p.residual) ld_global_256(p.residual + out_row * p.ldr + n, w);
      if (p.bias) {
#pragma unroll
        for (int i = 0; i < 4; ++i) b4[i] = __ldg(reinterpret_cast<const float4*>(p.bias + n) + i);
      } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) b4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
      if (rb) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float4 r4 = __ldg(reinterpret_cast<const float4*>(rb + n) + i);
          b4[i].x += r4.x; b4[i].y += r4.y; b4[i].z += r4.z; b4[i].w += r4.w;
        }
      }
    }
    tmem_ld_wait();
    if (ok) {
      float f[16];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        f[4 * i] = __uint_as_float(v[4 * i]) + b4[i].x;
        f[4 * i + 1] = __uint_as_float(v[4 * i + 1]) + b4[i].y;
        f[4 * i + 2] = __uint_as_float(v[4 * i + 2]) + b4[i].z;
        f[4 * i + 3] = __uint_as_float(v[4 * i + 3]) + b4[i].w;
      }
      if (p.residual) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          f[2 * i] += bf16_lo(w[i]);
          f[2 * i + 1] += bf16_hi(w[i]);
        }
      }
      uint32_t o[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) o[i] = pack_bf16x2(f[2 * i], f[2 * i + 1]);
      st_global_256(p.out + out_row * p.ldo + n, o);
    }
    __syncwarp();
  }
}

// Lean epilogue with the residual PREFETCHED: these GEMMs are bound by the memory-level parallelism of their epilogue --
// every warp had one 32-byte-per-lane residual read in flight at a time (16 KB per SM against ~1 us of latency = the
// 2.5 TB/s the K = 320 projections reach).  The persistent kernel's epilogue warps request the residual of the first kPf
// chunks of their NEXT tile before they wait for its accumulator, so the reads overlap that tile's main loop.
// One-tile kernels: the 256 epilogue threads (idle while the main loop runs) leave bias[n] + rowbias[image][n] of the tile's BN
// columns in shared memory (`bsum`, behind the barriers) and meet on their named barrier.  Valid when every row of the tile takes
// the same row bias (GemmParams::bias_smem: conv tiles of one image, or no row bias at all).
__device__ __forceinline__ void lean_stage_bias(const GemmParams& p, const int BN, const int n0, int image, float* bsum) {
  if (p.conv && image >= p.B) image = p.B - 1;  // the odd tail tile of a CTA pair lies past the last image (it stores nothing)
  const int e = (int)threadIdx.x - 64;
  if (e < BN) {
    const int n = n0 + e;
    float v = 0.f;
    if (n < p.N) {
      if (p.bias) v = __ldg(p.bias + n);
      if (p.rowbias) v += __ldg(p.rowbias + (long long)image * p.ld_rowbias + n);
    }
    bsum[e] = v;
  }
  asm volatile("bar.sync 1, 256;" ::: "memory");
}
template <int kPf>
__device__ __forceinline__ void lean_prefetch_residual(const GemmParams& p, const int BN, const int n0, const long long out_row,
                                                       const int ehalf, uint32_t (&wres)[kPf][8]) {
  if (p.residual == nullptr || out_row < 0) return;
#pragma unroll
  for (int k = 0; k < kPf; ++k) {
    const int c = ehalf * 16 + k * 32;
    if (c < BN && n0 + c < p.N) ld_global_256(p.residual + out_row * p.ldr + n0 + c, wres[k]);
  }
}
template <int kPf>
__device__ __forceinline__ void gemm_epilogue_tile_lean_pf(const GemmParams& p, const int BN, const int n0, const long long out_row,
                                                           const int batch, const uint32_t t_lane, const int ehalf,
                                                           uint32_t (&wres)[kPf][8], const float* bsum = nullptr) {
  // bsum (one-tile kernels): bias + time-embedding row bias of the tile's columns, summed into shared memory by the epilogue
  // warps while the main loop ran (lean_stage_bias) -- a broadcast LDS instead of eight L2-latency loads per chunk
  const float* rb = p.rowbias ? p.rowbias + (long long)batch * p.ld_rowbias : nullptr;
  auto chunk = [&](const int c, uint32_t* wpre) {
    uint32_t v[16];
    tmem_ld16(t_lane + (uint32_t)c, v);
    const int n = n0 + c;
    const bool ok = out_row >= 0 && n < p.N;
    uint32_t wl[8];
    float4 b4[4];
    if (bsum != nullptr) {
      if (ok && p.residual && wpre == nullptr) ld_global_256(p.residual + out_row * p.ldr + n, wl);
#pragma unroll
      for (int i = 0; i < 4; ++i) b4[i] = *reinterpret_cast<const float4*>(bsum + c + 4 * i);
    } else if (ok) {
      if (p.residual && wpre == nullptr) ld_global_256(p.residual + out_row * p.ldr + n, wl);
      if (p.bias) {
#pragma unroll
        for (int i = 0; i < 4; ++i) b4[i] = __ldg(reinterpret_cast<const float4*>(p.bias + n) + i);
      } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) b4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
      if (rb) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float4 r4 = __ldg(reinterpret_cast<const float4*>(rb + n) + i);
          b4[i].x += r4.x; b4[i].y += r4.y; b4[i].z += r4.z; b4[i].w += r4.w;
        }
      }
    }
    tmem_ld_wait();
    if (ok) {
      float f[16];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        f[4 * i] = __uint_as_float(v[4 * i]) + b4[i].x;
        f[4 * i + 1] = __uint_as_float(v[4 * i + 1]) + b4[i].y;
        f[4 * i + 2] = __uint_as_float(v[4 * i + 2]) + b4[i].z;
        f[4 * i + 3] = __uint_as_float(v[4 * i + 3]) + b4[i].w;
      }
      if (p.residual) {
        const uint32_t* w = wpre ? wpre : wl;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          f[2 * i] += bf16_lo(w[i]);
          f[2 * i + 1] += bf16_hi(w[i]);
        }
      }
      uint32_t o[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) o[i] = pack_bf16x2(f[2 * i], f[2 * i + 1]);
      st_global_256(p.out + out_row * p.ldo + n, o);
    }
    __syncwarp();
  };
#pragma unroll
  for (int k = 0; k < kPf; ++k) {
    const int c = ehalf * 16 + k * 32;
    if (c < BN) chunk(c, wres[k]);
  }
  for (int c = ehalf * 16 + kPf * 32; c < BN; c += 32) chunk(c, nullptr);
}

// Lean epilogue that also takes the GroupNorm statistics of the tile it stores (statistics in the producer's
// epilogue).  The consumer of a ResBlock conv's output is a GroupNorm over the same 32-group channel split, whose statistics
// kernel re-read the whole tensor; here every epilogue thread (one output pixel) adds the fp32 values it is about to round
// and store to a running (sum, sum of squares) of the current group.  Compile-time shape: BN = 160 (every SD1.5 conv:
// N = 320 / 640 / 1280) and kCpg = N / 32 channels per group (10 / 20 / 40, all dividing 160), so that which group a column
// belongs to -- and where a group ends -- is known after unrolling.  A finished group's partial goes to a per-warp scratch
// matrix [entry = group * 2 + moment][lane] in shared memory -- the operand ring, which is dead once the accumulator is
// complete (one tile per CTA) --; at the end lane e of the warp adds up the 32 pixels of entry e in a fixed order and leaves
// the total as 64-bit FIXED-POINT (the format of norm.cu's accumulators: integer addition is associative, so the totals stay
// bit-deterministic) in the warp's row of a small table; the kernel adds the eight rows and then the tile's partial to the
// (batch, group) accumulators of the GroupNorm instance.  No shuffles, no shared-memory atomics (a 64-bit shared-memory add is
// a compare-and-swap loop): the first version, built on both, cost ~6 us per level-0 conv (profiles/r2_experiments.md section 25).
// All 128 rows of the tile belong to one image (the plan checks BB = 1).
constexpr int kGnScratchStride = 33;                                   // floats per entry: 32 lanes + 1 (bank-conflict-free both ways)
constexpr int kGnScratchWarpBytes = 32 * kGnScratchStride * 4;          // 4224 B per epilogue warp
constexpr int kGnScratchBytes = 8 * kGnScratchWarpBytes + 8 * 32 * 8;   // + the [8 warps][32 entries] table of 64-bit totals

// bit g set: this warp (epilogue half kEh) stores columns of group g of the tile
template <int kCpg, int kEh>
__host__ __device__ constexpr uint32_t gn_group_mask() {
  uint32_t m = 0;
  for (int k = 0; k < 5; ++k)
    for (int i = 0; i < 16; ++i) m |= 1u << ((kEh * 16 + k * 32 + i) / kCpg);
  return m;
}

template <int kPf, int kCpg, int kEh>
__device__ __forceinline__ void lean_pf_gn_half(const GemmParams& p, const int n0, const long long out_row, const int batch,
                                                const uint32_t t_lane, uint32_t (&wres)[kPf][8], float* scratch,
                                                unsigned long long* table_row, const float* bsum) {
  const float* rb = p.rowbias ? p.rowbias + (long long)batch * p.ld_rowbias : nullptr;
  const int lane = threadIdx.x & 31;
  const bool ok = out_row >= 0;
  float gs = 0.f, gq = 0.f;
#pragma unroll
  for (int k = 0; k < 5; ++k) {
    const int c = kEh * 16 + k * 32;
    uint32_t v[16];
    tmem_ld16(t_lane + (uint32_t)c, v);
    const int n = n0 + c;
    uint32_t wl[8];
    float4 b4[4];
    if (bsum != nullptr) {
      if (ok && p.residual && k >= kPf) ld_global_256(p.residual + out_row * p.ldr + n, wl);
#pragma unroll
      for (int i = 0; i < 4; ++i) b4[i] = *reinterpret_cast<const float4*>(bsum + c + 4 * i);
    } else if (ok) {
      if (p.residual && k >= kPf) ld_global_256(p.residual + out_row * p.ldr + n, wl);
      if (p.bias) {
#pragma unroll
        for (int i = 0; i < 4; ++i) b4[i] = __ldg(reinterpret_cast<const float4*>(p.bias + n) + i);
      } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) b4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
      if (rb) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float4 r4 = __ldg(reinterpret_cast<const float4*>(rb + n) + i);
          b4[i].x += r4.x; b4[i].y += r4.y; b4[i].z += r4.z; b4[i].w += r4.w;
        }
      }
    }
    tmem_ld_wait();
    float f[16];
    if (ok) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        f[4 * i] = __uint_as_float(v[4 * i]) + b4[i].x;
        f[4 * i + 1] = __uint_as_float(v[4 * i + 1]) + b4[i].y;
        f[4 * i + 2] = __uint_as_float(v[4 * i + 2]) + b4[i].z;
        f[4 * i + 3] = __uint_as_float(v[4 * i + 3]) + b4[i].w;
      }
      if (p.residual) {
        const uint32_t* w = k < kPf ? wres[k < kPf ? k : 0] : wl;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          f[2 * i] += bf16_lo(w[i]);
          f[2 * i + 1] += bf16_hi(w[i]);
        }
      }
      uint32_t o[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) o[i] = pack_bf16x2(f[2 * i], f[2 * i + 1]);
      st_global_256(p.out + out_row * p.ldo + n, o);
    } else {
#pragma unroll
      for (int i = 0; i < 16; ++i) f[i] = 0.f;  // pixels past the image edge count for nothing
    }
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      gs += f[i];
      gq = fmaf(f[i], f[i], gq);
      // the next column this warp accumulates: the neighbour, the first column of its next chunk, or none
      const int next = i < 15 ? c + i + 1 : (k < 4 ? c + 32 : -1);
      if (next < 0 || next / kCpg != (c + i) / kCpg) {
        const int e = 2 * ((c + i) / kCpg);  // a warp meets every group at most once (its columns ascend)
        scratch[e * kGnScratchStride + lane] = gs;
        scratch[(e + 1) * kGnScratchStride + lane] = gq;
        gs = 0.f;
        gq = 0.f;
      }
    }
  }
  __syncwarp();
  // lane e: entry e = (group e / 2, moment e % 2) over the warp's 32 pixels, in a fixed order
  constexpr uint32_t kMask = gn_group_mask<kCpg, kEh>();
  unsigned long long total = 0ull;
  if ((kMask >> (lane >> 1)) & 1u) {
    const float* col = scratch + lane * kGnScratchStride;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
    for (int l = 0; l < 32; l += 4) {
      a0 += col[l];
      a1 += col[l + 1];
      a2 += col[l + 2];
      a3 += col[l + 3];
    }
    const float t = (a0 + a1) + (a2 + a3);
    total = (unsigned long long)__float2ll_rn(t * ((lane & 1) ? (float)LDN_GN_SQ_SCALE : (float)LDN_GN_SUM_SCALE));
  }
  table_row[lane] = total;
}
// scratch_base: kGnScratchBytes of shared memory nobody else touches any more (1024-byte aligned start of the operand ring)
template <int kPf, int kCpg>
__device__ __forceinline__ void gemm_epilogue_tile_lean_pf_gn(const GemmParams& p, const int n0, const long long out_row,
                                                              const int batch, const uint32_t t_lane, const int ehalf,
                                                              uint32_t (&wres)[kPf][8], uint8_t* scratch_base,
                                                              const float* bsum = nullptr) {
  const int ew = (int)(threadIdx.x >> 5) - 2;  // epilogue warp 0..7
  float* scratch = reinterpret_cast<float*>(scratch_base + ew * kGnScratchWarpBytes);
  unsigned long long* table_row = reinterpret_cast<unsigned long long*>(scratch_base + 8 * kGnScratchWarpBytes) + ew * 32;
  if (ehalf == 0)
    lean_pf_gn_half<kPf, kCpg, 0>(p, n0, out_row, batch, t_lane, wres, scratch, table_row, bsum);
  else
    lean_pf_gn_half<kPf, kCpg, 1>(p, n0, out_row, batch, t_lane, wres, scratch, table_row, bsum);
}
// after a barrier over the 256 epilogue threads: thread e < 2 * groups-per-tile adds the eight warps' totals of entry e and then
// the tile's partial to the (batch, group) accumulators
template <int kCpg>
__device__ __forceinline__ void gn_flush_tile(const GemmParams& p, const int n0, const int image, const uint8_t* scratch_base) {
  const int e = (int)threadIdx.x - 64;
  if (e < 2 * (160 / kCpg)) {
    const unsigned long long* table = reinterpret_cast<const unsigned long long*>(scratch_base + 8 * kGnScratchWarpBytes);
    unsigned long long t = 0ull;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += table[w * 32 + e];
    atomicAdd(&p.gn_acc[((size_t)image * 32 + (size_t)(n0 / kCpg)) * 2 + (size_t)e], t);
  }
}

// GEGLU without the MUFU: Phi(x) = 0.5 + x Q(min(x^2, 16)), Q a degree-7 minimax polynomial (max |dPhi| 3.2e-5 -- the
// cut-off at |x| = 4, where 1 - Phi = 3.2e-5 --, max |dGELU| 1.3e-4, an order of magnitude below the bf16 rounding of the
// output), saturated to [0, 1] by the FMA itself.  The level-0 GEGLU projection (K = 320) was bound by the epilogue's four MUFU
// per value pair (rcp + ex2 of the erf formula above): ~4100 XU-clk per 128 x 256 tile against 2560 clk of main loop.
// Same instruction count, no MUFU.  Returns a * gelu(g) element-wise.
__device__ __forceinline__ float2 geglu2_poly(float2 a, float2 g) {
  float2 t = fmul2(g, g);
  t.x = fminf(t.x, 16.0f);
  t.y = fminf(t.y, 16.0f);
  float2 q = ffma2(t, make_float2(-1.580773833e-09f, -1.580773833e-09f), make_float2(1.217103702e-07f, 1.217103702e-07f));
  q = ffma2(q, t, make_float2(-4.100848923e-06f, -4.100848923e-06f));
  q = ffma2(q, t, make_float2(8.066718382e-05f, 8.066718382e-05f));
  q = ffma2(q, t, make_float2(-1.048203032e-03f, -1.048203032e-03f));
  q = ffma2(q, t, make_float2(9.664869643e-03f, 9.664869643e-03f));
  q = ffma2(q, t, make_float2(-6.617537275e-02f, -6.617537275e-02f));
  q = ffma2(q, t, make_float2(3.988475075e-01f, 3.988475075e-01f));
  float2 phi;
  asm("fma.rn.sat.f32 %0, %1, %2, 0f3F000000;" : "=f"(phi.x) : "f"(g.x), "f"(q.x));
  asm("fma.rn.sat.f32 %0, %1, %2, 0f3F000000;" : "=f"(phi.y) : "f"(g.y), "f"(q.y));
  return fmul2(a, fmul2(g, phi));
}

// Drains one 128 x BN fp32 accumulator tile from TMEM (columns starting at t_lane) through the fused epilogue.
// Executed by the 8 epilogue warps; `ehalf` selects which alternate 16-column chunks this warp handles.
// `m_row`: the GEMM row index of this thread (token / channel) before any head-slot remapping, or -1 (conv).
// kLn (compile time, so that the plain epilogue carries none of it -- these epilogues are sensitive to every register and
// instruction): 0 no folded LayerNorm; 1 consumer, row mode; 2 consumer, column mode; 3 producer of row statistics.
template <int kLn = 0>
__device__ __forceinline__ void gemm_epilogue_tile(const GemmParams& p, const int BN, const int n0, const long long out_row,
                                                   const int batch, const uint32_t t_lane, const int ehalf,
                                                   const int split_z, const int m_row = -1) {
    // ---- folded LayerNorm, consumer side: per-row (rstd, -rstd * mean) from the producer's partial sums (fixed order)
    float ln_a = 1.f, ln_b = 0.f;
    if (kLn == 1 && m_row >= 0 && m_row < p.M) {
      float sum = 0.f, sq = 0.f;
      const float2* pp = p.ln_parts + (long long)m_row * p.ln_nparts;
      for (int i = 0; i < p.ln_nparts; ++i) {
        const float2 v = pp[i];
        sum += v.x;
        sq += v.y;
      }
      const float mean = sum * p.ln_inv_k;
      const float var = fmaxf(sq * p.ln_inv_k - mean * mean, 0.f);
      ln_a = rsqrtf(var + p.ln_eps);
      ln_b = -ln_a * mean;
      if (p.ln_final_out != nullptr && n0 == 0 && ehalf == 0) p.ln_final_out[m_row] = make_float2(ln_a, ln_b);
    }
    float ln_cv = 0.f, ln_dv = 0.f;  // column mode: this output row's c / d
    if (kLn == 2 && m_row >= 0 && m_row < p.M) {
      ln_cv = p.ln_c[m_row];
      ln_dv = p.ln_d[m_row];
    }
    float st_sum = 0.f, st_sq = 0.f;  // producer side: statistics of the values this thread stores

    if (p.splits > 1) {
      // split-K: raw fp32 partial tile; bias / residual are applied by splitk_reduce_kernel
      float* wbase = p.ws + (long long)split_z * p.ws_split_stride;
      for (int c = ehalf * 16; c < BN; c += 32) {
        uint32_t v[16];
        tmem_ld16(t_lane + (uint32_t)c, v);
        tmem_ld_wait();
        const int n = n0 + c;
        if (out_row >= 0 && n < p.N) {
          float4* op = reinterpret_cast<float4*>(wbase + out_row * p.N + n);
#pragma unroll
          for (int i = 0; i < 4; ++i)
            op[i] = make_float4(__uint_as_float(v[4 * i]), __uint_as_float(v[4 * i + 1]), __uint_as_float(v[4 * i + 2]),
                                __uint_as_float(v[4 * i + 3]));
        }
        __syncwarp();
      }
    } else if (p.epi == 0) {
      for (int c = ehalf * 16; c < BN; c += 32) {
        uint32_t v[16];
        tmem_ld16(t_lane + (uint32_t)c, v);
        tmem_ld_wait();
        const int n = n0 + c;
        if (out_row >= 0 && n < p.N) {
        float f[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) f[i] = __uint_as_float(v[i]);
        if (kLn == 1) {
#pragma unroll
          for (int i = 0; i < 16; i += 4) {
            const float4 cv = __ldg(reinterpret_cast<const float4*>(p.ln_c + n + i));
            const float4 dv = __ldg(reinterpret_cast<const float4*>(p.ln_d + n + i));
            f[i] = fmaf(ln_a, f[i], fmaf(ln_b, cv.x, dv.x));
            f[i + 1] = fmaf(ln_a, f[i + 1], fmaf(ln_b, cv.y, dv.y));
            f[i + 2] = fmaf(ln_a, f[i + 2], fmaf(ln_b, cv.z, dv.z));
            f[i + 3] = fmaf(ln_a, f[i + 3], fmaf(ln_b, cv.w, dv.w));
          }
        } else if (kLn == 2) {
#pragma unroll
          for (int i = 0; i < 16; i += 2) {  // (rstd, -rstd * mean) of two tokens per 16-byte load
            const float4 ab = __ldg(reinterpret_cast<const float4*>(p.ln_final_in + n + i));
            f[i] = fmaf(ab.x, f[i], fmaf(ab.y, ln_cv, ln_dv));
            f[i + 1] = fmaf(ab.z, f[i + 1], fmaf(ab.w, ln_cv, ln_dv));
          }
        }
        if (p.bias) {
#pragma unroll
          for (int i = 0; i < 16; i += 4) {
            const float4 bv = *reinterpret_cast<const float4*>(p.bias + n + i);
            f[i] += bv.x; f[i + 1] += bv.y; f[i + 2] += bv.z; f[i + 3] += bv.w;
          }
        }
        if (p.rowbias) {
          const float* rb = p.rowbias + (long long)batch * p.ld_rowbias + n;
#pragma unroll
          for (int i = 0; i < 16; i += 4) {
            const float4 bv = *reinterpret_cast<const float4*>(rb + i);
            f[i] += bv.x; f[i + 1] += bv.y; f[i + 2] += bv.z; f[i + 3] += bv.w;
          }
        }
        if (p.act == 1) {
          // x * sigmoid(1.702 x) = x * (0.5 + 0.5 tanh(0.851 x)): one MUFU (tanh.approx, rel. error 2^-11) instead of ex2 + rcp
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float hx = 0.5f * f[i];
            f[i] = fmaf(hx, tanh_approx(0.851f * f[i]), hx);
          }
        } else if (p.act == 2) {
#pragma unroll
          for (int i = 0; i < 16; ++i) f[i] = fmaxf(f[i], 0.f);
        } else if (p.act == 4) {
          // GELU, tanh approximation: 0.5 x (1 + tanh(u)) = x / (1 + exp(-2u)), u = sqrt(2/pi) (x + 0.044715 x^3)
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float x = f[i];
            const float u = 0.7978845608028654f * fmaf(0.044715f * x * x, x, x);
            const float hx = 0.5f * x;
            f[i] = fmaf(hx, tanh_approx(u), hx);  // one MUFU instead of ex2 + rcp
          }
        }
        if (p.colgate) {
          const float* gp = p.colgate + (long long)batch * p.ld_colgate + n;
#pragma unroll
          for (int i = 0; i < 16; i += 4) {
            const float4 gv = *reinterpret_cast<const float4*>(gp + i);
            f[i] *= gv.x; f[i + 1] *= gv.y; f[i + 2] *= gv.z; f[i + 3] *= gv.w;
          }
        }
        if (p.residual) {
          uint32_t w[8];
          if (p.epi_opt & 1) {
            ld_global_256(p.residual + out_row * p.ldr + n, w);  // one full 32-byte sector per thread
          } else {
            const uint4* rp = reinterpret_cast<const uint4*>(p.residual + out_row * p.ldr + n);
            const uint4 r0 = rp[0], r1 = rp[1];
            w[0] = r0.x; w[1] = r0.y; w[2] = r0.z; w[3] = r0.w; w[4] = r1.x; w[5] = r1.y; w[6] = r1.z; w[7] = r1.w;
          }
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            f[2 * i] += bf16_lo(w[i]);
            f[2 * i + 1] += bf16_hi(w[i]);
          }
        }
        if (p.act == 3) {  // ReLU after the residual add (TAESD Block: relu(conv(x) + skip(x)))
#pragma unroll
          for (int i = 0; i < 16; ++i) f[i] = fmaxf(f[i], 0.f);
        }
        if (kLn == 3) {
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            st_sum += f[i];
            st_sq = fmaf(f[i], f[i], st_sq);
          }
        }
        if (p.out_f32) {
          float4* op = reinterpret_cast<float4*>(p.out_f32 + out_row * p.ldo + n);
#pragma unroll
          for (int i = 0; i < 4; ++i) op[i] = make_float4(f[4 * i], f[4 * i + 1], f[4 * i + 2], f[4 * i + 3]);
        } else if ((p.epi_opt & 1) && p.head_dim == 0) {
          uint32_t o[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) o[i] = pack_bf16x2(f[2 * i], f[2 * i + 1]);
          st_global_256(p.out + out_row * p.ldo + n, o);
        } else {
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            int col = n + h * 8;
            if (p.head_dim > 0) col = (col / p.head_dim) * p.head_slot + (col % p.head_dim);
            uint4 ov;
            ov.x = pack_bf16x2(f[h * 8 + 0], f[h * 8 + 1]);
            ov.y = pack_bf16x2(f[h * 8 + 2], f[h * 8 + 3]);
            ov.z = pack_bf16x2(f[h * 8 + 4], f[h * 8 + 5]);
            ov.w = pack_bf16x2(f[h * 8 + 6], f[h * 8 + 7]);
            *reinterpret_cast<uint4*>(p.out + out_row * p.ldo + col) = ov;
          }
        }
        }
        __syncwarp();
      }
      if (kLn == 3 && m_row >= 0 && m_row < p.M)
        p.rowstat_out[(long long)m_row * p.rowstat_parts + (n0 / BN) * 2 + ehalf] = make_float2(st_sum, st_sq);
    } else {
      // GEGLU: weight rows were interleaved at load time so that this tile holds BN/2 value columns followed
      // by the BN/2 matching gate columns (Activation.py:30-31: x, gate = proj(x).chunk(2); x * gelu(gate)).
      const int half = BN / 2;
      const int o0 = n0 / 2;
      for (int c = ehalf * 16; c < half; c += 32) {
        uint32_t va[16], vg[16];
        tmem_ld16(t_lane + (uint32_t)c, va);
        tmem_ld16(t_lane + (uint32_t)(half + c), vg);
        float ba[16], bg[16];  // bias vectors fetched while the TMEM loads are in flight
        if (kLn == 1 && (n0 + c) < p.N) {
          // folded LayerNorm: value / gate = ln_a * acc + (ln_b * c[n] + d[n]); the bias travels inside d
#pragma unroll
          for (int i = 0; i < 16; i += 4) {
            const float4 ca = __ldg(reinterpret_cast<const float4*>(p.ln_c + n0 + c + i));
            const float4 da = __ldg(reinterpret_cast<const float4*>(p.ln_d + n0 + c + i));
            const float4 cg = __ldg(reinterpret_cast<const float4*>(p.ln_c + n0 + half + c + i));
            const float4 dg = __ldg(reinterpret_cast<const float4*>(p.ln_d + n0 + half + c + i));
            ba[i] = fmaf(ln_b, ca.x, da.x); ba[i + 1] = fmaf(ln_b, ca.y, da.y); ba[i + 2] = fmaf(ln_b, ca.z, da.z); ba[i + 3] = fmaf(ln_b, ca.w, da.w);
            bg[i] = fmaf(ln_b, cg.x, dg.x); bg[i + 1] = fmaf(ln_b, cg.y, dg.y); bg[i + 2] = fmaf(ln_b, cg.z, dg.z); bg[i + 3] = fmaf(ln_b, cg.w, dg.w);
          }
        } else if (p.bias && (n0 + c) < p.N) {
#pragma unroll
          for (int i = 0; i < 16; i += 4) {
            const float4 x4 = __ldg(reinterpret_cast<const float4*>(p.bias + n0 + c + i));
            const float4 y4 = __ldg(reinterpret_cast<const float4*>(p.bias + n0 + half + c + i));
            ba[i] = x4.x; ba[i + 1] = x4.y; ba[i + 2] = x4.z; ba[i + 3] = x4.w;
            bg[i] = y4.x; bg[i + 1] = y4.y; bg[i + 2] = y4.z; bg[i + 3] = y4.w;
          }
        }
        tmem_ld_wait();
        if (out_row >= 0 && (n0 + c) < p.N) {
        float f[16];
#pragma unroll
        if (!(p.epi_opt & 2)) {
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            float a = __uint_as_float(va[i]);
            float g = __uint_as_float(vg[i]);
            if (kLn == 1) {
              a = fmaf(ln_a, a, ba[i]);
              g = fmaf(ln_a, g, bg[i]);
            } else if (p.bias) {
              a += ba[i];
              g += bg[i];
            }
            f[i] = a * gelu_erf(g);
          }
        } else
#pragma unroll
        for (int i = 0; i < 16; i += 2) {
          float2 a = make_float2(__uint_as_float(va[i]), __uint_as_float(va[i + 1]));
          float2 g = make_float2(__uint_as_float(vg[i]), __uint_as_float(vg[i + 1]));
          if (kLn == 1) {
            a = ffma2(a, make_float2(ln_a, ln_a), make_float2(ba[i], ba[i + 1]));
            g = ffma2(g, make_float2(ln_a, ln_a), make_float2(bg[i], bg[i + 1]));
          } else if (p.bias) {
            a = fadd2(a, make_float2(ba[i], ba[i + 1]));
            g = fadd2(g, make_float2(bg[i], bg[i + 1]));
          }
          const float2 r2 = (p.epi_opt & 4) ? geglu2_poly(a, g) : geglu2(a, g);
          f[i] = r2.x;
          f[i + 1] = r2.y;
        }
        if (p.epi_opt & 1) {
          uint32_t o[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) o[i] = pack_bf16x2(f[2 * i], f[2 * i + 1]);
          st_global_256(p.out + out_row * p.ldo + o0 + c, o);
        } else {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          uint4 ov;
          ov.x = pack_bf16x2(f[h * 8 + 0], f[h * 8 + 1]);
          ov.y = pack_bf16x2(f[h * 8 + 2], f[h * 8 + 3]);
          ov.z = pack_bf16x2(f[h * 8 + 4], f[h * 8 + 5]);
          ov.w = pack_bf16x2(f[h * 8 + 6], f[h * 8 + 7]);
          *reinterpret_cast<uint4*>(p.out + out_row * p.ldo + o0 + c + h * 8) = ov;
        }
        }
        }
        __syncwarp();
      }
    }
}

}  // namespace ldn
