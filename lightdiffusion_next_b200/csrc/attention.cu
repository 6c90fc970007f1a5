// Flash-style fused attention on tcgen05 for sm_100a:  out = softmax(Q K^T * scale) V   per (batch, head).
//
// Replaces optimized_attention (xformers memory_efficient_attention / torch SDPA) on the reference path:
//   src/Attention/Attention.py:34-41,118-124, src/Attention/AttentionMethods.py:16-52,107-134  [reference file:line]
// head_dim on the SD1.5 UNet is 40 / 80 / 160 (src/SD15/SD15.py:25-28, unet.py:478), 64 for CLIP-L.
//
// Data layout (produced by the projection GEMMs, see engine.cu):
//   Q  [B*Nq,     heads*slot] bf16, head h occupies columns [h*slot, h*slot+d), zero padded up to a multiple of 16
//   K  [B*nk_pad, heads*slot] bf16, same slotting
//   Vt [heads*d,  B*nk_pad]   bf16, i.e. V transposed (keys contiguous) so that P*V has a K-major B operand
// One CTA = 128 query rows of one (batch, head):
//   warp 0     TMA producer (Q once; K / Vt tiles of 128 keys through a ring)
//   warp 1     lane 0 issues S = Q K^T (M=128, N=128, K=dqk) and PV = P V (M=128, N=DV, K=128) with tcgen05.mma,
//              accumulators in TMEM (S: columns [0,128), PV: columns [128,128+DV))
//   warps 2..5 softmax: one query row per thread (TMEM lane == row), online max / sum in fp32, P written to
//              shared memory as bf16 in the 128B-swizzled K-major layout tcgen05 reads, O accumulated in registers.
// Two CTAs are resident per SM for d <= 64 so one CTA's MMAs overlap the other's exponentials.
#include "common.h"
#include "ptx.cuh"

#include <cstdlib>

namespace ldn {

static constexpr int kAttnThreads = 192;
static constexpr int kTileQ = 128;
static constexpr int kTileK = 128;

__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// kBias: an additive logit bias (fp32, already multiplied by log2 e, [heads, bias_rows, bias_ld], padded to whole
// 128 x 128 tiles so every read of a tile is in range) is added to scale * q.k before the softmax -- T5's relative
// position bias (t5.cu).  The kBias = false instantiations are unchanged.
template <int DV, bool kBias = false>
__global__ void __launch_bounds__(kAttnThreads, (DV <= 80) ? 2 : 1) attn_tc_kernel(const __grid_constant__ AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int qt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int q0 = qt * kTileQ;
  const int nqk_atoms = (p.dqk + 63) / 64;
  const int nqk_ksteps = p.dqk / 16;
  const int stages = p.kv_stages;
  const uint32_t atom_bytes = 128 * 128;            // 128 rows x 128 B
  const uint32_t q_bytes = nqk_atoms * atom_bytes;
  const uint32_t k_bytes = q_bytes;
  const uint32_t vt_atom_bytes = DV * 128;          // DV rows x 64 keys
  const uint32_t stage_bytes = k_bytes + 2 * vt_atom_bytes;

  uint8_t* q_smem = smem;
  uint8_t* p_smem = smem + q_bytes;                  // 2 atoms
  uint8_t* kv_smem = p_smem + 2 * atom_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(kv_smem + (size_t)stages * stage_bytes);
  uint64_t* q_full = bars;
  uint64_t* s_full = bars + 1;
  uint64_t* p_full = bars + 2;
  uint64_t* pv_full = bars + 3;
  uint64_t* kv_full = bars + 4;
  uint64_t* kv_empty = kv_full + stages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(kv_empty + stages);

  constexpr uint32_t kTmemCols = (128 + DV <= 256) ? 256 : 512;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmQ);
    tma_prefetch_desc(&p.tmK);
    tma_prefetch_desc(&p.tmVt);
    mbar_init(q_full, 1);
    mbar_init(s_full, 1);
    mbar_init(p_full, 128);
    mbar_init(pv_full, 1);
    for (int s = 0; s < stages; ++s) {
      mbar_init(&kv_full[s], 1);
      mbar_init(&kv_empty[s], 1);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_s = tmem_base;
  const uint32_t tmem_pv = tmem_base + 128;

  // keys visible to this query tile
  int nk_eff = p.Nk;
  if (p.causal) nk_eff = min(p.Nk, q0 + kTileQ);
  const int n_tiles = (nk_eff + kTileK - 1) / kTileK;

  if (warp == 0) {
    if (lane == 0) {
      mbar_arrive_expect_tx(q_full, q_bytes);
      for (int a = 0; a < nqk_atoms; ++a)
        tma_load_2d(q_smem + a * atom_bytes, &p.tmQ, q_full, h * p.slot + a * 64, b * p.Nq + q0);
      for (int j = 0; j < n_tiles; ++j) {
        const int s = j % stages;
        const uint32_t ph = (uint32_t)(j / stages) & 1u;
        mbar_wait(&kv_empty[s], ph ^ 1u);
        uint8_t* k_dst = kv_smem + (size_t)s * stage_bytes;
        uint8_t* v_dst = k_dst + k_bytes;
        mbar_arrive_expect_tx(&kv_full[s], stage_bytes);
        const int key0 = b * p.nk_pad + j * kTileK;
        const int krow0 = b * p.k_batch_stride + j * kTileK;
        for (int a = 0; a < nqk_atoms; ++a)
          tma_load_2d(k_dst + a * atom_bytes, &p.tmK, &kv_full[s], h * p.slot + a * 64, krow0);
        tma_load_2d(v_dst, &p.tmVt, &kv_full[s], key0, h * p.d);
        tma_load_2d(v_dst + vt_atom_bytes, &p.tmVt, &kv_full[s], key0 + 64, h * p.d);
      }
    }
  } else if (warp == 1) {
    // whole warp walks the loop (warp-uniform descriptors in uniform registers); one elected lane issues
    {
      const uint32_t idesc_s = make_idesc_bf16(128, 128);
      const uint32_t idesc_pv = make_idesc_bf16(128, DV);
      const uint64_t qd = make_smem_desc_sw128(smem_u32(q_smem));
      const uint64_t pd = make_smem_desc_sw128(smem_u32(p_smem));
      const uint64_t kv0 = make_smem_desc_sw128(smem_u32(kv_smem));
      mbar_wait(q_full, 0);
      int s = 0;
      uint32_t ph = 0;
      for (int j = 0; j < n_tiles; ++j) {
        mbar_wait(&kv_full[s], ph);
        tc_fence_after();
        const uint64_t kd = kv0 + (uint64_t)(((uint32_t)s * stage_bytes) >> 4);
        const uint64_t vd = kd + (uint64_t)(k_bytes >> 4);
        if (elect_one()) {
          // S = Q K^T
          for (int ks = 0; ks < nqk_ksteps; ++ks) {
            const uint64_t off = (uint64_t)(((uint32_t)(ks >> 2) * atom_bytes) >> 4) + (uint64_t)(2 * (ks & 3));
            tc_mma_bf16(tmem_s, qd + off, kd + off, idesc_s, ks > 0 ? 1u : 0u);
          }
          tc_commit(s_full);
        }
        __syncwarp();
        // wait for P (also implies S and the previous PV were consumed)
        mbar_wait(p_full, (uint32_t)j & 1u);
        tc_fence_after();
        if (elect_one()) {
#pragma unroll
          for (int ks = 0; ks < kTileK / 16; ++ks) {
            const uint64_t aoff = (uint64_t)(((uint32_t)(ks >> 2) * atom_bytes) >> 4) + (uint64_t)(2 * (ks & 3));
            const uint64_t boff = (uint64_t)(((uint32_t)(ks >> 2) * vt_atom_bytes) >> 4) + (uint64_t)(2 * (ks & 3));
            tc_mma_bf16(tmem_pv, pd + aoff, vd + boff, idesc_pv, ks > 0 ? 1u : 0u);
          }
          tc_commit(&kv_empty[s]);
          tc_commit(pv_full);
        }
        __syncwarp();
        if (++s == stages) {
          s = 0;
          ph ^= 1u;
        }
      }
    }
  } else {
    // ---------------------------------------------------------------- softmax warps
    const int qd = warp & 3;
    const int r = qd * 32 + lane;              // row in tile == TMEM lane
    const uint32_t lane_off = (uint32_t)(qd * 32) << 16;
    const int q_idx = q0 + r;
    float m_run = -INFINITY;                   // running max (already multiplied by scale*log2e)
    float l_run = 0.f;
    float o[DV];
#pragma unroll
    for (int i = 0; i < DV; ++i) o[i] = 0.f;
    const float sc = p.scale_log2;
    const uint32_t p_row = smem_u32(p_smem) + (uint32_t)r * 128;
    const uint32_t sw = (uint32_t)(r & 7);

    for (int j = 0; j < n_tiles; ++j) {
      mbar_wait(s_full, (uint32_t)j & 1u);
      tc_fence_after();
      const int kbase = j * kTileK;
      int limit = nk_eff - kbase;              // keys [0, limit) of this tile are valid
      if (p.causal) limit = min(limit, q_idx - kbase + 1);
      const bool need_mask = limit < kTileK;
      // pass 1: row max
      float mxa[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};  // independent chains (ILP)
      const float* brow = nullptr;
      if constexpr (kBias) brow = p.bias + ((long long)h * p.bias_rows + q_idx) * p.bias_ld + kbase;
#pragma unroll
      for (int c = 0; c < kTileK; c += 32) {
        uint32_t v[32];
        tmem_ld32(tmem_s + lane_off + (uint32_t)c, v);
        tmem_ld_wait();
        if constexpr (kBias) {  // max of the biased, scaled logits
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            const float4 bv = *reinterpret_cast<const float4*>(brow + c + i);
            const float t0 = fmaf(__uint_as_float(v[i]), sc, bv.x), t1 = fmaf(__uint_as_float(v[i + 1]), sc, bv.y);
            const float t2 = fmaf(__uint_as_float(v[i + 2]), sc, bv.z), t3 = fmaf(__uint_as_float(v[i + 3]), sc, bv.w);
            if (c + i < limit) mxa[0] = fmaxf(mxa[0], t0);
            if (c + i + 1 < limit) mxa[1] = fmaxf(mxa[1], t1);
            if (c + i + 2 < limit) mxa[2] = fmaxf(mxa[2], t2);
            if (c + i + 3 < limit) mxa[3] = fmaxf(mxa[3], t3);
          }
        } else if (need_mask) {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (c + i < limit) mxa[i & 3] = fmaxf(mxa[i & 3], __uint_as_float(v[i]));
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) mxa[i & 3] = fmaxf(mxa[i & 3], __uint_as_float(v[i]));
        }
      }
      const float mx = fmaxf(fmaxf(mxa[0], mxa[1]), fmaxf(mxa[2], mxa[3]));
      const float m_new = fmaxf(m_run, kBias ? mx : mx * sc);
      const float m_use = (m_new == -INFINITY) ? 0.f : m_new;
      const float alpha = fast_exp2(m_run - m_use);   // m_run = -inf -> 0
      float rsa[4] = {0.f, 0.f, 0.f, 0.f};
      // pass 2: p = exp2(s*sc - m), write bf16 P tile (K-major, 128B swizzle)
#pragma unroll
      for (int c = 0; c < kTileK; c += 32) {
        uint32_t v[32];
        tmem_ld32(tmem_s + lane_off + (uint32_t)c, v);
        tmem_ld_wait();
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          float sh0 = -m_use, sh1 = -m_use;
          if constexpr (kBias) {
            const float2 bv = *reinterpret_cast<const float2*>(brow + c + i);
            sh0 += bv.x;
            sh1 += bv.y;
          }
          float p0 = fast_exp2(fmaf(__uint_as_float(v[i]), sc, sh0));
          float p1 = fast_exp2(fmaf(__uint_as_float(v[i + 1]), sc, sh1));
          if (need_mask) {
            if (c + i >= limit) p0 = 0.f;
            if (c + i + 1 >= limit) p1 = 0.f;
          }
          rsa[(i >> 1) & 3] += p0 + p1;
          pk[i >> 1] = pack_bf16x2(p0, p1);
        }
        // 32 keys = 4 x 16B chunks; chunk index within the 64-key atom row: (c/8 + t) & 7, atom = c / 64
        const uint32_t atom_off = (uint32_t)(c >> 6) * atom_bytes;
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const uint32_t chunk = ((uint32_t)((c & 63) >> 3) + t) ^ sw;
          const uint32_t addr = p_row + atom_off + chunk * 16;
          asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(pk[4 * t]), "r"(pk[4 * t + 1]),
                       "r"(pk[4 * t + 2]), "r"(pk[4 * t + 3])
                       : "memory");
        }
      }
      l_run = l_run * alpha + ((rsa[0] + rsa[1]) + (rsa[2] + rsa[3]));
      m_run = m_new;
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(p_full);
      // accumulate O = O*alpha + PV
      mbar_wait(pv_full, (uint32_t)j & 1u);
      tc_fence_after();
#pragma unroll
      for (int c = 0; c < DV; c += 16) {
        uint32_t v[16];
        tmem_ld16(tmem_pv + lane_off + (uint32_t)c, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; ++i) o[c + i] = fmaf(o[c + i], alpha, __uint_as_float(v[i]));
      }
      tc_fence_before();
    }
    // epilogue: normalise and store bf16
    if (q_idx < p.Nq) {
      const float inv = l_run > 0.f ? 1.f / l_run : 0.f;
      bf16* orow = p.out + ((long long)b * p.Nq + q_idx) * p.ldo + (long long)h * p.d;
#pragma unroll
      for (int c = 0; c < DV; c += 8) {
        if (c < p.d) {
          uint4 ov;
          ov.x = pack_bf16x2(o[c + 0] * inv, o[c + 1] * inv);
          ov.y = pack_bf16x2(o[c + 2] * inv, o[c + 3] * inv);
          ov.z = pack_bf16x2(o[c + 4] * inv, o[c + 5] * inv);
          ov.w = pack_bf16x2(o[c + 6] * inv, o[c + 7] * inv);
          *reinterpret_cast<uint4*>(orow + c) = ov;
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

// ----------------------------------------------------------------------------------- host side

AttnPlan make_attn_plan(const AttnArgs& a) {
  AttnPlan plan;
  AttnParams& p = plan.p;
  memset(&p, 0, sizeof(p));
  LDN_CHECK(a.d % 8 == 0, "attention: head dim must be a multiple of 8");
  LDN_CHECK(a.slot % 64 == 0 && a.slot >= a.d, "attention: slot must be a multiple of 64 and >= d");
  const int dp = (a.d + 15) / 16 * 16;
  LDN_CHECK(dp == 48 || dp == 64 || dp == 80 || dp == 128 || dp == 160, "attention: unsupported head dim");
  p.heads = a.heads;
  p.Nq = a.Nq;
  p.Nk = a.Nk;
  p.nk_pad = a.nk_pad;
  p.k_batch_stride = a.kv_batch_stride > 0 ? a.kv_batch_stride : a.nk_pad;
  LDN_CHECK(a.nk_pad % 8 == 0, "attention: V^T batch stride must be a multiple of 8 (16-byte TMA box alignment)");
  p.d = a.d;
  p.dqk = dp;
  p.dv = dp;
  p.slot = a.slot;
  p.causal = a.causal;
  p.scale_log2 = a.fold ? 1.0f : a.scale * 1.4426950408889634f;  // folded operands: Q already carries scale * log2(e)
  p.out = a.out;
  p.ldo = a.ldo;
  p.bias = a.bias;
  p.bias_rows = a.bias_rows;
  p.bias_ld = a.bias_ld;
  if (a.bias) {
    LDN_CHECK(a.d == 64 && !a.causal && a.vt_head_stride == 0, "attention: logit bias needs d = 64, no causal mask, plain V^T");
    LDN_CHECK(a.bias_rows >= (a.Nq + kTileQ - 1) / kTileQ * kTileQ && a.bias_ld >= (a.Nk + kTileK - 1) / kTileK * kTileK &&
                  a.bias_ld % 4 == 0,
              "attention: logit bias must be padded to whole 128 x 128 tiles");
  }
  p.tmQ = make_tmap_2d(a.Q, (uint64_t)a.B * a.Nq, (uint64_t)a.heads * a.slot, a.ldq, 128);
  p.tmK = make_tmap_2d(a.K, (uint64_t)a.B * p.k_batch_stride, (uint64_t)a.heads * a.slot, a.ldk, 128);
  p.tmVt = make_tmap_2d(a.Vt, (uint64_t)a.vt_rows, (uint64_t)a.B * a.nk_pad, a.ldvt, a.vt_head_stride > 0 ? a.vt_head_stride : dp);
  const int nqk_atoms = (dp + 63) / 64;
  const int q_bytes = nqk_atoms * 16384;
  const int stage_bytes = q_bytes + 2 * dp * 128;
  const int fixed = q_bytes + 2 * 16384 + 1024 + 256;
  const int n_tiles = (a.Nk + kTileK - 1) / kTileK;
  int budget = (dp <= 80) ? (113 * 1024) : (226 * 1024);
  int stages = (budget - fixed) / stage_bytes;
  if (stages < 1) {  // does not fit twice per SM: take the whole SM
    budget = 226 * 1024;
    stages = (budget - fixed) / stage_bytes;
  }
  LDN_CHECK(stages >= 1, "attention: tile does not fit in shared memory");
  if (stages > 4) stages = 4;
  if (stages > n_tiles) stages = n_tiles;
  p.kv_stages = stages;
  plan.smem_bytes = fixed + stages * stage_bytes;
  plan.grid = dim3((a.Nq + kTileQ - 1) / kTileQ, a.heads, a.B);
  p.variant = 1;
  p.vt_head_stride = a.vt_head_stride > 0 ? a.vt_head_stride : a.d;
  static const bool force_v1 = getenv("LDN_ATTN_V1") != nullptr;
  static const int poly_mod = getenv("LDN_ATTN_POLY") ? atoi(getenv("LDN_ATTN_POLY")) : 3;  // 37.5 % of the ex2 on the FMA pipes (measured best)
  p.poly_mod = poly_mod;
  if (a.d == 40 && p.vt_head_stride == 48) {
    // ones-row V^T. (An experiment with two softmax threads per row / 16 softmax warps measured slower -- 2.21 ms vs
    // 1.92 ms at N = 16384 -- and was dropped: the limiter was MMA issue, not softmax latency.)
    // generation 5 (P kept in tensor memory, TS-form P*V); generation 3 (P through shared memory) was retired once v5
    // had replaced it on every path
    // generation 9 (attention9.cu: 64-key steps, two CTAs per SM) for long key sequences: 1.207 ms vs 1.319 ms at N = 16384,
    // B*H = 16 (profiles/r2_attention9.md); generation 5 keeps only key sequences shorter than one 64-key step
    static const int d40_gen = getenv("LDN_ATTN_D40") ? atoi(getenv("LDN_ATTN_D40")) : 9;
    static const int fold_on = getenv("LDN_ATTN_FOLD") ? atoi(getenv("LDN_ATTN_FOLD")) : 1;
    static const int gen9_min_nk = getenv("LDN_ATTN9_MIN_NK") ? atoi(getenv("LDN_ATTN9_MIN_NK")) : 64;  // cross-attention (Nk = 77) included: step 15.86 -> 15.79 ms
    if (d40_gen == 9 && a.Nk >= gen9_min_nk) {
      finish_attn9_plan(plan, a);
      p.fold = (a.fold && fold_on) ? 1 : 0;  // (other kernels ignore the ones column: column 40 of Q is zero in global memory)
    } else {
      finish_attn5_plan(plan, a.Nq, a.Nk, a.heads, a.B);
    }
  } else if (a.d == 80 && p.vt_head_stride == 96 && !a.causal) {
    finish_attn6_plan(plan, a.Nq, a.Nk, a.heads, a.B);  // generation 6: ones-row V^T, P aliased over S in TMEM
  } else if (a.d == 128 && p.vt_head_stride == 128 && !a.causal && !force_v1) {
    finish_attn6_plan(plan, a.Nq, a.Nk, a.heads, a.B);  // generation 6 at d = 128 (Flux): row sum in registers
  } else {
    LDN_CHECK(p.vt_head_stride == a.d, "attention: vt_head_stride is only supported as 48 (d = 40) or 96 (d = 80)");
    if (dp <= 64 && !force_v1 && !a.bias) finish_attn2_plan(plan, a.Nq, a.Nk, a.heads, a.B);
  }
  return plan;
}

template <int DV, bool kBias = false>
static void launch_attn_t(const AttnPlan& plan, cudaStream_t stream) {
  static bool attr_set = false;
  if (!attr_set) {
    LDN_CUDA(cudaFuncSetAttribute(attn_tc_kernel<DV, kBias>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set = true;
  }
  attn_tc_kernel<DV, kBias><<<plan.grid, kAttnThreads, plan.smem_bytes, stream>>>(plan.p);
  LDN_CUDA(cudaGetLastError());
}

void launch_attn(const AttnPlan& plan, cudaStream_t stream) {
  if (plan.p.variant == 6) return launch_attn6(plan, stream);
  if (plan.p.variant == 9) return launch_attn9(plan, stream);
  if (plan.p.variant == 5) return launch_attn5(plan, stream);
  if (plan.p.variant == 2) return launch_attn2(plan, stream);
  if (plan.p.bias) {
    LDN_CHECK(plan.p.dv == 64, "attention: the logit-bias variant is built for head dim 64 only");
    return launch_attn_t<64, true>(plan, stream);
  }
  switch (plan.p.dv) {
    case 48: launch_attn_t<48>(plan, stream); break;
    case 64: launch_attn_t<64>(plan, stream); break;
    case 80: launch_attn_t<80>(plan, stream); break;
    case 128: launch_attn_t<128>(plan, stream); break;
    case 160: launch_attn_t<160>(plan, stream); break;
    default: LDN_CHECK(false, "attention: unsupported dv");
  }
}

}  // namespace ldn
