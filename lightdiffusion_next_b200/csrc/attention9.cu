// Attention kernel, ninth generation, head dim 40, long key sequences (level-0 self-attention of the SD1.5 UNet at
// 1024^2: N = 16384, 39 % of the sampler step).
//
// Generation 5 (attention5.cu) is bound by neither pipe: per 128 x 128 score tile it needs ~1450 clk against ~640 clk of
// MUFU and ~700 issue slots per scheduler (profiles/r1_attention_ncu_full.md) -- with ONE CTA per SM there are only two
// softmax warps per scheduler, and when both sit in a dependent stretch (TMEM round trip, max chain, MUFU results) the
// scheduler idles (issue slots 55 % busy, `not_selected` 20 % + stalls 49 % of the samples).  Splitting a row over two
// threads (generation 7) and software-pipelining one thread (generation 8) both lost to their own overhead
// (profiles/r2_experiments.md).  This generation keeps generation 5's thread-per-row softmax but halves every per-CTA
// resource so that TWO CTAs are resident per SM -- four independent softmax warps per scheduler that drift apart naturally:
//   * 64-key steps: S is 128 x 64 fp32 = 64 TMEM columns per query tile; P (bf16, two keys per 32-bit cell) is written
//     OVER the first 32 columns of the S it was computed from (generation 6's aliasing), O takes 48 columns:
//     2 tiles x (64 + 64) = 256 TMEM columns per CTA;
//   * K / V^T arrive as 64-key TMA stages of 14 KB (five of them + the two 16 KB query tiles = 102 KB of shared memory);
//   * the MMA-issuing thread of a tile issues P*V of step j and, right behind it, S of step j + 1 into the columns P*V is
//     still reading (tcgen05.mma of one thread execute in order), so no S-free / P-free barriers exist: per step and tile
//     there is one commit (S ready) and one 128-thread arrival (P ready);
//   * 96 registers per softmax thread (64 scores + packed probabilities in flight) via setmaxnreg.
// Everything else is generation 5: one MMA-issuing warp per query tile, TS-form P*V, the ones row of V^T producing the
// row sums, lazy rescale of O (threshold 2^8), a share of the exponentials as a degree-3 polynomial on the FMA pipes,
// packed fp32-pair arithmetic.
#include "common.h"
#include "ptx.cuh"
#include "attn_softmax.cuh"

#include <algorithm>
#include <cstdlib>
#include <type_traits>

namespace ldn {
namespace a9 {

static constexpr int kThreads = 384;
static constexpr int kQ = 128;     // query rows per tile (2 tiles per CTA)
static constexpr int kStep = 64;   // keys per stage / softmax step
static constexpr int kDV = 48;     // 40 value rows + ones row + 7 zero rows
static constexpr float kRescaleThreshold = 8.0f;  // log2 units

using namespace asm_sm;

template <uint32_t kPolyMask, bool kFold>
__global__ void __launch_bounds__(kThreads, 2) attn9_tc_kernel(const __grid_constant__ AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int h = blockIdx.y, b = blockIdx.z;
  const int q0 = blockIdx.x * (2 * kQ);
  const int stages = p.kv_stages;
  constexpr uint32_t q_bytes = 128 * 128;           // 128 rows x 64 bf16 (128B-swizzled)
  constexpr uint32_t k_bytes = kStep * 128;         // 64 keys x 64 bf16
  constexpr uint32_t vt_bytes = kDV * 128;          // 48 rows x 64 keys
  constexpr uint32_t stage_bytes = k_bytes + vt_bytes;

  uint8_t* q_smem = smem;  // 2 query tiles
  uint8_t* kv_smem = smem + 2 * q_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(kv_smem + (size_t)stages * stage_bytes);
  uint64_t* q_full = bars;          // 1
  uint64_t* s_full = bars + 1;      // [tile] one commit per step: S(j) is in TMEM (and every earlier P*V has landed in O)
  uint64_t* p_full = bars + 3;      // [tile] 128 arrivals per step: P(j) is in TMEM
  uint64_t* o_done = bars + 5;      // [tile] one commit: the last P*V has landed
  uint64_t* kv_full = bars + 7;
  uint64_t* kv_empty = kv_full + stages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(kv_empty + stages);
  constexpr uint32_t kTmemCols = 256;  // S / P of tile t: [t*64, +64); O of tile t: [128 + t*64, +48)

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmQ);
    tma_prefetch_desc(&p.tmK);
    tma_prefetch_desc(&p.tmVt);
    mbar_init(q_full, 1);
    for (int t = 0; t < 2; ++t) {
      mbar_init(&s_full[t], 1);
      mbar_init(&p_full[t], 128);
      mbar_init(&o_done[t], 1);
    }
    for (int s = 0; s < stages; ++s) {
      mbar_init(&kv_full[s], 1);
      mbar_init(&kv_empty[s], 2);
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
  const int n_steps = (p.Nk + kStep - 1) / kStep;

  if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
    if (warp == 0) {
      if (lane == 0) {
        mbar_arrive_expect_tx(q_full, 2 * q_bytes);
        tma_load_2d(q_smem, &p.tmQ, q_full, h * p.slot, b * p.Nq + q0);
        tma_load_2d(q_smem + q_bytes, &p.tmQ, q_full, h * p.slot, b * p.Nq + q0 + kQ);
        int s = 0;
        uint32_t ph = 0;
        for (int j = 0; j < n_steps; ++j) {
          mbar_wait(&kv_empty[s], ph ^ 1u);
          uint8_t* k_dst = kv_smem + (size_t)s * stage_bytes;
          mbar_arrive_expect_tx(&kv_full[s], stage_bytes);
          tma_load_2d(k_dst, &p.tmK, &kv_full[s], h * p.slot, b * p.k_batch_stride + j * kStep);
          tma_load_2d(k_dst + k_bytes, &p.tmVt, &kv_full[s], b * p.nk_pad + j * kStep, h * kDV);
          if (++s == stages) {
            s = 0;
            ph ^= 1u;
          }
        }
      }
    } else if (warp == 1 || warp == 2) {
      // one MMA-issuing warp per query tile; the whole warp walks the loop (uniform registers), one elected lane issues
      const int t = warp - 1;
      const uint32_t idesc_s = make_idesc_bf16(128, kStep);
      const uint32_t idesc_pv = make_idesc_bf16(128, kDV);
      const uint32_t tm_s = tmem_base + (uint32_t)t * 64;   // S, and P over its first 32 columns
      const uint32_t tm_o = tmem_base + 128 + (uint32_t)t * 64;
      const uint64_t qd0 = make_smem_desc_sw128(smem_u32(q_smem) + (uint32_t)t * q_bytes);
      const uint64_t kv0 = make_smem_desc_sw128(smem_u32(kv_smem));  // stage 0's K tile; stages / V^T are offsets
      mbar_wait(q_full, 0);
      mbar_wait(&kv_full[0], 0);
      tc_fence_after();
      if (elect_one()) {
        tc_mma_bf16(tm_s, qd0, kv0, idesc_s, 0u);
        tc_mma_bf16(tm_s, qd0 + 2, kv0 + 2, idesc_s, 1u);
        tc_mma_bf16(tm_s, qd0 + 4, kv0 + 4, idesc_s, 1u);
        tc_commit(&s_full[t]);
      }
      __syncwarp();
      int s = 0;
      uint32_t ph = 0;
      for (int j = 0; j < n_steps; ++j) {
        const uint64_t vd = kv0 + (uint64_t)(((uint32_t)s * stage_bytes + k_bytes) >> 4);
        int s1 = s + 1;
        uint32_t ph1 = ph;
        if (s1 == stages) {
          s1 = 0;
          ph1 ^= 1u;
        }
        const bool more = j + 1 < n_steps;
        if (more) mbar_wait(&kv_full[s1], ph1);  // normally long complete
        mbar_wait(&p_full[t], (uint32_t)j & 1u);
        tc_fence_after();
        if (elect_one()) {
          // O += P V: A = P from TMEM (k-step ks = keys [16 ks, 16 ks + 16) = 8 packed columns), B = this stage's V^T atom
          tc_mma_bf16_ts(tm_o, tm_s + 0, vd, idesc_pv, j > 0 ? 1u : 0u);
          tc_mma_bf16_ts(tm_o, tm_s + 8, vd + 2, idesc_pv, 1u);
          tc_mma_bf16_ts(tm_o, tm_s + 16, vd + 4, idesc_pv, 1u);
          tc_mma_bf16_ts(tm_o, tm_s + 24, vd + 6, idesc_pv, 1u);
          tc_commit(&kv_empty[s]);  // 2 arrivals per stage: one from each tile's issuing thread
          if (more) {
            // S of the next step, into the columns P*V above is still reading: MMAs of one thread execute in issue order
            const uint64_t kn = kv0 + (uint64_t)(((uint32_t)s1 * stage_bytes) >> 4);
            tc_mma_bf16(tm_s, qd0, kn, idesc_s, 0u);
            tc_mma_bf16(tm_s, qd0 + 2, kn + 2, idesc_s, 1u);
            tc_mma_bf16(tm_s, qd0 + 4, kn + 4, idesc_s, 1u);
            tc_commit(&s_full[t]);
          } else {
            tc_commit(&o_done[t]);
          }
        }
        __syncwarp();
        s = s1;
        ph = ph1;
      }
    }
  } else {
    // ---------------------------------------------------------------- softmax: one thread per score row, 64 scores per step
    asm volatile("setmaxnreg.inc.sync.aligned.u32 96;");
    const int t = (warp - 4) >> 2;
    const int qd = warp & 3;
    const int r = qd * 32 + lane;
    const uint32_t lane_off = (uint32_t)(qd * 32) << 16;
    const uint32_t tmem_s = tmem_base + (uint32_t)t * 64 + lane_off;
    const uint32_t tmem_o = tmem_base + 128 + (uint32_t)t * 64 + lane_off;
    const int q_idx = q0 + t * kQ + r;
    const float sc = p.scale_log2;
    uint64_t* const my_s_full = &s_full[t];
    uint64_t* const my_p_full = &p_full[t];
    float m_used = 0.f;  // exponent offset currently baked into O (scaled log2 units)

    // one 64-key step; `mask_tag` carries the compile-time polynomial mask of the step's eight 8-score chunks
    auto step = [&](const int j, auto mask_tag) {
      constexpr uint32_t mask8 = decltype(mask_tag)::value;
      mbar_wait(my_s_full, (uint32_t)j & 1u);
      tc_fence_after();
      uint32_t sv[64];
      tmem_ld32(tmem_s + 0, sv + 0);
      tmem_ld32(tmem_s + 32, sv + 32);
      tmem_ld_wait();
      const int limit = p.Nk - j * kStep;
      if (limit < kStep) {
#pragma unroll
        for (int i = 0; i < 64; ++i)
          if (i >= limit) sv[i] = 0xff800000u;  // -inf
      }
      float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
      for (int i = 0; i < 64; i += 8) {
        mx0 = fmaxf(mx0, fmaxf(__uint_as_float(sv[i]), __uint_as_float(sv[i + 1])));
        mx1 = fmaxf(mx1, fmaxf(__uint_as_float(sv[i + 2]), __uint_as_float(sv[i + 3])));
        mx2 = fmaxf(mx2, fmaxf(__uint_as_float(sv[i + 4]), __uint_as_float(sv[i + 5])));
        mx3 = fmaxf(mx3, fmaxf(__uint_as_float(sv[i + 6]), __uint_as_float(sv[i + 7])));
      }
      if constexpr (kFold) {
        // Folded variant: the MMA already delivered s' = s * scale * log2(e) - m_used (Q is pre-scaled; column 40 of the Q
        // tile holds -m_used of the row and column 40 of K holds ones), so the common path has no scale-subtract at all.
        // When the row maximum outgrows the offset (always at step 0, where m_used = 0), the offset is moved: subtract the
        // (bf16-exact) increment from the scores in registers, rescale O, and publish the new -m to the Q tile in shared
        // memory for the S MMAs to come (none is in flight: S(j + 1) is issued behind this thread's arrival on p_full).
        const float mx = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));
        const bool upd = (j == 0) || (mx > kRescaleThreshold);
        if (__any_sync(0xffffffffu, upd)) {
          const float m_new = upd ? __bfloat162float(__float2bfloat16_rn(m_used + mx)) : m_used;
          const float delta = m_new - m_used;  // exact: both are bf16 values of similar magnitude
          m_used = m_new;
          if (j > 0) {
            const float f = ex2m(-delta);  // 1 for rows that do not move
#pragma unroll
            for (int c = 0; c < kDV; c += 16) {
              uint32_t v[16];
              tmem_ld16(tmem_o + (uint32_t)c, v);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 16; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * f);
              tmem_st16(tmem_o + (uint32_t)c, v);
            }
          }
          const float2 nd = make_float2(-delta, -delta);
#pragma unroll
          for (int i = 0; i < 64; i += 2) {
            const float2 e = fadd2(make_float2(__uint_as_float(sv[i]), __uint_as_float(sv[i + 1])), nd);
            sv[i] = __float_as_uint(e.x);
            sv[i + 1] = __float_as_uint(e.y);
          }
          // row r of the 128B-swizzled tile: 16-byte chunk 5 (columns 40..47) lands at chunk 5 ^ (r % 8)
          const __nv_bfloat16 nm = __float2bfloat16_rn(-m_new);
          *reinterpret_cast<__nv_bfloat16*>(q_smem + (size_t)t * q_bytes + (size_t)r * 128 + (size_t)((5 ^ (r & 7)) << 4)) = nm;
          fence_proxy_async_smem();
        }
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          uint32_t w[4];
          exp8_pack_raw(sv + c * 8, kPolyMask == 0x10000u ? 2 : (int)((mask8 >> c) & 1u), w);
          asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};" ::"r"(tmem_s + (uint32_t)(c * 4)), "r"(w[0]),
                       "r"(w[1]), "r"(w[2]), "r"(w[3])
                       : "memory");
        }
      } else {
        const float mx = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3)) * sc;
        if (j == 0) {
          m_used = mx;
        } else {
          // lazy rescale: only when this row's max outgrew the offset baked into O by more than 2^8.  S(j) was committed
          // behind P*V(j - 1), so every earlier P*V has landed and none is in flight until this thread arrives on p_full.
          const bool need = mx > m_used + kRescaleThreshold;
          if (__any_sync(0xffffffffu, need)) {
            const float m_new = need ? mx : m_used;
            const float f = ex2m(m_used - m_new);  // 1 for rows that do not need it
            m_used = m_new;
#pragma unroll
            for (int c = 0; c < kDV; c += 16) {
              uint32_t v[16];
              tmem_ld16(tmem_o + (uint32_t)c, v);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 16; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * f);
              tmem_st16(tmem_o + (uint32_t)c, v);
            }
          }
        }
        const float m_off = m_used;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          uint32_t w[4];
          exp8_pack(sv + c * 8, sc, -m_off, kPolyMask == 0x10000u ? 2 : (int)((mask8 >> c) & 1u), w);
          asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};" ::"r"(tmem_s + (uint32_t)(c * 4)), "r"(w[0]),
                       "r"(w[1]), "r"(w[2]), "r"(w[3])
                       : "memory");
        }
      }
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(my_p_full);
    };
    for (int j = 0; j < n_steps; j += 2) {
      step(j, std::integral_constant<uint32_t, (kPolyMask & 0xffu)>{});
      if (j + 1 < n_steps) step(j + 1, std::integral_constant<uint32_t, ((kPolyMask >> 8) & 0xffu)>{});
    }
    // epilogue: O[:, 0:40] / O[:, 40]
    if (n_steps > 0) {
      mbar_wait(&o_done[t], 0);
      tc_fence_after();
      uint32_t v[48];
      tmem_ld16(tmem_o + 0, v + 0);
      tmem_ld16(tmem_o + 16, v + 16);
      tmem_ld16(tmem_o + 32, v + 32);
      tmem_ld_wait();
      if (q_idx < p.Nq) {
        const float l = __uint_as_float(v[40]);
        const float inv = l > 0.f ? 1.f / l : 0.f;
        bf16* orow = p.out + ((long long)b * p.Nq + q_idx) * p.ldo + (long long)h * 40;
#pragma unroll
        for (int c = 0; c < 40; c += 8) {
          uint4 ov;
          ov.x = pack_bf16x2(__uint_as_float(v[c + 0]) * inv, __uint_as_float(v[c + 1]) * inv);
          ov.y = pack_bf16x2(__uint_as_float(v[c + 2]) * inv, __uint_as_float(v[c + 3]) * inv);
          ov.z = pack_bf16x2(__uint_as_float(v[c + 4]) * inv, __uint_as_float(v[c + 5]) * inv);
          ov.w = pack_bf16x2(__uint_as_float(v[c + 6]) * inv, __uint_as_float(v[c + 7]) * inv);
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

}  // namespace a9
using namespace a9;

template <uint32_t kPolyMask, bool kFold>
static void launch_attn9_t2(const AttnPlan& plan, cudaStream_t stream) {
  static bool attr_set = false;
  if (!attr_set) {
    LDN_CUDA(cudaFuncSetAttribute(attn9_tc_kernel<kPolyMask, kFold>, cudaFuncAttributeMaxDynamicSharedMemorySize, 113 * 1024));
    attr_set = true;
  }
  attn9_tc_kernel<kPolyMask, kFold><<<plan.grid, kThreads, plan.smem_bytes, stream>>>(plan.p);
  LDN_CUDA(cudaGetLastError());
}
template <uint32_t kPolyMask>
static void launch_attn9_t(const AttnPlan& plan, cudaStream_t stream) {
  if (plan.p.fold)
    launch_attn9_t2<kPolyMask, true>(plan, stream);
  else
    launch_attn9_t2<kPolyMask, false>(plan, stream);
}

// kPolyMask: bit c (even steps) / bit 8 + c (odd steps) set = chunk c of a step's 64 scores takes the polynomial ex2
void launch_attn9(const AttnPlan& plan, cudaStream_t stream) {
  switch (plan.p.poly_mod) {
    case 2: return launch_attn9_t<0xAAAAu>(plan, stream);  // 50 % polynomial
    case 3: return launch_attn9_t<0x9249u>(plan, stream);  // 37.5 %
    case 4: return launch_attn9_t<0x8888u>(plan, stream);  // 25 %
    case 8: return launch_attn9_t<0x8080u>(plan, stream);  // 12.5 %
    case 99: return launch_attn9_t<0x10000u>(plan, stream);  // experiment: no exponential (wrong results)
    default: return launch_attn9_t<0u>(plan, stream);
  }
}

void finish_attn9_plan(AttnPlan& plan, const AttnArgs& a) {
  AttnParams& p = plan.p;
  LDN_CHECK(p.d == 40 && p.dv == 48 && p.dqk == 48 && !p.causal, "attention9: d = 40, non-causal only");
  // K arrives in 64-key boxes here (generation 5 uses 128-key boxes)
  p.tmK = make_tmap_2d(a.K, (uint64_t)a.B * p.k_batch_stride, (uint64_t)a.heads * a.slot, a.ldk, kStep);
  const int stage_bytes = kStep * 128 + kDV * 128;
  const int fixed = 2 * 16384 + 1024 + 512;
  const int n_steps = (a.Nk + kStep - 1) / kStep;
  int stages = (112 * 1024 - fixed) / stage_bytes;  // two CTAs per SM
  if (stages > 6) stages = 6;
  if (stages > n_steps) stages = n_steps;
  if (getenv("LDN_ATTN_STAGES")) stages = std::min(stages, atoi(getenv("LDN_ATTN_STAGES")));
  if (stages < 2 && n_steps >= 2) stages = 2;
  if (stages < 1) stages = 1;
  p.kv_stages = stages;
  p.variant = 9;
  plan.smem_bytes = fixed + stages * stage_bytes;
  plan.grid = dim3((a.Nq + 2 * kQ - 1) / (2 * kQ), a.heads, a.B);
}

}  // namespace ldn
