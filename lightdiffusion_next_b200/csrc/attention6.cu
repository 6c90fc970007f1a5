// Attention kernel, sixth generation, head dims 80 and 128 (80: level-1 self-attention of the SD1.5 UNet, N = 4096 tokens at
// 1024^2; 128: Flux.1 joint attention, N = 4352, src/BlackForest/Flux.py:18-33;
// reference call site CrossAttention.forward src/Attention/Attention.py:100-124 -> attention_pytorch
// src/Attention/AttentionMethods.py:107-134).
//
// Same organisation as generation 5 (attention5.cu): two 128-row query tiles per CTA, one MMA-issuing warp and four
// softmax warps per tile, O and the softmax row sum (ones row of V^T) resident in tensor memory, lazy rescaling,
// 37.5 % of the exponentials evaluated as a degree-3 polynomial with packed FFMA2 / FADD2.
// What changes at d = 80: O needs 96 columns (80 + ones row + pad), so a separate P buffer no longer fits in the 512
// TMEM columns. P (bf16, 64 packed columns) is therefore written over the first half of the S tile it was computed from
// -- every thread has pulled its whole S row into registers before it stores P -- and the next S = Q K^T of the same query
// tile is issued right behind P*V by the same thread, so the tensor pipe's issue order protects P from being overwritten.
// The issue bubble this leaves in one tile's softmax (P*V + next S, ~0.9k clk) is covered by the other tile's softmax.
//
// At d = 128 O takes 128 columns per tile (2 x (128 + 128) = all 512 TMEM columns), so there is no room for a ones row:
// the softmax row sum is accumulated in registers from the fp32 exponentials (packed FADD2) and rescaled with O.
// TMEM columns (d = 80): S0/P0 [0,128)  S1/P1 [128,256)  O0 [256,352)  O1 [352,448);  (d = 128): O0 [256,384)  O1 [384,512).
// Shared memory: Q 2 tiles x 2 atoms x 16 KB; per ring stage K 2 atoms x 16 KB + V^T 2 atoms x 12 KB (96 rows x 64 keys).
#include "common.h"
#include "ptx.cuh"
#include "attn_softmax.cuh"

#include <algorithm>
#include <cstdlib>

namespace ldn {
namespace a6 {
using namespace asm_sm;

static constexpr int kThreads = 384;
static constexpr int kQ = 128;
static constexpr int kK = 128;
// kD = 80: V^T carries 96 rows per head (80 values, row 80 = ones, rows 81..95 = zeros); kD = 128: plain 128-row V^T
static constexpr float kRescaleThreshold = 8.0f;  // log2 units
static constexpr uint32_t kPolyMask = 0x9249u;    // chunks (of 8 keys) whose exponentials run on the FMA pipes

template <int kD>
__global__ void __launch_bounds__(kThreads, 1) attn6_tc_kernel(const __grid_constant__ AttnParams p) {
  constexpr int kDV = kD == 80 ? 96 : 128;
  constexpr bool kOnes = kD == 80;
  constexpr int kQKSteps = kD / 16;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int h = blockIdx.y, b = blockIdx.z;
  const int q0 = blockIdx.x * (2 * kQ);
  const int stages = p.kv_stages;
  constexpr uint32_t atom_bytes = 128 * 128;      // 128 rows x 64 bf16
  constexpr uint32_t q_tile_bytes = 2 * atom_bytes;
  constexpr uint32_t k_bytes = 2 * atom_bytes;
  constexpr uint32_t vt_atom_bytes = kDV * 128;   // 96 rows x 64 keys
  constexpr uint32_t stage_bytes = k_bytes + 2 * vt_atom_bytes;

  uint8_t* q_smem = smem;  // 2 query tiles
  uint8_t* kv_smem = smem + 2 * q_tile_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(kv_smem + (size_t)stages * stage_bytes);
  uint64_t* q_full = bars;       // 1
  uint64_t* s_full = bars + 1;   // [2]
  uint64_t* p_full = bars + 3;   // [2] 128 arrivals
  uint64_t* pv_done = bars + 5;  // [2]
  uint64_t* kv_full = bars + 7;
  uint64_t* kv_empty = kv_full + stages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(kv_empty + stages);
  constexpr uint32_t kTmemCols = 512;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmQ);
    tma_prefetch_desc(&p.tmK);
    tma_prefetch_desc(&p.tmVt);
    mbar_init(q_full, 1);
    for (int t = 0; t < 2; ++t) {
      mbar_init(&s_full[t], 1);
      mbar_init(&p_full[t], 128);
      mbar_init(&pv_done[t], 1);
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
  const int n_tiles = (p.Nk + kK - 1) / kK;

  if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
    if (warp == 0) {
      if (lane == 0) {
        mbar_arrive_expect_tx(q_full, 2 * q_tile_bytes);
        for (int t = 0; t < 2; ++t)
          for (int a = 0; a < 2; ++a)
            tma_load_2d(q_smem + t * q_tile_bytes + a * atom_bytes, &p.tmQ, q_full, h * p.slot + a * 64,
                        b * p.Nq + q0 + t * kQ);
        for (int j = 0; j < n_tiles; ++j) {
          const int s = j % stages;
          const uint32_t ph = (uint32_t)(j / stages) & 1u;
          mbar_wait(&kv_empty[s], ph ^ 1u);
          uint8_t* k_dst = kv_smem + (size_t)s * stage_bytes;
          uint8_t* v_dst = k_dst + k_bytes;
          mbar_arrive_expect_tx(&kv_full[s], stage_bytes);
          const int key0 = b * p.nk_pad + j * kK;
          const int krow0 = b * p.k_batch_stride + j * kK;
          tma_load_2d(k_dst, &p.tmK, &kv_full[s], h * p.slot, krow0);
          tma_load_2d(k_dst + atom_bytes, &p.tmK, &kv_full[s], h * p.slot + 64, krow0);
          tma_load_2d(v_dst, &p.tmVt, &kv_full[s], key0, h * kDV);
          tma_load_2d(v_dst + vt_atom_bytes, &p.tmVt, &kv_full[s], key0 + 64, h * kDV);
        }
      }
    } else if (warp == 1 || warp == 2) {
      // one MMA-issuing warp per query tile; the whole warp walks the loop, one elected lane issues
      const int t = warp - 1;
      const uint32_t idesc_s = make_idesc_bf16(128, 128);
      const uint32_t idesc_pv = make_idesc_bf16(128, kDV);
      const uint32_t tm_s = tmem_base + (uint32_t)t * 128;
      const uint32_t tm_o = tmem_base + 256 + (uint32_t)t * kDV;
      const uint32_t tm_p = tm_s;  // P overwrites the first 64 columns of S
      const uint64_t qd0 = make_smem_desc_sw128(smem_u32(q_smem) + (uint32_t)t * q_tile_bytes);
      const uint64_t kv0 = make_smem_desc_sw128(smem_u32(kv_smem));
      constexpr uint64_t atom_off = atom_bytes >> 4;
      uint64_t* const my_s_full = &s_full[t];
      uint64_t* const my_p_full = &p_full[t];
      mbar_wait(q_full, 0);
      mbar_wait(&kv_full[0], 0);
      tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int ks = 0; ks < kQKSteps; ++ks) {
          const uint64_t off = (ks < 4) ? (uint64_t)(2 * ks) : atom_off + (uint64_t)(2 * (ks - 4));
          tc_mma_bf16(tm_s, qd0 + off, kv0 + off, idesc_s, ks > 0 ? 1u : 0u);
        }
        tc_commit(my_s_full);
      }
      __syncwarp();
      int s = 0;
      for (int j = 0; j < n_tiles; ++j) {
        const uint64_t kd = kv0 + (uint64_t)(((uint32_t)s * stage_bytes) >> 4);
        const uint64_t vd0 = kd + (uint64_t)(k_bytes >> 4);
        const uint64_t vd1 = vd0 + (uint64_t)(vt_atom_bytes >> 4);
        const int s1 = (s + 1 == stages) ? 0 : s + 1;
        mbar_wait(my_p_full, (uint32_t)j & 1u);
        tc_fence_after();
        if (elect_one()) {
          // O (+)= P V: A = P from TMEM, k-step ks covers keys [16 ks, 16 ks + 16) = 8 packed columns
          tc_mma_bf16_ts(tm_o, tm_p + 0, vd0, idesc_pv, j > 0 ? 1u : 0u);
          tc_mma_bf16_ts(tm_o, tm_p + 8, vd0 + 2, idesc_pv, 1u);
          tc_mma_bf16_ts(tm_o, tm_p + 16, vd0 + 4, idesc_pv, 1u);
          tc_mma_bf16_ts(tm_o, tm_p + 24, vd0 + 6, idesc_pv, 1u);
          tc_mma_bf16_ts(tm_o, tm_p + 32, vd1, idesc_pv, 1u);
          tc_mma_bf16_ts(tm_o, tm_p + 40, vd1 + 2, idesc_pv, 1u);
          tc_mma_bf16_ts(tm_o, tm_p + 48, vd1 + 4, idesc_pv, 1u);
          tc_mma_bf16_ts(tm_o, tm_p + 56, vd1 + 6, idesc_pv, 1u);
          tc_commit(&pv_done[t]);
          tc_commit(&kv_empty[s]);  // 2 arrivals per stage: one from each tile's issuing thread
        }
        __syncwarp();
        if (j + 1 < n_tiles) {
          mbar_wait(&kv_full[s1], (uint32_t)((j + 1) / stages) & 1u);
          tc_fence_after();
          const uint64_t kn = kv0 + (uint64_t)(((uint32_t)s1 * stage_bytes) >> 4);
          if (elect_one()) {
            // issued behind P*V by the same thread: executes after P has been consumed
#pragma unroll
            for (int ks = 0; ks < kQKSteps; ++ks) {
              const uint64_t off = (ks < 4) ? (uint64_t)(2 * ks) : atom_off + (uint64_t)(2 * (ks - 4));
              tc_mma_bf16(tm_s, qd0 + off, kn + off, idesc_s, ks > 0 ? 1u : 0u);
            }
            tc_commit(my_s_full);
          }
          __syncwarp();
        }
        s = s1;
      }
    }
  } else {
    // ---------------------------------------------------------------- softmax warpgroups
    asm volatile("setmaxnreg.inc.sync.aligned.u32 224;");
    const int t = (warp - 4) >> 2;
    const int qd = warp & 3;
    const int r = qd * 32 + lane;
    const uint32_t lane_off = (uint32_t)(qd * 32) << 16;
    const uint32_t tmem_s = tmem_base + (uint32_t)t * 128 + lane_off;
    const uint32_t tmem_o = tmem_base + 256 + (uint32_t)t * kDV + lane_off;
    const uint32_t tmem_p = tmem_s;
    const int q_idx = q0 + t * kQ + r;
    const float sc = p.scale_log2;
    uint64_t* const my_s_full = &s_full[t];
    uint64_t* const my_p_full = &p_full[t];
    uint64_t* const my_pv_done = &pv_done[t];
    float m_used = 0.f;  // exponent offset currently baked into O (scaled log2 units)
    float l_run = 0.f;   // softmax row sum (kept in registers when V^T has no ones row)

    for (int j = 0; j < n_tiles; ++j) {
      mbar_wait(my_s_full, (uint32_t)j & 1u);  // S(j) landed -- and, by issue order, P*V(j-1) finished too
      tc_fence_after();
      uint32_t sv[128];
      tmem_ld32(tmem_s + 0, sv + 0);
      tmem_ld32(tmem_s + 32, sv + 32);
      tmem_ld32(tmem_s + 64, sv + 64);
      tmem_ld32(tmem_s + 96, sv + 96);
      tmem_ld_wait();

      const int limit = p.Nk - j * kK;
      if (limit < kK) {
#pragma unroll
        for (int i = 0; i < 128; ++i)
          if (i >= limit) sv[i] = 0xff800000u;  // -inf
      }
      float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
      for (int i = 0; i < 128; i += 8) {
        mx0 = fmaxf(mx0, fmaxf(__uint_as_float(sv[i]), __uint_as_float(sv[i + 1])));
        mx1 = fmaxf(mx1, fmaxf(__uint_as_float(sv[i + 2]), __uint_as_float(sv[i + 3])));
        mx2 = fmaxf(mx2, fmaxf(__uint_as_float(sv[i + 4]), __uint_as_float(sv[i + 5])));
        mx3 = fmaxf(mx3, fmaxf(__uint_as_float(sv[i + 6]), __uint_as_float(sv[i + 7])));
      }
      const float mx = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3)) * sc;
      if (j == 0) {
        m_used = mx;
      } else {
        // lazy rescale: only when this row's max outgrew the offset baked into O by more than 2^8
        const bool need = mx > m_used + kRescaleThreshold;
        if (__any_sync(0xffffffffu, need)) {
          const float m_new = need ? mx : m_used;
          const float f = ex2m(m_used - m_new);  // 1 for rows that do not need it
          m_used = m_new;
          l_run *= f;
#pragma unroll
          for (int c = 0; c < kDV; c += 16) {
            uint32_t v[16];
            tmem_ld16(tmem_o + (uint32_t)c, v);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * f);
            tmem_st16(tmem_o + (uint32_t)c, v);
          }
          tmem_st_wait();
        }
      }
      const float m_off = m_used;
#pragma unroll
      float2 lacc = make_float2(0.f, 0.f);
#pragma unroll
      for (int c = 0; c < 16; ++c) {
        uint32_t w[4];
        if (kOnes) exp8_pack(sv + c * 8, sc, -m_off, (int)((kPolyMask >> c) & 1u), w);
        else exp8_pack_sum(sv + c * 8, sc, -m_off, (int)((kPolyMask >> c) & 1u), w, lacc);
        asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};" ::"r"(tmem_p + (uint32_t)(c * 4)), "r"(w[0]),
                     "r"(w[1]), "r"(w[2]), "r"(w[3])
                     : "memory");
      }
      if (!kOnes) l_run += lacc.x + lacc.y;
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(my_p_full);
    }
    // epilogue: O[:, 0:80] / O[:, 80]
    if (n_tiles > 0) {
      mbar_wait(my_pv_done, (uint32_t)(n_tiles - 1) & 1u);
      tc_fence_after();
      uint32_t v[kDV];
#pragma unroll
      for (int c = 0; c < kDV; c += 16) tmem_ld16(tmem_o + (uint32_t)c, v + c);
      tmem_ld_wait();
      if (q_idx < p.Nq) {
        const float l = kOnes ? __uint_as_float(v[kOnes ? kD : 0]) : l_run;
        const float inv = l > 0.f ? 1.f / l : 0.f;
        bf16* orow = p.out + ((long long)b * p.Nq + q_idx) * p.ldo + (long long)h * kD;
#pragma unroll
        for (int c = 0; c < kD; c += 8) {
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

}  // namespace a6

template <int kD>
static void launch_attn6_t(const AttnPlan& plan, cudaStream_t stream) {
  static bool attr_set = false;
  if (!attr_set) {
    LDN_CUDA(cudaFuncSetAttribute(a6::attn6_tc_kernel<kD>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set = true;
  }
  a6::attn6_tc_kernel<kD><<<plan.grid, a6::kThreads, plan.smem_bytes, stream>>>(plan.p);
  LDN_CUDA(cudaGetLastError());
}

void launch_attn6(const AttnPlan& plan, cudaStream_t stream) {
  if (plan.p.d == 128) launch_attn6_t<128>(plan, stream);
  else launch_attn6_t<80>(plan, stream);
}

void finish_attn6_plan(AttnPlan& plan, int Nq, int Nk, int heads, int B) {
  AttnParams& p = plan.p;
  LDN_CHECK(!p.causal && ((p.d == 80 && p.vt_head_stride == 96) || (p.d == 128 && p.vt_head_stride == 128)),
            "attention6: d = 80 with 96-row V^T heads or d = 128, non-causal only");
  const int dv = p.d == 80 ? 96 : 128;
  const int stage_bytes = 2 * 16384 + 2 * dv * 128;
  const int fixed = 4 * 16384 + 1024 + 512;
  const int n_tiles = (Nk + a6::kK - 1) / a6::kK;
  int stages = (226 * 1024 - fixed) / stage_bytes;
  if (stages > 4) stages = 4;
  if (stages > n_tiles) stages = n_tiles;
  if (stages < 1) stages = 1;
  p.kv_stages = stages;
  p.variant = 6;
  plan.smem_bytes = fixed + stages * stage_bytes;
  plan.grid = dim3((Nq + 2 * a6::kQ - 1) / (2 * a6::kQ), heads, B);
}

}  // namespace ldn
