// Attention kernel, seventh generation, head dim 40 (level-0 self-attention of the SD1.5 UNet: 42 % of the step's FLOPs).
//
// Same data path as generation 5 (attention5.cu: two 128-row query tiles per CTA share every K / V^T tile, S, O and the
// bf16 probabilities P live in tensor memory, P*V is issued in the TS form, a ones row in V^T makes the tensor core
// produce the softmax row sums, O is rescaled lazily), but the softmax is spread over TWICE as many threads:
// every score row of a 128-key tile is split between two threads (keys [0, 64) and [64, 128)) that sit in two different
// warps of the same TMEM lane quarter.  Generation 5 ran 2 softmax warps per scheduler, each holding 128 scores per
// thread in registers; its profile (profiles/r1_attention_ncu_full.md) showed the schedulers issuing 55 % of the cycles
// with the MUFU at 49 % and the tensor pipe at 29 % -- neither pipe saturated, the loop was bound by dependent-instruction
// latency with too few warps to hide it.  Here 4 softmax warps per scheduler hold 64 scores each (<= 104 registers), so
// one warp's TMEM load / max reduction / barrier waits overlap the other warps' exponentials.  The two halves of a row
// agree on the running maximum through a 4 KB shared-memory exchange and a 64-thread named barrier per key tile.
#include "common.h"
#include "ptx.cuh"
#include "attn_softmax.cuh"

#include <algorithm>
#include <cstdlib>

namespace ldn {
namespace a7 {

static constexpr int kA7Threads = 640;  // 4 helper warps + 16 softmax warps
static constexpr int kQ3 = 128;
static constexpr int kK3 = 128;
static constexpr int kDV3 = 48;
static constexpr float kRescaleThreshold = 8.0f;  // log2 units

using namespace asm_sm;

template <uint32_t kPolyMask>
__global__ void __launch_bounds__(kA7Threads, 1) attn7_tc_kernel(const __grid_constant__ AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int h = blockIdx.y, b = blockIdx.z;
  const int q0 = blockIdx.x * (2 * kQ3);
  const int stages = p.kv_stages;
  constexpr uint32_t atom_bytes = 128 * 128;
  constexpr uint32_t vt_atom_bytes = kDV3 * 128;
  constexpr uint32_t stage_bytes = atom_bytes + 2 * vt_atom_bytes;

  uint8_t* q_smem = smem;                       // 2 query tiles
  uint8_t* kv_smem = smem + 2 * atom_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(kv_smem + (size_t)stages * stage_bytes);
  uint64_t* q_full = bars;         // 1
  uint64_t* s_full = bars + 1;     // [2]
  uint64_t* s_free = bars + 3;     // [2] 128 arrivals
  uint64_t* p_full = bars + 5;     // [2] 128 arrivals
  uint64_t* pv_done = bars + 7;    // [2] one completion per key tile
  uint64_t* p_free = bars + 9;     // [2 tiles][2 parities] one completion per use of the P buffer
  uint64_t* kv_full = bars + 13;
  uint64_t* kv_empty = kv_full + stages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(kv_empty + stages);
  float* half_max = reinterpret_cast<float*>(tmem_slot + 4);  // [2 parities][2 tiles][2 key halves][128 rows]
  constexpr uint32_t kTmemCols = 512;  // S0 [0,128) S1 [128,256) O0 [256,304) O1 [320,368) P0 [384,448) P1 [448,512)

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmQ);
    tma_prefetch_desc(&p.tmK);
    tma_prefetch_desc(&p.tmVt);
    mbar_init(q_full, 1);
    for (int t = 0; t < 2; ++t) {
      mbar_init(&s_full[t], 1);
      mbar_init(&s_free[t], 256);
      mbar_init(&p_full[t], 256);
      mbar_init(&pv_done[t], 1);
      mbar_init(&p_free[2 * t], 1);
      mbar_init(&p_free[2 * t + 1], 1);
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
  const int n_tiles = (p.Nk + kK3 - 1) / kK3;

  if (warp < 4) {
    // Register budget: the CTA is launched with 96 registers x 640 threads = 61440; setmaxnreg can only hand out what other
    // warps released (it never draws on unallocated file space), so 128 x 48 + 512 x 104 = 59392 must stay below that --
    // asking for 112 here deadlocked the last softmax warps at their setmaxnreg.inc.
    asm volatile("setmaxnreg.dec.sync.aligned.u32 48;");
    if (warp == 0) {
      if (lane == 0) {
        mbar_arrive_expect_tx(q_full, 2 * atom_bytes);
        tma_load_2d(q_smem, &p.tmQ, q_full, h * p.slot, b * p.Nq + q0);
        tma_load_2d(q_smem + atom_bytes, &p.tmQ, q_full, h * p.slot, b * p.Nq + q0 + kQ3);
        for (int j = 0; j < n_tiles; ++j) {
          const int s = j % stages;
          const uint32_t ph = (uint32_t)(j / stages) & 1u;
          mbar_wait(&kv_empty[s], ph ^ 1u);
          uint8_t* k_dst = kv_smem + (size_t)s * stage_bytes;
          uint8_t* v_dst = k_dst + atom_bytes;
          mbar_arrive_expect_tx(&kv_full[s], stage_bytes);
          const int key0 = b * p.nk_pad + j * kK3;
          const int krow0 = b * p.k_batch_stride + j * kK3;
          tma_load_2d(k_dst, &p.tmK, &kv_full[s], h * p.slot, krow0);
          tma_load_2d(v_dst, &p.tmVt, &kv_full[s], key0, h * kDV3);
          tma_load_2d(v_dst + vt_atom_bytes, &p.tmVt, &kv_full[s], key0 + 64, h * kDV3);
        }
      }
    } else if (warp == 1 || warp == 2) {
      // One MMA-issuing thread PER QUERY TILE (warp 1 -> tile 0, warp 2 -> tile 1). A resource-binding experiment
      // (exponentials removed: same run time) showed that a single issuing thread was the bottleneck of generations
      // 1-3: at d = 40 the MMAs are tiny (N = 48: 24 tensor-clk each), so descriptor arithmetic + issue latency of 22 MMAs
      // per key tile on ONE thread cost more than the softmax. All descriptors that do not depend on the ring stage
      // are built once, the per-stage ones with a single 32-bit add.
      {
        // the whole warp walks the loop (warp-uniform values stay in uniform registers); one elected lane issues
        const int t = warp - 1;
        const uint32_t idesc_s = make_idesc_bf16(128, 128);
        const uint32_t idesc_pv = make_idesc_bf16(128, kDV3);
        const uint32_t kv_addr = smem_u32(kv_smem);
        const uint32_t tm_s = tmem_base + (uint32_t)t * 128;
        const uint32_t tm_o = tmem_base + 256 + (uint32_t)t * 64;
        const uint64_t qd0 = make_smem_desc_sw128(smem_u32(q_smem) + (uint32_t)t * atom_bytes);
        const uint32_t tm_p = tmem_base + 384 + (uint32_t)t * 64;
        const uint64_t kv0 = make_smem_desc_sw128(kv_addr);  // descriptor of stage 0's K tile; stages / V^T are offsets
        uint64_t* const my_s_full = &s_full[t];
        uint64_t* const my_s_free = &s_free[t];
        uint64_t* const my_p_full = &p_full[t];
        mbar_wait(q_full, 0);
        mbar_wait(&kv_full[0], 0);
        tc_fence_after();
        if (elect_one()) {
          tc_mma_bf16(tm_s, qd0, kv0, idesc_s, 0u);
          tc_mma_bf16(tm_s, qd0 + 2, kv0 + 2, idesc_s, 1u);
          tc_mma_bf16(tm_s, qd0 + 4, kv0 + 4, idesc_s, 1u);
          tc_commit(my_s_full);
        }
        __syncwarp();
        int s = 0;
        for (int j = 0; j < n_tiles; ++j) {
          const uint64_t kd = kv0 + (uint64_t)(((uint32_t)s * stage_bytes) >> 4);
          const uint64_t vd0 = kd + (uint64_t)(atom_bytes >> 4);
          const uint64_t vd1 = vd0 + (uint64_t)(vt_atom_bytes >> 4);
          if (j + 1 < n_tiles) {
            const int s1 = (s + 1 == stages) ? 0 : s + 1;
            mbar_wait(&kv_full[s1], (uint32_t)((j + 1) / stages) & 1u);
            const uint64_t kn = kv0 + (uint64_t)(((uint32_t)s1 * stage_bytes) >> 4);
            mbar_wait(my_s_free, (uint32_t)j & 1u);
            tc_fence_after();
            if (elect_one()) {
              tc_mma_bf16(tm_s, qd0, kn, idesc_s, 0u);
              tc_mma_bf16(tm_s, qd0 + 2, kn + 2, idesc_s, 1u);
              tc_mma_bf16(tm_s, qd0 + 4, kn + 4, idesc_s, 1u);
              tc_commit(my_s_full);
            }
            __syncwarp();
          }
          mbar_wait(my_p_full, (uint32_t)j & 1u);
          tc_fence_after();
          if (elect_one()) {
          // A = P from TMEM: k-step ks covers keys [16 ks, 16 ks + 16) = 8 packed columns
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
          s = (s + 1 == stages) ? 0 : s + 1;
        }
      }
    }
  } else {
    // ---------------------------------------------------------------- softmax: 16 warps, two threads per score row
    asm volatile("setmaxnreg.inc.sync.aligned.u32 104;");
    const int sw = warp - 4;
    const int t = sw >> 3;           // query tile
    const int hsel = (sw >> 2) & 1;  // which 64 keys of the 128-key tile this thread owns
    const int qd = warp & 3;         // TMEM lane quarter (a warp may only touch lanes 32 * (warp % 4) ...)
    const int r = qd * 32 + lane;
    const uint32_t lane_off = (uint32_t)(qd * 32) << 16;
    const uint32_t tmem_s = tmem_base + (uint32_t)t * 128 + (uint32_t)hsel * 64 + lane_off;
    const uint32_t tmem_o = tmem_base + 256 + (uint32_t)t * 64 + lane_off;
    const uint32_t tmem_p = tmem_base + 384 + (uint32_t)t * 64 + (uint32_t)hsel * 32 + lane_off;
    const int q_idx = q0 + t * kQ3 + r;
    const float sc = p.scale_log2;
    uint64_t* const my_s_full = &s_full[t];
    uint64_t* const my_s_free = &s_free[t];
    uint64_t* const my_p_full = &p_full[t];
    uint64_t* const my_pv_done = &pv_done[t];
    const int pair_bar = 1 + t * 4 + qd;  // named barrier shared by the two warps that split these 32 rows
    const uint32_t hm_addr = smem_u32(half_max) + (uint32_t)t * 1024u + (uint32_t)r * 4u;
    float m_used = 0.f;  // exponent offset currently baked into O (scaled log2 units); identical in both halves of a row

    for (int j = 0; j < n_tiles; ++j) {
      mbar_wait(my_s_full, (uint32_t)j & 1u);
      tc_fence_after();
      uint32_t sv[64];
      tmem_ld32(tmem_s + 0, sv + 0);
      tmem_ld32(tmem_s + 32, sv + 32);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(my_s_free);

      const int limit = p.Nk - j * kK3 - hsel * 64;
      if (limit < 64) {
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
      // the other half of the row: write mine, 64-thread barrier, read the partner's (buffers alternate with the tile
      // parity, so a write for tile j + 2 cannot overtake the partner's read for tile j: the barrier of tile j + 1 sits between)
      const uint32_t hm = hm_addr + (uint32_t)(j & 1) * 2048u;  // shared-space address: [parity][tile][half][row]
      const float mine = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));
      asm volatile("st.shared.f32 [%0], %1;" ::"r"(hm + (uint32_t)hsel * 512u), "f"(mine) : "memory");
      asm volatile("bar.sync %0, 64;" ::"r"(pair_bar) : "memory");
      float other;
      asm volatile("ld.shared.f32 %0, [%1];" : "=f"(other) : "r"(hm + (uint32_t)(hsel ^ 1) * 512u) : "memory");
      const float mx = fmaxf(mine, other) * sc;
      if (j == 0) {
        m_used = mx;
      } else {
        // lazy rescale: only when this row's max outgrew the offset baked into O by more than 2^8. Both halves of a row see
        // the same mx and m_used, so both warps of the pair take this branch together and each rescales its share of O.
        const bool need = mx > m_used + kRescaleThreshold;
        if (__any_sync(0xffffffffu, need)) {
          mbar_wait(my_pv_done, (uint32_t)(j - 1) & 1u);  // every earlier P*V has landed in O
          tc_fence_after();
          const float m_new = need ? mx : m_used;
          const float f = ex2m(m_used - m_new);  // 1 for rows that do not need it
          m_used = m_new;
          // columns [0, 32) by the first half's warp, [32, 48) by the second's
          if (hsel == 0) {
            uint32_t v[32];
            tmem_ld32(tmem_o, v);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * f);
            tmem_st16(tmem_o, v);
            tmem_st16(tmem_o + 16, v + 16);
          } else {
            uint32_t v[16];
            tmem_ld16(tmem_o + 32, v);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * f);
            tmem_st16(tmem_o + 32, v);
          }
          tmem_st_wait();
        }
      }
      const float m_off = m_used;
      if (j >= 1) mbar_wait(my_pv_done, (uint32_t)(j - 1) & 1u);  // P*V of the previous tile has consumed the P buffer
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        uint32_t w[4];
        exp8_pack(sv + c * 8, sc, -m_off, kPolyMask == 0x10000u ? 2 : (int)((kPolyMask >> c) & 1u), w);
        asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};" ::"r"(tmem_p + (uint32_t)(c * 4)), "r"(w[0]),
                     "r"(w[1]), "r"(w[2]), "r"(w[3])
                     : "memory");
      }
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(my_p_full);
    }
    // epilogue: O[:, 0:40] / O[:, 40]; the first half's warp writes channels [0, 24), the second's [24, 40)
    if (n_tiles > 0) {
      mbar_wait(my_pv_done, (uint32_t)(n_tiles - 1) & 1u);
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
          if ((c < 24) == (hsel == 0)) {
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
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

}  // namespace a7
using namespace a7;

template <uint32_t kPolyMask>
static void launch_attn7_t(const AttnPlan& plan, cudaStream_t stream) {
  static bool attr_set = false;
  if (!attr_set) {
    LDN_CUDA(cudaFuncSetAttribute(attn7_tc_kernel<kPolyMask>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set = true;
  }
  attn7_tc_kernel<kPolyMask><<<plan.grid, kA7Threads, plan.smem_bytes, stream>>>(plan.p);
  LDN_CUDA(cudaGetLastError());
}

// kPolyMask: bit c set = chunk c (8 of a thread's 64 scores) takes the polynomial ex2 on the FMA pipes instead of the MUFU
void launch_attn7(const AttnPlan& plan, cudaStream_t stream) {
  switch (plan.p.poly_mod) {
    case 2: return launch_attn7_t<0xAAu>(plan, stream);  // 50 % polynomial
    case 3: return launch_attn7_t<0x49u>(plan, stream);  // 37.5 %
    case 4: return launch_attn7_t<0x88u>(plan, stream);  // 25 %
    case 8: return launch_attn7_t<0x80u>(plan, stream);  // 12.5 %
    case 99: return launch_attn7_t<0x10000u>(plan, stream);  // experiment: no exponential (wrong results)
    default: return launch_attn7_t<0u>(plan, stream);
  }
}

void finish_attn7_plan(AttnPlan& plan, int Nq, int Nk, int heads, int B) {
  AttnParams& p = plan.p;
  LDN_CHECK(p.d == 40 && p.dv == 48 && p.dqk == 48 && !p.causal, "attention7: d = 40, non-causal only");
  const int stage_bytes = 16384 + 2 * kDV3 * 128;
  const int fixed = 2 * 16384 + 1024 + 512 + 4096;  // + the half-row maximum exchange buffer
  const int n_tiles = (Nk + kK3 - 1) / kK3;
  int stages = (226 * 1024 - fixed) / stage_bytes;
  if (stages > 6) stages = 6;
  if (stages > n_tiles) stages = n_tiles;
  if (getenv("LDN_ATTN_STAGES")) stages = std::min(stages, atoi(getenv("LDN_ATTN_STAGES")));
  if (stages < 2 && n_tiles >= 2) stages = 2;
  if (stages < 1) stages = 1;
  p.kv_stages = stages;
  p.variant = 7;
  p.p_bufs = 1;
  p.pingpong = 0;
  plan.smem_bytes = fixed + stages * stage_bytes;
  plan.grid = dim3((Nq + 2 * kQ3 - 1) / (2 * kQ3), heads, B);
}

}  // namespace ldn
