// Attention kernel, fourth generation, head dim 40 — generation 3 (attention3.cu: ones-row V^T so the tensor core
// produces the softmax row sums, O resident in TMEM with lazy rescaling, two 128-row query tiles per CTA) with TWICE the
// softmax threads: two threads per query row, each owning 64 of the 128 score columns of a key tile.
//
// Why: the generation-3 ncu capture (profiles/) shows the kernel neither issue-bound (32 % issue slots), MUFU-bound (53 %)
// nor tensor-bound (20 %): with only two softmax warps per SM sub-partition the fixed-latency dependency stalls
// ("stall_wait") are not covered. Here every sub-partition holds four softmax warps of half the length:
//   warps  4.. 7 : tile 0, columns  0..63      warps  8..11 : tile 0, columns 64..127
//   warps 12..15 : tile 1, columns  0..63      warps 16..19 : tile 1, columns 64..127
// (a warp may only touch TMEM lanes 32*(warp%4).., so the two halves of a row live in warps w and w+4).
// The two half-rows exchange their partial maxima through shared memory + a 256-thread named barrier per tile; the
// P tile is written one 64-key swizzle atom per half; O rescaling (rare) splits the 48 O columns between the halves.
#include "common.h"
#include "ptx.cuh"

#include <algorithm>
#include <cstdlib>

namespace ldn {
namespace a4 {

static constexpr int kA4Threads = 640;  // 4 control warps + 16 softmax warps
static constexpr int kQ4 = 128;
static constexpr int kK4 = 128;
static constexpr int kDV4 = 48;
static constexpr float kRescale4 = 8.0f;

__device__ __forceinline__ float ex2m4(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint32_t pin4(uint32_t v) {
  asm volatile("mov.u32 %0, %0;" : "+r"(v));
  return v;
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t* v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(v[0]), "r"(v[1]),
               "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t* v) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
      "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t* v) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "r"(taddr)
               : "memory");
}

__global__ void __launch_bounds__(kA4Threads, 1) attn4_tc_kernel(const __grid_constant__ AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int h = blockIdx.y, b = blockIdx.z;
  const int q0 = blockIdx.x * (2 * kQ4);
  const int stages = p.kv_stages;
  constexpr uint32_t atom_bytes = 128 * 128;
  constexpr uint32_t vt_atom_bytes = kDV4 * 128;
  constexpr uint32_t stage_bytes = atom_bytes + 2 * vt_atom_bytes;

  uint8_t* q_smem = smem;                       // 2 query tiles
  uint8_t* p_smem = smem + 2 * atom_bytes;      // [tile][2 atoms]
  uint8_t* kv_smem = p_smem + 4 * atom_bytes;
  uint8_t* after = kv_smem + (size_t)stages * stage_bytes;
  float* mx_smem = reinterpret_cast<float*>(after);  // [2 tiles][2 halves][128 rows]
  uint64_t* bars = reinterpret_cast<uint64_t*>(after + 2 * 2 * 128 * sizeof(float));
  uint64_t* q_full = bars;         // 1
  uint64_t* s_full = bars + 1;     // [2]
  uint64_t* s_free = bars + 3;     // [2] 256 arrivals
  uint64_t* p_full = bars + 5;     // [2] 256 arrivals
  uint64_t* pv_done = bars + 7;    // [2]
  uint64_t* kv_full = bars + 9;
  uint64_t* kv_empty = kv_full + stages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(kv_empty + stages);
  constexpr uint32_t kTmemCols = 512;  // S0 [0,128) S1 [128,256) O0 [256,304) O1 [320,368)

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
    }
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
  const int n_tiles = (p.Nk + kK4 - 1) / kK4;

  if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
    if (warp == 0) {
      if (lane == 0) {
        mbar_arrive_expect_tx(q_full, 2 * atom_bytes);
        tma_load_2d(q_smem, &p.tmQ, q_full, h * p.slot, b * p.Nq + q0);
        tma_load_2d(q_smem + atom_bytes, &p.tmQ, q_full, h * p.slot, b * p.Nq + q0 + kQ4);
        for (int j = 0; j < n_tiles; ++j) {
          const int s = j % stages;
          const uint32_t ph = (uint32_t)(j / stages) & 1u;
          mbar_wait(&kv_empty[s], ph ^ 1u);
          uint8_t* k_dst = kv_smem + (size_t)s * stage_bytes;
          uint8_t* v_dst = k_dst + atom_bytes;
          mbar_arrive_expect_tx(&kv_full[s], stage_bytes);
          const int key0 = b * p.nk_pad + j * kK4;
          const int krow0 = b * p.k_batch_stride + j * kK4;
          tma_load_2d(k_dst, &p.tmK, &kv_full[s], h * p.slot, krow0);
          tma_load_2d(v_dst, &p.tmVt, &kv_full[s], key0, h * kDV4);
          tma_load_2d(v_dst + vt_atom_bytes, &p.tmVt, &kv_full[s], key0 + 64, h * kDV4);
        }
      }
    } else if (warp == 1) {
      if (lane == 0) {
        const uint32_t idesc_s = make_idesc_bf16(128, 128);
        const uint32_t idesc_pv = make_idesc_bf16(128, kDV4);
        const uint32_t q_addr = smem_u32(q_smem);
        const uint32_t p_addr = smem_u32(p_smem);
        const uint32_t kv_addr = smem_u32(kv_smem);
        auto issue_s = [&](int t, uint32_t k_addr) {
          const uint64_t a0 = make_smem_desc_sw128(q_addr + (uint32_t)t * atom_bytes);
          const uint64_t b0 = make_smem_desc_sw128(k_addr);
#pragma unroll
          for (int ks = 0; ks < 3; ++ks)
            tc_mma_bf16(tmem_base + (uint32_t)t * 128, a0 + (uint64_t)(2 * ks), b0 + (uint64_t)(2 * ks), idesc_s,
                        ks > 0 ? 1u : 0u);
          tc_commit(&s_full[t]);
        };
        mbar_wait(q_full, 0);
        mbar_wait(&kv_full[0], 0);
        tc_fence_after();
        issue_s(0, kv_addr);
        issue_s(1, kv_addr);
        for (int j = 0; j < n_tiles; ++j) {
          const int s = j % stages;
          const uint32_t v_addr = kv_addr + (uint32_t)s * stage_bytes + atom_bytes;
          if (j + 1 < n_tiles) {
            const int s1 = (j + 1) % stages;
            mbar_wait(&kv_full[s1], (uint32_t)((j + 1) / stages) & 1u);
            const uint32_t k_next = kv_addr + (uint32_t)s1 * stage_bytes;
            for (int t = 0; t < 2; ++t) {
              mbar_wait(&s_free[t], (uint32_t)j & 1u);
              tc_fence_after();
              issue_s(t, k_next);
            }
          }
          for (int t = 0; t < 2; ++t) {
            mbar_wait(&p_full[t], (uint32_t)j & 1u);
            tc_fence_after();
            const uint32_t pa = p_addr + (uint32_t)t * 2 * atom_bytes;
#pragma unroll
            for (int ks = 0; ks < kK4 / 16; ++ks) {
              const uint64_t adesc =
                  make_smem_desc_sw128(pa + (uint32_t)(ks >> 2) * atom_bytes) + (uint64_t)(2 * (ks & 3));
              const uint64_t bdesc =
                  make_smem_desc_sw128(v_addr + (uint32_t)(ks >> 2) * vt_atom_bytes) + (uint64_t)(2 * (ks & 3));
              tc_mma_bf16(tmem_base + 256 + (uint32_t)t * 64, adesc, bdesc, idesc_pv, (j > 0 || ks > 0) ? 1u : 0u);
            }
            tc_commit(&pv_done[t]);
          }
          tc_commit(&kv_empty[s]);
        }
      }
    }
  } else {
    // ---------------------------------------------------------------- softmax: 16 warps, two threads per row
    asm volatile("setmaxnreg.inc.sync.aligned.u32 104;");
    const int sw = warp - 4;
    const int t = sw >> 3;                 // query tile
    const int half = (sw >> 2) & 1;        // column half of the score tile
    const int qd = warp & 3;               // TMEM lane quarter
    const int r = qd * 32 + lane;
    const uint32_t lane_off = (uint32_t)(qd * 32) << 16;
    const uint32_t tmem_s = tmem_base + (uint32_t)t * 128 + (uint32_t)half * 64 + lane_off;
    const uint32_t tmem_o = tmem_base + 256 + (uint32_t)t * 64 + lane_off;
    const int q_idx = q0 + t * kQ4 + r;
    const float sc = p.scale_log2;
    // this half's swizzle atom of the P tile
    const uint32_t p_row = pin4(smem_u32(p_smem) + (uint32_t)(t * 2 + half) * atom_bytes + (uint32_t)r * 128);
    const uint32_t sw16 = (uint32_t)(r & 7) << 4;
    float* const mx_mine = mx_smem + (t * 2 + half) * 128 + r;
    const float* const mx_other = mx_smem + (t * 2 + (half ^ 1)) * 128 + r;
    uint64_t* const my_s_full = &s_full[t];
    uint64_t* const my_s_free = &s_free[t];
    uint64_t* const my_p_full = &p_full[t];
    uint64_t* const my_pv_done = &pv_done[t];
    const int bar_id = 1 + t;
    float m_used = 0.f;

    for (int j = 0; j < n_tiles; ++j) {
      mbar_wait(my_s_full, (uint32_t)j & 1u);
      tc_fence_after();
      uint32_t sv[64];
      tmem_ld32(tmem_s + 0, sv + 0);
      tmem_ld32(tmem_s + 32, sv + 32);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(my_s_free);

      const int limit = p.Nk - j * kK4 - half * 64;
      if (limit < 64) {
#pragma unroll
        for (int i = 0; i < 64; ++i)
          if (i >= limit) sv[i] = 0xff800000u;
      }
      float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
      for (int i = 0; i < 64; i += 8) {
        mx0 = fmaxf(mx0, fmaxf(__uint_as_float(sv[i]), __uint_as_float(sv[i + 1])));
        mx1 = fmaxf(mx1, fmaxf(__uint_as_float(sv[i + 2]), __uint_as_float(sv[i + 3])));
        mx2 = fmaxf(mx2, fmaxf(__uint_as_float(sv[i + 4]), __uint_as_float(sv[i + 5])));
        mx3 = fmaxf(mx3, fmaxf(__uint_as_float(sv[i + 6]), __uint_as_float(sv[i + 7])));
      }
      float mx = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));
      // exchange the half-row maxima (the barrier also orders the reuse of mx_smem across tiles: two syncs per tile)
      *mx_mine = mx;
      asm volatile("bar.sync %0, 256;" ::"r"(bar_id) : "memory");
      mx = fmaxf(mx, *mx_other) * sc;
      if (j == 0) {
        m_used = mx;
      } else {
        const bool need = mx > m_used + kRescale4;
        if (__any_sync(0xffffffffu, need)) {
          mbar_wait(my_pv_done, (uint32_t)(j - 1) & 1u);
          tc_fence_after();
          const float m_new = need ? mx : m_used;
          const float f = ex2m4(m_used - m_new);
          m_used = m_new;
          // this half rescales O columns [24*half, 24*half + 24)
          uint32_t v[24];
          const uint32_t o0 = tmem_o + (uint32_t)(half * 24);
          tmem_ld16(o0, v);
          tmem_ld8(o0 + 16, v + 16);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 24; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * f);
          tmem_st16(o0, v);
          tmem_st8(o0 + 16, v + 16);
          asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        }
      }
      // second sync: both halves have read the exchanged maxima (mx_smem may be rewritten next tile) ...
      asm volatile("bar.sync %0, 256;" ::"r"(bar_id) : "memory");
      // ... and the previous tile's P*V has finished reading the single P buffer
      if (j >= 1) mbar_wait(my_pv_done, (uint32_t)(j - 1) & 1u);
      const float m_off = m_used;
#pragma unroll
      for (int c = 0; c < 64; c += 8) {
        float e[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) e[i] = ex2m4(fmaf(__uint_as_float(sv[c + i]), sc, -m_off));
        const uint32_t addr = p_row + ((((uint32_t)c >> 3) << 4) ^ sw16);
        asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(pack_bf16x2(e[0], e[1])),
                     "r"(pack_bf16x2(e[2], e[3])), "r"(pack_bf16x2(e[4], e[5])), "r"(pack_bf16x2(e[6], e[7]))
                     : "memory");
      }
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(my_p_full);
    }
    // epilogue: half 0 writes the row (it needs all 41 columns: 40 values + the row sum in column 40)
    if (n_tiles > 0 && half == 0) {
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

}  // namespace a4
using namespace a4;

void launch_attn4(const AttnPlan& plan, cudaStream_t stream) {
  static bool attr_set = false;
  if (!attr_set) {
    LDN_CUDA(cudaFuncSetAttribute(attn4_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set = true;
  }
  attn4_tc_kernel<<<plan.grid, kA4Threads, plan.smem_bytes, stream>>>(plan.p);
  LDN_CUDA(cudaGetLastError());
}

void finish_attn4_plan(AttnPlan& plan, int Nq, int Nk, int heads, int B) {
  AttnParams& p = plan.p;
  LDN_CHECK(p.d == 40 && p.dv == 48 && p.dqk == 48 && !p.causal, "attention4: d = 40, non-causal only");
  const int stage_bytes = 16384 + 2 * kDV4 * 128;
  const int fixed = 2 * 16384 + 4 * 16384 + 2 * 2 * 128 * 4 + 1024 + 256;
  const int n_tiles = (Nk + kK4 - 1) / kK4;
  int stages = (226 * 1024 - fixed) / stage_bytes;
  if (stages > 5) stages = 5;
  if (stages > n_tiles) stages = n_tiles;
  if (stages < 2 && n_tiles >= 2) stages = 2;
  if (stages < 1) stages = 1;
  p.kv_stages = stages;
  p.variant = 4;
  plan.smem_bytes = fixed + stages * stage_bytes;
  plan.grid = dim3((Nq + 2 * kQ4 - 1) / (2 * kQ4), heads, B);
}

}  // namespace ldn
