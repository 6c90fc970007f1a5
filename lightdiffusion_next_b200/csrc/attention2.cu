// Attention kernel, second generation, for head dims <= 64 (SD1.5 level 0: d = 40, 42 % of the step's FLOPs; CLIP d = 64).
//
// Same contract and data layout as attention.cu. What changes is the schedule:
//   * one CTA owns TWO 128-row query tiles of one (batch, head) and shares every K / V^T tile between them
//     (half the TMA / shared-memory traffic per query row, 1 CTA per SM);
//   * each query tile has its own softmax warpgroup (1 thread = 1 row) and its own S / PV accumulators in TMEM
//     (S0 | S1 | PV0 | PV1);
//   * the whole 128-wide S row is read from TMEM ONCE into registers; as soon as it is there the warpgroup
//     releases the S accumulator (s_free) and the MMA warp issues S = Q K^T of the NEXT key tile, so that MMA runs
//     underneath the exponentials and is never on the critical path;
//   * O += PV of tile j-1 is folded in at the end of tile j (PV had a whole softmax phase to finish);
//   * registers are re-partitioned with setmaxnreg (control warpgroup 56, softmax warpgroups 224);
//   * a compile-time subset of the exponentials is evaluated with a degree-3 polynomial on the FMA pipes
//     (Cody-Waite range reduction + exponent splice): at d = 40 the kernel is bound by the 16 ex2/clk/SM MUFU rate
//     (1024 clk per 128x128 tile), not by the tensor core (384 clk).
// The per-tile softmax body is straight-line code kept under the 32 KB instruction cache (the first cut had both
// exp variants in every chunk and stalled on instruction fetch).
#include "common.h"
#include "ptx.cuh"

namespace ldn {

static constexpr int kA2Threads = 384;  // warpgroup 0: warp 0 TMA, warp 1 MMA (2, 3 idle); warpgroups 1 / 2: softmax of tile 0 / 1
static constexpr int kTQ = 128;
static constexpr int kTK = 128;

__device__ __forceinline__ float ex2_mufu(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// 2^x for x <= 0 on the FMA / ALU pipes: n = round(x), f = x - n in [-0.5, 0.5], 2^f by a degree-3 minimax polynomial
// (max rel. error 1.1e-4, far below the bf16 rounding of P), exponent spliced in with an integer add.
__device__ __forceinline__ float ex2_poly(float x) {
  x = fmaxf(x, -125.0f);
  const float t = x + 12582912.0f;  // 1.5 * 2^23: low mantissa bits now hold round(x)
  const float n = t - 12582912.0f;
  const float f = x - n;
  float p = fmaf(f, 0.0555041086f, 0.2402265070f);
  p = fmaf(p, f, 0.6931471806f);
  p = fmaf(p, f, 1.0f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));
}
// keeps a loop-invariant shared-memory address in a register (stops ptxas re-deriving it with S2UR/ULEA per use)
__device__ __forceinline__ uint32_t pin_u32(uint32_t v) {
  asm volatile("mov.u32 %0, %0;" : "+r"(v));
  return v;
}

// kPolyMask: bit c set -> the c-th group of 8 columns (of 16 per key tile) uses ex2_poly instead of the MUFU
template <int DV, uint32_t kPolyMask>
__global__ void __launch_bounds__(kA2Threads, 1) attn2_tc_kernel(const __grid_constant__ AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int h = blockIdx.y, b = blockIdx.z;
  const int q0 = blockIdx.x * (2 * kTQ);
  const int nqk_ksteps = p.dqk / 16;  // dqk <= 64: one swizzle atom
  const int stages = p.kv_stages;
  constexpr uint32_t atom_bytes = 128 * 128;
  constexpr uint32_t vt_atom_bytes = DV * 128;
  constexpr uint32_t stage_bytes = atom_bytes + 2 * vt_atom_bytes;

  uint8_t* q_smem = smem;                        // 2 tiles x 16 KB
  uint8_t* p_smem = smem + 2 * atom_bytes;       // 2 tiles x 2 buffers (key tile parity) x 2 atoms x 16 KB
  uint8_t* kv_smem = p_smem + 8 * atom_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(kv_smem + (size_t)stages * stage_bytes);
  uint64_t* q_full = bars;          // 1
  uint64_t* s_full = bars + 1;      // [2]
  uint64_t* s_free = bars + 3;      // [2], 128 arrivals: S row is in registers
  uint64_t* p_full = bars + 5;      // [2], 128 arrivals: P tile written, previous PV consumed
  uint64_t* pv_full = bars + 7;     // [2]
  uint64_t* kv_full = bars + 9;
  uint64_t* kv_empty = kv_full + stages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(kv_empty + stages);
  constexpr uint32_t kTmemCols = 512;  // S0 [0,128) S1 [128,256) PV0 [256,256+DV) PV1 [320,320+DV)

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmQ);
    tma_prefetch_desc(&p.tmK);
    tma_prefetch_desc(&p.tmVt);
    mbar_init(q_full, 1);
    for (int t = 0; t < 2; ++t) {
      mbar_init(&s_full[t], 1);
      mbar_init(&s_free[t], 128);
      mbar_init(&p_full[t], 128);
      mbar_init(&pv_full[t], 1);
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

  int nk_eff = p.Nk;
  if (p.causal) nk_eff = min(p.Nk, q0 + 2 * kTQ);
  const int n_tiles = (nk_eff + kTK - 1) / kTK;

  if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
    if (warp == 0) {
      if (lane == 0) {
        mbar_arrive_expect_tx(q_full, 2 * atom_bytes);
        tma_load_2d(q_smem, &p.tmQ, q_full, h * p.slot, b * p.Nq + q0);
        tma_load_2d(q_smem + atom_bytes, &p.tmQ, q_full, h * p.slot, b * p.Nq + q0 + kTQ);
        for (int j = 0; j < n_tiles; ++j) {
          const int s = j % stages;
          const uint32_t ph = (uint32_t)(j / stages) & 1u;
          mbar_wait(&kv_empty[s], ph ^ 1u);
          uint8_t* k_dst = kv_smem + (size_t)s * stage_bytes;
          uint8_t* v_dst = k_dst + atom_bytes;
          mbar_arrive_expect_tx(&kv_full[s], stage_bytes);
          const int key0 = b * p.nk_pad + j * kTK;
          const int krow0 = b * p.k_batch_stride + j * kTK;
          tma_load_2d(k_dst, &p.tmK, &kv_full[s], h * p.slot, krow0);
          tma_load_2d(v_dst, &p.tmVt, &kv_full[s], key0, h * p.d);
          tma_load_2d(v_dst + vt_atom_bytes, &p.tmVt, &kv_full[s], key0 + 64, h * p.d);
        }
      }
    } else if (warp == 1) {
      if (lane == 0) {
        const uint32_t idesc_s = make_idesc_bf16(128, 128);
        const uint32_t idesc_pv = make_idesc_bf16(128, DV);
        const uint32_t q_addr = smem_u32(q_smem);
        const uint32_t p_addr = smem_u32(p_smem);
        const uint32_t kv_addr = smem_u32(kv_smem);
        auto issue_s = [&](int t, uint32_t k_addr) {
          const uint64_t a0 = make_smem_desc_sw128(q_addr + (uint32_t)t * atom_bytes);
          const uint64_t b0 = make_smem_desc_sw128(k_addr);
          for (int ks = 0; ks < nqk_ksteps; ++ks)
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
          // 1) as soon as a warpgroup holds S^j in registers, run S^{j+1} underneath its exponentials
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
          // 2) P^j ready -> PV^j
          for (int t = 0; t < 2; ++t) {
            mbar_wait(&p_full[t], (uint32_t)j & 1u);
            tc_fence_after();
            const uint32_t pa = p_addr + (uint32_t)(t * 2 + (j & 1)) * 2 * atom_bytes;
#pragma unroll
            for (int ks = 0; ks < kTK / 16; ++ks) {
              const uint64_t adesc =
                  make_smem_desc_sw128(pa + (uint32_t)(ks >> 2) * atom_bytes) + (uint64_t)(2 * (ks & 3));
              const uint64_t bdesc =
                  make_smem_desc_sw128(v_addr + (uint32_t)(ks >> 2) * vt_atom_bytes) + (uint64_t)(2 * (ks & 3));
              tc_mma_bf16(tmem_base + 256 + (uint32_t)t * 64, adesc, bdesc, idesc_pv, ks > 0 ? 1u : 0u);
            }
            tc_commit(&pv_full[t]);
          }
          tc_commit(&kv_empty[s]);
        }
      }
    }
  } else {
    // ---------------------------------------------------------------- softmax warpgroups
    asm volatile("setmaxnreg.inc.sync.aligned.u32 224;");
    const int t = (warp - 4) >> 2;             // query tile 0 / 1
    const int qd = warp & 3;                   // TMEM lane quarter
    const int r = qd * 32 + lane;
    const uint32_t lane_off = (uint32_t)(qd * 32) << 16;
    const uint32_t tmem_s = tmem_base + (uint32_t)t * 128 + lane_off;
    const uint32_t tmem_pv = tmem_base + 256 + (uint32_t)t * 64 + lane_off;
    const int q_idx = q0 + t * kTQ + r;
    float m_run = -INFINITY, l_run = 0.f, alpha_prev = 0.f;
    constexpr int DO = (DV == 48) ? 40 : DV;   // columns actually kept (d = 40 pads to 48 only for the MMA shape)
    float o[DO];
#pragma unroll
    for (int i = 0; i < DO; ++i) o[i] = 0.f;
    const float sc = p.scale_log2;
    // P row base with the 128B swizzle pre-applied for 16-byte chunk 0; chunk c lives at (c ^ (r & 7)) * 16
    const uint32_t p_row0 = pin_u32(smem_u32(p_smem) + (uint32_t)t * 4 * atom_bytes + (uint32_t)r * 128);
    const uint32_t sw16 = (uint32_t)(r & 7) << 4;
    uint64_t* const my_s_full = &s_full[t];
    uint64_t* const my_s_free = &s_free[t];
    uint64_t* const my_p_full = &p_full[t];
    uint64_t* const my_pv_full = &pv_full[t];

    auto accumulate_o = [&](int j, float alpha) {
      mbar_wait(my_pv_full, (uint32_t)j & 1u);
      tc_fence_after();
      uint32_t v[16];
#pragma unroll
      for (int c = 0; c < DO; c += 16) {
        if (c + 16 <= DO) {
          tmem_ld16(tmem_pv + (uint32_t)c, v);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; ++i) o[c + i] = fmaf(o[c + i], alpha, __uint_as_float(v[i]));
        } else {
          asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                       : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                       : "r"(tmem_pv + (uint32_t)c)
                       : "memory");
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 8; ++i) o[c + i] = fmaf(o[c + i], alpha, __uint_as_float(v[i]));
        }
      }
      tc_fence_before();
    };

    if (t == 1 && n_tiles > 0) asm volatile("bar.arrive 1, 256;" ::: "memory");  // warpgroup 0 takes the first turn
    for (int j = 0; j < n_tiles; ++j) {
      mbar_wait(my_s_full, (uint32_t)j & 1u);
      tc_fence_after();
      uint32_t sv[128];
      tmem_ld32(tmem_s + 0, sv + 0);
      tmem_ld32(tmem_s + 32, sv + 32);
      tmem_ld32(tmem_s + 64, sv + 64);
      tmem_ld32(tmem_s + 96, sv + 96);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(my_s_free);  // S accumulator may be overwritten by the next key tile

      const int kbase = j * kTK;
      int limit = nk_eff - kbase;
      if (p.causal) limit = min(limit, q_idx - kbase + 1);
      if (limit < kTK) {
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
      const float mx = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));
      const float m_new = fmaxf(m_run, mx * sc);
      const float m_use = (m_new == -INFINITY) ? 0.f : m_new;
      const float alpha = ex2_mufu(m_run - m_use);
      // The two softmax warpgroups take turns on the MUFU (named barriers 1 / 2, 256 threads): while one runs its
      // exp phase at the full 16 ex2/clk/SM, the other loads S, reduces the row max, folds PV into O and waits
      // for the tensor core -- the same ping-pong FlashAttention-3 uses between GEMM and softmax, here between
      // the MUFU-bound and the non-MUFU halves of the softmax itself.
      asm volatile("bar.sync %0, 256;" ::"r"(1 + t) : "memory");
      // Software-pipelined exp phase: while the MUFU works on chunk c, the FMA pipe prepares chunk c+1 and the
      // ALU / LSU consume chunk c-1 (row sum, bf16 pack, swizzled store) -- three independent streams per iteration.
      const uint32_t p_row = p_row0 + (uint32_t)(j & 1) * 2 * atom_bytes;
      float rs0 = 0.f, rs1 = 0.f, rs2 = 0.f, rs3 = 0.f;
      float x[8], e[8], d[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) x[i] = fmaf(__uint_as_float(sv[i]), sc, -m_use);
#pragma unroll
      for (int c = 0; c <= 16; ++c) {
        if (c < 16) {
          if ((kPolyMask >> c) & 1u) {
#pragma unroll
            for (int i = 0; i < 8; ++i) e[i] = ex2_poly(x[i]);
          } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) e[i] = ex2_mufu(x[i]);
          }
        }
        if (c + 1 < 16) {
#pragma unroll
          for (int i = 0; i < 8; ++i) x[i] = fmaf(__uint_as_float(sv[(c + 1) * 8 + i]), sc, -m_use);
        }
        if (c > 0) {
          const int cc = (c - 1) * 8;
          rs0 += d[0] + d[4];
          rs1 += d[1] + d[5];
          rs2 += d[2] + d[6];
          rs3 += d[3] + d[7];
          const uint32_t addr = p_row + (uint32_t)(cc >> 6) * atom_bytes + ((((uint32_t)(cc & 63) >> 3) << 4) ^ sw16);
          asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(pack_bf16x2(d[0], d[1])),
                       "r"(pack_bf16x2(d[2], d[3])), "r"(pack_bf16x2(d[4], d[5])), "r"(pack_bf16x2(d[6], d[7]))
                       : "memory");
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) d[i] = e[i];
      }
      if (!(t == 1 && j == n_tiles - 1)) asm volatile("bar.arrive %0, 256;" ::"r"(1 + (t ^ 1)) : "memory");
      // fold in PV of the previous key tile (issued a whole exp phase ago) before PV^j may overwrite the accumulator
      if (j > 0) accumulate_o(j - 1, alpha_prev);
      alpha_prev = alpha;
      l_run = l_run * alpha + ((rs0 + rs1) + (rs2 + rs3));
      m_run = m_new;
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(my_p_full);
    }
    if (n_tiles > 0) accumulate_o(n_tiles - 1, alpha_prev);
    if (q_idx < p.Nq) {
      const float inv = l_run > 0.f ? 1.f / l_run : 0.f;
      bf16* orow = p.out + ((long long)b * p.Nq + q_idx) * p.ldo + (long long)h * p.d;
#pragma unroll
      for (int c = 0; c < DO; c += 8) {
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

template <int DV, uint32_t kPolyMask>
static void launch_attn2_t(const AttnPlan& plan, cudaStream_t stream) {
  static bool attr_set = false;
  if (!attr_set) {
    LDN_CUDA(cudaFuncSetAttribute(attn2_tc_kernel<DV, kPolyMask>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  227 * 1024));
    attr_set = true;
  }
  attn2_tc_kernel<DV, kPolyMask><<<plan.grid, kA2Threads, plan.smem_bytes, stream>>>(plan.p);
  LDN_CUDA(cudaGetLastError());
}

void launch_attn2(const AttnPlan& plan, cudaStream_t stream) {
  const int pm = plan.p.poly_mod;
  if (plan.p.dv == 48) {
    if (pm == 2) return launch_attn2_t<48, 0xAAAAu>(plan, stream);   // 50 % polynomial
    if (pm == 3) return launch_attn2_t<48, 0x9249u>(plan, stream);   // 37.5 %
    if (pm == 4) return launch_attn2_t<48, 0x8888u>(plan, stream);   // 25 %
    return launch_attn2_t<48, 0u>(plan, stream);
  }
  if (plan.p.dv == 64) return launch_attn2_t<64, 0u>(plan, stream);
  LDN_CHECK(false, "attention2: unsupported dv");
}

// fills the schedule-specific fields of a plan made by make_attn_plan (tensor maps are shared with version 1)
void finish_attn2_plan(AttnPlan& plan, int Nq, int Nk, int heads, int B) {
  AttnParams& p = plan.p;
  const int dv = p.dv;
  const int stage_bytes = 16384 + 2 * dv * 128;
  const int fixed = 2 * 16384 + 8 * 16384 + 1024 + 256;
  const int n_tiles = (Nk + kTK - 1) / kTK;
  int stages = (226 * 1024 - fixed) / stage_bytes;
  if (stages > 4) stages = 4;
  if (stages > n_tiles) stages = n_tiles;
  if (stages < 1) stages = 1;
  p.kv_stages = stages;
  p.variant = 2;
  plan.smem_bytes = fixed + stages * stage_bytes;
  plan.grid = dim3((Nq + 2 * kTQ - 1) / (2 * kTQ), heads, B);
}

}  // namespace ldn
