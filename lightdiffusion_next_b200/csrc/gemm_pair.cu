// tcgen05 GEMM / implicit-GEMM conv3x3 on CTA PAIRS (cluster of two CTAs on one TPC, `tcgen05.mma.cta_group::2`).
//
// Why: the one-CTA kernels (gemm.cu) top out at 1.0-1.3 PFLOP/s because a 128 x 160 x 64 k-chunk costs 36 KB of
// L2 -> shared-memory traffic for 2.6 MFLOP (71 FLOP/B) and an SM sustains only ~40-64 B/clk from L2
// (B300_MICROARCH: LTS cap ~6300 B/clk chip-wide) -- the tensor pipe waits for operands. A CTA pair computes a
// 256 x BN tile: each CTA loads its own 128 rows of A but only HALF of the B tile (BN/2 rows); the MMA issued by the leader
// reads both halves. 26 KB per 2.6 MFLOP at BN = 160 (100 FLOP/B), 32 KB per 4.2 MFLOP at BN = 256 (131 FLOP/B).
//
// Structure = the persistent kernel of gemm.cu (static tile list, two accumulator stages in TMEM, epilogue of tile i
// overlapping the main loop of tile i+1) with the pair protocol:
//   * both CTAs run a TMA producer; completion bytes of both are signalled on the LEADER's full barrier (count 2);
//   * only the leader's warp 1 issues MMAs; `tcgen05.commit ... multicast::cluster` releases the stage in both CTAs and
//     publishes the accumulator to both CTAs' epilogue warps;
//   * each CTA drains its own 128 TMEM lanes (rows); the peer's epilogue warps release the accumulator stage with a
//     remote arrive on the leader's barrier.
#include "common.h"
#include "ptx.cuh"
#include "gemm_epilogue.cuh"

#include <cstdlib>

namespace ldn {

// kOcc = 1: persistent (one pair per TPC walks a tile list, two accumulator stages); kOcc = 2: one tile per pair, two CTAs
// (of different pairs) per SM overlap each other's epilogue and main loop -- the arrangement that wins for long-K convs.
// kMode: 0 general epilogue; 1 lean epilogue with the residual prefetched behind the main loop; 10 / 20 / 40 (kOcc = 2 only):
// lean epilogue that also accumulates the GroupNorm statistics of the output (gemm.cu / gemm_epilogue.cuh).
template <int kOcc, int kMode>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kGemmThreads, kOcc)
    gemm_tc_pair_kernel(const __grid_constant__ GemmParams p) {
  constexpr bool kGn = kMode >= 2;
  static_assert(!kGn || kOcc == 2, "GroupNorm statistics: one tile per CTA pair");
  extern __shared__ uint8_t smem_raw[];

  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int BN = p.BN;
  const uint32_t a_bytes = kBM * kBK * 2;
  const uint32_t b_bytes = (uint32_t)(BN / 2) * kBK * 2;  // this CTA's half of the B tile
  const uint32_t stage_bytes = a_bytes + b_bytes;
  const int stages = p.stages;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + (size_t)stages * stage_bytes);  // used in the leader only
  uint64_t* empty_bar = full_bar + stages;
  uint64_t* tfull_bar = empty_bar + stages;   // [2] accumulator ready (both CTAs)
  uint64_t* tempty_bar = tfull_bar + 2;       // [2] accumulator drained (leader only: 16 arrivals, 8 per CTA)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
  const uint32_t acc_stride = kOcc == 1 ? (uint32_t)p.tmem_cols / 2 : 0u;  // kOcc = 2: a single accumulator stage

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmA0);
    tma_prefetch_desc(&p.tmB2);
    if (!p.conv && p.a0_chunks < p.num_k_chunks) tma_prefetch_desc(&p.tmA1);
    for (int s = 0; s < stages; ++s) {
      mbar_init(&full_bar[s], 2);   // one arrive.expect_tx per CTA of the pair
      mbar_init(&empty_bar[s], 1);  // one multicast commit from the leader's MMA warp
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull_bar[a], 1);
      mbar_init(&tempty_bar[a], 16);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc2(tmem_slot, (uint32_t)p.tmem_cols);
    tmem_relinquish2();
  }
  tc_fence_before();
  cluster_sync_all();  // barriers of both CTAs are initialised before anyone signals across the pair
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int n_tiles = p.grid_n;
  const int m_pairs = (p.grid_m + 1) / 2;
  const int mn_tiles = n_tiles * m_pairs, total = mn_tiles * p.splits;
  const int n_clusters = gridDim.x / 2, cluster_id = blockIdx.x / 2;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer (both CTAs)
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (int w = cluster_id; w < total; w += n_clusters) {
        const int z = w / mn_tiles, rem = w - z * mn_tiles;
        const int mp = rem / n_tiles, nt = rem - mp * n_tiles;
        const int mt = 2 * mp + (int)rank;
        const int n0 = nt * BN + (int)rank * (BN / 2);
        int m0 = 0, x0 = 0, y0 = 0, b0 = 0;
        if (p.conv) {
          int t = mt;
          const int tx = t % p.tiles_x;
          t /= p.tiles_x;
          const int ty = t % p.tiles_y;
          x0 = tx * p.BW;
          y0 = ty * p.BH;
          b0 = (t / p.tiles_y) * p.BB;  // past the last batch for the odd tail tile: TMA zero-fills
        } else {
          m0 = mt * kBM;
        }
        const int kc_begin = z * p.chunks_per_split;
        const int kc_end = min(p.num_k_chunks, kc_begin + p.chunks_per_split);
        for (int kc = kc_begin; kc < kc_end; ++kc) {
          mbar_wait(&empty_bar[s], ph ^ 1u);
          uint8_t* a_dst = smem + (size_t)s * stage_bytes;
          uint8_t* b_dst = a_dst + a_bytes;
          const uint32_t full_leader = mapa_u32(smem_u32(&full_bar[s]), 0);
          mbar_arrive_expect_tx_cluster(full_leader, stage_bytes);
          if (p.conv) {
            const int tap = kc / p.cin_chunks;
            const int cc = kc - tap * p.cin_chunks;
            const int dy = tap / 3 - 1;
            const int dx = tap - (tap / 3) * 3 - 1;
            tma2_load_4d(a_dst, &p.tmA0, full_leader, cc * kBK, x0 + dx, y0 + dy, b0);
          } else if (kc < p.a0_chunks) {
            tma2_load_2d(a_dst, &p.tmA0, full_leader, kc * kBK, m0);
          } else {
            tma2_load_2d(a_dst, &p.tmA1, full_leader, (kc - p.a0_chunks) * kBK, m0);
          }
          tma2_load_2d(b_dst, &p.tmB2, full_leader, kc * kBK, n0);
          if (++s == stages) {
            s = 0;
            ph ^= 1u;
          }
        }
        if constexpr (kOcc == 2) break;  // one tile per pair: tells the compiler the loop state dies here
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer (leader CTA only)
    if (rank == 0) {
      const uint32_t idesc = make_idesc_bf16(2 * kBM, (uint32_t)BN);
      const uint32_t smem_base = smem_u32(smem);
      int s = 0;
      uint32_t ph = 0;
      int tl = 0;
      for (int w = cluster_id; w < total; w += n_clusters, ++tl) {
        const int z = w / mn_tiles;
        const int kc_begin = z * p.chunks_per_split;
        const int kc_end = min(p.num_k_chunks, kc_begin + p.chunks_per_split);
        const int acc = tl & 1;
        mbar_wait(&tempty_bar[acc], (((uint32_t)tl >> 1) & 1u) ^ 1u);  // both CTAs have drained this accumulator stage
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)acc * acc_stride;
        for (int kc = kc_begin; kc < kc_end; ++kc) {
          mbar_wait(&full_bar[s], ph);
          tc_fence_after();
          const uint32_t a_addr = smem_base + (uint32_t)s * stage_bytes;
          const uint64_t a_desc = make_smem_desc_sw128(a_addr);
          const uint64_t b_desc = make_smem_desc_sw128(a_addr + a_bytes);
          if (elect_one()) {
            tc_mma2_bf16(d_tmem, a_desc, b_desc, idesc, kc > kc_begin ? 1u : 0u);
            tc_mma2_bf16(d_tmem, a_desc + 2, b_desc + 2, idesc, 1u);
            tc_mma2_bf16(d_tmem, a_desc + 4, b_desc + 4, idesc, 1u);
            tc_mma2_bf16(d_tmem, a_desc + 6, b_desc + 6, idesc, 1u);
            tc_commit2(&empty_bar[s], 3);
            if (kc + 1 == kc_end) tc_commit2(&tfull_bar[acc], 3);
          }
          __syncwarp();
          if (++s == stages) {
            s = 0;
            ph ^= 1u;
          }
        }
        if constexpr (kOcc == 2) break;
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue (warps 2..9, both CTAs: own 128 rows)
    const int q = warp & 3;
    const int ehalf = (warp - 2) >> 2;
    const int r = q * 32 + lane;
    int tl = 0;
    for (int w = cluster_id; w < total; w += n_clusters, ++tl) {
      const int z = w / mn_tiles, rem = w - z * mn_tiles;
      const int mp = rem / n_tiles, nt = rem - mp * n_tiles;
      const int mt = 2 * mp + (int)rank;
      const int n0 = nt * BN;
      long long out_row;
      int batch;
      if (p.conv) {
        int t = mt;
        const int tx = t % p.tiles_x;
        t /= p.tiles_x;
        const int ty = t % p.tiles_y;
        const int tb = t / p.tiles_y;
        const int bx = r % p.BW;
        const int by = (r / p.BW) % p.BH;
        const int bb = r / (p.BW * p.BH);
        const int x = tx * p.BW + bx, y = ty * p.BH + by, b = tb * p.BB + bb;
        const bool ok = (x < p.W) && (y < p.H) && (b < p.B);
        out_row = ok ? ((long long)(b * p.H + y) * p.W + x) : -1;
        batch = b;
      } else {
        const int m = mt * kBM + r;
        out_row = (m < p.M) ? m : -1;
        if (p.row_head_dim > 0 && out_row >= 0)
          out_row = (m / p.row_head_dim) * p.row_head_slot + (m % p.row_head_dim);
        batch = p.rows_per_batch > 0 ? m / p.rows_per_batch : 0;
      }
      const int acc = tl & 1;
      const uint32_t t_lane = tmem_base + (uint32_t)acc * acc_stride + ((uint32_t)(q * 32) << 16);
      if constexpr (kMode == 0) {
        mbar_wait(&tfull_bar[acc], ((uint32_t)tl >> 1) & 1u);
        tc_fence_after();
        gemm_epilogue_tile(p, BN, n0, out_row, batch, t_lane, ehalf, z);
      } else {
        constexpr int kPf = 3;
        uint32_t wres[kPf][8];
        lean_prefetch_residual<kPf>(p, BN, n0, out_row, ehalf, wres);  // in flight while the main loop runs
        float* bsum = nullptr;
        if (kOcc == 2 && p.bias_smem) {  // (uniform) behind the barriers: plan reserved 1 KB there
          bsum = reinterpret_cast<float*>(smem + (size_t)stages * stage_bytes + 512);
          lean_stage_bias(p, BN, n0, p.conv ? mt / (p.tiles_x * p.tiles_y) : 0, bsum);
        }
        mbar_wait(&tfull_bar[acc], ((uint32_t)tl >> 1) & 1u);
        tc_fence_after();
        if constexpr (kGn) {
          // the pair's accumulator is complete (multicast commit): every MMA has read both CTAs' operands, this CTA's ring
          // (>= 3 stages of 26 KB) is free scratch space
          gemm_epilogue_tile_lean_pf_gn<kPf, kMode>(p, n0, out_row, batch, t_lane, ehalf, wres, smem, bsum);
          asm volatile("bar.sync 1, 256;" ::: "memory");  // the eight epilogue warps
          // this CTA's 128 pixels belong to image tb (BB = 1; the odd tail tile of a pair lies past the last image and stored nothing)
          const int tb = mt / (p.tiles_x * p.tiles_y);
          if (tb < p.B) gn_flush_tile<kMode>(p, n0, tb, smem);
        } else {
          gemm_epilogue_tile_lean_pf<kPf>(p, BN, n0, out_row, batch, t_lane, ehalf, wres, bsum);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(mapa_u32(smem_u32(&tempty_bar[acc]), 0));
      if constexpr (kOcc == 2) break;
    }
  }

  tc_fence_before();
  cluster_sync_all();  // nobody leaves (or frees TMEM) while the peer may still signal or read
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc2(tmem_base, (uint32_t)p.tmem_cols);
  }
}

void launch_gemm_pair(const GemmPlan& plan, cudaStream_t stream) {
  static bool attr = false;
  if (!attr) {
    LDN_CUDA(cudaFuncSetAttribute(gemm_tc_pair_kernel<1, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    LDN_CUDA(cudaFuncSetAttribute(gemm_tc_pair_kernel<1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    LDN_CUDA(cudaFuncSetAttribute(gemm_tc_pair_kernel<2, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 114 * 1024));
    LDN_CUDA(cudaFuncSetAttribute(gemm_tc_pair_kernel<2, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 114 * 1024));
    LDN_CUDA(cudaFuncSetAttribute(gemm_tc_pair_kernel<2, 10>, cudaFuncAttributeMaxDynamicSharedMemorySize, 114 * 1024));
    LDN_CUDA(cudaFuncSetAttribute(gemm_tc_pair_kernel<2, 20>, cudaFuncAttributeMaxDynamicSharedMemorySize, 114 * 1024));
    LDN_CUDA(cudaFuncSetAttribute(gemm_tc_pair_kernel<2, 40>, cudaFuncAttributeMaxDynamicSharedMemorySize, 114 * 1024));
    attr = true;
  }
  auto go = [&](auto kern) { kern<<<plan.pgrid, kGemmThreads, plan.pair_smem_bytes, stream>>>(plan.p); };
  static const int pair_lean = getenv("LDN_GEMM_PAIR_LEAN") ? atoi(getenv("LDN_GEMM_PAIR_LEAN")) : 1;
  const bool lean = plan.lean && pair_lean;
  if (plan.pair_occ2) {
    if (plan.gn_cpg == 10) go(gemm_tc_pair_kernel<2, 10>);
    else if (plan.gn_cpg == 20) go(gemm_tc_pair_kernel<2, 20>);
    else if (plan.gn_cpg == 40) go(gemm_tc_pair_kernel<2, 40>);
    else if (lean) go(gemm_tc_pair_kernel<2, 1>);
    else go(gemm_tc_pair_kernel<2, 0>);
  } else {
    if (lean) go(gemm_tc_pair_kernel<1, 1>);
    else go(gemm_tc_pair_kernel<1, 0>);
  }
  LDN_CUDA(cudaGetLastError());
}

}  // namespace ldn
