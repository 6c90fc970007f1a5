// tcgen05 GEMM and implicit-GEMM conv3x3 for sm_100a.
//
//   out[M, N] = A[M, K] * Wt[N, K]^T  (+ bias[N]) (+ rowbias[batch(row), N]) (+ residual[M, N])
//
// Replaces, on the reference's UNet path, the cuBLAS/cuDNN library calls made through
//   src/cond/cast.py:107 (F.linear) and :174 (Conv2d._conv_forward)            [reference file:line]
// for: ResBlock conv3x3 (src/AutoEncoders/ResBlock.py:251-292), 1x1 skip / proj_in / proj_out
// (ResBlock.py:294-299, src/NeuralNetwork/transformer.py:294-335), attention projections
// (src/Attention/Attention.py:85-98) and the GEGLU feed-forward (src/cond/Activation.py:6-31).
//
// Design (one 128 x BN output tile per CTA, 2 CTAs resident per SM so one tile's epilogue overlaps the
// other's main loop):
//   warp 0      TMA producer: A tile 128 rows x 64 K (bf16, 128B swizzle) + B tile BN x 64 K per stage.
//               conv3x3 mode: A is a 4-D NHWC tensor map; K chunk = (tap, 64 input channels); the tile is
//               a (BB, BH, BW) pixel box shifted by the tap offset; TMA out-of-bounds zero fill supplies
//               the conv padding -> no im2col buffer ever exists.
//               GEMM mode: A may be a virtual concat of two matrices along K (skip-connection concat).
//   warp 1      allocates TMEM; lane 0 issues tcgen05.mma (M=128, N=BN, K=16) x 4 per stage, fp32 accum in TMEM,
//               tcgen05.commit releases the smem stage / publishes the accumulator.
//   warps 2..9  epilogue (two warps per TMEM lane quarter, alternate 16-column chunks): tcgen05.ld 32 lanes x 16 columns, fused bias / time-embedding row bias / residual /
//               GEGLU (value * gelu_erf(gate)), bf16 (or fp32) store.
#include "common.h"
#include "ptx.cuh"
#include "gemm_epilogue.cuh"

#include <cstdlib>

namespace ldn {

// kMode: 0 general epilogue; 1 lean epilogue (bias / row bias / residual, bf16 out); 10 / 20 / 40: lean epilogue that also
// accumulates the GroupNorm statistics of the output (kMode = channels per group; gemm_epilogue.cuh).
template <int kMode>
__global__ void __launch_bounds__(kGemmThreads, 2) gemm_tc_kernel(const __grid_constant__ GemmParams p) {
  constexpr bool kLean = kMode != 0;
  constexpr bool kGn = kMode >= 2;
  extern __shared__ uint8_t smem_raw[];

  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int BN = p.BN;
  const uint32_t a_bytes = kBM * kBK * 2;
  const uint32_t b_bytes = (uint32_t)BN * kBK * 2;
  const uint32_t stage_bytes = a_bytes + b_bytes;
  const int stages = p.stages;

  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + (size_t)stages * stage_bytes);
  uint64_t* empty_bar = full_bar + stages;
  uint64_t* acc_bar = empty_bar + stages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_bar + 1);

  // tile coordinates
  const int n0 = blockIdx.x * BN;
  int m0 = 0, x0 = 0, y0 = 0, b0 = 0;
  if (p.conv) {
    int t = blockIdx.y;
    int tx = t % p.tiles_x;
    t /= p.tiles_x;
    int ty = t % p.tiles_y;
    int tb = t / p.tiles_y;
    x0 = tx * p.BW;
    y0 = ty * p.BH;
    b0 = tb * p.BB;
  } else {
    m0 = blockIdx.y * kBM;
  }

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmA0);
    tma_prefetch_desc(&p.tmB);
    if (!p.conv && p.a0_chunks < p.num_k_chunks) tma_prefetch_desc(&p.tmA1);
    for (int s = 0; s < stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(acc_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int kc_begin = blockIdx.z * p.chunks_per_split;
  const int kc_end = min(p.num_k_chunks, kc_begin + p.chunks_per_split);

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      for (int kc = kc_begin; kc < kc_end; ++kc) {
        const int it = kc - kc_begin;
        const int s = it % stages;
        const uint32_t ph = (uint32_t)(it / stages) & 1u;
        mbar_wait(&empty_bar[s], ph ^ 1u);
        uint8_t* a_dst = smem + (size_t)s * stage_bytes;
        uint8_t* b_dst = a_dst + a_bytes;
        mbar_arrive_expect_tx(&full_bar[s], stage_bytes);
        if (p.conv) {
          const int tap = kc / p.cin_chunks;
          const int cc = kc - tap * p.cin_chunks;
          const int dy = tap / 3 - 1;
          const int dx = tap - (tap / 3) * 3 - 1;
          tma_load_4d(a_dst, &p.tmA0, &full_bar[s], cc * kBK, x0 + dx, y0 + dy, b0);
        } else if (kc < p.a0_chunks) {
          tma_load_2d(a_dst, &p.tmA0, &full_bar[s], kc * kBK, m0);
        } else {
          tma_load_2d(a_dst, &p.tmA1, &full_bar[s], (kc - p.a0_chunks) * kBK, m0);
        }
        tma_load_2d(b_dst, &p.tmB, &full_bar[s], kc * kBK, n0);
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    // The whole warp walks the loop so that addresses / descriptors are warp-uniform (uniform registers); one elected
    // lane issues the MMAs and the commits.
    {
      const uint32_t idesc = make_idesc_bf16(kBM, (uint32_t)BN);
      const uint32_t smem_base = smem_u32(smem);
      int s = 0;
      uint32_t ph = 0;
      for (int kc = kc_begin; kc < kc_end; ++kc) {
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        const uint32_t a_addr = smem_base + (uint32_t)s * stage_bytes;
        const uint64_t a_desc = make_smem_desc_sw128(a_addr);
        const uint64_t b_desc = make_smem_desc_sw128(a_addr + a_bytes);
        if (elect_one()) {
          // advance 16 bf16 = 32 bytes along K inside the 128B swizzle atom: +2 in the (addr >> 4) field
          tc_mma_bf16(tmem_base, a_desc, b_desc, idesc, kc > kc_begin ? 1u : 0u);
          tc_mma_bf16(tmem_base, a_desc + 2, b_desc + 2, idesc, 1u);
          tc_mma_bf16(tmem_base, a_desc + 4, b_desc + 4, idesc, 1u);
          tc_mma_bf16(tmem_base, a_desc + 6, b_desc + 6, idesc, 1u);
          tc_commit(&empty_bar[s]);
          if (kc + 1 == kc_end) tc_commit(acc_bar);
        }
        __syncwarp();
        if (++s == stages) {
          s = 0;
          ph ^= 1u;
        }
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue (warps 2..5)
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    const int ehalf = (warp - 2) >> 2;  // the two warps of a quarter take alternate 16-column chunks
    const int r = q * 32 + lane;
    if constexpr (!kLean) {
      mbar_wait(acc_bar, 0);
      tc_fence_after();
    }

    long long out_row;  // output row index (pixel / token), -1 if masked
    int batch;
    int m_row = -1;
    if (p.conv) {
      const int bx = r % p.BW;
      const int by = (r / p.BW) % p.BH;
      const int bb = r / (p.BW * p.BH);
      const int x = x0 + bx, y = y0 + by, b = b0 + bb;
      const bool ok = (x < p.W) && (y < p.H) && (b < p.B);
      out_row = ok ? ((long long)(b * p.H + y) * p.W + x) : -1;
      batch = b;
    } else {
      const int m = m0 + r;
      m_row = m;
      out_row = (m < p.M) ? m : -1;
      if (p.row_head_dim > 0 && out_row >= 0) out_row = (m / p.row_head_dim) * p.row_head_slot + (m % p.row_head_dim);
      batch = p.rows_per_batch > 0 ? m / p.rows_per_batch : 0;
    }
    const uint32_t t_lane = tmem_base + ((uint32_t)(q * 32) << 16);
    if constexpr (kLean) {
      // lean epilogue: the residual of the first chunks is requested before the accumulator wait (the whole main loop)
      constexpr int kPf = 3;
      uint32_t wres[kPf][8];
      lean_prefetch_residual<kPf>(p, BN, n0, out_row, ehalf, wres);
      float* bsum = nullptr;
      if (p.bias_smem) {  // (uniform) behind the barriers: plan reserved 1 KB there
        bsum = reinterpret_cast<float*>(smem + (size_t)stages * stage_bytes + 256);
        lean_stage_bias(p, BN, n0, b0, bsum);
      }
      mbar_wait(acc_bar, 0);
      tc_fence_after();
      if constexpr (kGn) {
        // the accumulator is complete: every MMA has read its operands, the ring (>= 3 stages of 36 KB) is free scratch space
        gemm_epilogue_tile_lean_pf_gn<kPf, kMode>(p, n0, out_row, batch, t_lane, ehalf, wres, smem, bsum);
        asm volatile("bar.sync 1, 256;" ::: "memory");  // the eight epilogue warps
        gn_flush_tile<kMode>(p, n0, b0, smem);  // the tile's 128 pixels belong to image b0 (BB = 1)
      } else {
        gemm_epilogue_tile_lean_pf<kPf>(p, BN, n0, out_row, batch, t_lane, ehalf, wres, bsum);
      }
    } else {
      gemm_epilogue_tile<0>(p, BN, n0, out_row, batch, t_lane, ehalf, (int)blockIdx.z, m_row);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
  }
}

// ----------------------------------------------------------------------------------- persistent variant
// One CTA per SM walks a static list of (split, m-tile, n-tile) work items: barrier init, TMEM allocation and descriptor
// prefetch happen once per SM instead of once per tile; the TMA producer runs ahead into the next tile's K chunks while
// the epilogue warps drain the previous accumulator (two accumulator stages in TMEM: tmem_full / tmem_empty), and the
// shared-memory ring is as deep as the whole SM allows. Short-K GEMMs (1x1 projections, K = 320) were dominated by the
// per-CTA prologue + pipeline ramp of the one-tile-per-CTA kernel.
template <int kOcc, int kLn>
__global__ void __launch_bounds__(kGemmThreads, kOcc) gemm_tc_persist_kernel(const __grid_constant__ GemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int BN = p.BN;
  const uint32_t a_bytes = kBM * kBK * 2;
  const uint32_t b_bytes = (uint32_t)BN * kBK * 2;
  const uint32_t stage_bytes = a_bytes + b_bytes;
  const int stages = p.stages;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + (size_t)stages * stage_bytes);
  uint64_t* empty_bar = full_bar + stages;
  uint64_t* tfull_bar = empty_bar + stages;   // [2] accumulator ready
  uint64_t* tempty_bar = tfull_bar + 2;       // [2] accumulator drained (8 arrivals: one per epilogue warp)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
  // two accumulator stages (tile i + 1 accumulates while tile i drains), or one when two CTAs share the SM and a tile
  // needs more than 128 of the CTA's 256 columns (the other CTA then fills the tensor pipe during the drain)
  const bool two_acc = p.acc_stages == 2;
  const uint32_t acc_stride = two_acc ? (uint32_t)p.tmem_cols / 2 : 0u;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmA0);
    tma_prefetch_desc(&p.tmB);
    if (!p.conv && p.a0_chunks < p.num_k_chunks) tma_prefetch_desc(&p.tmA1);
    for (int s = 0; s < stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull_bar[a], 1);
      mbar_init(&tempty_bar[a], 8);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int n_tiles = p.grid_n, mn_tiles = p.grid_n * p.grid_m, total = mn_tiles * p.splits;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (int w = blockIdx.x; w < total; w += gridDim.x) {
        const int z = w / mn_tiles, rem = w - z * mn_tiles;
        const int mt = rem / n_tiles, nt = rem - mt * n_tiles;
        const int n0 = nt * BN;
        int m0 = 0, x0 = 0, y0 = 0, b0 = 0;
        if (p.conv) {
          int t = mt;
          const int tx = t % p.tiles_x;
          t /= p.tiles_x;
          const int ty = t % p.tiles_y;
          x0 = tx * p.BW;
          y0 = ty * p.BH;
          b0 = (t / p.tiles_y) * p.BB;
        } else {
          m0 = mt * kBM;
        }
        const int kc_begin = z * p.chunks_per_split;
        const int kc_end = min(p.num_k_chunks, kc_begin + p.chunks_per_split);
        for (int kc = kc_begin; kc < kc_end; ++kc) {
          mbar_wait(&empty_bar[s], ph ^ 1u);
          uint8_t* a_dst = smem + (size_t)s * stage_bytes;
          uint8_t* b_dst = a_dst + a_bytes;
          mbar_arrive_expect_tx(&full_bar[s], stage_bytes);
          if (p.conv) {
            const int tap = kc / p.cin_chunks;
            const int cc = kc - tap * p.cin_chunks;
            const int dy = tap / 3 - 1;
            const int dx = tap - (tap / 3) * 3 - 1;
            tma_load_4d(a_dst, &p.tmA0, &full_bar[s], cc * kBK, x0 + dx, y0 + dy, b0);
          } else if (kc < p.a0_chunks) {
            tma_load_2d(a_dst, &p.tmA0, &full_bar[s], kc * kBK, m0);
          } else {
            tma_load_2d(a_dst, &p.tmA1, &full_bar[s], (kc - p.a0_chunks) * kBK, m0);
          }
          tma_load_2d(b_dst, &p.tmB, &full_bar[s], kc * kBK, n0);
          if (++s == stages) {
            s = 0;
            ph ^= 1u;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer (warp-uniform loop, one elected lane issues)
    const uint32_t idesc = make_idesc_bf16(kBM, (uint32_t)BN);
    const uint32_t smem_base = smem_u32(smem);
    int s = 0;
    uint32_t ph = 0;
    int tl = 0;
    for (int w = blockIdx.x; w < total; w += gridDim.x, ++tl) {
      const int z = w / mn_tiles;
      const int kc_begin = z * p.chunks_per_split;
      const int kc_end = min(p.num_k_chunks, kc_begin + p.chunks_per_split);
      const int acc = two_acc ? (tl & 1) : 0;
      mbar_wait(&tempty_bar[acc], ((two_acc ? ((uint32_t)tl >> 1) : (uint32_t)tl) & 1u) ^ 1u);  // epilogue has drained this accumulator stage
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)acc * acc_stride;
      for (int kc = kc_begin; kc < kc_end; ++kc) {
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        const uint32_t a_addr = smem_base + (uint32_t)s * stage_bytes;
        const uint64_t a_desc = make_smem_desc_sw128(a_addr);
        const uint64_t b_desc = make_smem_desc_sw128(a_addr + a_bytes);
        if (elect_one()) {
          tc_mma_bf16(d_tmem, a_desc, b_desc, idesc, kc > kc_begin ? 1u : 0u);
          tc_mma_bf16(d_tmem, a_desc + 2, b_desc + 2, idesc, 1u);
          tc_mma_bf16(d_tmem, a_desc + 4, b_desc + 4, idesc, 1u);
          tc_mma_bf16(d_tmem, a_desc + 6, b_desc + 6, idesc, 1u);
          tc_commit(&empty_bar[s]);
          if (kc + 1 == kc_end) tc_commit(&tfull_bar[acc]);
        }
        __syncwarp();
        if (++s == stages) {
          s = 0;
          ph ^= 1u;
        }
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue (warps 2..9)
    const int q = warp & 3;
    const int ehalf = (warp - 2) >> 2;
    const int r = q * 32 + lane;
    int tl = 0;
    for (int w = blockIdx.x; w < total; w += gridDim.x, ++tl) {
      const int z = w / mn_tiles, rem = w - z * mn_tiles;
      const int mt = rem / n_tiles, nt = rem - mt * n_tiles;
      const int n0 = nt * BN;
      long long out_row;
      int batch;
      int m_row = -1;
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
        m_row = m;
        out_row = (m < p.M) ? m : -1;
        if (p.row_head_dim > 0 && out_row >= 0)
          out_row = (m / p.row_head_dim) * p.row_head_slot + (m % p.row_head_dim);
        batch = p.rows_per_batch > 0 ? m / p.rows_per_batch : 0;
      }
      const int acc = two_acc ? (tl & 1) : 0;
      const uint32_t acc_parity = (two_acc ? ((uint32_t)tl >> 1) : (uint32_t)tl) & 1u;
      const uint32_t t_lane = tmem_base + (uint32_t)acc * acc_stride + ((uint32_t)(q * 32) << 16);
      if constexpr (kLn == 5) {
        // lean epilogue with residual prefetch: the reads of this tile's residual are in flight while its main loop runs
        constexpr int kPf = kOcc == 1 ? 5 : 3;
        uint32_t wres[kPf][8];
        lean_prefetch_residual<kPf>(p, BN, n0, out_row, ehalf, wres);
        mbar_wait(&tfull_bar[acc], acc_parity);
        tc_fence_after();
        gemm_epilogue_tile_lean_pf<kPf>(p, BN, n0, out_row, batch, t_lane, ehalf, wres);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty_bar[acc]);
        continue;
      }
      mbar_wait(&tfull_bar[acc], acc_parity);
      tc_fence_after();
      gemm_epilogue_tile<kLn>(p, BN, n0, out_row, batch, t_lane, ehalf, z, m_row);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[acc]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
  }
}

// out[row, n] = sum_z ws[z][row][n] (fixed order => deterministic) + bias + rowbias + residual, 8 columns per thread
__global__ void splitk_reduce_kernel(const float* __restrict__ ws, long long split_stride, int splits, int rows, int N,
                                     const float* __restrict__ bias, const float* __restrict__ rowbias, int ld_rowbias,
                                     int rows_per_batch, const bf16* __restrict__ residual, long long ldr,
                                     bf16* __restrict__ out, long long ldo) {
  const int nv = N >> 3;
  const long long total = (long long)rows * nv;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int n = (int)(idx % nv) * 8;
    const long long row = idx / nv;
    float f[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) f[i] = 0.f;
    for (int z = 0; z < splits; ++z) {
      const float4* src = reinterpret_cast<const float4*>(ws + z * split_stride + row * N + n);
      const float4 a = src[0], b = src[1];
      f[0] += a.x; f[1] += a.y; f[2] += a.z; f[3] += a.w;
      f[4] += b.x; f[5] += b.y; f[6] += b.z; f[7] += b.w;
    }
    if (bias) {
#pragma unroll
      for (int i = 0; i < 8; ++i) f[i] += bias[n + i];
    }
    if (rowbias) {
      const float* rb = rowbias + (row / rows_per_batch) * ld_rowbias + n;
#pragma unroll
      for (int i = 0; i < 8; ++i) f[i] += rb[i];
    }
    if (residual) {
      const uint4 rv = *reinterpret_cast<const uint4*>(residual + row * ldr + n);
      const uint32_t w[4] = {rv.x, rv.y, rv.z, rv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        f[2 * i] += bf16_lo(w[i]);
        f[2 * i + 1] += bf16_hi(w[i]);
      }
    }
    uint4 ov;
    ov.x = pack_bf16x2(f[0], f[1]);
    ov.y = pack_bf16x2(f[2], f[3]);
    ov.z = pack_bf16x2(f[4], f[5]);
    ov.w = pack_bf16x2(f[6], f[7]);
    *reinterpret_cast<uint4*>(out + row * ldo + n) = ov;
  }
}

// ----------------------------------------------------------------------------------- host side

static int pick_bn(int N, bool geglu, long long m_tiles) {
  // Prefer tiles that divide N exactly (N in the UNet is 5*64*{1,2,4,8,...}); 160 wastes nothing for 320/640/1280.
  // The widest MMA (N = 256) is the most efficient one (measured: 1.38 PFLOP/s against ~1.1 at 160 and ~0.9 at 128): take it
  // whenever it divides N and still leaves at least two waves of tiles.
  if (N % 256 == 0 && (N / 256) * m_tiles >= 2 * 148) return 256;
  if (geglu) {
    const int cands[] = {160, 128, 256, 192, 96, 64, 32};
    for (int c : cands)
      if (N % c == 0) return c;
    return 128;
  }
  if (N <= 256 && N % 16 == 0) return N;
  const int cands[] = {160, 128, 192, 256, 224, 96, 64};
  for (int c : cands)
    if (N % c == 0) return c;
  return 128;
}

GemmPlan make_gemm_plan(const GemmArgs& a) {
  GemmPlan plan;
  GemmParams& p = plan.p;
  memset(&p, 0, sizeof(p));
  LDN_CHECK(a.Wt && a.A0 && (a.out || a.out_f32), "gemm: null operand");
  LDN_CHECK(a.N % 16 == 0, "gemm: N must be a multiple of 16");
  const long long m_rows = a.conv ? (long long)a.B * a.H * a.W : a.M;
  int BN = a.BN ? a.BN : pick_bn(a.N, a.epi == 1, (m_rows + kBM - 1) / kBM);
  if (a.epi == 1) LDN_CHECK(a.N % BN == 0 && BN % 32 == 0, "geglu: N must be a multiple of BN, BN of 32");
  LDN_CHECK(BN % 16 == 0 && BN >= 16 && BN <= 256, "gemm: bad BN");
  p.BN = BN;
  p.M = a.M;
  p.N = a.N;
  p.conv = a.conv ? 1 : 0;
  int K;
  if (a.conv) {
    LDN_CHECK(a.Cin % kBK == 0, "conv3x3: Cin must be a multiple of 64");
    K = 9 * a.Cin;
    p.cin_chunks = a.Cin / kBK;
    p.num_k_chunks = 9 * p.cin_chunks;
    p.a0_chunks = p.num_k_chunks;
    p.H = a.H;
    p.W = a.W;
    p.B = a.B;
    // pixel box of 128 rows
    int BW = 1;
    while (BW < a.W && BW < 128) BW <<= 1;
    int BH = 1;
    while (BW * BH < 128 && BH < a.H) BH <<= 1;
    int BB = 128 / (BW * BH);
    LDN_CHECK(BW * BH * BB == 128, "conv3x3: cannot form a 128-pixel box");
    p.BW = BW;
    p.BH = BH;
    p.BB = BB;
    p.tiles_x = (a.W + BW - 1) / BW;
    p.tiles_y = (a.H + BH - 1) / BH;
    const int tiles_b = (a.B + BB - 1) / BB;
    p.tmA0 = make_tmap_nhwc(a.A0, a.B, a.H, a.W, a.Cin, BB, BH, BW);
    p.tmA1 = p.tmA0;
    p.M = a.B * a.H * a.W;
    plan.grid = dim3((a.N + BN - 1) / BN, p.tiles_x * p.tiles_y * tiles_b, 1);
  } else {
    LDN_CHECK(a.K0 % 8 == 0 && a.K1 % 8 == 0, "gemm: K must be a multiple of 8");
    if (a.A1) LDN_CHECK(a.K0 % kBK == 0, "gemm: first K segment must be a multiple of 64");
    K = a.K0 + (a.A1 ? a.K1 : 0);
    p.a0_chunks = (a.K0 + kBK - 1) / kBK;
    p.num_k_chunks = p.a0_chunks + (a.A1 ? (a.K1 + kBK - 1) / kBK : 0);
    p.tmA0 = make_tmap_2d(a.A0, a.M, a.K0, a.lda0, kBM);
    p.tmA1 = a.A1 ? make_tmap_2d(a.A1, a.M, a.K1, a.lda1, kBM) : p.tmA0;
    plan.grid = dim3((a.N + BN - 1) / BN, (a.M + kBM - 1) / kBM, 1);
  }
  p.tmB = make_tmap_2d(a.Wt, a.wt_rows > 0 ? a.wt_rows : a.N, K, a.wt_ld > 0 ? a.wt_ld : K, BN);  // rows past wt_rows read as zero
  p.epi = a.epi;
  p.out = a.out;
  p.ldo = a.ldo;
  p.out_f32 = a.out_f32;
  p.bias = a.bias;
  p.rowbias = a.rowbias;
  p.colgate = a.colgate;
  p.ld_colgate = a.ld_colgate;
  p.ld_rowbias = a.ld_rowbias;
  p.rows_per_batch = a.rows_per_batch;
  p.residual = a.residual;
  p.ldr = a.ldr;
  p.head_dim = a.head_dim;
  p.head_slot = a.head_slot;
  p.act = a.act;
  p.rowstat_out = a.rowstat_out;
  p.rowstat_parts = 2 * (int)plan.grid.x;
  p.ln_parts = a.ln_parts;
  p.ln_nparts = a.ln_nparts;
  p.ln_inv_k = a.ln_width > 0 ? 1.0f / (float)a.ln_width : 0.f;
  p.ln_eps = a.ln_eps;
  p.ln_c = a.ln_c;
  p.ln_d = a.ln_d;
  p.ln_final_out = a.ln_final_out;
  p.ln_final_in = a.ln_final_in;
  if (a.ln_parts || a.ln_final_in) LDN_CHECK(a.ln_c && a.ln_d && !a.bias && !a.conv, "gemm: folded LayerNorm needs ln_c / ln_d and carries the bias in ln_d");
  if (a.rowstat_out) LDN_CHECK(!a.conv && a.epi == 0 && a.head_dim == 0 && a.row_head_dim == 0 && !a.out_f32, "gemm: row statistics are produced by plain epilogues only");
  static const int epi_opt = getenv("LDN_GEMM_EPI_OPT") ? atoi(getenv("LDN_GEMM_EPI_OPT")) : 7;
  p.epi_opt = epi_opt;
  {  // 256-bit epilogue accesses need 32-byte aligned rows
    const long long ldo_out = a.epi == 1 ? a.ldo : a.ldo;
    const bool ok = (reinterpret_cast<uintptr_t>(a.out) % 32 == 0) && (ldo_out % 16 == 0) &&
                    (!a.residual || ((reinterpret_cast<uintptr_t>(a.residual) % 32 == 0) && (a.ldr % 16 == 0))) &&
                    (a.epi != 1 || (a.N / 2) % 16 == 0);
    if (!ok) p.epi_opt &= ~1;
  }
  p.row_head_dim = a.row_head_dim;
  p.row_head_slot = a.row_head_slot;
  static const int lean_on = getenv("LDN_GEMM_LEAN") ? atoi(getenv("LDN_GEMM_LEAN")) : 1;
  plan.lean = lean_on && a.epi == 0 && a.act == 0 && !a.colgate && !a.out_f32 && a.head_dim == 0 && (p.epi_opt & 1) &&
              !a.rowstat_out && !a.ln_parts && !a.ln_final_in;  // (split-K is decided below and clears it)
  if (a.head_dim > 0) LDN_CHECK(a.head_dim % 8 == 0, "gemm: head_dim must be a multiple of 8");

  int cols = 32;
  while (cols < BN) cols <<= 1;
  p.tmem_cols = cols;
  const int stage_bytes = kBM * kBK * 2 + BN * kBK * 2;
  // 2 CTAs per SM: (227 KB - 2 KB reserve) / 2 per CTA, minus alignment slack and barriers.
  const int budget = 113 * 1024 - 1024 - 256;
  int stages = budget / stage_bytes;
  if (stages > 6) stages = 6;
  if (stages < 2) stages = 2;
  if (stages > p.num_k_chunks) stages = p.num_k_chunks < 1 ? 1 : p.num_k_chunks;
  // split-K when the tile grid cannot fill the machine and K is long (L2 / L3 / middle-block convs: M = 512..2048)
  p.splits = 1;
  p.chunks_per_split = p.num_k_chunks;
  p.total_rows = p.M;
  const int tiles = plan.grid.x * plan.grid.y;
  static const bool no_split = getenv("LDN_GEMM_NOSPLIT") != nullptr;  // experiments only
  static const int split_max_tiles = getenv("LDN_GEMM_SPLIT_MAX_TILES") ? atoi(getenv("LDN_GEMM_SPLIT_MAX_TILES")) : 74;
  if (!no_split && a.splitk_ws && !a.rowstat_out && !a.ln_parts && !a.ln_final_in && a.epi == 0 && a.head_dim == 0 && a.row_head_dim == 0 && a.act == 0 && !a.colgate && !a.out_f32 && tiles <= split_max_tiles && p.num_k_chunks >= 40) {
    int splits = (2 * 148 + tiles - 1) / tiles;
    if (splits > p.num_k_chunks / 8) splits = p.num_k_chunks / 8;
    if (splits > 16) splits = 16;
    while (splits > 1 && (size_t)splits * p.M * a.N * sizeof(float) > a.splitk_ws_bytes) --splits;
    if (splits > 1) {
      plan.lean = false;
      p.chunks_per_split = (p.num_k_chunks + splits - 1) / splits;
      splits = (p.num_k_chunks + p.chunks_per_split - 1) / p.chunks_per_split;
      p.splits = splits;
      p.ws = a.splitk_ws;
      p.ws_split_stride = (long long)p.M * a.N;
      plan.grid.z = splits;
      if (p.conv) p.rows_per_batch = a.H * a.W;
    }
  }
  // Experiment (off): a grid of at most one CTA per SM can never have two CTAs resident, so it could take the whole SM's
  // shared memory for a deeper ring.  The level-2 convs (128 tiles) did not get faster: their ~935 clk per 36 KB chunk is
  // the SM's L2 -> shared-memory ingest rate, not the TMA round trip.
  static const int deep_single = getenv("LDN_GEMM_DEEP_SINGLE") ? atoi(getenv("LDN_GEMM_DEEP_SINGLE")) : 0;  // measured: no gain (L2 convs 85.7 -> 96.1 us, profiles/r2_experiments.md section 11)
  if (deep_single && (long long)plan.grid.x * plan.grid.y * plan.grid.z <= 148) {
    stages = (225 * 1024 - 1024 - 256) / stage_bytes;
    if (stages > 8) stages = 8;
  }
  if (stages > p.chunks_per_split) stages = p.chunks_per_split;
  p.stages = stages;
  plan.smem_bytes = stages * stage_bytes + 1024 + 256;
  // persistent variant: one CTA per SM, two accumulator stages in TMEM, ring as deep as shared memory allows
  p.grid_n = plan.grid.x;
  p.grid_m = plan.grid.y;
  static const bool force_v1 = getenv("LDN_GEMM_V1") != nullptr;
  // persistent (1 CTA / SM, single MMA stream) wins when K is short (per-CTA prologue + pipeline ramp dominate); for
  // long-K problems (3x3 convs) two independent CTAs per SM keep the tensor pipe busier (measured 1018 vs 796 TFLOP/s)
  static const int persist_max_chunks = getenv("LDN_GEMM_PERSIST_MAX") ? atoi(getenv("LDN_GEMM_PERSIST_MAX")) : 20;
  plan.persistent = !force_v1 && p.chunks_per_split <= persist_max_chunks;
  // Two persistent CTAs per SM (LDN_GEMM_OCC2): sixteen epilogue warps per SM instead of eight.  The short-K GEMMs are bound
  // by their epilogue (TMEM drain + residual fetch + stores, all latency), not by the tensor pipe.
  // Measured (scripts/dev_gemm_graph.py, profiles/r2_experiments.md sections 9 and 17): 28.8 -> 24.9 us for the 320 x 320 projections
  // with a residual (M = 32768) -- more epilogue warps = more residual reads in flight.  The residual-prefetching lean epilogue
  // gets the same parallelism from one CTA per SM (23.3 us) without the single-accumulator penalty, so the mode is off by
  // default since.  0: never, 1: every persistent GEMM, 2: residual GEMMs with K <= 640 only.
  static const int occ2_mode = getenv("LDN_GEMM_OCC2") ? atoi(getenv("LDN_GEMM_OCC2")) : 0;
  const bool occ2 = occ2_mode == 1 || (occ2_mode == 2 && a.residual && a.epi == 0 && p.num_k_chunks <= 10 && m_rows >= 8192);
  p.acc_stages = 2;
  plan.persist_occ = 1;
  if (plan.persistent) {
    int acc_stride = 32;
    while (acc_stride < BN) acc_stride <<= 1;
    const int total = (int)(plan.grid.x * plan.grid.y * plan.grid.z);
    if (occ2 && total > 148) {
      plan.persist_occ = 2;
      p.acc_stages = 2 * acc_stride <= 256 ? 2 : 1;
      p.tmem_cols = p.acc_stages * acc_stride;  // <= 256
      int pst = (112 * 1024 - 2048) / stage_bytes;
      if (pst > 8) pst = 8;
      if (pst < 2) pst = 2;
      p.stages = pst;
      plan.smem_bytes = pst * stage_bytes + 1024 + 512;
      plan.pgrid = total < 296 ? total : 296;
    } else {
      p.tmem_cols = 2 * acc_stride;  // <= 512
      int pst = (225 * 1024 - 2048) / stage_bytes;
      if (pst > 8) pst = 8;
      p.stages = pst;
      plan.smem_bytes = pst * stage_bytes + 1024 + 512;
      plan.pgrid = total < 148 ? total : 148;
    }
  }
  // CTA-pair variant (gemm_pair.cu): 256 x BN tiles, each CTA loads half of the B tile
  // 0: never; 1: every eligible GEMM; 2: long-K problems (those that would otherwise run the one-tile-per-CTA kernel: the 3x3
  // convs) on the persistent pair kernel; 4 (default since round 2, profiles/r2_experiments.md section 24): long-K problems,
  // one 256 x BN tile per pair, two CTAs of different pairs per SM -- with the lean prefetching epilogue and the GroupNorm
  // statistics in it, 15.26 -> 14.94 ms per step on one box.
  static const int pair_mode = getenv("LDN_GEMM_PAIR") ? atoi(getenv("LDN_GEMM_PAIR")) : 4;
  const bool pair_ok = BN >= 32 && BN % 16 == 0 && plan.grid.y >= 2 && !force_v1;
  // (5, experiment: 4 plus the persistent pair kernel for the short-K problems)
  plan.pair = pair_ok && (pair_mode == 1 || (pair_mode == 2 && !plan.persistent) || (pair_mode == 3 && plan.persistent) ||
                          (pair_mode == 4 && !plan.persistent) || pair_mode == 5);
  if (plan.pair) {
    p.tmB2 = make_tmap_2d(a.Wt, a.wt_rows > 0 ? a.wt_rows : a.N, K, a.wt_ld > 0 ? a.wt_ld : K, BN / 2);
    int acc_stride = 32;
    while (acc_stride < BN) acc_stride <<= 1;
    const int stage2 = kBM * kBK * 2 + (BN / 2) * kBK * 2;
    const int m_pairs = ((int)plan.grid.y + 1) / 2;
    const int total_pairs = (int)plan.grid.x * m_pairs * (int)plan.grid.z;
    plan.pair_occ2 = pair_mode == 4 || (pair_mode == 5 && !plan.persistent);
    if (plan.pair_occ2) {
      p.tmem_cols = acc_stride;  // <= 256: two CTAs per SM share the 512 columns
      int pst = (113 * 1024 - 1024 - 512) / stage2;
      if (pst > p.chunks_per_split) pst = p.chunks_per_split;
      p.stages = pst;
      plan.pair_smem_bytes = pst * stage2 + 1024 + 512;
      plan.pgrid = 2 * total_pairs;
    } else {
      p.tmem_cols = 2 * acc_stride;
      int pst = (225 * 1024 - 2048) / stage2;
      if (pst > 10) pst = 10;
      p.stages = pst;
      plan.pair_smem_bytes = pst * stage2 + 1024 + 512;
      plan.pgrid = 2 * (total_pairs < 74 ? total_pairs : 74);
    }
  }
  // One-tile lean kernels: bias + row bias of the tile's columns staged in shared memory by the idle epilogue warps (needs one
  // row bias for the whole tile: conv tiles of one image, or no row bias) -- 1 KB more shared memory behind the barriers.
  static const int bias_smem_on = getenv("LDN_GEMM_BIAS_SMEM") ? atoi(getenv("LDN_GEMM_BIAS_SMEM")) : 1;
  if (bias_smem_on && plan.lean && !plan.persistent && (!plan.pair || plan.pair_occ2) && p.splits == 1 && (a.bias || a.rowbias) &&
      (a.rowbias == nullptr || (a.conv && p.BB == 1))) {
    p.bias_smem = 1;
    plan.smem_bytes += 1024;
    plan.pair_smem_bytes += 1024;
  }
  // GroupNorm statistics of the output in the epilogue: the lean one-tile-per-CTA kernel on BN = 160 tiles whose
  // 128 pixels belong to one image, N / 32 channels per group dividing the tile (10 / 20 / 40).  Anything else: the caller
  // runs the statistics kernel (GemmPlan::gn_cpg stays 0).
  const int gn_fuse = getenv("LDN_GN_FUSE") ? atoi(getenv("LDN_GN_FUSE")) : 1;  // read per plan (tests build both programs in one process)
  if (gn_fuse && a.gn_acc && a.conv && plan.lean && !plan.persistent && (!plan.pair || plan.pair_occ2) && p.splits == 1 &&
      BN == 160 && a.N % 160 == 0 && p.BB == 1 && (a.N == 320 || a.N == 640 || a.N == 1280) &&
      (size_t)p.stages * (plan.pair ? (kBM * kBK * 2 + (BN / 2) * kBK * 2) : stage_bytes) >= (size_t)kGnScratchBytes) {
    plan.gn_cpg = a.N / 32;
    p.gn_acc = a.gn_acc;
    p.gn_cpg = plan.gn_cpg;
  }
  return plan;
}

void launch_gemm(const GemmPlan& plan, cudaStream_t stream) {
  static bool attr_set = false;
  if (!attr_set) {
    LDN_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    LDN_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    LDN_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<10>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    LDN_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<20>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    LDN_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<40>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set = true;
  }
  if (plan.pair) {
    launch_gemm_pair(plan, stream);
  } else if (plan.persistent) {
    // folded-LayerNorm role of this GEMM (compile-time variants of the epilogue): 1 row consumer, 2 column consumer, 3 producer
    static const int lean_pf = getenv("LDN_GEMM_LEAN_PF") ? atoi(getenv("LDN_GEMM_LEAN_PF")) : 1;
    const int ln = plan.p.ln_parts ? 1 : plan.p.ln_final_in ? 2 : plan.p.rowstat_out ? 3 : plan.lean ? ((lean_pf && plan.p.residual) ? 5 : 4) : 0;
    static bool attr2 = false;
    if (!attr2) {
      auto prime = [](auto kern, int smem_max) { LDN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_max)); };
      prime(gemm_tc_persist_kernel<1, 0>, 227 * 1024); prime(gemm_tc_persist_kernel<1, 1>, 227 * 1024);
      prime(gemm_tc_persist_kernel<1, 2>, 227 * 1024); prime(gemm_tc_persist_kernel<1, 3>, 227 * 1024);
      prime(gemm_tc_persist_kernel<2, 0>, 113 * 1024); prime(gemm_tc_persist_kernel<2, 1>, 113 * 1024);
      prime(gemm_tc_persist_kernel<2, 2>, 113 * 1024); prime(gemm_tc_persist_kernel<2, 3>, 113 * 1024);
      prime(gemm_tc_persist_kernel<1, 4>, 227 * 1024); prime(gemm_tc_persist_kernel<2, 4>, 113 * 1024);
      prime(gemm_tc_persist_kernel<1, 5>, 227 * 1024); prime(gemm_tc_persist_kernel<2, 5>, 113 * 1024);
      attr2 = true;
    }
    auto go = [&](auto kern, int) { kern<<<plan.pgrid, kGemmThreads, plan.smem_bytes, stream>>>(plan.p); };
    if (plan.persist_occ == 2) {
      switch (ln) {
        case 1: go(gemm_tc_persist_kernel<2, 1>, 113 * 1024); break;
        case 2: go(gemm_tc_persist_kernel<2, 2>, 113 * 1024); break;
        case 3: go(gemm_tc_persist_kernel<2, 3>, 113 * 1024); break;
        case 4: go(gemm_tc_persist_kernel<2, 4>, 113 * 1024); break;
        case 5: go(gemm_tc_persist_kernel<2, 5>, 113 * 1024); break;
        default: go(gemm_tc_persist_kernel<2, 0>, 113 * 1024); break;
      }
    } else {
      switch (ln) {
        case 1: go(gemm_tc_persist_kernel<1, 1>, 227 * 1024); break;
        case 2: go(gemm_tc_persist_kernel<1, 2>, 227 * 1024); break;
        case 3: go(gemm_tc_persist_kernel<1, 3>, 227 * 1024); break;
        case 4: go(gemm_tc_persist_kernel<1, 4>, 227 * 1024); break;
        case 5: go(gemm_tc_persist_kernel<1, 5>, 227 * 1024); break;
        default: go(gemm_tc_persist_kernel<1, 0>, 227 * 1024); break;
      }
    }
  } else {
    LDN_CHECK(!plan.p.ln_parts && !plan.p.ln_final_in && !plan.p.rowstat_out, "gemm: folded LayerNorm runs on the persistent kernel only (K <= 1280)");
    if (plan.gn_cpg == 10)
      gemm_tc_kernel<10><<<plan.grid, kGemmThreads, plan.smem_bytes, stream>>>(plan.p);
    else if (plan.gn_cpg == 20)
      gemm_tc_kernel<20><<<plan.grid, kGemmThreads, plan.smem_bytes, stream>>>(plan.p);
    else if (plan.gn_cpg == 40)
      gemm_tc_kernel<40><<<plan.grid, kGemmThreads, plan.smem_bytes, stream>>>(plan.p);
    else if (plan.lean)
      gemm_tc_kernel<1><<<plan.grid, kGemmThreads, plan.smem_bytes, stream>>>(plan.p);
    else
      gemm_tc_kernel<0><<<plan.grid, kGemmThreads, plan.smem_bytes, stream>>>(plan.p);
  }
  LDN_CUDA(cudaGetLastError());
  if (plan.p.splits > 1) {
    const GemmParams& p = plan.p;
    const long long total = (long long)p.total_rows * (p.N / 8);
    int blocks = (int)((total + 255) / 256);
    if (blocks > 148 * 8) blocks = 148 * 8;
    splitk_reduce_kernel<<<blocks, 256, 0, stream>>>(p.ws, p.ws_split_stride, p.splits, p.total_rows, p.N, p.bias,
                                                    p.rowbias, p.ld_rowbias, p.rows_per_batch > 0 ? p.rows_per_batch : 1,
                                                    p.residual, p.ldr, p.out, p.ldo);
    LDN_CUDA(cudaGetLastError());
  }
}

}  // namespace ldn
