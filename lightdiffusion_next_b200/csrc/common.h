// Host-side shared declarations for the engine's kernels (internal; the public C ABI is include/ldn.h).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdexcept>
#include <string>

namespace ldn {

typedef __nv_bfloat16 bf16;

// ---- error handling: internal code throws, the C ABI catches and records ldn_last_error().
struct Error : public std::runtime_error {
  explicit Error(const std::string& s) : std::runtime_error(s) {}
};
void set_last_error(const std::string& s);
#define LDN_CHECK(cond, msg)                                                                       \
  do {                                                                                             \
    if (!(cond)) throw ::ldn::Error(std::string(__FILE__) + ":" + std::to_string(__LINE__) + ": " + (msg)); \
  } while (0)
#define LDN_CUDA(expr)                                                                   \
  do {                                                                                   \
    cudaError_t _e = (expr);                                                             \
    if (_e != cudaSuccess)                                                               \
      throw ::ldn::Error(std::string(__FILE__) + ":" + std::to_string(__LINE__) + ": " + \
                         cudaGetErrorString(_e) + " in " #expr);                         \
  } while (0)

// ---- TMA tensor maps (driver entry point resolved lazily; libcuda is not linked)
// 2-D row-major bf16 matrix [rows, cols] with leading dimension ld (elements); box = [box_rows, 64 cols], 128B swizzle.
CUtensorMap make_tmap_2d(const bf16* base, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows);
// 4-D NHWC bf16 activation [B, H, W, C]; box = [bb, bh, bw, 64 channels], 128B swizzle, OOB -> 0 (conv padding).
CUtensorMap make_tmap_nhwc(const bf16* base, int B, int H, int W, int C, int bb, int bh, int bw);

// ---- GEMM / implicit-GEMM conv3x3 on tcgen05 (gemm.cu)
struct GemmParams {
  CUtensorMap tmA0, tmA1, tmB;
  CUtensorMap tmB2;  // CTA-pair kernel: same matrix as tmB with a box of BN/2 rows (each CTA of the pair loads half the B tile)
  int M, N;          // logical output rows / weight rows
  int BN;            // N tile (multiple of 16, <= 256)
  int num_k_chunks;  // 64-wide K chunks in total
  int a0_chunks;     // GEMM mode: chunks taken from tmA0, the rest from tmA1 (virtual concat along K)
  int stages;
  int tmem_cols;
  // conv3x3 geometry (conv != 0): A is a 4-D NHWC map, K = 9 taps x Cin
  int conv;
  int H, W, B, cin_chunks, BW, BH, BB, tiles_x, tiles_y;
  // epilogue
  int epi;  // 0: out = acc (+bias) (+rowbias) (+residual);  1: GEGLU (value | gate halves of the tile)
  bf16* out;
  long long ldo;
  float* out_f32;  // if non-null, write fp32 here instead of bf16 `out`
  const float* bias;
  const float* rowbias;  // [batch, ld_rowbias] added per (batch(row), col)
  const float* colgate;  // [batch, ld_colgate] multiplies (acc + bias) per (batch(row), col) before the residual (Flux gates)
  int ld_colgate;
  int ld_rowbias;
  int rows_per_batch;
  const bf16* residual;
  long long ldr;
  int head_dim, head_slot;  // head_dim > 0: col n -> (n / head_dim) * head_slot + n % head_dim
  // LayerNorm folded into the GEMM that consumes it (transformer.py:154-157 norm1/2/3 -> q/k/v, q, GEGLU projections):
  //   LN(x) W^T = rstd * (x W'^T - mean * c) + d   with W' = W * gamma (bf16), c[n] = sum_k W'[n,k], d[n] = sum_k W[n,k] beta[k] (+ bias)
  // The PRODUCER of x (proj_in / attention out-projections) leaves per-row partial sums, the consumer finishes them.
  float2* rowstat_out;        // producer: [M, rowstat_parts] (sum, sum of squares) over the columns this (n-tile, epilogue half) stored
  int rowstat_parts;          // = 2 * n-tiles of the producer
  const float2* ln_parts;     // consumer, row mode: the producer's partials for the rows of A
  int ln_nparts;
  float ln_inv_k, ln_eps;     // 1 / normalised width, epsilon
  const float* ln_c;          // row mode: [N]; column mode: [M]
  const float* ln_d;
  float2* ln_final_out;       // row mode, optional: n-tile 0 / half 0 stores (rstd, -rstd * mean) per row for a later column-mode GEMM
  const float2* ln_final_in;  // column mode (operand-swapped V^T GEMM: output COLUMNS are tokens): per column (rstd, -rstd * mean)
  // GroupNorm statistics of the OUTPUT taken in this conv's epilogue (gemm_epilogue_tile_lean_pf_gn): the
  // 64-bit fixed-point (batch, group) accumulators of the consuming GroupNorm instance (norm.cu), or null
  unsigned long long* gn_acc;
  int gn_cpg;                 // output channels per group (10 / 20 / 40)
  int bias_smem;              // one-tile lean kernels: bias + row bias of the tile's columns staged in shared memory (gemm_epilogue.cuh)
  int acc_stages;    // persistent kernel: accumulator stages in TMEM (2, or 1 when two CTAs share the SM and BN > 128)
  int epi_opt;       // bit 0: 256-bit epilogue accesses, bit 1: packed-pair GEGLU arithmetic, bit 2: MUFU-free polynomial Phi in the packed GEGLU
  int act;  // epi 0 only: 0 none, 1 quick_gelu x*sigmoid(1.702x) applied after the bias (CLIP MLP, src/clip/Clip.py:74-77),
            // 2 ReLU after the bias, 3 ReLU after the residual add (TAESD, src/AutoEncoders/taesd.py:39-63),
            // 4 GELU (tanh approximation) after the bias (Flux MLPs, src/BlackForest/Flux.py:283-294)
  int row_head_dim, row_head_slot;  // GEMM mode, row_head_dim > 0: output row m -> (m / dim) * slot + m % dim
  // split-K (small-M problems): grid.z splits, each writes an fp32 partial tile; splitk_reduce_kernel finishes
  int splits, chunks_per_split;
  int grid_n, grid_m;  // tile grid (persistent kernel walks it)
  float* ws;
  long long ws_split_stride;  // elements between consecutive splits (= rows * N)
  int total_rows;
};

struct GemmPlan {
  GemmParams p;
  dim3 grid;
  int smem_bytes;
  bool persistent = false;
  int gn_cpg = 0;  // > 0: the lean one-tile kernel variant that also accumulates GroupNorm statistics of its output
  bool lean = false;  // plain bf16 epilogue (bias / rowbias / residual only): compile-time lean variant of the kernels
  int pgrid = 0;  // CTAs of the persistent kernel (<= SM count x persist_occ)
  int persist_occ = 1;  // persistent CTAs per SM (1 or 2)
  bool pair = false;  // CTA-pair kernel (gemm_pair.cu); pgrid is then an even CTA count
  int pair_smem_bytes = 0;
  bool pair_occ2 = false;  // one tile per CTA pair, two CTAs per SM (single accumulator stage)
};

struct GemmArgs {
  // operands
  const bf16* A0 = nullptr;  // [M, K0] (ld lda0), or NHWC activation for conv
  long long lda0 = 0;
  int K0 = 0;
  const bf16* A1 = nullptr;  // optional second K-segment [M, K1]
  long long lda1 = 0;
  int K1 = 0;
  const bf16* Wt = nullptr;  // [N, K] K-major (conv: K = tap*Cin + c)
  int M = 0, N = 0;
  // conv
  bool conv = false;
  int B = 0, H = 0, W = 0, Cin = 0;
  // epilogue
  int epi = 0;
  bf16* out = nullptr;
  long long ldo = 0;
  float* out_f32 = nullptr;
  const float* bias = nullptr;
  const float* rowbias = nullptr;
  const float* colgate = nullptr;  // per-(batch, column) gate applied before the residual add
  int ld_colgate = 0;
  int ld_rowbias = 0;
  int rows_per_batch = 0;
  const bf16* residual = nullptr;
  long long ldr = 0;
  int head_dim = 0, head_slot = 0;
  int row_head_dim = 0, row_head_slot = 0;
  int act = 0;
  // folded LayerNorm (see GemmParams)
  float2* rowstat_out = nullptr;
  const float2* ln_parts = nullptr;
  int ln_nparts = 0;
  int ln_width = 0;
  float ln_eps = 1e-5f;
  const float* ln_c = nullptr;
  const float* ln_d = nullptr;
  float2* ln_final_out = nullptr;
  const float2* ln_final_in = nullptr;
  // optional: accumulate the GroupNorm statistics (32 groups) of the output into these (batch, group) fixed-point
  // accumulators; make_gemm_plan reports in GemmPlan::gn_cpg whether the plan does it (else the caller runs gn_stats)
  unsigned long long* gn_acc = nullptr;
  int BN = 0;  // 0 = choose
  long long wt_ld = 0;  // leading dimension of Wt in elements (0 = K)
  int wt_rows = 0;  // valid rows of Wt if fewer than N (the rest are zero-filled by TMA)
  float* splitk_ws = nullptr;  // optional fp32 workspace enabling split-K for problems with too few tiles
  size_t splitk_ws_bytes = 0;
};

GemmPlan make_gemm_plan(const GemmArgs& a);
void launch_gemm(const GemmPlan& plan, cudaStream_t stream);
void launch_gemm_pair(const GemmPlan& plan, cudaStream_t stream);

// ---- attention (attention.cu)
struct AttnParams {
  CUtensorMap tmQ, tmK, tmVt;
  int heads;
  int Nq, Nk;        // tokens per (batch) for queries / keys
  int nk_pad;        // columns per batch in Vt (>= Nk)
  int k_batch_stride; // key rows per batch in the K buffer
  int d;             // head dim (output columns per head)
  int dqk;           // K extent of S = Q K^T, multiple of 16 (>= d)
  int dv;            // N extent of O = P V, multiple of 16 (>= d)
  int slot;          // per-head column slot width in the Q / K buffers (multiple of 64)
  int causal;
  int kv_stages;
  float scale_log2;  // softmax scale * log2(e)
  int vt_head_stride; // rows per head in Vt (d, or 48 when a ones row at index d supplies the softmax row sum)
  int variant;       // 1: attention.cu (one query tile per CTA); 2: attention2.cu (two query tiles, d <= 64); 5: attention5.cu (d = 40, S / O / P in TMEM); 6: attention6.cu (d = 80)
  int pingpong;      // variant 3: softmax warpgroups alternate on the MUFU
  int p_bufs;        // variant 3: P buffers per query tile in shared memory (1 or 2)
  int fold;          // variant 9: Q arrives pre-scaled by scale * log2(e) and column d of every K row holds 1.0: the kernel keeps -m in column d of its Q tile
  int poly_mod;      // variant 2: every poly_mod-th group of 8 exponentials runs on the FMA pipes (0 = all MUFU)
  bf16* out;         // [B*Nq, heads*d]
  long long ldo;
  const float* bias; // variant 1 only: additive logit bias * log2(e), [heads, bias_rows, bias_ld] (nullptr = none)
  int bias_rows, bias_ld;
};
struct AttnPlan {
  AttnParams p;
  dim3 grid;
  int smem_bytes;
};
struct AttnArgs {
  const bf16* Q;   // [B*Nq, heads*slot]
  long long ldq;
  const bf16* K;   // [B*nk_pad, heads*slot]
  long long ldk;
  const bf16* Vt;  // [heads*d (+pad), B*nk_pad], keys contiguous
  long long ldvt;
  long long vt_rows;
  int B, heads, Nq, Nk, nk_pad, d, slot;
  int kv_batch_stride = 0;  // K rows per batch (0 = nk_pad)
  int vt_head_stride = 0;   // rows per head in Vt (0 = d). d = 40 with stride 48 / d = 80 with stride 96: row d of every head must be all ones
  int causal = 0;
  int fold = 0;             // d = 40 only: Q is pre-scaled by scale * log2(e) and column 40 of every K slot holds 1.0 (see attention9.cu)
  float scale;
  bf16* out;
  long long ldo;
  const float* bias = nullptr;  // additive logit bias * log2(e), fp32 [heads, bias_rows, bias_ld], shared by all batches; rows / columns padded to multiples of 128 (d = 64, non-causal only)
  int bias_rows = 0, bias_ld = 0;
};
AttnPlan make_attn_plan(const AttnArgs& a);
void finish_attn2_plan(AttnPlan& plan, int Nq, int Nk, int heads, int B);
void launch_attn2(const AttnPlan& plan, cudaStream_t stream);
void finish_attn5_plan(AttnPlan& plan, int Nq, int Nk, int heads, int B);
void launch_attn5(const AttnPlan& plan, cudaStream_t stream);
void finish_attn6_plan(AttnPlan& plan, int Nq, int Nk, int heads, int B);
void launch_attn6(const AttnPlan& plan, cudaStream_t stream);
void finish_attn9_plan(AttnPlan& plan, const AttnArgs& a);
void launch_attn9(const AttnPlan& plan, cudaStream_t stream);
void launch_attn(const AttnPlan& plan, cudaStream_t stream);

// ---- normalisation / pointwise kernels (norm.cu, pointwise.cu)
// GroupNorm over NHWC bf16 input that may be a virtual concat of two tensors along C.
// stats_ws: workspace of groupnorm_ws_bytes(B) bytes: LDN_GN_SLOTS statistics slots ([B][32] x two 64-bit fixed-point totals).
// Every GroupNorm instance of a program uses its own slot; the whole workspace is zeroed once per program execution.
#define LDN_GN_SLOTS 160
// fixed-point scales of the accumulators: sum * 2^16, sum of squares * 2^12 (range: see norm.cu)
#define LDN_GN_SUM_SCALE 65536.0
#define LDN_GN_SQ_SCALE 4096.0
inline size_t groupnorm_ws_bytes(int B) { return (size_t)LDN_GN_SLOTS * B * 64 * sizeof(unsigned long long); }
void launch_groupnorm(const bf16* x0, int C0, const bf16* x1, int C1, int B, int HW, int groups, float eps,
                      const float* gamma, const float* beta, bool silu, bf16* out, float* stats_ws, int slot,
                      cudaStream_t stream, bool have_stats = false);  // have_stats: the slot already holds the totals (taken in the
                                                                      // producing conv's epilogue): only the apply kernel runs
inline unsigned long long* groupnorm_slot(float* stats_ws, int slot, int B) {
  return reinterpret_cast<unsigned long long*>(stats_ws) + (size_t)slot * B * 64;
}
void launch_layernorm(const bf16* x, int rows, int C, float eps, const float* gamma, const float* beta, bf16* out,
                      cudaStream_t stream, float* out_f32 = nullptr);
// out[b, n] = act_in(x[b, :]) . W[n, :] + bias[n]   (tiny-M linear; W bf16 [N, K], x fp32 [Bn, K])
void launch_small_linear(const float* x, int Bn, int K, const bf16* W, const float* bias, int N, bool silu_in,
                         bool silu_out, float* out, cudaStream_t stream, long long ldw = 0);  // ldw: weight row stride (0 = K)
// sigma[B] -> nearest discrete timestep index -> sinusoidal embedding [B, dim] fp32
void launch_timestep_embed(const float* sigma, int Bn, const float* log_sigmas, int n_sigmas, int dim, float* out,
                           float* t_index_out, cudaStream_t stream);
// x NCHW fp32 [B,C,H,W] * 1/sqrt(sigma^2+1) -> NHWC bf16 padded to Cpad channels (zeros)
void launch_scale_in(const float* x, const float* sigma, int B, int C, int H, int W, int Cpad, bf16* out,
                     cudaStream_t stream);
// conv_in: 3x3, Cin=4 (NCHW fp32 input scaled by 1/sqrt(sigma^2+1)), Cout=N -> NHWC bf16
void launch_conv_in(const float* x, const float* sigma, const bf16* Wt, const float* bias, int B, int H, int W,
                    int Cin, int Cout, bf16* out, cudaStream_t stream, int flags = 0, int ldo = 0);
// flags: 1 = tanh(x/3)*3 on the input (TAESD Clamp), 2 = ReLU on the output; ldo: output row pitch (0 = Cout) so that a wide
// layer can be produced in channel slices whose fp32 weights fit in shared memory
// conv_out: 3x3 Cin -> 4 on NHWC bf16 input; writes denoised = x - eps*sigma (NCHW fp32) and optionally eps
void launch_conv_out_finish(const float* acc16, const float* bias, const float* x, const float* sigma, int B, int HW,
                            int cout, float* denoised, cudaStream_t stream);
void launch_conv_out(const bf16* h, const bf16* Wt, const float* bias, const float* x, const float* sigma, int B,
                     int H, int W, int Cin, int Cout, float* denoised, float* eps_out, cudaStream_t stream);
void launch_upsample2x(const bf16* x, int B, int H, int W, int C, bf16* out, cudaStream_t stream);
// stride-2 3x3 pad-1 im2col gather: [B,H,W,C] -> [B*(H/2)*(W/2), 9*C]
void launch_im2col_s2(const bf16* x, int B, int H, int W, int C, bf16* out, cudaStream_t stream, int pad_before = 1);
void launch_vae_rgb_finish(const float* acc16, const float* bias, size_t npix, int cout, float* out, cudaStream_t stream,
                           int raw = 0);  // raw = 1: acc + bias without the [0,1] image mapping (TAESD)
// ---- Flux DiT helpers (flux_kernels.cu)
void launch_flux_temb(const float* t, int B, float* out, cudaStream_t stream);  // timestep_embedding_flux(t, 256)
void launch_vec_add3(const float* a, const float* b, const float* c, int n, float* out, cudaStream_t stream);
// out[r, :] = (1 + scale) * LayerNorm(x[r, :], eps 1e-6, no affine) + shift   (shift / scale: fp32 [C])
void launch_modln(const bf16* x, int rows, int C, const float* shift, const float* scale, bf16* out, cudaStream_t stream);
// in place on rows [0, rows) of a [*, ld] buffer holding `heads` q heads then `heads` k heads of 128: RMS-norm with the
// learned scales, then RoPE with pe[row, 64, (cos, sin)]
void launch_qk_norm_rope(bf16* qk, long long ld, int rows, int heads, const float* q_scale, const float* k_scale,
                         const float* pe, cudaStream_t stream);
void launch_vae_moments_finish(const float* acc16, const float* bc, const float* Wq, const float* bq, int B, int HW,
                               int zc2, float* out, cudaStream_t stream);
// dst[c, b*nk_pad + k] = src[c, b*N + k]: V^T re-laid with 16-byte aligned per-batch column offsets
void launch_pad_vt_cols(const bf16* src, int ld_src, int C, int Bn, int N, int nk_pad, bf16* dst, cudaStream_t stream);
void launch_fill_bf16(bf16* p, size_t n, float v, cudaStream_t stream);
// weight ingest: src is fp32 or fp16 (src_dtype 0 = f32, 1 = f16, 2 = bf16)
void launch_convert_to_bf16(const void* src, int src_dtype, size_t n, bf16* dst, cudaStream_t stream);
void launch_convert_to_f32(const void* src, int src_dtype, size_t n, float* dst, cudaStream_t stream);
// OIHW [O,I,kh,kw] -> [O, kh, kw, I] bf16
void launch_repack_conv_weight(const void* src, int src_dtype, int O, int I, int kh, int kw, bf16* dst,
                               cudaStream_t stream);
// CFG combine + sampler update (fp32, elementwise):  see pointwise.cu
void launch_cfg_step(const float* x, const float* den_uncond, const float* den_cond, float cfg, int mode, float c0,
                     float c1, float c2, const float* noise, float* x_out, float* denoised_out, size_t n,
                     cudaStream_t stream);
// F.interpolate(bilinear, align_corners=False) on fp32 [planes, h, w] -> [planes, oh, ow]
void launch_resample_bilinear(const float* src, float* dst, int planes, int h, int w, int oh, int ow, cudaStream_t stream);
// rgb[B,H,W,3] fp32 = clamp((conv3x3(h) + 1) / 2, 0, 1)   (VAE conv_out + VAE.process_output)
void launch_conv_out_rgb(const bf16* h, const bf16* Wt, const float* bias, int B, int H, int W, int Cin, float* rgb,
                         cudaStream_t stream);
// y[b,o,p] = sum_c W[o,c] x[b,c,p] + bias[o] on fp32 NCHW (VAE post_quant_conv, 4 -> 4 channels)
void launch_conv1x1_f32(const float* x, const float* W, const float* bias, int B, int Cin, int Cout, int HW, float* y,
                        cudaStream_t stream);
// out[r, :] = (ids[r] < vocab ? tok_emb[ids[r], :] : extra[ids[r] - vocab, :]) + pos_emb[r % T, :]  -> bf16
void launch_clip_embed(const long long* ids, const float* tok_emb, int vocab, const float* extra, const int* extra_n,
                       const float* pos_emb, int rows, int T, int C, bf16* out, cudaStream_t stream);
void launch_bislerp(const float* src, float* tmp, float* dst, int n, int c, int h, int w, int H, int W, cudaStream_t stream);
void launch_softmax_rows(const bf16* in, long long ld_in, bf16* out, long long ld_out, int rows, int cols, float scale,
                         cudaStream_t stream);
void launch_softmax_rows_f32(const float* in, long long ld_in, bf16* out, long long ld_out, int rows, int cols, float scale,
                             cudaStream_t stream);

}  // namespace ldn
