// SD1.5 UNet forward as a fixed launch program over NHWC / token-major bf16 activations.
//
// Mirrors (structure, not code) UNetModel1.forward          src/NeuralNetwork/unet.py:679-770
//   ResBlock1._forward                                      src/AutoEncoders/ResBlock.py:315-335
//   SpatialTransformer.forward                              src/NeuralNetwork/transformer.py:342-377
//   BasicTransformerBlock._forward                          src/NeuralNetwork/transformer.py:186-245
//   CrossAttention.forward                                  src/Attention/Attention.py:100-124
//   FeedForward / GEGLU                                     transformer.py:19-70, src/cond/Activation.py:6-31
//   BaseModel.apply_model (EPS scaling in/out)              src/Model/ModelBase.py:72-133, src/sample/sampling.py:26-56
//
// B200-first choices: activations stay NHWC so a conv output *is* the token matrix (no rearrange copies);
// the skip concat is virtual (GroupNorm reads two sources, 1x1 skip conv reads two K segments); Q and K are one
// GEMM; V is produced already transposed by swapping GEMM operands; cross-attention K/V are computed once per
// context (ldn_set_context), not once per step; all 22 time-embedding projections are one launch; the whole
// forward is a CUDA graph.
#include <cmath>
#include <cstring>
#include <cstdlib>

#include "engine.h"

namespace ldn {

// ---------------------------------------------------------------- small device helpers local to the UNet
// dst row r = interleaved (value | gate) rows of the GEGLU projection in blocks of `half`
__global__ void geglu_interleave_kernel(const bf16* __restrict__ src, const float* __restrict__ bsrc, int inner,
                                        int K, int half, bf16* __restrict__ dst, float* __restrict__ bdst) {
  const int r = blockIdx.x;  // dst row in [0, 2*inner)
  const int bn = 2 * half;
  const int t = r / bn, j = r % bn;
  const int s = (j < half) ? (t * half + j) : (inner + t * half + (j - half));
  for (int k = threadIdx.x; k < K; k += blockDim.x) dst[(size_t)r * K + k] = src[(size_t)s * K + k];
  if (threadIdx.x == 0) bdst[r] = bsrc[s];
}

// ctx fp32 [rows, tokens, D] -> bf16 [rows, nk_pad, D] zero padded
__global__ void pad_context_kernel(const float* __restrict__ ctx, int rows, int tokens, int nk_pad, int D,
                                   bf16* __restrict__ out) {
  const size_t total = (size_t)rows * nk_pad * D;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int d = (int)(i % D);
    const size_t rt = i / D;
    const int t = (int)(rt % nk_pad);
    const int r = (int)(rt / nk_pad);
    out[i] = __float2bfloat16(t < tokens ? ctx[((size_t)r * tokens + t) * D + d] : 0.f);
  }
}

// dst[c, b*nk_pad + k] = src[c, b*N + k]  (only used when N % 8 != 0: TMA box starts must be 16-byte aligned)
__global__ void pad_vt_cols_kernel(const bf16* __restrict__ src, int ld_src, int C, int Bn, int N, int nk_pad,
                                   bf16* __restrict__ dst) {
  const size_t total = (size_t)C * Bn * N;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int k = (int)(i % N);
    const size_t cb = i / N;
    const int b = (int)(cb % Bn);
    const size_t c = cb / Bn;
    dst[(c * Bn + b) * nk_pad + k] = src[c * ld_src + (size_t)b * N + k];
  }
}

void launch_pad_vt_cols(const bf16* src, int ld_src, int C, int Bn, int N, int nk_pad, bf16* dst, cudaStream_t stream) {
  pad_vt_cols_kernel<<<64, 256, 0, stream>>>(src, ld_src, C, Bn, N, nk_pad, dst);
  LDN_CUDA(cudaGetLastError());
}

// V^T buffers of head-dim-40 layers carry 48 rows per head; row 40 of every head is all ones (softmax row sums on the
// tensor core, attention5.cu / attention6.cu), the remaining pad rows stay zero.
__global__ void fill_ones_rows_kernel(bf16* __restrict__ vt, int heads, long long ld, int head_stride, int row) {
  const long long total = (long long)heads * ld;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int h = (int)(i / ld);
    vt[((long long)h * head_stride + row) * ld + (i % ld)] = __float2bfloat16(1.0f);
  }
}

// Folded attention operands (attention9.cu): column d of every head slot in the K half of a [Q | K] buffer holds 1.0, so that
// column d of the kernel's Q tile (where it keeps -m of the row) is added to every score by the tensor core.
__global__ void fill_k_ones_kernel(bf16* __restrict__ qk, long long rows, long long ld, int heads, int slot, int k_off, int col) {
  const long long total = rows * heads;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long row = i / heads;
    const int h = (int)(i % heads);
    qk[row * ld + k_off + h * slot + col] = __float2bfloat16(1.0f);
  }
}

// Folded LayerNorm (gemm_epilogue.cuh): LN(x) W^T = rstd * (x W'^T - mean * c) + d.  One warp per weight row, in place:
//   d[n] = sum_k W[n,k] beta[k] (+ bias[n]);  W'[n,k] = bf16(W[n,k] * gamma[k]);  c[n] = sum_k W'[n,k] (the ROUNDED values the MMA will see).
__global__ void ln_fold_prepare_kernel(bf16* __restrict__ W, int rows, int K, const float* __restrict__ gamma,
                                       const float* __restrict__ beta, const float* __restrict__ bias,
                                       float* __restrict__ c, float* __restrict__ d) {
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  bf16* w = W + (size_t)row * K;
  float cs = 0.f, ds = 0.f;
  for (int k = lane; k < K; k += 32) {
    const float x = __bfloat162float(w[k]);
    ds = fmaf(x, beta[k], ds);
    const bf16 y = __float2bfloat16(x * gamma[k]);
    w[k] = y;
    cs += __bfloat162float(y);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    cs += __shfl_xor_sync(0xffffffffu, cs, o);
    ds += __shfl_xor_sync(0xffffffffu, ds, o);
  }
  if (lane == 0) {
    c[row] = cs;
    d[row] = ds + (bias ? bias[row] : 0.f);
  }
}

}  // namespace ldn

using namespace ldn;

namespace {

struct ResW {
  std::string prefix;
  int cin, cout;
  bool has_skip;
  int emb_off;  // column offset into the batched time-embedding projection
};
struct STW {
  std::string prefix;
  int C, d, slot;
  bf16* Wqk = nullptr;    // [2C, C]
  float* qk_gate = nullptr;  // d = 40: per-column multiplier of the [Q | K] projection, scale * log2(e) on the Q half (attention9.cu, folded operands)
  bf16* Wff1 = nullptr;   // interleaved [8C, C]
  float* bff1 = nullptr;  // interleaved [8C]
  // folded LayerNorms (norm1 -> [q|k] and v^T, norm2 -> cross-attention q, norm3 -> GEGLU): gamma-scaled weight copies + c / d vectors
  bool ln_fold = false;
  bf16 *Wqk_f = nullptr, *Wv_f = nullptr, *Wq2_f = nullptr, *Wff1_f = nullptr;
  float *cqk = nullptr, *dqk = nullptr, *cv = nullptr, *dv = nullptr, *cq2 = nullptr, *dq2 = nullptr, *cff = nullptr, *dff = nullptr;
  int ff_bn = 256;
  int index;  // 0..15, order of execution
};

enum BlockKind { B_CONV_IN, B_RES, B_ST, B_DOWN, B_UP };
struct BlockDesc {
  BlockKind kind;
  int res = -1, st = -1;
  std::string prefix;  // for down / up conv
  int ch = 0;
};

}  // namespace

struct ldn_engine::UNetState {
  // architecture (SD1.5: src/SD15/SD15.py:17-28, unet.py:941-1080)
  int model_ch = 320, heads = 8, ctx_dim = 768, temb_dim = 1280, in_ch = 4, out_ch = 4;
  std::vector<int> channel_mult = {1, 2, 4, 4};
  int num_res = 2;
  std::vector<bool> attn_level = {true, true, true, false};

  std::vector<ResW> res;
  std::vector<STW> sts;
  std::vector<std::vector<BlockDesc>> input_blocks, output_blocks;
  std::vector<BlockDesc> middle;
  std::vector<int> input_block_chans;

  bf16* Wemb_all = nullptr;  // [sum Cout, 1280]
  float* bemb_all = nullptr;
  int emb_total = 0;

  Arena arena;  // derived weights + context buffers
  // context
  int ctx_rows = 0, ctx_tokens = 0, nk_pad = 0;
  bf16* ctx_pad = nullptr;
  std::vector<bf16*> kctx, vtctx;  // per ST
  int ctx_cap_rows = 0, ctx_cap_pad = 0;

  std::map<std::tuple<int, int, int, int, int>, std::unique_ptr<Program>> programs;
  std::vector<std::unique_ptr<Arena>> program_arenas;
  int last_launches = 0;
};

namespace ldn {

static int slot_of(int d) { return (d + 63) / 64 * 64; }

void unet_finalize(ldn_engine* e, cudaStream_t stream) {
  LDN_CHECK(!e->w[0].empty(), "UNet weights not loaded");
  e->unet.reset(new ldn_engine::UNetState());
  auto& U = *e->unet;
  // ---- build the block table exactly as the LDM constructor does (unet.py:344-677)
  int ch = U.model_ch;
  U.input_blocks.push_back({BlockDesc{B_CONV_IN, -1, -1, "input_blocks.0.0", ch}});
  U.input_block_chans.push_back(ch);
  int emb_off = 0;
  auto add_res = [&](const std::string& prefix, int cin, int cout) {
    ResW r;
    r.prefix = prefix;
    r.cin = cin;
    r.cout = cout;
    r.has_skip = cin != cout;
    r.emb_off = emb_off;
    emb_off += cout;
    U.res.push_back(r);
    return (int)U.res.size() - 1;
  };
  auto add_st = [&](const std::string& prefix, int C) {
    STW s;
    s.prefix = prefix;
    s.C = C;
    s.d = C / U.heads;
    s.slot = slot_of(s.d);
    s.index = (int)U.sts.size();
    U.sts.push_back(s);
    return (int)U.sts.size() - 1;
  };
  int bi = 1;
  const int nlev = (int)U.channel_mult.size();
  for (int level = 0; level < nlev; ++level) {
    for (int nr = 0; nr < U.num_res; ++nr) {
      std::vector<BlockDesc> blk;
      const std::string p = "input_blocks." + std::to_string(bi);
      const int cout = U.channel_mult[level] * U.model_ch;
      BlockDesc r{B_RES};
      r.res = add_res(p + ".0", ch, cout);
      r.ch = cout;
      blk.push_back(r);
      ch = cout;
      if (U.attn_level[level]) {
        BlockDesc s{B_ST};
        s.st = add_st(p + ".1", ch);
        s.ch = ch;
        blk.push_back(s);
      }
      U.input_blocks.push_back(blk);
      U.input_block_chans.push_back(ch);
      ++bi;
    }
    if (level != nlev - 1) {
      BlockDesc d{B_DOWN};
      d.prefix = "input_blocks." + std::to_string(bi) + ".0.op";
      d.ch = ch;
      U.input_blocks.push_back({d});
      U.input_block_chans.push_back(ch);
      ++bi;
    }
  }
  {
    BlockDesc r0{B_RES};
    r0.res = add_res("middle_block.0", ch, ch);
    r0.ch = ch;
    BlockDesc s{B_ST};
    s.st = add_st("middle_block.1", ch);
    s.ch = ch;
    BlockDesc r1{B_RES};
    r1.res = add_res("middle_block.2", ch, ch);
    r1.ch = ch;
    U.middle = {r0, s, r1};
  }
  std::vector<int> chans = U.input_block_chans;
  int oi = 0;
  for (int level = nlev - 1; level >= 0; --level) {
    for (int i = 0; i <= U.num_res; ++i) {
      std::vector<BlockDesc> blk;
      const std::string p = "output_blocks." + std::to_string(oi);
      const int ich = chans.back();
      chans.pop_back();
      const int cout = U.model_ch * U.channel_mult[level];
      BlockDesc r{B_RES};
      r.res = add_res(p + ".0", ch + ich, cout);
      r.ch = cout;
      blk.push_back(r);
      ch = cout;
      int sub = 1;
      if (U.attn_level[level]) {
        BlockDesc s{B_ST};
        s.st = add_st(p + "." + std::to_string(sub), ch);
        s.ch = ch;
        blk.push_back(s);
        ++sub;
      }
      if (level > 0 && i == U.num_res) {
        BlockDesc u{B_UP};
        u.prefix = p + "." + std::to_string(sub) + ".conv";
        u.ch = ch;
        blk.push_back(u);
      }
      U.output_blocks.push_back(blk);
      ++oi;
    }
  }
  U.emb_total = emb_off;

  // ---- derived weights
  U.Wemb_all = U.arena.get<bf16>((size_t)U.emb_total * U.temb_dim);
  U.bemb_all = U.arena.get<float>(U.emb_total);
  for (auto& r : U.res) {
    const DevTensor& w = e->W(0, r.prefix + ".emb_layers.1.weight");
    const DevTensor& b = e->W(0, r.prefix + ".emb_layers.1.bias");
    LDN_CHECK(w.shape[0] == r.cout && w.shape[1] == U.temb_dim, "emb_layers shape mismatch at " + r.prefix);
    LDN_CUDA(cudaMemcpyAsync(U.Wemb_all + (size_t)r.emb_off * U.temb_dim, w.p, w.numel() * sizeof(bf16),
                             cudaMemcpyDeviceToDevice, stream));
    LDN_CUDA(cudaMemcpyAsync(U.bemb_all + r.emb_off, b.p, b.numel() * sizeof(float), cudaMemcpyDeviceToDevice,
                             stream));
    const DevTensor& c1 = e->W(0, r.prefix + ".in_layers.2.weight");
    LDN_CHECK(c1.shape[0] == r.cout && c1.shape[1] == 9 * r.cin, "conv1 shape mismatch at " + r.prefix);
  }
  for (auto& s : U.sts) {
    const std::string tb = s.prefix + ".transformer_blocks.0";
    const int C = s.C;
    const DevTensor& wq = e->W(0, tb + ".attn1.to_q.weight");
    const DevTensor& wk = e->W(0, tb + ".attn1.to_k.weight");
    LDN_CHECK(wq.shape[0] == C && wq.shape[1] == C, "to_q shape mismatch at " + s.prefix);
    s.Wqk = U.arena.get<bf16>((size_t)2 * C * C);
    LDN_CUDA(cudaMemcpyAsync(s.Wqk, wq.p, (size_t)C * C * sizeof(bf16), cudaMemcpyDeviceToDevice, stream));
    LDN_CUDA(cudaMemcpyAsync(s.Wqk + (size_t)C * C, wk.p, (size_t)C * C * sizeof(bf16), cudaMemcpyDeviceToDevice,
                             stream));
    if (s.d == 40) {
      std::vector<float> gate((size_t)2 * C, 1.0f);
      const float qs = (1.0f / sqrtf((float)s.d)) * 1.4426950408889634f;
      for (int i = 0; i < C; ++i) gate[i] = qs;
      s.qk_gate = U.arena.get<float>((size_t)2 * C);
      LDN_CUDA(cudaMemcpyAsync(s.qk_gate, gate.data(), gate.size() * sizeof(float), cudaMemcpyHostToDevice, stream));
      LDN_CUDA(cudaStreamSynchronize(stream));  // `gate` is a pageable temporary
    }
    const DevTensor& wf = e->W(0, tb + ".ff.net.0.proj.weight");
    const DevTensor& bf = e->W(0, tb + ".ff.net.0.proj.bias");
    LDN_CHECK(wf.shape[0] == 8 * C && wf.shape[1] == C, "GEGLU proj shape mismatch at " + s.prefix);
    s.ff_bn = 256;  // GEGLU tile: 128 value + 128 gate columns (N = 8C is a multiple of 256 for every SD1.5 level)
    LDN_CHECK((8 * C) % s.ff_bn == 0, "GEGLU width not a multiple of the tile");
    s.Wff1 = U.arena.get<bf16>((size_t)8 * C * C);
    s.bff1 = U.arena.get<float>((size_t)8 * C);
    geglu_interleave_kernel<<<8 * C, 128, 0, stream>>>(wf.b(), bf.f(), 4 * C, C, s.ff_bn / 2, s.Wff1, s.bff1);
    LDN_CUDA(cudaGetLastError());
    // SURVEY K5, built and measured (profiles/r2_experiments.md section 16): 48 launches fewer, same accuracy, but the K = C projections
    // are bound by their epilogues and the extra ~50 instructions per 16-column chunk cost more than the LayerNorm kernels did
    // (15.87 vs 15.81 ms per step).  Opt-in.
    static const int ln_fold_on = getenv("LDN_LN_FOLD") ? atoi(getenv("LDN_LN_FOLD")) : 0;
    if (ln_fold_on) {
      s.ln_fold = true;
      auto fold = [&](const bf16* src, int rows, const std::string& norm, const float* bias, bf16*& Wf, float*& cvec, float*& dvec) {
        Wf = U.arena.get<bf16>((size_t)rows * C);
        cvec = U.arena.get<float>((size_t)rows);
        dvec = U.arena.get<float>((size_t)rows);
        LDN_CUDA(cudaMemcpyAsync(Wf, src, (size_t)rows * C * sizeof(bf16), cudaMemcpyDeviceToDevice, stream));
        ln_fold_prepare_kernel<<<(rows * 32 + 255) / 256, 256, 0, stream>>>(Wf, rows, C, e->W(0, tb + norm + ".weight").f(),
                                                                           e->W(0, tb + norm + ".bias").f(), bias, cvec, dvec);
        LDN_CUDA(cudaGetLastError());
      };
      fold(s.Wqk, 2 * C, ".norm1", nullptr, s.Wqk_f, s.cqk, s.dqk);
      fold(e->W(0, tb + ".attn1.to_v.weight").b(), C, ".norm1", nullptr, s.Wv_f, s.cv, s.dv);
      fold(e->W(0, tb + ".attn2.to_q.weight").b(), C, ".norm2", nullptr, s.Wq2_f, s.cq2, s.dq2);
      fold(s.Wff1, 8 * C, ".norm3", s.bff1, s.Wff1_f, s.cff, s.dff);
    }
  }
  // context buffers (capacity from the config)
  const int cap_rows = e->cfg.max_rows > 0 ? e->cfg.max_rows : 2;
  const int cap_tok = e->cfg.max_ctx_tokens > 0 ? e->cfg.max_ctx_tokens : 77;
  U.ctx_cap_rows = cap_rows;
  U.ctx_cap_pad = (cap_tok + 127) / 128 * 128;
  U.ctx_pad = U.arena.get<bf16>((size_t)cap_rows * U.ctx_cap_pad * U.ctx_dim);
  for (auto& s : U.sts) {
    U.kctx.push_back(U.arena.get<bf16>((size_t)cap_rows * U.ctx_cap_pad * U.heads * s.slot, true));
    U.vtctx.push_back(U.arena.get<bf16>((size_t)(s.d == 40 ? U.heads * 48 : s.C) * cap_rows * U.ctx_cap_pad, true));
  }
  LDN_CUDA(cudaStreamSynchronize(stream));
  e->finalized[0] = true;
}

void unet_set_context(ldn_engine* e, const float* ctx, int rows, int tokens, cudaStream_t stream) {
  auto& U = *e->unet;
  const int nk_pad = (tokens + 127) / 128 * 128;
  LDN_CHECK(rows <= U.ctx_cap_rows && nk_pad <= U.ctx_cap_pad,
            "ldn_set_context: rows/tokens exceed the capacity given to ldn_create");
  if (rows != U.ctx_rows || tokens != U.ctx_tokens) {
    // layout (leading dimensions) changes: zero K buffers so slot padding columns / padded keys are 0
    for (size_t i = 0; i < U.sts.size(); ++i) {
      LDN_CUDA(cudaMemsetAsync(U.kctx[i], 0, (size_t)rows * nk_pad * U.heads * U.sts[i].slot * sizeof(bf16), stream));
      if (U.sts[i].d == 40) {
        const long long ld = (long long)rows * nk_pad;
        LDN_CUDA(cudaMemsetAsync(U.vtctx[i], 0, (size_t)U.heads * 48 * ld * sizeof(bf16), stream));
        fill_ones_rows_kernel<<<32, 256, 0, stream>>>(U.vtctx[i], U.heads, ld, 48, 40);
        LDN_CUDA(cudaGetLastError());
      }
    }
  }
  U.ctx_rows = rows;
  U.ctx_tokens = tokens;
  U.nk_pad = nk_pad;
  const size_t total = (size_t)rows * nk_pad * U.ctx_dim;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  pad_context_kernel<<<blocks, 256, 0, stream>>>(ctx, rows, tokens, nk_pad, U.ctx_dim, U.ctx_pad);
  LDN_CUDA(cudaGetLastError());
  const int M = rows * nk_pad;
  for (size_t i = 0; i < U.sts.size(); ++i) {
    const STW& s = U.sts[i];
    const std::string tb = s.prefix + ".transformer_blocks.0";
    // K_ctx = ctx * Wk^T, scattered into per-head slots
    GemmArgs a;
    a.A0 = U.ctx_pad; a.lda0 = U.ctx_dim; a.K0 = U.ctx_dim;
    a.Wt = e->W(0, tb + ".attn2.to_k.weight").b();
    a.M = M; a.N = s.C;
    a.out = U.kctx[i]; a.ldo = (long long)U.heads * s.slot;
    a.head_dim = s.d; a.head_slot = s.slot;
    launch_gemm(make_gemm_plan(a), stream);
    // V_ctx^T = Wv * ctx^T  -> [C, rows*nk_pad]
    GemmArgs v;
    v.A0 = e->W(0, tb + ".attn2.to_v.weight").b(); v.lda0 = U.ctx_dim; v.K0 = U.ctx_dim;
    v.Wt = U.ctx_pad;
    v.M = s.C; v.N = M;
    v.out = U.vtctx[i]; v.ldo = M;
    if (s.d == 40) { v.row_head_dim = 40; v.row_head_slot = 48; }
    launch_gemm(make_gemm_plan(v), stream);
  }
}

// ---------------------------------------------------------------- program construction
namespace {
struct Builder {
  ldn_engine* e;
  ldn_engine::UNetState& U;
  Program& P;
  Arena& A;
  int B, H0, W0;
  // scratch (sized for the largest layer)
  bf16 *sA = nullptr, *sB = nullptr, *sC = nullptr, *sO = nullptr, *sG = nullptr, *sVt = nullptr, *sCol = nullptr;
  bf16* sVtPad = nullptr;  // zero-initialised, only used for levels whose token count is not a multiple of 8
  std::vector<bf16*> sQK;  // per level
  std::vector<bf16*> sVt40;  // per level: [heads*48|96, Tld] V^T with a ones row per head (head dims 40 / 80)
  size_t vt_pad_elems = 0;
  float* splitk_ws = nullptr;
  size_t splitk_ws_bytes = 0;
  float* gn_ws = nullptr;
  int gn_slots = 0;
  float2* ln_parts = nullptr;  // [tokens, 16] row-statistics partials of the token stream (folded LayerNorm)
  float2* ln_final = nullptr;  // [tokens] (rstd, -rstd * mean) of norm1, for the operand-swapped V^T GEMM
  float *temb = nullptr, *emb1 = nullptr, *emb = nullptr, *emb_all = nullptr;

  void add(const std::string& name, Step s, int launches = 1) {
    P.steps.push_back(std::move(s));
    P.names.push_back(name);
    P.launches += launches;
  }
  // GroupNorm statistics taken in the producing conv's epilogue: `stats_of` is the tensor whose 32-group
  // statistics slot `stats_slot` already holds (or will hold when the program reaches the consumer); a GroupNorm over exactly
  // that tensor (no concat) then runs its apply kernel only.
  const bf16* stats_of = nullptr;
  int stats_slot = -1, stats_c = 0;
  bool last_gemm_took_stats = false;
  int new_gn_slot() {
    LDN_CHECK(gn_slots < LDN_GN_SLOTS, "too many GroupNorm instances for the statistics workspace");
    return gn_slots++;
  }
  // returns the number of row-statistics partials per row the plan writes (meaningful when a0.rowstat_out is set)
  int gemm(const std::string& name, const GemmArgs& a0) {
    GemmArgs a = a0;
    a.splitk_ws = splitk_ws;
    a.splitk_ws_bytes = splitk_ws_bytes;
    GemmPlan plan = make_gemm_plan(a);
    last_gemm_took_stats = plan.gn_cpg > 0;
    const long long Mm = a.conv ? (long long)a.B * a.H * a.W : a.M;
    const long long Kk = a.conv ? 9LL * a.Cin : (long long)a.K0 + a.K1;
    add(name + " [M=" + std::to_string(Mm) + " N=" + std::to_string(a.N) + " K=" + std::to_string(Kk) + "]",
        [plan](cudaStream_t st) { launch_gemm(plan, st); }, plan.p.splits > 1 ? 2 : 1);  // split-K adds its reduce kernel
    if (a.rowstat_out) LDN_CHECK(plan.p.rowstat_parts <= 16, "row-statistics buffer holds 16 partials per row");
    return plan.p.rowstat_parts;
  }
  void groupnorm(const std::string& name, const bf16* x0, int C0, const bf16* x1, int C1, int HW, float eps,
                 const std::string& wprefix, bool silu, bf16* out) {
    const float* g = e->W(0, wprefix + ".weight").f();
    const float* b = e->W(0, wprefix + ".bias").f();
    float* ws = gn_ws;
    int Bn = B;
    // every GroupNorm instance accumulates its statistics in its own (pre-zeroed) slot -- unless the conv that produced its
    // input already did (then only the apply kernel runs, on the producer's slot)
    const bool have = x1 == nullptr && x0 == stats_of && C0 == stats_c && stats_slot >= 0;
    const int slot = have ? stats_slot : new_gn_slot();
    add(name, [=](cudaStream_t st) { launch_groupnorm(x0, C0, x1, C1, Bn, HW, 32, eps, g, b, silu, out, ws, slot, st, have); }, have ? 1 : 2);  // (stats +) apply
  }
  void layernorm(const std::string& name, const bf16* x, int rows, int C, const std::string& wprefix, bf16* out) {
    const float* g = e->W(0, wprefix + ".weight").f();
    const float* b = e->W(0, wprefix + ".bias").f();
    add(name, [=](cudaStream_t st) { launch_layernorm(x, rows, C, 1e-5f, g, b, out, st); });
  }

  // ResBlock: x (C0) [+ skip (C1)] -> new buffer (Cout)
  bf16* resblock(const ResW& r, const bf16* x, int C0, const bf16* skip, int C1, int H, int W) {
    const int HW = H * W, M = B * HW;
    LDN_CHECK(C0 + C1 == r.cin, "resblock channel mismatch at " + r.prefix);
    bf16* out = A.get<bf16>((size_t)M * r.cout);
    groupnorm(r.prefix + ".gn1", x, C0, skip, C1, HW, 1e-5f, r.prefix + ".in_layers.0", true, sA);
    {
      GemmArgs a;
      a.conv = true; a.A0 = sA; a.B = B; a.H = H; a.W = W; a.Cin = r.cin;
      a.Wt = e->W(0, r.prefix + ".in_layers.2.weight").b(); a.N = r.cout;
      a.bias = e->W(0, r.prefix + ".in_layers.2.bias").f();
      a.rowbias = emb_all + r.emb_off; a.ld_rowbias = U.emb_total;
      a.out = sB; a.ldo = r.cout;
      const int slot = new_gn_slot();  // statistics of conv1's output for gn2, taken in conv1's epilogue where the plan allows
      a.gn_acc = groupnorm_slot(gn_ws, slot, B);
      gemm(r.prefix + ".conv1", a);
      stats_of = last_gemm_took_stats ? sB : nullptr;
      stats_slot = slot;
      stats_c = r.cout;
    }
    groupnorm(r.prefix + ".gn2", sB, r.cout, nullptr, 0, HW, 1e-5f, r.prefix + ".out_layers.0", true, sA);
    const bf16* res = x;
    if (r.has_skip) {
      GemmArgs a;
      a.A0 = x; a.lda0 = C0; a.K0 = C0;
      if (skip) { a.A1 = skip; a.lda1 = C1; a.K1 = C1; }
      a.Wt = e->W(0, r.prefix + ".skip_connection.weight").b();
      a.M = M; a.N = r.cout;
      a.bias = e->W(0, r.prefix + ".skip_connection.bias").f();
      a.out = sC; a.ldo = r.cout;
      gemm(r.prefix + ".skip", a);
      res = sC;
    } else {
      LDN_CHECK(skip == nullptr, "identity skip with concat input");
    }
    {
      GemmArgs a;
      a.conv = true; a.A0 = sA; a.B = B; a.H = H; a.W = W; a.Cin = r.cout;
      a.Wt = e->W(0, r.prefix + ".out_layers.3.weight").b(); a.N = r.cout;
      a.bias = e->W(0, r.prefix + ".out_layers.3.bias").f();
      a.residual = res; a.ldr = r.cout;
      a.out = out; a.ldo = r.cout;
      const int slot = new_gn_slot();  // statistics of the block's output for a GroupNorm that reads exactly it (transformer.norm,
      a.gn_acc = groupnorm_slot(gn_ws, slot, B);  // the next ResBlock's gn1 where nothing is concatenated)
      gemm(r.prefix + ".conv2", a);
      stats_of = last_gemm_took_stats ? out : nullptr;
      stats_slot = slot;
      stats_c = r.cout;
    }
    return out;
  }

  // SpatialTransformer (depth 1): x [B,H,W,C] -> new buffer
  bf16* transformer(const STW& s, const bf16* x, int H, int W, int level) {
    const int N = H * W, T = B * N, C = s.C;
    const int Tld = (T + 15) / 16 * 16;  // leading dimension of V^T (16-byte rows, whole 16-column GEMM chunks)
    const std::string tb = s.prefix + ".transformer_blocks.0";
    bf16* out = A.get<bf16>((size_t)T * C);
    bf16* X = sB;  // token stream
    bf16* QK = sQK[level];
    const long long ldqk = 2LL * U.heads * s.slot;
    const float scale = 1.0f / sqrtf((float)s.d);
    // LayerNorm folded into the consuming projections (norm1 / norm2 / norm3 have no kernel of their own): the producers of
    // the token stream X leave per-row partial sums in `ln_parts`, the consumers read the raw X.  Needs 16-token-aligned rows.
    const bool fold = s.ln_fold && (T % 16 == 0);
    int nparts = 0;
    groupnorm(s.prefix + ".norm", x, C, nullptr, 0, N, 1e-6f, s.prefix + ".norm", false, sA);
    {
      GemmArgs a;
      a.A0 = sA; a.lda0 = C; a.K0 = C; a.Wt = e->W(0, s.prefix + ".proj_in.weight").b(); a.M = T; a.N = C;
      a.bias = e->W(0, s.prefix + ".proj_in.bias").f(); a.out = X; a.ldo = C;
      if (fold) a.rowstat_out = ln_parts;  // statistics of X for norm1
      nparts = gemm(s.prefix + ".proj_in", a);
    }
    // ---- self-attention
    if (!fold) layernorm(tb + ".norm1", X, T, C, tb + ".norm1", sA);
    {
      GemmArgs a;
      a.A0 = fold ? X : sA; a.lda0 = C; a.K0 = C; a.Wt = fold ? s.Wqk_f : s.Wqk; a.M = T; a.N = 2 * C;
      if (fold) {
        a.ln_parts = ln_parts; a.ln_nparts = nparts; a.ln_width = C; a.ln_c = s.cqk; a.ln_d = s.dqk;
        a.ln_final_out = ln_final;  // (rstd, -rstd * mean) per token for the V^T GEMM below
      }
      a.out = QK; a.ldo = ldqk; a.head_dim = s.d; a.head_slot = s.slot;
      a.colgate = s.qk_gate; a.ld_colgate = 0;  // d = 40: Q leaves the projection already scaled by scale * log2(e)
      gemm(tb + ".attn1.qk", a);
      const bool ones = sVt40[level] != nullptr;  // V^T with a ones row per head: 48 (d = 40) / 96 (d = 80) rows
      const int hs = s.d == 40 ? 48 : 96;
      bf16* Vt = ones ? sVt40[level] : sVt;
      GemmArgs v;
      v.A0 = fold ? s.Wv_f : e->W(0, tb + ".attn1.to_v.weight").b(); v.lda0 = C; v.K0 = C; v.Wt = fold ? X : sA; v.M = C; v.N = Tld;
      v.wt_rows = T;
      if (fold) { v.ln_final_in = ln_final; v.ln_c = s.cv; v.ln_d = s.dv; }  // output columns are tokens here
      v.out = Vt; v.ldo = Tld;
      if (ones) { v.row_head_dim = s.d; v.row_head_slot = hs; }
      gemm(tb + ".attn1.vt", v);
      AttnArgs at;
      at.Q = QK; at.ldq = ldqk; at.K = QK + (size_t)U.heads * s.slot; at.ldk = ldqk;
      at.Vt = Vt; at.ldvt = Tld; at.vt_rows = ones ? U.heads * hs : C;
      at.vt_head_stride = ones ? hs : 0;
      at.B = B; at.heads = U.heads; at.Nq = N; at.Nk = N; at.nk_pad = N; at.d = s.d; at.slot = s.slot;
      at.fold = s.qk_gate ? 1 : 0;
      if (N % 8 != 0) {
        LDN_CHECK(!ones, "head-dim-40 level with a token count that is not a multiple of 8");
        // per-batch key offsets b*N would start TMA boxes at non-16B-aligned addresses: re-lay V^T with padded batches
        const int npad = (N + 7) / 8 * 8;
        LDN_CHECK((size_t)C * B * npad <= vt_pad_elems, "V^T pad scratch too small");
        const bf16* src = sVt;
        bf16* dst = sVtPad;
        const int Bn = B;
        add(tb + ".attn1.vt_pad", [=](cudaStream_t st) {
          pad_vt_cols_kernel<<<64, 256, 0, st>>>(src, Tld, C, Bn, N, npad, dst);
          LDN_CUDA(cudaGetLastError());
        });
        at.Vt = sVtPad; at.ldvt = (long long)B * npad;
        at.kv_batch_stride = N;  // K rows stay dense (row offsets have no alignment constraint)
        at.nk_pad = npad;
      }
      at.scale = scale; at.out = sO; at.ldo = C;
      AttnPlan plan = make_attn_plan(at);
      add(tb + ".attn1.sdpa", [plan](cudaStream_t st) { launch_attn(plan, st); });
      GemmArgs o;
      o.A0 = sO; o.lda0 = C; o.K0 = C; o.Wt = e->W(0, tb + ".attn1.to_out.0.weight").b(); o.M = T; o.N = C;
      o.bias = e->W(0, tb + ".attn1.to_out.0.bias").f(); o.residual = X; o.ldr = C; o.out = X; o.ldo = C;
      if (fold) o.rowstat_out = ln_parts;  // statistics of the updated X for norm2
      nparts = gemm(tb + ".attn1.out", o);
    }
    // ---- cross-attention (K/V precomputed by ldn_set_context)
    if (!fold) layernorm(tb + ".norm2", X, T, C, tb + ".norm2", sA);
    {
      GemmArgs a;
      a.A0 = fold ? X : sA; a.lda0 = C; a.K0 = C; a.Wt = fold ? s.Wq2_f : e->W(0, tb + ".attn2.to_q.weight").b(); a.M = T; a.N = C;
      if (fold) { a.ln_parts = ln_parts; a.ln_nparts = nparts; a.ln_width = C; a.ln_c = s.cq2; a.ln_d = s.dq2; }
      a.out = QK; a.ldo = ldqk; a.head_dim = s.d; a.head_slot = s.slot;
      gemm(tb + ".attn2.q", a);
      AttnArgs at;
      at.Q = QK; at.ldq = ldqk; at.K = U.kctx[s.index]; at.ldk = (long long)U.heads * s.slot;
      at.Vt = U.vtctx[s.index]; at.ldvt = (long long)U.ctx_rows * U.nk_pad; at.vt_rows = s.d == 40 ? U.heads * 48 : C;
      at.vt_head_stride = s.d == 40 ? 48 : 0;
      at.B = B; at.heads = U.heads; at.Nq = N; at.Nk = U.ctx_tokens; at.nk_pad = U.nk_pad; at.d = s.d;
      at.slot = s.slot; at.scale = scale; at.out = sO; at.ldo = C;
      AttnPlan plan = make_attn_plan(at);
      add(tb + ".attn2.sdpa", [plan](cudaStream_t st) { launch_attn(plan, st); });
      GemmArgs o;
      o.A0 = sO; o.lda0 = C; o.K0 = C; o.Wt = e->W(0, tb + ".attn2.to_out.0.weight").b(); o.M = T; o.N = C;
      o.bias = e->W(0, tb + ".attn2.to_out.0.bias").f(); o.residual = X; o.ldr = C; o.out = X; o.ldo = C;
      if (fold) o.rowstat_out = ln_parts;  // statistics of the updated X for norm3
      nparts = gemm(tb + ".attn2.out", o);
    }
    // ---- GEGLU feed-forward
    if (!fold) layernorm(tb + ".norm3", X, T, C, tb + ".norm3", sA);
    {
      GemmArgs a;
      a.A0 = fold ? X : sA; a.lda0 = C; a.K0 = C; a.Wt = fold ? s.Wff1_f : s.Wff1; a.M = T; a.N = 8 * C; a.epi = 1;
      if (fold) { a.ln_parts = ln_parts; a.ln_nparts = nparts; a.ln_width = C; a.ln_c = s.cff; a.ln_d = s.dff; }
      else a.bias = s.bff1;
      a.BN = s.ff_bn; a.out = sG; a.ldo = 4 * C;
      gemm(tb + ".ff.geglu", a);
      GemmArgs o;
      o.A0 = sG; o.lda0 = 4 * C; o.K0 = 4 * C; o.Wt = e->W(0, tb + ".ff.net.2.weight").b(); o.M = T; o.N = C;
      o.bias = e->W(0, tb + ".ff.net.2.bias").f(); o.residual = X; o.ldr = C; o.out = X; o.ldo = C;
      gemm(tb + ".ff.out", o);
    }
    {
      GemmArgs a;
      a.A0 = X; a.lda0 = C; a.K0 = C; a.Wt = e->W(0, s.prefix + ".proj_out.weight").b(); a.M = T; a.N = C;
      a.bias = e->W(0, s.prefix + ".proj_out.bias").f(); a.residual = x; a.ldr = C; a.out = out; a.ldo = C;
      gemm(s.prefix + ".proj_out", a);
    }
    return out;
  }
};
}  // namespace

static Program* build_unet_program(ldn_engine* e, int B, int H, int W) {
  auto& U = *e->unet;
  LDN_CHECK(H % 8 == 0 && W % 8 == 0, "latent height/width must be multiples of 8");
  LDN_CHECK(U.ctx_rows == B, "ldn_set_context rows must equal the denoise rows");
  std::unique_ptr<Program> prog(new Program());
  U.program_arenas.emplace_back(new Arena());
  Arena& A = *U.program_arenas.back();
  prog->arena = &A;
  Builder bd{e, U, *prog, A, B, H, W};
  const int nlev = (int)U.channel_mult.size();
  // ---- scratch sizing
  size_t max_act = 0, max_geglu = 0, max_col = 0;
  {
    int h = H, w = W;
    for (int level = 0; level < nlev; ++level) {
      const int c = U.model_ch * U.channel_mult[level];
      const size_t M = (size_t)B * h * w;
      // widest tensors at this resolution: concat inputs on the way up (prev level channels + this level's)
      const int cprev = U.model_ch * U.channel_mult[std::min(level + 1, nlev - 1)];
      max_act = std::max(max_act, M * (size_t)(c + cprev));
      if (U.attn_level[level]) max_geglu = std::max(max_geglu, M * (size_t)4 * c);
      if (level != nlev - 1) max_col = std::max(max_col, (M / 4) * (size_t)9 * c);
      if (level != nlev - 1) {
        h /= 2;
        w /= 2;
      }
    }
  }
  // upsampled tensors: (2h x 2w) x C of the coarser level
  max_act = std::max(max_act, (size_t)B * H * W * (size_t)(U.model_ch * U.channel_mult[std::min(1, nlev - 1)]));
  max_act += (size_t)16 * 2560;  // slack: V^T rows are padded to 16-column multiples
  bd.sA = A.get<bf16>(max_act);
  bd.sB = A.get<bf16>(max_act);
  bd.sC = A.get<bf16>(max_act);
  bd.sO = A.get<bf16>(max_act);
  bd.sVt = A.get<bf16>(max_act);
  bd.splitk_ws_bytes = (size_t)64 << 20;
  bd.splitk_ws = reinterpret_cast<float*>(A.alloc(bd.splitk_ws_bytes));
  bd.vt_pad_elems = (size_t)1280 * B * 8 * 8;  // only tiny token counts (< 64) can be non-multiples of 8
  bd.sVtPad = A.get<bf16>(bd.vt_pad_elems, true);
  bd.sG = A.get<bf16>(std::max<size_t>(max_geglu, 16));
  bd.sCol = A.get<bf16>(std::max<size_t>(max_col, 16));
  {
    int h = H, w = W;
    for (int level = 0; level < nlev; ++level) {
      const int c = U.model_ch * U.channel_mult[level];
      const int slot = slot_of(c / U.heads);
      bd.sQK.push_back(U.attn_level[level] ? A.get<bf16>((size_t)B * h * w * 2 * U.heads * slot, true) : nullptr);
      bf16* vt40 = nullptr;
      const int dh = c / U.heads;
      if (U.attn_level[level] && dh == 40) {
        // the K half's pad column 40 is never written by the projections (they scatter columns 0..39 of every head slot)
        fill_k_ones_kernel<<<64, 256>>>(bd.sQK.back(), (long long)B * h * w, 2LL * U.heads * slot, U.heads, slot, U.heads * slot, dh);
        LDN_CUDA(cudaGetLastError());
        LDN_CUDA(cudaDeviceSynchronize());
      }
      if (U.attn_level[level] && (dh == 40 || dh == 80) && (h * w) % 8 == 0) {
        const int hs = dh == 40 ? 48 : 96;
        const long long tld = ((long long)B * h * w + 15) / 16 * 16;
        vt40 = A.get<bf16>((size_t)U.heads * hs * tld, true);
        fill_ones_rows_kernel<<<64, 256>>>(vt40, U.heads, tld, hs, dh);
        LDN_CUDA(cudaGetLastError());
        LDN_CUDA(cudaDeviceSynchronize());
      }
      bd.sVt40.push_back(vt40);
      if (level != nlev - 1) {
        h /= 2;
        w /= 2;
      }
    }
    // the middle block runs at the last level's resolution with its channel count
    if (!U.attn_level[nlev - 1]) {
      const int c = U.model_ch * U.channel_mult[nlev - 1];
      bd.sQK[nlev - 1] = A.get<bf16>((size_t)B * h * w * 2 * U.heads * slot_of(c / U.heads), true);
    }
  }
  bd.gn_ws = reinterpret_cast<float*>(A.alloc(groupnorm_ws_bytes(B), true));
  {  // first node of the program: zero the GroupNorm statistics slots (fixed-point accumulators, norm.cu)
    float* ws = bd.gn_ws;
    const size_t bytes = groupnorm_ws_bytes(B);
    bd.add("groupnorm.zero_statistics", [=](cudaStream_t st) { LDN_CUDA(cudaMemsetAsync(ws, 0, bytes, st)); }, 0);
  }
  bd.ln_parts = A.get<float2>((size_t)B * H * W * 16, true);
  bd.ln_final = A.get<float2>((size_t)B * H * W + 16, true);
  bd.temb = A.get<float>((size_t)B * U.model_ch);
  bd.emb1 = A.get<float>((size_t)B * U.temb_dim);
  bd.emb = A.get<float>((size_t)B * U.temb_dim);
  bd.emb_all = A.get<float>((size_t)B * U.emb_total);
  prog->io_elems = (size_t)B * U.in_ch * H * W;
  prog->in_x = A.get<float>(prog->io_elems);
  prog->in_sigma = A.get<float>(B);
  prog->out = A.get<float>(prog->io_elems);

  // ---- time embedding (unet.py:705-708; ResBlock.py:270-278 emb_layers batched into one launch)
  {
    const float* sig = prog->in_sigma;
    const float* ls = e->log_sigmas;
    const int ns = e->n_sigmas;
    LDN_CHECK(ls != nullptr, "ldn_set_sigmas has not been called");
    float *temb = bd.temb, *emb1 = bd.emb1, *emb = bd.emb, *emb_all = bd.emb_all;
    const int mc = U.model_ch, td = U.temb_dim, et = U.emb_total;
    const bf16* w0 = e->W(0, "time_embed.0.weight").b();
    const float* b0 = e->W(0, "time_embed.0.bias").f();
    const bf16* w2 = e->W(0, "time_embed.2.weight").b();
    const float* b2 = e->W(0, "time_embed.2.bias").f();
    const bf16* wall = U.Wemb_all;
    const float* ball = U.bemb_all;
    bd.add("time_embed", [=](cudaStream_t st) {
      launch_timestep_embed(sig, B, ls, ns, mc, temb, nullptr, st);
      launch_small_linear(temb, B, mc, w0, b0, td, false, true, emb1, st);
      launch_small_linear(emb1, B, td, w2, b2, td, false, false, emb, st);
      launch_small_linear(emb, B, td, wall, ball, et, true, false, emb_all, st);
    }, 4);
  }
  // ---- input blocks
  std::vector<const bf16*> hs;
  std::vector<int> hs_c;
  const bf16* h = nullptr;
  int ch = 0, ch_h = H, ch_w = W, level = 0;
  for (auto& blk : U.input_blocks) {
    for (auto& bdsc : blk) {
      if (bdsc.kind == B_CONV_IN) {
        bf16* o = A.get<bf16>((size_t)B * H * W * U.model_ch);
        const float* x = prog->in_x;
        const float* sig = prog->in_sigma;
        const bf16* wt = e->W(0, "input_blocks.0.0.weight").b();
        const float* bias = e->W(0, "input_blocks.0.0.bias").f();
        const int cin = U.in_ch, cout = U.model_ch;
        bd.add("conv_in", [=](cudaStream_t st) { launch_conv_in(x, sig, wt, bias, B, H, W, cin, cout, o, st); });
        h = o;
        ch = cout;
      } else if (bdsc.kind == B_RES) {
        const ResW& r = U.res[bdsc.res];
        h = bd.resblock(r, h, ch, nullptr, 0, ch_h, ch_w);
        ch = r.cout;
      } else if (bdsc.kind == B_ST) {
        h = bd.transformer(U.sts[bdsc.st], h, ch_h, ch_w, level);
      } else if (bdsc.kind == B_DOWN) {
        const int ho = ch_h / 2, wo = ch_w / 2;
        bf16* o = A.get<bf16>((size_t)B * ho * wo * ch);
        const bf16* src = h;
        bf16* col = bd.sCol;
        const int hh = ch_h, ww = ch_w, cc = ch;
        bd.add(bdsc.prefix + ".gather", [=](cudaStream_t st) { launch_im2col_s2(src, B, hh, ww, cc, col, st); });
        GemmArgs a;
        a.A0 = col; a.lda0 = 9LL * ch; a.K0 = 9 * ch; a.Wt = e->W(0, bdsc.prefix + ".weight").b();
        a.M = B * ho * wo; a.N = ch; a.bias = e->W(0, bdsc.prefix + ".bias").f(); a.out = o; a.ldo = ch;
        bd.gemm(bdsc.prefix, a);
        h = o;
        ch_h = ho;
        ch_w = wo;
        ++level;
      }
    }
    hs.push_back(h);
    hs_c.push_back(ch);
  }
  // ---- middle
  for (auto& bdsc : U.middle) {
    if (bdsc.kind == B_RES) {
      h = bd.resblock(U.res[bdsc.res], h, ch, nullptr, 0, ch_h, ch_w);
    } else {
      h = bd.transformer(U.sts[bdsc.st], h, ch_h, ch_w, level);
    }
  }
  // ---- output blocks
  for (auto& blk : U.output_blocks) {
    const bf16* skip = hs.back();
    const int skip_c = hs_c.back();
    hs.pop_back();
    hs_c.pop_back();
    for (auto& bdsc : blk) {
      if (bdsc.kind == B_RES) {
        const ResW& r = U.res[bdsc.res];
        h = bd.resblock(r, h, ch, skip, skip_c, ch_h, ch_w);
        ch = r.cout;
      } else if (bdsc.kind == B_ST) {
        h = bd.transformer(U.sts[bdsc.st], h, ch_h, ch_w, level);
      } else if (bdsc.kind == B_UP) {
        // nearest 2x then conv3x3 (Upsample1, ResBlock.py:106-138)
        bf16* up = bd.sA;
        const bf16* src = h;
        const int hh = ch_h, ww = ch_w, cc = ch;
        bd.add(bdsc.prefix + ".nearest", [=](cudaStream_t st) { launch_upsample2x(src, B, hh, ww, cc, up, st); });
        ch_h *= 2;
        ch_w *= 2;
        --level;
        bf16* o = A.get<bf16>((size_t)B * ch_h * ch_w * ch);
        GemmArgs a;
        a.conv = true; a.A0 = up; a.B = B; a.H = ch_h; a.W = ch_w; a.Cin = ch;
        a.Wt = e->W(0, bdsc.prefix + ".weight").b(); a.N = ch; a.bias = e->W(0, bdsc.prefix + ".bias").f();
        a.out = o; a.ldo = ch;
        bd.gemm(bdsc.prefix, a);
        h = o;
      }
    }
  }
  LDN_CHECK(hs.empty() && ch_h == H && ch_w == W && ch == U.model_ch, "UNet program: skip stack / shape mismatch");
  // ---- head: GroupNorm + SiLU + conv3x3 -> eps; denoised = x - eps * sigma (fp32 NCHW)
  bd.groupnorm("out.gn", h, ch, nullptr, 0, H * W, 1e-5f, "out.0", true, bd.sA);
  {
    const float* bias = e->W(0, "out.2.bias").f();
    const float* x = prog->in_x;
    const float* sig = prog->in_sigma;
    float* out = prog->out;
    const int cin = ch, cout = U.out_ch;
    if (cin % 64 == 0 && cout <= 16) {
      // tensor-core path: implicit-GEMM conv with the 4 output channels padded to one 16-column MMA (weight rows past
      // cout read as zero through the tensor map), fp32 accumulators to a [pixels, 16] scratch, then the fused
      // bias / x - eps * sigma / NCHW pass
      float* acc16 = A.get<float>((size_t)B * H * W * 16);
      GemmArgs a;
      a.conv = true; a.A0 = bd.sA; a.B = B; a.H = H; a.W = W; a.Cin = cin;
      a.Wt = e->W(0, "out.2.weight").b(); a.N = 16; a.wt_rows = cout; a.BN = 16;
      a.out_f32 = acc16; a.ldo = 16;
      bd.gemm("out.conv", a);
      const int HW = H * W;
      bd.add("conv_out_finish", [=](cudaStream_t st) { launch_conv_out_finish(acc16, bias, x, sig, B, HW, cout, out, st); });
    } else {
      const bf16* src = bd.sA;
      const bf16* wt = e->W(0, "out.2.weight").b();
      bd.add("conv_out", [=](cudaStream_t st) {
        launch_conv_out(src, wt, bias, x, sig, B, H, W, cin, cout, out, nullptr, st);
      });
    }
  }
  return prog.release();
}

void unet_denoise(ldn_engine* e, const float* x, const float* sigma, float* out, int rows, int h, int w,
                  cudaStream_t stream) {
  auto& U = *e->unet;
  LDN_CHECK(U.ctx_rows > 0, "ldn_set_context must be called before ldn_unet_denoise");
  auto key = std::make_tuple(rows, h, w, U.ctx_rows, U.ctx_tokens);
  auto it = U.programs.find(key);
  if (it == U.programs.end()) {
    it = U.programs.emplace(key, std::unique_ptr<Program>(build_unet_program(e, rows, h, w))).first;
  }
  Program& P = *it->second;
  if (getenv("LDN_DEBUG_HASH") && P.arena)  // identical start state for every run so per-step checksums are comparable
    for (size_t b = 0; b < P.arena->blocks.size(); ++b) cudaMemsetAsync(P.arena->blocks[b], 0, P.arena->sizes[b], stream);
  LDN_CUDA(cudaMemcpyAsync(P.in_x, x, P.io_elems * sizeof(float), cudaMemcpyDeviceToDevice, stream));
  LDN_CUDA(cudaMemcpyAsync(P.in_sigma, sigma, rows * sizeof(float), cudaMemcpyDeviceToDevice, stream));
  run_program(P, e->cfg.use_graph != 0, stream);
  LDN_CUDA(cudaMemcpyAsync(out, P.out, P.io_elems * sizeof(float), cudaMemcpyDeviceToDevice, stream));
  U.last_launches = P.launches;
}

int unet_last_launches(ldn_engine* e) { return e->unet ? e->unet->last_launches : 0; }

}  // namespace ldn
