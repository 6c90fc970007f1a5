// SD1.x VAE decoder (AutoencoderKL) as a launch program over NHWC bf16 activations.
//
// Mirrors (structure, not code): VAE.decode / process_output    src/AutoEncoders/VariationalAE.py:602-604, 690-722
//   AutoencodingEngine.decode (post_quant_conv + Decoder)        VariationalAE.py:130-145
//   Decoder.forward                                              VariationalAE.py:532-567
//   ResnetBlock.forward (GroupNorm eps 1e-6 + swish)             src/AutoEncoders/ResBlock.py:383-406
//   AttnBlock.forward (single head, d = 512)                     src/Attention/Attention.py:159-178
//   Upsample (nearest 2x + conv3x3)                              VariationalAE.py:192-221
//
// Reuses the UNet's kernels: implicit-GEMM conv3x3 and GEMM on tcgen05, deterministic GroupNorm+SiLU. The mid-block
// attention (N = h*w tokens, one head of 512) is done with two GEMMs around a row softmax on a materialised score
// matrix per image, in query blocks of at most 1 GiB of fp32 logits (the whole N x N at 1024^2, 16 blocks at 2048^2);
// V is produced transposed by swapping GEMM operands, and its bias is folded into proj_out's bias
// (softmax rows sum to 1:  P (X Wv^T + 1 bv^T) Wp^T + bp = P X Wv^T Wp^T + (Wp bv + bp)).
#include <algorithm>
#include <cmath>
#include <map>

#include "engine.h"

using namespace ldn;

struct ldn_engine::VaeState {
  int ch = 128, zc = 4, out_ch = 3, num_res = 2;  // zc is re-read from decoder.conv_in at finalize
  std::vector<int> ch_mult = {1, 2, 4, 4};
  Arena arena;
  float* attn_bias = nullptr;      // decoder mid attention: Wp bv + bp
  float* attn_bias_enc = nullptr;  // encoder mid attention (only when encoder weights were loaded)
  float* quant_w = nullptr;        // quant_conv weight as fp32 [8, 8]
  bool has_decoder = false, has_encoder = false, has_post_quant = false;
  std::map<std::tuple<int, int, int>, std::unique_ptr<Program>> programs;
  std::map<std::tuple<int, int, int>, std::unique_ptr<Program>> enc_programs;
  std::vector<std::unique_ptr<Arena>> program_arenas;
};

namespace ldn {

void vae_finalize(ldn_engine* e, cudaStream_t stream) {
  LDN_CHECK(!e->w[1].empty(), "VAE weights not loaded");
  e->vae.reset(new ldn_engine::VaeState());
  auto& V = *e->vae;
  const int C = V.ch * V.ch_mult.back();
  V.has_decoder = e->has(1, "decoder.conv_in.weight");
  if (V.has_decoder) V.zc = (int)(e->W(1, "decoder.conv_in.weight").shape[1] / 9);  // 4: SD1.x, 16: Flux
  V.has_post_quant = e->has(1, "post_quant_conv.weight");  // absent in the Flux VAE (AutoencodingEngine flux=True)
  V.has_encoder = e->has(1, "encoder.conv_in.weight");
  LDN_CHECK(V.has_decoder || V.has_encoder, "VAE weights hold neither decoder.* nor encoder.* tensors");
  // attn_bias = proj_out.weight @ v.bias + proj_out.bias  (tiny mat-vec on the device)
  if (V.has_decoder) {
    const std::string p = "decoder.mid.attn_1";
    V.attn_bias = V.arena.get<float>(C);
    launch_small_linear(e->W(1, p + ".v.bias").f(), 1, C, e->W(1, p + ".proj_out.weight").b(),
                        e->W(1, p + ".proj_out.bias").f(), C, false, false, V.attn_bias, stream);
  }
  if (V.has_encoder) {
    const std::string p = "encoder.mid.attn_1";
    V.attn_bias_enc = V.arena.get<float>(C);
    launch_small_linear(e->W(1, p + ".v.bias").f(), 1, C, e->W(1, p + ".proj_out.weight").b(),
                        e->W(1, p + ".proj_out.bias").f(), C, false, false, V.attn_bias_enc, stream);
    const int z2 = 2 * V.zc;
    V.quant_w = V.arena.get<float>(z2 * z2);
    launch_convert_to_f32(e->W(1, "quant_conv.weight").p, 2, (size_t)z2 * z2, V.quant_w, stream);
  }
  LDN_CUDA(cudaStreamSynchronize(stream));
  e->finalized[1] = true;
}

namespace {
struct VB {
  ldn_engine* e;
  ldn_engine::VaeState& V;
  Program& P;
  Arena& A;
  int B;
  bf16 *sA = nullptr, *sB = nullptr, *sC = nullptr;
  float* gn_ws = nullptr;
  int gn_slots = 0;

  void add(const std::string& name, Step s, int launches = 1) {
    P.steps.push_back(std::move(s));
    P.names.push_back(name);
    P.launches += launches;
  }
  void gemm(const std::string& name, const GemmArgs& a) {
    GemmPlan plan = make_gemm_plan(a);
    const long long Mm = a.conv ? (long long)a.B * a.H * a.W : a.M;
    const long long Kk = a.conv ? 9LL * a.Cin : (long long)a.K0 + a.K1;
    add(name + " [M=" + std::to_string(Mm) + " N=" + std::to_string(a.N) + " K=" + std::to_string(Kk) + "]",
        [plan](cudaStream_t st) { launch_gemm(plan, st); });
  }
  void gn(const std::string& name, const bf16* x, int C, int HW, const std::string& wp, bool silu, bf16* out) {
    const float* g = e->W(1, wp + ".weight").f();
    const float* b = e->W(1, wp + ".bias").f();
    float* ws = gn_ws;
    const int Bn = B;
    const int slot = gn_slots++;
    LDN_CHECK(slot < LDN_GN_SLOTS, "too many GroupNorm instances for the statistics workspace");
    add(name, [=](cudaStream_t st) { launch_groupnorm(x, C, nullptr, 0, Bn, HW, 32, 1e-6f, g, b, silu, out, ws, slot, st); }, 2);
  }
  void conv(const std::string& name, const std::string& wp, const bf16* x, int H, int W, int Cin, int Cout,
            const bf16* residual, bf16* out) {
    GemmArgs a;
    a.conv = true; a.A0 = x; a.B = B; a.H = H; a.W = W; a.Cin = Cin;
    a.Wt = e->W(1, wp + ".weight").b(); a.N = Cout; a.bias = e->W(1, wp + ".bias").f();
    a.residual = residual; a.ldr = Cout; a.out = out; a.ldo = Cout;
    gemm(name, a);
  }
  bf16* resnet(const std::string& p, const bf16* x, int H, int W, int Cin, int Cout) {
    const int HW = H * W;
    bf16* out = A.get<bf16>((size_t)B * HW * Cout);
    gn(p + ".norm1", x, Cin, HW, p + ".norm1", true, sA);
    conv(p + ".conv1", p + ".conv1", sA, H, W, Cin, Cout, nullptr, sB);
    gn(p + ".norm2", sB, Cout, HW, p + ".norm2", true, sA);
    const bf16* res = x;
    if (Cin != Cout) {
      GemmArgs a;
      a.A0 = x; a.lda0 = Cin; a.K0 = Cin; a.Wt = e->W(1, p + ".nin_shortcut.weight").b(); a.M = B * HW; a.N = Cout;
      a.bias = e->W(1, p + ".nin_shortcut.bias").f(); a.out = sC; a.ldo = Cout;
      gemm(p + ".nin_shortcut", a);
      res = sC;
    }
    conv(p + ".conv2", p + ".conv2", sA, H, W, Cout, Cout, res, out);
    return out;
  }
  // AttnBlock (single head of C channels over all h*w pixels; src/Attention/Attention.py:159-178).
  // Any token count: per image, Q / K / V^T / S live in layouts padded to Np = roundup(N, 16) tokens. Pad rows of K and pad
  // columns of V^T are produced as exact zeros by the GEMMs themselves (operand rows past N read as zero through the
  // tensor map), so the pad columns of S are 0 before and after the in-place softmax over the first N columns and
  // contribute nothing to P V.
  const bf16* attn(const std::string& p, const bf16* x, int h, int w, int C, const float* attn_bias) {
    const int N = h * w, T = B * N;
    const int Np = (N + 15) / 16 * 16;
    bf16* q = A.get<bf16>((size_t)B * Np * C, true);
    bf16* k = A.get<bf16>((size_t)B * Np * C, true);
    bf16* vt = A.get<bf16>((size_t)C * B * Np + 64, true);
    bf16* o = A.get<bf16>((size_t)T * C);
    // The score matrix is materialised one QUERY BLOCK at a time, as FP32 LOGITS (QB x Np floats, capped at 1 GiB) plus the
    // bf16 probabilities the softmax writes (QB x Np, 0.5 GiB): QB = N up to 16384 tokens = a 1024^2 image, 16 blocks of 4096
    // queries at 2048^2 (where the whole N x N matrix would be 17 GB in fp32). Rows of a softmax are independent, so
    // blocking changes nothing numerically; fp32 logits keep the exponent's argument exact where a bf16 logit of
    // magnitude ~30 would carry an absolute error of ~0.06.
    int QB = N;
    {
      const long long cap_rows = ((1LL << 28) / Np) / 128 * 128;  // 2^28 elements = 1 GiB of fp32
      if (cap_rows >= 128 && cap_rows < N) QB = (int)cap_rows;
    }
    float* S = A.get<float>((size_t)QB * Np, true);
    bf16* P = A.get<bf16>((size_t)QB * Np, true);  // pad columns N..Np-1 stay zero: the softmax writes the first N only
    bf16* out = A.get<bf16>((size_t)T * C);
    gn(p + ".norm", x, C, N, p + ".norm", false, sA);
    const float scale = 1.0f / sqrtf((float)C);
    for (int b = 0; b < B; ++b) {
      const bf16* xb = sA + (size_t)b * N * C;
      for (const char* nm : {"q", "k"}) {
        GemmArgs a;
        a.A0 = xb; a.lda0 = C; a.K0 = C; a.Wt = e->W(1, p + "." + nm + ".weight").b(); a.M = N; a.N = C;
        a.bias = e->W(1, p + "." + nm + ".bias").f(); a.out = ((nm[0] == 'q') ? q : k) + (size_t)b * Np * C; a.ldo = C;
        gemm(p + "." + nm, a);
      }
      {
        GemmArgs a;  // V^T = Wv * X^T (bias folded into proj_out); token columns N..Np-1 come out as zeros
        a.A0 = e->W(1, p + ".v.weight").b(); a.lda0 = C; a.K0 = C; a.Wt = xb; a.wt_rows = N; a.M = C; a.N = Np;
        a.out = vt + (size_t)b * Np; a.ldo = (long long)B * Np;
        gemm(p + ".vt", a);
      }
      for (int q0 = 0; q0 < N; q0 += QB) {
        const int rows = std::min(QB, N - q0);
        GemmArgs s;
        s.A0 = q + ((size_t)b * Np + q0) * C; s.lda0 = C; s.K0 = C; s.Wt = k + (size_t)b * Np * C; s.M = rows; s.N = Np;
        s.out_f32 = S; s.ldo = Np;
        gemm(p + ".scores", s);
        float* Sp = S;
        bf16* Pp = P;
        add(p + ".softmax", [=](cudaStream_t st) { launch_softmax_rows_f32(Sp, Np, Pp, Np, rows, N, scale, st); });
        GemmArgs pv;
        pv.A0 = P; pv.lda0 = Np; pv.K0 = Np; pv.Wt = vt + (size_t)b * Np; pv.M = rows; pv.N = C;
        pv.out = o + ((size_t)b * N + q0) * C; pv.ldo = C;
        pv.wt_ld = (long long)B * Np;  // V^T rows are B*Np apart: describe the weight operand with its true leading dimension
        gemm(p + ".pv", pv);
      }
    }
    GemmArgs po;
    po.A0 = o; po.lda0 = C; po.K0 = C; po.Wt = e->W(1, p + ".proj_out.weight").b(); po.M = T; po.N = C;
    po.bias = attn_bias; po.residual = x; po.ldr = C; po.out = out; po.ldo = C;
    gemm(p + ".proj_out", po);
    return out;
  }
};
}  // namespace

static Program* build_vae_program(ldn_engine* e, int B, int h, int w) {
  auto& V = *e->vae;
  std::unique_ptr<Program> prog(new Program());
  V.program_arenas.emplace_back(new Arena());
  Arena& A = *V.program_arenas.back();
  prog->arena = &A;
  VB vb{e, V, *prog, A, B};
  const int nlev = (int)V.ch_mult.size();
  // largest activation: the tensor entering the last level after its upsample (ch*mult[1] channels at full resolution)
  size_t max_act = 0;
  {
    int hh = h, ww = w;
    int c = V.ch * V.ch_mult[nlev - 1];
    for (int lvl = nlev - 1; lvl >= 0; --lvl) {
      const int cout = V.ch * V.ch_mult[lvl];
      max_act = std::max(max_act, (size_t)B * hh * ww * std::max(c, cout));
      c = cout;
      if (lvl != 0) {
        hh *= 2;
        ww *= 2;
        max_act = std::max(max_act, (size_t)B * hh * ww * c);
      }
    }
  }
  vb.sA = A.get<bf16>(max_act);
  vb.sB = A.get<bf16>(max_act);
  vb.sC = A.get<bf16>(max_act);
  vb.gn_ws = reinterpret_cast<float*>(A.alloc(groupnorm_ws_bytes(B), true));
  {  // first node of the program: zero the GroupNorm statistics slots
    float* ws = vb.gn_ws;
    const size_t bytes = groupnorm_ws_bytes(B);
    vb.add("groupnorm.zero_statistics", [=](cudaStream_t st) { LDN_CUDA(cudaMemsetAsync(ws, 0, bytes, st)); }, 0);
  }
  prog->io_elems = (size_t)B * V.zc * h * w;
  prog->in_x = A.get<float>(prog->io_elems);
  float* zq = A.get<float>(prog->io_elems);
  const size_t out_elems = (size_t)B * (8 * h) * (8 * w) * 3;
  prog->out = A.get<float>(out_elems);

  const int C = V.ch * V.ch_mult[nlev - 1];
  // post_quant_conv (1x1 on the fp32 latent; SD1.x only), conv_in zc -> 512
  const float* zin = prog->in_x;
  if (V.has_post_quant) {
    const float* x = prog->in_x;
    const int zc = V.zc, HW = h * w;
    // post_quant_conv weights are stored bf16 [zc,zc] by the generic ingest; convert once into fp32 scratch
    float* wq = A.get<float>(zc * zc);
    launch_convert_to_f32(e->W(1, "post_quant_conv.weight").p, 2, (size_t)zc * zc, wq, 0);
    LDN_CUDA(cudaDeviceSynchronize());
    const float* bq = e->W(1, "post_quant_conv.bias").f();
    vb.add("post_quant_conv", [=](cudaStream_t st) { launch_conv1x1_f32(x, wq, bq, B, zc, zc, HW, zq, st); });
    zin = zq;
  }
  bf16* hcur = A.get<bf16>((size_t)B * h * w * C);
  {
    const bf16* wt = e->W(1, "decoder.conv_in.weight").b();
    const float* bias = e->W(1, "decoder.conv_in.bias").f();
    const int zc = V.zc;
    // channel slices so that the slice's fp32 weights (slice * 9 * zc floats) fit in shared memory (16-channel Flux latents)
    int slice = C;
    while ((size_t)slice * 9 * zc * sizeof(float) > 160 * 1024 && slice % 16 == 0) slice /= 2;
    for (int c0 = 0; c0 < C; c0 += slice) {
      const bf16* wts = wt + (size_t)c0 * 9 * zc;
      const float* bs = bias + c0;
      bf16* o = hcur + c0;
      const int n = std::min(slice, C - c0);
      vb.add("decoder.conv_in", [=](cudaStream_t st) { launch_conv_in(zin, nullptr, wts, bs, B, h, w, zc, n, o, st, 0, C); });
    }
  }
  const bf16* x = vb.resnet("decoder.mid.block_1", hcur, h, w, C, C);
  x = vb.attn("decoder.mid.attn_1", x, h, w, C, V.attn_bias);
  x = vb.resnet("decoder.mid.block_2", x, h, w, C, C);
  int hh = h, ww = w, c = C;
  for (int lvl = nlev - 1; lvl >= 0; --lvl) {
    const int cout = V.ch * V.ch_mult[lvl];
    for (int i = 0; i <= V.num_res; ++i) {
      x = vb.resnet("decoder.up." + std::to_string(lvl) + ".block." + std::to_string(i), x, hh, ww, c, cout);
      c = cout;
    }
    if (lvl != 0) {
      const std::string p = "decoder.up." + std::to_string(lvl) + ".upsample.conv";
      const bf16* src = x;
      bf16* up = vb.sA;
      const int h0 = hh, w0 = ww, cc = c;
      vb.add(p + ".nearest", [=](cudaStream_t st) { launch_upsample2x(src, B, h0, w0, cc, up, st); });
      hh *= 2;
      ww *= 2;
      bf16* o = A.get<bf16>((size_t)B * hh * ww * c);
      vb.conv(p, p, up, hh, ww, c, c, nullptr, o);
      x = o;
    }
  }
  vb.gn("decoder.norm_out", x, c, hh * ww, "decoder.norm_out", true, vb.sA);
  {
    const float* bias = e->W(1, "decoder.conv_out.bias").f();
    float* out = prog->out;
    const int H8 = hh, W8 = ww, cin = c, cout = V.out_ch;
    if (cin % 64 == 0 && cout <= 4) {
      // tensor-core path: 3 output channels padded to one 16-column MMA (weight rows past 3 read as zero), fp32 scratch
      float* acc16 = A.get<float>((size_t)B * H8 * W8 * 16);
      GemmArgs a;
      a.conv = true; a.A0 = vb.sA; a.B = B; a.H = H8; a.W = W8; a.Cin = cin;
      a.Wt = e->W(1, "decoder.conv_out.weight").b(); a.N = 16; a.wt_rows = cout; a.BN = 16;
      a.out_f32 = acc16; a.ldo = 16;
      vb.gemm("decoder.conv_out", a);
      const size_t npix = (size_t)B * H8 * W8;
      vb.add("decoder.rgb", [=](cudaStream_t st) { launch_vae_rgb_finish(acc16, bias, npix, cout, out, st); });
    } else {
      const bf16* src = vb.sA;
      const bf16* wt = e->W(1, "decoder.conv_out.weight").b();
      vb.add("decoder.conv_out", [=](cudaStream_t st) { launch_conv_out_rgb(src, wt, bias, B, H8, W8, cin, out, st); });
    }
  }
  return prog.release();
}

// ------------------------------------------------------------------------------------------------------------------
// Encoder (VAE.encode -> AutoencodingEngine.encode -> Encoder.forward; VariationalAE.py:725-760, 148-172, 377-413).
// Input: fp32 NCHW [B, 3, H, W] already mapped to [-1, 1] (process_input); output: Gaussian moments fp32 NCHW
// [B, 8, H/8, W/8] (mean | logvar) after quant_conv. The reparameterised sample stays with the host (torch RNG).
static Program* build_vae_enc_program(ldn_engine* e, int B, int H, int W) {
  auto& V = *e->vae;
  std::unique_ptr<Program> prog(new Program());
  V.program_arenas.emplace_back(new Arena());
  Arena& A = *V.program_arenas.back();
  prog->arena = &A;
  VB vb{e, V, *prog, A, B};
  const int nlev = (int)V.ch_mult.size();
  size_t max_act = 0, max_col = 0;
  {
    int hh = H, ww = W, c = V.ch;
    for (int lvl = 0; lvl < nlev; ++lvl) {
      const int cout = V.ch * V.ch_mult[lvl];
      max_act = std::max(max_act, (size_t)B * hh * ww * std::max(c, cout));
      c = cout;
      if (lvl != nlev - 1) {
        hh /= 2;
        ww /= 2;
        max_col = std::max(max_col, (size_t)B * hh * ww * 9 * c);
      }
    }
  }
  vb.sA = A.get<bf16>(max_act);
  vb.sB = A.get<bf16>(max_act);
  vb.sC = A.get<bf16>(max_act);
  bf16* col = A.get<bf16>(std::max<size_t>(max_col, 16));
  vb.gn_ws = reinterpret_cast<float*>(A.alloc(groupnorm_ws_bytes(B), true));
  {  // first node of the program: zero the GroupNorm statistics slots
    float* ws = vb.gn_ws;
    const size_t bytes = groupnorm_ws_bytes(B);
    vb.add("groupnorm.zero_statistics", [=](cudaStream_t st) { LDN_CUDA(cudaMemsetAsync(ws, 0, bytes, st)); }, 0);
  }
  prog->io_elems = (size_t)B * 3 * H * W;
  prog->in_x = A.get<float>(prog->io_elems);

  int hh = H, ww = W, c = V.ch;
  bf16* h0 = A.get<bf16>((size_t)B * H * W * c);
  {
    const float* x = prog->in_x;
    const bf16* wt = e->W(1, "encoder.conv_in.weight").b();
    const float* bias = e->W(1, "encoder.conv_in.bias").f();
    const int cc = c;
    vb.add("encoder.conv_in", [=](cudaStream_t st) { launch_conv_in(x, nullptr, wt, bias, B, H, W, 3, cc, h0, st); });
  }
  const bf16* x = h0;
  for (int lvl = 0; lvl < nlev; ++lvl) {
    const int cout = V.ch * V.ch_mult[lvl];
    for (int i = 0; i < V.num_res; ++i) {
      x = vb.resnet("encoder.down." + std::to_string(lvl) + ".block." + std::to_string(i), x, hh, ww, c, cout);
      c = cout;
    }
    if (lvl != nlev - 1) {
      const std::string p = "encoder.down." + std::to_string(lvl) + ".downsample.conv";
      const int ho = hh / 2, wo = ww / 2, h1 = hh, w1 = ww, cc = c;
      const bf16* src = x;
      vb.add(p + ".gather", [=](cudaStream_t st) { launch_im2col_s2(src, B, h1, w1, cc, col, st, 0); });
      bf16* o = A.get<bf16>((size_t)B * ho * wo * c);
      GemmArgs a;
      a.A0 = col; a.lda0 = 9LL * c; a.K0 = 9 * c; a.Wt = e->W(1, p + ".weight").b(); a.M = B * ho * wo; a.N = c;
      a.bias = e->W(1, p + ".bias").f(); a.out = o; a.ldo = c;
      vb.gemm(p, a);
      x = o;
      hh = ho;
      ww = wo;
    }
  }
  x = vb.resnet("encoder.mid.block_1", x, hh, ww, c, c);
  x = vb.attn("encoder.mid.attn_1", x, hh, ww, c, V.attn_bias_enc);
  x = vb.resnet("encoder.mid.block_2", x, hh, ww, c, c);
  vb.gn("encoder.norm_out", x, c, hh * ww, "encoder.norm_out", true, vb.sA);
  const int z2 = 2 * V.zc;
  prog->out = A.get<float>((size_t)B * z2 * hh * ww);
  {
    float* acc16 = A.get<float>((size_t)B * hh * ww * 16);
    GemmArgs a;
    a.conv = true; a.A0 = vb.sA; a.B = B; a.H = hh; a.W = ww; a.Cin = c;
    a.Wt = e->W(1, "encoder.conv_out.weight").b(); a.N = 16; a.wt_rows = z2; a.BN = 16;
    a.out_f32 = acc16; a.ldo = 16;
    vb.gemm("encoder.conv_out", a);
    const float* bc = e->W(1, "encoder.conv_out.bias").f();
    const float* wq = V.quant_w;
    const float* bq = e->W(1, "quant_conv.bias").f();
    float* out = prog->out;
    const int HW = hh * ww;
    vb.add("quant_conv", [=](cudaStream_t st) { launch_vae_moments_finish(acc16, bc, wq, bq, B, HW, z2, out, st); });
  }
  return prog.release();
}

void vae_encode(ldn_engine* e, const float* pixels, float* moments, int B, int H, int W, cudaStream_t stream) {
  auto& V = *e->vae;
  LDN_CHECK(V.has_encoder, "ldn_vae_encode: the loaded VAE weights hold no encoder.* tensors");
  LDN_CHECK(H >= 8 && W >= 8, "ldn_vae_encode: image smaller than one latent pixel");
  auto key = std::make_tuple(B, H, W);
  auto it = V.enc_programs.find(key);
  if (it == V.enc_programs.end())
    it = V.enc_programs.emplace(key, std::unique_ptr<Program>(build_vae_enc_program(e, B, H, W))).first;
  Program& P = *it->second;
  LDN_CUDA(cudaMemcpyAsync(P.in_x, pixels, P.io_elems * sizeof(float), cudaMemcpyDeviceToDevice, stream));
  run_program(P, e->cfg.use_graph != 0, stream);
  const size_t n = (size_t)B * 2 * V.zc * (H / 8) * (W / 8);
  LDN_CUDA(cudaMemcpyAsync(moments, P.out, n * sizeof(float), cudaMemcpyDeviceToDevice, stream));
}

void vae_decode(ldn_engine* e, const float* z, float* rgb, int B, int h, int w, cudaStream_t stream) {
  auto& V = *e->vae;
  LDN_CHECK(V.has_decoder, "ldn_vae_decode: the loaded VAE weights hold no decoder.* tensors");
  auto key = std::make_tuple(B, h, w);
  auto it = V.programs.find(key);
  if (it == V.programs.end()) it = V.programs.emplace(key, std::unique_ptr<Program>(build_vae_program(e, B, h, w))).first;
  Program& P = *it->second;
  LDN_CUDA(cudaMemcpyAsync(P.in_x, z, P.io_elems * sizeof(float), cudaMemcpyDeviceToDevice, stream));
  run_program(P, e->cfg.use_graph != 0, stream);
  LDN_CUDA(cudaMemcpyAsync(rgb, P.out, (size_t)B * 8 * h * 8 * w * 3 * sizeof(float), cudaMemcpyDeviceToDevice, stream));
}

}  // namespace ldn
