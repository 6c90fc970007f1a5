// Flux.1 DiT forward (double-stream + single-stream blocks) as a launch program over token-major bf16 activations.
//
// Mirrors (structure, not code): Flux3.forward_orig src/BlackForest/Flux.py:658-730, DoubleStreamBlock :260-348,
// SingleStreamBlock :351-417, LastLayer :420-471, Modulation :231-257, QKNorm :173-200, attention / rope :18-82.
//
// One program serves one sequence (one batch row); ldn_flux_forward loops over the rows of a call. Layout decisions:
//   * ONE residual buffer X [Nt + Ni, C] holds the text rows first and the image rows after them from the start, so the
//     torch.cat((txt, img), 1) between the double and the single blocks (Flux.py:715) never happens; the double blocks work
//     on the two row ranges, the single blocks on the whole buffer.
//   * qkv projections are split by weight rows: [q | k] -> one GEMM into the joint [N, 2C] buffer the attention kernel
//     reads (head dim 128 = one 128-column slot per head), v -> produced already transposed by swapping the GEMM operands;
//     the v bias is folded into the following projection's bias (softmax rows sum to one).
//   * RMS q/k-norm + RoPE: one in-place pass over [q | k]; LayerNorm + (1 + scale) x + shift: one pass; the gates ride in
//     the GEMM epilogue (acc + bias) * gate + residual, GELU(tanh) in the epilogue of the MLP's first GEMM.
//   * single blocks: linear1 is three GEMMs over row slices of its weight ([q|k], v^T, mlp with GELU), linear2 reads the
//     virtual concat [attention | mlp] as two K segments.
#include <cmath>
#include <map>

#include "engine.h"

using namespace ldn;

struct ldn_engine::FluxState {
  int C = 0, heads = 0, M = 0, depth = 0, depth_single = 0, ctx_dim = 0, vec_dim = 0, in_dim = 0;
  bool guidance = false;
  Arena arena;
  std::vector<float*> proj_bias_img, proj_bias_txt, lin2_bias;  // projection biases with the v bias folded in
  // every Modulation.lin / adaLN weight stacked into one [mod_total, C] matrix: all of them are functions of the same
  // conditioning vector, so ONE weight-streaming mat-vec at the start of the forward replaces 77 launches
  bf16* mod_w = nullptr;
  float* mod_b = nullptr;
  long long mod_total = 0;
  std::vector<long long> mod_off_img, mod_off_txt, mod_off_single;
  long long mod_off_final = 0;
  std::map<std::tuple<int, int>, std::unique_ptr<Program>> programs;
  std::vector<std::unique_ptr<Arena>> arenas;
  // per-program I/O (indexed like programs)
  struct IO {
    float *img = nullptr, *ctx = nullptr, *pe = nullptr, *t = nullptr, *g = nullptr, *y = nullptr, *out = nullptr;
  };
  std::map<std::tuple<int, int>, IO> io;
};

namespace ldn {

static void flux_finalize(ldn_engine* e, cudaStream_t stream) {
  LDN_CHECK(!e->w[4].empty(), "Flux weights not loaded");
  e->flux.reset(new ldn_engine::FluxState());
  auto& F = *e->flux;
  const DevTensor& w_in = e->W(4, "img_in.weight");
  F.C = (int)w_in.shape[0];
  F.in_dim = (int)w_in.shape[1];
  LDN_CHECK(F.C % 128 == 0, "Flux: hidden size must be a multiple of the 128-wide head");
  F.heads = F.C / 128;
  F.M = (int)e->W(4, "double_blocks.0.img_mlp.0.weight").shape[0];
  F.ctx_dim = (int)e->W(4, "txt_in.weight").shape[1];
  F.vec_dim = (int)e->W(4, "vector_in.in_layer.weight").shape[1];
  F.guidance = e->has(4, "guidance_in.in_layer.weight");
  while (e->has(4, "double_blocks." + std::to_string(F.depth) + ".img_mod.lin.weight")) ++F.depth;
  while (e->has(4, "single_blocks." + std::to_string(F.depth_single) + ".linear1.weight")) ++F.depth_single;
  LDN_CHECK(F.depth > 0 && F.depth_single >= 0, "Flux: no double blocks found");
  const int C = F.C;
  // fold the v bias into the projection that follows the attention: W_p (o + b_v) + b_p = W_p o + (W_p b_v + b_p)
  for (int b = 0; b < F.depth; ++b) {
    for (int s = 0; s < 2; ++s) {
      const std::string p = "double_blocks." + std::to_string(b) + (s == 0 ? ".img_attn" : ".txt_attn");
      float* dst = F.arena.get<float>(C);
      launch_small_linear(e->W(4, p + ".qkv.bias").f() + 2 * C, 1, C, e->W(4, p + ".proj.weight").b(),
                          e->W(4, p + ".proj.bias").f(), C, false, false, dst, stream);
      (s == 0 ? F.proj_bias_img : F.proj_bias_txt).push_back(dst);
    }
  }
  for (int b = 0; b < F.depth_single; ++b) {
    const std::string p = "single_blocks." + std::to_string(b);
    float* dst = F.arena.get<float>(C);
    launch_small_linear(e->W(4, p + ".linear1.bias").f() + 2 * C, 1, C, e->W(4, p + ".linear2.weight").b(),
                        e->W(4, p + ".linear2.bias").f(), C, false, false, dst, stream, (long long)C + F.M);
    F.lin2_bias.push_back(dst);
  }
  {
    std::vector<std::pair<std::string, long long>> parts;  // key, rows
    for (int b = 0; b < F.depth; ++b) {
      const std::string p = "double_blocks." + std::to_string(b);
      F.mod_off_img.push_back(F.mod_total);
      parts.push_back({p + ".img_mod.lin", 6LL * C});
      F.mod_total += 6LL * C;
      F.mod_off_txt.push_back(F.mod_total);
      parts.push_back({p + ".txt_mod.lin", 6LL * C});
      F.mod_total += 6LL * C;
    }
    for (int b = 0; b < F.depth_single; ++b) {
      F.mod_off_single.push_back(F.mod_total);
      parts.push_back({"single_blocks." + std::to_string(b) + ".modulation.lin", 3LL * C});
      F.mod_total += 3LL * C;
    }
    F.mod_off_final = F.mod_total;
    parts.push_back({"final_layer.adaLN_modulation.1", 2LL * C});
    F.mod_total += 2LL * C;
    F.mod_w = F.arena.get<bf16>((size_t)F.mod_total * C);
    F.mod_b = F.arena.get<float>((size_t)F.mod_total);
    long long off = 0;
    for (auto& pr : parts) {
      const DevTensor& w = e->W(4, pr.first + ".weight");
      LDN_CHECK((long long)w.shape[0] == pr.second && (int)w.shape[1] == C, "Flux: unexpected modulation weight shape: " + pr.first);
      LDN_CUDA(cudaMemcpyAsync(F.mod_w + (size_t)off * C, w.p, (size_t)pr.second * C * sizeof(bf16), cudaMemcpyDeviceToDevice, stream));
      LDN_CUDA(cudaMemcpyAsync(F.mod_b + off, e->W(4, pr.first + ".bias").p, (size_t)pr.second * sizeof(float),
                               cudaMemcpyDeviceToDevice, stream));
      off += pr.second;
    }
  }
  LDN_CUDA(cudaStreamSynchronize(stream));
  e->finalized[4] = true;
}

static Program* build_flux_program(ldn_engine* e, int Ni, int Nt, ldn_engine::FluxState::IO& io) {
  auto& F = *e->flux;
  std::unique_ptr<Program> prog(new Program());
  F.arenas.emplace_back(new Arena());
  Arena& A = *F.arenas.back();
  prog->arena = &A;
  Program& P = *prog;
  const int C = F.C, H = F.heads, M = F.M, N = Nt + Ni;
  const int Np = (N + 15) / 16 * 16;
  auto add = [&](const std::string& name, Step s) {
    P.steps.push_back(std::move(s));
    P.names.push_back(name);
    P.launches += 1;
  };
  auto gemm = [&](const std::string& name, const GemmArgs& a) {
    GemmPlan plan = make_gemm_plan(a);
    const long long Kk = (long long)a.K0 + a.K1;
    add(name + " [M=" + std::to_string(a.M) + " N=" + std::to_string(a.N) + " K=" + std::to_string(Kk) + "]",
        [plan](cudaStream_t st) { launch_gemm(plan, st); });
  };
  auto small = [&](const std::string& name, const float* x, int K, const std::string& wkey, int n_out, bool silu_in,
                   bool silu_out, float* out) {
    const bf16* w = e->W(4, wkey + ".weight").b();
    const float* b = e->W(4, wkey + ".bias").f();
    add(name, [=](cudaStream_t st) { launch_small_linear(x, 1, K, w, b, n_out, silu_in, silu_out, out, st); });
  };
  // ---- I/O staging (graph-stable addresses)
  io.img = A.get<float>((size_t)Ni * F.in_dim);
  io.ctx = A.get<float>((size_t)Nt * F.ctx_dim);
  io.pe = A.get<float>((size_t)N * 128);
  io.t = A.get<float>(1);
  io.g = A.get<float>(1);
  io.y = A.get<float>(F.vec_dim);
  io.out = A.get<float>((size_t)Ni * F.in_dim);
  // ---- activations
  bf16* X = A.get<bf16>((size_t)N * C);            // residual stream: text rows, then image rows
  bf16* sA = A.get<bf16>((size_t)N * C);           // modulated LayerNorm output
  bf16* QK = A.get<bf16>((size_t)Np * 2 * C, true);  // [q | k], head dim 128
  bf16* Vt = A.get<bf16>((size_t)C * Np + 64, true);
  bf16* O = A.get<bf16>((size_t)N * C);
  bf16* Hm = A.get<bf16>((size_t)N * M);
  bf16* in16 = A.get<bf16>((size_t)std::max((size_t)Ni * F.in_dim, (size_t)Nt * F.ctx_dim));
  float* te = A.get<float>(256);
  float* h1 = A.get<float>(C);
  float *v_t = A.get<float>(C), *v_g = A.get<float>(C), *v_y = A.get<float>(C), *vec = A.get<float>(C);
  float* mod_all = A.get<float>((size_t)F.mod_total);

  // ---- conditioning vector (Flux.py:676-689): time_in(temb(t)) [+ guidance_in(temb(g))] + vector_in(y)
  add("temb.t", [=](cudaStream_t st) { launch_flux_temb(io.t, 1, te, st); });
  small("time_in.in", te, 256, "time_in.in_layer", C, false, true, h1);
  small("time_in.out", h1, C, "time_in.out_layer", C, false, false, v_t);
  if (F.guidance) {
    add("temb.g", [=](cudaStream_t st) { launch_flux_temb(io.g, 1, te, st); });
    small("guidance_in.in", te, 256, "guidance_in.in_layer", C, false, true, h1);
    small("guidance_in.out", h1, C, "guidance_in.out_layer", C, false, false, v_g);
  }
  small("vector_in.in", io.y, F.vec_dim, "vector_in.in_layer", C, false, true, h1);
  small("vector_in.out", h1, C, "vector_in.out_layer", C, false, false, v_y);
  {
    const float* g = F.guidance ? v_g : nullptr;
    add("vec", [=](cudaStream_t st) { launch_vec_add3(v_t, g, v_y, C, vec, st); });
  }
  {
    const bf16* mw = F.mod_w;
    const float* mb = F.mod_b;
    const int total = (int)F.mod_total;
    add("modulation.all", [=](cudaStream_t st) { launch_small_linear(vec, 1, C, mw, mb, total, true, false, mod_all, st); });
  }
  // ---- img_in / txt_in into the two row ranges of X
  bf16* Xt = X;
  bf16* Xi = X + (size_t)Nt * C;
  {
    const size_t n = (size_t)Nt * F.ctx_dim;
    add("txt.cast", [=](cudaStream_t st) { launch_convert_to_bf16(io.ctx, 0, n, in16, st); });
    GemmArgs a;
    a.A0 = in16; a.lda0 = F.ctx_dim; a.K0 = F.ctx_dim; a.Wt = e->W(4, "txt_in.weight").b(); a.M = Nt; a.N = C;
    a.bias = e->W(4, "txt_in.bias").f(); a.out = Xt; a.ldo = C;
    gemm("txt_in", a);
  }
  {
    const size_t n = (size_t)Ni * F.in_dim;
    add("img.cast", [=](cudaStream_t st) { launch_convert_to_bf16(io.img, 0, n, in16, st); });
    GemmArgs a;
    a.A0 = in16; a.lda0 = F.in_dim; a.K0 = F.in_dim; a.Wt = e->W(4, "img_in.weight").b(); a.M = Ni; a.N = C;
    a.bias = e->W(4, "img_in.bias").f(); a.out = Xi; a.ldo = C;
    gemm("img_in", a);
  }
  const float scale = 1.0f / sqrtf(128.0f);
  auto attention = [&](const std::string& name) {
    AttnArgs at;
    at.Q = QK; at.ldq = 2LL * C; at.K = QK + C; at.ldk = 2LL * C;
    at.Vt = Vt; at.ldvt = Np; at.vt_rows = C; at.vt_head_stride = 0;
    at.B = 1; at.heads = H; at.Nq = N; at.Nk = N; at.nk_pad = Np; at.d = 128; at.slot = 128;
    at.scale = scale; at.out = O; at.ldo = C;
    AttnPlan plan = make_attn_plan(at);
    add(name, [plan](cudaStream_t st) { launch_attn(plan, st); });
  };
  // [q | k] and v^T of `rows` tokens starting at row r0, from weight rows [0, 2C) / [2C, 3C) of `wkey`
  auto qkv = [&](const std::string& name, const std::string& wkey, int r0, int rows, const std::string& norm_prefix) {
    const bf16* W = e->W(4, wkey + ".weight").b();
    const float* b = e->W(4, wkey + ".bias").f();
    GemmArgs a;
    a.A0 = sA + (size_t)r0 * C; a.lda0 = C; a.K0 = C; a.Wt = W; a.M = rows; a.N = 2 * C; a.bias = b;
    a.out = QK + (size_t)r0 * 2 * C; a.ldo = 2LL * C;
    gemm(name + ".qk", a);
    GemmArgs v;  // v^T = W_v * A^T; token columns past `rows` come out as zeros (and are overwritten by the next range)
    v.A0 = W + (size_t)2 * C * C; v.lda0 = C; v.K0 = C; v.Wt = sA + (size_t)r0 * C; v.wt_rows = rows; v.M = C;
    v.N = (rows + 15) / 16 * 16; v.out = Vt + r0; v.ldo = Np;
    gemm(name + ".vt", v);
    const float* qs = e->W(4, norm_prefix + ".query_norm.scale").f();
    const float* ks = e->W(4, norm_prefix + ".key_norm.scale").f();
    bf16* qk = QK + (size_t)r0 * 2 * C;
    const float* pe = io.pe + (size_t)r0 * 128;
    add(name + ".norm_rope", [=](cudaStream_t st) { launch_qk_norm_rope(qk, 2LL * C, rows, H, qs, ks, pe, st); });
  };
  // ---- double-stream blocks
  for (int b = 0; b < F.depth; ++b) {
    const std::string p = "double_blocks." + std::to_string(b);
    float* mod_i = mod_all + F.mod_off_img[b];
    float* mod_t = mod_all + F.mod_off_txt[b];
    // image rows must write v^T after the text rows (a text range that is not a multiple of 16 zero-pads into the image columns)
    add(p + ".txt.ln1", [=](cudaStream_t st) { launch_modln(Xt, Nt, C, mod_t, mod_t + C, sA, st); });
    qkv(p + ".txt", p + ".txt_attn.qkv", 0, Nt, p + ".txt_attn.norm");
    add(p + ".img.ln1", [=](cudaStream_t st) { launch_modln(Xi, Ni, C, mod_i, mod_i + C, sA + (size_t)Nt * C, st); });
    qkv(p + ".img", p + ".img_attn.qkv", Nt, Ni, p + ".img_attn.norm");
    attention(p + ".sdpa");
    for (int s = 0; s < 2; ++s) {
      const std::string t = s == 0 ? ".img" : ".txt";
      bf16* Xs = s == 0 ? Xi : Xt;
      const int r0 = s == 0 ? Nt : 0, rows = s == 0 ? Ni : Nt;
      const float* mod = s == 0 ? mod_i : mod_t;
      GemmArgs a;
      a.A0 = O + (size_t)r0 * C; a.lda0 = C; a.K0 = C; a.Wt = e->W(4, p + t + "_attn.proj.weight").b(); a.M = rows; a.N = C;
      a.bias = s == 0 ? F.proj_bias_img[b] : F.proj_bias_txt[b];
      a.colgate = mod + 2 * C; a.ld_colgate = 0; a.residual = Xs; a.ldr = C; a.out = Xs; a.ldo = C;
      gemm(p + t + ".proj", a);
      bf16* sAs = sA + (size_t)r0 * C;
      add(p + t + ".ln2", [=](cudaStream_t st) { launch_modln(Xs, rows, C, mod + 3 * C, mod + 4 * C, sAs, st); });
      GemmArgs m0;
      m0.A0 = sAs; m0.lda0 = C; m0.K0 = C; m0.Wt = e->W(4, p + t + "_mlp.0.weight").b(); m0.M = rows; m0.N = M;
      m0.bias = e->W(4, p + t + "_mlp.0.bias").f(); m0.act = 4; m0.out = Hm + (size_t)r0 * M; m0.ldo = M;
      gemm(p + t + ".mlp0", m0);
      GemmArgs m2;
      m2.A0 = Hm + (size_t)r0 * M; m2.lda0 = M; m2.K0 = M; m2.Wt = e->W(4, p + t + "_mlp.2.weight").b(); m2.M = rows; m2.N = C;
      m2.bias = e->W(4, p + t + "_mlp.2.bias").f(); m2.colgate = mod + 5 * C; m2.residual = Xs; m2.ldr = C; m2.out = Xs;
      m2.ldo = C;
      gemm(p + t + ".mlp2", m2);
    }
  }
  // ---- single-stream blocks over the whole buffer
  for (int b = 0; b < F.depth_single; ++b) {
    const std::string p = "single_blocks." + std::to_string(b);
    float* mod_i = mod_all + F.mod_off_single[b];
    add(p + ".ln", [=](cudaStream_t st) { launch_modln(X, N, C, mod_i, mod_i + C, sA, st); });
    qkv(p, p + ".linear1", 0, N, p + ".norm");
    {
      const bf16* W = e->W(4, p + ".linear1.weight").b();
      GemmArgs m0;
      m0.A0 = sA; m0.lda0 = C; m0.K0 = C; m0.Wt = W + (size_t)3 * C * C; m0.M = N; m0.N = M;
      m0.bias = e->W(4, p + ".linear1.bias").f() + 3 * C; m0.act = 4; m0.out = Hm; m0.ldo = M;
      gemm(p + ".mlp", m0);
    }
    attention(p + ".sdpa");
    GemmArgs l2;
    l2.A0 = O; l2.lda0 = C; l2.K0 = C; l2.A1 = Hm; l2.lda1 = M; l2.K1 = M; l2.Wt = e->W(4, p + ".linear2.weight").b();
    l2.M = N; l2.N = C; l2.bias = F.lin2_bias[b]; l2.colgate = mod_i + 2 * C; l2.residual = X; l2.ldr = C; l2.out = X; l2.ldo = C;
    gemm(p + ".linear2", l2);
  }
  // ---- LastLayer on the image rows (Flux.py:458-471): shift, scale = adaLN(silu(vec)).chunk(2)
  float* mod_i = mod_all + F.mod_off_final;
  add("final.ln", [=](cudaStream_t st) { launch_modln(Xi, Ni, C, mod_i, mod_i + C, sA, st); });
  {
    GemmArgs a;
    a.A0 = sA; a.lda0 = C; a.K0 = C; a.Wt = e->W(4, "final_layer.linear.weight").b(); a.M = Ni; a.N = F.in_dim;
    a.bias = e->W(4, "final_layer.linear.bias").f(); a.out_f32 = io.out; a.ldo = F.in_dim;
    gemm("final.linear", a);
  }
  return prog.release();
}

void flux_forward(ldn_engine* e, const float* img, const float* ctx, const float* pe, const float* t, const float* guidance,
                  const float* y, float* out, int B, int n_img, int n_txt, cudaStream_t stream) {
  if (!e->finalized[4]) flux_finalize(e, stream);
  auto& F = *e->flux;
  LDN_CHECK(!F.guidance || guidance, "ldn_flux_forward: this model is guidance-distilled and needs a guidance value");
  LDN_CHECK(n_img > 0 && n_txt > 0 && B > 0, "ldn_flux_forward: empty input");
  LDN_CHECK(n_txt % 8 == 0, "ldn_flux_forward: the text length must be a multiple of 8 (16-byte aligned v^T column ranges)");
  auto key = std::make_tuple(n_img, n_txt);
  auto it = F.programs.find(key);
  if (it == F.programs.end()) {
    auto& io = F.io[key];
    it = F.programs.emplace(key, std::unique_ptr<Program>(build_flux_program(e, n_img, n_txt, io))).first;
  }
  Program& P = *it->second;
  const auto& io = F.io[key];
  const int N = n_img + n_txt;
  LDN_CUDA(cudaMemcpyAsync(io.pe, pe, (size_t)N * 128 * sizeof(float), cudaMemcpyDeviceToDevice, stream));
  for (int b = 0; b < B; ++b) {
    LDN_CUDA(cudaMemcpyAsync(io.img, img + (size_t)b * n_img * F.in_dim, (size_t)n_img * F.in_dim * sizeof(float),
                             cudaMemcpyDeviceToDevice, stream));
    LDN_CUDA(cudaMemcpyAsync(io.ctx, ctx + (size_t)b * n_txt * F.ctx_dim, (size_t)n_txt * F.ctx_dim * sizeof(float),
                             cudaMemcpyDeviceToDevice, stream));
    LDN_CUDA(cudaMemcpyAsync(io.t, t + b, sizeof(float), cudaMemcpyDeviceToDevice, stream));
    if (F.guidance) LDN_CUDA(cudaMemcpyAsync(io.g, guidance + b, sizeof(float), cudaMemcpyDeviceToDevice, stream));
    LDN_CUDA(cudaMemcpyAsync(io.y, y + (size_t)b * F.vec_dim, (size_t)F.vec_dim * sizeof(float), cudaMemcpyDeviceToDevice, stream));
    run_program(P, e->cfg.use_graph != 0, stream);
    LDN_CUDA(cudaMemcpyAsync(out + (size_t)b * n_img * F.in_dim, io.out, (size_t)n_img * F.in_dim * sizeof(float),
                             cudaMemcpyDeviceToDevice, stream));
  }
}

}  // namespace ldn
