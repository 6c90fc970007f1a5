// CLIP-L text encoder (12 layers, width 768, 12 heads x 64, quick_gelu MLP, causal mask) as a launch program.
//
// Mirrors (structure, not code): CLIPTextModel_.forward   src/clip/CLIPTextModel.py:51-107
//   CLIPEncoder / CLIPLayer / CLIPAttention / CLIPMLP      src/clip/Clip.py:14-180
//   embeddings (token + position)                          src/clip/Clip.py:182-294
// Output: final_layer_norm of the penultimate layer's hidden state (what SD1.5 conditions on: CLIPSetLastLayer(-2),
// src/clip/Clip.py:592-608, SDClipModel.forward src/SD15/SDClip.py:269-336) and of the last layer's.
//
// The reference computes CLIP in fp32 from fp16 weights; here GEMMs take bf16 operands with fp32 accumulation (the output
// is consumed as a bf16 cross-attention context anyway). Q and K projections are one GEMM written into per-head slots,
// V is produced transposed by swapping GEMM operands with its bias folded into out_proj's bias
// (softmax rows sum to 1), attention is the causal tcgen05 kernel.
#include <cmath>
#include <map>

#include "engine.h"

using namespace ldn;

struct ldn_engine::ClipState {
  int layers = 12, width = 768, heads = 12, mlp = 3072, T = 77;
  Arena arena;
  std::vector<bf16*> Wqk;        // [2W, W]
  std::vector<float*> bqk;       // [2W]
  std::vector<float*> out_bias;  // Wo bv + bo
  std::map<int, std::unique_ptr<Program>> programs;
  std::map<int, long long*> in_ids;
  std::map<int, float*> out_pen, out_last;
  std::vector<std::unique_ptr<Arena>> program_arenas;
};

namespace ldn {

void clip_finalize(ldn_engine* e, cudaStream_t stream) {
  LDN_CHECK(!e->w[2].empty(), "CLIP weights not loaded");
  e->clip.reset(new ldn_engine::ClipState());
  auto& C = *e->clip;
  const int W = C.width;
  for (int i = 0; i < C.layers; ++i) {
    const std::string p = "encoder.layers." + std::to_string(i) + ".self_attn";
    bf16* wqk = C.arena.get<bf16>((size_t)2 * W * W);
    float* bqk = C.arena.get<float>(2 * W);
    LDN_CUDA(cudaMemcpyAsync(wqk, e->W(2, p + ".q_proj.weight").p, (size_t)W * W * sizeof(bf16), cudaMemcpyDeviceToDevice, stream));
    LDN_CUDA(cudaMemcpyAsync(wqk + (size_t)W * W, e->W(2, p + ".k_proj.weight").p, (size_t)W * W * sizeof(bf16),
                             cudaMemcpyDeviceToDevice, stream));
    LDN_CUDA(cudaMemcpyAsync(bqk, e->W(2, p + ".q_proj.bias").p, W * sizeof(float), cudaMemcpyDeviceToDevice, stream));
    LDN_CUDA(cudaMemcpyAsync(bqk + W, e->W(2, p + ".k_proj.bias").p, W * sizeof(float), cudaMemcpyDeviceToDevice, stream));
    float* ob = C.arena.get<float>(W);
    launch_small_linear(e->W(2, p + ".v_proj.bias").f(), 1, W, e->W(2, p + ".out_proj.weight").b(),
                        e->W(2, p + ".out_proj.bias").f(), W, false, false, ob, stream);
    C.Wqk.push_back(wqk);
    C.bqk.push_back(bqk);
    C.out_bias.push_back(ob);
  }
  if (!e->clip_extra) {
    e->clip_extra = e->weights_arena.get<float>((size_t)ldn_engine::kClipExtraCap * W, true);
    e->clip_extra_n = reinterpret_cast<int*>(e->weights_arena.alloc(sizeof(int), true));
  }
  LDN_CUDA(cudaStreamSynchronize(stream));
  e->finalized[2] = true;
}

static Program* build_clip_program(ldn_engine* e, int S) {
  auto& C = *e->clip;
  std::unique_ptr<Program> prog(new Program());
  C.program_arenas.emplace_back(new Arena());
  Arena& A = *C.program_arenas.back();
  prog->arena = &A;
  const int W = C.width, T = C.T, H = C.heads, d = W / H, M = S * T;
  const int slot = 64;
  LDN_CHECK(d == 64, "CLIP head dim must be 64");
  const int Mld = (M + 15) / 16 * 16;
  const int nk_pad = (T + 7) / 8 * 8;  // V^T columns per sequence (16-byte aligned TMA box starts)
  long long* ids = A.get<long long>(M);
  bf16* X = A.get<bf16>((size_t)M * W);
  bf16* N1 = A.get<bf16>((size_t)M * W);
  bf16* QK = A.get<bf16>((size_t)M * 2 * H * slot, true);
  bf16* Vt = A.get<bf16>((size_t)W * Mld, true);
  bf16* VtP = A.get<bf16>((size_t)W * S * nk_pad, true);
  bf16* O = A.get<bf16>((size_t)M * W);
  bf16* F1 = A.get<bf16>((size_t)M * C.mlp);
  bf16* Xpen = A.get<bf16>((size_t)M * W);
  float* out_pen = A.get<float>((size_t)M * W);
  float* out_last = A.get<float>((size_t)M * W);
  C.in_ids[S] = ids;
  C.out_pen[S] = out_pen;
  C.out_last[S] = out_last;

  auto add = [&](const std::string& name, Step s) {
    prog->steps.push_back(std::move(s));
    prog->names.push_back(name);
    prog->launches += 1;
  };
  auto gemm = [&](const std::string& name, const GemmArgs& a) {
    GemmPlan plan = make_gemm_plan(a);
    add(name, [plan](cudaStream_t st) { launch_gemm(plan, st); });
  };
  {
    const DevTensor& tokw = e->W(2, "embeddings.token_embedding.weight");
    const float* tok = tokw.f();
    const int vocab = (int)tokw.shape[0];
    const float* pos = e->W(2, "embeddings.position_embedding.weight").f();
    const float* extra = e->clip_extra;
    const int* extra_n = e->clip_extra_n;
    add("embeddings", [=](cudaStream_t st) { launch_clip_embed(ids, tok, vocab, extra, extra_n, pos, M, T, W, X, st); });
  }
  const float scale = 1.0f / sqrtf((float)d);
  for (int i = 0; i < C.layers; ++i) {
    const std::string p = "encoder.layers." + std::to_string(i);
    {
      const float* g = e->W(2, p + ".layer_norm1.weight").f();
      const float* b = e->W(2, p + ".layer_norm1.bias").f();
      add(p + ".ln1", [=](cudaStream_t st) { launch_layernorm(X, M, W, 1e-5f, g, b, N1, st); });
    }
    GemmArgs qk;
    qk.A0 = N1; qk.lda0 = W; qk.K0 = W; qk.Wt = C.Wqk[i]; qk.M = M; qk.N = 2 * W; qk.bias = C.bqk[i];
    qk.out = QK; qk.ldo = 2LL * H * slot; qk.head_dim = d; qk.head_slot = slot;
    gemm(p + ".qk", qk);
    GemmArgs v;
    v.A0 = e->W(2, p + ".self_attn.v_proj.weight").b(); v.lda0 = W; v.K0 = W; v.Wt = N1; v.M = W; v.N = Mld; v.wt_rows = M;
    v.out = Vt; v.ldo = Mld;
    gemm(p + ".vt", v);
    // re-lay V^T with 80-column sequences (77 is not a multiple of 8)
    add(p + ".vt_pad", [=](cudaStream_t st) { launch_pad_vt_cols(Vt, Mld, W, S, T, nk_pad, VtP, st); });
    AttnArgs at;
    at.Q = QK; at.ldq = 2LL * H * slot; at.K = QK + (size_t)H * slot; at.ldk = 2LL * H * slot;
    at.Vt = VtP; at.ldvt = (long long)S * nk_pad; at.vt_rows = W;
    at.B = S; at.heads = H; at.Nq = T; at.Nk = T; at.nk_pad = nk_pad; at.kv_batch_stride = T; at.d = d; at.slot = slot;
    at.causal = 1; at.scale = scale; at.out = O; at.ldo = W;
    AttnPlan ap = make_attn_plan(at);
    add(p + ".sdpa", [ap](cudaStream_t st) { launch_attn(ap, st); });
    GemmArgs o;
    o.A0 = O; o.lda0 = W; o.K0 = W; o.Wt = e->W(2, p + ".self_attn.out_proj.weight").b(); o.M = M; o.N = W;
    o.bias = C.out_bias[i]; o.residual = X; o.ldr = W; o.out = X; o.ldo = W;
    gemm(p + ".out_proj", o);
    {
      const float* g = e->W(2, p + ".layer_norm2.weight").f();
      const float* b = e->W(2, p + ".layer_norm2.bias").f();
      add(p + ".ln2", [=](cudaStream_t st) { launch_layernorm(X, M, W, 1e-5f, g, b, N1, st); });
    }
    GemmArgs f1;
    f1.A0 = N1; f1.lda0 = W; f1.K0 = W; f1.Wt = e->W(2, p + ".mlp.fc1.weight").b(); f1.M = M; f1.N = C.mlp;
    f1.bias = e->W(2, p + ".mlp.fc1.bias").f(); f1.act = 1; f1.out = F1; f1.ldo = C.mlp;
    gemm(p + ".fc1", f1);
    GemmArgs f2;
    f2.A0 = F1; f2.lda0 = C.mlp; f2.K0 = C.mlp; f2.Wt = e->W(2, p + ".mlp.fc2.weight").b(); f2.M = M; f2.N = W;
    f2.bias = e->W(2, p + ".mlp.fc2.bias").f(); f2.residual = X; f2.ldr = W; f2.out = X; f2.ldo = W;
    gemm(p + ".fc2", f2);
    if (i == C.layers - 2) {
      add(p + ".keep_penultimate", [=](cudaStream_t st) {
        LDN_CUDA(cudaMemcpyAsync(Xpen, X, (size_t)M * W * sizeof(bf16), cudaMemcpyDeviceToDevice, st));
      });
    }
  }
  {
    const float* g = e->W(2, "final_layer_norm.weight").f();
    const float* b = e->W(2, "final_layer_norm.bias").f();
    add("final_ln.penultimate", [=](cudaStream_t st) { launch_layernorm(Xpen, M, W, 1e-5f, g, b, nullptr, st, out_pen); });
    add("final_ln.last", [=](cudaStream_t st) { launch_layernorm(X, M, W, 1e-5f, g, b, nullptr, st, out_last); });
  }
  return prog.release();
}

void clip_encode(ldn_engine* e, const int64_t* ids, int S, float* out_pen, float* out_last, cudaStream_t stream) {
  auto& C = *e->clip;
  auto it = C.programs.find(S);
  if (it == C.programs.end()) it = C.programs.emplace(S, std::unique_ptr<Program>(build_clip_program(e, S))).first;
  Program& P = *it->second;
  const size_t M = (size_t)S * C.T;
  LDN_CUDA(cudaMemcpyAsync(C.in_ids[S], ids, M * sizeof(long long), cudaMemcpyDeviceToDevice, stream));
  run_program(P, e->cfg.use_graph != 0, stream);
  if (out_pen)
    LDN_CUDA(cudaMemcpyAsync(out_pen, C.out_pen[S], M * C.width * sizeof(float), cudaMemcpyDeviceToDevice, stream));
  if (out_last)
    LDN_CUDA(cudaMemcpyAsync(out_last, C.out_last[S], M * C.width * sizeof(float), cudaMemcpyDeviceToDevice, stream));
}

}  // namespace ldn
