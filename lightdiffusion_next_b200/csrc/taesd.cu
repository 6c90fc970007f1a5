// TAESD preview decoder (Tiny AutoEncoder for SD) as a launch program over NHWC bf16 activations.
//
// Mirrors (structure, not code): Decoder2 / Block / Clamp, src/AutoEncoders/taesd.py:25-136 -- the network the reference's
// sampler loops run every 5 steps on the current latent for on-screen previews (samplers.py:959-960, taesd_preview
// taesd.py:219-255). State-dict keys are the nn.Sequential indices of `taesd_decoder.safetensors`:
//   1: conv 4->64   3,4,5: Block   7: conv (no bias)   8,9,10: Block   12: conv   13,14,15: Block   17: conv   18: Block
//   19: conv 64->3   (0 = Clamp, 2 = ReLU, 6/11/16 = nearest 2x Upsample)
// Everything maps onto kernels the UNet / VAE already use: the tiny-Cin first conv with the Clamp (tanh(x/3)*3) and the ReLU
// fused, implicit-GEMM conv3x3 on tcgen05 with ReLU before / after the residual in the epilogue, nearest 2x upsample, and
// the 3-channel head conv as a 16-column tensor-core tile.
#include <map>

#include "engine.h"

using namespace ldn;

struct ldn_engine::TaesdState {
  std::map<std::tuple<int, int, int>, std::unique_ptr<Program>> programs;
  std::vector<std::unique_ptr<Arena>> arenas;
};

namespace ldn {

static Program* build_taesd_program(ldn_engine* e, int B, int h, int w) {
  auto& T = *e->taesd;
  std::unique_ptr<Program> prog(new Program());
  T.arenas.emplace_back(new Arena());
  Arena& A = *T.arenas.back();
  prog->arena = &A;
  Program& P = *prog;
  const int C = 64;
  auto add = [&](const std::string& name, Step s) {
    P.steps.push_back(std::move(s));
    P.names.push_back(name);
    P.launches += 1;
  };
  auto conv = [&](const std::string& key, const bf16* x, int H, int W, int act, const bf16* residual, bf16* out) {
    GemmArgs a;
    a.conv = true; a.A0 = x; a.B = B; a.H = H; a.W = W; a.Cin = C;
    a.Wt = e->W(3, key + ".weight").b(); a.N = C;
    a.bias = e->has(3, key + ".bias") ? e->W(3, key + ".bias").f() : nullptr;
    a.act = act; a.residual = residual; a.ldr = C; a.out = out; a.ldo = C;
    GemmPlan plan = make_gemm_plan(a);
    add(key, [plan](cudaStream_t st) { launch_gemm(plan, st); });
  };
  const size_t max_act = (size_t)B * (8 * h) * (8 * w) * C;
  bf16* t0 = A.get<bf16>(max_act);
  bf16* t1 = A.get<bf16>(max_act);
  bf16* xa = A.get<bf16>(max_act);
  bf16* xb = A.get<bf16>(max_act);
  prog->io_elems = (size_t)B * 4 * h * w;
  prog->in_x = A.get<float>(prog->io_elems);
  prog->out = A.get<float>((size_t)B * 8 * h * 8 * w * 3);
  {
    const float* z = prog->in_x;
    const bf16* wt = e->W(3, "1.weight").b();
    const float* bias = e->W(3, "1.bias").f();
    add("1", [=](cudaStream_t st) { launch_conv_in(z, nullptr, wt, bias, B, h, w, 4, C, xa, st, /*clamp + relu*/ 3); });
  }
  bf16* x = xa;
  bf16* y = xb;
  int H = h, W = w;
  auto block = [&](int idx) {  // relu(conv(relu(conv(relu(conv(x))))) + x)
    const std::string p = std::to_string(idx) + ".conv.";
    conv(p + "0", x, H, W, 2, nullptr, t0);
    conv(p + "2", t0, H, W, 2, nullptr, t1);
    conv(p + "4", t1, H, W, 3, x, y);
    std::swap(x, y);
  };
  int idx = 3;
  for (int stage = 0; stage < 4; ++stage) {
    const int nblocks = stage == 3 ? 1 : 3;
    for (int i = 0; i < nblocks; ++i) block(idx++);
    if (stage < 3) {
      const bf16* src = x;
      const int H0 = H, W0 = W;
      add(std::to_string(idx) + ".upsample", [=](cudaStream_t st) { launch_upsample2x(src, B, H0, W0, C, t0, st); });
      ++idx;
      H *= 2;
      W *= 2;
      conv(std::to_string(idx), t0, H, W, 0, nullptr, y);
      std::swap(x, y);
      ++idx;
    }
  }
  {
    float* acc16 = A.get<float>((size_t)B * H * W * 16);
    GemmArgs a;
    a.conv = true; a.A0 = x; a.B = B; a.H = H; a.W = W; a.Cin = C;
    a.Wt = e->W(3, std::to_string(idx) + ".weight").b(); a.N = 16; a.wt_rows = 3; a.BN = 16;
    a.out_f32 = acc16; a.ldo = 16;
    GemmPlan plan = make_gemm_plan(a);
    add(std::to_string(idx), [plan](cudaStream_t st) { launch_gemm(plan, st); });
    const float* bias = e->W(3, std::to_string(idx) + ".bias").f();
    float* out = prog->out;
    const size_t npix = (size_t)B * H * W;
    add("rgb", [=](cudaStream_t st) { launch_vae_rgb_finish(acc16, bias, npix, 3, out, st, /*raw*/ 1); });
  }
  return prog.release();
}

void taesd_decode(ldn_engine* e, const float* z, float* rgb, int B, int h, int w, cudaStream_t stream) {
  LDN_CHECK(!e->w[3].empty(), "ldn_taesd_decode: TAESD decoder weights not loaded");
  if (!e->taesd) e->taesd.reset(new ldn_engine::TaesdState());
  if (!e->finalized[3]) {  // new weights: programs hold pointers into the old ones
    e->taesd->programs.clear();
    e->finalized[3] = true;
  }
  auto& T = *e->taesd;
  auto key = std::make_tuple(B, h, w);
  auto it = T.programs.find(key);
  if (it == T.programs.end()) it = T.programs.emplace(key, std::unique_ptr<Program>(build_taesd_program(e, B, h, w))).first;
  Program& P = *it->second;
  LDN_CUDA(cudaMemcpyAsync(P.in_x, z, P.io_elems * sizeof(float), cudaMemcpyDeviceToDevice, stream));
  run_program(P, e->cfg.use_graph != 0, stream);
  LDN_CUDA(cudaMemcpyAsync(rgb, P.out, (size_t)B * 8 * h * 8 * w * 3 * sizeof(float), cudaMemcpyDeviceToDevice, stream));
}

}  // namespace ldn
