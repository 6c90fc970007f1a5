// Engine lifecycle, weight ingest and program execution (C ABI side: ldn_create / ldn_load_weights / ...).
#include "engine.h"

#include <cmath>
#include <cstring>
#include <cstdlib>

namespace ldn {

void* Arena::alloc(size_t bytes, bool zero) {
  if (bytes == 0) bytes = 16;
  bytes = (bytes + 255) & ~size_t(255);
  void* p = nullptr;
  LDN_CUDA(cudaMalloc(&p, bytes));
  if (zero) LDN_CUDA(cudaMemset(p, 0, bytes));
  blocks.push_back(p);
  sizes.push_back(bytes);
  total += bytes;
  return p;
}
void Arena::free_block(void* p) {
  for (size_t i = 0; i < blocks.size(); ++i)
    if (blocks[i] == p) {
      cudaFree(p);  // implicit device synchronisation: no kernel can still be reading the tensor
      total -= sizes[i];
      blocks.erase(blocks.begin() + i);
      sizes.erase(sizes.begin() + i);
      return;
    }
}
void Arena::release() {
  for (void* p : blocks) cudaFree(p);
  blocks.clear();
  sizes.clear();
  total = 0;
}

__global__ void checksum_kernel(const uint32_t* __restrict__ p, size_t nwords, unsigned long long* out) {
  unsigned long long acc = 0;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nwords; i += (size_t)gridDim.x * blockDim.x)
    acc += (unsigned long long)p[i] * (unsigned long long)((i % 1000003) + 1);
  atomicAdd(out, acc);
}

// Debug aid (LDN_DEBUG_HASH=1): order-independent checksum of every activation buffer after each step, so two runs
// on identical inputs can be diffed to find the first non-deterministic step.
static void debug_hash(Program& prog, size_t step, cudaStream_t stream) {
  static unsigned long long* d = nullptr;
  if (!d) cudaMalloc(&d, 8);
  cudaMemsetAsync(d, 0, 8, stream);
  for (size_t b = 0; b < prog.arena->blocks.size(); ++b)
    checksum_kernel<<<256, 256, 0, stream>>>((const uint32_t*)prog.arena->blocks[b], prog.arena->sizes[b] / 4, d);
  unsigned long long h = 0;
  cudaMemcpyAsync(&h, d, 8, cudaMemcpyDeviceToHost, stream);
  cudaStreamSynchronize(stream);
  printf("LDNHASH %zu %s %016llx\n", step, prog.names[step].c_str(), h);
}

void run_program(Program& prog, bool use_graph, cudaStream_t stream) {
  if (!prog.warmed || !use_graph) {
    static const bool debug_sync = getenv("LDN_DEBUG_SYNC") != nullptr;
    static const bool debug_hash_on = getenv("LDN_DEBUG_HASH") != nullptr;
    static const bool profile_on = getenv("LDN_PROFILE") != nullptr;
    if (profile_on && prog.warmed) {
      // Per-step device timing (eager mode only): one CUDA event pair per step, printed as "LDNPROF idx ms name".
      std::vector<cudaEvent_t> ev(prog.steps.size() + 1);
      for (auto& e : ev) cudaEventCreate(&e);
      cudaEventRecord(ev[0], stream);
      for (size_t i = 0; i < prog.steps.size(); ++i) {
        prog.steps[i](stream);
        cudaEventRecord(ev[i + 1], stream);
      }
      LDN_CUDA(cudaStreamSynchronize(stream));
      for (size_t i = 0; i < prog.steps.size(); ++i) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, ev[i], ev[i + 1]);
        printf("LDNPROF %zu %.4f %s\n", i, ms, prog.names[i].c_str());
      }
      fflush(stdout);
      for (auto& e : ev) cudaEventDestroy(e);
      return;
    }
    for (size_t i = 0; i < prog.steps.size(); ++i) {
      prog.steps[i](stream);
      if (debug_hash_on && prog.arena) debug_hash(prog, i, stream);
      if (debug_sync) {
        cudaError_t e = cudaStreamSynchronize(stream);
        if (e != cudaSuccess)
          throw Error("step " + std::to_string(i) + " (" + prog.names[i] + ") failed: " + cudaGetErrorString(e));
      }
    }
    prog.warmed = true;
    return;
  }
  if (!prog.graph) {
    // Capture on a private stream so the caller's stream state is untouched, then launch on the caller's stream.
    cudaStream_t cs;
    LDN_CUDA(cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking));
    cudaGraph_t g = nullptr;
    LDN_CUDA(cudaStreamSynchronize(stream));
    LDN_CUDA(cudaStreamBeginCapture(cs, cudaStreamCaptureModeThreadLocal));
    try {
      for (auto& s : prog.steps) s(cs);
    } catch (...) {
      cudaStreamEndCapture(cs, &g);
      if (g) cudaGraphDestroy(g);
      cudaStreamDestroy(cs);
      throw;
    }
    LDN_CUDA(cudaStreamEndCapture(cs, &g));
    LDN_CUDA(cudaGraphInstantiate(&prog.graph, g, 0));
    cudaGraphDestroy(g);
    cudaStreamDestroy(cs);
  }
  LDN_CUDA(cudaGraphLaunch(prog.graph, stream));
}

}  // namespace ldn

using namespace ldn;

ldn_engine::ldn_engine() {}
ldn_engine::~ldn_engine() {}

const DevTensor& ldn_engine::W(int which, const std::string& name) const {
  auto it = w[which].find(name);
  LDN_CHECK(it != w[which].end(), "missing weight: " + name);
  return it->second;
}

#define LDN_API_BEGIN try {
#define LDN_API_END                       \
  }                                       \
  catch (const std::exception& e) {       \
    ldn::set_last_error(e.what());        \
    return 1;                             \
  }                                       \
  catch (...) {                           \
    ldn::set_last_error("unknown error"); \
    return 2;                             \
  }                                       \
  return 0;

extern "C" {

int ldn_create(const ldn_config* cfg, ldn_handle* out) {
  LDN_API_BEGIN
  LDN_CHECK(cfg && out, "ldn_create: null argument");
  int ndev = 0;
  LDN_CUDA(cudaGetDeviceCount(&ndev));
  LDN_CHECK(ndev > 0, "ldn_create: no CUDA device (this engine has no CPU fallback)");
  int dev = 0;
  LDN_CUDA(cudaGetDevice(&dev));
  cudaDeviceProp prop;
  LDN_CUDA(cudaGetDeviceProperties(&prop, dev));
  LDN_CHECK(prop.major == 10, std::string("ldn_create: needs an sm_100 (B200) device, found sm_") +
                                  std::to_string(prop.major) + std::to_string(prop.minor));
  ldn_engine* e = new ldn_engine();
  e->cfg = *cfg;
  *out = e;
  LDN_API_END
}

void ldn_destroy(ldn_handle h) {
  if (h) delete h;
}

int ldn_load_weights(ldn_handle h, int which, const ldn_tensor* tensors, int n, void* stream_) {
  LDN_API_BEGIN
  LDN_CHECK(h && tensors && which >= 0 && which < 6, "ldn_load_weights: bad argument");
  cudaStream_t stream = (cudaStream_t)stream_;
  for (int i = 0; i < n; ++i) {
    const ldn_tensor& t = tensors[i];
    LDN_CHECK(t.name && t.data && t.ndim >= 1 && t.ndim <= 4, "ldn_load_weights: malformed tensor");
    std::string name(t.name);
    DevTensor d;
    size_t numel = 1;
    for (int k = 0; k < t.ndim; ++k) numel *= (size_t)t.shape[k];
    const bool keep_f32 = t.ndim == 1 || name.find("embedding") != std::string::npos ||
                          name.find("relative_attention_bias") != std::string::npos;  // T5 logit-bias table [32, heads]
    // Re-loading a name (another checkpoint / LoRA into the same engine) reuses the old allocation when the size is
    // unchanged and frees it otherwise: nothing accumulates in the weight arena across reloads.
    void* reuse = nullptr;
    {
      auto old = h->w[which].find(name);
      if (old != h->w[which].end()) {
        const size_t old_bytes = old->second.numel() * (old->second.is_bf16 ? sizeof(bf16) : sizeof(float));
        const size_t new_bytes = numel * (keep_f32 ? sizeof(float) : sizeof(bf16));
        if (old_bytes == new_bytes) {
          LDN_CUDA(cudaStreamSynchronize(stream));
          reuse = old->second.p;
        } else {
          h->weights_arena.free_block(old->second.p);
        }
        h->w[which].erase(old);
      }
    }
    auto walloc = [&](size_t bytes) { return reuse ? reuse : h->weights_arena.alloc(bytes); };
    if (keep_f32) {
      d.is_bf16 = false;
      d.shape.assign(t.shape, t.shape + t.ndim);
      d.p = walloc(numel * sizeof(float));
      launch_convert_to_f32(t.data, t.dtype, numel, d.f(), stream);
    } else if (t.ndim == 4 && t.shape[2] * t.shape[3] > 1) {
      // OIHW -> [O, kh*kw*I], K index = (ky*kw + kx)*I + c  (K-major operand of the implicit GEMM)
      d.is_bf16 = true;
      d.shape = {t.shape[0], t.shape[2] * t.shape[3] * t.shape[1]};
      d.p = walloc(numel * sizeof(bf16));
      launch_repack_conv_weight(t.data, t.dtype, (int)t.shape[0], (int)t.shape[1], (int)t.shape[2], (int)t.shape[3],
                                d.b(), stream);
    } else {
      d.is_bf16 = true;
      if (t.ndim == 4)
        d.shape = {t.shape[0], t.shape[1]};
      else
        d.shape.assign(t.shape, t.shape + t.ndim);
      d.p = walloc(numel * sizeof(bf16));
      launch_convert_to_bf16(t.data, t.dtype, numel, d.b(), stream);
    }
    h->w[which][name] = d;
  }
  h->finalized[which] = false;
  LDN_CUDA(cudaStreamSynchronize(stream));
  LDN_API_END
}

int ldn_set_sigmas(ldn_handle h, const float* sigmas_host, const float* log_sigmas_host, int n) {
  LDN_API_BEGIN
  LDN_CHECK(h && sigmas_host && log_sigmas_host && n > 0, "ldn_set_sigmas: bad argument");
  h->log_sigmas = h->weights_arena.get<float>(n);
  h->n_sigmas = n;
  LDN_CUDA(cudaMemcpy(h->log_sigmas, log_sigmas_host, sizeof(float) * n, cudaMemcpyHostToDevice));
  LDN_API_END
}

int ldn_set_context(ldn_handle h, const float* ctx, int rows, int tokens, void* stream) {
  LDN_API_BEGIN
  LDN_CHECK(h && ctx, "ldn_set_context: bad argument");
  if (!h->finalized[0]) unet_finalize(h, (cudaStream_t)stream);
  unet_set_context(h, ctx, rows, tokens, (cudaStream_t)stream);
  LDN_API_END
}

int ldn_unet_denoise(ldn_handle h, const float* x, const float* sigma, float* out, int rows, int lat_h, int lat_w,
                     void* stream) {
  LDN_API_BEGIN
  LDN_CHECK(h && x && sigma && out, "ldn_unet_denoise: bad argument");
  if (!h->finalized[0]) unet_finalize(h, (cudaStream_t)stream);
  unet_denoise(h, x, sigma, out, rows, lat_h, lat_w, (cudaStream_t)stream);
  LDN_API_END
}

int ldn_unet_last_launches(ldn_handle h) { return h ? ldn::unet_last_launches(h) : 0; }

int ldn_cfg_step(const float* x, const float* den_uncond, const float* den_cond, float cfg, int mode, float c0,
                 float c1, float c2, const float* noise, float* x_out, float* denoised_out, int64_t n, void* stream) {
  LDN_API_BEGIN
  LDN_CHECK(den_uncond && den_cond, "ldn_cfg_step: bad argument");
  launch_cfg_step(x, den_uncond, den_cond, cfg, mode, c0, c1, c2, noise, x_out, denoised_out, (size_t)n,
                  (cudaStream_t)stream);
  LDN_API_END
}

int ldn_resample_bilinear(const float* src, float* dst, int planes, int h, int w, int oh, int ow, void* stream) {
  LDN_API_BEGIN
  LDN_CHECK(src && dst && planes >= 0 && h > 0 && w > 0 && oh > 0 && ow > 0, "ldn_resample_bilinear: bad argument");
  launch_resample_bilinear(src, dst, planes, h, w, oh, ow, (cudaStream_t)stream);
  LDN_API_END
}

int ldn_bislerp(const float* src, float* tmp, float* dst, int n, int c, int h, int w, int oh, int ow, void* stream) {
  LDN_API_BEGIN
  LDN_CHECK(src && tmp && dst && n >= 0 && c > 0 && h > 0 && w > 0 && oh > 0 && ow > 0, "ldn_bislerp: bad argument");
  if (n > 0) launch_bislerp(src, tmp, dst, n, c, h, w, oh, ow, (cudaStream_t)stream);
  LDN_API_END
}

int ldn_vae_decode(ldn_handle h, const float* z, float* rgb, int B, int lat_h, int lat_w, void* stream) {
  LDN_API_BEGIN
  LDN_CHECK(h && z && rgb, "ldn_vae_decode: bad argument");
  if (!h->finalized[1]) vae_finalize(h, (cudaStream_t)stream);
  vae_decode(h, z, rgb, B, lat_h, lat_w, (cudaStream_t)stream);
  LDN_API_END
}

int ldn_vae_encode(ldn_handle h, const float* pixels, float* moments, int B, int H, int W, void* stream) {
  LDN_API_BEGIN
  LDN_CHECK(h && pixels && moments, "ldn_vae_encode: bad argument");
  if (!h->finalized[1]) vae_finalize(h, (cudaStream_t)stream);
  vae_encode(h, pixels, moments, B, H, W, (cudaStream_t)stream);
  LDN_API_END
}

int ldn_flux_forward(ldn_handle h, const float* img, const float* ctx, const float* pe, const float* t,
                     const float* guidance, const float* y, float* out, int B, int n_img, int n_txt, void* stream) {
  LDN_API_BEGIN
  LDN_CHECK(h && img && ctx && pe && t && y && out, "ldn_flux_forward: bad argument");
  flux_forward(h, img, ctx, pe, t, guidance, y, out, B, n_img, n_txt, (cudaStream_t)stream);
  LDN_API_END
}

int ldn_taesd_decode(ldn_handle h, const float* z, float* rgb, int B, int lat_h, int lat_w, void* stream) {
  LDN_API_BEGIN
  LDN_CHECK(h && z && rgb, "ldn_taesd_decode: bad argument");
  taesd_decode(h, z, rgb, B, lat_h, lat_w, (cudaStream_t)stream);
  LDN_API_END
}

int ldn_clip_encode(ldn_handle h, const int64_t* ids, int S, float* out_penultimate, float* out_last, void* stream) {
  LDN_API_BEGIN
  LDN_CHECK(h && ids, "ldn_clip_encode: bad argument");
  if (!h->finalized[2]) clip_finalize(h, (cudaStream_t)stream);
  clip_encode(h, ids, S, out_penultimate, out_last, (cudaStream_t)stream);
  LDN_API_END
}

int ldn_clip_set_extra_embeddings(ldn_handle h, const float* vectors, int n, void* stream) {
  LDN_API_BEGIN
  LDN_CHECK(h && n >= 0 && (n == 0 || vectors), "ldn_clip_set_extra_embeddings: bad argument");
  LDN_CHECK(!h->w[2].empty(), "ldn_clip_set_extra_embeddings: CLIP weights not loaded");
  LDN_CHECK(n <= ldn_engine::kClipExtraCap, "ldn_clip_set_extra_embeddings: more than 256 textual-inversion vectors");
  if (!h->finalized[2]) clip_finalize(h, (cudaStream_t)stream);
  const int width = (int)h->W(2, "embeddings.token_embedding.weight").shape[1];
  if (n > 0)
    LDN_CUDA(cudaMemcpyAsync(h->clip_extra, vectors, (size_t)n * width * sizeof(float), cudaMemcpyDeviceToDevice,
                             (cudaStream_t)stream));
  LDN_CUDA(cudaMemcpyAsync(h->clip_extra_n, &n, sizeof(int), cudaMemcpyHostToDevice, (cudaStream_t)stream));
  LDN_CUDA(cudaStreamSynchronize((cudaStream_t)stream));  // `n` lives on this stack frame
  LDN_API_END
}

int ldn_t5_encode(ldn_handle h, const int64_t* ids, const int32_t* rel_buckets, int S, int n, float* out, void* stream) {
  LDN_API_BEGIN
  LDN_CHECK(h && ids && rel_buckets && out, "ldn_t5_encode: bad argument");
  LDN_CHECK(!h->w[5].empty(), "ldn_t5_encode: T5 weights not loaded");
  t5_encode(h, ids, rel_buckets, S, n, out, (cudaStream_t)stream);
  LDN_API_END
}

int ldn_groupnorm_bf16(const void* x0, int C0, const void* x1, int C1, int B, int HW, int groups, float eps,
                       const float* gamma, const float* beta, int silu, void* out, void* stream) {
  LDN_API_BEGIN
  static thread_local float* ws = nullptr;
  static thread_local int ws_b = 0;
  if (ws_b < B) {
    if (ws) cudaFree(ws);
    LDN_CUDA(cudaMalloc(&ws, groupnorm_ws_bytes(B)));
    LDN_CUDA(cudaMemset(ws, 0, groupnorm_ws_bytes(B)));
    ws_b = B;
  }
  LDN_CUDA(cudaMemsetAsync(ws, 0, (size_t)B * 64 * sizeof(unsigned long long), (cudaStream_t)stream));  // statistics slot 0
  launch_groupnorm((const bf16*)x0, C0, (const bf16*)x1, C1, B, HW, groups, eps, gamma, beta, silu != 0, (bf16*)out, ws, 0,
                   (cudaStream_t)stream);
  LDN_API_END
}

int ldn_conv3x3_groupnorm_bf16(const void* x, const void* Wt, int B, int H, int W, int Cin, int Cout, const float* bias,
                               const float* rowbias, int ld_rowbias, const void* residual, float eps, const float* gamma,
                               const float* beta, int silu, void* conv_out, void* gn_out, int* fused, void* stream) {
  LDN_API_BEGIN
  LDN_CHECK(x && Wt && gamma && beta && conv_out && gn_out && B > 0 && H > 0 && W > 0, "ldn_conv3x3_groupnorm_bf16: bad argument");
  static thread_local float* ws = nullptr;
  static thread_local int ws_b = 0;
  static thread_local float* sk = nullptr;
  const size_t sk_bytes = (size_t)64 << 20;
  if (ws_b < B) {
    if (ws) cudaFree(ws);
    LDN_CUDA(cudaMalloc(&ws, groupnorm_ws_bytes(B)));
    ws_b = B;
  }
  if (!sk) LDN_CUDA(cudaMalloc(&sk, sk_bytes));
  LDN_CUDA(cudaMemsetAsync(ws, 0, (size_t)B * 64 * sizeof(unsigned long long), (cudaStream_t)stream));  // statistics slot 0
  GemmArgs a;
  a.conv = true;
  a.A0 = (const bf16*)x; a.B = B; a.H = H; a.W = W; a.Cin = Cin;
  a.Wt = (const bf16*)Wt; a.N = Cout; a.M = B * H * W;
  a.bias = bias; a.rowbias = rowbias; a.ld_rowbias = ld_rowbias;
  a.residual = (const bf16*)residual; a.ldr = Cout;
  a.out = (bf16*)conv_out; a.ldo = Cout;
  a.splitk_ws = sk; a.splitk_ws_bytes = sk_bytes;
  a.gn_acc = groupnorm_slot(ws, 0, B);
  GemmPlan plan = make_gemm_plan(a);
  launch_gemm(plan, (cudaStream_t)stream);
  const bool have = plan.gn_cpg > 0;
  if (fused) *fused = have ? 1 : 0;
  launch_groupnorm((const bf16*)conv_out, Cout, nullptr, 0, B, H * W, 32, eps, gamma, beta, silu != 0, (bf16*)gn_out, ws, 0,
                   (cudaStream_t)stream, have);
  LDN_API_END
}

int ldn_layernorm_bf16(const void* x, int rows, int C, float eps, const float* gamma, const float* beta, void* out,
                       void* stream) {
  LDN_API_BEGIN
  launch_layernorm((const bf16*)x, rows, C, eps, gamma, beta, (bf16*)out, (cudaStream_t)stream);
  LDN_API_END
}

}  // extern "C"
