// Engine state shared by the UNet / VAE / CLIP graph builders (internal).
#pragma once
#include <functional>
#include <map>
#include <memory>
#include <string>
#include <tuple>
#include <unordered_map>
#include <vector>

#include "../../include/ldn.h"
#include "common.h"

namespace ldn {

struct DevTensor {
  void* p = nullptr;
  bool is_bf16 = false;  // else fp32
  std::vector<int64_t> shape;
  size_t numel() const {
    size_t n = 1;
    for (auto s : shape) n *= (size_t)s;
    return n;
  }
  bf16* b() const { return reinterpret_cast<bf16*>(p); }
  float* f() const { return reinterpret_cast<float*>(p); }
};

typedef std::function<void(cudaStream_t)> Step;

// A fixed launch sequence over fixed buffers for one problem shape; optionally frozen into a CUDA graph.
struct Arena;
struct Program {
  std::vector<Step> steps;
  std::vector<std::string> names;  // one label per step (profiling / debugging)
  cudaGraphExec_t graph = nullptr;
  bool warmed = false;
  // static I/O staging buffers (graph-stable addresses)
  float* in_x = nullptr;
  float* in_sigma = nullptr;
  float* out = nullptr;
  size_t io_elems = 0;
  int launches = 0;  // kernels launched per run (reported as gpu_launches by bench.py)
  struct Arena* arena = nullptr;  // activation arena (debug checksums)
};

struct Arena {
  std::vector<void*> blocks;
  std::vector<size_t> sizes;
  size_t total = 0;
  void* alloc(size_t bytes, bool zero = false);
  void free_block(void* p);  // releases one block (a replaced weight tensor)
  template <typename T>
  T* get(size_t n, bool zero = false) {
    return reinterpret_cast<T*>(alloc(n * sizeof(T), zero));
  }
  void release();
  ~Arena() { release(); }
};

}  // namespace ldn

struct ldn_engine {
  ldn_config cfg;
  ldn::Arena weights_arena;
  std::unordered_map<std::string, ldn::DevTensor> w[6];  // 0 unet, 1 vae, 2 clip, 3 taesd (preview decoder), 4 flux DiT, 5 T5 text encoder
  bool finalized[6] = {false, false, false, false, false, false};
  float* log_sigmas = nullptr;
  int n_sigmas = 0;
  // textual-inversion rows of the CLIP token table (ids vocab, vocab + 1, ...): fixed capacity, graph-stable addresses
  static constexpr int kClipExtraCap = 256;
  float* clip_extra = nullptr;   // [kClipExtraCap, 768] fp32
  int* clip_extra_n = nullptr;   // device scalar: rows in use

  // ---- UNet derived weights / context (unet.cu)
  struct UNetState;
  std::shared_ptr<UNetState> unet;
  // ---- VAE / CLIP (vae.cu / clip.cu)
  struct VaeState;
  std::shared_ptr<VaeState> vae;
  struct ClipState;
  std::shared_ptr<ClipState> clip;
  struct TaesdState;
  std::shared_ptr<TaesdState> taesd;
  struct FluxState;
  std::shared_ptr<FluxState> flux;
  struct T5State;
  std::shared_ptr<T5State> t5;

  ldn_engine();
  ~ldn_engine();
  const ldn::DevTensor& W(int which, const std::string& name) const;
  bool has(int which, const std::string& name) const { return w[which].count(name) != 0; }
};

namespace ldn {
// run a program: eager on first call (sets func attributes, validates), graph-captured afterwards if enabled
void run_program(Program& prog, bool use_graph, cudaStream_t stream);

void unet_finalize(ldn_engine* e, cudaStream_t stream);
void unet_set_context(ldn_engine* e, const float* ctx, int rows, int tokens, cudaStream_t stream);
void unet_denoise(ldn_engine* e, const float* x, const float* sigma, float* out, int rows, int h, int w,
                  cudaStream_t stream);
int unet_last_launches(ldn_engine* e);
void vae_finalize(ldn_engine* e, cudaStream_t stream);
void vae_decode(ldn_engine* e, const float* z, float* rgb, int B, int h, int w, cudaStream_t stream);
void vae_encode(ldn_engine* e, const float* pixels, float* moments, int B, int H, int W, cudaStream_t stream);
void flux_forward(ldn_engine* e, const float* img, const float* ctx, const float* pe, const float* t, const float* guidance,
                  const float* y, float* out, int B, int n_img, int n_txt, cudaStream_t stream);
void taesd_decode(ldn_engine* e, const float* z, float* rgb, int B, int h, int w, cudaStream_t stream);
void clip_finalize(ldn_engine* e, cudaStream_t stream);
void clip_encode(ldn_engine* e, const int64_t* ids, int S, float* out_pen, float* out_last, cudaStream_t stream);
void t5_encode(ldn_engine* e, const int64_t* ids, const int32_t* rel_buckets, int S, int n, float* out, cudaStream_t stream);
}  // namespace ldn
