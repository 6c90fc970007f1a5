// T5 text encoder (T5-XXL for Flux: 24 blocks, width 4096, 64 heads x 64, gated tanh-GELU feed-forward of 10240, RMS norms,
// relative-position logit bias shared by all blocks, no attention mask, no 1/sqrt(d) scale) as a launch program.
//
// Mirrors (structure, not code): T5.forward / T5Stack.forward   src/clip/FluxClip.py:457-562
//   T5Block = T5LayerSelfAttention + T5LayerFF                    src/clip/FluxClip.py:58-98, 272-400
//   T5Attention (compute_bias, k pre-scaled to cancel the scale)  src/clip/FluxClip.py:101-270
//   T5LayerNorm (RMS, eps 1e-6)                                   src/clip/FluxClip.py:616-643
// Output: final_layer_norm of the last block's hidden state in fp32 (SDClipModel layer="last", src/SD15/SDClip.py:269-336).
//
// The reference runs the stack in fp32 / fp16 from (de)quantised weights; here GEMMs take bf16 operands with fp32
// accumulation and the residual stream is bf16.  q and k projections write the per-head slots the attention kernel reads,
// v is produced transposed by swapping GEMM operands, attention is the generation-1 tcgen05 kernel with the additive
// logit-bias variant (attention.cu, kBias), the bias table [heads, n, n] is expanded once per sequence length from the
// 32-bucket embedding and the host-computed bucket of every relative distance (bit-identical to the reference's own
// fp32 bucket arithmetic, which a device logf could miss by one bucket at the boundaries).
//
// STATUS (round 1): written after the round's GPU budget was spent -- compiles for sm_100a, not yet executed on a GPU.
#include <cmath>
#include <map>

#include "engine.h"

using namespace ldn;

struct ldn_engine::T5State {
  int layers = 0, width = 0, heads = 0, ff = 0, vocab = 0;
  std::map<std::pair<int, int>, std::unique_ptr<Program>> programs;  // key: (rows S, tokens n)
  std::map<std::pair<int, int>, long long*> in_ids;
  std::map<std::pair<int, int>, int*> in_buckets;
  std::map<std::pair<int, int>, float*> out;
  std::vector<std::unique_ptr<Arena>> program_arenas;
};

namespace ldn {

// ------------------------------------------------------------------------------------------------ kernels
// X[row, :] = table[ids[row], :]   (bf16 rows of W elements, W % 8 == 0; ids outside the table are clamped)
__global__ void t5_embed_kernel(const long long* __restrict__ ids, const bf16* __restrict__ table, int rows, int W, int vocab,
                                bf16* __restrict__ X) {
  const int vec = W / 8;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < (long long)rows * vec;
       i += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(i / vec), c = (int)(i % vec);
    long long id = ids[r];
    id = id < 0 ? 0 : (id >= vocab ? vocab - 1 : id);
    reinterpret_cast<uint4*>(X + (size_t)r * W)[c] = reinterpret_cast<const uint4*>(table + (size_t)id * W)[c];
  }
}

// out[row, :] = w * x[row, :] * rsqrt(mean(x^2) + eps); one CTA per row, fp32 statistics; bf16 or fp32 output
__global__ void __launch_bounds__(256) t5_rms_kernel(const bf16* __restrict__ x, const float* __restrict__ w, int W, float eps,
                                                     bf16* __restrict__ out, float* __restrict__ out_f32) {
  __shared__ float red[8];
  const int row = blockIdx.x;
  const bf16* xr = x + (size_t)row * W;
  float ss = 0.f;
  for (int c = threadIdx.x * 2; c < W; c += 512) {
    const float2 v = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(xr + c));
    ss = fmaf(v.x, v.x, fmaf(v.y, v.y, ss));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ss;
  __syncthreads();
  float tot = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) tot += red[i];  // same order in every thread: deterministic
  const float inv = rsqrtf(tot / (float)W + eps);
  for (int c = threadIdx.x * 2; c < W; c += 512) {
    const float2 v = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(xr + c));
    const float a = w[c] * (v.x * inv), b = w[c + 1] * (v.y * inv);
    if (out_f32) {
      *reinterpret_cast<float2*>(out_f32 + (size_t)row * W + c) = make_float2(a, b);
    } else {
      *reinterpret_cast<__nv_bfloat162*>(out + (size_t)row * W + c) = __floats2bfloat162_rn(a, b);
    }
  }
}

// bias[h, i, j] = log2(e) * table[bucket[j - i + n - 1], h] for i, j < n; 0 in the padding (rows / columns up to the
// next multiple of 128, masked or unused by the attention kernel)
__global__ void t5_bias_kernel(const float* __restrict__ table, const int* __restrict__ bucket, int H, int n, int rows_pad,
                               int ld, float* __restrict__ bias) {
  const long long total = (long long)H * rows_pad * ld;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int j = (int)(t % ld);
    const int i = (int)((t / ld) % rows_pad);
    const int h = (int)(t / ((long long)ld * rows_pad));
    float v = 0.f;
    if (i < n && j < n) {
      int b = bucket[j - i + n - 1];
      b = b < 0 ? 0 : (b > 31 ? 31 : b);
      v = table[b * H + h] * 1.4426950408889634f;
    }
    bias[t] = v;
  }
}

// g *= u (bf16 pairs)
__global__ void t5_mul_kernel(bf16* __restrict__ g, const bf16* __restrict__ u, size_t pairs) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < pairs; i += (size_t)gridDim.x * blockDim.x) {
    const float2 a = __bfloat1622float2(reinterpret_cast<const __nv_bfloat162*>(g)[i]);
    const float2 b = __bfloat1622float2(reinterpret_cast<const __nv_bfloat162*>(u)[i]);
    reinterpret_cast<__nv_bfloat162*>(g)[i] = __floats2bfloat162_rn(a.x * b.x, a.y * b.y);
  }
}

// ------------------------------------------------------------------------------------------------ host side
static void t5_finalize(ldn_engine* e) {
  LDN_CHECK(!e->w[5].empty(), "T5 weights not loaded");
  e->t5.reset(new ldn_engine::T5State());
  auto& T = *e->t5;
  const DevTensor& emb = e->W(5, "shared.weight");
  LDN_CHECK(emb.is_bf16 && emb.shape.size() == 2, "T5: shared.weight must be a matrix");
  T.vocab = (int)emb.shape[0];
  T.width = (int)emb.shape[1];
  const DevTensor& rb = e->W(5, "encoder.block.0.layer.0.SelfAttention.relative_attention_bias.weight");
  LDN_CHECK(!rb.is_bf16 && rb.shape.size() == 2 && rb.shape[0] == 32, "T5: relative_attention_bias must be fp32 [32, heads]");
  T.heads = (int)rb.shape[1];
  T.ff = (int)e->W(5, "encoder.block.0.layer.1.DenseReluDense.wi_0.weight").shape[0];
  while (e->has(5, "encoder.block." + std::to_string(T.layers) + ".layer.0.SelfAttention.q.weight")) ++T.layers;
  LDN_CHECK(T.layers > 0 && T.width % T.heads == 0 && T.width / T.heads == 64,
            "T5: the attention kernel's logit-bias variant is built for 64-wide heads");
  LDN_CHECK(T.width % 16 == 0 && T.ff % 16 == 0, "T5: widths must be multiples of 16");
  e->finalized[5] = true;
}

static Program* build_t5_program(ldn_engine* e, int S, int n) {
  auto& T = *e->t5;
  std::unique_ptr<Program> prog(new Program());
  T.program_arenas.emplace_back(new Arena());
  Arena& A = *T.program_arenas.back();
  prog->arena = &A;
  const int W = T.width, H = T.heads, d = 64, slot = 64, F = T.ff, M = S * n;
  const int Mld = (M + 63) / 64 * 64;  // V^T row length: a multiple of 64 so that the v^T GEMM has whole N tiles only
  const int nk_pad = (n + 7) / 8 * 8;
  const int bias_rows = (n + 127) / 128 * 128, bias_ld = bias_rows;
  const auto key = std::make_pair(S, n);
  long long* ids = A.get<long long>(M);
  int* buckets = A.get<int>(2 * n - 1);
  bf16* X = A.get<bf16>((size_t)M * W);
  bf16* N1 = A.get<bf16>((size_t)M * W);
  bf16* QK = A.get<bf16>((size_t)M * 2 * W, true);
  bf16* Vt = A.get<bf16>((size_t)W * Mld, true);
  bf16* VtP = A.get<bf16>((size_t)W * S * nk_pad, true);
  bf16* O = A.get<bf16>((size_t)M * W);
  bf16* G = A.get<bf16>((size_t)M * F);
  bf16* U = A.get<bf16>((size_t)M * F);
  float* bias = A.get<float>((size_t)H * bias_rows * bias_ld, true);
  float* out = A.get<float>((size_t)M * W);
  T.in_ids[key] = ids;
  T.in_buckets[key] = buckets;
  T.out[key] = out;

  auto add = [&](const std::string& name, Step s) {
    prog->steps.push_back(std::move(s));
    prog->names.push_back(name);
    prog->launches += 1;
  };
  auto gemm = [&](const std::string& name, const GemmArgs& a) {
    GemmPlan plan = make_gemm_plan(a);
    add(name, [plan](cudaStream_t st) { launch_gemm(plan, st); });
  };
  auto rms = [&](const std::string& name, const bf16* x, const float* w, bf16* o, float* o32) {
    add(name, [=](cudaStream_t st) {
      t5_rms_kernel<<<M, 256, 0, st>>>(x, w, W, 1e-6f, o, o32);
      LDN_CUDA(cudaGetLastError());
    });
  };
  {
    const bf16* table = e->W(5, "shared.weight").b();
    const int vocab = T.vocab;
    add("embed", [=](cudaStream_t st) {
      t5_embed_kernel<<<296, 256, 0, st>>>(ids, table, M, W, vocab, X);
      LDN_CUDA(cudaGetLastError());
    });
    const float* rtab = e->W(5, "encoder.block.0.layer.0.SelfAttention.relative_attention_bias.weight").f();
    add("relative_bias", [=](cudaStream_t st) {
      t5_bias_kernel<<<592, 256, 0, st>>>(rtab, buckets, H, n, bias_rows, bias_ld, bias);
      LDN_CUDA(cudaGetLastError());
    });
  }
  for (int i = 0; i < T.layers; ++i) {
    const std::string p = "encoder.block." + std::to_string(i) + ".layer";
    rms(p + ".0.rms", X, e->W(5, p + ".0.layer_norm.weight").f(), N1, nullptr);
    GemmArgs q;
    q.A0 = N1; q.lda0 = W; q.K0 = W; q.Wt = e->W(5, p + ".0.SelfAttention.q.weight").b(); q.M = M; q.N = W;
    q.out = QK; q.ldo = 2LL * W; q.head_dim = d; q.head_slot = slot;
    gemm(p + ".0.q", q);
    GemmArgs k = q;
    k.Wt = e->W(5, p + ".0.SelfAttention.k.weight").b(); k.out = QK + W;
    gemm(p + ".0.k", k);
    GemmArgs v;
    v.A0 = e->W(5, p + ".0.SelfAttention.v.weight").b(); v.lda0 = W; v.K0 = W; v.Wt = N1; v.M = W; v.N = Mld; v.wt_rows = M;
    v.out = Vt; v.ldo = Mld;
    gemm(p + ".0.vt", v);
    const bf16* vt_use = Vt;
    long long ldvt = Mld;
    if (nk_pad != n) {  // re-lay V^T so that every sequence starts on a 16-byte boundary
      add(p + ".0.vt_pad", [=](cudaStream_t st) { launch_pad_vt_cols(Vt, Mld, W, S, n, nk_pad, VtP, st); });
      vt_use = VtP;
      ldvt = (long long)S * nk_pad;
    }
    AttnArgs at;
    at.Q = QK; at.ldq = 2LL * W; at.K = QK + W; at.ldk = 2LL * W;
    at.Vt = vt_use; at.ldvt = ldvt; at.vt_rows = W;
    at.B = S; at.heads = H; at.Nq = n; at.Nk = n; at.nk_pad = nk_pad; at.kv_batch_stride = n; at.d = d; at.slot = slot;
    at.causal = 0; at.scale = 1.0f; at.out = O; at.ldo = W;
    at.bias = bias; at.bias_rows = bias_rows; at.bias_ld = bias_ld;
    AttnPlan ap = make_attn_plan(at);
    add(p + ".0.sdpa", [ap](cudaStream_t st) { launch_attn(ap, st); });
    GemmArgs o;
    o.A0 = O; o.lda0 = W; o.K0 = W; o.Wt = e->W(5, p + ".0.SelfAttention.o.weight").b(); o.M = M; o.N = W;
    o.residual = X; o.ldr = W; o.out = X; o.ldo = W;
    gemm(p + ".0.o", o);
    rms(p + ".1.rms", X, e->W(5, p + ".1.layer_norm.weight").f(), N1, nullptr);
    GemmArgs g;
    g.A0 = N1; g.lda0 = W; g.K0 = W; g.Wt = e->W(5, p + ".1.DenseReluDense.wi_0.weight").b(); g.M = M; g.N = F;
    g.act = 4; g.out = G; g.ldo = F;  // GELU, tanh approximation
    gemm(p + ".1.wi_0", g);
    GemmArgs u = g;
    u.Wt = e->W(5, p + ".1.DenseReluDense.wi_1.weight").b(); u.act = 0; u.out = U;
    gemm(p + ".1.wi_1", u);
    const size_t pairs = (size_t)M * F / 2;
    add(p + ".1.gate", [=](cudaStream_t st) {
      t5_mul_kernel<<<592, 256, 0, st>>>(G, U, pairs);
      LDN_CUDA(cudaGetLastError());
    });
    GemmArgs wo;
    wo.A0 = G; wo.lda0 = F; wo.K0 = F; wo.Wt = e->W(5, p + ".1.DenseReluDense.wo.weight").b(); wo.M = M; wo.N = W;
    wo.residual = X; wo.ldr = W; wo.out = X; wo.ldo = W;
    gemm(p + ".1.wo", wo);
  }
  rms("final_rms", X, e->W(5, "encoder.final_layer_norm.weight").f(), nullptr, out);
  return prog.release();
}

void t5_encode(ldn_engine* e, const int64_t* ids, const int32_t* rel_buckets, int S, int n, float* out, cudaStream_t stream) {
  if (!e->finalized[5]) t5_finalize(e);
  auto& T = *e->t5;
  LDN_CHECK(S >= 1 && n >= 1, "T5: empty batch");
  const auto key = std::make_pair(S, n);
  auto it = T.programs.find(key);
  if (it == T.programs.end()) it = T.programs.emplace(key, std::unique_ptr<Program>(build_t5_program(e, S, n))).first;
  Program& P = *it->second;
  const size_t M = (size_t)S * n;
  LDN_CUDA(cudaMemcpyAsync(T.in_ids[key], ids, M * sizeof(long long), cudaMemcpyDeviceToDevice, stream));
  LDN_CUDA(cudaMemcpyAsync(T.in_buckets[key], rel_buckets, (size_t)(2 * n - 1) * sizeof(int), cudaMemcpyDeviceToDevice, stream));
  run_program(P, e->cfg.use_graph != 0, stream);
  LDN_CUDA(cudaMemcpyAsync(out, T.out[key], M * T.width * sizeof(float), cudaMemcpyDeviceToDevice, stream));
}

}  // namespace ldn
