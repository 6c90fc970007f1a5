/* libldn — C ABI of the B200-native SD1.5 sampling engine.
 *
 * Drop-in boundary for LightDiffusion-Next's sampler hot path.  Every entry point takes plain
 * pointers / sizes (device pointers unless stated), launches asynchronously on the CUDA stream
 * passed as `stream` (a cudaStream_t cast to void*; NULL = legacy default stream), returns 0 on
 * success and non-zero on failure with a message available from ldn_last_error().
 *
 * Reference interfaces replaced (paths relative to the reference repo):
 *   ldn_unet_denoise   <- BaseModel.apply_model          src/Model/ModelBase.py:72-133
 *                         (called through model_options["model_function_wrapper"], src/cond/cond.py:254-265)
 *   ldn_set_context    <- CrossAttention.to_k/to_v        src/Attention/Attention.py:118-121 (hoisted out of the loop)
 *   ldn_cfg_step       <- cfg_function + sampler update   src/sample/CFG.py:55-60, src/sample/samplers.py:728-732,952-953
 *   ldn_vae_decode     <- VAE.decode                      src/AutoEncoders/VariationalAE.py:690-722
 *   ldn_vae_encode     <- AutoencodingEngine.encode (w/o the sampling step)  src/AutoEncoders/VariationalAE.py:148-172, 377-413
 *   ldn_flux_forward   <- Flux3.forward_orig                src/BlackForest/Flux.py:658-730
 *   ldn_taesd_decode   <- TAESD.decode (preview)           src/AutoEncoders/taesd.py:104-136,190-197
 *   ldn_clip_encode    <- CLIPTextModel_.forward          src/clip/CLIPTextModel.py:51-107
 *   ldn_t5_encode      <- T5.forward (Flux text encoder)  src/clip/FluxClip.py:457-562
 *   op-level entries   <- the torch library calls of      src/cond/cast.py:107,174,241,281 and
 *                         optimized_attention             src/Attention/Attention.py:34-41
 */
#ifndef LDN_H
#define LDN_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct ldn_engine* ldn_handle;

/* dtype codes for weight ingest */
enum { LDN_F32 = 0, LDN_F16 = 1, LDN_BF16 = 2 };

typedef struct ldn_tensor {
  const char* name;   /* SD1.5 / LDM state-dict key, e.g. "input_blocks.1.0.in_layers.2.weight" */
  const void* data;   /* device pointer, contiguous */
  int dtype;          /* LDN_F32 | LDN_F16 | LDN_BF16 */
  int ndim;
  int64_t shape[4];
} ldn_tensor;

typedef struct ldn_config {
  int max_rows;       /* max UNet batch rows (cond+uncond rows, i.e. 2*bs) */
  int max_h, max_w;   /* max latent height / width */
  int max_ctx_tokens; /* max context tokens per row (77*k) */
  int use_graph;      /* capture the UNet forward into a CUDA graph per (rows,h,w) */
} ldn_config;

const char* ldn_last_error(void);
int ldn_version(void);

/* ---- engine lifecycle */
int ldn_create(const ldn_config* cfg, ldn_handle* out);
void ldn_destroy(ldn_handle h);
/* which: 0 = UNet ("model.diffusion_model." prefix stripped), 1 = VAE ("first_stage_model." stripped; decoder.* +
 *        post_quant_conv.* and/or encoder.* + quant_conv.*), 2 = CLIP-L text model ("...text_model." stripped),
 *        3 = TAESD preview decoder (keys of taesd_decoder.safetensors), 4 = Flux.1 DiT (Flux3 state-dict keys),
 *        5 = T5 text encoder (keys of the reference's T5 module, src/clip/FluxClip.py:501-531).
 * Matrices are stored as bf16; vectors, embedding tables named "*embedding*" and the T5 relative_attention_bias table as fp32. */
int ldn_load_weights(ldn_handle h, int which, const ldn_tensor* tensors, int n, void* stream);
/* Discrete schedule tables (ModelSamplingDiscrete.sigmas / .log_sigmas, src/sample/sampling.py:221-356): host
 * pointers, n entries each. log_sigmas is passed separately because the reference computes it in float64. */
int ldn_set_sigmas(ldn_handle h, const float* sigmas_host, const float* log_sigmas_host, int n);

/* ---- UNet hot path */
/* ctx: [rows, tokens, 768] fp32 (device). Pre-computes cross-attention K/V for all 16 transformer blocks. */
int ldn_set_context(ldn_handle h, const float* ctx, int rows, int tokens, void* stream);
/* x: [rows,4,h,w] fp32 NCHW; sigma: [rows] fp32; out: denoised = x - eps*sigma, [rows,4,h,w] fp32. */
int ldn_unet_denoise(ldn_handle h, const float* x, const float* sigma, float* out, int rows, int lat_h, int lat_w,
                     void* stream);
/* number of kernels the last ldn_unet_denoise program launches per call (as counted when it was built) */
int ldn_unet_last_launches(ldn_handle h);
/* Fused CFG combine + solver update, fp32 elementwise over n elements.
 *  denoised = uncond + (cond - uncond) * cfg
 *  mode 0 (dpmpp_2m_cfgpp as executed by the reference, first order): x' = c0*x - c1*denoised
 *  mode 1 (euler ancestral): x' = x + (x - denoised) / c2 * c0 + noise * c1   (c2 = sigma)
 *  mode 2: only write denoised */
int ldn_cfg_step(const float* x, const float* den_uncond, const float* den_cond, float cfg, int mode, float c0,
                 float c1, float c2, const float* noise, float* x_out, float* denoised_out, int64_t n, void* stream);

/* Bilinear resample of fp32 planes, src [planes,h,w] -> dst [planes,oh,ow], with the arithmetic of
 * F.interpolate(mode="bilinear", align_corners=False): the down / up-scaling around the reference's half-resolution
 * (multiscale) sampler steps, src/sample/samplers.py:821-835 (dpmpp_2m_cfgpp), :1035-1051 (dpmpp_sde_cfgpp). */
int ldn_resample_bilinear(const float* src, float* dst, int planes, int h, int w, int oh, int ow, void* stream);
/* `bislerp` of the reference's LatentUpscale (src/Utilities/upscale.py:5-128, 144-166: the HiresFix branch of pipeline()):
 * src [n, c, h, w] fp32 -> dst [n, c, oh, ow] fp32, a separable resize (width first) whose two-tap blend is a spherical
 * interpolation of the c-vectors with the tap positions / ratios of bilinear resampling.  tmp: n * c * h * ow floats of
 * scratch.  All pointers are device pointers. */
int ldn_bislerp(const float* src, float* tmp, float* dst, int n, int c, int h, int w, int oh, int ow, void* stream);

/* ---- VAE decode / CLIP encode */
/* z: [B,zc,h,w] fp32, already un-scaled by the latent format (SD1.x: zc = 4, z / 0.18215; Flux VAE: zc = 16, no
 * post_quant_conv -- both read off the loaded decoder weights); rgb: [B,8h,8w,3] fp32 in [0,1] */
int ldn_vae_decode(ldn_handle h, const float* z, float* rgb, int B, int lat_h, int lat_w, void* stream);
/* pixels: [B,3,H,W] fp32 already mapped to [-1,1] (process_input, VariationalAE.py:601); moments: [B,8,H/8,W/8] fp32
 * (mean | logvar after quant_conv). The reparameterised sample mean + exp(0.5*clamp(logvar,-30,20))*randn stays with the
 * caller, which owns the RNG (DiagonalGaussianDistribution.sample, VariationalAE.py:42-51). Needs encoder.* weights. */
int ldn_vae_encode(ldn_handle h, const float* pixels, float* moments, int B, int H, int W, void* stream);
/* Flux.1 DiT forward (Flux3.forward_orig, src/BlackForest/Flux.py:658-730) on already patchified tokens.
 * img: [B, n_img, 64] fp32 (2x2 patches of the 16-channel latent), ctx: [B, n_txt, ctx_dim] fp32 (T5 states),
 * pe: [n_txt + n_img, 64, 2] fp32 (cos, sin) rotary table of EmbedND for the concatenated (txt, img) ids -- shared by all rows,
 * t / guidance: [B] fp32 (sigma in [0,1]; guidance may be NULL for non-distilled models), y: [B, vec_dim] fp32 (CLIP pooled),
 * out: [B, n_img, 64] fp32. n_txt must be a multiple of 8. Weights: ldn_load_weights(which = 4), Flux3 state-dict keys. */
int ldn_flux_forward(ldn_handle h, const float* img, const float* ctx, const float* pe, const float* t,
                     const float* guidance, const float* y, float* out, int B, int n_img, int n_txt, void* stream);
/* TAESD preview decoder (src/AutoEncoders/taesd.py:104-136, TAESD.decode :190-197 before its sub(0.5).mul(2)):
 * z: [B,4,h,w] fp32 raw latent; rgb: [B,8h,8w,3] fp32, the decoder's raw output (~[0,1], not clamped).
 * Weights: ldn_load_weights(which = 3) with the keys of taesd_decoder.safetensors (nn.Sequential indices). */
int ldn_taesd_decode(ldn_handle h, const float* z, float* rgb, int B, int lat_h, int lat_w, void* stream);
/* ids: [S,77] int64 (device); out_last: [S,77,768] fp32 final-LN of last layer (may be NULL);
 * out_penultimate: [S,77,768] fp32 final-LN of layer -2 (what SD1.5 uses) */
int ldn_clip_encode(ldn_handle h, const int64_t* ids, int S, float* out_penultimate, float* out_last, void* stream);
/* Textual-inversion vectors for the CLIP token table: vectors [n, 768] fp32 (device) become token ids vocab, vocab + 1, ...
 * of the loaded table, which itself is untouched (SDClipModel.set_up_textual_embeddings, src/SD15/SDClip.py:213-268).
 * n <= 256; n = 0 clears them. No re-upload of the table, no program rebuild. */
int ldn_clip_set_extra_embeddings(ldn_handle h, const float* vectors, int n, void* stream);
/* T5 text encoder of the Flux path (T5.forward / T5Stack.forward, src/clip/FluxClip.py:457-562; no attention mask, layer
 * "last" + final RMS norm as T5XXLModel configures SDClipModel, :565-590).  ids: [S,n] int64 (device), any n >= 1;
 * rel_buckets: [2n-1] int32 (device), the reference's relative-position bucket (T5Attention._relative_position_bucket,
 * :153-205) of every distance key - query = -(n-1) .. n-1, computed by the host mirror with the reference's own fp32
 * arithmetic; out: [S,n,d_model] fp32.  Weights: ldn_load_weights(which = 5) with the state-dict keys of the reference's
 * T5 module (shared.weight, encoder.block.N.layer.{0,1}.*, encoder.final_layer_norm.weight); heads must be 64 wide. */
int ldn_t5_encode(ldn_handle h, const int64_t* ids, const int32_t* rel_buckets, int S, int n, float* out, void* stream);

/* ---- op-level entries (used by the parity tests; same kernels the engine runs) */
/* out[M,N] = [A0 | A1][M,K0+K1] * Wt[N,K]^T (+bias) (+rowbias[row / rows_per_batch]) (+residual); bf16 in/out.
 * epi: 0 plain, 1 GEGLU (Wt/bias rows pre-interleaved in blocks of BN/2; out has N/2 columns).
 * head_dim > 0: scatter output column n to (n / head_dim) * head_slot + n % head_dim. */
int ldn_gemm_bf16(const void* A0, int64_t lda0, int K0, const void* A1, int64_t lda1, int K1, const void* Wt, int M,
                  int N, const float* bias, const float* rowbias, int ld_rowbias, int rows_per_batch,
                  const void* residual, int64_t ldr, void* out, int64_t ldo, float* out_f32, int epi, int head_dim,
                  int head_slot, int BN, void* stream);
/* x: NHWC bf16 [B,H,W,Cin]; Wt: [Cout, 3,3, Cin] bf16; out NHWC bf16 [B,H,W,Cout]; stride 1, pad 1. */
int ldn_conv3x3_bf16(const void* x, const void* Wt, int B, int H, int W, int Cin, int Cout, const float* bias,
                     const float* rowbias, int ld_rowbias, const void* residual, void* out, void* stream);
/* Q: [B*Nq, heads*slot], K: [B*nk_pad, heads*slot], Vt: [vt_rows, B*nk_pad] (all bf16); out: [B*Nq, heads*d].
 * vt_head_stride: rows per head in Vt; 0 or d = plain V^T. For d = 40 (stride 48) and d = 80 (stride 96) the padded
 * layout selects the fastest kernels and requires row d of every head to be all ones (remaining pad rows zero): the
 * softmax row sum is then computed by the tensor core.
 * causal: bit 0 = causal mask; bit 1 (d = 40 with the padded layout only) = FOLDED operands: Q already carries
 * scale * log2(e) and column 40 of every K head slot holds 1.0 -- the long-sequence kernel then keeps the running offset
 * -m in column 40 of its Q tile, so the tensor core delivers scaled, offset scores and the softmax needs no scale-subtract
 * (`scale` is ignored; this is how the UNet program calls its level-0 self-attention). */
int ldn_attention_bf16(const void* Q, int64_t ldq, const void* K, int64_t ldk, const void* Vt, int64_t ldvt,
                       int64_t vt_rows, int vt_head_stride, int B, int heads, int Nq, int Nk, int nk_pad, int d,
                       int slot, int causal, float scale, void* out, int64_t ldo, void* stream);
/* GroupNorm (+SiLU) over NHWC bf16, input may be a channel concat of x0 (C0) and x1 (C1, may be NULL/0). */
int ldn_groupnorm_bf16(const void* x0, int C0, const void* x1, int C1, int B, int HW, int groups, float eps,
                       const float* gamma, const float* beta, int silu, void* out, void* stream);
/* The ResBlock pair conv3x3 -> GroupNorm(32 groups)(+SiLU) as the UNet program runs it (src/AutoEncoders/ResBlock.py:251-292:
 * in_layers[2] -> out_layers[0..1]): where the plan allows (whole 128-pixel tiles of one image, Cout 320 / 640 /
 * 1280, no split-K) the conv's epilogue accumulates the GroupNorm statistics of its own output and only the apply kernel
 * follows; otherwise the statistics kernel runs.  conv_out: the conv result (NHWC bf16), gn_out: the normalised result;
 * *fused (host int, may be NULL) reports which of the two happened. */
int ldn_conv3x3_groupnorm_bf16(const void* x, const void* Wt, int B, int H, int W, int Cin, int Cout, const float* bias,
                               const float* rowbias, int ld_rowbias, const void* residual, float eps, const float* gamma,
                               const float* beta, int silu, void* conv_out, void* gn_out, int* fused, void* stream);
int ldn_layernorm_bf16(const void* x, int rows, int C, float eps, const float* gamma, const float* beta, void* out,
                       void* stream);

#ifdef __cplusplus
}
#endif
#endif /* LDN_H */
