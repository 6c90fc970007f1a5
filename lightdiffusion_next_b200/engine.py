"""Python host for libldn.so: owns an engine handle, feeds it torch CUDA tensors by pointer.

PyTorch is plumbing here (device memory, streams, RNG); every FLOP of the UNet / VAE / CLIP forward runs in the
hand-written sm_100a kernels behind the C ABI (include/ldn.h).  There is no CPU or eager fallback.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional

import torch

from . import _lib as L
from .schedule import DiscreteSchedule

_DTYPE_CODE = {torch.float32: 0, torch.float16: 1, torch.bfloat16: 2}
UNET, VAE, CLIP, TAESD, FLUX, T5 = 0, 1, 2, 3, 4, 5


class Engine:
    """One engine per GPU (per process).  Thread-compatible, not thread-safe."""

    def __init__(self, max_rows: int = 2, max_h: int = 128, max_w: int = 128, max_ctx_tokens: int = 77,
                 use_graph: bool = True, device: Optional[torch.device] = None):
        if not torch.cuda.is_available():
            raise L.LdnError("no CUDA device: the B200 engine has no CPU fallback")
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        self.lib = L.load()
        cfg = L.ldn_config(max_rows, max_h, max_w, max_ctx_tokens, int(use_graph))
        h = C.c_void_p()
        with torch.cuda.device(self.device):
            L.check(self.lib.ldn_create(C.byref(cfg), C.byref(h)))
        self.h = h
        self.schedule = DiscreteSchedule()
        sig = self.schedule.sigmas.contiguous()
        lsig = self.schedule.log_sigmas.contiguous()
        L.check(self.lib.ldn_set_sigmas(self.h, C.cast(sig.data_ptr(), C.POINTER(C.c_float)),
                                        C.cast(lsig.data_ptr(), C.POINTER(C.c_float)), sig.numel()))
        self._ctx_key = None
        self._keep = []
        self.weights_epoch: Dict[int, int] = {}
        self.context_uploads = 0  # ldn_set_context calls (each: pad + 32 K/V projections)

    def close(self) -> None:
        if getattr(self, "h", None):
            self.lib.ldn_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ weights
    def load_weights(self, which: int, state_dict: Dict[str, torch.Tensor], chunk: int = 64) -> None:
        """state_dict: LDM key names with the model prefix stripped; tensors on any device, fp32/fp16/bf16."""
        items = list(state_dict.items())
        with torch.cuda.device(self.device):
            stream = L.cur_stream()
            for i in range(0, len(items), chunk):
                part = items[i:i + chunk]
                arr = (L.ldn_tensor * len(part))()
                keep = []
                for j, (name, t) in enumerate(part):
                    if t.dtype not in _DTYPE_CODE:
                        t = t.float()
                    t = t.detach().to(self.device, non_blocking=True).contiguous()
                    keep.append(t)
                    arr[j].name = name.encode()
                    arr[j].data = t.data_ptr()
                    arr[j].dtype = _DTYPE_CODE[t.dtype]
                    arr[j].ndim = t.dim()
                    for k, s in enumerate(t.shape):
                        arr[j].shape[k] = s
                L.check(self.lib.ldn_load_weights(self.h, which, arr, len(part), stream))
                del keep
        self._ctx_key = None
        self.weights_epoch[which] = self.weights_epoch.get(which, 0) + 1  # a UNet reload drops the uploaded context

    def load_unet(self, state_dict: Dict[str, torch.Tensor]) -> None:
        self.load_weights(UNET, state_dict)

    def load_vae(self, state_dict: Dict[str, torch.Tensor]) -> None:
        self.load_weights(VAE, state_dict)

    def load_clip(self, state_dict: Dict[str, torch.Tensor]) -> None:
        self.load_weights(CLIP, state_dict)
        self._clip_tok = state_dict.get("embeddings.token_embedding.weight")  # dtype / vocabulary size of the table
        self._clip_ti_key = None
        self._clip_extra_n = 0

    def clip_vocab(self) -> int:
        tok = getattr(self, "_clip_tok", None)
        if tok is None:
            raise L.LdnError("CLIP weights not loaded (Engine.load_clip)")
        return int(tok.shape[0])

    def set_clip_extra_embeddings(self, vectors) -> None:
        """Textual-inversion vectors become token ids vocab, vocab + 1, ... (what the reference's
        set_up_textual_embeddings does with a temporary Embedding, src/SD15/SDClip.py:247-259).  They go into a small
        separate device table (ldn_clip_set_extra_embeddings): the checkpoint's 150 MB token table is neither re-uploaded
        nor replaced, and no program is rebuilt.  The vectors are rounded to the table's dtype exactly as there."""
        self.clip_vocab()
        vec = torch.stack([v.detach().cpu() for v in vectors]).to(self._clip_tok.dtype).float().contiguous() \
            if len(vectors) else torch.zeros(0, self._clip_tok.shape[1])
        key = hash(vec.numpy().tobytes())
        if key != self._clip_ti_key:
            dv = vec.to(self.device)
            with torch.cuda.device(self.device):
                L.check(self.lib.ldn_clip_set_extra_embeddings(self.h, dv.data_ptr(), int(vec.shape[0]), L.cur_stream()))
            self._clip_ti_key = key
            self._clip_extra_n = int(vec.shape[0])

    def load_taesd(self, state_dict: Dict[str, torch.Tensor]) -> None:
        """TAESD preview decoder weights (keys of `taesd_decoder.safetensors`: nn.Sequential indices, taesd.py:104-136)."""
        self.load_weights(TAESD, state_dict)

    def load_flux(self, state_dict: Dict[str, torch.Tensor]) -> None:
        """Flux.1 DiT weights (Flux3 state-dict keys, src/BlackForest/Flux.py:548-656)."""
        from . import flux as FX

        FX.validate_state_dict(state_dict)  # fails loudly on a malformed checkpoint
        self.load_weights(FLUX, state_dict)
        self._flux_pe = {}

    def load_t5(self, state_dict: Dict[str, torch.Tensor]) -> None:
        """T5 text-encoder weights (state-dict keys of the reference's T5 module, src/clip/FluxClip.py:501-531; see
        t5.t5_shapes).  Heads must be 64 wide (T5-XXL: 4096 / 64)."""
        from . import t5 as T5H

        cfg = T5H.validate_state_dict(state_dict)
        self.load_weights(T5, {k: state_dict[k] for k in T5H.t5_shapes(**cfg)})  # encoder tensors only
        self._t5_width = cfg["d_model"]

    def load_checkpoint(self, path: str, lora_path: Optional[str] = None, strength_model: float = 1.0,
                        strength_clip: float = 1.0) -> Dict[str, int]:
        """SD1.5 checkpoint file (.safetensors / .ckpt), optionally with a LoRA folded in -> UNet / VAE / CLIP weights
        (see checkpoint.py; the reference side is load_checkpoint_guess_config, src/FileManaging/Loader.py:11-111)."""
        from . import checkpoint

        return checkpoint.load_sd15(self, path, lora_path, strength_model, strength_clip)

    # ------------------------------------------------------------------ UNet hot path
    def set_context(self, ctx: torch.Tensor) -> None:
        """ctx: [rows, tokens, 768]; rows ordered as the UNet batch rows (uncond first, then cond)."""
        ctx = ctx.to(self.device, torch.float32).contiguous()
        rows, tokens, _ = ctx.shape
        with torch.cuda.device(self.device):
            L.check(self.lib.ldn_set_context(self.h, ctx.data_ptr(), rows, tokens, L.cur_stream()))
        self._keep = [ctx]
        self.context_uploads += 1

    def denoise(self, x: torch.Tensor, sigma: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """x [rows,4,h,w] fp32, sigma [rows] fp32 -> denoised = x - eps*sigma (BaseModel.apply_model semantics)."""
        assert x.is_cuda and x.dtype == torch.float32 and x.is_contiguous()
        sigma = sigma.to(self.device, torch.float32).contiguous()
        if out is None:
            out = torch.empty_like(x)
        rows, _, h, w = x.shape
        with torch.cuda.device(self.device):
            L.check(self.lib.ldn_unet_denoise(self.h, x.data_ptr(), sigma.data_ptr(), out.data_ptr(), rows, h, w,
                                              L.cur_stream()))
        return out

    def cfg_step(self, x, den_uncond, den_cond, cfg: float, mode: int, c0: float = 0.0, c1: float = 0.0,
                 c2: float = 0.0, noise=None, x_out=None, denoised_out=None) -> None:
        n = den_uncond.numel()
        with torch.cuda.device(self.device):
            L.check(self.lib.ldn_cfg_step(L.ptr(x), den_uncond.data_ptr(), den_cond.data_ptr(), float(cfg), mode,
                                          float(c0), float(c1), float(c2), L.ptr(noise), L.ptr(x_out),
                                          L.ptr(denoised_out), n, L.cur_stream()))

    def resample_bilinear(self, x: torch.Tensor, size, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """F.interpolate(x, size=size, mode="bilinear", align_corners=False) on fp32 NCHW (the resampling around the
        reference's half-resolution sampler steps, samplers.py:821-835), as one launch of the engine's own kernel."""
        assert x.is_cuda and x.dtype == torch.float32
        x = x.contiguous()
        B, Cc, h, w = x.shape
        oh, ow = int(size[0]), int(size[1])
        if out is None:
            out = torch.empty(B, Cc, oh, ow, device=x.device, dtype=torch.float32)
        with torch.cuda.device(self.device):
            L.check(self.lib.ldn_resample_bilinear(x.data_ptr(), out.data_ptr(), B * Cc, h, w, oh, ow, L.cur_stream()))
        return out

    def bislerp(self, x: torch.Tensor, width: int, height: int) -> torch.Tensor:
        """`bislerp(samples, width, height)` of the reference (src/Utilities/upscale.py:5-128) on the device: [n,c,h,w] fp32 ->
        [n,c,height,width] fp32, two launches of the engine's separable slerp kernel (width first, as the reference)."""
        assert x.is_cuda
        xf = x.to(torch.float32).contiguous()
        n, c, h, w = xf.shape
        tmp = torch.empty(n, c, h, int(width), device=xf.device, dtype=torch.float32)
        out = torch.empty(n, c, int(height), int(width), device=xf.device, dtype=torch.float32)
        with torch.cuda.device(self.device):
            L.check(self.lib.ldn_bislerp(xf.data_ptr(), tmp.data_ptr(), out.data_ptr(), n, c, h, w, int(height), int(width),
                                         L.cur_stream()))
        return out.to(x.dtype)

    # ------------------------------------------------------------------ VAE / CLIP
    def vae_decode(self, z: torch.Tensor) -> torch.Tensor:
        """z [B,4,h,w] fp32 (already / 0.18215) -> [B,8h,8w,3] fp32 in [0,1] (VAE.decode semantics)."""
        z = z.to(self.device, torch.float32).contiguous()
        B, _, h, w = z.shape
        out = torch.empty(B, 8 * h, 8 * w, 3, device=self.device, dtype=torch.float32)
        with torch.cuda.device(self.device):
            L.check(self.lib.ldn_vae_decode(self.h, z.data_ptr(), out.data_ptr(), B, h, w, L.cur_stream()))
        return out

    def vae_encode_moments(self, pixels: torch.Tensor) -> torch.Tensor:
        """pixels [B,H,W,3] fp32 in [0,1] -> Gaussian moments [B,8,H/8,W/8] fp32 on the device (quant_conv(Encoder(2x-1))).
        Like the reference (whose vae_encode_crop_pixels is a no-op, VariationalAE.py:677-688) nothing is cropped."""
        x = (pixels[..., :3].to(self.device, torch.float32).movedim(-1, 1) * 2.0 - 1.0).contiguous()
        B, _, H, W = x.shape
        out = torch.empty(B, 8, H // 8, W // 8, device=self.device, dtype=torch.float32)
        with torch.cuda.device(self.device):
            L.check(self.lib.ldn_vae_encode(self.h, x.data_ptr(), out.data_ptr(), B, H, W, L.cur_stream()))
        return out

    def vae_encode(self, pixels: torch.Tensor, noise: Optional[torch.Tensor] = None) -> torch.Tensor:
        """VAE.encode semantics (VariationalAE.py:725-760): a posterior SAMPLE, fp32 on the CPU. The noise is drawn like
        DiagonalGaussianDistribution.sample does -- torch.randn(mean.shape) from the global CPU generator -- unless given."""
        m = self.vae_encode_moments(pixels)
        mean, logvar = m.chunk(2, dim=1)
        if noise is None:
            noise = torch.randn(mean.shape)
        std = torch.exp(0.5 * torch.clamp(logvar, -30.0, 20.0))
        return (mean + std * noise.to(m.device)).float().cpu()

    def taesd_decode(self, z: torch.Tensor) -> torch.Tensor:
        """Raw latent [B,4,h,w] -> the TAESD decoder's raw output [B,8h,8w,3] fp32 on the device (~[0,1], unclamped).
        TAESD.decode returns this * 2 - 1 (taesd.py:190-197); taesd_preview maps it back and clamps for display."""
        z = z.to(self.device, torch.float32).contiguous()
        B, _, h, w = z.shape
        out = torch.empty(B, 8 * h, 8 * w, 3, device=self.device, dtype=torch.float32)
        with torch.cuda.device(self.device):
            L.check(self.lib.ldn_taesd_decode(self.h, z.data_ptr(), out.data_ptr(), B, h, w, L.cur_stream()))
        return out

    def flux_forward(self, x: torch.Tensor, timestep: torch.Tensor, context: torch.Tensor, y: torch.Tensor,
                     guidance: Optional[torch.Tensor] = None, axes_dim=(16, 56, 56), theta: int = 10000) -> torch.Tensor:
        """Flux3.forward (Flux.py:732-778): latent [B,16,h,w] (even h, w), timestep [B], context [B,Nt,ctx], y [B,vec],
        guidance [B] -> model output [B,16,h,w] fp32. Patchify / unpatchify and the rotary table are boundary work in torch."""
        from . import flux as FX

        x = x.to(self.device, torch.float32)
        B, c, h, w = x.shape
        img, hl, wl = FX.patchify(x)
        Nt = context.shape[1]
        key = (hl, wl, Nt, tuple(axes_dim), theta)
        cache = self.__dict__.setdefault("_flux_pe", {})
        pe = cache.get(key)
        if pe is None:
            pe = FX.rope_table(hl, wl, Nt, axes_dim, theta).to(self.device).contiguous()
            cache[key] = pe
        ctx = context.to(self.device, torch.float32).contiguous()
        t = timestep.to(self.device, torch.float32).contiguous()
        yv = y.to(self.device, torch.float32).contiguous()
        g = guidance.to(self.device, torch.float32).contiguous() if guidance is not None else None
        out = torch.empty_like(img)
        with torch.cuda.device(self.device):
            L.check(self.lib.ldn_flux_forward(self.h, img.data_ptr(), ctx.data_ptr(), pe.data_ptr(), t.data_ptr(),
                                              g.data_ptr() if g is not None else 0, yv.data_ptr(), out.data_ptr(), B,
                                              hl * wl, Nt, L.cur_stream()))
        self._keep = [img, ctx, t, yv, g, pe]
        return FX.unpatchify(out, hl, wl)[:, :, :h, :w]

    def clip_encode(self, ids: torch.Tensor):
        """ids [S,77] int64 -> (penultimate-layer output after final LN, last-layer output after final LN)."""
        if not ids.is_cuda:  # token ids come from the host tokenizer: range-check them before they index device memory
            hi = self.clip_vocab() + getattr(self, "_clip_extra_n", 0)
            if ids.numel() and (int(ids.min()) < 0 or int(ids.max()) >= hi):
                raise L.LdnError(f"CLIP token id outside [0, {hi}) (vocabulary + loaded textual-inversion vectors)")
        ids = ids.to(self.device, torch.int64).contiguous()
        S = ids.shape[0]
        pen = torch.empty(S, 77, 768, device=self.device, dtype=torch.float32)
        last = torch.empty(S, 77, 768, device=self.device, dtype=torch.float32)
        with torch.cuda.device(self.device):
            L.check(self.lib.ldn_clip_encode(self.h, ids.data_ptr(), S, pen.data_ptr(), last.data_ptr(),
                                             L.cur_stream()))
        return pen, last

    def t5_encode(self, ids: torch.Tensor) -> torch.Tensor:
        """ids [S,n] int64 -> final-RMS-norm of the last T5 block's states, [S,n,d_model] fp32 (T5.forward without a mask,
        src/clip/FluxClip.py:457-562)."""
        from . import t5 as T5H

        width = getattr(self, "_t5_width", None)
        if width is None:
            raise L.LdnError("T5 weights not loaded (Engine.load_t5)")
        ids = ids.to(self.device, torch.int64).contiguous()
        S, n = ids.shape
        buckets = T5H.relative_position_buckets(n).to(self.device)
        with torch.cuda.device(self.device):
            out = torch.empty(S, n, width, device=self.device, dtype=torch.float32)
            L.check(self.lib.ldn_t5_encode(self.h, ids.data_ptr(), buckets.data_ptr(), S, n, out.data_ptr(), L.cur_stream()))
        self._keep = [ids, buckets]
        return out

