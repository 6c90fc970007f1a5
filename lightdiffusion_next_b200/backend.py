"""Drop-in boundary: the reference's own plugin seam for the UNet step.

LightDiffusion-Next consults `model_options["model_function_wrapper"]` in calc_cond_batch (src/cond/cond.py:254-265);
it is installed with `ModelPatcher.set_model_unet_function_wrapper` (src/Model/ModelPatcher.py:138-144) on a clone,
exactly like its Stable-Fast integration (src/StableFast/StableFast.py:230-274).  `EngineWrapper` is that callable:

    fn(model_function, params) -> denoised
      params["input"]      [k*B, 4, h, w] fp32 on the load device
      params["timestep"]   [k*B] fp32 — sigma repeated per row (NOT a timestep index)
      params["c"]          {"c_crossattn": [k*B, 77m, 768] fp32, "transformer_options": {...}}
      params["cond_or_uncond"]  e.g. [1, 0]  (rows: uncond first)
    returns x - eps*sigma as fp32 [k*B, 4, h, w], rows in the same order (the caller chunks and applies CFG itself).

`model_function` (BaseModel.apply_model) is never called: the engine replaces it.  Nothing here falls back to it.
"""
from __future__ import annotations

from typing import Any, Dict, Optional

import torch

from .engine import Engine

UNET_PREFIX = "model.diffusion_model."


class EngineWrapper:
    def __init__(self, engine: Engine):
        self.engine = engine
        self._ctx_ident = None    # identity of the last tensor object seen (fast path)
        self._ctx_copy = None     # device copy of the context the engine currently holds
        self._ctx_epoch = None
        self.calls = 0

    def _set_context(self, ctx: torch.Tensor) -> None:
        """ldn_set_context (pad + 32 K/V projection GEMMs) only when the conditioning really changed.  calc_cond_batch
        builds `c_crossattn` with a fresh torch.cat on EVERY step (CONDCrossAttn.concat, src/cond/cond.py:100-126, 219-226), so
        tensor identity never repeats through this seam: the cache is keyed on CONTENT -- one 473 KB device compare per
        step against the copy uploaded last (the reference's own loop already synchronises every step, samplers.py:928).
        A UNet weight reload (Engine.weights_epoch) invalidates it, because the engine then drops its K/V buffers."""
        eng = self.engine
        epoch = eng.weights_epoch.get(0, 0)
        try:
            version = ctx._version
        except RuntimeError:   # inference tensors (the reference samples under torch.inference_mode) track no version:
            version = None     # an in-place change would go unseen, so identity alone proves nothing -> compare content
        ident = (id(ctx), version, ctx.data_ptr(), tuple(ctx.shape))
        if self._ctx_copy is not None and self._ctx_epoch == epoch:
            if version is not None and ident == self._ctx_ident[0] and self._ctx_ident[1] is ctx:
                return
            c = ctx.to(eng.device, torch.float32)
            if c.shape == self._ctx_copy.shape and torch.equal(c, self._ctx_copy):
                self._ctx_ident = (ident, ctx)
                return
        c = ctx.detach().to(eng.device, torch.float32).clone()
        eng.set_context(c)
        self._ctx_copy, self._ctx_ident, self._ctx_epoch = c, (ident, ctx), epoch

    def __call__(self, model_function, params: Dict[str, Any]) -> torch.Tensor:
        x = params["input"]
        sigma = params["timestep"]
        c = params["c"]
        for k in c:
            if k not in ("c_crossattn", "transformer_options") and c[k] is not None:
                raise NotImplementedError(f"conditioning {k!r} is not supported by the B200 engine (SD1.5 txt2img path)")
        ctx = c["c_crossattn"]
        if ctx.shape[0] != x.shape[0]:
            raise ValueError("c_crossattn rows must match input rows")
        dev = self.engine.device
        x = x.to(dev, torch.float32).contiguous()
        self._set_context(ctx)
        self.calls += 1
        return self.engine.denoise(x, sigma.to(dev, torch.float32))

    def to(self, arg):
        # ModelPatcher.model_patches_to calls .to(device) and .to(dtype) on load/unload and stores the result
        # (src/Model/ModelPatcher.py:165-175): the engine owns its own weights, so both are no-ops.
        return self


def unet_state_dict_from_model(base_model, model_patcher=None) -> Dict[str, torch.Tensor]:
    """Weights stay owned by BaseModel.diffusion_model; the engine keeps its own repacked bf16 copy.  Weight patches queued
    on the ModelPatcher (LoRAs -- the reference's default pipeline loads one, src/user/pipeline.py:283-291) are what the
    reference would apply in patch_model before running the UNet (ModelPatcher.py:267-300, 515-548): they are folded into the
    copy with the patcher's OWN calculate_weight, on an fp32 temporary rounded once to the storage dtype, exactly as
    patch_weight_to_device does.  The reference's module is left untouched."""
    return patched_state_dict(base_model.diffusion_model.state_dict(), model_patcher, "diffusion_model.")


def patched_state_dict(state_dict, model_patcher, prefix: str) -> Dict[str, torch.Tensor]:
    """A copy of `state_dict` with the patches a reference ModelPatcher has queued for keys `prefix + name` applied the way
    patch_weight_to_device applies them (fp32 temporary, the patcher's own calculate_weight, one rounding)."""
    sd = {k: v for k, v in state_dict.items()}
    patches = getattr(model_patcher, "patches", None) if model_patcher is not None else None
    # If the patcher is currently PATCHED (the reference already sampled with this model: patch_model has run and the live
    # module holds W + LoRA), the unpatched weights are in `backup` (ModelPatcher.py:267-300 saves them before patching);
    # start from those, otherwise the patches would be applied twice.
    backup = getattr(model_patcher, "backup", None) or {}
    if patches:
        for key, plist in patches.items():
            k = key[len(prefix):] if key.startswith(prefix) else None
            if k is not None and k in sd:
                w = backup[key] if key in backup else sd[k]
                sd[k] = model_patcher.calculate_weight(plist, w.to(torch.float32, copy=True), key).to(w.dtype)
    return sd


def clip_state_dict_from_clip(clip) -> Dict[str, torch.Tensor]:
    """CLIP-L text-model weights of a reference `CLIP` object (src/clip/Clip.py:297-404) for `Engine.load_clip`, with the
    patches queued on `clip.patcher` (the CLIP half of a LoRA, keys `clip_l.transformer.text_model.*`, LoRas.py:60-84) folded
    in like the UNet's."""
    tm = clip.cond_stage_model.clip_l.transformer.text_model
    sd = patched_state_dict(tm.state_dict(), getattr(clip, "patcher", None), "clip_l.transformer.text_model.")
    sd.pop("embeddings.position_ids", None)
    return sd


def install(model_patcher, engine: Optional[Engine] = None, max_rows: int = 2, max_h: int = 128, max_w: int = 128,
            max_ctx_tokens: int = 77):
    """Returns a clone of `model_patcher` whose UNet step runs on the B200 engine (the reference's
    ApplyStableFastUnet.apply_stable_fast pattern)."""
    if engine is None:
        engine = Engine(max_rows=max_rows, max_h=max_h, max_w=max_w, max_ctx_tokens=max_ctx_tokens)
        engine.load_unet(unet_state_dict_from_model(model_patcher.model, model_patcher))
    m = model_patcher.clone()
    m.set_model_unet_function_wrapper(EngineWrapper(engine))
    return m


class EngineVAE:
    """Coarser seam: stands in for the reference's `VAE` object (src/AutoEncoders/VariationalAE.py:570-760) wherever nodes
    call `vae.decode(samples)` / `vae.encode(pixels)` (VAEDecode / VAEEncode, :771-801).  Same arguments, layouts and
    devices: decode takes latents [B,C,h,w] (4 channels, or 16 with the Flux autoencoder loaded) and returns [B,8h,8w,3] fp32 in
    [0,1] on the CPU; encode takes [B,H,W,3] in [0,1] and returns a posterior sample [B,C,H/8,W/8] fp32 on the CPU.  `flux`
    is accepted like the reference's keyword; which autoencoder runs is decided by the weights that were loaded."""

    def __init__(self, engine: Engine):
        self.engine = engine

    def decode(self, samples_in: torch.Tensor, flux: bool = False) -> torch.Tensor:
        return self.engine.vae_decode(samples_in).cpu()

    def encode(self, pixel_samples: torch.Tensor, flux: bool = False) -> torch.Tensor:
        return self.engine.vae_encode(pixel_samples)


class EngineCLIP:
    """Coarser seam: stands in for the reference's `CLIP` object (src/clip/Clip.py:297-446) for `clip.tokenize(text)` +
    `clip.encode_from_tokens(tokens, return_pooled, return_dict, flux_enabled)` (CLIPTextEncode :574-589, CLIPTextEncodeFlux
    src/Quantize/Quantizer.py:960-990).  Tokenisation is delegated to the reference's own tokenizer object (host Python with
    its vocabulary files); everything after the token ids runs on the engine.  SD1.5: tokens = list of 77-long (id, weight)
    rows -> cond [1, 77k, 768] (layer -2 + final LN, as CLIPSetLastLayer(-2) configures the reference's pipeline).  Flux:
    tokens = {"l": rows, "t5xxl": rows} -> cond = T5 states, pooled = CLIP-L's pooled vector (FluxClip.py:704-718)."""

    def __init__(self, engine: Engine, tokenizer=None):
        from .pipeline import Pipeline

        self.engine = engine
        self.tokenizer = tokenizer
        self._pipe = Pipeline(engine)

    def tokenize(self, text: str, return_word_ids: bool = False):
        if self.tokenizer is None:
            raise RuntimeError("EngineCLIP was built without a tokenizer (pass the reference's SD1Tokenizer / FluxTokenizer)")
        return self.tokenizer.tokenize_with_weights(text, return_word_ids)

    def encode_from_tokens(self, tokens, return_pooled: bool = False, return_dict: bool = False, flux_enabled: bool = False):
        if isinstance(tokens, dict) and "t5xxl" in tokens:
            from . import t5 as T5H

            cond = T5H.encode_token_weights(self.engine, tokens["t5xxl"])
            _, pooled = self._pipe.encode(tokens["l"], return_pooled=True)
        else:
            rows = tokens["l"] if isinstance(tokens, dict) else tokens
            cond, pooled = self._pipe.encode(rows, return_pooled=True)
        cond, pooled = cond.cpu(), pooled.cpu()  # Device.intermediate_device()
        if return_dict:
            return {"cond": cond, "pooled_output": pooled}
        if return_pooled:
            return cond, pooled
        return cond


def engine_sampler_function(engine: Engine, sampler_name: str, interrupt=None):
    """Coarser seam: a sampler function for the reference's registry -- what `sampling.ksampler(name)` wraps in `KSAMPLER`
    (src/sample/sampling.py:500-534) and `KSAMPLER.sample` calls as
    `fn(model_k, x, sigmas, extra_args=, callback=, disable=, pipeline=, **extra_options) -> x` (:445-497).  `model_k` is the
    reference's `KSamplerX0Inpaint` around its `CFGGuider`; the contexts and the cfg scale are read off the guider
    (`.conds["positive" / "negative"][0]["model_conds"]["c_crossattn"].cond`, `.cfg`; CFG.py:164-235, cond.py:74-147) and
    the whole loop -- model calls, CFG, solver update -- then runs on the engine instead of calling `model_k` per step.
    x arrives already noise-scaled on the load device and is returned in the model's latent space, as the reference's
    samplers do.  `interrupt`: a callable polled before every step unless the call is made with pipeline=True -- pass
    `lambda: app_instance.app.interrupt_flag` to keep the reference's cancel button working (samplers.py:884-889).
    Use: `KSAMPLER(engine_sampler_function(engine, "dpmpp_2m_cfgpp"), extra_options)`."""
    from . import sampling as S

    if sampler_name not in S.SAMPLERS:
        raise ValueError(f"sampler {sampler_name!r} is not built (have {S.SAMPLERS})")

    def _ctx(conds) -> torch.Tensor:
        if conds is None or len(conds) != 1:
            raise NotImplementedError("the B200 engine samples one positive and one negative conditioning (no areas / masks)")
        c = conds[0]["model_conds"]["c_crossattn"]
        return c.cond if hasattr(c, "cond") else c

    def fn(model, x, sigmas, extra_args=None, callback=None, disable=None, pipeline=False, **extra_options):
        guider = model.inner_model
        dev = engine.device
        S.set_contexts(engine, _ctx(guider.conds.get("positive")), _ctx(guider.conds.get("negative")), x.shape[0])
        xs = x.detach().to(dev, torch.float32, copy=True).contiguous()  # the loops ping-pong between buffers: never the caller's
        cfg = float(guider.cfg)
        opts = dict(extra_options)
        allowed = set(S.SAMPLER_OPTIONS[sampler_name])
        if sampler_name == "dpmpp_sde_cfgpp":
            opts.setdefault("seed", (extra_args or {}).get("seed"))
            allowed.add("seed")
        unknown = set(opts) - allowed
        if unknown:
            raise ValueError(f"unknown {sampler_name} options {sorted(unknown)}")
        run = {"dpmpp_2m_cfgpp": S.sample_dpmpp_2m_cfgpp, "dpmpp_sde_cfgpp": S.sample_dpmpp_sde_cfgpp,
               "euler_ancestral_cfgpp": S.sample_euler_ancestral_cfgpp, "euler_cfgpp": S.sample_euler_cfgpp}[sampler_name]
        return run(engine, xs, sigmas, cfg, callback=callback, interrupt=None if pipeline else interrupt, **opts).to(x.device)

    return fn
