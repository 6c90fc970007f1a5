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
        self._ctx_ref = None      # (tensor object, version) of the last context uploaded
        self.calls = 0

    def _set_context(self, ctx: torch.Tensor) -> None:
        key = (id(ctx), ctx._version, tuple(ctx.shape), ctx.data_ptr())
        if self._ctx_ref is not None and self._ctx_ref[0] == key and self._ctx_ref[1] is ctx:
            return
        self.engine.set_context(ctx)
        self._ctx_ref = (key, ctx)

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


def unet_state_dict_from_model(base_model) -> Dict[str, torch.Tensor]:
    """Weights stay owned by BaseModel.diffusion_model; the engine keeps its own repacked bf16 copy."""
    return {k: v for k, v in base_model.diffusion_model.state_dict().items()}


def install(model_patcher, engine: Optional[Engine] = None, max_rows: int = 2, max_h: int = 128, max_w: int = 128,
            max_ctx_tokens: int = 77):
    """Returns a clone of `model_patcher` whose UNet step runs on the B200 engine (the reference's
    ApplyStableFastUnet.apply_stable_fast pattern)."""
    if engine is None:
        engine = Engine(max_rows=max_rows, max_h=max_h, max_w=max_w, max_ctx_tokens=max_ctx_tokens)
        engine.load_unet(unet_state_dict_from_model(model_patcher.model))
    m = model_patcher.clone()
    m.set_model_unet_function_wrapper(EngineWrapper(engine))
    return m
