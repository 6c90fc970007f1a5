"""Flux.1 sampling on the engine: the host-side mirror of the reference's Flux branch of `pipeline()` /
`KSampler.sample` (src/user/pipeline.py:251-264: 20 steps, cfg 1, sampler `euler_cfgpp`, scheduler `beta`) around
`ldn_flux_forward`.

Pieces and their reference counterparts: ModelSamplingFlux (sigma table, timestep = sigma; src/sample/sampling.py:172-218)
-> schedule.FluxSchedule; CONST (noise_scaling sigma * noise + (1 - sigma) * latent, denoised = x - v * sigma; :100-156);
latent format Flux1 (scale 0.3611, shift 0.1159; src/Utilities/Latent.py:114-148); cond/uncond batched in one call with rows
[uncond, cond] (calc_cond_batch, src/cond/cond.py:150-288 -- the cfgpp samplers disable the cfg == 1 shortcut, so both rows
are always evaluated); sampler = sampling.sample_euler_cfgpp.
"""
from __future__ import annotations

from typing import Callable, Dict, Optional, Tuple

import torch

from . import sampling as S
from .engine import Engine
from .schedule import FluxSchedule, calculate_sigmas

LATENT_SCALE = 0.3611
LATENT_SHIFT = 0.1159


def flux_pair_fn(engine: Engine, ctx_neg: torch.Tensor, ctx_pos: torch.Tensor, y_neg: torch.Tensor, y_pos: torch.Tensor,
                 guidance: float) -> Callable:
    """(x, sigma) -> (denoised_uncond, denoised_cond) for B latents: one flux_forward over rows [uncond.., cond..]."""
    def fn(x: torch.Tensor, sigma: float) -> Tuple[torch.Tensor, torch.Tensor]:
        B = x.shape[0]
        dev = x.device
        xx = torch.cat([x, x])
        ctx = torch.cat([ctx_neg.expand(B, -1, -1), ctx_pos.expand(B, -1, -1)]).to(dev)
        y = torch.cat([y_neg.expand(B, -1), y_pos.expand(B, -1)]).to(dev)
        t = torch.full((2 * B,), float(sigma), device=dev)
        g = torch.full((2 * B,), float(guidance), device=dev)
        v = engine.flux_forward(xx, t, ctx, y, g)
        den = xx - v * float(sigma)  # CONST.calculate_denoised
        return den[:B].contiguous(), den[B:].contiguous()
    return fn


def sample_flux(engine: Engine, seed: int, steps: int, positive: Tuple[torch.Tensor, torch.Tensor],
                negative: Tuple[torch.Tensor, torch.Tensor], latent_image: Dict[str, torch.Tensor], cfg: float = 1.0,
                guidance: float = 3.5, scheduler: str = "beta", shift: float = 1.15,
                noise: Optional[torch.Tensor] = None, callback: Optional[Callable] = None,
                interrupt: Optional[Callable[[], bool]] = None) -> Tuple[Dict[str, torch.Tensor]]:
    """positive / negative: (T5 states [1,Nt,4096], pooled CLIP vector [1,768]); latent_image {"samples": [B,16,h,w]}.
    Returns ({"samples": latents in the VAE's space (process_out applied), fp32 on the CPU},)."""
    latent = latent_image["samples"]
    sigmas = calculate_sigmas(FluxSchedule(shift), scheduler, steps)
    if noise is None:
        noise = S.prepare_noise(latent, seed)
    lat = (latent - LATENT_SHIFT) * LATENT_SCALE if torch.count_nonzero(latent) > 0 else latent  # process_in (CFG.py:266-269)
    x = (sigmas[0] * noise + (1.0 - sigmas[0]) * lat).to(engine.device, torch.float32).contiguous()  # CONST.noise_scaling
    fn = flux_pair_fn(engine, negative[0], positive[0], negative[1], positive[1], guidance)
    x = S.sample_euler_cfgpp(engine, x, sigmas, cfg, callback=callback, pair_fn=fn, interrupt=interrupt)
    out = (x / LATENT_SCALE + LATENT_SHIFT).to(torch.float32).cpu()  # process_out
    return ({"samples": out},)
