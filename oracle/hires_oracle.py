"""ORACLE — TEST INFRASTRUCTURE ONLY (see sd15_oracle.py).  HiresFix pieces of the reference path (BASELINE config 5):
LatentUpscale / bislerp (src/Utilities/upscale.py:5-166) and KSampler with denoise < 1 on a non-empty latent
(src/sample/sampling.py:655-675 set_steps; src/sample/CFG.py:266-269 latent scaling; sampling.py:58-83 noise_scaling).
Pinned against tests/golden/hires_small.pt (generated from the reference by tests/golden/make_golden_hires.py)."""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

from . import sd15_oracle as O


def _bilinear_coords(n_old: int, n_new: int):
    """Source indices (left, right) and blend ratio per destination index, as F.interpolate(bilinear) places them."""
    base = torch.arange(n_old, dtype=torch.float32).reshape(1, 1, 1, -1)
    left = F.interpolate(base, size=(1, n_new), mode="bilinear")
    ratio = left - left.floor()
    nxt = base + 1
    nxt[..., -1] -= 1
    right = F.interpolate(nxt, size=(1, n_new), mode="bilinear")
    return left.to(torch.int64).flatten(), right.to(torch.int64).flatten(), ratio.flatten()


def _slerp(a: torch.Tensor, b: torch.Tensor, r: torch.Tensor) -> torch.Tensor:
    """Spherical interpolation of channel vectors a, b [n, c] with ratios r [n, 1] (upscale.py:6-27)."""
    c = a.shape[-1]
    na, nb = a.norm(dim=-1, keepdim=True), b.norm(dim=-1, keepdim=True)
    ua, ub = a / na, b / nb
    ua[na.expand(-1, c) == 0.0] = 0.0
    ub[nb.expand(-1, c) == 0.0] = 0.0
    dot = (ua * ub).sum(1)
    omega = torch.acos(dot)
    so = torch.sin(omega)
    res = (torch.sin((1.0 - r.squeeze(1)) * omega) / so).unsqueeze(1) * ua + (torch.sin(r.squeeze(1) * omega) / so).unsqueeze(1) * ub
    res = res * (na * (1.0 - r) + nb * r).expand(-1, c)
    res[dot > 1 - 1e-5] = a[dot > 1 - 1e-5]
    res[dot < 1e-5 - 1] = (a * (1.0 - r) + b * r)[dot < 1e-5 - 1]
    return res


def bislerp(x: torch.Tensor, width: int, height: int) -> torch.Tensor:
    """[n,c,h,w] -> [n,c,height,width]: slerp along W, then along H."""
    x = x.float()
    n, c, h, w = x.shape
    l, r, t = _bilinear_coords(w, width)
    a = x[..., l].permute(0, 2, 3, 1).reshape(-1, c)
    b = x[..., r].permute(0, 2, 3, 1).reshape(-1, c)
    y = _slerp(a, b, t.repeat(n * h).reshape(-1, 1)).reshape(n, h, width, c).permute(0, 3, 1, 2)
    l, r, t = _bilinear_coords(h, height)
    a = y[:, :, l, :].permute(0, 2, 3, 1).reshape(-1, c)
    b = y[:, :, r, :].permute(0, 2, 3, 1).reshape(-1, c)
    tt = t.reshape(1, -1, 1).expand(n, -1, width).reshape(-1, 1)
    return _slerp(a, b, tt).reshape(n, height, width, c).permute(0, 3, 1, 2)


def latent_upscale(samples: torch.Tensor, width_px: int, height_px: int) -> torch.Tensor:
    return bislerp(samples, max(64, width_px) // 8, max(64, height_px) // 8)


def sigmas_for_denoise(scheduler: str, steps: int, denoise: float) -> torch.Tensor:
    if denoise > 0.9999:
        return O.calculate_sigmas(scheduler, steps)
    return O.calculate_sigmas(scheduler, int(steps / denoise))[-(steps + 1):]


def ksample(sd, seed: int, steps: int, cfg: float, sampler: str, scheduler: str, cond, uncond, latent: torch.Tensor,
            denoise: float = 1.0) -> torch.Tensor:
    """KSampler.sample on a (possibly non-empty) latent with denoise <= 1; returns samples / 0.18215."""
    tables = O.make_sigma_tables()
    sigmas = sigmas_for_denoise(scheduler, steps, denoise)
    noise = O.prepare_noise(latent.shape, seed)
    lat = latent * O.LATENT_SCALE if torch.count_nonzero(latent) > 0 else latent
    maxd = math.isclose(float(tables[0][-1]), float(sigmas[0]), rel_tol=1e-05) or float(sigmas[0]) > float(tables[0][-1])
    x = noise * torch.sqrt(1.0 + sigmas[0] ** 2.0) if maxd else noise * sigmas[0]
    x = x + lat

    def denoise_fn(xx, ss):
        return O.cfg_denoise(lambda a, b, c: O.apply_model(sd, a, b, c, tables), xx, ss, cond, uncond, cfg)

    if sampler == "euler_ancestral_cfgpp":
        x = O.sample_euler_ancestral(denoise_fn, x, sigmas, lambda t: torch.randn_like(t))
    else:
        x = O.sample_dpmpp_2m(denoise_fn, x, sigmas)
    return x / O.LATENT_SCALE
