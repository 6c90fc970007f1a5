"""Latent-space helpers around the sampler: LatentUpscale for HiresFix (BASELINE config 5).

`latent_upscale` mirrors `LatentUpscale.upscale` / `bislerp` (src/Utilities/upscale.py:5-166): a separable resize whose
two-tap blend is a *spherical* interpolation of the 4-channel latent vectors (norms blended linearly), with the tap
positions of bilinear resampling.  One-off host-side work on a few KB (the reference also runs it on the CPU)."""
from __future__ import annotations

import torch
import torch.nn.functional as F


def _taps(n_old: int, n_new: int):
    pos = torch.arange(n_old, dtype=torch.float32).view(1, 1, 1, n_old)
    lo = F.interpolate(pos, size=(1, n_new), mode="bilinear")
    frac = (lo - lo.floor()).flatten()
    hi_src = pos + 1
    hi_src[..., n_old - 1] = n_old - 1
    hi = F.interpolate(hi_src, size=(1, n_new), mode="bilinear")
    return lo.flatten().long(), hi.flatten().long(), frac


def _slerp_last_axis(x: torch.Tensor, n_new: int) -> torch.Tensor:
    """x [..., C, L] -> [..., C, n_new]; channel vectors (dim -2) are slerped between neighbouring positions of dim -1."""
    lo, hi, frac = _taps(x.shape[-1], n_new)
    a, b = x[..., lo], x[..., hi]                      # [..., C, n_new]
    na, nb = a.norm(dim=-2, keepdim=True), b.norm(dim=-2, keepdim=True)
    ua = torch.where(na == 0, torch.zeros_like(a), a / na)
    ub = torch.where(nb == 0, torch.zeros_like(b), b / nb)
    dot = (ua * ub).sum(dim=-2, keepdim=True)
    omega = torch.acos(dot)
    so = torch.sin(omega)
    r = frac.view(*([1] * (x.dim() - 1)), n_new)
    out = (torch.sin((1.0 - r) * omega) / so) * ua + (torch.sin(r * omega) / so) * ub
    out = out * (na * (1.0 - r) + nb * r)
    out = torch.where(dot > 1 - 1e-5, a, out)
    out = torch.where(dot < 1e-5 - 1, a * (1.0 - r) + b * r, out)
    return out


def bislerp(samples: torch.Tensor, width: int, height: int) -> torch.Tensor:
    x = samples.float()                                              # [n, c, h, w]
    x = _slerp_last_axis(x.permute(0, 2, 1, 3), width)               # [n, h, c, W]
    x = _slerp_last_axis(x.permute(0, 3, 2, 1), height)              # [n, W, c, H]
    return x.permute(0, 2, 3, 1).contiguous().to(samples.dtype)      # [n, c, H, W]


def latent_upscale(latent: dict, width: int, height: int, engine=None) -> dict:
    """LatentUpscale node: pixel sizes in, latent dict out (sizes clamped to >= 64 px like the reference).  With an engine
    and device-resident samples the resize runs on the engine's bislerp kernel (`ldn_bislerp`), otherwise on the host
    restatement above (which is what the reference itself does with its CPU-resident latents)."""
    if width == 0 and height == 0:
        return latent
    out = dict(latent)
    s = latent["samples"]
    if engine is not None and s.is_cuda:
        out["samples"] = engine.bislerp(s, max(64, width) // 8, max(64, height) // 8)
    else:
        out["samples"] = bislerp(s, max(64, width) // 8, max(64, height) // 8)
    return out
