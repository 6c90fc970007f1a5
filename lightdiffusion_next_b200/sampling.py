"""Sampler loop of the engine — the host-side mirror of the reference's sampling driver for the path that
BASELINE.json measures.

Public surface mirrors `sampling.KSampler().sample(model, seed, steps, cfg, sampler_name, scheduler, positive,
negative, latent_image, denoise)` (src/sample/sampling.py:773-887) and the sampler functions
`sample_dpmpp_2m_cfgpp` / `sample_euler_ancestral_dy_cfg_pp` (src/sample/samplers.py:755-962, 612-740) *as they
execute* (the CFG++ branches are never taken; fact 9: multiscale is on by default for dpmpp_2m).

Per step the host issues exactly two engine calls — `ldn_unet_denoise` (one CUDA graph: the 2B-row UNet forward with
EPS scaling) and `ldn_cfg_step` (CFG lerp + solver update) — and never synchronises: solver coefficients are computed
up front on the CPU in fp32 with the reference's op order.
"""
from __future__ import annotations

from typing import Callable, Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F

from .engine import Engine
from .schedule import calculate_sigmas, get_ancestral_step, max_denoise

LATENT_SCALE = 0.18215  # src/Utilities/Latent.py:41-62
SAMPLERS = ("dpmpp_2m_cfgpp", "euler_ancestral_cfgpp", "dpmpp_sde_cfgpp", "euler_cfgpp")


def prepare_noise(latent: torch.Tensor, seed: int) -> torch.Tensor:
    """ksampler_util.py:274-295: global CPU generator, one draw for the whole batch (so seeds reproduce)."""
    g = torch.manual_seed(seed)
    return torch.randn(latent.size(), dtype=latent.dtype, layout=latent.layout, generator=g, device="cpu")


def _multiscale_fullres(step: int, n_steps: int, start: int = 5, end: int = 8, intermittent: bool = True) -> bool:
    if step < start or step >= n_steps - end:
        return True
    if intermittent:
        return (step - start) % 2 == 0
    return False


class SamplerLoop:
    """Reusable loop state for one (batch, resolution): device buffers are allocated once.  The sample_* loops below
    ping-pong between the tensor they are given and `x_next`, i.e. they CONSUME their `x` argument (sample() and the seam
    functions in backend.py always hand them a private copy)."""

    def __init__(self, engine: Engine, batch: int, lat_h: int, lat_w: int):
        self.e = engine
        dev = engine.device
        self.B = batch
        self.x2 = torch.empty(2 * batch, 4, lat_h, lat_w, device=dev)  # [uncond rows | cond rows] input
        self.sig2 = torch.empty(2 * batch, device=dev)
        self.den2 = torch.empty_like(self.x2)
        self.x_next = torch.empty(batch, 4, lat_h, lat_w, device=dev)

    def denoise_pair(self, x: torch.Tensor, sigma: float) -> Tuple[torch.Tensor, torch.Tensor]:
        """calc_cond_batch (src/cond/cond.py:150-288): one batched call, rows [uncond.., cond..]."""
        B = self.B
        self.x2[:B].copy_(x)
        self.x2[B:].copy_(x)
        self.sig2.fill_(sigma)
        self.e.denoise(self.x2, self.sig2, out=self.den2)
        return self.den2[:B], self.den2[B:]


def sample_dpmpp_2m_cfgpp(engine: Engine, x: torch.Tensor, sigmas: torch.Tensor, cfg: float,
                          enable_multiscale: bool = True, multiscale_factor: float = 0.5,
                          callback: Optional[Callable] = None, multiscale_fullres_start: int = 5,
                          multiscale_fullres_end: int = 8, multiscale_intermittent_fullres: bool = True,
                          interrupt: Optional[Callable[[], bool]] = None) -> torch.Tensor:
    """x <- (sigma_{i+1}/sigma_i) x - expm1(-h_i) * lerp(uncond, cond, cfg)  (first order, as the reference executes).
    The multiscale_* options are the sampler's own keyword arguments (samplers.py:768-773), reachable in the reference
    through `ksampler(name, extra_options=...)`; the defaults are what `KSampler.sample` always runs with."""
    B, _, oh, ow = x.shape
    sh = int(max(8, ((oh * multiscale_factor) // 8) * 8)) if enable_multiscale else oh
    sw = int(max(8, ((ow * multiscale_factor) // 8) * 8)) if enable_multiscale else ow
    active = enable_multiscale and (sh != oh or sw != ow)
    sig = sigmas.float().cpu()
    t = -torch.log(sig)
    sigma_steps = torch.exp(-t)
    ratios = sigma_steps[1:] / sigma_steps[:-1]
    h = t[1:] - t[:-1]
    hexp = torch.expm1(-h)
    n = len(sig) - 1
    full = SamplerLoop(engine, B, oh, ow)
    low = SamplerLoop(engine, B, sh, sw) if active else None
    den = torch.empty_like(x)
    for i in range(n):
        if interrupt is not None and interrupt():
            return x  # the reference's interrupt_flag poll (samplers.py:884-889): hand back the current latent
        if (not active) or _multiscale_fullres(i, n, multiscale_fullres_start, multiscale_fullres_end,
                                               multiscale_intermittent_fullres):
            du, dc = full.denoise_pair(x, float(sig[i]))
            engine.cfg_step(x, du, dc, cfg, 0, c0=float(ratios[i]), c1=float(hexp[i]), x_out=full.x_next,
                            denoised_out=den)
            x, full.x_next = full.x_next, x
        else:
            xp = F.interpolate(x, size=(sh, sw), mode="bilinear", align_corners=False)
            du, dc = low.denoise_pair(xp, float(sig[i]))
            d_low = torch.empty_like(xp)
            engine.cfg_step(None, du, dc, cfg, 2, denoised_out=d_low)
            den = F.interpolate(d_low, size=(oh, ow), mode="bilinear", align_corners=False)
            x = float(ratios[i]) * x - float(hexp[i]) * den
        if callback is not None:
            callback({"x": x, "i": i, "sigma": sig[i], "denoised": den})
    return x


def sample_euler_ancestral_cfgpp(engine: Engine, x: torch.Tensor, sigmas: torch.Tensor, cfg: float,
                                 noise_sampler: Optional[Callable] = None,
                                 callback: Optional[Callable] = None,
                                 interrupt: Optional[Callable[[], bool]] = None, eta: float = 1.0,
                                 s_noise: float = 1.0) -> torch.Tensor:
    """d = (x - D)/sigma; x += d (sigma_down - sigma); x += noise * s_noise * sigma_up, (sigma_down, sigma_up) =
    get_ancestral_step(sigma_i, sigma_{i+1}, eta)   (sample_euler_ancestral_dy_cfg_pp as executed, samplers.py:612-740)."""
    B = x.shape[0]
    sig = sigmas.float().cpu()
    n = len(sig) - 1
    loop = SamplerLoop(engine, B, x.shape[2], x.shape[3])
    den = torch.empty_like(x)
    for i in range(n):
        if interrupt is not None and interrupt():
            return x  # the reference's interrupt_flag poll (samplers.py:884-889): hand back the current latent
        du, dc = loop.denoise_pair(x, float(sig[i]))
        sd, su = get_ancestral_step(sig[i], sig[i + 1], eta)
        noise = None
        if sig[i + 1] > 0:
            # the reference's convention (samplers.py:633-636, 732): noise_sampler(sigma, sigma_next) -> noise like x
            noise = noise_sampler(sig[i], sig[i + 1]).to(x.device) if noise_sampler is not None else torch.randn_like(x)
        engine.cfg_step(x, du, dc, cfg, 1, c0=float(sd - sig[i]), c1=float(su) * s_noise, c2=float(sig[i]), noise=noise,
                        x_out=loop.x_next, denoised_out=den)
        x, loop.x_next = loop.x_next, x
        if callback is not None:
            callback({"x": x, "i": i, "sigma": sig[i], "denoised": den})
    return x


def sample_euler_cfgpp(engine: Engine, x: torch.Tensor, sigmas: torch.Tensor, cfg: float, cfg_scale: float = 7.5,
                       cfg_min: float = 1.0, callback: Optional[Callable] = None,
                       pair_fn: Optional[Callable] = None, interrupt: Optional[Callable[[], bool]] = None) -> torch.Tensor:
    """`euler_cfgpp` = sample_euler_dy_cfg_pp as the reference executes it (samplers.py:470-608, the sampler of its Flux
    pipeline, also selectable for SD1.5): a plain Euler step on the guider's CFG result (the sampler's own CFG++ bookkeeping
    is reset to None every step, :548-550), plus -- while i // 2 == 1 and sigma_{i+1} > 0 -- the dynamic step
    dy_sampling_step_cfg_pp (:362-466): the (1,1) pixel of every 2x2 block is denoised again at half resolution, at sigma_i,
    with the CFG result extrapolated once more from the true uncond by current_cfg = cfg_scale + (cfg_min - cfg_scale) i / n
    (the sampler's own default cfg_scale = 7.5, not the user's cfg), i.e. uncond + cfg * current_cfg * (cond - uncond).
    pair_fn(x, sigma) -> (denoised_uncond, denoised_cond) replaces the SD1.5 UNet call (used by the Flux path)."""
    B, _, H, W = x.shape
    sig = sigmas.float().cpu()
    n = len(sig) - 1
    loop = SamplerLoop(engine, B, H, W) if pair_fn is None else None
    half = None
    x_next = torch.empty_like(x)
    den = torch.empty_like(x)
    for i in range(n):
        if interrupt is not None and interrupt():
            return x  # the reference's interrupt_flag poll (samplers.py:884-889): hand back the current latent
        du, dc = loop.denoise_pair(x, float(sig[i])) if pair_fn is None else pair_fn(x, float(sig[i]))
        engine.cfg_step(x, du, dc, cfg, 1, c0=float(sig[i + 1] - sig[i]), c1=0.0, c2=float(sig[i]), noise=None,
                        x_out=x_next, denoised_out=den)
        x, x_next = x_next, x
        if callback is not None:
            callback({"x": x, "i": i, "sigma": sig[i], "denoised": den})
        if sig[i + 1] > 0 and i // 2 == 1:
            m, k = H // 2, W // 2
            sub = x[:, :, 1:2 * m:2, 1:2 * k:2].contiguous()
            if pair_fn is None:
                if half is None:
                    half = SamplerLoop(engine, B, m, k)
                du2, dc2 = half.denoise_pair(sub, float(sig[i]))
            else:
                du2, dc2 = pair_fn(sub, float(sig[i]))
            current_cfg = cfg_scale + (cfg_min - cfg_scale) * (i / n)
            sub_next = torch.empty_like(sub)
            engine.cfg_step(sub, du2, dc2, cfg * current_cfg, 1, c0=float(sig[i + 1] - sig[i]), c1=0.0, c2=float(sig[i]),
                            noise=None, x_out=sub_next, denoised_out=None)
            x[:, :, 1:2 * m:2, 1:2 * k:2] = sub_next
    return x


class BrownianIntervalNoise:
    """Default noise source of dpmpp_sde when none is injected: increments of one Brownian path W over sigma
    ("time" = sigma, as BrownianTreeNoiseSampler uses it, src/sample/sampling_util.py:239-287), normalised by
    sqrt(|interval|).  The two queries of a step, (sigma_i, sigma_s) and (sigma_i, sigma_{i+1}), overlap; their
    increments are built from the same two independent pieces so they are correlated as on a true path.  Drawn on the
    device; statistically equivalent to the reference's CPU Brownian tree, not sample-identical (inject `noise_sampler`
    for that)."""

    def __init__(self, x: torch.Tensor, seed: Optional[int] = None):
        self.shape, self.device = x.shape, x.device
        self.gen = torch.Generator(device=x.device)
        self.gen.manual_seed(0 if seed is None else int(seed))
        self._left = None   # (sigma_hi, sigma_mid, increment over [mid, hi])

    def _draw(self, var: float) -> torch.Tensor:
        return torch.randn(self.shape, generator=self.gen, device=self.device) * (var ** 0.5)

    def __call__(self, sigma, sigma_next) -> torch.Tensor:
        hi, lo = float(sigma), float(sigma_next)
        if self._left is not None and abs(self._left[0] - hi) < 1e-12 and lo < self._left[1]:
            inc = self._left[2] + self._draw(self._left[1] - lo)   # extend the path from sigma_s down to sigma_{i+1}
            self._left = None
        else:
            inc = self._draw(hi - lo)
            self._left = (hi, lo, inc)
        return inc / ((hi - lo) ** 0.5)


def sample_dpmpp_sde_cfgpp(engine: Engine, x: torch.Tensor, sigmas: torch.Tensor, cfg: float,
                           noise_sampler: Optional[Callable] = None, seed: Optional[int] = None, eta: float = 1.0,
                           r: float = 0.5, enable_multiscale: bool = False, multiscale_factor: float = 0.5,
                           callback: Optional[Callable] = None,
                           interrupt: Optional[Callable[[], bool]] = None, s_noise: float = 1.0,
                           multiscale_fullres_start: int = 5, multiscale_fullres_end: int = 8,
                           multiscale_intermittent_fullres: bool = False) -> torch.Tensor:
    """DPM-Solver++ (SDE) as the reference executes it (samplers.py:966-1254): two CFG-batched UNet evaluations per
    step (at sigma_i and at the midpoint in log-sigma), ancestral noise from `noise_sampler(sigma, sigma_next)`.  The
    keyword defaults are the reference sampler's own (:973-989: multiscale margins 5 / 8, no intermittent full-res steps)."""
    B, _, oh, ow = x.shape
    sh = int(max(8, ((oh * multiscale_factor) // 8) * 8)) if enable_multiscale else oh
    sw = int(max(8, ((ow * multiscale_factor) // 8) * 8)) if enable_multiscale else ow
    active = enable_multiscale and (sh != oh or sw != ow)
    sig = sigmas.float().cpu()
    n = len(sig) - 1
    if noise_sampler is None:
        noise_sampler = BrownianIntervalNoise(x, seed)
    full = SamplerLoop(engine, B, oh, ow)
    low = SamplerLoop(engine, B, sh, sw) if active else None
    sigma_fn = lambda t: t.neg().exp()
    t_fn = lambda s: s.log().neg()

    def denoised_at(xx: torch.Tensor, sigma: float, fullres: bool) -> torch.Tensor:
        loop = full if fullres else low
        xin = xx if fullres else F.interpolate(xx, size=(sh, sw), mode="bilinear", align_corners=False)
        du, dc = loop.denoise_pair(xin, sigma)
        d = torch.empty_like(xin)
        engine.cfg_step(None, du, dc, cfg, 2, denoised_out=d)
        return d if fullres else F.interpolate(d, size=(oh, ow), mode="bilinear", align_corners=False)

    for i in range(n):
        if interrupt is not None and interrupt():
            return x  # the reference's interrupt_flag poll (samplers.py:884-889): hand back the current latent
        fullres = (not active) or _multiscale_fullres(i, n, multiscale_fullres_start, multiscale_fullres_end,
                                                      multiscale_intermittent_fullres)
        den = denoised_at(x, float(sig[i]), fullres)
        if sig[i + 1] == 0:
            x = x + (x - den) / float(sig[i]) * float(sig[i + 1] - sig[i])
        else:
            t, t_next = t_fn(sig[i]), t_fn(sig[i + 1])
            s = t + (t_next - t) * r
            sd, su = get_ancestral_step(sigma_fn(t), sigma_fn(s), eta)
            s_ = t_fn(sd)
            n1 = noise_sampler(sigma_fn(t), sigma_fn(s)).to(x.device)
            x_2 = float(sigma_fn(s_) / sigma_fn(t)) * x - float((t - s_).expm1()) * den + n1 * (float(su) * s_noise)
            den_2 = denoised_at(x_2, float(sigma_fn(s)), fullres)
            sd, su = get_ancestral_step(sigma_fn(t), sigma_fn(t_next), eta)
            t_next_ = t_fn(sd)
            d_mix = (1 - 1 / (2 * r)) * den + (1 / (2 * r)) * den_2
            n2 = noise_sampler(sigma_fn(t), sigma_fn(t_next)).to(x.device)
            x = float(sigma_fn(t_next_) / sigma_fn(t)) * x - float((t - t_next_).expm1()) * d_mix + n2 * (float(su) * s_noise)
        if callback is not None:
            callback({"x": x, "i": i, "sigma": sig[i], "denoised": den})
    return x


def set_contexts(engine: Engine, positive: torch.Tensor, negative: torch.Tensor, batch: int) -> None:
    """Uploads the cross-attention contexts of a CFG pair for `batch` images: rows [uncond.., cond..] (cond.py:194);
    contexts of unequal token length are tiled to their least common multiple before batching (cond.py:100-126)."""
    tn, tp = negative.shape[1], positive.shape[1]
    if tn != tp:
        import math
        lcm = tn * tp // math.gcd(tn, tp)
        negative = negative.repeat(1, lcm // tn, 1)
        positive = positive.repeat(1, lcm // tp, 1)
    ctx = torch.cat([negative.expand(batch, -1, -1), positive.expand(batch, -1, -1)]).to(engine.device)
    engine.set_context(ctx)


def sample(engine: Engine, seed: int, steps: int, cfg: float, sampler_name: str, scheduler: str,
           positive: torch.Tensor, negative: torch.Tensor, latent_image: Dict[str, torch.Tensor],
           denoise: float = 1.0, enable_multiscale: bool = True, noise: Optional[torch.Tensor] = None,
           callback: Optional[Callable] = None, noise_sampler: Optional[Callable] = None,
           sampler_options: Optional[Dict[str, object]] = None,
           interrupt: Optional[Callable[[], bool]] = None) -> Tuple[Dict[str, torch.Tensor]]:
    """Drop-in for KSampler.sample on the measured path. positive / negative: [1 or B, 77k, 768] conditioning tensors.
    sampler_options: the `extra_options` of the reference's `ksampler(name, extra_options)` seam (sampling.py:500-534) for
    dpmpp_2m_cfgpp: multiscale_factor / multiscale_fullres_start / multiscale_fullres_end / multiscale_intermittent_fullres.
    interrupt: polled before every step like the reference polls app.interrupt_flag (samplers.py:884-889); when it returns
    True the loop stops and the current latent is returned.
    Returns ({"samples": latents / 0.18215 on the CPU},) like the reference node."""
    if sampler_name not in SAMPLERS:
        raise ValueError(f"sampler {sampler_name!r} is not built (have {SAMPLERS})")
    latent = latent_image["samples"]
    B = latent.shape[0]
    dev = engine.device
    if denoise is None or denoise > 0.9999:
        sigmas = calculate_sigmas(engine.schedule, scheduler, steps)
    else:
        # KSampler1.set_steps (sampling.py:655-675): the tail of a longer schedule (HiresFix second pass, img2img)
        if denoise <= 0.0:
            return ({"samples": latent.clone()},)
        sigmas = calculate_sigmas(engine.schedule, scheduler, int(steps / denoise))[-(steps + 1):]
    if noise is None:
        noise = prepare_noise(latent, seed)
    lat = latent * LATENT_SCALE if torch.count_nonzero(latent) > 0 else latent  # CFG.py:266-269
    if max_denoise(engine.schedule, sigmas):
        x = noise * torch.sqrt(1.0 + sigmas[0] ** 2.0)
    else:
        x = noise * sigmas[0]
    x = (x + lat).to(dev, torch.float32).contiguous()
    set_contexts(engine, positive, negative, B)
    if sampler_name == "dpmpp_2m_cfgpp":
        opts = dict(sampler_options or {})
        unknown = set(opts) - {"multiscale_factor", "multiscale_fullres_start", "multiscale_fullres_end", "multiscale_intermittent_fullres"}
        if unknown:
            raise ValueError(f"unknown dpmpp_2m_cfgpp options {sorted(unknown)}")
        x = sample_dpmpp_2m_cfgpp(engine, x, sigmas, cfg, enable_multiscale=enable_multiscale, callback=callback,
                                  interrupt=interrupt, **opts)
    elif sampler_name == "dpmpp_sde_cfgpp":
        x = sample_dpmpp_sde_cfgpp(engine, x, sigmas, cfg, noise_sampler=noise_sampler, seed=seed,
                                   enable_multiscale=enable_multiscale, callback=callback, interrupt=interrupt)
    elif sampler_name == "euler_cfgpp":
        x = sample_euler_cfgpp(engine, x, sigmas, cfg, callback=callback, interrupt=interrupt)
    else:
        x = sample_euler_ancestral_cfgpp(engine, x, sigmas, cfg, noise_sampler=noise_sampler, callback=callback,
                                         interrupt=interrupt)
    out = (x / LATENT_SCALE).to(torch.float32).cpu()
    return ({"samples": out},)
