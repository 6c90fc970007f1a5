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

from .engine import Engine
from .schedule import calculate_sigmas, get_ancestral_step, max_denoise

LATENT_SCALE = 0.18215  # src/Utilities/Latent.py:41-62
SAMPLERS = ("dpmpp_2m_cfgpp", "euler_ancestral_cfgpp", "dpmpp_sde_cfgpp", "euler_cfgpp")
_MS = {"enable_multiscale", "multiscale_factor", "multiscale_fullres_start", "multiscale_fullres_end",
       "multiscale_intermittent_fullres"}
# keyword arguments of the reference's sampler functions that this engine honours (samplers.py:755-773, 612-631, 966-989,
# 470-489); everything else is rejected loudly rather than silently dropped
SAMPLER_OPTIONS = {"dpmpp_2m_cfgpp": set(_MS), "dpmpp_sde_cfgpp": _MS | {"noise_sampler", "eta", "r", "s_noise"},
                   "euler_ancestral_cfgpp": {"noise_sampler", "eta", "s_noise"}, "euler_cfgpp": {"cfg_scale", "cfg_min"}}
# what KSampler.sample -> common_ksampler -> sample1 always passes for dpmpp_sde_cfgpp (sampling.py:795-799, 949-964)
KSAMPLER_SDE_MULTISCALE = {"multiscale_factor": 0.5, "multiscale_fullres_start": 3, "multiscale_fullres_end": 8,
                           "multiscale_intermittent_fullres": False}


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
            # half-resolution step (samplers.py:821-835, 905-930): bilinear down, denoise, bilinear up, same update
            xp = engine.resample_bilinear(x, (sh, sw))
            du, dc = low.denoise_pair(xp, float(sig[i]))
            d_low = torch.empty_like(xp)
            engine.cfg_step(None, du, dc, cfg, 2, denoised_out=d_low)
            den = engine.resample_bilinear(d_low, (oh, ow), out=den)
            engine.cfg_step(x, den, den, 1.0, 0, c0=float(ratios[i]), c1=float(hexp[i]), x_out=full.x_next)
            x, full.x_next = full.x_next, x
        if callback is not None:
            callback({"x": x, "i": i, "sigma": sig[i], "denoised": den})
    return x


def default_noise_sampler(x: torch.Tensor, batch_slice: Optional[Tuple[int, int, int]] = None) -> Callable:
    """sampling_util.default_noise_sampler (src/sample/sampling_util.py:154-165): torch.randn_like(x) from the device's
    global generator (seeded by prepare_noise's torch.manual_seed, as in the reference).  batch_slice = (lo, hi, total):
    this process holds images [lo, hi) of a batch of `total` -- the draw is made for the WHOLE batch and sliced, so that
    every rank of a sharded run consumes the generator exactly like the single-process batch does (each rank must have
    been seeded alike: distributed.sample_sharded calls torch.manual_seed(seed) on every rank)."""
    if batch_slice is None:
        return lambda sigma, sigma_next: torch.randn_like(x)
    lo, hi, total = batch_slice
    shape = (total,) + tuple(x.shape[1:])
    return lambda sigma, sigma_next: torch.randn(shape, dtype=x.dtype, device=x.device)[lo:hi]


def sample_euler_ancestral_cfgpp(engine: Engine, x: torch.Tensor, sigmas: torch.Tensor, cfg: float,
                                 noise_sampler: Optional[Callable] = None,
                                 callback: Optional[Callable] = None,
                                 interrupt: Optional[Callable[[], bool]] = None, eta: float = 1.0,
                                 s_noise: float = 1.0,
                                 batch_slice: Optional[Tuple[int, int, int]] = None) -> torch.Tensor:
    """d = (x - D)/sigma; x += d (sigma_down - sigma); x += noise * s_noise * sigma_up, (sigma_down, sigma_up) =
    get_ancestral_step(sigma_i, sigma_{i+1}, eta)   (sample_euler_ancestral_dy_cfg_pp as executed, samplers.py:612-740)."""
    B = x.shape[0]
    sig = sigmas.float().cpu()
    n = len(sig) - 1
    if noise_sampler is None:
        noise_sampler = default_noise_sampler(x, batch_slice)
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
            noise = noise_sampler(sig[i], sig[i + 1]).to(x.device, torch.float32).contiguous()
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

    def __init__(self, x: torch.Tensor, seed: Optional[int] = None, batch_slice: Optional[Tuple[int, int, int]] = None):
        """batch_slice = (lo, hi, total): x holds images [lo, hi) of a batch of `total`; every draw is made for the whole
        batch from the same seeded generator and sliced, so the shards of a multi-GPU run get DIFFERENT, and the same,
        noise as the rows of the single-process batch (one path per image)."""
        self.lo, self.hi, total = batch_slice if batch_slice is not None else (0, x.shape[0], x.shape[0])
        self.shape, self.device = (total,) + tuple(x.shape[1:]), x.device
        self.gen = torch.Generator(device=x.device)
        self.gen.manual_seed(0 if seed is None else int(seed))
        self._left = None   # (sigma_hi, sigma_mid, increment over [mid, hi])

    def _draw(self, var: float) -> torch.Tensor:
        return torch.randn(self.shape, generator=self.gen, device=self.device)[self.lo:self.hi] * (var ** 0.5)

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
                           multiscale_intermittent_fullres: bool = False,
                           batch_slice: Optional[Tuple[int, int, int]] = None) -> torch.Tensor:
    """DPM-Solver++ (SDE) as the reference executes it (samplers.py:966-1254): two CFG-batched UNet evaluations per
    step (at sigma_i and at the midpoint in log-sigma), ancestral noise from `noise_sampler(sigma, sigma_next)`.  The
    multiscale margins default to the reference sampler's own (:973-989: 5 / 8, no intermittent full-res steps); note that
    KSampler.sample always overrides them for this sampler (3 / 8) -- `sample()` below does the same."""
    B, _, oh, ow = x.shape
    sh = int(max(8, ((oh * multiscale_factor) // 8) * 8)) if enable_multiscale else oh
    sw = int(max(8, ((ow * multiscale_factor) // 8) * 8)) if enable_multiscale else ow
    active = enable_multiscale and (sh != oh or sw != ow)
    sig = sigmas.float().cpu()
    n = len(sig) - 1
    if noise_sampler is None:
        noise_sampler = BrownianIntervalNoise(x, seed, batch_slice)
    full = SamplerLoop(engine, B, oh, ow)
    low = SamplerLoop(engine, B, sh, sw) if active else None
    sigma_fn = lambda t: t.neg().exp()
    t_fn = lambda s: s.log().neg()

    def denoised_at(xx: torch.Tensor, sigma: float, fullres: bool) -> torch.Tensor:
        loop = full if fullres else low
        xin = xx if fullres else engine.resample_bilinear(xx, (sh, sw))
        du, dc = loop.denoise_pair(xin, sigma)
        d = torch.empty_like(xin)
        engine.cfg_step(None, du, dc, cfg, 2, denoised_out=d)
        return d if fullres else engine.resample_bilinear(d, (oh, ow))

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
           interrupt: Optional[Callable[[], bool]] = None,
           batch_slice: Optional[Tuple[int, int, int]] = None) -> Tuple[Dict[str, torch.Tensor]]:
    """Drop-in for KSampler.sample on the measured path. positive / negative: [1 or B, 77k, 768] conditioning tensors.
    sampler_options: the `extra_options` of the reference's `ksampler(name, extra_options)` seam (sampling.py:500-534),
    i.e. keyword arguments of the sampler function (SAMPLER_OPTIONS lists what each sampler takes; anything else raises).
    batch_slice = (lo, hi, total): this call samples images [lo, hi) of a batch of `total` (multi-GPU sharding) -- the
    default per-step noise of the ancestral / SDE samplers is then drawn for the whole batch and sliced.
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
    opts = dict(sampler_options or {})
    unknown = set(opts) - SAMPLER_OPTIONS[sampler_name]
    if unknown:
        raise ValueError(f"unknown {sampler_name} options {sorted(unknown)} (accepted: {sorted(SAMPLER_OPTIONS[sampler_name])})")
    if sampler_name == "dpmpp_2m_cfgpp":
        # sample1's multiscale_supported_samplers lists "sample_dpmpp_2m_cfgpp", which never matches this name, so the
        # reference runs this sampler with the sampler function's OWN defaults (5 / 8 / intermittent)
        opts.setdefault("enable_multiscale", enable_multiscale)
        x = sample_dpmpp_2m_cfgpp(engine, x, sigmas, cfg, callback=callback, interrupt=interrupt, **opts)
    elif sampler_name == "dpmpp_sde_cfgpp":
        # ... whereas "dpmpp_sde_cfgpp" IS in that list: KSampler.sample / common_ksampler / sample1 always hand it
        # extra_options with their own defaults (sampling.py:795-799, 949-964)
        for k, v in KSAMPLER_SDE_MULTISCALE.items():
            opts.setdefault(k, v)
        opts.setdefault("enable_multiscale", enable_multiscale)
        opts.setdefault("noise_sampler", noise_sampler)
        x = sample_dpmpp_sde_cfgpp(engine, x, sigmas, cfg, seed=seed, callback=callback, interrupt=interrupt,
                                   batch_slice=batch_slice, **opts)
    elif sampler_name == "euler_cfgpp":
        x = sample_euler_cfgpp(engine, x, sigmas, cfg, callback=callback, interrupt=interrupt, **opts)
    else:
        opts.setdefault("noise_sampler", noise_sampler)
        x = sample_euler_ancestral_cfgpp(engine, x, sigmas, cfg, callback=callback, interrupt=interrupt,
                                         batch_slice=batch_slice, **opts)
    out = (x / LATENT_SCALE).to(torch.float32).cpu()
    return ({"samples": out},)
