"""ORACLE — TEST INFRASTRUCTURE ONLY (see sd15_oracle.py).  dpmpp_sde_cfgpp as the reference executes it
(src/sample/samplers.py:966-1254; SURVEY.md fact 7: the CFG++ momentum branches are never taken, so this is plain
DPM-Solver++ (SDE), eta = 1, r = 1/2, s_noise = 1, two model evaluations per step).  The noise sampler is injected
(the reference's default BrownianTree needs torchsde).  Pinned against tests/golden/sde_small.pt."""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

from . import sd15_oracle as O


def _fullres(step, n_steps, start=5, end=8):
    # sample_dpmpp_sde_cfgpp defaults: multiscale_intermittent_fullres=False
    return step < start or step >= n_steps - end


def sample_dpmpp_sde(denoise, x, sigmas, noise_sampler, multiscale=False, factor=0.5, eta=1.0, r=0.5):
    oh, ow = x.shape[-2:]
    sh = int(max(8, ((oh * factor) // 8) * 8)) if multiscale else oh
    sw = int(max(8, ((ow * factor) // 8) * 8)) if multiscale else ow
    active = multiscale and (sh != oh or sw != ow)
    n = len(sigmas) - 1
    s_in = x.new_ones([x.shape[0]])
    sigma_fn = lambda t: t.neg().exp()
    t_fn = lambda s: s.log().neg()

    def model(xx, sig, full):
        if full:
            return denoise(xx, sig * s_in)
        d = denoise(F.interpolate(xx, size=(sh, sw), mode="bilinear", align_corners=False), sig * s_in)
        return F.interpolate(d, size=(oh, ow), mode="bilinear", align_corners=False)

    for i in range(n):
        full = (not active) or _fullres(i, n)
        den = model(x, sigmas[i], full)
        if sigmas[i + 1] == 0:
            x = x + (x - den) / sigmas[i] * (sigmas[i + 1] - sigmas[i])
        else:
            t, t_next = t_fn(sigmas[i]), t_fn(sigmas[i + 1])
            s = t + (t_next - t) * r
            sd, su = O.ancestral_step(sigma_fn(t), sigma_fn(s), eta)
            s_ = t_fn(sd)
            x_2 = (sigma_fn(s_) / sigma_fn(t)) * x - (t - s_).expm1() * den + noise_sampler(sigma_fn(t), sigma_fn(s)) * su
            den_2 = model(x_2, sigma_fn(s), full)
            sd, su = O.ancestral_step(sigma_fn(t), sigma_fn(t_next), eta)
            t_next_ = t_fn(sd)
            d_mix = (1 - 1 / (2 * r)) * den + (1 / (2 * r)) * den_2
            x = (sigma_fn(t_next_) / sigma_fn(t)) * x - (t - t_next_).expm1() * d_mix + noise_sampler(sigma_fn(t), sigma_fn(t_next)) * su
    return x


def ksample_sde(sd, seed, steps, cfg, scheduler, cond, uncond, latent, noise_sampler, multiscale=False):
    tables = O.make_sigma_tables()
    sigmas = O.calculate_sigmas(scheduler, steps)
    noise = O.prepare_noise(latent.shape, seed)
    maxd = math.isclose(float(tables[0][-1]), float(sigmas[0]), rel_tol=1e-05) or float(sigmas[0]) > float(tables[0][-1])
    x = noise * torch.sqrt(1.0 + sigmas[0] ** 2.0) if maxd else noise * sigmas[0]
    x = x + latent

    def denoise(xx, ss):
        return O.cfg_denoise(lambda a, b, c: O.apply_model(sd, a, b, c, tables), xx, ss, cond, uncond, cfg)

    return sample_dpmpp_sde(denoise, x, sigmas, noise_sampler, multiscale=multiscale) / O.LATENT_SCALE
