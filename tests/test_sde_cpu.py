"""CPU: dpmpp_sde_cfgpp (the reference pipeline's default sampler) — oracle and host loop vs the reference golden, with
the same injected deterministic noise sampler the golden was generated with."""
import os

import pytest
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm())


class SeqNoise:
    def __init__(self, shape, seed):
        self.g = torch.Generator().manual_seed(seed)
        self.shape = shape
        self.calls = []

    def __call__(self, sigma, sigma_next):
        self.calls.append((float(sigma), float(sigma_next)))
        return torch.randn(self.shape, generator=self.g)


@pytest.mark.parametrize("name,steps,ms", [("sde", 4, False), ("sde_ms", 15, True)])
def test_dpmpp_sde_matches_reference(unet_sd, name, steps, ms):
    from fake_engine import FakeEngine
    from lightdiffusion_next_b200 import sampling as S
    from oracle import sde_oracle as D
    g = torch.load(os.path.join(GOLDEN, "sde_small.pt"))
    lat = torch.zeros(1, 4, 16, 16)
    ns = SeqNoise(lat.shape, 99)
    o = D.ksample_sde(unet_sd, 42, steps, 7.0, "karras", g["ctx_pos"], g["ctx_neg"], lat, ns, multiscale=ms)
    assert rel(o, g[f"{name}_final"]) < 1e-4
    assert torch.allclose(torch.tensor(ns.calls), g[f"{name}_calls"], rtol=1e-5)
    ns2 = SeqNoise(lat.shape, 99)
    eng = FakeEngine(unet_sd)
    e = S.sample(eng, 42, steps, 7.0, "dpmpp_sde_cfgpp", "karras", g["ctx_pos"], g["ctx_neg"], {"samples": lat},
                 enable_multiscale=ms, noise_sampler=ns2,
                 # this golden went through the ksampler() seam with only enable_multiscale set, i.e. with the sampler
                 # function's own margin (5); KSampler.sample's 3 is pinned by test_round2_cpu.py
                 sampler_options={"multiscale_fullres_start": 5})[0]["samples"]
    assert rel(e, g[f"{name}_final"]) < 1e-4
    assert eng.denoise_calls == 2 * steps - 1   # two model evaluations per step except the last


def test_default_brownian_noise_is_unit_variance_and_path_consistent():
    from lightdiffusion_next_b200.sampling import BrownianIntervalNoise
    x = torch.zeros(64, 4, 32, 32)
    bn = BrownianIntervalNoise(x, seed=3)
    a = bn(10.0, 6.0)       # [sigma_s, sigma_i]
    b = bn(10.0, 4.0)       # [sigma_next, sigma_i] extends the same path
    assert abs(float(a.std()) - 1) < 0.02 and abs(float(b.std()) - 1) < 0.02
    # corr(W(10)-W(6), W(10)-W(4)) / (sqrt(4) sqrt(6)) = 4 / sqrt(24)
    corr = float((a * b).mean())
    assert abs(corr - 4 / 24 ** 0.5) < 0.02
