"""CPU: HiresFix pieces (bislerp LatentUpscale, denoise < 1 schedules) — oracle and host code vs the reference golden."""
import os

import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm())


def test_bislerp_oracle_and_product_match_reference():
    from lightdiffusion_next_b200 import latent as L
    from oracle import hires_oracle as H
    g = torch.load(os.path.join(GOLDEN, "hires_small.pt"))
    for fn in (H.bislerp, L.bislerp):
        up = fn(g["lat"], 16, 16)
        assert up.shape == g["up"].shape and rel(up, g["up"]) < 1e-6, fn.__module__
        up2 = fn(g["lat"][:1], 24, 16)
        assert up2.shape == g["up_rect"].shape and rel(up2, g["up_rect"]) < 1e-6
    assert torch.equal(L.latent_upscale({"samples": g["lat"]}, 0, 0)["samples"], g["lat"])
    assert L.latent_upscale({"samples": g["lat"]}, 128, 128)["samples"].shape == (2, 4, 16, 16)


def test_denoise_schedule_and_second_pass(unet_sd):
    from fake_engine import FakeEngine
    from lightdiffusion_next_b200 import sampling as S
    from oracle import hires_oracle as H
    g = torch.load(os.path.join(GOLDEN, "hires_small.pt"))
    assert torch.equal(H.sigmas_for_denoise("normal", 4, 0.45), g["sched_normal_4_d045"])
    ref = g["hires_final"]
    o = H.ksample(unet_sd, 43, 4, 8.0, "dpmpp_2m_cfgpp", "normal", g["ctx_pos"], g["ctx_neg"], g["up"][:1], denoise=0.45)
    assert rel(o, ref) < 1e-4
    e = S.sample(FakeEngine(unet_sd), 43, 4, 8.0, "dpmpp_2m_cfgpp", "normal", g["ctx_pos"], g["ctx_neg"],
                 {"samples": g["up"][:1]}, denoise=0.45)[0]["samples"]
    assert rel(e, ref) < 1e-4


def test_unequal_context_lengths_are_tiled_to_lcm(unet_sd):
    from fake_engine import FakeEngine
    from lightdiffusion_next_b200 import sampling as S
    g = torch.Generator().manual_seed(0)
    pos = torch.randn(1, 154, 768, generator=g)
    neg = torch.randn(1, 77, 768, generator=g)
    eng = FakeEngine(unet_sd)
    S.sample(eng, 1, 1, 7.0, "dpmpp_2m_cfgpp", "karras", pos, neg, {"samples": torch.zeros(1, 4, 16, 16)})
    assert eng.ctx.shape == (2, 154, 768)
    assert torch.equal(eng.ctx[0, :77], neg[0]) and torch.equal(eng.ctx[0, 77:], neg[0]) and torch.equal(eng.ctx[1], pos[0])


def test_inpaint_noise_mask_is_inert_like_the_reference(unet_sd):
    """`latent["noise_mask"]` (SetLatentNoiseMask -> common_ksampler, sampling.py:1203-1221): the reference hands it to
    KSamplerX0Inpaint, whose __call__ forwards to the model without blending (sampling.py:363-378), so for the SD1.5 UNet a
    mask changes nothing -- recorded from a reference run by tests/golden/make_golden_mask.py.  The engine's sample() takes
    the same latent dict and reproduces the reference's result bit for bit in behaviour: the mask is accepted and inert."""
    from fake_engine import FakeEngine
    from lightdiffusion_next_b200 import sampling as S
    g = torch.load(os.path.join(GOLDEN, "hires_small.pt"))
    m = torch.load(os.path.join(GOLDEN, "mask_small.pt"))
    assert m["equals_unmasked"] and torch.equal(m["masked_final"], g["hires_final"])
    e = S.sample(FakeEngine(unet_sd), 43, 4, 8.0, "dpmpp_2m_cfgpp", "normal", g["ctx_pos"], g["ctx_neg"],
                 {"samples": g["up"][:1], "noise_mask": m["mask"]}, denoise=0.45)[0]["samples"]
    assert rel(e, m["masked_final"]) < 1e-4
