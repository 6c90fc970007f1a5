"""CPU: pins the oracle (oracle/sd15_oracle.py) against fixtures produced by the reference itself
(tests/golden/make_golden.py, run in the build container against /root/reference)."""
import torch

from oracle import sd15_oracle as O


def rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm())


def test_sigma_tables_bit_exact(golden_unet):
    sigmas, log_sigmas = O.make_sigma_tables()
    assert torch.equal(sigmas, golden_unet["sigmas"])
    assert torch.equal(log_sigmas, golden_unet["log_sigmas"])
    assert abs(float(sigmas[0]) - 0.02917) < 1e-4 and abs(float(sigmas[-1]) - 14.6146) < 1e-3


def test_schedules_bit_exact(golden_unet):
    for name, steps in (("karras", 20), ("karras", 30), ("normal", 10), ("normal", 22)):
        assert torch.equal(O.calculate_sigmas(name, steps), golden_unet[f"sched_{name}_{steps}"]), (name, steps)


def test_timestep_index(golden_unet):
    _, ls = O.make_sigma_tables()
    for hw in (16, 32):
        t = O.timestep_index(golden_unet[f"apply_sigma_{hw}"], ls).float()
        assert torch.equal(t, golden_unet[f"apply_t_{hw}"])


def test_apply_model_matches_reference(golden_unet, unet_sd):
    # the reference computes fp32 on CPU from fp16-stored weights (manual_cast); the oracle does the same maths
    for hw in (16, 32):
        out = O.apply_model(unet_sd, golden_unet[f"apply_x_{hw}"], golden_unet[f"apply_sigma_{hw}"],
                            golden_unet[f"apply_ctx_{hw}"])
        r = rel(out, golden_unet[f"apply_out_{hw}"])
        assert r < 2e-5, (hw, r)


def test_seam_record_matches_oracle(golden_sample, unet_sd):
    # what crosses model_function_wrapper (cond.py:254-265): rows are [uncond, cond], timestep carries sigma
    g = golden_sample
    assert g["euler_a_seam_cond_or_uncond"].tolist() == [1, 0]
    assert torch.equal(g["euler_a_seam_ctx"][0:1], g["ctx_neg"]) and torch.equal(g["euler_a_seam_ctx"][1:2], g["ctx_pos"])
    out = O.apply_model(unet_sd, g["euler_a_seam_input"], g["euler_a_seam_timestep"], g["euler_a_seam_ctx"])
    assert rel(out, g["euler_a_seam_output"]) < 2e-5


def _run(g, unet_sd, sampler, sched, steps, hw=16):
    return O.ksample(unet_sd, seed=42, steps=steps, cfg=7.0, sampler=sampler, scheduler=sched, cond=g["ctx_pos"],
                     uncond=g["ctx_neg"], latent=torch.zeros(1, 4, hw, hw))


def test_ksample_euler_ancestral(golden_sample, unet_sd):
    out = _run(golden_sample, unet_sd, "euler_ancestral_cfgpp", "karras", 4)
    assert rel(out, golden_sample["euler_a_final"]) < 1e-4


def test_ksample_euler_ancestral_normal(golden_sample, unet_sd):
    out = _run(golden_sample, unet_sd, "euler_ancestral_cfgpp", "normal", 3)
    assert rel(out, golden_sample["euler_a_normal_final"]) < 1e-4


def test_ksample_dpmpp_2m(golden_sample, unet_sd):
    out = _run(golden_sample, unet_sd, "dpmpp_2m_cfgpp", "karras", 6)
    assert rel(out, golden_sample["dpmpp_2m_final"]) < 1e-4


def test_ksample_dpmpp_2m_multiscale_default_on(golden_sample, unet_sd):
    # 15 steps: steps 6 is low-res (8x8) under the sampler's own defaults
    out = _run(golden_sample, unet_sd, "dpmpp_2m_cfgpp", "karras", 15)
    assert rel(out, golden_sample["dpmpp_2m_ms_final"]) < 1e-4


def test_all_scheduler_names_bit_exact():
    """Every scheduler the reference's calculate_sigmas knows (karras, normal, simple, beta) at several step counts:
    oracle and the product's host schedule code both reproduce the reference bit for bit."""
    import os
    import torch
    from oracle import sd15_oracle as O
    from lightdiffusion_next_b200 import schedule as S
    gold = torch.load(os.path.join(os.path.dirname(__file__), "golden", "schedules.pt"))
    ms = S.DiscreteSchedule()
    fm = S.FluxSchedule()
    assert torch.equal(fm.sigmas, gold["flux_sigmas"])  # ModelSamplingFlux table (shift 1.15, 10000 entries)
    for key, ref in gold.items():
        if key == "flux_sigmas":
            continue
        if key.startswith("flux__"):
            name, steps = key[len("flux__"):].rsplit("_", 1)
            assert torch.equal(S.calculate_sigmas(fm, name, int(steps)), ref), key
            continue
        name, steps = key.rsplit("_", 1)
        assert torch.equal(O.calculate_sigmas(name, int(steps)), ref), key
        assert torch.equal(S.calculate_sigmas(ms, name, int(steps)), ref), key


def test_oracle_euler_cfgpp_vs_reference(unet_sd):
    """The reference's 4th named sampler (euler_cfgpp = sample_euler_dy_cfg_pp with its dynamic half-resolution steps):
    oracle trajectory vs full reference KSampler runs (tests/golden/make_golden_euler_cfgpp.py)."""
    import os
    import torch
    from oracle import sd15_oracle as O
    gold = torch.load(os.path.join(os.path.dirname(__file__), "golden", "euler_cfgpp_small.pt"))
    for name in ("a", "b"):
        a = gold[f"{name}_args"]
        out = O.ksample(unet_sd, 42, a["steps"], a["cfg"], "euler_cfgpp", a["scheduler"], gold["ctx_pos"], gold["ctx_neg"],
                        torch.zeros(1, 4, a["hw"], a["hw"]))
        ref = gold[f"{name}_final"]
        err = ((out - ref).norm() / ref.norm()).item()
        assert err < 1e-4, (name, err)
