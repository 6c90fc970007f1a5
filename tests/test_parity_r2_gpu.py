"""GPU parity, round 2: the holes the round-1 review listed.

 * euler_ancestral_cfgpp and dpmpp_sde_cfgpp: NUMERIC final-latent parity against reference runs, with the reference's own
   per-step noise injected (Euler-a: the global CPU generator the reference draws from after prepare_noise; SDE: the seeded
   sequence the goldens were generated with).
 * BASELINE config 1 at full size (512x512, 20 steps Euler-a, seed 42): all 20 UNet calls under teacher forcing against the
   reference's recorded rows, and the free-running trajectory.
 * VAE decode at the BASELINE image size (128x128 latent -> 1024x1024) against the fp32 oracle.
 * HiresFix second pass / img2img-style partial denoise, the bilinear resample kernel, the seam's context cache on the GPU.

Tolerances: one bf16 forward <= 1.5e-2 rel-L2 on eps (the reference's own bf16 path: 1.06e-2; measured here
~0.9e-2); free-running trajectories amplify per-step error: <= 6e-2 up to 15 steps, <= 1e-1 for the 20-step ancestral run."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
EPS_TOL = 1.5e-2
TRAJ_TOL = 6e-2


def rel(a, b):
    return float((a.double().cpu() - b.double().cpu()).norm() / b.double().cpu().norm())


class SeqNoise:
    def __init__(self, shape, seed):
        self.g, self.shape, self.calls = torch.Generator().manual_seed(seed), tuple(shape), []

    def __call__(self, sigma, sigma_next):
        self.calls.append((float(sigma), float(sigma_next)))
        return torch.randn(self.shape, generator=self.g)


@pytest.fixture(scope="module")
def engine(unet_sd):
    from lightdiffusion_next_b200.engine import Engine
    eng = Engine(max_rows=2, max_h=64, max_w=64, max_ctx_tokens=77)
    eng.load_unet(unet_sd)
    return eng


def cpu_global_noise(shape):
    """What the reference's default_noise_sampler does on the CPU: torch.randn_like(x) from the global generator, which
    prepare_noise seeded (and advanced by the initial-noise draw) -- sample() replays prepare_noise, so the order matches."""
    return lambda sigma, sigma_next: torch.randn(shape)


@pytest.mark.parametrize("name,sched,steps", [("euler_a", "karras", 4), ("euler_a_normal", "normal", 3)])
def test_euler_ancestral_numeric_vs_reference(engine, golden_sample, name, sched, steps):
    from lightdiffusion_next_b200 import sampling as S
    g = golden_sample
    out = S.sample(engine, 42, steps, 7.0, "euler_ancestral_cfgpp", sched, g["ctx_pos"], g["ctx_neg"],
                   {"samples": torch.zeros(1, 4, 16, 16)}, noise_sampler=cpu_global_noise((1, 4, 16, 16)))[0]["samples"]
    assert rel(out, g[f"{name}_final"]) < TRAJ_TOL, rel(out, g[f"{name}_final"])


@pytest.mark.parametrize("name,steps,ms", [("sde", 4, False), ("sde_ms", 15, True)])
def test_dpmpp_sde_numeric_vs_reference(engine, name, steps, ms):
    """dpmpp_sde_cfgpp (the pipeline default sampler; two UNet evaluations per step) through the ksampler() seam options the
    golden was generated with (enable_multiscale only => the sampler's own margin 5)."""
    from lightdiffusion_next_b200 import sampling as S
    g = torch.load(os.path.join(GOLDEN, "sde_small.pt"))
    ns = SeqNoise((1, 4, 16, 16), 99)
    out = S.sample(engine, 42, steps, 7.0, "dpmpp_sde_cfgpp", "karras", g["ctx_pos"], g["ctx_neg"],
                   {"samples": torch.zeros(1, 4, 16, 16)}, enable_multiscale=ms, noise_sampler=ns,
                   sampler_options={"multiscale_fullres_start": 5})[0]["samples"]
    assert torch.allclose(torch.tensor(ns.calls), g[f"{name}_calls"], rtol=1e-5)
    assert rel(out, g[f"{name}_final"]) < TRAJ_TOL, rel(out, g[f"{name}_final"])


@pytest.mark.parametrize("name", ["a", "b"])
def test_dpmpp_sde_through_ksampler_defaults(engine, name):
    """... and as KSampler.sample runs it (margins 3 / 8, golden recorded through KSampler.sample itself)."""
    from lightdiffusion_next_b200 import sampling as S
    g = torch.load(os.path.join(GOLDEN, "sde_ksampler_small.pt"))
    a = g[f"{name}_args"]
    ns = SeqNoise((1, 4, a["hw"], a["hw"]), 99)
    out = S.sample(engine, a["seed"], a["steps"], a["cfg"], "dpmpp_sde_cfgpp", "karras", g["ctx_pos"], g["ctx_neg"],
                   {"samples": torch.zeros(1, 4, a["hw"], a["hw"])}, noise_sampler=ns)[0]["samples"]
    assert torch.allclose(torch.tensor(ns.calls), g[f"{name}_noise_calls"], rtol=1e-5)
    # 2 x steps chained UNet evaluations with injected noise of sigma-scale magnitude: wider than the 15-step bound
    assert rel(out, g[f"{name}_final"]) < 1e-1, rel(out, g[f"{name}_final"])


def test_config1_teacher_forced_and_free_running(engine):
    """BASELINE config 1: 512x512, euler_ancestral_cfgpp, karras, 20 steps, cfg 7, seed 42."""
    from lightdiffusion_next_b200 import sampling as S
    g = torch.load(os.path.join(GOLDEN, "config1_euler_a_512.pt"))
    ctx = torch.cat([g["ctx_neg"], g["ctx_pos"]]).cuda()
    engine.set_context(ctx)
    worst = 0.0
    for i in range(20):
        x2 = g["x"][i:i + 1].repeat(2, 1, 1, 1)
        s = float(g["sigma"][i])
        out = engine.denoise(x2.cuda(), g["sigma"][i].repeat(2).cuda()).cpu()
        ref = torch.stack([g["den_uncond"][i], g["den_cond"][i]])
        e = rel((x2 - out) / s, (x2 - ref) / s)
        worst = max(worst, e)
        assert e < EPS_TOL, (i, e)
    a = g["args"]
    out = S.sample(engine, a["seed"], a["steps"], a["cfg"], a["sampler"], a["scheduler"], g["ctx_pos"], g["ctx_neg"],
                   {"samples": torch.zeros(1, 4, a["hw"], a["hw"])},
                   noise_sampler=cpu_global_noise((1, 4, a["hw"], a["hw"])))[0]["samples"]
    traj = rel(out, g["final"])
    print(f"config 1: worst teacher-forced eps rel-L2 {worst:.3e} over 20 steps, free-running final latent rel-L2 {traj:.3e}")
    assert traj < 1e-1, traj


def test_vae_decode_at_1024(unet_sd):
    """VAE decode at the BASELINE size (latent 128x128 -> 1024x1024 image; mid-block attention N = 16384, d = 512) against
    the fp32 oracle -- the big-N attention path differs from the small-latent one the goldens cover."""
    import time
    from lightdiffusion_next_b200.engine import Engine
    from oracle import sd15_oracle as O
    vsd = O.synth_state_dict(O.vae_decoder_param_shapes(), seed=4321)
    eng = Engine(max_rows=2, max_h=8, max_w=8)
    eng.load_vae(vsd)
    g = torch.Generator().manual_seed(17)
    z = torch.randn(1, 4, 128, 128, generator=g)
    out = eng.vae_decode(z.cuda()).cpu()
    assert out.shape == (1, 1024, 1024, 3) and torch.isfinite(out).all()
    torch.set_num_threads(os.cpu_count() or 1)
    t0 = time.time()
    ref = O.vae_decode(vsd, z)
    r, mx = rel(out, ref), float((out - ref).abs().max())
    print(f"VAE decode 1024^2: rel-L2 {r:.3e}, max-abs {mx:.3e} (oracle {time.time() - t0:.0f} s on the host)")
    assert r < 2e-2, r
    eng.close()


def test_hires_second_pass_and_helper(engine):
    """HiresFix second pass as recorded from the reference (KSampler on a bislerp-upscaled latent, denoise 0.45), and the
    pipeline's hires_fix() helper == LatentUpscale x2 + KSampler(10 steps, cfg 8, euler_ancestral_cfgpp, normal, 0.45)."""
    import lightdiffusion_next_b200.sampling as S
    from lightdiffusion_next_b200.latent import latent_upscale
    from lightdiffusion_next_b200.pipeline import Pipeline, hires_fix
    g = torch.load(os.path.join(GOLDEN, "hires_small.pt"))
    out = S.sample(engine, 43, 4, 8.0, "dpmpp_2m_cfgpp", "normal", g["ctx_pos"], g["ctx_neg"], {"samples": g["up"][:1]},
                   denoise=0.45)[0]["samples"]
    assert rel(out, g["hires_final"]) < TRAJ_TOL, rel(out, g["hires_final"])
    lat = g["lat"][:1]
    up = latent_upscale({"samples": lat}, 128, 128)["samples"]
    assert torch.equal(up, g["up"][:1])   # bislerp: bit-level parity with the reference's LatentUpscale
    zero = lambda s0, s1: torch.zeros(1, 4, 16, 16)   # isolates the composition from where the ancestral noise is drawn
    direct = S.sample(engine, 9, 10, 8.0, "euler_ancestral_cfgpp", "normal", g["ctx_pos"], g["ctx_neg"], {"samples": up},
                      denoise=0.45, noise_sampler=zero)[0]["samples"]
    orig = S.default_noise_sampler
    S.default_noise_sampler = lambda x, bs=None: (lambda s0, s1: torch.zeros_like(x))
    try:
        via_helper = hires_fix(Pipeline(engine), lat, g["ctx_pos"], g["ctx_neg"], 64, 64, seed=9)
    finally:
        S.default_noise_sampler = orig
    assert torch.equal(via_helper, direct)


def test_hires_partial_denoise_vs_oracle(engine, unet_sd):
    """denoise < 1 (schedule tail + noise added to a non-zero latent, CFG.py:266-269) on the GPU vs the fp32 oracle."""
    from lightdiffusion_next_b200 import sampling as S
    from oracle import hires_oracle as H
    g = torch.load(os.path.join(GOLDEN, "hires_small.pt"))
    up = g["up"][:1]
    ref = H.ksample(unet_sd, 21, 5, 8.0, "dpmpp_2m_cfgpp", "normal", g["ctx_pos"], g["ctx_neg"], up, denoise=0.6)
    out = S.sample(engine, 21, 5, 8.0, "dpmpp_2m_cfgpp", "normal", g["ctx_pos"], g["ctx_neg"], {"samples": up},
                   denoise=0.6)[0]["samples"]
    assert rel(out, ref) < TRAJ_TOL, rel(out, ref)


def test_bislerp_on_the_device(engine):
    """`ldn_bislerp` (the HiresFix LatentUpscale on the device) against the host restatement, which is bit-identical to the
    reference's `bislerp` (tests/golden/hires_small.pt): the golden upscale itself, integer and fractional ratios, a
    downscale, and a latent with a zero vector / parallel / opposite neighbours (the slerp's special cases)."""
    from lightdiffusion_next_b200.latent import bislerp
    g = torch.load(os.path.join(GOLDEN, "hires_small.pt"))
    up = engine.bislerp(g["lat"].cuda(), 16, 16).cpu()
    assert rel(up, g["up"]) < 1e-5 and float((up - g["up"]).abs().max()) < 1e-4
    gen = torch.Generator().manual_seed(7)
    for (n, c, h, w, H, W) in [(1, 4, 64, 64, 256, 256), (2, 4, 24, 40, 61, 77), (1, 16, 32, 32, 48, 80), (1, 4, 40, 24, 16, 16)]:
        x = torch.randn(n, c, h, w, generator=gen)
        if (n, c, h, w) == (2, 4, 24, 40):
            x[0, :, 3, 5] = 0.0                     # zero vector
            x[0, :, 7, 9] = 2.5 * x[0, :, 7, 8]     # parallel neighbours (first tap wins)
            x[1, :, 11, 4] = -0.5 * x[1, :, 11, 3]  # opposite neighbours (linear blend)
        ref = bislerp(x, W, H)
        out = engine.bislerp(x.cuda(), W, H).cpu()
        assert out.shape == ref.shape
        assert rel(out, ref) < 1e-5, (n, c, h, w, H, W, rel(out, ref))
        assert float((out - ref).abs().max()) < 1e-3

def test_pipeline_img2img_vs_oracle(engine, unet_sd):
    """`Pipeline.img2img` (VAEEncode -> KSampler with denoise < 1, the img2img branch of the reference's pipeline,
    src/user/pipeline.py:120-216 / VariationalAE.py:787-801) on the GPU: the encoded posterior sample against the oracle's
    encoder with the same noise, and the partial-denoise pass from that latent against the fp32 oracle sampler."""
    from lightdiffusion_next_b200.pipeline import Pipeline
    from oracle import hires_oracle as H
    from oracle import sd15_oracle as O
    g = torch.load(os.path.join(GOLDEN, "hires_small.pt"))
    vsd = dict(O.synth_state_dict(O.vae_decoder_param_shapes(), seed=4321))
    vsd.update(O.synth_state_dict(O.vae_encoder_param_shapes(), seed=2468))
    engine.load_vae(vsd)
    pipe = Pipeline(engine)
    pixels = torch.rand(1, 128, 128, 3, generator=torch.Generator().manual_seed(3))
    noise = torch.randn(1, 4, 16, 16, generator=torch.Generator().manual_seed(4))
    mean, logvar = O.vae_encode_moments(vsd, pixels).chunk(2, dim=1)
    enc_ref = mean + torch.exp(0.5 * torch.clamp(logvar, -30.0, 20.0)) * noise
    enc = engine.vae_encode(pixels, noise=noise)
    assert rel(enc, enc_ref) < 2e-2, rel(enc, enc_ref)
    # the pipeline call draws the posterior noise from the global CPU generator, like the reference
    torch.manual_seed(11)
    expect_noise = torch.randn(1, 4, 16, 16)
    torch.manual_seed(11)
    out = pipe.img2img(pixels, g["ctx_pos"], g["ctx_neg"], seed=21, steps=5, cfg=8.0, denoise=0.6, sampler_name="dpmpp_2m_cfgpp",
                       scheduler="normal")  # (reference-default multiscale schedule, as the oracle's sampler runs it)
    lat0 = engine.vae_encode(pixels, noise=expect_noise)
    ref = H.ksample(unet_sd, 21, 5, 8.0, "dpmpp_2m_cfgpp", "normal", g["ctx_pos"], g["ctx_neg"], lat0, denoise=0.6)
    assert out.shape == (1, 4, 16, 16) and torch.isfinite(out).all()
    assert rel(out, ref) < TRAJ_TOL, rel(out, ref)


def test_resample_bilinear_matches_aten(engine):
    import torch.nn.functional as F
    g = torch.Generator().manual_seed(1)
    for shape, size in (((2, 4, 64, 64), (32, 32)), ((2, 4, 32, 32), (64, 64)), ((1, 4, 40, 24), (16, 8)),
                        ((1, 4, 16, 8), (40, 24)), ((3, 4, 128, 128), (64, 64)), ((1, 4, 24, 40), (24, 40))):
        x = torch.randn(shape, generator=g).cuda()
        ref = F.interpolate(x, size=size, mode="bilinear", align_corners=False)
        out = engine.resample_bilinear(x, size)
        assert out.shape == ref.shape
        assert float((out - ref).abs().max()) < 1e-5, (shape, size)


def test_seam_context_cache_on_the_engine(engine, golden_sample):
    """Through the primary seam the reference passes a fresh c_crossattn every step: ldn_set_context must run once."""
    from lightdiffusion_next_b200.backend import EngineWrapper
    g = golden_sample
    w = EngineWrapper(engine)
    n0 = engine.context_uploads
    outs = []
    for _ in range(4):
        params = {"input": g["dpmpp_2m_seam_input"].cuda(), "timestep": g["dpmpp_2m_seam_timestep"].cuda(),
                  "c": {"c_crossattn": g["dpmpp_2m_seam_ctx"].clone().cuda(), "transformer_options": {}},
                  "cond_or_uncond": [1, 0]}
        outs.append(w(None, params))
    assert engine.context_uploads == n0 + 1
    assert all(torch.equal(o, outs[0]) for o in outs)
    x, s = g["dpmpp_2m_seam_input"], g["dpmpp_2m_seam_timestep"].view(-1, 1, 1, 1)
    assert rel((x - outs[0].cpu()) / s, (x - g["dpmpp_2m_seam_output"]) / s) < EPS_TOL
    # reloading the UNet drops the engine's K/V buffers: the wrapper must upload again instead of failing
    from oracle import sd15_oracle as O
    engine.load_unet(O.synth_state_dict(O.unet_param_shapes()))
    params["c"]["c_crossattn"] = g["dpmpp_2m_seam_ctx"].clone().cuda()
    again = w(None, params)
    assert engine.context_uploads == n0 + 2 and torch.equal(again, outs[0])
