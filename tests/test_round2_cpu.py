"""CPU: round-2 host logic -- KSampler-routed dpmpp_sde_cfgpp options, sampler option validation, the config-1 golden vs
the oracle, the content-keyed context cache of the seam, pipeline(prompt: str, ...), shard-invariant per-step noise, and the
seam installed on the LIVE reference (only where /root/reference exists: the build container)."""
import os
import subprocess
import sys
import types

import pytest
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
has_reference = os.path.isdir(os.path.join(REF, "src"))


def rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm())


class SeqNoise:
    def __init__(self, shape, seed):
        self.g, self.shape, self.calls = torch.Generator().manual_seed(seed), tuple(shape), []

    def __call__(self, sigma, sigma_next):
        self.calls.append((float(sigma), float(sigma_next)))
        return torch.randn(self.shape, generator=self.g)


def test_sde_through_sample_uses_the_ksampler_multiscale_margins(unet_sd):
    """sample(..., "dpmpp_sde_cfgpp") == the reference's KSampler.sample for that sampler, which always injects
    multiscale_fullres_start=3 / end=8 / no intermittent steps (sampling.py:795-799, 949-964) -- NOT the sampler
    function's own 5 / 8.  Golden: tests/golden/make_golden_round2.py (KSampler.sample with the Brownian tree class replaced
    by a seeded sequence).  Also checks which UNet calls ran at half resolution."""
    from fake_engine import FakeEngine
    from lightdiffusion_next_b200 import sampling as S
    g = torch.load(os.path.join(GOLDEN, "sde_ksampler_small.pt"))
    a = g["a_args"]
    lat = torch.zeros(1, 4, a["hw"], a["hw"])
    eng = FakeEngine(unet_sd)
    seen_hw = []
    inner = eng.denoise
    eng.denoise = lambda x, s, out=None: (seen_hw.append(x.shape[-1]), inner(x, s, out))[1]
    ns = SeqNoise(lat.shape, 99)
    e = S.sample(eng, a["seed"], a["steps"], a["cfg"], "dpmpp_sde_cfgpp", "karras", g["ctx_pos"], g["ctx_neg"],
                 {"samples": lat}, noise_sampler=ns)[0]["samples"]
    assert seen_hw == g["a_call_hw"].tolist()
    assert torch.allclose(torch.tensor(ns.calls), g["a_noise_calls"], rtol=1e-5)
    assert rel(e, g["a_final"]) < 1e-4
    # the sampler function's own defaults (ksampler() seam without options) are different: low-res from step 5
    assert S.sample_dpmpp_sde_cfgpp.__kwdefaults__ is None
    import inspect
    d = inspect.signature(S.sample_dpmpp_sde_cfgpp).parameters
    assert d["multiscale_fullres_start"].default == 5 and S.KSAMPLER_SDE_MULTISCALE["multiscale_fullres_start"] == 3


def test_sampler_options_are_validated_for_every_sampler(unet_sd):
    from fake_engine import FakeEngine
    from lightdiffusion_next_b200 import sampling as S
    eng = FakeEngine(unet_sd)
    ctx = torch.zeros(1, 77, 768)
    lat = {"samples": torch.zeros(1, 4, 16, 16)}
    for name in S.SAMPLERS:
        with pytest.raises(ValueError, match=f"unknown {name} options"):
            S.sample(eng, 1, 1, 7.0, name, "karras", ctx, ctx, lat, sampler_options={"bogus": 1})
    # a valid option for one sampler is still rejected for another
    with pytest.raises(ValueError, match="unknown euler_cfgpp options"):
        S.sample(eng, 1, 1, 7.0, "euler_cfgpp", "karras", ctx, ctx, lat, sampler_options={"eta": 0.5})
    assert eng.denoise_calls == 0


def test_config1_golden_pins_the_oracle_at_512(unet_sd):
    """BASELINE config 1 (512^2, 20 steps Euler-a, seed 42) recorded from the live reference: the oracle reproduces the
    reference's uncond / cond denoised rows on the latents the reference itself visited (three of the 20 steps here to keep
    the CPU suite short; the GPU test covers all 20), and the recorded trajectory is self-consistent with the solver."""
    from lightdiffusion_next_b200.schedule import DiscreteSchedule, calculate_sigmas, get_ancestral_step
    from oracle import sd15_oracle as O
    g = torch.load(os.path.join(GOLDEN, "config1_euler_a_512.pt"))
    assert g["x"].shape == (20, 4, 64, 64) and g["args"]["steps"] == 20
    sig = calculate_sigmas(DiscreteSchedule(), "karras", 20)
    assert torch.allclose(g["sigma"], sig[:20])
    ctx = torch.cat([g["ctx_neg"], g["ctx_pos"]])
    for i in (0, 11, 19):
        x2 = g["x"][i:i + 1].repeat(2, 1, 1, 1)
        ref = torch.stack([g["den_uncond"][i], g["den_cond"][i]])
        out = O.apply_model(unet_sd, x2, g["sigma"][i].repeat(2), ctx)
        s = float(g["sigma"][i])
        assert rel((x2 - out) / s, (x2 - ref) / s) < 5e-5, i
    # the reference's x_{i+1} follows from x_i, its denoised rows and ITS noise (global CPU generator seeded by prepare_noise)
    torch.manual_seed(42)
    noise0 = torch.randn(1, 4, 64, 64)
    assert rel(noise0[0] * torch.sqrt(1.0 + sig[0] ** 2), g["x"][0]) < 1e-6
    x = g["x"][0:1].clone()
    for i in range(20):
        den = torch.lerp(g["den_uncond"][i:i + 1], g["den_cond"][i:i + 1], 7.0)
        sd_, su = get_ancestral_step(sig[i], sig[i + 1], 1.0)
        x = x + (x - den) / sig[i] * (sd_ - sig[i])
        if sig[i + 1] > 0:
            x = x + torch.randn(1, 4, 64, 64) * su
        if i + 1 < 20:
            assert rel(x[0], g["x"][i + 1]) < 1e-5, i
    assert rel(x / 0.18215, g["final"]) < 1e-5


def test_seam_context_cache_is_keyed_on_content(unet_sd):
    """calc_cond_batch hands the wrapper a FRESH c_crossattn tensor every step (torch.cat, cond.py:219-226): the engine's
    K/V precompute must still run once per conditioning, again when the content changes, and again after a weight reload."""
    from fake_engine import FakeEngine
    from lightdiffusion_next_b200.backend import EngineWrapper
    eng = FakeEngine(unet_sd)
    w = EngineWrapper(eng)
    g = torch.Generator().manual_seed(3)
    a, b = torch.randn(1, 77, 768, generator=g), torch.randn(1, 77, 768, generator=g)
    x = torch.randn(2, 4, 8, 8, generator=g)

    def call(neg, pos):
        ctx = torch.cat([neg, pos])  # a new tensor object every call, like the reference
        return w(None, {"input": x, "timestep": torch.tensor([1.5, 1.5]),
                        "c": {"c_crossattn": ctx, "transformer_options": {}}, "cond_or_uncond": [1, 0]})

    r1 = call(a, b)
    for _ in range(4):
        r = call(a, b)
    assert eng.context_uploads == 1 and torch.equal(r, r1) and w.calls == 5
    call(b, a)
    assert eng.context_uploads == 2
    call(b, a)
    assert eng.context_uploads == 2
    eng.weights_epoch[0] = 1   # Engine.load_weights(UNET) bumps this: the engine dropped its K/V buffers
    call(b, a)
    assert eng.context_uploads == 3
    call(torch.cat([b, b], 1), torch.cat([a, a], 1))   # another token count
    assert eng.context_uploads == 4
    # the reference samples under torch.inference_mode (pipeline.py:281): inference tensors have no version counter
    with torch.inference_mode():
        for _ in range(3):
            call(a, b)
    assert eng.context_uploads == 5


class RecordedTokenizer:
    """Answers tokenize_with_weights from rows the reference's SD1Tokenizer produced (tests/golden/clip_small.pt)."""

    def __init__(self, g):
        self.table = {"a (red:1.3) cube on a (blue:0.7) sphere": "weighted", "": "empty",
                      "a photograph of an astronaut riding a horse": "plain"}
        self.g = g

    def tokenize_with_weights(self, text, return_word_ids=False):
        n = self.table[text]
        return {"l": [list(zip(r.tolist(), w.tolist())) for r, w in zip(self.g[f"{n}_ids"], self.g[f"{n}_weights"])]}


def test_pipeline_prompt_str_surface(tiny_unet_sd):
    """pipeline(engine, prompt: str, w, h, ...) == tokenizer -> CLIPTextEncode x2 -> KSampler -> VAEDecode composed by hand
    (src/user/pipeline.py:31-55, 278-372), with the reference's argument names; out-of-scope flags raise."""
    from fake_engine import FakeEngine
    from lightdiffusion_next_b200 import pipeline as P
    from lightdiffusion_next_b200 import sampling as S
    from oracle import sd15_oracle as O
    vsd = O.synth_state_dict(O.vae_decoder_param_shapes(), seed=4321)
    csd = O.synth_state_dict(O.clip_param_shapes(), seed=777)
    from conftest import TINY_UNET
    eng = FakeEngine(tiny_unet_sd, vsd, csd, unet_cfg=TINY_UNET)  # composition test: a narrow UNet stands in for the denoiser
    g = torch.load(os.path.join(GOLDEN, "clip_small.pt"))
    tok = RecordedTokenizer(g)
    prompt, negative = "a (red:1.3) cube on a (blue:0.7) sphere", "a photograph of an astronaut riding a horse"
    imgs = P.pipeline(eng, prompt, 64, 64, number=1, batch=1, prio_speed=True, negative_prompt=negative, tokenizer=tok, seed=5)
    assert len(imgs) == 1 and imgs[0].shape == (1, 64, 64, 3)
    assert P.last_seed == 5
    # by hand, from the reference's own conditioning tensors
    pipe = P.Pipeline(eng)
    pos, neg = pipe.encode(tok.tokenize_with_weights(prompt)["l"]), pipe.encode(tok.tokenize_with_weights(negative)["l"])
    assert rel(pos, g["weighted_cond"]) < 1e-4 and rel(neg, g["plain_cond"]) < 1e-4
    lat = S.sample(eng, 5, 20, 7.0, "dpmpp_2m_cfgpp", "karras", pos, neg, {"samples": torch.zeros(1, 4, 8, 8)})[0]["samples"]
    assert torch.equal(imgs[0], O.vae_decode(vsd, lat))
    # reuse_seed reproduces the image; a second `number` repeats the generation like the reference's loop
    again = P.pipeline(eng, prompt, 64, 64, number=2, prio_speed=True, negative_prompt=negative, tokenizer=tok, reuse_seed=True)
    assert len(again) == 2 and torch.equal(again[0], imgs[0]) and torch.equal(again[1], imgs[0])
    for flag in ("adetailer", "enhance_prompt", "autohdr", "img2img", "flux_enabled"):
        with pytest.raises(NotImplementedError, match=flag):
            P.pipeline(eng, prompt, 64, 64, tokenizer=tok, **{flag: True})
    with pytest.raises(ValueError, match="tokenizer"):
        P.pipeline(eng, prompt, 64, 64)
    assert P.MULTISCALE_PRESETS["performance"] == (True, 0.25, 5, 8, True)


@pytest.mark.skipif(not has_reference, reason="needs the reference checkout (build container only)")
def test_pipeline_with_the_reference_tokenizer_object():
    """The tokenizer argument is the reference's own SD1Tokenizer; prompt weights and the default negative prompt's missing
    textual-inversion files (skipped with a warning by the tokenizer) flow through unchanged."""
    script = r"""
import os, sys, types, tempfile, torch
sys.path.insert(0, %r); sys.path.insert(0, os.path.join(%r, "tests")); sys.path.insert(0, %r)
sys.modules.setdefault("torchsde", types.ModuleType("torchsde"))
from src.SD15 import SDToken
from fake_engine import FakeEngine
from lightdiffusion_next_b200 import pipeline as P
from oracle import sd15_oracle as O
tok = SDToken.SD1Tokenizer(tokenizer=lambda embedding_directory=None: SDToken.SDTokenizer(
    tokenizer_path=os.path.join(%r, "include", "sd1_tokenizer/"), embedding_directory=embedding_directory))
g = torch.load(os.path.join(%r, "clip_small.pt"))
rows = P._token_rows(tok, "a (red:1.3) cube on a (blue:0.7) sphere")
assert [t for t, _ in rows[0]] == g["weighted_ids"][0].tolist()
assert torch.allclose(torch.tensor([w for _, w in rows[0]]), g["weighted_weights"][0])
neg = P._token_rows(tok, P.DEFAULT_NEGATIVE_PROMPT)
assert len(neg) == 1 and len(neg[0]) == 77
eng = FakeEngine(None, None, O.synth_state_dict(O.clip_param_shapes(), seed=777))
cond = P.Pipeline(eng).encode(rows)
assert float((cond - g["weighted_cond"]).norm() / g["weighted_cond"].norm()) < 1e-4
print("OK")
""" % (ROOT, ROOT, REF, REF, GOLDEN)
    r = subprocess.run([sys.executable, "-c", script], capture_output=True, text=True, timeout=600, cwd=REF)
    assert r.returncode == 0 and "OK" in r.stdout, r.stdout[-1500:] + r.stderr[-3000:]


@pytest.mark.skipif(not has_reference, reason="needs the reference checkout (build container only)")
def test_install_on_the_live_reference_reproduces_its_own_run(tmp_path):
    """backend.install() on a live reference ModelPatcher: KSampler.sample then runs every UNet step through EngineWrapper
    (the reference's model_function_wrapper hook) and must reproduce the golden the unmodified reference produced -- with
    ONE context upload for the whole run although calc_cond_batch rebuilds c_crossattn every step.  The engine behind the
    wrapper is the CPU stand-in (oracle); the GPU engine's numerics are covered by tests -m gpu."""
    script = r"""
import os, sys, torch
sys.path.insert(0, os.path.join(%r, "tests", "golden")); sys.path.insert(0, os.path.join(%r, "tests"))
import _ref_setup as R
R.enter_reference()
from fake_engine import FakeEngine
from lightdiffusion_next_b200 import backend
from src.sample import sampling
model, mp, sd = R.build_reference_unet()
eng = FakeEngine(sd)
g = torch.load(os.path.join(%r, "sample_small.pt"))

def forbidden(*a, **k):
    raise AssertionError("BaseModel.apply_model must not run: the engine replaces it")
m = backend.install(mp, engine=eng)
model.apply_model = forbidden
assert m is not mp and "model_function_wrapper" not in mp.model_options
with torch.inference_mode():   # as the reference's pipeline() runs it (src/user/pipeline.py:281)
    res = sampling.KSampler().sample(model=m, seed=42, steps=6, cfg=7.0, sampler_name="dpmpp_2m_cfgpp", scheduler="karras",
                                     denoise=1.0, positive=[[g["ctx_pos"], {}]], negative=[[g["ctx_neg"], {}]],
                                     latent_image={"samples": torch.zeros(1, 4, 16, 16)}, pipeline=True)
out, ref = res[0]["samples"], g["dpmpp_2m_final"]
err = float((out - ref).norm() / ref.norm())
print("err", err, "uploads", eng.context_uploads, "denoise", eng.denoise_calls)
assert err < 1e-4 and eng.denoise_calls == 6 and eng.context_uploads == 1
print("OK")
""" % (ROOT, ROOT, GOLDEN)
    r = subprocess.run([sys.executable, "-c", script], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "OK" in r.stdout, r.stdout[-1500:] + r.stderr[-3000:]


def test_default_step_noise_is_shard_invariant():
    """every rank used to draw the same per-step noise.  With batch_slice = (lo, hi, total) the default noise
    samplers draw the whole batch from identically seeded generators and slice: shards get the rows the unsharded batch
    gets (and therefore different noise per image)."""
    from lightdiffusion_next_b200 import sampling as S
    x = torch.zeros(4, 4, 8, 8)
    torch.manual_seed(11)
    full = [S.default_noise_sampler(x)(1.0, 0.5) for _ in range(2)]
    for lo, hi in ((0, 1), (1, 4)):
        torch.manual_seed(11)
        f = S.default_noise_sampler(x[lo:hi], (lo, hi, 4))
        for k in range(2):
            assert torch.equal(f(1.0, 0.5), full[k][lo:hi])
    assert not torch.equal(full[0][0], full[0][1])
    bf = S.BrownianIntervalNoise(x, seed=3)
    ref = [bf(10.0, 6.0), bf(10.0, 4.0)]
    bs = S.BrownianIntervalNoise(x[1:3], seed=3, batch_slice=(1, 3, 4))
    assert torch.equal(bs(10.0, 6.0), ref[0][1:3]) and torch.equal(bs(10.0, 4.0), ref[1][1:3])
