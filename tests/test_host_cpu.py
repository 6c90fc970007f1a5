"""CPU: the C-ABI library loads and exports what include/ldn.h declares; host-side logic (schedules, synthetic weights,
sampler loop, batch sharding over 2 gloo ranks) matches the oracle / reference goldens.  No GPU compute here."""
import ctypes
import os
import re
import subprocess
import sys

import pytest
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm())


def test_library_exports_every_declared_symbol():
    from lightdiffusion_next_b200 import _lib
    lib = _lib.load()
    assert _lib.MISSING == []
    header = open(os.path.join(ROOT, "include", "ldn.h")).read()
    declared = set(re.findall(r"\b(ldn_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 15
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/ldn.h but not exported"
        assert name in _lib.SIGNATURES, f"{name} has no ctypes prototype"
    assert lib.ldn_version() >= 100


@pytest.mark.skipif(torch.cuda.is_available(), reason="CPU-only check of the no-fallback rule")
def test_create_fails_loudly_without_gpu():
    from lightdiffusion_next_b200 import _lib
    lib = _lib.load()
    cfg = _lib.ldn_config(2, 32, 32, 77, 1)
    h = ctypes.c_void_p()
    rc = lib.ldn_create(ctypes.byref(cfg), ctypes.byref(h))
    assert rc != 0 and lib.ldn_last_error()
    from lightdiffusion_next_b200.engine import Engine
    with pytest.raises(_lib.LdnError):
        Engine()


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "lightdiffusion_next_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert "oracle" not in re.sub(r'""".*?"""', "", src, flags=re.S), fn


def test_schedule_matches_reference_tables(golden_unet):
    from lightdiffusion_next_b200.schedule import DiscreteSchedule, calculate_sigmas
    ms = DiscreteSchedule()
    assert torch.equal(ms.sigmas, golden_unet["sigmas"]) and torch.equal(ms.log_sigmas, golden_unet["log_sigmas"])
    for name, steps in (("karras", 20), ("karras", 30), ("normal", 10), ("normal", 22)):
        assert torch.equal(calculate_sigmas(ms, name, steps), golden_unet[f"sched_{name}_{steps}"])
    for hw in (16, 32):
        assert torch.equal(ms.timestep(golden_unet[f"apply_sigma_{hw}"]).float(), golden_unet[f"apply_t_{hw}"])


def test_synth_weights_match_oracle_generator():
    from lightdiffusion_next_b200 import synth
    from oracle import sd15_oracle as O
    a, b = synth.unet_shapes(), O.unet_param_shapes()
    assert a == b and len(a) == 686
    assert sum(int(torch.tensor(s).prod()) for s in a.values()) == 859_520_964
    for k in ("time_embed.0.weight", "out.0.bias", "middle_block.1.transformer_blocks.0.ff.net.2.weight"):
        assert torch.equal(synth.synth_tensor(k, a[k]), O.synth_state_dict({k: b[k]})[k])


@pytest.mark.parametrize("name,sampler,sched,steps", [("euler_a", "euler_ancestral_cfgpp", "karras", 4),
                                                      ("dpmpp_2m", "dpmpp_2m_cfgpp", "karras", 6),
                                                      ("dpmpp_2m_ms", "dpmpp_2m_cfgpp", "karras", 15)])
def test_sampler_loop_host_logic(golden_sample, unet_sd, name, sampler, sched, steps):
    """The engine's sampler loop (coefficients, CFG row order, noise, multiscale schedule) driven by a CPU fake engine
    reproduces the reference's final latents."""
    from lightdiffusion_next_b200 import sampling as S
    from fake_engine import FakeEngine
    g = golden_sample
    eng = FakeEngine(unet_sd)
    out = S.sample(eng, 42, steps, 7.0, sampler, sched, g["ctx_pos"], g["ctx_neg"], {"samples": torch.zeros(1, 4, 16, 16)})
    assert rel(out[0]["samples"], g[f"{name}_final"]) < 1e-4
    assert eng.denoise_calls == steps


def test_shard_range_ragged():
    from lightdiffusion_next_b200.distributed import shard_range
    for batch in (0, 1, 3, 8, 32, 33):
        for world in (1, 2, 3, 8):
            spans = [shard_range(batch, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == batch
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


_WORKER = r"""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "tests"))
from lightdiffusion_next_b200 import distributed as D, sampling as S
from lightdiffusion_next_b200.synth import unet_shapes, synth_tensor
from fake_engine import FakeEngine
from conftest import TINY_UNET
from oracle import sd15_oracle as O
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
torch.set_num_threads(4)
shapes = O.unet_param_shapes(TINY_UNET)  # narrow UNet of the SD1.5 topology: this test pins host logic, not numerics
sd = D.broadcast_state_dict(shapes, synth_tensor, "cpu")
# rank 1 received exactly what rank 0 generated
chk = synth_tensor("out.2.weight", shapes["out.2.weight"])
assert torch.equal(sd["out.2.weight"], chk)
eng = FakeEngine(sd, unet_cfg=TINY_UNET)
g = torch.Generator().manual_seed(1234)
pos = torch.randn(1, 77, 768, generator=g); neg = torch.randn(1, 77, 768, generator=g)
B = 3  # ragged over 2 ranks
res = D.sample_sharded(eng, 42, 2, 7.0, "dpmpp_2m_cfgpp", "karras", pos, neg, {{"samples": torch.zeros(B, 4, 16, 16)}})
# ancestral sampler: the per-step noise must be the unsharded batch's rows, not the same draw on every rank
anc = D.sample_sharded(eng, 42, 2, 7.0, "euler_ancestral_cfgpp", "karras", pos, neg, {{"samples": torch.zeros(B, 4, 8, 8)}})
if rank == 0:
    torch.save({{"dpmpp_2m": res[0]["samples"], "euler_a": anc[0]["samples"]}}, {out!r})
dist.barrier(); dist.destroy_process_group()
"""


def test_two_rank_gloo_sharded_sampling_equals_single_process(tmp_path):
    """world_size=2 over gloo on CPU: weights broadcast, noise scattered, latents gathered; the sharded batch equals
    the single-process batch bit-for-bit in structure (same noise rows) and numerically in value."""
    from lightdiffusion_next_b200 import sampling as S
    from lightdiffusion_next_b200.synth import synth_tensor
    from fake_engine import FakeEngine
    from conftest import TINY_UNET
    from oracle import sd15_oracle as O
    tiny_sd = {k: synth_tensor(k, shp) for k, shp in O.unet_param_shapes(TINY_UNET).items()}
    out = str(tmp_path / "sharded.pt")
    script = tmp_path / "worker.py"
    script.write_text(_WORKER.format(root=ROOT, out=out))
    env = dict(os.environ, OMP_NUM_THREADS="4")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29531", str(script)], env=env,
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    sharded = torch.load(out)
    g = torch.Generator().manual_seed(1234)
    pos = torch.randn(1, 77, 768, generator=g)
    neg = torch.randn(1, 77, 768, generator=g)
    single = S.sample(FakeEngine(tiny_sd, unet_cfg=TINY_UNET), 42, 2, 7.0, "dpmpp_2m_cfgpp", "karras", pos, neg,
                      {"samples": torch.zeros(3, 4, 16, 16)})[0]["samples"]
    assert sharded["dpmpp_2m"].shape == single.shape
    assert rel(sharded["dpmpp_2m"], single) < 1e-5
    single_a = S.sample(FakeEngine(tiny_sd, unet_cfg=TINY_UNET), 42, 2, 7.0, "euler_ancestral_cfgpp", "karras", pos, neg,
                        {"samples": torch.zeros(3, 4, 8, 8)})[0]["samples"]
    assert rel(sharded["euler_a"], single_a) < 1e-5   # exact noise rows: shard-invariant per-step draws
    assert rel(single_a[0], single_a[1]) > 0.5        # ... and different images get different noise


def test_synth_shape_tables_match_oracle():
    """The package-side VAE / CLIP layout tables (used by bench.py, which may not import the oracle on the product arm)
    name exactly the tensors the oracle's restatement of the reference modules consumes."""
    from oracle import sd15_oracle as O
    from lightdiffusion_next_b200 import synth
    assert synth.vae_decoder_shapes() == O.vae_decoder_param_shapes()
    assert synth.clip_shapes() == O.clip_param_shapes()
    assert synth.vae_encoder_shapes() == O.vae_encoder_param_shapes()
    assert synth.taesd_decoder_shapes() == O.taesd_decoder_param_shapes()


@pytest.mark.parametrize("name", ["perf", "block"])
def test_dpmpp_2m_multiscale_options_match_reference(unet_sd, name):
    """Non-default multiscale options of sample_dpmpp_2m_cfgpp (factor 0.25 / contiguous low-resolution block / other
    full-resolution margins), as injected through the reference's `ksampler(name, extra_options)` seam
    (tests/golden/make_golden_msopts.py): oracle and the product's host loop vs the reference's final latents."""
    from fake_engine import FakeEngine
    from lightdiffusion_next_b200 import sampling as S
    from oracle import sd15_oracle as O
    g = torch.load(os.path.join(GOLDEN, "msopts_small.pt"))
    a = g[f"{name}_args"]
    lat = torch.zeros(1, 4, a["hw"], a["hw"])
    o = a["opts"]
    ref = g[f"{name}_final"]
    orc = O.ksample(unet_sd, 42, a["steps"], 7.0, "dpmpp_2m_cfgpp", "karras", g["ctx_pos"], g["ctx_neg"], lat,
                    ms_options=dict(factor=o["multiscale_factor"], start=o["multiscale_fullres_start"],
                                    end=o["multiscale_fullres_end"], intermittent=o["multiscale_intermittent_fullres"]))
    assert float((orc - ref).norm() / ref.norm()) < 1e-4
    eng = FakeEngine(unet_sd)
    e = S.sample(eng, 42, a["steps"], 7.0, "dpmpp_2m_cfgpp", "karras", g["ctx_pos"], g["ctx_neg"], {"samples": lat},
                 sampler_options=o)[0]["samples"]
    assert float((e - ref).norm() / ref.norm()) < 1e-4
    with pytest.raises(ValueError, match="unknown dpmpp_2m_cfgpp options"):
        S.sample(eng, 42, 1, 7.0, "dpmpp_2m_cfgpp", "karras", g["ctx_pos"], g["ctx_neg"], {"samples": lat},
                 sampler_options={"multiscale_typo": 1})


def test_entry_points_reject_null_arguments_with_an_error_code():
    """Error convention of include/ldn.h: non-zero return + ldn_last_error() message, never a crash -- checked on the
    argument validation every engine entry runs before touching the device (no compute call is made)."""
    from lightdiffusion_next_b200 import _lib
    lib = _lib.load()
    calls = {
        "ldn_load_weights": (None, 0, None, 0, None),
        "ldn_set_context": (None, None, 2, 77, None),
        "ldn_unet_denoise": (None, None, None, None, 2, 32, 32, None),
        "ldn_vae_decode": (None, None, None, 1, 8, 8, None),
        "ldn_vae_encode": (None, None, None, 1, 64, 64, None),
        "ldn_taesd_decode": (None, None, None, 1, 8, 8, None),
        "ldn_flux_forward": (None, None, None, None, None, None, None, None, 1, 16, 16, None),
        "ldn_clip_encode": (None, None, 1, None, None, None),
        "ldn_t5_encode": (None, None, None, 1, 16, None, None),
        "ldn_resample_bilinear": (None, None, 1, 8, 8, 4, 4, None),
        "ldn_bislerp": (None, None, None, 1, 4, 8, 8, 16, 16, None),
        "ldn_conv3x3_groupnorm_bf16": (None, None, 1, 8, 8, 64, 320, None, None, 0, None, 1e-5, None, None, 1, None, None, None, None),
    }
    for name, args in calls.items():
        rc = getattr(lib, name)(*args)
        assert rc != 0, name
        msg = lib.ldn_last_error()
        assert msg and b"bad argument" in msg, (name, msg)
        with pytest.raises(_lib.LdnError, match="bad argument"):
            _lib.check(rc)


def test_sampler_registry_seam_function(unet_sd):
    """backend.engine_sampler_function: the reference's sampler-function signature (KSAMPLER.sample -> fn(model_k, x, sigmas,
    extra_args=, callback=, disable=, pipeline=, **extra_options), sampling.py:445-534) driven the way KSAMPLER.sample drives it
    -- noise-scaled x in, latent-space x out -- reproduces the reference's final latents (golden of the dpmpp_2m run)."""
    import types
    from fake_engine import FakeEngine
    from lightdiffusion_next_b200 import backend, sampling as S
    from lightdiffusion_next_b200.schedule import calculate_sigmas
    g = torch.load(os.path.join(GOLDEN, "msopts_small.pt"))
    a = g["block_args"]
    eng = FakeEngine(unet_sd)
    cond = lambda t: [{"model_conds": {"c_crossattn": types.SimpleNamespace(cond=t)}}]
    guider = types.SimpleNamespace(conds={"positive": cond(g["ctx_pos"]), "negative": cond(g["ctx_neg"])}, cfg=7.0)
    model_k = types.SimpleNamespace(inner_model=guider)
    lat = torch.zeros(1, 4, a["hw"], a["hw"])
    sigmas = calculate_sigmas(eng.schedule, "karras", a["steps"])
    x = S.prepare_noise(lat, 42) * torch.sqrt(1.0 + sigmas[0] ** 2.0)       # noise_scaling at max denoise (KSAMPLER.sample)
    fn = backend.engine_sampler_function(eng, "dpmpp_2m_cfgpp")
    seen = []
    out = fn(model_k, x, sigmas, extra_args={"seed": 42}, callback=lambda d: seen.append(d["i"]), disable=True, pipeline=True,
             **a["opts"])
    ref = g["block_final"]                                                    # process_latent_out applied by the reference
    assert float((out / S.LATENT_SCALE - ref).norm() / ref.norm()) < 1e-4
    assert seen == list(range(a["steps"]))
    with pytest.raises(ValueError, match="unknown dpmpp_2m_cfgpp options"):
        fn(model_k, x, sigmas, bogus=1)
    with pytest.raises(ValueError, match="not built"):
        backend.engine_sampler_function(eng, "euler")


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the CPU arm of the measurement contract): one JSON line with the contract's keys, rank 0
    only; non-zero ranks print nothing and exit 0.  Run at a small size so the CPU suite stays short."""
    import json
    cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0", "--size", "128"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env={**os.environ, "RANK": "0"})
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["unit"] == "it/s" and d["value"] > 0 and d["vs_baseline"] is None
    # "reference" = the unmodified reference from the baseline/_ref snapshot (build() makes it where /root/reference exists),
    # "port" = the oracle restatement when the snapshot is absent
    have_ref = os.path.isdir(os.path.join(ROOT, "baseline", "_ref", "reference", "src"))
    assert d["cpu_baseline"]["kind"] == ("reference" if have_ref else "port")
    assert d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["sample"]
    assert d["steps"] == 1 and d["warmup"] == 0
    assert d["e2e"] == {"value": d["value"], "unit": "it/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]
    r1 = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env={**os.environ, "RANK": "1"})
    assert r1.returncode == 0 and not [l for l in r1.stdout.splitlines() if l.startswith("{")]


def test_interrupt_poll_returns_the_current_latent(unet_sd):
    """The reference's samplers poll app.interrupt_flag before every step (unless pipeline=True) and hand back the current x
    (samplers.py:884-889).  Same contract here: `interrupt()` is polled before each step; stopping before step k returns
    exactly the latent an uninterrupted run holds after step k - 1; pipeline=True at the registry seam disables the poll."""
    import types
    from fake_engine import FakeEngine
    from lightdiffusion_next_b200 import backend, sampling as S
    g = torch.load(os.path.join(GOLDEN, "msopts_small.pt"))
    lat = torch.zeros(1, 4, 16, 16)
    states = []
    S.sample(FakeEngine(unet_sd), 42, 3, 7.0, "euler_ancestral_cfgpp", "karras", g["ctx_pos"], g["ctx_neg"], {"samples": lat},
             noise_sampler=lambda s0, s1: torch.zeros(1, 4, 16, 16), callback=lambda d: states.append(d["x"].clone()))
    polls = []
    eng = FakeEngine(unet_sd)
    out = S.sample(eng, 42, 3, 7.0, "euler_ancestral_cfgpp", "karras", g["ctx_pos"], g["ctx_neg"], {"samples": lat},
                   noise_sampler=lambda s0, s1: torch.zeros(1, 4, 16, 16), interrupt=lambda: (polls.append(1), len(polls) > 2)[1])[0]["samples"]
    assert len(polls) == 3 and eng.denoise_calls == 2           # stopped before the third step
    assert torch.equal(out, states[1] / S.LATENT_SCALE)
    # registry seam: polled when pipeline=False, ignored when pipeline=True (as in the reference)
    cond = lambda t: [{"model_conds": {"c_crossattn": types.SimpleNamespace(cond=t)}}]
    model_k = types.SimpleNamespace(inner_model=types.SimpleNamespace(
        conds={"positive": cond(g["ctx_pos"]), "negative": cond(g["ctx_neg"])}, cfg=7.0))
    sig = torch.tensor([2.0, 1.0, 0.0])
    x0 = torch.randn(1, 4, 16, 16, generator=torch.Generator().manual_seed(0))
    fn = backend.engine_sampler_function(FakeEngine(unet_sd), "dpmpp_2m_cfgpp", interrupt=lambda: True)
    assert torch.equal(fn(model_k, x0, sig, pipeline=False), x0)
    assert not torch.equal(fn(model_k, x0, sig, pipeline=True), x0)


def test_euler_ancestral_injected_noise_sampler_matches_reference(unet_sd):
    """euler_ancestral_cfgpp with the noise sampler injected the way the reference's seam injects it
    (ksampler(name, extra_options={"noise_sampler": f}); called as f(sigma, sigma_next), samplers.py:732): the product's host
    loop reproduces the reference's final latent and calls f with the same sigma pairs (not after the last step)."""
    from fake_engine import FakeEngine
    from lightdiffusion_next_b200 import sampling as S
    g = torch.load(os.path.join(GOLDEN, "msopts_small.pt"))

    class SeqNoise:
        def __init__(self, shape, seed):
            self.g, self.shape, self.calls = torch.Generator().manual_seed(seed), shape, []

        def __call__(self, sigma, sigma_next):
            self.calls.append((float(sigma), float(sigma_next)))
            return torch.randn(self.shape, generator=self.g)

    lat = torch.zeros(1, 4, 16, 16)
    ns = SeqNoise(lat.shape, 7)
    e = S.sample(FakeEngine(unet_sd), 42, 3, 7.0, "euler_ancestral_cfgpp", "karras", g["ctx_pos"], g["ctx_neg"], {"samples": lat},
                 noise_sampler=ns)[0]["samples"]
    ref = g["anc_final"]
    assert float((e - ref).norm() / ref.norm()) < 1e-4
    assert torch.allclose(torch.tensor(ns.calls), g["anc_calls"], rtol=1e-5)
    # eta / s_noise, through the registry-seam function with the reference's extra_options
    import types
    from lightdiffusion_next_b200 import backend
    from lightdiffusion_next_b200.schedule import calculate_sigmas
    eng = FakeEngine(unet_sd)
    cond = lambda t: [{"model_conds": {"c_crossattn": types.SimpleNamespace(cond=t)}}]
    model_k = types.SimpleNamespace(inner_model=types.SimpleNamespace(
        conds={"positive": cond(g["ctx_pos"]), "negative": cond(g["ctx_neg"])}, cfg=7.0))
    sigmas = calculate_sigmas(eng.schedule, "karras", 3)
    x = S.prepare_noise(lat, 42) * torch.sqrt(1.0 + sigmas[0] ** 2.0)
    out = backend.engine_sampler_function(eng, "euler_ancestral_cfgpp")(
        model_k, x, sigmas, pipeline=True, noise_sampler=SeqNoise(lat.shape, 7), eta=0.6, s_noise=1.1)
    ref = g["anc_eta_final"]
    assert float((out / S.LATENT_SCALE - ref).norm() / ref.norm()) < 1e-4


def test_schedules_sweep_product_equals_oracle_and_is_well_formed():
    """Every scheduler name x 1..40 steps: the product's host schedule equals the (reference-pinned) oracle's bit for bit,
    starts at or below sigma_max, decreases strictly and ends in exactly 0; same for the Flux table with `beta` / `simple`."""
    from lightdiffusion_next_b200.schedule import DiscreteSchedule, FluxSchedule, calculate_sigmas
    from oracle import sd15_oracle as O
    sched = DiscreteSchedule()
    for name in ("karras", "normal", "simple", "beta"):
        for steps in range(1, 41):
            s = calculate_sigmas(sched, name, steps)
            assert torch.equal(s, O.calculate_sigmas(name, steps)), (name, steps)
            assert s[-1] == 0 and s.dtype == torch.float32 and 2 <= len(s) <= steps + 1, (name, steps)
            assert bool((s[:-1] > s[1:]).all()), (name, steps)
            assert float(s[0]) <= 14.6147
    fs = FluxSchedule(1.15)
    for name in ("beta", "simple"):
        for steps in (1, 2, 4, 20, 50):
            s = calculate_sigmas(fs, name, steps)
            assert s[-1] == 0 and float(s[0]) <= 1.0 and bool((s[:-1] > s[1:]).all()), (name, steps)


def test_t5_bucket_properties():
    """Relative-position buckets: 16 buckets per direction, exact below 8, logarithmic up to 128, saturating beyond; mirrored
    distances differ by exactly the direction offset; non-decreasing in |distance|."""
    from lightdiffusion_next_b200 import t5 as T5H
    n = 700
    b = T5H.relative_position_buckets(n).long()
    c = n - 1                                    # index of distance 0
    assert int(b[c]) == 0 and b.min() == 0 and b.max() == 31
    d = torch.arange(1, n)
    assert torch.equal(b[c + d], b[c - d] + 16)  # key after query = same magnitude bucket + 16
    mag = b[c - d]
    assert torch.equal(mag[:7], torch.arange(1, 8)) and bool((mag[1:] >= mag[:-1]).all())
    assert int(mag[89]) == 14 and bool((mag[90:] == 15).all())    # last bucket: |distance| >= 8 * 16^(7/8) = 90.5, incl. everything >= 128
