#!/usr/bin/env python
"""Times the UNMODIFIED reference (LightDiffusion-Next) on this box -- measurement harness, not product code.

The reference is pure Python; `__graft_entry__.build()` snapshots its `src/` and `include/` trees into `baseline/_ref/`
(git-ignored, never committed; it travels to the GPU box with the gpurun snapshot exactly like a built .so).  This script
imports that copy, fills the reference's own `UNetModel1` with the seeded synthetic SD1.5 weights the engine is benchmarked
with, and drives the reference's own sampling stack:

    sampling.ksampler("dpmpp_2m_cfgpp", extra_options={"enable_multiscale": False})  +  sampling.sample(...)
        (src/sample/sampling.py:500-590: the entry one level below KSampler.sample, the only public way to switch the
         half-resolution steps off for this sampler -- SURVEY fact 9), or KSampler.sample itself (`--default-schedule`).

  --device cpu    the reference's CPU path as it runs (fp16-stored weights, fp32 compute through manual_cast) on all host
                  cores: the `--impl reference` arm of bench.py and its `cpu_baseline`.
  --device cuda   the reference's GPU path (fp16 UNet, torch SDPA: xformers is not installable here, Attention.py:34-41),
                  the comparator north_star names (>= 1.5x).  bench.py's `secondary.gpu_reference`.
  --seam          same call, but with lightdiffusion_next_b200.backend.install() applied to the ModelPatcher first: the
                  reference's own loop with every UNet call answered by the B200 engine through model_function_wrapper.

Prints one JSON line.  Timing: a warm-up run of --warmup steps (model load, cuDNN autotune, graph capture), then a run of
--steps steps wall-clocked with a device synchronise on both sides; it/s = steps / seconds, i.e. what tqdm would report.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import tempfile
import time
import types

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.path.join(HERE, "_ref", "reference")

SD15_UNET_CONFIG = dict(image_size=32, in_channels=4, out_channels=4, model_channels=320, num_res_blocks=[2, 2, 2, 2],
                        channel_mult=[1, 2, 4, 4], transformer_depth=[1, 1, 1, 1, 1, 1, 0, 0],
                        transformer_depth_output=[1] * 9 + [0] * 3, transformer_depth_middle=1,
                        use_linear_in_transformer=False, context_dim=768, use_spatial_transformer=True, legacy=False,
                        use_checkpoint=False, adm_in_channels=None, use_temporal_attention=False,
                        use_temporal_resblock=False)


def available() -> bool:
    return os.path.isdir(os.path.join(REF, "src"))


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--device", default="cpu", choices=["cpu", "cuda"])
    ap.add_argument("--size", type=int, default=1024)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--bs", type=int, default=1)
    ap.add_argument("--seam", action="store_true")
    ap.add_argument("--default-schedule", action="store_true", help="KSampler.sample as is (reference-default multiscale)")
    ap.add_argument("--max-seconds", type=float, default=0.0, help="stop the timed run after this many seconds (CPU arm)")
    args = ap.parse_args()
    if not available():
        print(json.dumps({"unavailable": "baseline/_ref/reference missing (run __graft_entry__.build() where /root/reference exists)"}))
        return
    if args.device == "cpu":
        os.environ["CUDA_VISIBLE_DEVICES"] = ""  # the reference picks cuda whenever it sees one (Device.py:73-95)
    import torch

    for p in (ROOT, REF):
        if p not in sys.path:
            sys.path.insert(0, p)
    sys.modules.setdefault("torchsde", types.ModuleType("torchsde"))  # only dpmpp_sde's default noise sampler needs it
    work = tempfile.mkdtemp(prefix="ldn_ref_")
    os.makedirs(os.path.join(work, "include"), exist_ok=True)
    for sub in ("clip", "sd1_tokenizer"):
        os.symlink(os.path.join(REF, "include", sub), os.path.join(work, "include", sub))
    os.chdir(work)  # the reference writes ./output/preview and reads ./include/... relative to the cwd
    torch.set_grad_enabled(False)
    if args.device == "cpu":
        torch.set_num_threads(os.cpu_count() or 1)
    from src.user import app_instance
    app_instance.app.previewer_var.set(False)
    from src.Device import Device
    from src.Model import ModelPatcher
    from src.NeuralNetwork import unet
    from src.sample import ksampler_util, sampling

    from lightdiffusion_next_b200.synth import synth_state_dict, unet_shapes

    dev = Device.get_torch_device()
    assert dev.type == args.device, (dev, args.device)
    mc = unet.model_config_from_unet_config(dict(SD15_UNET_CONFIG))
    dt = unet.unet_dtype1()
    manual = Device.unet_manual_cast(dt, dev, mc.supported_inference_dtypes)
    mc.set_inference_dtype(dt, manual)
    model = mc.get_model({}, "", device=torch.device("cpu"))
    sd = synth_state_dict(unet_shapes())
    model.diffusion_model.load_state_dict(sd, strict=True)
    mp = ModelPatcher.ModelPatcher(model, load_device=dev, offload_device=Device.unet_offload_device())
    eng = None
    if args.seam:
        from lightdiffusion_next_b200 import backend
        from lightdiffusion_next_b200.engine import Engine
        lat_ = args.size // 8
        eng = Engine(max_rows=2 * args.bs, max_h=lat_, max_w=lat_, max_ctx_tokens=77)
        eng.load_unet(backend.unet_state_dict_from_model(model, mp))
        mp = backend.install(mp, engine=eng)
    lat = args.size // 8
    g = torch.Generator().manual_seed(1234)
    pos = [[torch.randn(1, 77, 768, generator=g), {}]]
    neg = [[torch.randn(1, 77, 768, generator=g), {}]]
    latent = torch.zeros(args.bs, 4, lat, lat)

    def sync():
        if args.device == "cuda":
            torch.cuda.synchronize()

    def run(steps: int):
        if args.default_schedule:
            return sampling.KSampler().sample(model=mp, seed=42, steps=steps, cfg=7.0, sampler_name="dpmpp_2m_cfgpp",
                                              scheduler="karras", denoise=1.0, positive=pos, negative=neg,
                                              latent_image={"samples": latent}, pipeline=True)[0]["samples"]
        sampler = sampling.ksampler("dpmpp_2m_cfgpp", pipeline=True, extra_options={"enable_multiscale": False})
        sigmas = ksampler_util.calculate_sigmas(model.model_sampling, "karras", steps).to(dev)
        noise = ksampler_util.prepare_noise(latent, 42)
        return sampling.sample(mp, noise, pos, neg, 7.0, dev, sampler, sigmas, mp.model_options, latent_image=latent,
                               seed=42, pipeline=True, disable_pbar=True)

    with torch.inference_mode():
        done_w = 0
        if args.warmup > 0:
            run(args.warmup)
            done_w = args.warmup
        sync()
        t0 = time.perf_counter()
        out = run(args.steps)
        sync()
        dt_s = time.perf_counter() - t0
    sdpa = "torch SDPA (xformers absent)" if not Device.xformers_enabled() else "xformers"
    line = {"impl": "reference", "device": args.device, "seam": bool(args.seam), "size": args.size, "bs": args.bs,
            "steps": args.steps, "warmup": done_w, "seconds": dt_s, "it_per_s": args.steps / dt_s,
            "ms_per_step": 1000.0 * dt_s / args.steps, "unet_dtype": str(dt), "manual_cast": str(manual), "attention": sdpa,
            "schedule": "reference default (multiscale)" if args.default_schedule else "multiscale off (every step full resolution)",
            "threads": torch.get_num_threads() if args.device == "cpu" else None,
            "finite": bool(torch.isfinite(out).all().item()), "torch": torch.__version__,
            "context_uploads": getattr(eng, "context_uploads", None)}
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
