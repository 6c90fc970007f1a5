"""Generates tests/golden/*.pt by running the UNMODIFIED reference (imported read-only from /root/reference).

Runs only in the build container (the GPU box has no /root/reference); the fixtures it writes are committed.
Recipe = SURVEY.md Appendix C.  Usage:  python tests/golden/make_golden.py [unet] [sample] [vae] [clip]
"""
import os
import sys
import tempfile
import types

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)
sys.modules.setdefault("torchsde", types.ModuleType("torchsde"))

# the reference needs ./include/clip/sd1_clip_config.json, ./include/sd1_tokenizer/ and a writable cwd
work = tempfile.mkdtemp(prefix="ldn_golden_")
os.makedirs(os.path.join(work, "include"), exist_ok=True)
for sub in ("clip", "sd1_tokenizer"):
    os.symlink(os.path.join(REF, "include", sub), os.path.join(work, "include", sub))
os.chdir(work)

from oracle import sd15_oracle as O  # noqa: E402

torch.manual_seed(0)
torch.set_grad_enabled(False)
what = set(sys.argv[1:]) or {"unet", "sample", "vae", "clip"}


def build_ref_unet():
    from src.user import app_instance
    app_instance.app.previewer_var.set(False)
    from src.NeuralNetwork import unet
    from src.Device import Device
    from src.Model import ModelPatcher
    cfg = dict(image_size=32, in_channels=4, out_channels=4, model_channels=320, num_res_blocks=[2, 2, 2, 2],
               channel_mult=[1, 2, 4, 4], transformer_depth=[1, 1, 1, 1, 1, 1, 0, 0],
               transformer_depth_output=[1] * 9 + [0] * 3, transformer_depth_middle=1,
               use_linear_in_transformer=False, context_dim=768, use_spatial_transformer=True, legacy=False,
               use_checkpoint=False, adm_in_channels=None, use_temporal_attention=False, use_temporal_resblock=False)
    mc = unet.model_config_from_unet_config(cfg)
    dt = unet.unet_dtype1()
    mc.set_inference_dtype(dt, Device.unet_manual_cast(dt, Device.get_torch_device(), mc.supported_inference_dtypes))
    model = mc.get_model({}, "", device=torch.device("cpu"))
    shapes = O.unet_param_shapes()
    ref_shapes = {k: tuple(v.shape) for k, v in model.diffusion_model.state_dict().items()}
    assert ref_shapes == shapes, (set(ref_shapes) ^ set(shapes))
    sd = O.synth_state_dict(shapes)
    model.diffusion_model.load_state_dict(sd, strict=True)
    mp = ModelPatcher.ModelPatcher(model, load_device=torch.device("cpu"), offload_device=torch.device("cpu"))
    return model, mp, sd


if "unet" in what or "sample" in what:
    model, mp, sd = build_ref_unet()
    g = torch.Generator().manual_seed(1234)
    ctx_pos = torch.randn(1, 77, 768, generator=g)
    ctx_neg = torch.randn(1, 77, 768, generator=g)

if "unet" in what:
    out = {}
    # sigma tables of the reference's ModelSamplingDiscrete
    ms = model.model_sampling
    out["sigmas"] = ms.sigmas.clone()
    out["log_sigmas"] = ms.log_sigmas.clone()
    from src.sample import ksampler_util
    for name, steps in (("karras", 20), ("karras", 30), ("normal", 10), ("normal", 22)):
        out[f"sched_{name}_{steps}"] = ksampler_util.calculate_sigmas(ms, name, steps).clone()
    # apply_model at two small latent sizes (rows: uncond first, cond second as calc_cond_batch batches them)
    for hw in (16, 32):
        gx = torch.Generator().manual_seed(7 + hw)
        x = torch.randn(2, 4, hw, hw, generator=gx) * 3.0
        sigma = torch.tensor([2.5, 0.7])
        ctx = torch.cat([ctx_neg, ctx_pos])
        den = model.apply_model(x, sigma, c_crossattn=ctx, transformer_options={})
        out[f"apply_x_{hw}"] = x
        out[f"apply_sigma_{hw}"] = sigma
        out[f"apply_ctx_{hw}"] = ctx
        out[f"apply_out_{hw}"] = den.float().clone()
        out[f"apply_t_{hw}"] = ms.timestep(sigma).float()
        print("apply_model", hw, float(den.std()), flush=True)
    torch.save(out, os.path.join(HERE, "unet_small.pt"))
    print("wrote unet_small.pt")

if "sample" in what:
    from src.sample import sampling
    out = {"ctx_pos": ctx_pos, "ctx_neg": ctx_neg}

    class Capture:
        """model_function_wrapper seam (src/cond/cond.py:254-265): pass-through that records one call."""
        def __init__(self):
            self.rec = None
        def __call__(self, model_function, params):
            r = model_function(params["input"], params["timestep"], **params["c"])
            if self.rec is None:
                self.rec = dict(input=params["input"].clone(), timestep=params["timestep"].clone(),
                                ctx=params["c"]["c_crossattn"].clone(),
                                cond_or_uncond=torch.tensor(params["cond_or_uncond"]), output=r.clone())
            return r
        def to(self, *_):
            return self

    pos = [[ctx_pos, {}]]
    neg = [[ctx_neg, {}]]
    for name, sampler, sched, steps, hw in (("euler_a", "euler_ancestral_cfgpp", "karras", 4, 16),
                                            ("dpmpp_2m", "dpmpp_2m_cfgpp", "karras", 6, 16),
                                            ("dpmpp_2m_ms", "dpmpp_2m_cfgpp", "karras", 15, 16),
                                            ("euler_a_normal", "euler_ancestral_cfgpp", "normal", 3, 16)):
        mp2 = mp.clone()
        cap = Capture()
        mp2.set_model_unet_function_wrapper(cap)
        res = sampling.KSampler().sample(model=mp2, seed=42, steps=steps, cfg=7.0, sampler_name=sampler,
                                         scheduler=sched, denoise=1.0, positive=pos, negative=neg,
                                         latent_image={"samples": torch.zeros(1, 4, hw, hw)}, pipeline=True)
        out[f"{name}_final"] = res[0]["samples"].clone()
        for k, v in cap.rec.items():
            out[f"{name}_seam_{k}"] = v
        print(name, float(res[0]["samples"].std()), flush=True)
    torch.save(out, os.path.join(HERE, "sample_small.pt"))
    print("wrote sample_small.pt")

if "vae" in what:
    from src.AutoEncoders import VariationalAE as V
    shapes = O.vae_decoder_param_shapes()
    sd = {k: v.float() for k, v in O.synth_state_dict(shapes, seed=4321).items()}
    # the reference VAE also needs encoder-side keys to construct; give it its own (unused here) random ones
    dd = dict(double_z=True, z_channels=4, resolution=256, in_channels=3, out_ch=3, ch=128, ch_mult=[1, 2, 4, 4],
              num_res_blocks=2, attn_resolutions=[], dropout=0.0)
    eng = V.AutoencodingEngine(V.Encoder(**dd), V.Decoder(**dd), V.DiagonalGaussianRegularizer())
    full = eng.state_dict()
    ref_dec = {k: tuple(v.shape) for k, v in full.items() if k.startswith("decoder.") or k.startswith("post_quant_conv")}
    assert ref_dec == shapes, (set(ref_dec) ^ set(shapes))
    g = torch.Generator().manual_seed(99)
    for k in full:
        full[k] = sd[k] if k in sd else torch.randn(full[k].shape, generator=g) * 0.02
    vae = V.VAE(sd=full)
    out = {}
    for hw in (8, 16):
        z = torch.randn(2, 4, hw, hw, generator=g)
        img = V.VAEDecode().decode(vae, {"samples": z})[0]
        out[f"z_{hw}"] = z
        out[f"img_{hw}"] = img.float().clone()
        print("vae", hw, tuple(img.shape), float(img.mean()), float(img.std()), flush=True)
    torch.save(out, os.path.join(HERE, "vae_small.pt"))
    print("wrote vae_small.pt")

if "clip" in what:
    from src.SD15 import SDClip, SDToken
    shapes = O.clip_param_shapes()
    sd = O.synth_state_dict(shapes, seed=777)
    tok = SDToken.SD1Tokenizer(tokenizer=lambda embedding_directory=None: SDToken.SDTokenizer(
        tokenizer_path=os.path.join(REF, "include", "sd1_tokenizer/"), embedding_directory=embedding_directory))
    m = SDClip.SD1ClipModel(device="cpu", dtype=torch.float16)
    ref_shapes = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    pref = "clip_l.transformer.text_model."
    mine = {pref + k: v for k, v in shapes.items()}
    missing = set(mine) - set(ref_shapes)
    assert not missing, missing
    extra = {k for k in ref_shapes if k not in mine}
    print("reference-only CLIP keys:", sorted(extra))
    full = m.state_dict()
    for k, v in sd.items():
        assert tuple(full[pref + k].shape) == tuple(v.shape), k
        full[pref + k] = v
    m.load_state_dict(full)
    m.set_clip_options({"layer": -2})
    out = {}
    for name, text in (("plain", "a photograph of an astronaut riding a horse"), ("empty", ""),
                       ("weighted", "a (red:1.3) cube on a (blue:0.7) sphere")):
        tokens = tok.tokenize_with_weights(text)
        cond, pooled = m.encode_token_weights(tokens)
        ids = torch.tensor([[t for t, _ in row] for row in tokens["l"]], dtype=torch.int64)
        wts = torch.tensor([[w for _, w in row] for row in tokens["l"]], dtype=torch.float32)
        out[f"{name}_ids"] = ids
        out[f"{name}_weights"] = wts
        out[f"{name}_cond"] = cond.float().clone()
        print("clip", name, tuple(ids.shape), tuple(cond.shape), float(cond.std()), flush=True)
    torch.save(out, os.path.join(HERE, "clip_small.pt"))
    print("wrote clip_small.pt")
