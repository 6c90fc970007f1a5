"""Generates tests/golden/sde_small.pt from the UNMODIFIED reference: dpmpp_sde_cfgpp (the pipeline default sampler) driven
through sampling.ksampler(..., extra_options={"noise_sampler": f}) with a deterministic injected noise sampler (the default
BrownianTree needs torchsde, which is not installed — SURVEY.md §8c).  Build container only."""
import os, sys, tempfile, types
import torch
HERE = os.path.dirname(os.path.abspath(__file__)); ROOT = os.path.dirname(os.path.dirname(HERE)); REF = "/root/reference"
sys.path.insert(0, ROOT); sys.path.insert(0, REF)
sys.modules.setdefault("torchsde", types.ModuleType("torchsde"))
work = tempfile.mkdtemp(prefix="ldn_golden_"); os.makedirs(os.path.join(work, "include"), exist_ok=True)
for sub in ("clip", "sd1_tokenizer"):
    os.symlink(os.path.join(REF, "include", sub), os.path.join(work, "include", sub))
os.chdir(work)
from oracle import sd15_oracle as O
torch.set_grad_enabled(False)
from src.user import app_instance; app_instance.app.previewer_var.set(False)
from src.NeuralNetwork import unet; from src.Device import Device; from src.Model import ModelPatcher
from src.sample import sampling, ksampler_util
cfg = dict(image_size=32, in_channels=4, out_channels=4, model_channels=320, num_res_blocks=[2, 2, 2, 2], channel_mult=[1, 2, 4, 4],
           transformer_depth=[1, 1, 1, 1, 1, 1, 0, 0], transformer_depth_output=[1] * 9 + [0] * 3, transformer_depth_middle=1,
           use_linear_in_transformer=False, context_dim=768, use_spatial_transformer=True, legacy=False, use_checkpoint=False,
           adm_in_channels=None, use_temporal_attention=False, use_temporal_resblock=False)
mc = unet.model_config_from_unet_config(cfg); dt = unet.unet_dtype1()
mc.set_inference_dtype(dt, Device.unet_manual_cast(dt, Device.get_torch_device(), mc.supported_inference_dtypes))
model = mc.get_model({}, "", device=torch.device("cpu"))
model.diffusion_model.load_state_dict(O.synth_state_dict(O.unet_param_shapes()), strict=True)
mp = ModelPatcher.ModelPatcher(model, load_device=torch.device("cpu"), offload_device=torch.device("cpu"))
g = torch.Generator().manual_seed(1234)
ctx_pos = torch.randn(1, 77, 768, generator=g); ctx_neg = torch.randn(1, 77, 768, generator=g)
out = {"ctx_pos": ctx_pos, "ctx_neg": ctx_neg}

class SeqNoise:
    """Deterministic stand-in for the Brownian tree: the n-th call returns the n-th draw of a seeded CPU generator."""
    def __init__(self, shape, seed): self.g = torch.Generator().manual_seed(seed); self.shape = shape; self.calls = []
    def __call__(self, sigma, sigma_next):
        self.calls.append((float(sigma), float(sigma_next)))
        return torch.randn(self.shape, generator=self.g)

for name, steps, ms in (("sde", 4, False), ("sde_ms", 15, True)):
    lat = torch.zeros(1, 4, 16, 16)
    ns = SeqNoise(lat.shape, 99)
    sampler = sampling.ksampler("dpmpp_sde_cfgpp", extra_options={"noise_sampler": ns, "enable_multiscale": ms})
    sigmas = ksampler_util.calculate_sigmas(model.model_sampling, "karras", steps)
    noise = ksampler_util.prepare_noise(lat, 42)
    res = sampling.sample(mp, noise, [[ctx_pos, {}]], [[ctx_neg, {}]], 7.0, torch.device("cpu"), sampler, sigmas,
                          latent_image=lat, seed=42, pipeline=True)
    out[f"{name}_final"] = res.clone(); out[f"{name}_calls"] = torch.tensor(ns.calls)
    print(name, tuple(res.shape), float(res.std()), len(ns.calls))
torch.save(out, os.path.join(HERE, "sde_small.pt")); print("wrote sde_small.pt")
