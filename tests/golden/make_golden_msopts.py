"""Generates tests/golden/msopts_small.pt: reference runs of sample_dpmpp_2m_cfgpp with NON-default multiscale options
(samplers.py:768-773) injected through the sampler seam `ksampler(name, extra_options=...)` (sampling.py:500-534) -- the
only way they reach the sampler in the reference -- on the SD1.5 UNet with seeded synthetic weights."""
import os
import sys
import tempfile
import types

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT); sys.path.insert(0, REF)
sys.modules.setdefault("torchsde", types.ModuleType("torchsde"))
work = tempfile.mkdtemp(prefix="ldn_golden_")
os.makedirs(os.path.join(work, "include"), exist_ok=True)
for sub in ("clip", "sd1_tokenizer"):
    os.symlink(os.path.join(REF, "include", sub), os.path.join(work, "include", sub))
os.chdir(work)
from oracle import sd15_oracle as O  # noqa: E402

torch.manual_seed(0); torch.set_grad_enabled(False)
from src.user import app_instance  # noqa: E402
app_instance.app.previewer_var.set(False)
from src.NeuralNetwork import unet  # noqa: E402
from src.Device import Device  # noqa: E402
from src.Model import ModelPatcher  # noqa: E402
from src.sample import ksampler_util, sampling  # noqa: E402

cfg = dict(image_size=32, in_channels=4, out_channels=4, model_channels=320, num_res_blocks=[2, 2, 2, 2],
           channel_mult=[1, 2, 4, 4], transformer_depth=[1, 1, 1, 1, 1, 1, 0, 0],
           transformer_depth_output=[1] * 9 + [0] * 3, transformer_depth_middle=1,
           use_linear_in_transformer=False, context_dim=768, use_spatial_transformer=True, legacy=False,
           use_checkpoint=False, adm_in_channels=None, use_temporal_attention=False, use_temporal_resblock=False)
mc = unet.model_config_from_unet_config(cfg); dt = unet.unet_dtype1()
mc.set_inference_dtype(dt, Device.unet_manual_cast(dt, Device.get_torch_device(), mc.supported_inference_dtypes))
model = mc.get_model({}, "", device=torch.device("cpu"))
model.diffusion_model.load_state_dict(O.synth_state_dict(O.unet_param_shapes()), strict=True)
mp = ModelPatcher.ModelPatcher(model, load_device=torch.device("cpu"), offload_device=torch.device("cpu"))
g = torch.Generator().manual_seed(1234)
ctx_pos = torch.randn(1, 77, 768, generator=g); ctx_neg = torch.randn(1, 77, 768, generator=g)
out = {"ctx_pos": ctx_pos, "ctx_neg": ctx_neg}
cases = {
    # "performance" preset values (factor 0.25) on a 24x24 latent: low-res steps at 8x8
    "perf": (6, 24, dict(multiscale_factor=0.25, multiscale_fullres_start=2, multiscale_fullres_end=2, multiscale_intermittent_fullres=True)),
    # contiguous low-res block (no intermittent full-res steps), "quality"-like late start
    "block": (7, 16, dict(multiscale_factor=0.5, multiscale_fullres_start=3, multiscale_fullres_end=1, multiscale_intermittent_fullres=False)),
}
for name, (steps, hw, opts) in cases.items():
    lat = torch.zeros(1, 4, hw, hw)
    sampler = sampling.ksampler("dpmpp_2m_cfgpp", extra_options=dict(opts))
    sigmas = ksampler_util.calculate_sigmas(model.model_sampling, "karras", steps)
    noise = ksampler_util.prepare_noise(lat, 42)
    res = sampling.sample(mp, noise, [[ctx_pos, {}]], [[ctx_neg, {}]], 7.0, torch.device("cpu"), sampler, sigmas,
                          latent_image=lat, seed=42, pipeline=True)
    out[f"{name}_final"] = res.clone(); out[f"{name}_args"] = dict(steps=steps, hw=hw, opts=opts)
    print(name, tuple(res.shape), float(res.std()), flush=True)


class SeqNoise:
    """n-th call returns the n-th draw of a seeded CPU generator; records the (sigma, sigma_next) it was called with."""
    def __init__(self, shape, seed): self.g = torch.Generator().manual_seed(seed); self.shape = shape; self.calls = []
    def __call__(self, sigma, sigma_next):
        self.calls.append((float(sigma), float(sigma_next)))
        return torch.randn(self.shape, generator=self.g)


# euler_ancestral_cfgpp with an injected noise sampler (extra_options={"noise_sampler": f}; samplers.py:621-636, 732)
lat = torch.zeros(1, 4, 16, 16)
ns = SeqNoise(lat.shape, 7)
sampler = sampling.ksampler("euler_ancestral_cfgpp", extra_options={"noise_sampler": ns})
sigmas = ksampler_util.calculate_sigmas(model.model_sampling, "karras", 3)
res = sampling.sample(mp, ksampler_util.prepare_noise(lat, 42), [[ctx_pos, {}]], [[ctx_neg, {}]], 7.0, torch.device("cpu"), sampler,
                      sigmas, latent_image=lat, seed=42, pipeline=True)
out["anc_final"] = res.clone(); out["anc_calls"] = torch.tensor(ns.calls)
print("anc", float(res.std()), ns.calls, flush=True)
# ... and with eta / s_noise (the sampler's own keyword arguments, samplers.py:619-620)
ns = SeqNoise(lat.shape, 7)
sampler = sampling.ksampler("euler_ancestral_cfgpp", extra_options={"noise_sampler": ns, "eta": 0.6, "s_noise": 1.1})
res = sampling.sample(mp, ksampler_util.prepare_noise(lat, 42), [[ctx_pos, {}]], [[ctx_neg, {}]], 7.0, torch.device("cpu"), sampler,
                      sigmas, latent_image=lat, seed=42, pipeline=True)
out["anc_eta_final"] = res.clone()
print("anc_eta", float(res.std()), flush=True)
torch.save(out, os.path.join(HERE, "msopts_small.pt")); print("wrote msopts_small.pt")
