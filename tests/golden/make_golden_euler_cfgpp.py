"""Generates tests/golden/euler_cfgpp_small.pt: full KSampler.sample runs of the UNMODIFIED reference with
sampler_name="euler_cfgpp" (sample_euler_dy_cfg_pp incl. its dynamic half-resolution steps, src/sample/samplers.py:362-608)
on the SD1.5 UNet with seeded synthetic weights (build container only)."""
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
work = tempfile.mkdtemp(prefix="ldn_golden_")
os.makedirs(os.path.join(work, "include"), exist_ok=True)
for sub in ("clip", "sd1_tokenizer"):
    os.symlink(os.path.join(REF, "include", sub), os.path.join(work, "include", sub))
os.chdir(work)
from oracle import sd15_oracle as O  # noqa: E402

torch.manual_seed(0)
torch.set_grad_enabled(False)
from src.user import app_instance  # noqa: E402
app_instance.app.previewer_var.set(False)
from src.NeuralNetwork import unet  # noqa: E402
from src.Device import Device  # noqa: E402
from src.Model import ModelPatcher  # noqa: E402
from src.sample import sampling  # noqa: E402

cfg = dict(image_size=32, in_channels=4, out_channels=4, model_channels=320, num_res_blocks=[2, 2, 2, 2],
           channel_mult=[1, 2, 4, 4], transformer_depth=[1, 1, 1, 1, 1, 1, 0, 0],
           transformer_depth_output=[1] * 9 + [0] * 3, transformer_depth_middle=1,
           use_linear_in_transformer=False, context_dim=768, use_spatial_transformer=True, legacy=False,
           use_checkpoint=False, adm_in_channels=None, use_temporal_attention=False, use_temporal_resblock=False)
mc = unet.model_config_from_unet_config(cfg)
dt = unet.unet_dtype1()
mc.set_inference_dtype(dt, Device.unet_manual_cast(dt, Device.get_torch_device(), mc.supported_inference_dtypes))
model = mc.get_model({}, "", device=torch.device("cpu"))
sd = O.synth_state_dict(O.unet_param_shapes())
model.diffusion_model.load_state_dict(sd, strict=True)
mp = ModelPatcher.ModelPatcher(model, load_device=torch.device("cpu"), offload_device=torch.device("cpu"))
g = torch.Generator().manual_seed(1234)
ctx_pos = torch.randn(1, 77, 768, generator=g)
ctx_neg = torch.randn(1, 77, 768, generator=g)
out = {"ctx_pos": ctx_pos, "ctx_neg": ctx_neg}
for name, sched, steps, hw, cfg_scale in (("a", "karras", 6, 16, 7.0), ("b", "normal", 5, 32, 4.0)):
    res = sampling.KSampler().sample(model=mp.clone(), seed=42, steps=steps, cfg=cfg_scale, sampler_name="euler_cfgpp",
                                     scheduler=sched, denoise=1.0, positive=[[ctx_pos, {}]], negative=[[ctx_neg, {}]],
                                     latent_image={"samples": torch.zeros(1, 4, hw, hw)}, pipeline=True)
    out[f"{name}_final"] = res[0]["samples"].clone()
    out[f"{name}_args"] = dict(scheduler=sched, steps=steps, hw=hw, cfg=cfg_scale)
    print(name, float(res[0]["samples"].std()), flush=True)
torch.save(out, os.path.join(HERE, "euler_cfgpp_small.pt"))
print("wrote euler_cfgpp_small.pt")
