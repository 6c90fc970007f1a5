"""Generates tests/golden/mask_small.pt: the reference's KSampler with an inpaint `noise_mask` in the latent dict
(common_ksampler, src/sample/sampling.py:1203-1221 -> KSAMPLER.sample -> KSamplerX0Inpaint, :363-378) on the img2img case of
hires_small.pt.  Finding recorded by this script: the reference's KSamplerX0Inpaint.__call__ forwards to the inner model and
never blends with the mask, so for a non-inpaint SD1.5 UNet the mask changes nothing (build container only)."""
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
from src.sample import sampling  # noqa: E402

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
g = torch.load(os.path.join(HERE, "hires_small.pt"))
up = g["up"][:1]
mask = torch.zeros(1, 1, up.shape[2], up.shape[3]); mask[:, :, 4:12, 2:9] = 1.0
res = sampling.KSampler().sample(model=mp, seed=43, steps=4, cfg=8.0, sampler_name="dpmpp_2m_cfgpp", scheduler="normal",
                                 denoise=0.45, positive=[[g["ctx_pos"], {}]], negative=[[g["ctx_neg"], {}]],
                                 latent_image={"samples": up, "noise_mask": mask}, pipeline=True)
final = res[0]["samples"].clone()
same = bool(torch.equal(final, g["hires_final"]))
print("masked run equals unmasked run:", same, float((final - g["hires_final"]).abs().max()))
torch.save({"mask": mask, "masked_final": final, "equals_unmasked": same}, os.path.join(HERE, "mask_small.pt"))
print("wrote mask_small.pt")
