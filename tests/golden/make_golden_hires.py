"""Generates tests/golden/hires_small.pt from the UNMODIFIED reference: LatentUpscale (bislerp) and a KSampler second pass
with denoise 0.45 on the upscaled latent (the HiresFix branch of src/user/pipeline.py:346-366).  Build container only."""
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
from src.sample import sampling; from src.Utilities import upscale
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
lat = torch.randn(2, 4, 8, 8, generator=g) * 4.0
lat[1, :, 3, 3] = 0.0   # a zero vector exercises the slerp degenerate branch
up = upscale.LatentUpscale().upscale({"samples": lat}, 128, 128)[0]["samples"]
out["lat"] = lat; out["up"] = up.clone()
up2 = upscale.LatentUpscale().upscale({"samples": lat[:1]}, 192, 128)[0]["samples"]
out["up_rect"] = up2.clone()
res = sampling.KSampler().sample(model=mp, seed=43, steps=4, cfg=8.0, sampler_name="dpmpp_2m_cfgpp", scheduler="normal",
                                 denoise=0.45, positive=[[ctx_pos, {}]], negative=[[ctx_neg, {}]],
                                 latent_image={"samples": up[:1]}, pipeline=True)
out["hires_final"] = res[0]["samples"].clone()
from src.sample import ksampler_util
out["sched_normal_4_d045"] = ksampler_util.calculate_sigmas(model.model_sampling, "normal", int(4 / 0.45))[-5:].clone()
torch.save(out, os.path.join(HERE, "hires_small.pt")); print("wrote hires_small.pt", float(res[0]["samples"].std()))
