"""Generates tests/golden/lora_apply.pt: a small kohya-style LoRA (diffusers-named and LDM-named UNet modules) applied by the
UNMODIFIED reference -- LoRas.model_lora_keys_unet + load_lora (src/Model/LoRas.py:15-121), ModelPatcher.add_patches /
patch_model / calculate_weight (src/Model/ModelPatcher.py:186-300, 515-548, 621-650) -- to the SD1.5 UNet with seeded synthetic
fp16 weights; records the patched weights (build container only)."""
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
os.chdir(tempfile.mkdtemp(prefix="ldn_golden_"))
from oracle import sd15_oracle as O  # noqa: E402

torch.set_grad_enabled(False)
from src.NeuralNetwork import unet  # noqa: E402
from src.Device import Device  # noqa: E402
from src.Model import LoRas, ModelPatcher  # noqa: E402

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
g = torch.Generator().manual_seed(9)
r = 4
mods = {"lora_unet_down_blocks_0_attentions_0_transformer_blocks_0_attn2_to_k": ("input_blocks.1.1.transformer_blocks.0.attn2.to_k.weight", None, 2.0),
        "lora_unet_input_blocks_1_1_proj_in": ("input_blocks.1.1.proj_in.weight", None, None),      # LDM-named, 1x1 conv, no alpha
        "lora_unet_conv_out": ("out.2.weight", None, 1.0)}                                           # diffusers-named 3x3 conv
lora = {}
full_sd = model.diffusion_model.state_dict()
for m, (key, _, alpha) in mods.items():
    shape = tuple(full_sd[key].shape)
    lora[m + ".lora_up.weight"] = (torch.randn((shape[0], r) + (1,) * (len(shape) - 2), generator=g) * 0.1).half()
    lora[m + ".lora_down.weight"] = (torch.randn((r,) + tuple(shape[1:]), generator=g) * 0.1).half()
    if alpha is not None:
        lora[m + ".alpha"] = torch.tensor(alpha)
key_map = LoRas.model_lora_keys_unet(model, {})
loaded = LoRas.load_lora(lora, key_map)
assert len(loaded) == 3, list(loaded)
mp2 = mp.clone()
mp2.add_patches(loaded, 0.7)
mp2.patch_model()
sd = model.diffusion_model.state_dict()
out = {"lora": lora, "strength": 0.7, "patched": {key: sd[key].clone() for _, (key, _, _) in mods.items()}}
mp2.unpatch_model()
torch.save(out, os.path.join(HERE, "lora_apply.pt")); print("wrote lora_apply.pt", os.path.getsize(os.path.join(HERE, "lora_apply.pt")))
