"""Generates tests/golden/lora_keys.pt: the reference's LoRA-name -> UNet-weight key map for the SD1.5 UNet
(model_lora_keys_unet, src/Model/LoRas.py:86-121, which relies on unet_to_diffusers, src/NeuralNetwork/unet.py:85-185) and
the diffusers map itself (build container only)."""
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
torch.set_grad_enabled(False)
from src.NeuralNetwork import unet  # noqa: E402
from src.Device import Device  # noqa: E402
from src.Model import LoRas  # noqa: E402

cfg = dict(image_size=32, in_channels=4, out_channels=4, model_channels=320, num_res_blocks=[2, 2, 2, 2],
           channel_mult=[1, 2, 4, 4], transformer_depth=[1, 1, 1, 1, 1, 1, 0, 0],
           transformer_depth_output=[1] * 9 + [0] * 3, transformer_depth_middle=1,
           use_linear_in_transformer=False, context_dim=768, use_spatial_transformer=True, legacy=False,
           use_checkpoint=False, adm_in_channels=None, use_temporal_attention=False, use_temporal_resblock=False)
mc = unet.model_config_from_unet_config(cfg); dt = unet.unet_dtype1()
mc.set_inference_dtype(dt, Device.unet_manual_cast(dt, Device.get_torch_device(), mc.supported_inference_dtypes))
model = mc.get_model({}, "", device=torch.device("meta"))
keys = LoRas.model_lora_keys_unet(model, {})
sdk = set(model.state_dict().keys())
out = {"lora_to_unet": dict(keys), "unet_keys": sorted(k for k in sdk if k.startswith("diffusion_model.")),
       "diffusers_map": dict(unet.unet_to_diffusers(mc.unet_config))}
print(len(keys), "lora names;", sum(v in sdk for v in keys.values()), "point at existing weights;", len(out["diffusers_map"]), "diffusers keys")
torch.save(out, os.path.join(HERE, "lora_keys.pt")); print("wrote lora_keys.pt", os.path.getsize(os.path.join(HERE, "lora_keys.pt")))
