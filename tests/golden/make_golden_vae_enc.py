"""Generates tests/golden/vae_enc_small.pt from the UNMODIFIED reference VAE encoder (imported read-only from
/root/reference; build container only).  Usage: python tests/golden/make_golden_vae_enc.py"""
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

torch.set_grad_enabled(False)
from src.AutoEncoders import VariationalAE as V  # noqa: E402

shapes = dict(O.vae_encoder_param_shapes())
sd = {k: v.float() for k, v in O.synth_state_dict(shapes, seed=2468).items()}
dd = dict(double_z=True, z_channels=4, resolution=256, in_channels=3, out_ch=3, ch=128, ch_mult=[1, 2, 4, 4],
          num_res_blocks=2, attn_resolutions=[], dropout=0.0)
eng = V.AutoencodingEngine(V.Encoder(**dd), V.Decoder(**dd), V.DiagonalGaussianRegularizer())
full = eng.state_dict()
ref_enc = {k: tuple(v.shape) for k, v in full.items() if k.startswith("encoder.") or k.startswith("quant_conv")}
assert ref_enc == shapes, (set(ref_enc) ^ set(shapes))
g = torch.Generator().manual_seed(99)
for k in full:
    full[k] = sd[k] if k in sd else torch.randn(full[k].shape, generator=g) * 0.02
vae = V.VAE(sd=full)
out = {}
for name, (H, W) in {"a": (64, 64), "b": (72, 132)}.items():  # b: not a multiple of 8 -> exercises the crop
    pixels = torch.rand(2 if name == "a" else 1, H, W, 3, generator=g)
    x = vae.vae_encode_crop_pixels(pixels).movedim(-1, 1)
    x = vae.process_input(x).to(vae.vae_dtype).to(vae.device)
    moments = vae.first_stage_model.encode(x, unregularized=True)[0].float()
    torch.manual_seed(777)
    latent = V.VAEEncode().encode(vae, pixels)[0]["samples"].float()
    out[f"pixels_{name}"] = pixels
    out[f"moments_{name}"] = moments.clone()
    out[f"latent_{name}"] = latent.clone()
    print(name, tuple(moments.shape), float(moments.std()), tuple(latent.shape), float(latent.std()), flush=True)
out["sample_seed"] = 777
torch.save(out, os.path.join(HERE, "vae_enc_small.pt"))
print("wrote vae_enc_small.pt")
