"""Generates tests/golden/flux_vae_small.pt: the reference's VAE decoder in its Flux configuration (16 latent channels, no
post_quant_conv; AutoencodingEngine(..., flux=True), src/AutoEncoders/VariationalAE.py:103-145) with seeded synthetic
weights (build container only)."""
import os
import sys
import tempfile
import types

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")
sys.modules.setdefault("torchsde", types.ModuleType("torchsde"))
os.chdir(tempfile.mkdtemp(prefix="ldn_golden_"))
from oracle import sd15_oracle as O  # noqa: E402

torch.set_grad_enabled(False)
from src.AutoEncoders import VariationalAE as V  # noqa: E402

shapes = O.vae_decoder_param_shapes(O.FLUX_VAE_CFG)
dd = dict(double_z=True, z_channels=16, resolution=256, in_channels=3, out_ch=3, ch=128, ch_mult=[1, 2, 4, 4],
          num_res_blocks=2, attn_resolutions=[], dropout=0.0)
eng = V.AutoencodingEngine(V.Encoder(**dd), V.Decoder(**dd), V.DiagonalGaussianRegularizer(), flux=True)
full = eng.state_dict()
ref_dec = {k: tuple(v.shape) for k, v in full.items() if k.startswith("decoder.")}
assert ref_dec == shapes, (set(ref_dec) ^ set(shapes))
sd = {k: v.float() for k, v in O.synth_state_dict(shapes, seed=9753).items()}
eng.load_state_dict(sd, strict=False)
eng = eng.float().eval()
g = torch.Generator().manual_seed(21)
out = {}
for name, shape in {"a": (1, 16, 8, 8), "b": (2, 16, 6, 10)}.items():
    z = torch.randn(shape, generator=g)
    img = torch.clamp((eng.decode(z, flux=True) + 1.0) / 2.0, min=0.0, max=1.0).movedim(1, -1)  # VAE.process_output
    out[f"z_{name}"] = z
    out[f"img_{name}"] = img.float().contiguous().clone()
    print(name, tuple(img.shape), float(img.mean()), float(img.std()), flush=True)
torch.save(out, os.path.join(HERE, "flux_vae_small.pt"))
print("wrote flux_vae_small.pt")
