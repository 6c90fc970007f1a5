"""Generates tests/golden/taesd_small.pt from the UNMODIFIED reference TAESD decoder (Decoder2, src/AutoEncoders/taesd.py)
with seeded synthetic weights (build container only)."""
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
from src.AutoEncoders import taesd  # noqa: E402

shapes = O.taesd_decoder_param_shapes()
dec = taesd.Decoder2(4)
ref_shapes = {k: tuple(v.shape) for k, v in dec.state_dict().items()}
assert ref_shapes == shapes, (set(ref_shapes) ^ set(shapes))
sd = {k: v.float() for k, v in O.synth_state_dict(shapes, seed=1357).items()}
dec.load_state_dict(sd, strict=True)
dec = dec.float()
g = torch.Generator().manual_seed(5)
out = {}
for name, shape in {"a": (2, 4, 8, 8), "b": (1, 4, 6, 10)}.items():
    z = torch.randn(shape, generator=g) * 4.0  # large enough for the Clamp to matter
    y = dec(z)
    out[f"z_{name}"] = z
    out[f"dec_{name}"] = y.movedim(1, -1).contiguous().clone()
    print(name, tuple(y.shape), float(y.mean()), float(y.std()), flush=True)
torch.save(out, os.path.join(HERE, "taesd_small.pt"))
print("wrote taesd_small.pt")
