"""Generates tests/golden/flux_small.pt from the UNMODIFIED reference Flux3 module (src/BlackForest/Flux.py) on a small
configuration with seeded synthetic weights (build container only)."""
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
from oracle import flux_oracle as FO  # noqa: E402
from oracle import sd15_oracle as O  # noqa: E402

torch.set_grad_enabled(False)
from src.BlackForest import Flux as RF  # noqa: E402
from src.cond import cast  # noqa: E402

cfg = FO.FLUX_TINY
model = RF.Flux3(dtype=torch.float32, device=torch.device("cpu"), operations=cast.disable_weight_init, **cfg)
shapes = FO.flux_param_shapes(cfg)
ref_shapes = {k: tuple(v.shape) for k, v in model.state_dict().items()}
assert ref_shapes == shapes, (sorted(set(ref_shapes) ^ set(shapes))[:10])
sd = {k: v.float() for k, v in O.synth_state_dict(shapes, seed=8642).items()}
for k in sd:  # RMSNorm scales around 1, like trained checkpoints
    if k.endswith(".scale"):
        sd[k] = (1.0 + 0.1 * torch.randn(sd[k].shape, generator=torch.Generator().manual_seed(len(k)))).half().float()
model.load_state_dict(sd, strict=True)
g = torch.Generator().manual_seed(11)
out = {"scales": {k: v for k, v in sd.items() if k.endswith(".scale")}}
for name, (B, h, w, nt) in {"a": (1, 8, 8, 16), "b": (2, 6, 10, 24)}.items():
    x = torch.randn(B, 16, h, w, generator=g)
    t = torch.rand(B, generator=g) * 0.9 + 0.05
    ctx = torch.randn(B, nt, cfg["context_in_dim"], generator=g)
    y = torch.randn(B, cfg["vec_in_dim"], generator=g)
    guid = torch.full((B,), 3.5)
    res = model(x, t, ctx, y, guid)
    out.update({f"x_{name}": x, f"t_{name}": t, f"ctx_{name}": ctx, f"y_{name}": y, f"g_{name}": guid, f"out_{name}": res.float().clone()})
    print(name, tuple(res.shape), float(res.mean()), float(res.std()), flush=True)
torch.save(out, os.path.join(HERE, "flux_small.pt"))
print("wrote flux_small.pt")
