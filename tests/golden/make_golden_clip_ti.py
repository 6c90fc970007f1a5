"""Generates tests/golden/clip_ti_small.pt: the reference's CLIP-L encode of a token row that carries textual-inversion
embedding VECTORS in place of token ids (what SDTokenizer emits for "embedding:name", SDToken.py:330-360; handled by
SDClipModel.set_up_textual_embeddings, src/SD15/SDClip.py:213-268) -- the reference pipeline's default negative prompt uses
four of them (src/user/pipeline.py:98).  Seeded synthetic weights (build container only)."""
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

torch.set_grad_enabled(False)
from src.SD15 import SDClip  # noqa: E402

shapes = O.clip_param_shapes()
sd = O.synth_state_dict(shapes, seed=777)
m = SDClip.SD1ClipModel(device="cpu", dtype=torch.float16)
pref = "clip_l.transformer.text_model."
full = m.state_dict()
for k, v in sd.items():
    full[pref + k] = v
m.load_state_dict(full)
m.set_clip_options({"layer": -2})
g = torch.Generator().manual_seed(31)
v1, v2, v3 = (torch.randn(768, generator=g) * 0.02 for _ in range(3))
bad = torch.randn(1024, generator=g)  # wrong width: ignored with a warning, the row is re-padded at its end
row = [(49406, 1.0), (320, 1.0), (v1, 1.0), (v2, 1.2), (1125, 1.0), (bad, 1.0), (v3, 0.8)] + [(49407, 1.0)] * 70
assert len(row) == 77
cond, pooled = m.encode_token_weights({"l": [row]})
out = {"vectors": [v1, v2, v3], "bad": bad, "cond": cond.float().clone(),
       "row_spec": [(t if isinstance(t, int) else ("v", i), w) for (t, w), i in zip(row, [None, None, 0, 1, None, "bad", 2] + [None] * 70)]}
print("clip_ti", tuple(cond.shape), float(cond.std()))
torch.save(out, os.path.join(HERE, "clip_ti_small.pt")); print("wrote clip_ti_small.pt")
