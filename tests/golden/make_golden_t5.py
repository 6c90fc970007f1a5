"""Generates tests/golden/t5_small.pt from the UNMODIFIED reference T5 encoder (src/clip/FluxClip.py: T5 / T5Stack /
T5Block, wrapped in the reference's own SDClipModel + ClipTokenWeightEncoder exactly as T5XXLModel does, :565-590) on a small
configuration with seeded synthetic weights (build container only)."""
import json
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
work = tempfile.mkdtemp(prefix="ldn_golden_")
os.chdir(work)
from oracle import sd15_oracle as O  # noqa: E402
from oracle import t5_oracle as TO  # noqa: E402

torch.set_grad_enabled(False)
from src.clip import FluxClip as RC  # noqa: E402
from src.cond import cast  # noqa: E402
from src.SD15 import SDClip  # noqa: E402

cfg = TO.T5_TINY
config = dict(json.load(open("/root/reference/src/clip/clip/t5_config_xxl.json")), d_model=cfg["d_model"], d_ff=cfg["d_ff"],
              num_heads=cfg["num_heads"], num_layers=cfg["num_layers"], vocab_size=cfg["vocab_size"], d_kv=cfg["d_model"] // cfg["num_heads"])
cfg_path = os.path.join(work, "t5_tiny.json")
json.dump(config, open(cfg_path, "w"))


class TinyT5(SDClip.SDClipModel):  # T5XXLModel (:565-590) with the small config file
    def __init__(self):
        super().__init__(device="cpu", layer="last", layer_idx=None, textmodel_json_config=cfg_path, dtype=torch.float32,
                         special_tokens={"end": 1, "pad": 0}, model_class=RC.T5,
                         model_options={"custom_operations": cast.disable_weight_init})


model = TinyT5()
shapes = TO.t5_param_shapes(cfg)
ref_shapes = {k: tuple(v.shape) for k, v in model.transformer.state_dict().items()}
assert ref_shapes == shapes, sorted(set(ref_shapes) ^ set(shapes))[:10]
sd = {k: v.float() for k, v in O.synth_state_dict(shapes, seed=2468).items()}
over = {}
for k in sd:
    g = torch.Generator().manual_seed(len(k) * 7 + 1)
    if k.endswith("layer_norm.weight"):  # RMS scales around 1, like trained checkpoints
        over[k] = (1.0 + 0.1 * torch.randn(sd[k].shape, generator=g)).half().float()
    elif k.endswith("relative_attention_bias.weight"):  # logit biases of order 1 so the buckets matter
        over[k] = torch.randn(sd[k].shape, generator=g).half().float()
    elif k.endswith("SelfAttention.q.weight"):
        # T5 applies no 1/sqrt(d) logit scale; trained checkpoints carry it in q. Unit-variance q and k over 64 dims would give
        # logits of std 8 -- a softmax so peaked that fp32-vs-bf16 differences of ANY implementation are amplified.
        over[k] = (sd[k] * 0.125).half().float()
    elif k == "shared.weight":
        over[k] = torch.randn(sd[k].shape, generator=g).half().float()
sd.update(over)
model.transformer.load_state_dict(sd, strict=True)
g = torch.Generator().manual_seed(5)
out = {"overrides": over}
# a: plain batch of ids through T5.forward (lengths beyond 128 exercise the log buckets and the clamp at max distance)
for name, (S, n) in {"a": (2, 40), "b": (1, 300)}.items():
    ids = torch.randint(2, cfg["vocab_size"], (S, n), generator=g)
    z, _ = model.transformer(ids, None, intermediate_output=None, final_layer_norm_intermediate=True, dtype=torch.float32)
    out[f"ids_{name}"], out[f"out_{name}"] = ids, z.float().clone()
    print(name, tuple(z.shape), float(z.mean()), float(z.std()), flush=True)
# c / d: the tokenizer's row format (ids + end, padded to 256) through encode_token_weights, without and with weights
prompt = torch.randint(2, cfg["vocab_size"], (9,), generator=g).tolist()
row = TO.pad_tokens(prompt)
z, pooled = model.encode_token_weights([row])
assert pooled is None
out["prompt"], out["out_c"] = prompt, z.float().clone()
rowd = [(t, (1.3 if 2 <= j < 5 else 0.6 if j == 7 else w)) for j, (t, w) in enumerate(row)]
z, _ = model.encode_token_weights([rowd])
out["weights_d"], out["out_d"] = [w for _, w in rowd], z.float().clone()
print("c/d", tuple(z.shape), float((out["out_d"] - out["out_c"]).abs().max()), flush=True)
out["buckets_300"] = RC.T5Attention._relative_position_bucket(torch.arange(300)[None, :] - torch.arange(300)[:, None]).to(torch.uint8)
# the reference's GGUF name map (clip_sd_map, src/Quantize/Quantizer.py:815-858) applied to llama.cpp's T5 encoder names
from src.Quantize import Quantizer as RQ  # noqa: E402
names = ["token_embd.weight", "enc.output_norm.weight"] + [
    f"enc.blk.{i}.{n}.weight" for i in (0, 23) for n in ("attn_q", "attn_k", "attn_v", "attn_o", "attn_norm", "attn_rel_b",
                                                         "ffn_up", "ffn_down", "ffn_gate", "ffn_norm")]
mapped = {}
for k in names:
    m = k
    for a, b in RQ.clip_sd_map.items():
        m = m.replace(a, b)
    mapped[k] = m
out["gguf_name_map"] = mapped
torch.save(out, os.path.join(HERE, "t5_small.pt"))
print("wrote t5_small.pt", os.path.getsize(os.path.join(HERE, "t5_small.pt")))
