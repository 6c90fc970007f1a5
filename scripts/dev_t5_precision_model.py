"""CPU model of csrc/t5.cu's storage precision (bf16 weights / activations / residual stream / P, fp32 accumulation and
statistics) in plain torch, against the reference goldens: predicts the rel-L2 the CUDA program should show and which
rounding dominates.  Dev tool (imports the oracle helpers; not part of the product)."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch.nn.functional as F
from oracle import t5_oracle as TO
from test_t5_cpu import t5_tiny_sd

sd, gold = t5_tiny_sd()


def run(ids, flags, cfg=TO.T5_TINY):
    rnd = lambda on: (lambda t: t.to(torch.bfloat16).float()) if on else (lambda t: t)
    bw, ba, br, bp, be = (rnd(c in flags) for c in "warpe")
    H = cfg["num_heads"]
    W = lambda k: bw(sd[k].float())
    x = be(sd["shared.weight"].float())[ids]
    S, n, d = x.shape
    bias = TO.attention_bias(sd, n)

    def rms(x, w, out_bf=True):
        y = sd[w].float() * (x * torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + 1e-6))
        return ba(y) if out_bf else y
    for i in range(cfg["num_layers"]):
        p = f"encoder.block.{i}.layer"
        h = rms(x, f"{p}.0.layer_norm.weight")
        q, k, v = (ba(F.linear(h, W(f"{p}.0.SelfAttention.{c}.weight"))).view(S, n, H, d // H).transpose(1, 2) for c in "qkv")
        s = q @ k.transpose(-1, -2) + bias
        pexp = torch.exp(s - s.max(-1, keepdim=True).values)
        o = ba((bp(pexp) @ v) / pexp.sum(-1, keepdim=True))
        x = br(x + F.linear(o.transpose(1, 2).reshape(S, n, d), W(f"{p}.0.SelfAttention.o.weight")))
        h = rms(x, f"{p}.1.layer_norm.weight")
        g = ba(F.gelu(F.linear(h, W(f"{p}.1.DenseReluDense.wi_0.weight")), approximate="tanh"))
        u = ba(F.linear(h, W(f"{p}.1.DenseReluDense.wi_1.weight")))
        x = br(x + F.linear(ba(g * u), W(f"{p}.1.DenseReluDense.wo.weight")))
    return rms(x, "encoder.final_layer_norm.weight", out_bf=False)


for name in "ab":
    ref = gold[f"out_{name}"]
    print(name, {f or "fp32": round(float((run(gold[f"ids_{name}"], f) - ref).norm() / ref.norm()), 5) for f in ("", "w", "a", "r", "p", "e", "warpe")})
