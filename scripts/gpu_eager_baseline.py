"""Comparator: the reference's GPU code path restated — PyTorch eager, fp16 weights/activations, cuDNN convs, cuBLAS
linears, torch SDPA (what LightDiffusion-Next runs on a GPU without xformers: SURVEY.md §2.3, facts 1 and 6) — timed on the
same B200 for the same UNet CFG step.  It is the oracle's UNet forward executed on cuda in fp16 (the reference itself cannot
be shipped to the GPU box).  Not part of the product path; run by hand:  python scripts/gpu_eager_baseline.py [size]"""
import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import sd15_oracle as O
import torch.nn.functional as F

size = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
lat = size // 8
dev = "cuda"
torch.backends.cuda.matmul.allow_tf32 = True   # reference sets TF32 on at import (src/Device/Device.py:11-12)
torch.backends.cudnn.allow_tf32 = True
sd = {k: v.to(dev) for k, v in O.synth_state_dict(O.unet_param_shapes()).items()}   # fp16, as the reference stores them
O._w = lambda s, k: s[k]                                                              # no per-call cast on GPU (Device.py:980-1012)
_attn = O.attention
def attention_sdpa(q, k, v, heads, mask=None):
    b, n, c = q.shape
    d = c // heads
    q, k, v = (t.view(b, -1, heads, d).transpose(1, 2) for t in (q, k, v))
    return F.scaled_dot_product_attention(q, k, v, attn_mask=mask).transpose(1, 2).reshape(b, n, c)
O.attention = attention_sdpa
_te = O.timestep_embedding
O.timestep_embedding = lambda t, dim, max_period=10000: _te(t.cpu(), dim, max_period).to(dev, torch.float16)
tables = tuple(t.to(dev) for t in O.make_sigma_tables())
g = torch.Generator().manual_seed(0)
x = (torch.randn(2, 4, lat, lat, generator=g) * 5).to(dev)
sigma = torch.tensor([5.0, 5.0], device=dev)
ctx = torch.randn(2, 77, 768, generator=g).to(dev, torch.float16)

def step():
    s = sigma.view(-1, 1, 1, 1)
    xc = (x / (s ** 2 + 1.0) ** 0.5).half()
    t = O.timestep_index(sigma, tables[1]).float()
    eps = O.unet_forward.__wrapped__(sd, xc, t, ctx) if hasattr(O.unet_forward, "__wrapped__") else unet_half(xc, t)
    den = x - eps.float() * s
    un, co = den.chunk(2)
    return torch.lerp(un, co, 7.0)

def unet_half(xc, t):
    # oracle forward, but without its .float() on the input
    orig = torch.Tensor.float
    try:
        return O.unet_forward(sd, _NoFloat(xc), t, ctx)
    finally:
        pass

class _NoFloat(torch.Tensor):
    @staticmethod
    def __new__(cls, t):
        return torch.Tensor._make_subclass(cls, t)
    def float(self):
        return torch.Tensor._make_subclass(torch.Tensor, self)

with torch.inference_mode():
    for _ in range(3):
        out = step()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    n = 10
    e0.record()
    for _ in range(n):
        out = step()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
print(f"torch-eager fp16 + SDPA restatement of the reference GPU path, {size}x{size} UNet batch 2: {ms:.2f} ms/step = {1000/ms:.2f} it/s, finite={torch.isfinite(out).all().item()}")
