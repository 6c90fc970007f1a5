"""Flux.1-dev sized DiT forward (BASELINE config 4: 1024x1024 -> 128x128x16 latent -> 4096 image tokens) with seeded
synthetic weights generated on the GPU (11.9 G parameters, bf16 in HBM). Prints ms / forward and achieved TFLOP/s.
Usage: python scripts/bench_flux.py [--txt 256] [--size 1024] [--profile]"""
import argparse, os, sys, time, zlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
ap = argparse.ArgumentParser(); ap.add_argument("--txt", type=int, default=256); ap.add_argument("--size", type=int, default=1024)
ap.add_argument("--reps", type=int, default=5); ap.add_argument("--no-graph", action="store_true")
args = ap.parse_args()
import torch
from lightdiffusion_next_b200.engine import Engine
from lightdiffusion_next_b200 import flux as FX
cfg = FX.FLUX_DEV
shapes = FX.flux_shapes(cfg)
eng = Engine(max_rows=1, max_h=8, max_w=8, use_graph=not args.no_graph)
t0 = time.time()
batch = {}
nbytes = 0
for k, shp in shapes.items():
    g = torch.Generator(device="cuda").manual_seed(zlib.crc32(k.encode()) & 0x7FFFFFFF)
    if len(shp) > 1:
        w = torch.randn(shp, generator=g, device="cuda", dtype=torch.float32 if shp[0] * shp[1] < 1 << 24 else torch.bfloat16)
        w = (w * (0.5 if any(s in k for s in ("proj.", "mlp.2", "linear2")) else 1.0) / shp[1] ** 0.5).to(torch.bfloat16)
    elif k.endswith(".scale"):
        w = (1.0 + 0.1 * torch.randn(shp, generator=g, device="cuda")).float()
    else:
        w = (0.02 * torch.randn(shp, generator=g, device="cuda")).float()
    batch[k] = w
    nbytes += w.numel() * w.element_size()
    if nbytes > 2 << 30:  # ingest in ~2 GB slices so that torch's copy and the engine's copy never coexist in full
        eng.load_weights(4, batch); batch = {}; nbytes = 0; torch.cuda.empty_cache()
if batch:
    eng.load_weights(4, batch); batch = {}
torch.cuda.empty_cache()
print(f"weights: {sum(int(torch.tensor(s).prod()) for s in shapes.values())/1e9:.2f} G params loaded in {time.time()-t0:.1f} s", flush=True)
lat = args.size // 8
x = torch.randn(1, 16, lat, lat, device="cuda"); ctx = torch.randn(1, args.txt, cfg["context_in_dim"], device="cuda")
y = torch.randn(1, cfg["vec_in_dim"], device="cuda"); t = torch.tensor([0.7], device="cuda"); g = torch.tensor([3.5], device="cuda")
for _ in range(2):
    out = eng.flux_forward(x, t, ctx, y, g)
torch.cuda.synchronize()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(args.reps):
    out = eng.flux_forward(x, t, ctx, y, g)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / args.reps
C, M, H = cfg["hidden_size"], int(cfg["hidden_size"] * cfg["mlp_ratio"]), cfg["num_heads"]
Ni, Nt = (lat // 2) ** 2, args.txt
N = Ni + Nt
dbl = cfg["depth"] * (2 * N * C * (3 * C + C + 2 * M) + 4 * N * N * C)
sgl = cfg["depth_single_blocks"] * (2 * N * C * (3 * C + M) + 2 * N * (C + M) * C + 4 * N * N * C)
flops = dbl + sgl
print(f"flux forward {args.size}^2, {Ni}+{Nt} tokens: {ms:.2f} ms = {flops/ms/1e9:.0f} TFLOP/s ({flops/1e12:.1f} TFLOP/forward)  finite={torch.isfinite(out).all().item()} std={out.std().item():.3f}", flush=True)
