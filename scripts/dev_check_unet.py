"""Dev check (GPU): full UNet denoise vs the CPU fp32 oracle on small latents + timing at 1024^2."""
import sys, os, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lightdiffusion_next_b200.engine import Engine
from oracle import sd15_oracle as O

t0 = time.time()
sd = O.synth_state_dict(O.unet_param_shapes())
print(f"synth weights {time.time()-t0:.1f}s", flush=True)
eng = Engine(max_rows=2, max_h=128, max_w=128, use_graph=("nograph" not in sys.argv))
eng.load_unet(sd)
print(f"loaded {time.time()-t0:.1f}s", flush=True)
g = torch.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "unet_small.pt"))
for hw in (16, 32):
    x = g[f"apply_x_{hw}"].cuda(); sigma = g[f"apply_sigma_{hw}"].cuda(); ctx = g[f"apply_ctx_{hw}"].cuda()
    eng.set_context(ctx)
    out = eng.denoise(x, sigma)
    out2 = eng.denoise(x, sigma)   # second call goes through the CUDA graph
    torch.cuda.synchronize()
    ref = g[f"apply_out_{hw}"]
    # compare on eps = (x - denoised)/sigma to remove the trivially-matching x term
    s = sigma.view(-1, 1, 1, 1).cpu()
    eps_ref = (g[f"apply_x_{hw}"] - ref) / s
    eps = (g[f"apply_x_{hw}"] - out.cpu()) / s
    eps2 = (g[f"apply_x_{hw}"] - out2.cpu()) / s
    r = ((eps - eps_ref).norm() / eps_ref.norm()).item()
    r2 = ((eps2 - eps_ref).norm() / eps_ref.norm()).item()
    print(f"unet {hw}x{hw}: eps rel-L2 vs reference golden = {r:.3e} (graph replay {r2:.3e}), denoised rel = "
          f"{((out.cpu()-ref).norm()/ref.norm()).item():.3e}, finite={torch.isfinite(out).all().item()}", flush=True)
if "time" in sys.argv:
    for hw in (64, 128):
        x = torch.randn(2, 4, hw, hw, device="cuda"); sigma = torch.tensor([3.0, 3.0], device="cuda")
        eng.set_context(torch.randn(2, 77, 768, device="cuda"))
        for _ in range(3): out = eng.denoise(x, sigma)
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): out = eng.denoise(x, sigma)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print(f"unet {hw*8}^2 B=2: {ms:.2f} ms/step = {1000/ms:.1f} it/s  finite={torch.isfinite(out).all().item()}", flush=True)
