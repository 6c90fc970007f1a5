import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lightdiffusion_next_b200.engine import Engine
from lightdiffusion_next_b200.synth import synth_state_dict, unet_shapes
sd = synth_state_dict(unet_shapes())
eng = Engine(max_rows=2, max_h=64, max_w=64, use_graph=False)
eng.load_unet(sd)
def rel(a,b): return ((a-b).norm()/b.norm()).item()
g = torch.Generator().manual_seed(9)
for hw in (16, 32, 64):
    x = torch.randn(2, 4, hw, hw, generator=g).cuda(); sigma = torch.tensor([1.5, 6.0]).cuda(); ctx = torch.randn(2, 77, 768, generator=g).cuda()
    eng.set_context(ctx)
    a = eng.denoise(x, sigma).clone(); b = eng.denoise(x, sigma).clone(); c = eng.denoise(x, sigma).clone()
    print(f"hw={hw} same-input repeat: {rel(a,b):.3e} {rel(a,c):.3e} bit-equal={torch.equal(a,b)}", flush=True)
    x2 = x.clone(); x2[1] = torch.randn(4, hw, hw, generator=g).cuda() * 5
    d = eng.denoise(x2, sigma)
    print(f"   change x[1] only: row0 diff {rel(d[:1], a[:1]):.3e}  row1 diff {rel(d[1:], a[1:]):.3e}", flush=True)
    ctx2 = ctx.clone(); ctx2[1] = torch.randn(77, 768, generator=g).cuda()
    eng.set_context(ctx2); e = eng.denoise(x, sigma)
    print(f"   change ctx[1] only: row0 diff {rel(e[:1], a[:1]):.3e}  row1 diff {rel(e[1:], a[1:]):.3e}", flush=True)
    eng.set_context(ctx); s2 = sigma.clone(); s2[1] = 0.3; f = eng.denoise(x, s2)
    print(f"   change sigma[1] only: row0 diff {rel(f[:1], a[:1]):.3e}  row1 diff {rel(f[1:], a[1:]):.3e}", flush=True)
