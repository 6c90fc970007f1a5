import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lightdiffusion_next_b200.engine import Engine
from lightdiffusion_next_b200.synth import synth_state_dict, unet_shapes
sd = synth_state_dict(unet_shapes())
eng = Engine(max_rows=2, max_h=64, max_w=64, use_graph=False)
eng.load_unet(sd)
g = torch.Generator().manual_seed(9)
hw = 16
x = torch.randn(2, 4, hw, hw, generator=g).cuda(); sigma = torch.tensor([1.5, 6.0]).cuda(); ctx = torch.randn(2, 77, 768, generator=g).cuda()
eng.set_context(ctx)
for r in range(3):
    print("RUN", r, flush=True)
    eng.denoise(x, sigma); torch.cuda.synchronize()
