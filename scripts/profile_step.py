"""One UNet CFG step (1024^2, B=2) bracketed by cudaProfilerStart/Stop for `ncu --profile-from-start off`."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lightdiffusion_next_b200.engine import Engine
from lightdiffusion_next_b200.synth import synth_state_dict, unet_shapes
size = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
lat = size // 8
eng = Engine(max_rows=2, max_h=lat, max_w=lat, use_graph=False)
eng.load_unet(synth_state_dict(unet_shapes()))
g = torch.Generator().manual_seed(0)
x = (torch.randn(2, 4, lat, lat, generator=g) * 5).cuda(); sigma = torch.tensor([5.0, 5.0]).cuda()
eng.set_context(torch.randn(2, 77, 768, generator=g).cuda())
for _ in range(2):
    out = eng.denoise(x, sigma)
torch.cuda.synchronize()
torch.cuda.profiler.start()
out = eng.denoise(x, sigma)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("finite", torch.isfinite(out).all().item())
