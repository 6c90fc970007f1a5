"""One GroupNorm launch sequence at the level-0 shape for ncu (LDN_GN_FUSED selects the kernel)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lightdiffusion_next_b200 import _lib as L
lib = L.load(); dev = "cuda"
B2, HW, C = 2, 16384, int(sys.argv[1]) if len(sys.argv) > 1 else 320
x = torch.randn(B2 * HW, C, device=dev).bfloat16(); y = torch.empty_like(x)
g, b = torch.ones(C, device=dev), torch.zeros(C, device=dev)
def run():
    L.check(lib.ldn_groupnorm_bf16(x.data_ptr(), C, None, 0, B2, HW, 32, 1e-5, g.data_ptr(), b.data_ptr(), 1, y.data_ptr(), L.cur_stream()))
for _ in range(3): run()
torch.cuda.synchronize()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): run()
e1.record(); torch.cuda.synchronize()
print(f"GN fused={os.environ.get('LDN_GN_FUSED','1')} C={C}: {e0.elapsed_time(e1)/20*1e3:.1f} us")
