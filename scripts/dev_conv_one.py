"""One 3x3 conv of the UNet (level 1: 640 -> 640 at 64x64, UNet batch 2) through the C ABI, for `ncu -k regex:gemm_tc`."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lightdiffusion_next_b200 import _lib as L
lib = L.load(); torch.manual_seed(0)
B, H, W, Cin, Cout = 2, 64, 64, 640, 640
x = torch.randn(B, H, W, Cin, device="cuda").bfloat16(); w = (torch.randn(Cout, 3, 3, Cin, device="cuda") / (9 * Cin) ** 0.5).bfloat16()
b = torch.randn(Cout, device="cuda"); out = torch.empty(B, H, W, Cout, device="cuda", dtype=torch.bfloat16)
for _ in range(4):
    L.check(lib.ldn_conv3x3_bf16(x.data_ptr(), w.data_ptr(), B, H, W, Cin, Cout, b.data_ptr(), 0, 0, 0, out.data_ptr(), L.cur_stream()))
torch.cuda.synchronize(); print("ok")
