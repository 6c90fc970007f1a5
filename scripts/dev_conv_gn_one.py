"""One ResBlock pair conv3x3 -> GroupNorm + SiLU of the UNet's level 0 (320 -> 320 at 128x128, UNet batch 2, time-embedding row
bias) through the C ABI (ldn_conv3x3_groupnorm_bf16: statistics in the conv epilogue), for `ncu -k regex:gemm_tc_pair`."""
import ctypes, sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lightdiffusion_next_b200 import _lib as L
lib = L.load(); torch.manual_seed(0)
B, H, W, Cin, Cout = 2, 128, 128, 320, 320
x = torch.randn(B, H, W, Cin, device="cuda").bfloat16(); w = (torch.randn(Cout, 3, 3, Cin, device="cuda") / (9 * Cin) ** 0.5).bfloat16()
b = torch.randn(Cout, device="cuda"); rb = torch.randn(B, Cout, device="cuda")
gamma = torch.ones(Cout, device="cuda"); beta = torch.zeros(Cout, device="cuda")
out = torch.empty(B, H, W, Cout, device="cuda", dtype=torch.bfloat16); gn = torch.empty_like(out)
fused = ctypes.c_int(-1)
for _ in range(4):
    L.check(lib.ldn_conv3x3_groupnorm_bf16(x.data_ptr(), w.data_ptr(), B, H, W, Cin, Cout, b.data_ptr(), rb.data_ptr(), Cout, 0, 1e-5,
                                           gamma.data_ptr(), beta.data_ptr(), 1, out.data_ptr(), gn.data_ptr(), ctypes.addressof(fused),
                                           L.cur_stream()))
torch.cuda.synchronize(); print("ok fused", fused.value)
