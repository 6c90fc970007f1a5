"""The Flux joint attention shape (N = 4352, 24 heads of 128) through the C ABI, for `ncu -k regex:attn6`."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lightdiffusion_next_b200 import _lib as L
lib = L.load(); torch.manual_seed(0)
H, N, d = 24, 4352, 128
QK = torch.randn(N, 2 * H * d, device="cuda").bfloat16(); Vt = torch.randn(H * d, N, device="cuda").bfloat16()
O = torch.empty(N, H * d, device="cuda", dtype=torch.bfloat16)
for _ in range(4):
    L.check(lib.ldn_attention_bf16(QK.data_ptr(), 2 * H * d, QK.data_ptr() + 2 * H * d, 2 * H * d, Vt.data_ptr(), N, H * d, 0, 1, H, N, N, N, d, d, 0, d ** -0.5, O.data_ptr(), H * d, L.cur_stream()))
torch.cuda.synchronize(); print("ok")
