import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lightdiffusion_next_b200 import _lib as L
lib = L.load(); torch.manual_seed(0); dev = "cuda"
B, H, Nq, Nk, d = 2, 8, 16384, 16384, 40
slot = 64
Qb = torch.zeros(B * Nq, 2 * H * slot, device=dev, dtype=torch.bfloat16)
Qb.view(B * Nq, 2 * H, slot)[:, :, :d] = torch.randn(B * Nq, 2 * H, d, device=dev).bfloat16()
Vt = torch.zeros(H * 48, B * Nk, device=dev, dtype=torch.bfloat16)
Vt.view(H, 48, B * Nk)[:, :d] = torch.randn(H, d, B * Nk, device=dev).bfloat16()
Vt.view(H, 48, B * Nk)[:, d] = 1.0
out = torch.zeros(B * Nq, H * d, device=dev, dtype=torch.bfloat16)
for _ in range(3):
    L.check(lib.ldn_attention_bf16(Qb.data_ptr(), 2 * H * slot, Qb.data_ptr() + 2 * H * slot, 2 * H * slot, Vt.data_ptr(), B * Nk, H * 48, 48, B, H, Nq, Nk, Nk, d, slot, 0, d ** -0.5, out.data_ptr(), H * d, L.cur_stream()))
torch.cuda.synchronize()
print("ok", torch.isfinite(out.float()).all().item())
