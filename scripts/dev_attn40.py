"""Head-dim-40 self-attention (the level-0 kernel): parity vs fp32 SDPA and CUDA-event timing for the generation selected
by LDN_ATTN_D40 (5 | 7) and the polynomial share LDN_ATTN_POLY.  Run once per setting (the choice is read at first use)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lightdiffusion_next_b200 import _lib as L
lib = L.load(); torch.manual_seed(0); dev = "cuda"
FOLD = int(os.environ.get("FOLD", "0"))  # 1: folded operands (Q pre-scaled, ones column in K; `causal` bit 1 of the C entry)
tag = f"gen={os.environ.get('LDN_ATTN_D40', '9')} poly={os.environ.get('LDN_ATTN_POLY', '3')} fold={FOLD}"
def attn(B, H, Nq, Nk, d=40, reps=10, spread=1.0):
    hs, slot = 48, 64
    nk_pad = (Nk + 127) // 128 * 128 if Nk % 8 else Nk
    q = (torch.randn(B, H, Nq, d, device=dev) * spread).bfloat16(); k = torch.randn(B, H, Nk, d, device=dev).bfloat16(); v = torch.randn(B, H, Nk, d, device=dev).bfloat16()
    Qb = torch.zeros(B * Nq, H * slot, device=dev, dtype=torch.bfloat16); Kb = torch.zeros(B * nk_pad, H * slot, device=dev, dtype=torch.bfloat16)
    qs = (q.float() * (d ** -0.5 * 1.4426950408889634)).bfloat16() if FOLD else q
    Qb.view(B, Nq, H, slot)[..., :d] = qs.permute(0, 2, 1, 3); Kb.view(B, nk_pad, H, slot)[:, :Nk, :, :d] = k.permute(0, 2, 1, 3)
    if FOLD: Kb.view(B, nk_pad, H, slot)[:, :Nk, :, d] = 1.0
    Vt = torch.zeros(H * hs, B * nk_pad, device=dev, dtype=torch.bfloat16); Vt.view(H, hs, B, nk_pad)[:, :d, :, :Nk] = v.permute(1, 3, 0, 2)
    Vt.view(H, hs, B, nk_pad)[:, d] = 1.0
    out = torch.zeros(B * Nq, H * d, device=dev, dtype=torch.bfloat16)
    def run():
        L.check(lib.ldn_attention_bf16(Qb.data_ptr(), H * slot, Kb.data_ptr(), H * slot, Vt.data_ptr(), B * nk_pad, H * hs, hs, B, H, Nq, Nk, nk_pad, d, slot, 2 if FOLD else 0, d ** -0.5, out.data_ptr(), H * d, L.cur_stream()))
    run(); torch.cuda.synchronize()
    ref = torch.nn.functional.scaled_dot_product_attention(q.float(), k.float(), v.float()).permute(0, 2, 1, 3).reshape(B * Nq, H * d)
    rel = ((out.float() - ref).norm() / ref.norm()).item()
    a = out.clone(); run(); torch.cuda.synchronize(); det = torch.equal(a, out)
    for _ in range(2): run()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): run()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f"[{tag}] B={B} H={H} Nq={Nq} Nk={Nk} spread={spread}: rel={rel:.3e} det={det} {ms:.4f} ms {4*B*H*Nq*Nk*d/ms/1e9:.1f} TFLOP/s", flush=True)
attn(2, 8, 16384, 16384, reps=20)
if "--quick" not in sys.argv:
    attn(1, 8, 4096, 4096)
    attn(2, 8, 1024, 1024)
    attn(1, 2, 300, 1100)          # ragged: partial last query tile and partial last key tile (both key halves masked differently)
    attn(1, 2, 200, 1090)          # last key tile holds 66 keys: the second half has 2 valid columns
    attn(1, 8, 4096, 4096, spread=6.0)   # large logits: exercises the lazy rescale path
    attn(4, 8, 16384, 16384, reps=5)     # UNet batch 4 (bs = 2 / GPU)
