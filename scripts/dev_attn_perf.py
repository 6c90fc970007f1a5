import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lightdiffusion_next_b200 import _lib as L
lib = L.load(); torch.manual_seed(0); dev = "cuda"
def attn(B, H, Nq, Nk, d, causal=False, ones=False):
    hs = {40: 48, 80: 96}[d] if ones else d
    slot = (d + 63) // 64 * 64
    nk_pad = (Nk + 127) // 128 * 128 if Nk % 8 else Nk
    q = torch.randn(B, H, Nq, d, device=dev).bfloat16(); k = torch.randn(B, H, Nk, d, device=dev).bfloat16(); v = torch.randn(B, H, Nk, d, device=dev).bfloat16()
    Qb = torch.zeros(B * Nq, H * slot, device=dev, dtype=torch.bfloat16); Kb = torch.zeros(B * nk_pad, H * slot, device=dev, dtype=torch.bfloat16)
    Qb.view(B, Nq, H, slot)[..., :d] = q.permute(0, 2, 1, 3); Kb.view(B, nk_pad, H, slot)[:, :Nk, :, :d] = k.permute(0, 2, 1, 3)
    Vt = torch.zeros(H * hs, B * nk_pad, device=dev, dtype=torch.bfloat16); Vt.view(H, hs, B, nk_pad)[:, :d, :, :Nk] = v.permute(1, 3, 0, 2)
    if ones: Vt.view(H, hs, B, nk_pad)[:, d] = 1.0
    out = torch.zeros(B * Nq, H * d, device=dev, dtype=torch.bfloat16)
    def run():
        L.check(lib.ldn_attention_bf16(Qb.data_ptr(), H * slot, Kb.data_ptr(), H * slot, Vt.data_ptr(), B * nk_pad, H * hs, hs if ones else 0, B, H, Nq, Nk, nk_pad, d, slot, int(causal), d ** -0.5, out.data_ptr(), H * d, L.cur_stream()))
    run(); torch.cuda.synchronize()
    ref = torch.nn.functional.scaled_dot_product_attention(q.float(), k.float(), v.float(), is_causal=causal).permute(0, 2, 1, 3).reshape(B * Nq, H * d)
    rel = ((out.float() - ref).norm() / ref.norm()).item()
    for _ in range(2): run()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): run()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"attn ones={ones} B={B} H={H} Nq={Nq} Nk={Nk} d={d} causal={causal}: rel={rel:.3e}  {ms:.3f} ms {4*B*H*Nq*Nk*d/ms/1e9:.1f} TFLOP/s", flush=True)
print("env", {k: v for k, v in os.environ.items() if k.startswith("LDN_")})
attn(2, 8, 4096, 4096, 80, ones=True)
attn(2, 8, 4096, 4096, 80)
attn(2, 8, 1000, 1000, 80, ones=True)
attn(2, 8, 16384, 16384, 80, ones=True)
attn(1, 3, 300, 77, 80, ones=True)
attn(2, 8, 16384, 16384, 40, ones=True)
attn(2, 8, 4096, 4096, 40, ones=True)
attn(2, 8, 16384, 77, 40, ones=True)
attn(2, 8, 16384, 16384, 40)
attn(2, 8, 4096, 4096, 40)
attn(2, 8, 1000, 1000, 40)
attn(2, 8, 16384, 77, 40)
attn(3, 12, 77, 77, 64, causal=True)
attn(1, 1, 300, 520, 64, causal=False)
