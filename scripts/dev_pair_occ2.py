"""Dev probe (GPU): 3x3 convs of the UNet under the GEMM mode in LDN_GEMM_PAIR (read once per process), correctness vs
torch and event-timed.  Run under `timeout -s KILL`."""
import sys, os, json, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lightdiffusion_next_b200 import _lib as L
lib = L.load(); torch.manual_seed(0); dev = "cuda"
res = {"mode": os.environ.get("LDN_GEMM_PAIR", "0"), "convs": []}
for (B, H, W, Cin, Cout) in [(2, 32, 32, 1280, 1280), (2, 64, 64, 640, 640), (2, 128, 128, 320, 320), (2, 128, 128, 960, 320), (2, 64, 64, 1920, 640)]:
    x = torch.randn(B, H, W, Cin, device=dev).bfloat16()
    w = (torch.randn(Cout, 3, 3, Cin, device=dev) / (9 * Cin) ** 0.5).bfloat16()
    b = torch.randn(Cout, device=dev); out = torch.zeros(B, H, W, Cout, device=dev, dtype=torch.bfloat16)
    call = lambda: L.check(lib.ldn_conv3x3_bf16(x.data_ptr(), w.data_ptr(), B, H, W, Cin, Cout, b.data_ptr(), None, 0, None, out.data_ptr(), L.cur_stream()))
    call(); torch.cuda.synchronize()
    ref = torch.nn.functional.conv2d(x.float().permute(0, 3, 1, 2), w.float().permute(0, 3, 1, 2), b, padding=1).permute(0, 2, 3, 1)
    rel = ((out.float() - ref).norm() / ref.norm()).item()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(3): call()
    e0.record()
    for _ in range(20): call()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1000 / 20
    res["convs"].append({"shape": [B, H, W, Cin, Cout], "rel": rel, "us": us, "tflops": 2 * B * H * W * 9 * Cin * Cout / us / 1e6})
    print(res["convs"][-1], flush=True)
os.makedirs("gpurun_out", exist_ok=True)
open(f"gpurun_out/pair_occ2_mode{res['mode']}.json", "w").write(json.dumps(res))
