"""Dev check (GPU): tcgen05 GEMM / conv3x3 vs torch. Run under `timeout`."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lightdiffusion_next_b200 import _lib as L
lib = L.load()
torch.manual_seed(0)
dev = "cuda"

def relerr(a, b):
    return ((a.float() - b.float()).norm() / b.float().norm()).item()

def gemm(M, N, K, bias=False, res=False, BN=0):
    A = torch.randn(M, K, device=dev).bfloat16()
    W = (torch.randn(N, K, device=dev) / K ** 0.5).bfloat16()
    b = torch.randn(N, device=dev) if bias else None
    R = torch.randn(M, N, device=dev).bfloat16() if res else None
    out = torch.zeros(M, N, device=dev, dtype=torch.bfloat16)
    L.check(lib.ldn_gemm_bf16(A.data_ptr(), K, K, None, 0, 0, W.data_ptr(), M, N, L.ptr(b), None, 0, 0, L.ptr(R), N,
                              out.data_ptr(), N, None, 0, 0, 0, BN, L.cur_stream()))
    torch.cuda.synchronize()
    ref = A.float() @ W.float().t()
    if bias: ref += b
    if res: ref += R.float()
    print(f"gemm M={M} N={N} K={K} bias={bias} res={res} BN={BN}: rel={relerr(out, ref):.3e}", flush=True)

def conv(B, H, W_, Cin, Cout, bias=True):
    x = torch.randn(B, Cin, H, W_, device=dev)
    w = torch.randn(Cout, Cin, 3, 3, device=dev) / (9 * Cin) ** 0.5
    b = torch.randn(Cout, device=dev) if bias else None
    xh = x.permute(0, 2, 3, 1).contiguous().bfloat16()
    wt = w.permute(0, 2, 3, 1).contiguous().bfloat16()
    out = torch.zeros(B, H, W_, Cout, device=dev, dtype=torch.bfloat16)
    L.check(lib.ldn_conv3x3_bf16(xh.data_ptr(), wt.data_ptr(), B, H, W_, Cin, Cout, L.ptr(b), None, 0, None,
                                 out.data_ptr(), L.cur_stream()))
    torch.cuda.synchronize()
    ref = torch.nn.functional.conv2d(xh.float().permute(0, 3, 1, 2), wt.float().permute(0, 3, 1, 2), b, padding=1)
    print(f"conv B={B} H={H} W={W_} Cin={Cin} Cout={Cout}: rel={relerr(out.permute(0,3,1,2), ref):.3e}", flush=True)

which = sys.argv[1] if len(sys.argv) > 1 else "all"
if which in ("all", "gemm"):
    gemm(128, 128, 64)
    gemm(128, 160, 256)
    gemm(256, 320, 320, bias=True)
    gemm(1000, 640, 1280, bias=True, res=True)
    gemm(4096, 1280, 2560, bias=True, res=True, BN=256)
    gemm(512, 64, 128, BN=64)
if which in ("all", "conv"):
    conv(2, 16, 16, 64, 64)
    conv(2, 8, 8, 128, 320)
    conv(1, 32, 32, 320, 320)
    conv(2, 128, 128, 320, 320)
    conv(2, 64, 64, 640, 640)
if which in ("all", "perf"):
    import time
    M, N, K = 32768, 320, 2880
    B, H, W_, Cin, Cout = 2, 128, 128, 320, 320
    xh = torch.randn(B, H, W_, Cin, device=dev).bfloat16(); wt = torch.randn(Cout, 3, 3, Cin, device=dev).bfloat16()
    out = torch.zeros(B, H, W_, Cout, device=dev, dtype=torch.bfloat16)
    for _ in range(3):
        lib.ldn_conv3x3_bf16(xh.data_ptr(), wt.data_ptr(), B, H, W_, Cin, Cout, None, None, 0, None, out.data_ptr(), L.cur_stream())
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        lib.ldn_conv3x3_bf16(xh.data_ptr(), wt.data_ptr(), B, H, W_, Cin, Cout, None, None, 0, None, out.data_ptr(), L.cur_stream())
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    print(f"conv3x3 L0 320->320: {ms:.3f} ms, {2*M*N*K/ms/1e9:.1f} TFLOP/s", flush=True)
    A = torch.randn(32768, 1280, device=dev).bfloat16(); Wm = torch.randn(320, 1280, device=dev).bfloat16()
    o = torch.zeros(32768, 320, device=dev, dtype=torch.bfloat16)
    for _ in range(3):
        lib.ldn_gemm_bf16(A.data_ptr(), 1280, 1280, None, 0, 0, Wm.data_ptr(), 32768, 320, None, None, 0, 0, None, 0, o.data_ptr(), 320, None, 0, 0, 0, 0, L.cur_stream())
    e0.record()
    for _ in range(20):
        lib.ldn_gemm_bf16(A.data_ptr(), 1280, 1280, None, 0, 0, Wm.data_ptr(), 32768, 320, None, None, 0, 0, None, 0, o.data_ptr(), 320, None, 0, 0, 0, 0, L.cur_stream())
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    print(f"gemm 32768x320x1280: {ms:.3f} ms, {2*32768*320*1280/ms/1e9:.1f} TFLOP/s", flush=True)
