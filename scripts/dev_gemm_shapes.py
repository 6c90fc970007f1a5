"""Time the UNet's linear-layer GEMM shapes in isolation through the C ABI (L2 flushed between launches).
Usage: python scripts/dev_gemm_shapes.py [only_index]"""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lightdiffusion_next_b200 import _lib as L
lib = L.load(); torch.manual_seed(0); dev = "cuda"
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
SHAPES = [  # name, M, N, K, epi, bias, residual
    ("L0 proj/out 320x320", 32768, 320, 320, 0, True, True),
    ("L0 qk", 32768, 640, 320, 0, False, False),
    ("L0 geglu", 32768, 2560, 320, 1, True, False),
    ("L0 ff.out", 32768, 320, 1280, 0, True, True),
    ("L1 out 640x640", 8192, 640, 640, 0, True, True),
    ("L1 geglu", 8192, 5120, 640, 1, True, False),
    ("L1 ff.out", 8192, 640, 2560, 0, True, True),
    ("L2 out 1280x1280", 2048, 1280, 1280, 0, True, True),
    ("L2 geglu", 2048, 10240, 1280, 1, True, False),
    ("L2 ff.out", 2048, 1280, 5120, 0, True, True),
    ("L3 out 1280x1280", 512, 1280, 1280, 0, True, True),
    ("L3 ff.out", 512, 1280, 5120, 0, True, True),
    ("L3 geglu", 512, 10240, 1280, 1, True, False),
]
only = int(sys.argv[1]) if len(sys.argv) > 1 else -1
BN = int(os.environ.get("BN", "0"))
for i, (name, M, N, K, epi, hb, hr) in enumerate(SHAPES):
    if only >= 0 and i != only: continue
    A = torch.randn(M, K, device=dev).bfloat16(); W = (torch.randn(N, K, device=dev) / K ** 0.5).bfloat16()
    bias = torch.randn(N, device=dev) if hb else None
    No = N // 2 if epi == 1 else N
    res = torch.randn(M, No, device=dev).bfloat16() if hr else None
    out = torch.empty(M, No, device=dev, dtype=torch.bfloat16)
    def run():
        L.check(lib.ldn_gemm_bf16(A.data_ptr(), K, K, 0, 0, 0, W.data_ptr(), M, N, bias.data_ptr() if hb else 0, 0, 0, 0,
                                  res.data_ptr() if hr else 0, No, out.data_ptr(), No, 0, epi, 0, 0, BN, L.cur_stream()))
    run(); torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        flush.zero_()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); run(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    # warm (back-to-back, inputs L2-resident where they fit)
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): run()
    e1.record(); torch.cuda.synchronize(); warm = e0.elapsed_time(e1) / 20
    cold = sorted(ts)[len(ts) // 2]
    fl = 2.0 * M * N * K
    print(f"{i} {name:22s} M={M} N={N} K={K}: cold {cold*1e3:7.1f} us {fl/cold/1e9:7.1f} TF/s | warm {warm*1e3:7.1f} us {fl/warm/1e9:7.1f} TF/s", flush=True)
