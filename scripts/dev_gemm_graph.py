"""Time the UNet's linear-layer GEMM shapes the way the UNet step runs them: 20 back-to-back launches captured in ONE CUDA
graph (no host launch overhead in the number), operands rotating over 4 buffer sets (L2-warm like the step's activations).
Usage: python scripts/dev_gemm_graph.py [only_index ...]"""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lightdiffusion_next_b200 import _lib as L
lib = L.load(); torch.manual_seed(0); dev = "cuda"
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
only = [int(a) for a in sys.argv[1:]]
BN = int(os.environ.get("BN", "0"))
REP, SETS = 20, 4
for i, (name, M, N, K, epi, hb, hr) in enumerate(SHAPES):
    if only and i not in only: continue
    No = N // 2 if epi == 1 else N
    As = [torch.randn(M, K, device=dev).bfloat16() for _ in range(SETS)]
    W = (torch.randn(N, K, device=dev) / K ** 0.5).bfloat16()
    bias = torch.randn(N, device=dev) if hb else None
    ress = [torch.randn(M, No, device=dev).bfloat16() if hr else None for _ in range(SETS)]
    outs = [torch.empty(M, No, device=dev, dtype=torch.bfloat16) for _ in range(SETS)]
    def run(j):
        A, res, out = As[j % SETS], ress[j % SETS], outs[j % SETS]
        L.check(lib.ldn_gemm_bf16(A.data_ptr(), K, K, 0, 0, 0, W.data_ptr(), M, N, bias.data_ptr() if hb else 0, 0, 0, 0,
                                  res.data_ptr() if hr else 0, No, out.data_ptr(), No, 0, epi, 0, 0, BN, L.cur_stream()))
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for j in range(3): run(j)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            for j in range(REP): run(j)
        for _ in range(3): g.replay()
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(s)
        for _ in range(5): g.replay()
        e1.record(s); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / (5 * REP) * 1e3
    fl = 2.0 * M * N * K
    byts = 2.0 * (M * K + N * K + M * No * (2 if hr else 1))
    print(f"{i} {name:22s} M={M} N={N} K={K}: graph {us:7.1f} us {fl/us/1e6:7.1f} TF/s {byts/us/1e3:7.0f} GB/s(alg)", flush=True)
