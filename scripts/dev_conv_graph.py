"""Time the UNet's 3x3 conv shapes (UNet batch 2, 1024^2 image) inside one CUDA graph each (20 launches, operands rotating over
4 buffer sets), through the C ABI.  Usage: python scripts/dev_conv_graph.py [only_index ...]"""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lightdiffusion_next_b200 import _lib as L
lib = L.load(); torch.manual_seed(0); dev = "cuda"
SHAPES = [  # name, H=W, Cin, Cout
    ("L0 320->320 @128", 128, 320, 320),
    ("L0 640->320 @128 (up)", 128, 640, 320),
    ("L0 960->320 @128 (up)", 128, 960, 320),
    ("L1 640->640 @64", 64, 640, 640),
    ("L1 320->640 @64", 64, 320, 640),
    ("L1 1280->640 @64 (up)", 64, 1280, 640),
    ("L1 1920->640 @64 (up)", 64, 1920, 640),
    ("L2 1280->1280 @32", 32, 1280, 1280),
    ("L2 640->1280 @32", 32, 640, 1280),
    ("L2 2560->1280 @32 (up)", 32, 2560, 1280),
    ("L3 1280->1280 @16", 16, 1280, 1280),
    ("L3 2560->1280 @16 (up)", 16, 2560, 1280),
]
only = [int(a) for a in sys.argv[1:]]
REP, SETS, B = 20, 4, 2
for i, (name, HW, Cin, Cout) in enumerate(SHAPES):
    if only and i not in only: continue
    xs = [torch.randn(B, HW, HW, Cin, device=dev).bfloat16() for _ in range(SETS)]
    w = (torch.randn(Cout, 3, 3, Cin, device=dev) / (9 * Cin) ** 0.5).bfloat16()
    bias = torch.randn(Cout, device=dev)
    outs = [torch.empty(B, HW, HW, Cout, device=dev, dtype=torch.bfloat16) for _ in range(SETS)]
    def run(j):
        L.check(lib.ldn_conv3x3_bf16(xs[j % SETS].data_ptr(), w.data_ptr(), B, HW, HW, Cin, Cout, bias.data_ptr(), 0, 0, 0,
                                     outs[j % SETS].data_ptr(), L.cur_stream()))
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
    fl = 2.0 * B * HW * HW * Cout * 9 * Cin
    print(f"{i} {name:26s}: graph {us:7.1f} us {fl/us/1e6:7.1f} TF/s", flush=True)
