"""Per-op device timing of one UNet CFG step (LDN_PROFILE=1, eager launches, CUDA events around every op).

Usage: python scripts/profile_ops.py [--size 1024] [--bs 1] > gpurun_out/ops.txt
Prints the per-op table of the last of three runs, sorted by time, with TFLOP/s for GEMM-shaped ops.
"""
import argparse, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ap = argparse.ArgumentParser(); ap.add_argument("--size", type=int, default=1024); ap.add_argument("--bs", type=int, default=1)
ap.add_argument("--child", action="store_true")
args = ap.parse_args()
if args.child:
    sys.path.insert(0, ROOT)
    import torch
    from lightdiffusion_next_b200.engine import Engine
    from lightdiffusion_next_b200.synth import synth_state_dict, unet_shapes
    h = args.size // 8; rows = 2 * args.bs
    eng = Engine(max_rows=rows, max_h=h, max_w=h, use_graph=False)
    eng.load_unet(synth_state_dict(unet_shapes()))
    x = torch.randn(rows, 4, h, h, device="cuda"); sigma = torch.full((rows,), 2.0, device="cuda")
    eng.set_context(torch.randn(rows, 77, 768, device="cuda"))
    for i in range(3):
        print("LDNRUN", i, flush=True)
        eng.denoise(x, sigma); torch.cuda.synchronize()
    sys.exit(0)
env = dict(os.environ, LDN_PROFILE="1")
out = subprocess.run([sys.executable, os.path.abspath(__file__), "--child", "--size", str(args.size), "--bs", str(args.bs)],
                     env=env, capture_output=True, text=True)
if out.returncode != 0:
    print(out.stdout[-2000:], out.stderr[-4000:]); sys.exit(1)
txt = out.stdout.split("LDNRUN 2")[-1]
rows = []
for line in txt.splitlines():
    m = re.match(r"LDNPROF (\d+) ([\d.]+) (.*)", line)
    if m: rows.append((int(m.group(1)), float(m.group(2)), m.group(3)))
tot = sum(r[1] for r in rows)
print("total %.3f ms over %d ops (eager, includes launch gaps)" % (tot, len(rows)))
cat = {}
for i, ms, name in rows:
    key = re.sub(r"^.*\.", "", name.split(" [")[0])
    cat.setdefault(key, [0, 0.0]); cat[key][0] += 1; cat[key][1] += ms
print("\n== by op kind")
for k, (n, ms) in sorted(cat.items(), key=lambda kv: -kv[1][1]):
    print("%-14s %4d  %8.3f ms  %5.1f%%" % (k, n, ms, 100 * ms / tot))
print("\n== every op (program order)")
for i, ms, name in rows:
    m = re.search(r"\[M=(\d+) N=(\d+) K=(\d+)\]", name)
    extra = ""
    if m:
        M, N, K = map(int, m.groups()); extra = "  %7.1f TFLOP/s" % (2.0 * M * N * K / (ms * 1e-3) / 1e12)
    print("%4d %8.4f  %s%s" % (i, ms, name, extra))
