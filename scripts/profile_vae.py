"""Per-op device timing of one VAE decode (1024^2) with LDN_PROFILE=1 (eager launches)."""
import os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if "--child" in sys.argv:
    sys.path.insert(0, ROOT)
    import torch
    from lightdiffusion_next_b200.engine import Engine
    from lightdiffusion_next_b200.synth import synth_state_dict, vae_decoder_shapes
    eng = Engine(max_rows=2, max_h=128, max_w=128, use_graph=False)
    eng.load_vae(synth_state_dict(vae_decoder_shapes(), seed=4321))
    z = torch.randn(1, 4, 128, 128).cuda()
    for i in range(3):
        print("LDNRUN", i, flush=True); eng.vae_decode(z); torch.cuda.synchronize()
    sys.exit(0)
out = subprocess.run([sys.executable, os.path.abspath(__file__), "--child"], env=dict(os.environ, LDN_PROFILE="1"), capture_output=True, text=True)
if out.returncode != 0:
    print(out.stdout[-2000:], out.stderr[-3000:]); sys.exit(1)
rows = [(int(m.group(1)), float(m.group(2)), m.group(3)) for m in re.finditer(r"LDNPROF (\d+) ([\d.]+) (.*)", out.stdout.split("LDNRUN 2")[-1])]
tot = sum(r[1] for r in rows); print("total %.3f ms over %d ops" % (tot, len(rows)))
for i, ms, name in sorted(rows, key=lambda r: -r[1])[:40]:
    m = re.search(r"\[M=(\d+) N=(\d+) K=(\d+)\]", name)
    extra = "  %7.1f TFLOP/s" % (2.0 * int(m.group(1)) * int(m.group(2)) * int(m.group(3)) / (ms * 1e-3) / 1e12) if m else ""
    print("%4d %8.4f  %s%s" % (i, ms, name, extra))
