"""Per-op device timing of one Flux.1-dev sized forward (LDN_PROFILE=1, eager launches)."""
import os, re, subprocess, sys, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "bench_flux.py"), "--reps", "1", "--no-graph"],
                     env=dict(os.environ, LDN_PROFILE="1"), capture_output=True, text=True)
if out.returncode != 0:
    print(out.stdout[-2000:], out.stderr[-3000:]); sys.exit(1)
blocks = out.stdout.split("LDNPROF 0 ")
rows = [(int(m.group(1)), float(m.group(2)), m.group(3)) for m in re.finditer(r"LDNPROF (\d+) ([\d.]+) (.*)", "LDNPROF 0 " + blocks[-1])]
tot = sum(r[1] for r in rows); print("total %.3f ms over %d ops" % (tot, len(rows)))
cat = collections.OrderedDict()
for i, ms, name in rows:
    base = name.split(" [")[0]
    key = re.sub(r"^(double_blocks|single_blocks)\.\d+", r"\1", base)
    m = re.search(r"\[M=(\d+) N=(\d+) K=(\d+)\]", name)
    fl = 2.0 * int(m.group(1)) * int(m.group(2)) * int(m.group(3)) if m else 0.0
    c = cat.setdefault(key, [0, 0.0, 0.0]); c[0] += 1; c[1] += ms; c[2] += fl
for k, (n, ms, fl) in sorted(cat.items(), key=lambda kv: -kv[1][1]):
    print("%-34s %4d %9.3f ms %5.1f%%  %s" % (k, n, ms, 100 * ms / tot, ("%7.0f TFLOP/s" % (fl / ms / 1e9)) if fl else ""))
