#!/bin/bash
# Round 2, GPU call 25: LayerNorm folded into the consuming projections (LDN_LN_FOLD=1) vs LayerNorm kernels (0).
set -u
O=gpurun_out/r2_call25; mkdir -p $O
for f in 1 0; do
  LDN_LN_FOLD=$f timeout -s KILL 600 python -m pytest tests/test_unet_gpu.py tests/test_fullsize_gpu.py tests/test_parity_r2_gpu.py -q -x -s -p no:cacheprovider 2>&1 | grep -E "rel-L2|passed|failed|Error|error" | tail -8 | sed "s/^/[ln_fold=$f] /" | tee -a $O/summary.txt
done
for f in 0 1; do
  LDN_LN_FOLD=$f timeout -s KILL 300 python bench.py --steps 20 --warmup 5 --no-secondary --no-cpu-baseline --no-config3 --no-gpu-reference > $O/bench_ln$f.json 2> $O/bench_ln$f.err
  python - <<PY | tee -a $O/summary.txt
import json
d=json.load(open("$O/bench_ln$f.json"))
print("LN_FOLD=$f", "it/s", round(d["value"],2), "ms", round(d["ms_per_step"],3), "finite", d["config"]["finite"], "launches/step", d["gpu_launches"]//20)
PY
done
python - <<'PY' 2>&1 | tee -a gpurun_out/r2_call25/summary.txt
# smoke-level accuracy: eps rel-L2 vs the reference golden with and without the fold
import os, subprocess, sys
for f in ("1", "0"):
    env = dict(os.environ, LDN_LN_FOLD=f)
    r = subprocess.run([sys.executable, "-c", "import __graft_entry__ as g; g.smoke()"], env=env, capture_output=True, text=True)
    print("LN_FOLD=" + f, (r.stdout + r.stderr).strip().splitlines()[-1][:300])
PY
