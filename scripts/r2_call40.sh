#!/bin/bash
# Round 2, GPU call 40: GroupNorm statistics accumulated as 64-bit fixed point (no arrival counter / fence / last-block fold).
set -u
O=gpurun_out/r2_call40; mkdir -p $O
for c in 320 640 1280; do python scripts/dev_gn_one.py $c 2>&1 | tail -1 | tee -a $O/summary.txt; done
timeout -s KILL 900 python -m pytest tests -m gpu -q -x -s -p no:cacheprovider 2>&1 | grep -E "rel-L2|passed|failed|Error" | tail -4 | tee -a $O/summary.txt
timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee -a $O/summary.txt
timeout -s KILL 200 python scripts/dev_determinism.py 2>&1 | tail -3 | tee -a $O/summary.txt
timeout -s KILL 300 python bench.py --steps 20 --warmup 5 --no-secondary --no-cpu-baseline --no-config3 --no-gpu-reference > $O/bench.json 2> $O/bench.err
python - <<PY | tee -a $O/summary.txt
import json
d=json.load(open("$O/bench.json"))
print("step", "it/s", round(d["value"],2), "ms", round(d["ms_per_step"],3), "finite", d["config"]["finite"], "gn", round(d["roofline_hbm"]["ms_per_launch"]*1e3,1), "us", round(d["roofline_hbm"]["frac"],3))
PY
