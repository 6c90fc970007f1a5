#!/bin/bash
# Round 2, GPU call 32: attention generation 9 with a pair / single CTA mix (LDN_ATTN9_MIX) that fills the 2 x 148 slots evenly.
set -u
O=gpurun_out/r2_call32; mkdir -p $O
for m in 0 1; do
  LDN_ATTN9_MIX=$m FOLD=1 timeout -s KILL 200 python scripts/dev_attn40.py 2>&1 | tail -7 | sed "s/^/[mix=$m] /" | tee -a $O/summary.txt
done
for sc in 0.5 0.65 0.8; do
  LDN_ATTN9_SINGLE_COST=$sc FOLD=1 timeout -s KILL 100 python scripts/dev_attn40.py --quick 2>&1 | tail -1 | sed "s/^/[single_cost=$sc] /" | tee -a $O/summary.txt
done
timeout -s KILL 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider 2>&1 | tail -3 | tee -a $O/summary.txt
for m in 0 1; do
  LDN_ATTN9_MIX=$m timeout -s KILL 300 python bench.py --steps 20 --warmup 5 --no-secondary --no-cpu-baseline --no-config3 --no-gpu-reference > $O/bench_mix$m.json 2> $O/bench_mix$m.err
  python - <<PY | tee -a $O/summary.txt
import json
d=json.load(open("$O/bench_mix$m.json"))
print("MIX=$m", "it/s", round(d["value"],2), "ms", round(d["ms_per_step"],3), "finite", d["config"]["finite"], "attn", round(d["roofline"]["ms_per_launch"],4), round(d["roofline"]["frac"],3))
PY
done
