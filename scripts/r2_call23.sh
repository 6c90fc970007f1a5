#!/bin/bash
# Round 2, GPU call 23: folded attention with overflow detection after the exponentials (LDN_ATTN_FOLD=2) vs row maximum first (1).
set -u
O=gpurun_out/r2_call23; mkdir -p $O
LDN_ATTN_FOLD=2 FOLD=1 timeout -s KILL 200 python scripts/dev_attn40.py > $O/attn40_fold2.log 2>&1; echo "fold2 rc=$?" | tee -a $O/summary.txt; tail -12 $O/attn40_fold2.log | tee -a $O/summary.txt
for poly in 0 8 4 2; do
  LDN_ATTN_FOLD=2 FOLD=1 LDN_ATTN_POLY=$poly timeout -s KILL 100 python scripts/dev_attn40.py --quick 2>&1 | tail -1 | tee -a $O/summary.txt
done
LDN_ATTN_FOLD=1 FOLD=1 timeout -s KILL 100 python scripts/dev_attn40.py --quick 2>&1 | tail -1 | tee -a $O/summary.txt
timeout -s KILL 600 python -m pytest tests -m gpu -q -x -p no:cacheprovider 2>&1 | tail -3 | tee -a $O/summary.txt
for f in 1 2; do
  LDN_ATTN_FOLD=$f timeout -s KILL 300 python bench.py --steps 20 --warmup 5 --no-secondary --no-cpu-baseline --no-config3 --no-gpu-reference > $O/bench_fold$f.json 2> $O/bench_fold$f.err
  python - <<PY | tee -a $O/summary.txt
import json
d=json.load(open("$O/bench_fold$f.json"))
print("ATTN_FOLD=$f", "it/s", round(d["value"],2), "ms", round(d["ms_per_step"],3), "finite", d["config"]["finite"])
PY
done
LDN_ATTN_FOLD=2 FOLD=1 timeout -s KILL 200 ncu --set full --clock-control none --import-source on -k regex:attn9 -s 2 -c 1 -o $O/attn9_fold2 python scripts/dev_attn40.py --quick > $O/ncu_attn9.log 2>&1; echo "ncu rc=$?" | tee -a $O/summary.txt
