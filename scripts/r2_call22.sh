#!/bin/bash
# Round 2, GPU call 22: attention generation 9 with FOLDED operands (scale and running offset delivered by the tensor core).
set -u
O=gpurun_out/r2_call22; mkdir -p $O
FOLD=1 timeout -s KILL 200 python scripts/dev_attn40.py > $O/attn40_fold.log 2>&1; echo "fold rc=$?" | tee -a $O/summary.txt; tail -12 $O/attn40_fold.log | tee -a $O/summary.txt
for poly in 0 8 4 2; do
  FOLD=1 LDN_ATTN_POLY=$poly timeout -s KILL 100 python scripts/dev_attn40.py --quick 2>&1 | tail -1 | tee -a $O/summary.txt
done
FOLD=0 timeout -s KILL 100 python scripts/dev_attn40.py --quick 2>&1 | tail -1 | tee -a $O/summary.txt
timeout -s KILL 600 python -m pytest tests -m gpu -q -x -p no:cacheprovider 2>&1 | tail -3 | tee -a $O/summary.txt
for f in 0 1; do
  LDN_ATTN_FOLD=$f timeout -s KILL 300 python bench.py --steps 20 --warmup 5 --no-secondary --no-cpu-baseline --no-config3 --no-gpu-reference > $O/bench_fold$f.json 2> $O/bench_fold$f.err
  python - <<PY | tee -a $O/summary.txt
import json
d=json.load(open("$O/bench_fold$f.json"))
print("ATTN_FOLD=$f", "it/s", round(d["value"],2), "ms", round(d["ms_per_step"],3), "finite", d["config"]["finite"])
PY
done
FOLD=1 timeout -s KILL 200 ncu --set full --clock-control none --import-source on -k regex:attn9 -s 2 -c 1 -o $O/attn9_fold python scripts/dev_attn40.py --quick > $O/ncu_attn9.log 2>&1; echo "ncu rc=$?" | tee -a $O/summary.txt
