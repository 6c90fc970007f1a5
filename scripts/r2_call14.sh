#!/bin/bash
# Round 2, GPU call 14: mbarrier try_wait with a suspend-time hint (no spinning): attention generations 5 / 9, GEMM shapes, the step.
set -u
O=gpurun_out/r2_call14; mkdir -p $O
for gen in 5 9; do
  LDN_ATTN_D40=$gen timeout -s KILL 100 python scripts/dev_attn40.py --quick 2>&1 | tail -1 | tee -a $O/summary.txt
done
for poly in 0 8 4 2; do
  LDN_ATTN_D40=9 LDN_ATTN_POLY=$poly timeout -s KILL 100 python scripts/dev_attn40.py --quick 2>&1 | tail -1 | tee -a $O/summary.txt
done
LDN_GEMM_EPI_OPT=3 timeout -s KILL 200 python scripts/dev_gemm_graph.py 0 1 2 3 4 7 2>&1 | tee -a $O/summary.txt
timeout -s KILL 600 python -m pytest tests -m gpu -q -x -p no:cacheprovider 2>&1 | tail -3 | tee -a $O/summary.txt
for gen in 5 9; do
  LDN_GEMM_EPI_OPT=3 LDN_ATTN_D40=$gen timeout -s KILL 300 python bench.py --steps 20 --warmup 5 --no-secondary --no-cpu-baseline --no-config3 --no-gpu-reference > $O/bench_gen$gen.json 2> $O/bench_gen$gen.err
  python - <<PY | tee -a $O/summary.txt
import json
d=json.load(open("$O/bench_gen$gen.json"))
print("gen=$gen", "it/s", round(d["value"],2), "ms", round(d["ms_per_step"],3), "finite", d["config"]["finite"], "attn ms", d["roofline"]["ms_per_launch"], d["roofline"]["kernel"][:30])
PY
done
LDN_ATTN_D40=9 timeout -s KILL 200 ncu --set full --clock-control none --import-source on -k regex:attn9 -s 2 -c 1 -o $O/attn9_hint python scripts/dev_attn40.py --quick > $O/ncu_attn9.log 2>&1; echo "ncu rc=$?" | tee -a $O/summary.txt
