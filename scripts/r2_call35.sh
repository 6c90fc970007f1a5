#!/bin/bash
# Round 2, GPU call 35: GEGLU epilogue with a MUFU-free polynomial Phi (LDN_GEMM_EPI_OPT=7) vs the erf formula (3).
set -u
O=gpurun_out/r2_call35; mkdir -p $O
for opt in 3 7; do
  LDN_GEMM_EPI_OPT=$opt timeout -s KILL 200 python scripts/dev_gemm_graph.py 2 5 8 12 2>&1 | sed "s/^/[epi_opt=$opt] /" | tee -a $O/summary.txt
done
timeout -s KILL 900 python -m pytest tests -m gpu -q -x -s -p no:cacheprovider 2>&1 | grep -E "rel-L2|passed|failed" | tail -4 | tee -a $O/summary.txt
timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee -a $O/summary.txt
for opt in 3 7; do
  LDN_GEMM_EPI_OPT=$opt timeout -s KILL 300 python bench.py --steps 20 --warmup 5 --no-secondary --no-cpu-baseline --no-config3 --no-gpu-reference > $O/bench_$opt.json 2> $O/bench_$opt.err
  python - <<PY | tee -a $O/summary.txt
import json
d=json.load(open("$O/bench_$opt.json"))
print("EPI_OPT=$opt", "it/s", round(d["value"],2), "ms", round(d["ms_per_step"],3), "finite", d["config"]["finite"])
PY
done
