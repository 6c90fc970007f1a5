#!/bin/bash
# Round 2, GPU call 28: lean compile-time epilogue variant for the plain bf16 GEMMs / convs (LDN_GEMM_LEAN=1) vs the general one (0).
set -u
O=gpurun_out/r2_call28; mkdir -p $O
for l in 0 1; do
  LDN_GEMM_LEAN=$l timeout -s KILL 200 python scripts/dev_gemm_graph.py 0 3 4 6 7 9 2>&1 | sed "s/^/[lean=$l] /" | tee -a $O/summary.txt
  LDN_GEMM_LEAN=$l timeout -s KILL 200 python scripts/dev_conv_graph.py 0 3 7 2>&1 | sed "s/^/[lean=$l] /" | tee -a $O/summary.txt
done
timeout -s KILL 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider 2>&1 | tail -3 | tee -a $O/summary.txt
for l in 0 1; do
  LDN_GEMM_LEAN=$l timeout -s KILL 300 python bench.py --steps 20 --warmup 5 --no-secondary --no-cpu-baseline --no-config3 --no-gpu-reference > $O/bench_lean$l.json 2> $O/bench_lean$l.err
  python - <<PY | tee -a $O/summary.txt
import json
d=json.load(open("$O/bench_lean$l.json"))
print("GEMM_LEAN=$l", "it/s", round(d["value"],2), "ms", round(d["ms_per_step"],3), "finite", d["config"]["finite"], "attn", round(d["roofline"]["ms_per_launch"],4), round(d["roofline"]["frac"],3))
PY
done
