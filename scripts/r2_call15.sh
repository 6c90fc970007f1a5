#!/bin/bash
# Round 2, GPU call 15: attention generation 9 as the default (full suite + bench), persistent GEMM with two CTAs per SM.
set -u
O=gpurun_out/r2_call15; mkdir -p $O
for occ in 0 1; do
  LDN_GEMM_OCC2=$occ timeout -s KILL 200 python scripts/dev_gemm_graph.py 0 1 2 3 4 5 6 7 8 2>&1 | sed "s/^/[occ2=$occ] /" | tee -a $O/summary.txt
done
timeout -s KILL 600 python -m pytest tests -m gpu -q -x -p no:cacheprovider 2>&1 | tail -3 | tee -a $O/summary.txt
LDN_GEMM_OCC2=1 timeout -s KILL 400 python -m pytest tests/test_ops_gpu.py tests/test_unet_gpu.py tests/test_fullsize_gpu.py -q -x -p no:cacheprovider 2>&1 | tail -3 | tee -a $O/summary.txt
for occ in 0 1; do
  LDN_GEMM_OCC2=$occ timeout -s KILL 300 python bench.py --steps 20 --warmup 5 --no-secondary --no-cpu-baseline --no-config3 --no-gpu-reference > $O/bench_occ$occ.json 2> $O/bench_occ$occ.err
  python - <<PY | tee -a $O/summary.txt
import json
d=json.load(open("$O/bench_occ$occ.json"))
print("OCC2=$occ", "it/s", round(d["value"],2), "ms", round(d["ms_per_step"],3), "finite", d["config"]["finite"], "attn ms", round(d["roofline"]["ms_per_launch"],4), "frac", round(d["roofline"]["frac"],3))
PY
done
