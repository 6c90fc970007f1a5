#!/bin/bash
# Round 2, GPU call 24: level-0 cross-attention (Nk = 77) on generation 9 instead of generation 5.
set -u
O=gpurun_out/r2_call24; mkdir -p $O
for mn in 512 64; do
  LDN_ATTN9_MIN_NK=$mn timeout -s KILL 300 python -m pytest tests/test_ops_gpu.py -q -k "attention" -p no:cacheprovider 2>&1 | tail -2 | tee -a $O/summary.txt
  LDN_ATTN9_MIN_NK=$mn timeout -s KILL 300 python bench.py --steps 20 --warmup 5 --no-secondary --no-cpu-baseline --no-config3 --no-gpu-reference > $O/bench_mn$mn.json 2> $O/bench_mn$mn.err
  python - <<PY | tee -a $O/summary.txt
import json
d=json.load(open("$O/bench_mn$mn.json"))
print("ATTN9_MIN_NK=$mn", "it/s", round(d["value"],2), "ms", round(d["ms_per_step"],3), "finite", d["config"]["finite"])
PY
done
