#!/bin/bash
# Round 2, GPU call 30: residual prefetch in the one-tile-per-CTA lean kernel as well (convs with a skip connection, long-K ff.out);
# two-CTA persistent mode off by default.
set -u
O=gpurun_out/r2_call30; mkdir -p $O
timeout -s KILL 200 python scripts/dev_gemm_graph.py 0 3 6 9 11 2>&1 | tee -a $O/summary.txt
timeout -s KILL 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider 2>&1 | tail -3 | tee -a $O/summary.txt
timeout -s KILL 300 python bench.py --steps 20 --warmup 5 --no-secondary --no-cpu-baseline --no-config3 --no-gpu-reference > $O/bench.json 2> $O/bench.err
python - <<PY | tee -a $O/summary.txt
import json
d=json.load(open("$O/bench.json"))
print("final", "it/s", round(d["value"],2), "ms", round(d["ms_per_step"],3), "finite", d["config"]["finite"], "attn frac", round(d["roofline"]["frac"],3))
PY
