#!/bin/bash
# Round 2, GPU call 19: conv shapes inside a graph; deeper smem ring for grids of <= 148 tiles (LDN_GEMM_DEEP_SINGLE).
set -u
O=gpurun_out/r2_call19; mkdir -p $O
for d in 0 1; do
  LDN_GEMM_DEEP_SINGLE=$d timeout -s KILL 300 python scripts/dev_conv_graph.py 2>&1 | sed "s/^/[deep=$d] /" | tee -a $O/summary.txt
done
timeout -s KILL 400 python -m pytest tests/test_ops_gpu.py tests/test_unet_gpu.py tests/test_fullsize_gpu.py tests/test_vae_clip_gpu.py -q -x -p no:cacheprovider 2>&1 | tail -3 | tee -a $O/summary.txt
for d in 0 1; do
  LDN_GEMM_DEEP_SINGLE=$d timeout -s KILL 300 python bench.py --steps 20 --warmup 5 --no-secondary --no-cpu-baseline --no-config3 --no-gpu-reference > $O/bench_deep$d.json 2> $O/bench_deep$d.err
  python - <<PY | tee -a $O/summary.txt
import json
d=json.load(open("$O/bench_deep$d.json"))
print("DEEP_SINGLE=$d", "it/s", round(d["value"],2), "ms", round(d["ms_per_step"],3), "finite", d["config"]["finite"])
PY
done
