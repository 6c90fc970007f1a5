#!/bin/bash
# Round 2, GPU call 20: split-K for grids up to 148 tiles (level-2 convs: 128 tiles, one CTA per SM today).
set -u
O=gpurun_out/r2_call20; mkdir -p $O
for mt in 74 148; do
  LDN_GEMM_SPLIT_MAX_TILES=$mt timeout -s KILL 300 python bench.py --steps 20 --warmup 5 --no-secondary --no-cpu-baseline --no-config3 --no-gpu-reference > $O/bench_mt$mt.json 2> $O/bench_mt$mt.err
  python - <<PY | tee -a $O/summary.txt
import json
d=json.load(open("$O/bench_mt$mt.json"))
print("SPLIT_MAX_TILES=$mt", "it/s", round(d["value"],2), "ms", round(d["ms_per_step"],3), "finite", d["config"]["finite"], "launches/step", d["gpu_launches"]//20)
PY
done
LDN_GEMM_SPLIT_MAX_TILES=148 timeout -s KILL 400 python -m pytest tests/test_unet_gpu.py tests/test_fullsize_gpu.py -q -x -p no:cacheprovider 2>&1 | tail -3 | tee -a $O/summary.txt
LDN_GEMM_SPLIT_MAX_TILES=148 LDN_PROFILE=1 timeout -s KILL 200 python scripts/profile_ops.py 2>&1 | sort -t: -k2 -n -r | head -5 > /dev/null
