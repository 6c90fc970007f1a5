#!/bin/bash
set -u
O=gpurun_out/r2_call9; mkdir -p $O
timeout -s KILL 400 python -m pytest tests/test_ops_gpu.py tests/test_unet_gpu.py tests/test_fullsize_gpu.py tests/test_vae_clip_gpu.py tests/test_flux_gpu.py -q -x -p no:cacheprovider 2>&1 | tail -6 | tee -a $O/summary.txt
for st in 0 1; do
  for i in 0 1 2 3 4 5 7 8; do LDN_GEMM_STAGED=$st timeout -s KILL 100 python scripts/dev_gemm_shapes.py $i 2>&1 | sed "s/^/[staged=$st] /" | tee -a $O/summary.txt; done
done
for st in 0 1; do
  LDN_GEMM_STAGED=$st timeout -s KILL 300 python bench.py --steps 20 --warmup 5 --no-secondary --no-cpu-baseline > $O/bench_st$st.json 2> $O/bench_st$st.err
  python - <<PY | tee -a $O/summary.txt
import json
d=json.load(open("$O/bench_st$st.json"))
print("STAGED=$st", "it/s", round(d["value"],2), "ms", round(d["ms_per_step"],3), "finite", d["config"]["finite"])
PY
done
