#!/bin/bash
set -u
O=gpurun_out/r2_call5; mkdir -p $O
LDN_GN_FUSED=0 timeout -s KILL 400 python -m pytest tests/test_ops_gpu.py tests/test_unet_gpu.py tests/test_fullsize_gpu.py tests/test_vae_clip_gpu.py tests/test_flux_gpu.py -q -x -p no:cacheprovider 2>&1 | tail -4 | tee -a $O/summary.txt
for pdl in 0 1; do
  LDN_GN_FUSED=0 LDN_PDL=$pdl timeout -s KILL 300 python bench.py --steps 20 --warmup 5 --no-secondary --no-cpu-baseline > $O/bench_pdl$pdl.json 2> $O/bench_pdl$pdl.err
  python - <<PY | tee -a $O/summary.txt
import json
d=json.load(open("$O/bench_pdl$pdl.json"))
print("PDL=$pdl", "it/s", round(d["value"],2), "ms", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],2), "finite", d["config"]["finite"])
PY
done
tail -3 $O/bench_pdl1.err | tee -a $O/summary.txt
