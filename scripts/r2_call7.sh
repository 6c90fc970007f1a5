#!/bin/bash
set -u
O=gpurun_out/r2_call7; mkdir -p $O
for f in 0 1 2; do for c in 320 640 1280; do LDN_GN_FUSED=$f python scripts/dev_gn_one.py $c 2>&1 | tail -1 | tee -a $O/summary.txt; done; done
timeout -s KILL 400 python -m pytest tests/test_ops_gpu.py tests/test_unet_gpu.py tests/test_fullsize_gpu.py tests/test_vae_clip_gpu.py -q -x -p no:cacheprovider 2>&1 | tail -4 | tee -a $O/summary.txt
for f in 0 2; do
  LDN_GN_FUSED=$f timeout -s KILL 300 python bench.py --steps 20 --warmup 5 --no-secondary --no-cpu-baseline > $O/bench_gn$f.json 2> $O/bench_gn$f.err
  python - <<PY | tee -a $O/summary.txt
import json
d=json.load(open("$O/bench_gn$f.json"))
print("GN_FUSED=$f", "it/s", round(d["value"],2), "ms", round(d["ms_per_step"],3), "gn", d["roofline_hbm"]["ms_per_launch"], d["roofline_hbm"]["frac"])
PY
done
LDN_GN_FUSED=2 timeout -s KILL 120 ncu --set full --clock-control none --import-source on -k regex:gn_bulk -s 3 -c 1 -o $O/gn_bulk python scripts/dev_gn_one.py 320 > $O/ncu_gn.log 2>&1; echo "ncu gn rc=$?" | tee -a $O/summary.txt
