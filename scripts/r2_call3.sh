#!/bin/bash
# Round 2, GPU call 3: single-launch GroupNorm, GEMM epilogue experiments, attention7 after clean-up.
set -u
O=gpurun_out/r2_call3; mkdir -p $O
# 1. fused GroupNorm: parity + timing
timeout -s KILL 400 python -m pytest tests/test_ops_gpu.py tests/test_unet_gpu.py tests/test_fullsize_gpu.py tests/test_vae_clip_gpu.py -q -x -p no:cacheprovider 2>&1 | tail -4 | tee -a $O/summary.txt
for f in 0 1; do
  LDN_GN_FUSED=$f timeout -s KILL 300 python bench.py --steps 20 --warmup 5 --no-secondary --no-cpu-baseline > $O/bench_gn$f.json 2> $O/bench_gn$f.err
  python - <<PY | tee -a $O/summary.txt
import json
d=json.load(open("$O/bench_gn$f.json"))
print("GN_FUSED=$f", "it/s", round(d["value"],2), "ms", round(d["ms_per_step"],3), "launches/step", d["gpu_launches"]//20, "gn", d["roofline_hbm"]["ms_per_launch"], d["roofline_hbm"]["frac"])
PY
done
# 2. GEMM epilogue experiments (bit 2: no residual read, bit 3: no store, bit 4: main loop only)
for opt in 3 7 11 15 19; do
  for i in 0 1 3; do LDN_GEMM_EPI_OPT=$opt timeout -s KILL 100 python scripts/dev_gemm_shapes.py $i 2>&1 | sed "s/^/[epi_opt=$opt] /" | tee -a $O/summary.txt; done
done
# 3. attention7
for poly in 0 8 3; do
  LDN_ATTN_D40=7 LDN_ATTN_POLY=$poly timeout -s KILL 100 python scripts/dev_attn40.py --quick 2>&1 | tail -1 | tee -a $O/summary.txt
done
LDN_ATTN_D40=7 LDN_ATTN_POLY=0 timeout -s KILL 300 python bench.py --steps 20 --warmup 5 --no-secondary --no-cpu-baseline > $O/bench_gen7p0.json 2> $O/bench_gen7p0.err; cut -c1-200 $O/bench_gen7p0.json | tee -a $O/summary.txt
