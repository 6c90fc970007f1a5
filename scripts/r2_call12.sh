#!/bin/bash
# Round 2, GPU call 12: launch-floor micro-benchmark; prefetching epilogue of the persistent GEMM (EPI_OPT 3 = old, 7 = new).
set -u
O=gpurun_out/r2_call12; mkdir -p $O
(cd scripts/micro && nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/launch_floor launch_floor.cu && /tmp/launch_floor) 2>&1 | tee -a $O/summary.txt
for opt in 3 7; do
  LDN_GEMM_EPI_OPT=$opt timeout -s KILL 200 python scripts/dev_gemm_graph.py 0 1 3 4 6 7 9 10 2>&1 | sed "s/^/[epi_opt=$opt] /" | tee -a $O/summary.txt
done
timeout -s KILL 400 python -m pytest tests/test_ops_gpu.py tests/test_unet_gpu.py tests/test_fullsize_gpu.py -q -x -p no:cacheprovider 2>&1 | tail -4 | tee -a $O/summary.txt
for opt in 3 7; do
  LDN_GEMM_EPI_OPT=$opt timeout -s KILL 300 python bench.py --steps 20 --warmup 5 --no-secondary --no-cpu-baseline --no-config3 --no-gpu-reference > $O/bench_epi$opt.json 2> $O/bench_epi$opt.err
  python - <<PY | tee -a $O/summary.txt
import json
d=json.load(open("$O/bench_epi$opt.json"))
print("EPI_OPT=$opt", "it/s", round(d["value"],2), "ms", round(d["ms_per_step"],3), "finite", d["config"]["finite"])
PY
done
