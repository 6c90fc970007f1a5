#!/bin/bash
# Round 2, GPU call 29: lean epilogue with the residual prefetched before the accumulator wait (LDN_GEMM_LEAN_PF=1) vs not (0).
set -u
O=gpurun_out/r2_call29; mkdir -p $O
for l in 0 1; do
  LDN_GEMM_LEAN_PF=$l timeout -s KILL 200 python scripts/dev_gemm_graph.py 0 3 4 6 7 9 10 11 2>&1 | sed "s/^/[pf=$l] /" | tee -a $O/summary.txt
done
LDN_GEMM_LEAN_PF=1 LDN_GEMM_OCC2=0 timeout -s KILL 200 python scripts/dev_gemm_graph.py 0 4 2>&1 | sed "s/^/[pf=1 occ2=0] /" | tee -a $O/summary.txt
timeout -s KILL 900 python -m pytest tests/test_ops_gpu.py tests/test_unet_gpu.py tests/test_fullsize_gpu.py -q -x -p no:cacheprovider 2>&1 | tail -3 | tee -a $O/summary.txt
for l in 0 1; do
  LDN_GEMM_LEAN_PF=$l timeout -s KILL 300 python bench.py --steps 20 --warmup 5 --no-secondary --no-cpu-baseline --no-config3 --no-gpu-reference > $O/bench_pf$l.json 2> $O/bench_pf$l.err
  python - <<PY | tee -a $O/summary.txt
import json
d=json.load(open("$O/bench_pf$l.json"))
print("LEAN_PF=$l", "it/s", round(d["value"],2), "ms", round(d["ms_per_step"],3), "finite", d["config"]["finite"])
PY
done
LDN_GEMM_LEAN_PF=1 LDN_GEMM_OCC2=0 timeout -s KILL 300 python bench.py --steps 20 --warmup 5 --no-secondary --no-cpu-baseline --no-config3 --no-gpu-reference > $O/bench_pf1_occ0.json 2> $O/bench_pf1_occ0.err
python - <<PY | tee -a $O/summary.txt
import json
d=json.load(open("$O/bench_pf1_occ0.json"))
print("LEAN_PF=1 OCC2=0", "it/s", round(d["value"],2), "ms", round(d["ms_per_step"],3), "finite", d["config"]["finite"])
PY
