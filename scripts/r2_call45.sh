#!/bin/bash
# Round 2, GPU call 45: conv shapes graph-timed under the CTA-pair default vs the one-CTA kernel; ncu --set full of the pair conv
# kernel with the GroupNorm statistics; compute-sanitizer on the conv + GroupNorm op; wider split-K A/B under the pair mode.
set -u
O=gpurun_out/r2_call45; mkdir -p $O
timeout -s KILL 200 python scripts/dev_conv_graph.py 2>&1 | tee $O/conv_graph_pair4.txt | tail -14 | sed 's/^/[pair4] /' | tee -a $O/summary.txt
LDN_GEMM_PAIR=0 timeout -s KILL 200 python scripts/dev_conv_graph.py 2>&1 | tee $O/conv_graph_pair0.txt | tail -14 | sed 's/^/[pair0] /' | tee -a $O/summary.txt
timeout -s KILL 240 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_pair -s 2 -c 1 -o $O/conv_pair_gn python scripts/dev_conv_gn_one.py > $O/ncu_conv.log 2>&1; echo "ncu rc=$?" | tee -a $O/summary.txt
for tool in memcheck racecheck synccheck; do
  timeout -s KILL 400 compute-sanitizer --tool $tool python -m pytest tests/test_ops_gpu.py -m gpu -q -x -p no:cacheprovider -k "conv3x3_groupnorm and (100-24 or 3-32-32)" > $O/sanitizer_$tool.log 2>&1; echo "$tool rc=$?" | tee -a $O/summary.txt
  grep -E "ERROR SUMMARY|passed|failed|RACECHECK SUMMARY" $O/sanitizer_$tool.log | tail -3 | tee -a $O/summary.txt
done
B="python bench.py --steps 30 --warmup 5 --no-secondary --no-cpu-baseline --no-config3 --no-gpu-reference"
run() {  # name, env...
  local name=$1; shift
  env "$@" timeout -s KILL 300 $B > $O/bench_$name.json 2> $O/bench_$name.err
  python - <<PY | tee -a $O/summary.txt
import json
try:
    d=json.load(open("$O/bench_$name.json"))
    print("$name", "it/s", round(d["value"],2), "ms", round(d["ms_per_step"],3), "finite", d["config"]["finite"], "launches/step", d["gpu_launches"]/d["steps"])
except Exception as e:
    print("$name", "failed", e)
PY
}
run default LDN_GEMM_PAIR=4
run split148 LDN_GEMM_SPLIT_MAX_TILES=148
run default_again LDN_GEMM_PAIR=4
