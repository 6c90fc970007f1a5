#!/bin/bash
# Round 2, GPU call 49: GroupNorm statistics epilogue without shuffles / shared-memory atomics (scratch in the dead operand ring):
# op-level and UNet-level parity, sanitizer, A/B against the statistics kernels on one box.
set -u
O=gpurun_out/r2_call49; mkdir -p $O
timeout -s KILL 600 python -m pytest tests/test_ops_gpu.py tests/test_unet_gpu.py tests/test_fullsize_gpu.py -m gpu -q -s -p no:cacheprovider 2>&1 | grep -E "rel-L2|passed|failed|Error|assert" | tail -8 | tee -a $O/summary.txt
LDN_GEMM_PAIR=0 timeout -s KILL 300 python -m pytest tests/test_ops_gpu.py -m gpu -q -p no:cacheprovider -k "conv3x3" 2>&1 | grep -E "passed|failed|Error|assert" | tail -4 | sed 's/^/[one-CTA kernel] /' | tee -a $O/summary.txt
for tool in memcheck racecheck; do
  LDN_GEMM_PAIR=0 timeout -s KILL 400 compute-sanitizer --tool $tool python -m pytest tests/test_ops_gpu.py -m gpu -q -x -p no:cacheprovider -k "conv3x3_groupnorm and (100-24 or 3-32-32)" > $O/sanitizer_single_$tool.log 2>&1; echo "single $tool rc=$?" | tee -a $O/summary.txt
  grep -E "ERROR SUMMARY|passed|failed|RACECHECK SUMMARY" $O/sanitizer_single_$tool.log | tail -2 | tee -a $O/summary.txt
  timeout -s KILL 400 compute-sanitizer --tool $tool python -m pytest tests/test_ops_gpu.py -m gpu -q -x -p no:cacheprovider -k "conv3x3_groupnorm and (100-24 or 3-32-32)" > $O/sanitizer_pair_$tool.log 2>&1; echo "pair $tool rc=$?" | tee -a $O/summary.txt
  grep -E "ERROR SUMMARY|passed|failed|RACECHECK SUMMARY" $O/sanitizer_pair_$tool.log | tail -2 | tee -a $O/summary.txt
done
B="python bench.py --steps 30 --warmup 5 --no-secondary --no-cpu-baseline --no-config3 --no-gpu-reference"
run() {  # name, env...
  local name=$1; shift
  env "$@" timeout -s KILL 300 $B > $O/bench_$name.json 2> $O/bench_$name.err
  python - <<PY | tee -a $O/summary.txt
import json
try:
    d=json.load(open("$O/bench_$name.json"))
    ac=d["roofline_hbm"]["after_conv"]
    print("$name", "it/s", round(d["value"],2), "ms", round(d["ms_per_step"],3), "finite", d["config"]["finite"], "launches/step", d["gpu_launches"]/d["steps"], "| conv", round(ac["conv_ms"]*1e3,1), "conv+gn", round(ac["conv_groupnorm_ms"]*1e3,1), "us fused", ac["statistics_in_conv_epilogue"])
except Exception as e:
    print("$name", "failed", e)
PY
}
run fused LDN_GN_FUSE=1
run unfused LDN_GN_FUSE=0
run fused_again LDN_GN_FUSE=1
run unfused_again LDN_GN_FUSE=0
