#!/bin/bash
# Round 2, GPU call 51: bias + time-embedding row bias of the tile staged in shared memory by the idle epilogue warps (one-tile
# lean kernels): parity, sanitizer, A/B in the step (LDN_GEMM_BIAS_SMEM=0 vs default) and on the conv shapes.
set -u
O=gpurun_out/r2_call51; mkdir -p $O
timeout -s KILL 600 python -m pytest tests/test_ops_gpu.py tests/test_unet_gpu.py tests/test_fullsize_gpu.py tests/test_vae_clip_gpu.py -m gpu -q -s -p no:cacheprovider 2>&1 | grep -E "passed|failed|Error|assert" | tail -8 | tee -a $O/summary.txt
timeout -s KILL 400 compute-sanitizer --tool memcheck python -m pytest tests/test_ops_gpu.py -m gpu -q -x -p no:cacheprovider -k "conv3x3" > $O/sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?" | tee -a $O/summary.txt
grep -E "ERROR SUMMARY|passed|failed" $O/sanitizer_memcheck.log | tail -2 | tee -a $O/summary.txt
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
run bias_smem LDN_GEMM_BIAS_SMEM=1
run bias_global LDN_GEMM_BIAS_SMEM=0
run bias_smem_again LDN_GEMM_BIAS_SMEM=1
run bias_global_again LDN_GEMM_BIAS_SMEM=0
