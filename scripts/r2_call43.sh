#!/bin/bash
# Round 2, GPU call 43: GroupNorm statistics in the conv epilogues -- op-level and UNet-level parity; the CTA-pair conv mode
# (LDN_GEMM_PAIR=4) with the lean prefetching epilogue and the statistics, A/B in the step on one box.
set -u
O=gpurun_out/r2_call43; mkdir -p $O
timeout -s KILL 600 python -m pytest tests/test_ops_gpu.py tests/test_unet_gpu.py tests/test_fullsize_gpu.py -m gpu -q -s -p no:cacheprovider 2>&1 | grep -E "rel-L2|passed|failed|Error|assert" | tail -12 | tee -a $O/summary.txt
LDN_GEMM_PAIR=4 timeout -s KILL 600 python -m pytest tests/test_ops_gpu.py tests/test_unet_gpu.py tests/test_fullsize_gpu.py -m gpu -q -s -p no:cacheprovider 2>&1 | grep -E "rel-L2|passed|failed|Error|assert" | tail -12 | sed 's/^/[pair4] /' | tee -a $O/summary.txt
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
run default LDN_GN_FUSE=1
run pair4_lean_gn LDN_GEMM_PAIR=4
run pair4_lean LDN_GEMM_PAIR=4 LDN_GN_FUSE=0
run pair4_old LDN_GEMM_PAIR=4 LDN_GN_FUSE=0 LDN_GEMM_PAIR_LEAN=0
run pair4_lean_gn_again LDN_GEMM_PAIR=4
run pair2_lean_gn LDN_GEMM_PAIR=2
