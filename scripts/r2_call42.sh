#!/bin/bash
# Round 2, GPU call 42: GroupNorm statistics taken in the ResBlock convs' epilogues: parity, A/B in the step
# (LDN_GN_FUSE=0 vs default), and the CTA-pair conv mode re-measured on the same box (LDN_GEMM_PAIR=4).
set -u
O=gpurun_out/r2_call42; mkdir -p $O
timeout -s KILL 600 python -m pytest tests/test_unet_gpu.py tests/test_fullsize_gpu.py tests/test_ops_gpu.py -m gpu -q -x -s -p no:cacheprovider 2>&1 | grep -E "rel-L2|passed|failed|Error|assert" | tail -8 | tee -a $O/summary.txt
timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee -a $O/summary.txt
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
run fused LDN_GN_FUSE=1
run unfused LDN_GN_FUSE=0
run fused_again LDN_GN_FUSE=1
run pair4 LDN_GN_FUSE=1 LDN_GEMM_PAIR=4
run pair4_unfused LDN_GN_FUSE=0 LDN_GEMM_PAIR=4
