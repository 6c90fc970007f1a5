#!/bin/bash
# Round 2, GPU call 54: LayerNorm with the row kept packed and two rows per warp (more bytes in flight): parity (bit-identical to the
# one-row kernel: the smoke figure must not move), A/B in the step (LDN_LN_ROWS=1 vs default).
set -u
O=gpurun_out/r2_call54; mkdir -p $O
timeout -s KILL 600 python -m pytest tests/test_ops_gpu.py tests/test_unet_gpu.py tests/test_vae_clip_gpu.py tests/test_flux_gpu.py -m gpu -q -p no:cacheprovider 2>&1 | grep -E "passed|failed|Error|assert" | tail -6 | tee -a $O/summary.txt
timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee -a $O/summary.txt
B="python bench.py --steps 30 --warmup 5 --no-secondary --no-cpu-baseline --no-config3 --no-gpu-reference"
run() {  # name, env...
  local name=$1; shift
  env "$@" timeout -s KILL 300 $B > $O/bench_$name.json 2> $O/bench_$name.err
  python - <<PY | tee -a $O/summary.txt
import json
try:
    d=json.load(open("$O/bench_$name.json"))
    ln=d["roofline_hbm"]["layernorm"]
    print("$name", "it/s", round(d["value"],2), "ms", round(d["ms_per_step"],3), "finite", d["config"]["finite"], "| layernorm [32768, 320]", round(ln["ms_per_launch"]*1e3,2), "us frac", round(ln["frac"],3))
except Exception as e:
    print("$name", "failed", e)
PY
}
run rows2 LDN_LN_ROWS=2
run rows1 LDN_LN_ROWS=1
run rows2_again LDN_LN_ROWS=2
run rows1_again LDN_LN_ROWS=1
