#!/bin/bash
# Round 2, GPU call 44 (pair-mode convs + GroupNorm statistics in the epilogue as the default): whole GPU suite, smoke, the
# contract bench line with default flags, pair mode 5 A/B, the Flux line, the ncu launch list of one step.
set -u
O=gpurun_out/r2_call44; mkdir -p $O
timeout -s KILL 900 python -m pytest tests -m gpu -q -s -rxXs -p no:cacheprovider --durations=8 > $O/gpu_tests.log 2>&1; echo "gpu tests rc=$?" | tee -a $O/summary.txt
grep -E "rel-L2|passed|failed|error" $O/gpu_tests.log | tail -8 | tee -a $O/summary.txt
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
run default LDN_GEMM_PAIR=4
run pair5 LDN_GEMM_PAIR=5
run default_again LDN_GEMM_PAIR=4
T0=$(date +%s)
timeout -s KILL 1200 python bench.py > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc=$?" | tee -a $O/summary.txt
T1=$(date +%s); echo "bench wall seconds: $((T1-T0))" | tee -a $O/summary.txt
cut -c1-260 $O/bench_n1.json | tee -a $O/summary.txt
timeout -s KILL 600 python bench.py --workload flux --steps 10 --warmup 3 > $O/bench_flux.json 2> $O/bench_flux.err; echo "flux bench rc=$?" | tee -a $O/summary.txt; cut -c1-200 $O/bench_flux.json | tee -a $O/summary.txt
timeout -s KILL 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $O/launches_step.csv python scripts/profile_step.py > $O/prof_step.log 2>&1; echo "ncu launches rc=$?" | tee -a $O/summary.txt
