#!/bin/bash
# Round 2, GPU call 37 / 41 (final tree): whole GPU suite, smoke, the contract bench line with default flags, the Flux line.
set -u
O=gpurun_out/r2_call41; mkdir -p $O
timeout -s KILL 900 python -m pytest tests -m gpu -q -s -rxXs -p no:cacheprovider --durations=8 > $O/gpu_tests.log 2>&1; echo "gpu tests rc=$?" | tee -a $O/summary.txt
grep -E "rel-L2|passed|failed|error" $O/gpu_tests.log | tail -6 | tee -a $O/summary.txt
timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee -a $O/summary.txt
T0=$(date +%s)
timeout -s KILL 1200 python bench.py > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc=$?" | tee -a $O/summary.txt
T1=$(date +%s); echo "bench wall seconds: $((T1-T0))" | tee -a $O/summary.txt
cut -c1-260 $O/bench_n1.json | tee -a $O/summary.txt
timeout -s KILL 600 python bench.py --workload flux --steps 10 --warmup 3 > $O/bench_flux.json 2> $O/bench_flux.err; echo "flux bench rc=$?" | tee -a $O/summary.txt; cut -c1-200 $O/bench_flux.json | tee -a $O/summary.txt
timeout -s KILL 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $O/launches_step.csv python scripts/profile_step.py > $O/prof_step.log 2>&1; echo "ncu launches rc=$?" | tee -a $O/summary.txt
