#!/bin/bash
# Round 2, GPU call 53 (final tree): whole GPU suite, smoke, sanitizer on the conv + GroupNorm op, the contract bench line with
# default flags, the Flux line, the ncu launch list of one step.
set -u
O=gpurun_out/r2_call53; mkdir -p $O
timeout -s KILL 900 python -m pytest tests -m gpu -q -s -rxXs -p no:cacheprovider --durations=8 > $O/gpu_tests.log 2>&1; echo "gpu tests rc=$?" | tee -a $O/summary.txt
grep -E "rel-L2|passed|failed|error" $O/gpu_tests.log | tail -8 | tee -a $O/summary.txt
timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee -a $O/summary.txt
for tool in memcheck racecheck synccheck; do
  timeout -s KILL 400 compute-sanitizer --tool $tool python -m pytest tests/test_ops_gpu.py -m gpu -q -x -p no:cacheprovider -k "conv3x3_groupnorm and (100-24 or 3-32-32)" > $O/sanitizer_$tool.log 2>&1; echo "$tool rc=$?" | tee -a $O/summary.txt
  grep -E "ERROR SUMMARY|passed|failed|RACECHECK SUMMARY" $O/sanitizer_$tool.log | tail -3 | tee -a $O/summary.txt
done
T0=$(date +%s)
timeout -s KILL 1200 python bench.py > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc=$?" | tee -a $O/summary.txt
T1=$(date +%s); echo "bench wall seconds: $((T1-T0))" | tee -a $O/summary.txt
cut -c1-260 $O/bench_n1.json | tee -a $O/summary.txt
timeout -s KILL 600 python bench.py --workload flux --steps 10 --warmup 3 > $O/bench_flux.json 2> $O/bench_flux.err; echo "flux bench rc=$?" | tee -a $O/summary.txt; cut -c1-200 $O/bench_flux.json | tee -a $O/summary.txt
timeout -s KILL 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $O/launches_step.csv python scripts/profile_step.py > $O/prof_step.log 2>&1; echo "ncu launches rc=$?" | tee -a $O/summary.txt
