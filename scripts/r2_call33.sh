#!/bin/bash
# Round 2, GPU call 33: the contract bench line (default flags) and the ncu launch list on the tree with the lean prefetching epilogue and
# the GroupNorm launch bounds.
set -u
O=gpurun_out/r2_call33; mkdir -p $O
T0=$(date +%s)
timeout -s KILL 1200 python bench.py > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc=$?" | tee -a $O/summary.txt
T1=$(date +%s); echo "bench wall seconds: $((T1-T0))" | tee -a $O/summary.txt
cut -c1-300 $O/bench_n1.json | tee -a $O/summary.txt
timeout -s KILL 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $O/launches_step.csv python scripts/profile_step.py > $O/prof_step.log 2>&1; echo "ncu launches rc=$?" | tee -a $O/summary.txt
timeout -s KILL 200 python scripts/dev_gemm_graph.py 2>&1 | tee -a $O/summary.txt
