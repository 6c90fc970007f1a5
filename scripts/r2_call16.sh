#!/bin/bash
# Round 2, GPU call 16: the contract bench line on the current tree (attention generation 9, selective two-CTA persistent GEMM),
# the ncu launch list of one step, the reference arm.
set -u
O=gpurun_out/r2_call16; mkdir -p $O
timeout -s KILL 900 python bench.py --steps 20 --warmup 5 > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc=$?" | tee -a $O/summary.txt; cut -c1-300 $O/bench_n1.json | tee -a $O/summary.txt; tail -3 $O/bench_n1.err
timeout -s KILL 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $O/launches_step.csv python scripts/profile_step.py > $O/prof_step.log 2>&1; echo "ncu launches rc=$?" | tee -a $O/summary.txt
timeout -s KILL 200 python scripts/dev_gemm_graph.py 2>&1 | tee -a $O/summary.txt
