#!/bin/bash
# Round 2, GPU call 27: the contract bench line with default flags (as the driver runs it), wall-clocked.
set -u
O=gpurun_out/r2_call27; mkdir -p $O
T0=$(date +%s)
timeout -s KILL 1200 python bench.py > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc=$?" | tee -a $O/summary.txt
T1=$(date +%s); echo "bench wall seconds: $((T1-T0))" | tee -a $O/summary.txt
cut -c1-300 $O/bench_n1.json | tee -a $O/summary.txt; tail -3 $O/bench_n1.err
