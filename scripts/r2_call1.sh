#!/bin/bash
# Round 2, GPU call 1 (one B200): the whole GPU suite, the contract bench line (both arms), and the launch list of one step.
#   gpurun --timeout 1500 -- 'bash scripts/r2_call1.sh'
set -u
O=gpurun_out/r2_call1; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/smi.txt 2>&1
timeout -s KILL 600 python -m pytest tests -m gpu -q -rxXs -p no:cacheprovider --durations=15 > $O/gpu_tests.log 2>&1; echo "gpu tests rc=$?" | tee -a $O/summary.txt; tail -25 $O/gpu_tests.log | tee -a $O/summary.txt
timeout -s KILL 600 python bench.py --steps 20 --warmup 5 > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc=$?" | tee -a $O/summary.txt; cut -c1-400 $O/bench_n1.json | tee -a $O/summary.txt; tail -5 $O/bench_n1.err
timeout -s KILL 300 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err; echo "ref arm rc=$?" | tee -a $O/summary.txt; cut -c1-600 $O/bench_ref.json | tee -a $O/summary.txt
timeout -s KILL 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $O/launches_step.csv python scripts/profile_step.py > $O/prof_step.log 2>&1; echo "ncu launches rc=$?" | tee -a $O/summary.txt
