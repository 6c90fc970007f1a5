#!/bin/bash
# First GPU call of round 2 (one B200, ~6 min): everything round 1 left unmeasured, outputs under gpurun_out/r2_first/.
#   gpurun --timeout 600 -- 'bash scripts/round2_first_gpu_call.sh'
set -u
O=gpurun_out/r2_first; mkdir -p $O
# 1. T5 program: first execution (eager, sanitizer), then the parity test, then XXL-size timing
timeout -s KILL 120 compute-sanitizer --tool memcheck python scripts/sanitizer_t5.py > $O/t5_memcheck.log 2>&1; echo "t5 memcheck rc=$?" | tee -a $O/summary.txt
timeout -s KILL 120 python -m pytest tests/test_zz_t5_gpu.py -q -rxX -p no:cacheprovider > $O/t5_test.log 2>&1; echo "t5 test rc=$?" | tee -a $O/summary.txt; tail -3 $O/t5_test.log | tee -a $O/summary.txt
timeout -s KILL 150 python scripts/bench_t5.py --tokens 256 > $O/t5_bench.log 2>&1; tail -1 $O/t5_bench.log | tee -a $O/summary.txt
# 2. the whole GPU suite and the contract bench line on this box
timeout -s KILL 200 python -m pytest tests -m gpu -x -q -rxX -p no:cacheprovider > $O/gpu_tests.log 2>&1; tail -2 $O/gpu_tests.log | tee -a $O/summary.txt
timeout -s KILL 150 python bench.py > $O/bench.json 2> $O/bench.err; cut -c1-300 $O/bench.json | tee -a $O/summary.txt
