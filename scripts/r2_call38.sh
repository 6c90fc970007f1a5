#!/bin/bash
# Round 2, GPU call 38: bislerp on the device (ldn_bislerp), HiresFix tests, config 5.
set -u
O=gpurun_out/r2_call38; mkdir -p $O
timeout -s KILL 600 python -m pytest tests/test_parity_r2_gpu.py tests/test_pipeline_gpu.py -q -x -p no:cacheprovider 2>&1 | tail -5 | tee -a $O/summary.txt
