#!/bin/bash
set -u
O=gpurun_out/r2_call6; mkdir -p $O
cd scripts/micro && nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/ex2_bench ex2_bench.cu && /tmp/ex2_bench | tee -a ../../$O/summary.txt
