#!/bin/bash
set -u
O=gpurun_out/r2_call8; mkdir -p $O
for opt in 3 19 131; do
  for i in 2 5 8; do LDN_GEMM_EPI_OPT=$opt timeout -s KILL 100 python scripts/dev_gemm_shapes.py $i 2>&1 | sed "s/^/[epi_opt=$opt] /" | tee -a $O/summary.txt; done
done
for bn in 128 160; do BN=$bn LDN_GEMM_EPI_OPT=3 timeout -s KILL 100 python scripts/dev_gemm_shapes.py 2 2>&1 | sed "s/^/[BN=$bn] /" | tee -a $O/summary.txt; done
LDN_GEMM_EPI_OPT=3 timeout -s KILL 120 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_persist -s 3 -c 1 -o $O/geglu_l0 python scripts/dev_gemm_shapes.py 2 > $O/ncu_geglu.log 2>&1; echo "ncu rc=$?" | tee -a $O/summary.txt
