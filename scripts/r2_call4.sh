#!/bin/bash
set -u
O=gpurun_out/r2_call4; mkdir -p $O
for f in 0 1; do for c in 320 640 1280; do LDN_GN_FUSED=$f python scripts/dev_gn_one.py $c 2>&1 | tail -1 | tee -a $O/summary.txt; done; done
LDN_GN_FUSED=1 timeout -s KILL 120 ncu --set full --clock-control none --import-source on -k regex:gn_fused -s 3 -c 1 -o $O/gn_fused python scripts/dev_gn_one.py 320 > $O/ncu_gn.log 2>&1; echo "ncu gn rc=$?" | tee -a $O/summary.txt
LDN_GN_FUSED=0 timeout -s KILL 120 ncu --set full --clock-control none -k regex:gn_ -s 6 -c 2 -o $O/gn_two python scripts/dev_gn_one.py 320 > $O/ncu_gn2.log 2>&1; echo "ncu gn2 rc=$?" | tee -a $O/summary.txt
for opt in 19 51 83 115; do
  for i in 0 3; do LDN_GEMM_EPI_OPT=$opt timeout -s KILL 100 python scripts/dev_gemm_shapes.py $i 2>&1 | sed "s/^/[epi_opt=$opt] /" | tee -a $O/summary.txt; done
done
LDN_GEMM_EPI_OPT=19 timeout -s KILL 120 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_persist -s 3 -c 1 -o $O/gemm320_mainloop python scripts/dev_gemm_shapes.py 0 > $O/ncu_gemm.log 2>&1; echo "ncu gemm rc=$?" | tee -a $O/summary.txt
