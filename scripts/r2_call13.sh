#!/bin/bash
# Round 2, GPU call 13: generation-9 head-dim-40 attention (64-key steps, two CTAs per SM) vs generation 5.
set -u
O=gpurun_out/r2_call13; mkdir -p $O
LDN_ATTN_D40=9 timeout -s KILL 200 python scripts/dev_attn40.py > $O/attn40_gen9.log 2>&1; echo "gen 9 rc=$?" | tee -a $O/summary.txt; tail -12 $O/attn40_gen9.log | tee -a $O/summary.txt
for poly in 0 8 4 2; do
  LDN_ATTN_D40=9 LDN_ATTN_POLY=$poly timeout -s KILL 100 python scripts/dev_attn40.py --quick 2>&1 | tail -1 | tee -a $O/summary.txt
done
for st in 3 4; do
  LDN_ATTN_D40=9 LDN_ATTN_STAGES=$st timeout -s KILL 100 python scripts/dev_attn40.py --quick 2>&1 | tail -1 | sed "s/^/[stages=$st] /" | tee -a $O/summary.txt
done
LDN_ATTN_D40=5 timeout -s KILL 100 python scripts/dev_attn40.py --quick 2>&1 | tail -1 | tee -a $O/summary.txt
LDN_ATTN_D40=9 timeout -s KILL 300 python -m pytest tests/test_ops_gpu.py -q -k "attention" -p no:cacheprovider 2>&1 | tail -3 | tee -a $O/summary.txt
LDN_ATTN_D40=9 timeout -s KILL 200 ncu --set full --clock-control none --import-source on -k regex:attn9 -s 2 -c 1 -o $O/attn9_full python scripts/dev_attn40.py --quick > $O/ncu_attn9.log 2>&1; echo "ncu rc=$?" | tee -a $O/summary.txt
