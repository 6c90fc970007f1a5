#!/bin/bash
# Round 2, GPU call 17: generation-10 head-dim-40 attention (32-key steps, two S buffers per tile, two CTAs per SM).
set -u
O=gpurun_out/r2_call17; mkdir -p $O
LDN_ATTN_D40=10 timeout -s KILL 200 python scripts/dev_attn40.py > $O/attn40_gen10.log 2>&1; echo "gen 10 rc=$?" | tee -a $O/summary.txt; tail -12 $O/attn40_gen10.log | tee -a $O/summary.txt
for poly in 0 8 4 2; do
  LDN_ATTN_D40=10 LDN_ATTN_POLY=$poly timeout -s KILL 100 python scripts/dev_attn40.py --quick 2>&1 | tail -1 | tee -a $O/summary.txt
done
LDN_ATTN_D40=9 timeout -s KILL 100 python scripts/dev_attn40.py --quick 2>&1 | tail -1 | tee -a $O/summary.txt
LDN_ATTN_D40=10 timeout -s KILL 300 python -m pytest tests/test_ops_gpu.py -q -k "attention" -p no:cacheprovider 2>&1 | tail -3 | tee -a $O/summary.txt
LDN_ATTN_D40=10 timeout -s KILL 200 ncu --set full --clock-control none --import-source on -k regex:attn10 -s 2 -c 1 -o $O/attn10_full python scripts/dev_attn40.py --quick > $O/ncu_attn10.log 2>&1; echo "ncu rc=$?" | tee -a $O/summary.txt
