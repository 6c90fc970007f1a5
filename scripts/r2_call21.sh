#!/bin/bash
# Round 2, GPU call 21: attention generation 9 with fp16 probabilities (ex2.approx.f16x2, F16 x BF16 P*V).
set -u
O=gpurun_out/r2_call21; mkdir -p $O
LDN_ATTN_HALFP=1 timeout -s KILL 200 python scripts/dev_attn40.py > $O/attn40_halfp.log 2>&1; echo "halfp rc=$?" | tee -a $O/summary.txt; tail -12 $O/attn40_halfp.log | tee -a $O/summary.txt
for poly in 0 8 4 2; do
  LDN_ATTN_HALFP=1 LDN_ATTN_POLY=$poly timeout -s KILL 100 python scripts/dev_attn40.py --quick 2>&1 | tail -1 | tee -a $O/summary.txt
done
LDN_ATTN_HALFP=0 timeout -s KILL 100 python scripts/dev_attn40.py --quick 2>&1 | tail -1 | tee -a $O/summary.txt
LDN_ATTN_HALFP=1 timeout -s KILL 300 python -m pytest tests/test_ops_gpu.py -q -k "attention" -p no:cacheprovider 2>&1 | tail -3 | tee -a $O/summary.txt
