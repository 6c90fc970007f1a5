#!/bin/bash
# Round 2, GPU call 2: generation-7 head-dim-40 attention (two softmax threads per row) vs generation 5.
set -u
O=gpurun_out/r2_call2; mkdir -p $O
for gen in 7; do
  LDN_ATTN_D40=$gen timeout -s KILL 200 python scripts/dev_attn40.py > $O/attn40_gen$gen.log 2>&1; echo "gen $gen rc=$?" | tee -a $O/summary.txt; cat $O/attn40_gen$gen.log | tee -a $O/summary.txt
done
for poly in 0 8 4 2; do
  LDN_ATTN_D40=7 LDN_ATTN_POLY=$poly timeout -s KILL 100 python scripts/dev_attn40.py --quick 2>&1 | tail -1 | tee -a $O/summary.txt
done
LDN_ATTN_D40=7 timeout -s KILL 300 python -m pytest tests/test_ops_gpu.py -q -k "attention" -p no:cacheprovider 2>&1 | tail -3 | tee -a $O/summary.txt
LDN_ATTN_D40=7 timeout -s KILL 300 python bench.py --steps 20 --warmup 5 --no-secondary --no-cpu-baseline > $O/bench_gen7.json 2> $O/bench_gen7.err; cut -c1-330 $O/bench_gen7.json | tee -a $O/summary.txt
LDN_ATTN_D40=7 timeout -s KILL 200 ncu --set full --clock-control none --import-source on -k regex:attn7 -s 2 -c 1 -o $O/attn7_full python scripts/dev_attn40.py --quick > $O/ncu_attn7.log 2>&1; echo "ncu rc=$?" | tee -a $O/summary.txt
