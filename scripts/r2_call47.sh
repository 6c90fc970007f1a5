#!/bin/bash
# Round 2, GPU call 47: racecheck attribution -- the hazard racecheck reports in the CTA-pair kernel (tcgen05.alloc.cta_group::2 result
# write) with and without the GroupNorm statistics, and the one-CTA kernel with the statistics.
set -u
O=gpurun_out/r2_call47; mkdir -p $O
K="conv3x3_groupnorm and (100-24 or 3-32-32)"
rc() { local name=$1; shift; env "$@" timeout -s KILL 300 compute-sanitizer --tool racecheck python -m pytest tests/test_ops_gpu.py -m gpu -q -x -p no:cacheprovider -k "$K" > $O/racecheck_$name.log 2>&1; echo "$name rc=$?" | tee -a $O/summary.txt; grep -E "RACECHECK SUMMARY|passed|failed" $O/racecheck_$name.log | tail -2 | tee -a $O/summary.txt; grep -E "Error: Race|and (Read|Write) access" $O/racecheck_$name.log | sort | uniq -c | head -6 | tee -a $O/summary.txt; }
rc pair_gn LDN_GEMM_PAIR=4
rc pair_nogn LDN_GEMM_PAIR=4 LDN_GN_FUSE=0
rc single_gn LDN_GEMM_PAIR=0
rc single_nogn LDN_GEMM_PAIR=0 LDN_GN_FUSE=0
