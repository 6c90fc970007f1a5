#!/bin/bash
# Round 2, GPU call 39: Pipeline.img2img on the GPU vs the oracle; bislerp; config 5 through bench.
set -u
O=gpurun_out/r2_call39; mkdir -p $O
timeout -s KILL 600 python -m pytest tests/test_parity_r2_gpu.py -q -x -p no:cacheprovider 2>&1 | tail -5 | tee -a $O/summary.txt
timeout -s KILL 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-reference > $O/bench.json 2> $O/bench.err; echo "bench rc=$?" | tee -a $O/summary.txt
python - <<PY | tee -a $O/summary.txt
import json
d=json.load(open("$O/bench.json"))
print(round(d["value"],2), {k:{kk:vv for kk,vv in v.items() if kk!='what'} for k,v in d["batch_configs"].items()})
PY
tail -3 $O/bench.err
