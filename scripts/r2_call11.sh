#!/bin/bash
# Round 2, GPU call 11: whole GPU suite on the current tree (fp32-logit VAE attention, query-blocked), VAE timings, Flux config 4 end to end.
set -u
O=gpurun_out/r2_call11; mkdir -p $O
timeout -s KILL 700 python -m pytest tests -m gpu -q -s -rxXs -p no:cacheprovider --durations=10 > $O/gpu_tests.log 2>&1; echo "gpu tests rc=$?" | tee -a $O/summary.txt
grep -E "rel-L2|passed|failed|error" $O/gpu_tests.log | tail -20 | tee -a $O/summary.txt
timeout -s KILL 200 python scripts/dev_time_vae_clip.py 2>&1 | tail -8 | tee -a $O/summary.txt
timeout -s KILL 200 python scripts/dev_hires_check.py 2>&1 | tail -6 | tee -a $O/summary.txt
timeout -s KILL 600 python bench.py --workload flux --steps 10 --warmup 3 > $O/bench_flux.json 2> $O/bench_flux.err; echo "flux bench rc=$?" | tee -a $O/summary.txt; tail -3 $O/bench_flux.err | tee -a $O/summary.txt
python - <<PY | tee -a $O/summary.txt
import json
d=json.load(open("$O/bench_flux.json"))
print("flux it/s", d["value"], "e2e", d.get("e2e"))
print(json.dumps(d.get("secondary"), indent=1))
PY
