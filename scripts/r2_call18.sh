#!/bin/bash
# Round 2, GPU call 18 (two B200s): the bench contract under torchrun at N = 2 (sharded config 3 / config 5 on NCCL), reference arm at N = 2.
set -u
O=gpurun_out/r2_call18; mkdir -p $O
nvidia-smi -L | tee -a $O/summary.txt
timeout -s KILL 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > $O/bench_n2.json 2> $O/bench_n2.err; echo "bench n2 rc=$?" | tee -a $O/summary.txt
tail -5 $O/bench_n2.err | tee -a $O/summary.txt
python - <<PY | tee -a $O/summary.txt
import json
for line in open("$O/bench_n2.json"):
    line=line.strip()
    if not line.startswith("{"): continue
    d=json.loads(line)
    print("n_gpus", d["n_gpus"], "it/s", round(d["value"],2), "ms", round(d["ms_per_step"],3), "e2e", d.get("e2e",{}).get("value"))
    print(json.dumps(d.get("batch_configs"), indent=1)[:1500])
PY
timeout -s KILL 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > $O/bench_ref_n2.json 2> $O/bench_ref_n2.err; echo "ref n2 rc=$?" | tee -a $O/summary.txt; cut -c1-300 $O/bench_ref_n2.json | tee -a $O/summary.txt
