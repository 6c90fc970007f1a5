#!/bin/bash
# Round 2, GPU call 52 (eight B200s): the bench contract under torchrun at N = 8 on the final tree -- CTA-pair convs, GroupNorm statistics in the conv epilogues -- (sharded config 3: 4 images per GPU;
# config 5: one HiresFix image per GPU), then the reference arm at N = 8 (rank 0 alone works).
set -u
O=gpurun_out/r2_call52; mkdir -p $O
N=$(nvidia-smi -L | wc -l); echo "gpus: $N" | tee -a $O/summary.txt
timeout -s KILL 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 20 --warmup 5 > $O/bench_n$N.json 2> $O/bench_n$N.err; echo "bench n$N rc=$?" | tee -a $O/summary.txt
python - <<PY | tee -a $O/summary.txt
import json
for line in open("$O/bench_n$N.json"):
    line=line.strip()
    if not line.startswith("{"): continue
    d=json.loads(line)
    print("n_gpus", d["n_gpus"], "it/s", round(d["value"],2), "ms", round(d["ms_per_step"],3), "e2e", d.get("e2e",{}).get("value"))
    print(json.dumps(d.get("sharded"), indent=1)[:1500])
PY
grep -v "Warning: \[PG ID" $O/bench_n$N.err | tail -5 | tee -a $O/summary.txt
