#!/bin/bash
N=${1:-2}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 \
   bench.py --gpus $N --steps 8 --warmup 3 --no-cpu > gpurun_out/bench_e2e_n$N.txt 2>&1
echo "rc=$?"
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/bench_e2e_n$N.txt").read().splitlines() if l.startswith("{")][-1])
print("value", round(d["value"],1), "e2e", d.get("e2e"))
PY
tail -3 gpurun_out/bench_e2e_n$N.txt | cut -c1-300
