#!/bin/bash
# Round 2, call J (gpurun --gpus 2): why does the e2e leg not scale with ranks?  (r1: 2.8 s per rank at every N; now N x slower)
mkdir -p gpurun_out
export CUDA_DEVICE_MAX_CONNECTIONS=32
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port"
B="bench.py --gpus 2 --steps 2 --warmup 1 --files 128 --no-cpu --no-prewarm"
run() { name=$1; port=$2; shift; shift; env "$@" timeout 400 $T $port $B > gpurun_out/r2j_$name.txt 2> gpurun_out/r2j_$name.err; python - "$name" <<'PY'
import json,re,sys
t=open(f"gpurun_out/r2j_{sys.argv[1]}.txt").read()
m=re.findall(r'\{.*\}',t)
if m:
    d=json.loads(m[-1]); e=d['e2e']; print(sys.argv[1], 'e2e', round(e['value'],1), 'GiB/s', round(e['seconds'],2), 's  threads', e.get('host_threads'), 'h2d peak', e.get('roofline',{}).get('peak'), flush=True)
else: print(sys.argv[1], 'FAILED', t[-300:], open(f"gpurun_out/r2j_{sys.argv[1]}.err").read()[-500:])
PY
}
run fixed 29521 X=1 | tee -a gpurun_out/r2j_summary.txt
B="bench.py --gpus 2 --steps 2 --warmup 1 --files 128 --no-cpu --no-prewarm"
