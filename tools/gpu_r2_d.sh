#!/bin/bash
# Round 2, call D (run with gpurun --gpus 2): the 2-rank NCCL merge test, the single-set expectation for cfg4
# (profiles/r02_cfg4_expected.json) and the N=2 bench line (cfg4 through pbsgpu_set_allgather).
mkdir -p gpurun_out
export CUDA_DEVICE_MAX_CONNECTIONS=32
nvidia-smi -L | tee gpurun_out/r2d_gpus.txt
timeout 600 python -m pytest tests/test_gpu_round2.py -q -k "allgather" > gpurun_out/r2d_pytest_allgather.txt 2>&1; echo "rc=$?" >> gpurun_out/r2d_pytest_allgather.txt; tail -5 gpurun_out/r2d_pytest_allgather.txt
timeout 900 python bench.py --workload cfg4verify --emulate-ranks 2,4,8 > gpurun_out/r2d_cfg4_expected.json 2> gpurun_out/r2d_cfg4_expected.err; tail -c 800 gpurun_out/r2d_cfg4_expected.json; tail -3 gpurun_out/r2d_cfg4_expected.err
mkdir -p profiles; python - <<'PY'
import json,re
t=open('gpurun_out/r2d_cfg4_expected.json').read()
m=re.findall(r'\{.*\}',t)
if m: open('profiles/r02_cfg4_expected.json','w').write(json.dumps(json.loads(m[-1]),indent=1)+"\n")
PY
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 --no-cpu > gpurun_out/r2d_bench_n2.txt 2> gpurun_out/r2d_bench_n2.err; tail -c 1500 gpurun_out/r2d_bench_n2.txt; tail -5 gpurun_out/r2d_bench_n2.err
cp profiles/r02_cfg4_expected.json gpurun_out/ 2>/dev/null
