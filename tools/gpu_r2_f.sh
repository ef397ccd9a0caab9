#!/bin/bash
# Round 2, call F: 3-way SM partition (dedicated scan partition): tests, then the pipelined headline for several partition sizes.
mkdir -p gpurun_out
export CUDA_DEVICE_MAX_CONNECTIONS=32
timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/r2f_pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r2f_pytest_gpu.txt; tail -5 gpurun_out/r2f_pytest_gpu.txt
B="python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu --no-verify --no-distinct"
run() { name=$1; shift; env "$@" timeout 300 $B > gpurun_out/r2f_sweep_$name.txt 2>&1; python - "$name" <<'PY'
import json,re,sys
t=open(f"gpurun_out/r2f_sweep_{sys.argv[1]}.txt").read()
m=re.findall(r'\{.*\}',t)
if m:
    d=json.loads(m[-1]); print(sys.argv[1], round(d['value'],1), 'GiB/s', round(d['ms_per_step'],1), 'ms/step  iso', round(d['single_batch_latency_ms']), d['config'].get('sm_partition(long,bulk,scan)'), flush=True)
else: print(sys.argv[1], 'FAILED', t[-300:])
PY
}
run scan24 PBSGPU_DEBUG=1 | tee -a gpurun_out/r2f_sweep.txt
run scan0 PBSGPU_SCAN_SMS=0 | tee -a gpurun_out/r2f_sweep.txt
run scan16 PBSGPU_SCAN_SMS=16 | tee -a gpurun_out/r2f_sweep.txt
run scan32 PBSGPU_SCAN_SMS=32 | tee -a gpurun_out/r2f_sweep.txt
run scan24_long16 PBSGPU_PARTITION_SMS=16 | tee -a gpurun_out/r2f_sweep.txt
run scan32_long16 PBSGPU_SCAN_SMS=32 PBSGPU_PARTITION_SMS=16 | tee -a gpurun_out/r2f_sweep.txt
run scan24_again X=1 | tee -a gpurun_out/r2f_sweep.txt
run scan0_again PBSGPU_SCAN_SMS=0 | tee -a gpurun_out/r2f_sweep.txt
run scan24_slots20 PBSGPU_SLOTS=13 X=2 | tee -a gpurun_out/r2f_sweep.txt
B="python bench.py --steps 32 --warmup 3 --no-e2e --no-cpu --no-verify --no-distinct"
run scan24_k32 X=1 | tee -a gpurun_out/r2f_sweep.txt
run scan0_k32 PBSGPU_SCAN_SMS=0 | tee -a gpurun_out/r2f_sweep.txt
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu --no-verify > gpurun_out/r2f_bench.txt 2>gpurun_out/r2f_bench.err; tail -c 1800 gpurun_out/r2f_bench.txt; tail -3 gpurun_out/r2f_bench.err
PBSGPU_SCAN_SMS=0 timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu --no-verify --no-e2e > gpurun_out/r2f_bench_scan0.txt 2>&1; tail -c 900 gpurun_out/r2f_bench_scan0.txt
