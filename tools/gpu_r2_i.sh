#!/bin/bash
# Round 2, call I: the high-priority "mid" class of the bulk SHA launch (PBSGPU_BULK_MID_X10): parity with it on, then the headline.
mkdir -p gpurun_out
export CUDA_DEVICE_MAX_CONNECTIONS=32
PBSGPU_BULK_MID_X10=15 timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py -q -x -k "cfg2 or cfg5 or chunk_digest or golden or structured or async or mostly_long or all_long or knobs" > gpurun_out/r2i_pytest_mid.txt 2>&1; echo "pytest(mid) rc=$?" >> gpurun_out/r2i_pytest_mid.txt; tail -3 gpurun_out/r2i_pytest_mid.txt
B="python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu --no-verify --no-distinct"
run() { name=$1; shift; env "$@" timeout 300 $B > gpurun_out/r2i_sweep_$name.txt 2>&1; python - "$name" <<'PY'
import json,re,sys
t=open(f"gpurun_out/r2i_sweep_{sys.argv[1]}.txt").read()
m=re.findall(r'\{.*\}',t)
if m:
    d=json.loads(m[-1]); print(sys.argv[1], round(d['value'],1), 'GiB/s', round(d['ms_per_step'],1), 'ms/step  iso', round(d['single_batch_latency_ms']), flush=True)
else: print(sys.argv[1], 'FAILED', t[-300:])
PY
}
for rep in 1 2; do
run mid0_$rep X=1 | tee -a gpurun_out/r2i_sweep.txt
run mid15_s10_$rep PBSGPU_BULK_MID_X10=15 PBSGPU_SLOTS=10 | tee -a gpurun_out/r2i_sweep.txt
run mid12_s10_$rep PBSGPU_BULK_MID_X10=12 PBSGPU_SLOTS=10 | tee -a gpurun_out/r2i_sweep.txt
run mid20_s10_$rep PBSGPU_BULK_MID_X10=20 PBSGPU_SLOTS=10 | tee -a gpurun_out/r2i_sweep.txt
run mid15_s13_$rep PBSGPU_BULK_MID_X10=15 | tee -a gpurun_out/r2i_sweep.txt
done
PBSGPU_BULK_MID_X10=15 PBSGPU_SLOTS=10 timeout 300 $B --timeline > gpurun_out/r2i_timeline_mid15.txt 2>&1
