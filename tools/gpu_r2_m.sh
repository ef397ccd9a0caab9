#!/bin/bash
# Round 2, call M: early-input tests again; value_distinct vs the size-class knobs (threshold of the long class, SMs of its partition).
mkdir -p gpurun_out
export CUDA_DEVICE_MAX_CONNECTIONS=32
timeout 900 python -m pytest tests/test_gpu_round2.py -q -x -k "early or wait_input or async or dense or knobs" > gpurun_out/r2m_pytest_early.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r2m_pytest_early.txt; tail -6 gpurun_out/r2m_pytest_early.txt
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_round2.py -q -x -k "overwrite or candidate_overflow" > gpurun_out/r2m_memcheck_early.txt 2>&1; echo "memcheck rc=$?" >> gpurun_out/r2m_memcheck_early.txt; tail -4 gpurun_out/r2m_memcheck_early.txt
run() {  # name, env...
  name=$1; shift
  env "$@" timeout 600 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu --no-verify --distinct-early ${EARLY:-0} --distinct-bufs 8 > gpurun_out/r2m_$name.txt 2>gpurun_out/r2m_$name.err
  python - "$name" <<'PY' | tee -a gpurun_out/r2m_sweep.txt
import json, sys
n = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/r2m_{n}.txt").read().strip().splitlines()[-1])
    v = d["value_distinct"]
    print(f"{n}: value {d['value']:.1f} GiB/s (K=10), value_distinct {v.get('value') or 0:.1f} GiB/s, single batch {d.get('single_batch_latency_ms', 0):.0f} ms", v.get("error", ""))
except Exception as ex:
    print(n, "failed", repr(ex))
PY
}
run A_thr25_p24
run B_thr15_p40_h64 PBSGPU_HYBRID_THR_X10=15 PBSGPU_PARTITION_SMS=40 PBSGPU_HYBRID_HEAD_PER_SM=64
run C_thr15_p32_h64 PBSGPU_HYBRID_THR_X10=15 PBSGPU_PARTITION_SMS=32 PBSGPU_HYBRID_HEAD_PER_SM=64
run D_thr10_p56_h96 PBSGPU_HYBRID_THR_X10=10 PBSGPU_PARTITION_SMS=56 PBSGPU_HYBRID_HEAD_PER_SM=96
run E_thr20_p32_h48 PBSGPU_HYBRID_THR_X10=20 PBSGPU_PARTITION_SMS=32 PBSGPU_HYBRID_HEAD_PER_SM=48
EARLY=1 run F_early_thr15_p40_h64 PBSGPU_HYBRID_THR_X10=15 PBSGPU_PARTITION_SMS=40 PBSGPU_HYBRID_HEAD_PER_SM=64 PBSGPU_ARENA_FRAC_X16=7 PBSGPU_ARENA_MB=49152
run A2_thr25_p24
