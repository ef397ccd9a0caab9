#!/bin/bash
# Round 2, call O: zstd-framed blobs (parity + sanitizer) and the ncu launch list of a short bench run (unpartitioned, see gpu_r2_k.sh).
mkdir -p gpurun_out
export CUDA_DEVICE_MAX_CONNECTIONS=32
timeout 600 python -m pytest tests/test_gpu_round2.py -q -x -k "blob" > gpurun_out/r2o_pytest_blob.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r2o_pytest_blob.txt; tail -12 gpurun_out/r2o_pytest_blob.txt
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_round2.py -q -x -k "blob_encode_batch_z_builds" > gpurun_out/r2o_memcheck_blob.txt 2>&1; echo "memcheck rc=$?" >> gpurun_out/r2o_memcheck_blob.txt; tail -3 gpurun_out/r2o_memcheck_blob.txt
PBSGPU_PARTITION_SMS=0 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2o_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu --no-verify --no-distinct --no-prewarm > gpurun_out/r2o_launches_bench.log 2>&1; echo "ncu rc=$?"; tail -c 300 gpurun_out/r2o_launches_bench.log; wc -l gpurun_out/r2o_launches.csv
python - <<'PY'
# throughput of the zstd path on a sparse image: 8 GiB, half of it zero runs
import time, numpy as np, torch, pbs_plus_b200 as pg
e = pg.Engine(0)
n = 8 << 30
d = torch.randint(0, 256, (n,), dtype=torch.uint8, device="cuda")
d[1 << 30: 5 << 30] = 0
L = 4 << 20
off = np.arange(0, n, L, dtype=np.uint64); ln = np.full(len(off), L, dtype=np.uint64)
for name, f in (("blob_encode_batch", e.blob_encode_batch), ("blob_encode_batch_z", e.blob_encode_batch_z)):
    f(d, off[:64], ln[:64])
    torch.cuda.synchronize(); t0 = time.perf_counter(); r = f(d, off, ln); dt = time.perf_counter() - t0
    size = sum(len(b) for b in r[0]) if isinstance(r[0], list) else len(r[0])
    print(f"{name}: {n / dt / 1e9:.1f} GB/s of input ({dt * 1e3:.0f} ms), output {size / 2**30:.2f} GiB")
e.close()
PY
