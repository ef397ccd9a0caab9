#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_sha -c 8 --csv --log-file gpurun_out/launches_hyb.csv \
   python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/ncu_hyb.log 2>&1
grep -E "k_sha" gpurun_out/launches_hyb.csv | awk -F'","' '{print $5, $(NF-1), $NF}' | head -12
