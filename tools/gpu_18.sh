#!/bin/bash
mkdir -p gpurun_out
PBSGPU_CRC_VARIANT=0 timeout 600 python -m pytest tests -x -q -m gpu -k "crc32" > gpurun_out/pytest_crc.txt 2>&1; tail -3 gpurun_out/pytest_crc.txt
PBSGPU_CRC_VARIANT=1 timeout 600 python -m pytest tests -x -q -m gpu -k "crc32" 2>&1 | tail -1
PBSGPU_CRC_VARIANT=0 timeout 300 python tools/crc_bench.py 2>&1 | tail -1 | tee gpurun_out/crc_bench.txt
PBSGPU_CRC_VARIANT=1 timeout 300 python tools/crc_bench.py 2>&1 | tail -1 | tee -a gpurun_out/crc_bench.txt
PBSGPU_CRC_VARIANT=0 PBSGPU_PARTITION_SMS=0 timeout 300 python tools/crc_bench.py 2>&1 | tail -1 | tee -a gpurun_out/crc_bench.txt
