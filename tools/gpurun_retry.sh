#!/bin/bash
# usage: tools/gpurun_retry.sh <timeout_s> <command> [gpus]  -- retries while the pod answers busy (exit code 3)
T=$1; CMD=$2; G=${3:-1}
for i in $(seq 1 40); do
  if [ "$G" -gt 1 ]; then /usr/local/graft/bin/gpurun --gpus "$G" --timeout "$T" -- "$CMD"; else /usr/local/graft/bin/gpurun --timeout "$T" -- "$CMD"; fi
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  echo "[retry] busy, attempt $i; sleeping 90 s"
  sleep 90
done
exit 3
