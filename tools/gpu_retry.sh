#!/bin/bash
# usage: tools/gpu_retry.sh <timeout> <gpus> '<command>'  -- retries gpurun while the pod answers busy (exit 3)
T=$1; G=$2; shift 2
for i in $(seq 1 20); do
  if [ "$G" = "1" ]; then /usr/local/graft/bin/gpurun --timeout $T -- "$@"; else /usr/local/graft/bin/gpurun --gpus $G --timeout $T -- "$@"; fi
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 45
done
exit 3
