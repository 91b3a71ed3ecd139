#!/bin/bash
# usage: tools/gpurun_retry.sh <timeout> <gpus> '<command>'   -- retries while the pod answers busy (exit 3 / "transient")
T=$1; G=$2; shift 2
for i in $(seq 1 20); do
  OUT=$(/usr/local/graft/bin/gpurun --gpus $G --timeout $T -- "$@" 2>&1)
  if echo "$OUT" | grep -q "status=transient"; then sleep 90; continue; fi
  echo "$OUT" | tail -30
  exit 0
done
echo "gave up: pod busy"
