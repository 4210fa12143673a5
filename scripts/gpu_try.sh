#!/bin/bash
# Development helper: retry a gpurun call while the pod answers "transient" (no slot free, nothing charged).
# usage: scripts/gpu_try.sh <timeout_s> <gpus> '<command>'
T=$1; G=$2; shift 2
for i in $(seq 1 40); do
  if [ "$G" = "1" ]; then out=$(/usr/local/graft/bin/gpurun --timeout "$T" -- "$@" 2>&1); else out=$(/usr/local/graft/bin/gpurun --gpus "$G" --timeout "$T" -- "$@" 2>&1); fi
  if echo "$out" | grep -q "status=transient\|exit code 3\|no box"; then sleep 60; continue; fi
  echo "$out" | tail -80
  exit 0
done
echo "gave up after 40 transient answers"
