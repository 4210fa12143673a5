#!/bin/bash
# Runs ON THE GPU BOX (via gpurun): plain bench run first, then the ncu launch list and one --set full capture of K1.
# Usage: gpurun --timeout 900 -- 'bash scripts/ncu_capture.sh r01'
set -u
TAG=${1:-r01}
# NOTE: K1 launches per bench run = 3 warm-up + 2 timed steps + 2 kernel-only + e2e chunks; -s 3 skips the warm-up
OUT=gpurun_out
mkdir -p $OUT
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$CMD > $OUT/plain_$TAG.log 2> $OUT/plain_$TAG.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_$TAG.csv $CMD > $OUT/ncu_launches_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:eval_kernel -s 3 -c 1 -o $OUT/k1_full_$TAG -f $CMD > $OUT/ncu_full_$TAG.log 2>&1
echo "rc=$?"
tail -3 $OUT/plain_$TAG.log
