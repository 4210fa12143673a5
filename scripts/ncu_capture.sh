#!/bin/bash
# Runs ON THE GPU BOX (via gpurun): plain bench run first, then the ncu launch list and one --set full capture.
# Usage: gpurun --timeout 900 -- 'bash scripts/ncu_capture.sh <tag> [workload] [kernel-regex] [skip]'
#   bench launches per run of the top kernel: 3 warm-up + 2 timed steps + 2 kernel-only (+ e2e chunks); `skip` jumps
#   over the warm-up (and, for K4, the 1-sample problem-preparation launch).
set -u
TAG=${1:-r01}
WL=${2:-cfg-synth-4-2-10}
RX=${3:-eval_kernel}
SKIP=${4:-3}
OUT=gpurun_out
mkdir -p $OUT
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --workload $WL"
$CMD > $OUT/plain_$TAG.log 2> $OUT/plain_$TAG.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_$TAG.csv $CMD > $OUT/ncu_launches_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:$RX -s $SKIP -c 1 -o $OUT/full_$TAG -f $CMD > $OUT/ncu_full_$TAG.log 2>&1
echo "rc=$?"
tail -3 $OUT/plain_$TAG.log
