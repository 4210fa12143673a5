#!/bin/bash
# Runs ON THE GPU BOX (via gpurun): plain sweep_probe run, then ONE ncu --set full pass that captures the three sweep
# kernels (K2a, K2b, K3) of the full-size launches (the 3 warm-up launches and the sampler / dlqr launches are skipped
# by the kernel regex + launch-skip).  Usage: bash scripts/ncu_sweep_capture.sh <tag> <N> [samples]
set -u
TAG=${1:-r02sweep}; N=${2:-50}; S=${3:-1000000}
OUT=gpurun_out; mkdir -p $OUT
CMD="python scripts/sweep_probe.py --N $N --samples $S"
$CMD > $OUT/plain_$TAG.log 2> $OUT/plain_$TAG.err &&
ncu --set full --clock-control none --import-source on -k regex:'mpc_solve_kernel|simulate_kernel|bounds_kernel' -s 3 -c 3 -o $OUT/full_$TAG -f $CMD > $OUT/ncu_full_$TAG.log 2>&1
echo "rc=$?"; cat $OUT/plain_$TAG.log
