#!/bin/bash
# Runs ON THE GPU BOX (via gpurun): the round-2 ncu evidence in one go. Every command runs plainly first.
#   launches_r02.csv        launch list of the default bench command (headline only, 2 steps)
#   full_r02_k1             --set full of eval_kernel<4,2>
#   full_r02_k4             --set full of tiled_eval_kernel<32,8>
#   full_r02_sweepN50       --set full of mpc_solve / simulate / bounds kernels at N = 50, 1e6 samples
#   full_r02_dyn            --set full of dyn_eval_kernel at (8, 2) (run-time-dimension route)
set -u
OUT=gpurun_out; mkdir -p $OUT
PART=${1:-all}     # a: launch lists + K1 + K4 + dyn, b: the three sweep kernels (gpurun merges at most 64 MiB back)
if [ "$PART" != "b" ]; then
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --headline-only"
$B > $OUT/plain_r02.log 2> $OUT/plain_r02.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_r02.csv $B > $OUT/ncu_launches_r02.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:eval_kernel -s 3 -c 1 -o $OUT/full_r02_k1 -f $B > $OUT/ncu_full_r02_k1.log 2>&1
B4="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --workload cfg-synth-32-8-30"
$B4 > $OUT/plain_r02_k4.log 2> $OUT/plain_r02_k4.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_r02k4.csv $B4 > $OUT/ncu_launches_r02k4.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:tiled_eval_kernel -s 4 -c 1 -o $OUT/full_r02_k4 -f $B4 > $OUT/ncu_full_r02_k4.log 2>&1
fi
if [ "$PART" = "b" ] || [ "$PART" = "all" ]; then
S="python scripts/sweep_probe.py --N 50 --samples 1000000"
$S > $OUT/plain_r02_sweepN50.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'mpc_solve_kernel|simulate_kernel|bounds_kernel' -s 3 -c 3 -o $OUT/full_r02_sweepN50 -f $S > $OUT/ncu_full_r02_sweepN50.log 2>&1
fi
if [ "$PART" != "b" ]; then
D="python scripts/dims_probe.py 8 2 10 100000"
LQMPC_FORCE_DYN=1 $D > $OUT/plain_r02_dyn.log 2>&1 &&
LQMPC_FORCE_DYN=1 ncu --set full --clock-control none --import-source on -k regex:dyn_eval_kernel -s 1 -c 1 -o $OUT/full_r02_dyn -f $D > $OUT/ncu_full_r02_dyn.log 2>&1
fi
echo done
