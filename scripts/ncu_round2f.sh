#!/bin/bash
# Runs ON THE GPU BOX (via gpurun): ncu evidence for the tree at the end of round 2 (every command runs plainly first).
#   launches_r02f.csv / full_r02f_k1      default bench command (headline only): launch list + eval_kernel<4,2>
#   launches_r02fk4.csv / full_r02f_k4    tiled_eval_kernel<32,8> (fully unrolled warp tiles)
#   full_r02f_sweepN50                    mpc_solve / simulate / bounds kernels at N = 50, 1e6 samples
set -u
OUT=gpurun_out; mkdir -p $OUT
PART=${1:-all}
if [ "$PART" != "b" ]; then
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --headline-only"
$B > $OUT/plain_r02f.log 2> $OUT/plain_r02f.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_r02f.csv $B > $OUT/ncu_launches_r02f.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:eval_kernel -s 3 -c 1 -o $OUT/full_r02f_k1 -f $B > $OUT/ncu_full_r02f_k1.log 2>&1
B4="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --workload cfg-synth-32-8-30"
$B4 > $OUT/plain_r02f_k4.log 2> $OUT/plain_r02f_k4.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_r02fk4.csv $B4 > $OUT/ncu_launches_r02fk4.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:tiled_eval_kernel -s 4 -c 1 -o $OUT/full_r02f_k4 -f $B4 > $OUT/ncu_full_r02f_k4.log 2>&1
fi
if [ "$PART" != "a" ]; then
S="python scripts/sweep_probe.py --N 50 --samples 1000000"
$S > $OUT/plain_r02f_sweepN50.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'mpc_solve_kernel|simulate_kernel|bounds_kernel' -s 3 -c 3 -o $OUT/full_r02f_sweepN50 -f $S > $OUT/ncu_full_r02f_sweepN50.log 2>&1
fi
echo done
