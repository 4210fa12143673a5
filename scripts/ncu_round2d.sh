#!/bin/bash
# Runs ON THE GPU BOX (via gpurun): ncu evidence for the last round-2 kernel changes (every command runs plainly first).
#   launches_r02dk4.csv / full_r02d_k4   tiled_eval_kernel<32,8> with the trace-based spectral-radius acceptance
#   full_r02d_group                      group_eval_kernel<8,2,4> (lane-group K1)
set -u
OUT=gpurun_out; mkdir -p $OUT
B4="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --workload cfg-synth-32-8-30"
$B4 > $OUT/plain_r02d_k4.log 2> $OUT/plain_r02d_k4.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_r02dk4.csv $B4 > $OUT/ncu_launches_r02dk4.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:tiled_eval_kernel -s 4 -c 1 -o $OUT/full_r02d_k4 -f $B4 > $OUT/ncu_full_r02d_k4.log 2>&1
D="python scripts/dims_probe.py 8 2 10 500000"
$D > $OUT/plain_r02d_group.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:group_eval_kernel -s 1 -c 1 -o $OUT/full_r02d_group -f $D > $OUT/ncu_full_r02d_group.log 2>&1
echo done
