python -m pytest tests -m gpu -x -q 2>&1 | tail -2
for v in "" $VARIANTS; do
  if [ -n "$v" ]; then export LQMPC_LIB=$PWD/lq_mpc_b200/_lib/variants/$v.so; fi
  python bench.py --workload cfg-sweep-f --no-cpu-baseline > gpurun_out/sweep_$v.json 2> gpurun_out/sweep_$v.err
  python -c "
import json,sys
d=json.loads(open('gpurun_out/sweep_$v.json').read().strip().splitlines()[-1])
print('$v', d['value'], d['ms_per_step'], d.get('phase_seconds'), d['roofline']['frac'])
"
done
