# development: headline workload of bench.py for the default library and the variants named in $VARIANTS
for v in "" $VARIANTS; do
  if [ -n "$v" ]; then export LQMPC_LIB=$PWD/lq_mpc_b200/_lib/variants/$v.so; fi
  python bench.py --headline-only --no-cpu-baseline --steps 30 > gpurun_out/k1_$v.json 2> gpurun_out/k1_$v.err
  python -c "
import json
d=json.loads(open('gpurun_out/k1_$v.json').read().strip().splitlines()[-1])
print('k1 variant [$v]', d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline'].get('kernel_ms'), d['clocks'])
"
done
