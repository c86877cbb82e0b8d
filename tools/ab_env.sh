#!/bin/bash
# bench.py under different environments / flags on one box: tools/ab_env.sh "VAR=1 VAR2=x|--flags" ...
for spec in "$@"; do
  envs="${spec%%|*}"; flags="${spec#*|}"
  env $envs python bench.py --no-cpu-baseline --steps 4 --warmup 3 $flags 2>/dev/null | python -c "
import json,sys
l=json.loads(sys.stdin.readline()); r=l['roofline']
print('[$spec]', round(l['value']), 'Mrays/s', round(l['ms_per_step'],2), 'ms/step e2e', round(l['e2e']['value']), 'launches', l['gpu_launches'], 'extend share', round(r['share_of_step'],3), 'shadow share', round(r['shadow_kernel']['share_of_step'],3))
"
done
