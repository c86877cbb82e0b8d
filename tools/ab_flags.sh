#!/bin/bash
# bench.py with different flag sets on one box: tools/ab_flags.sh "--queue 8388608" "--queue 33554432" ...
for f in "$@"; do
  python bench.py --no-cpu-baseline --steps 4 --warmup 3 $f 2>/dev/null | python -c "
import json,sys
l=json.loads(sys.stdin.readline()); r=l['roofline']
print('[$f]', round(l['value']), 'Mrays/s', round(l['ms_per_step'],2), 'ms/step e2e', round(l['e2e']['value']), 'launches', l['gpu_launches'], 'extend share', round(r['share_of_step'],3), 'shadow share', round(r['shadow_kernel']['share_of_step'],3))
"
done
