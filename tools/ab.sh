#!/bin/bash
# A/B of kernel variants on one box: tools/ab.sh variants/a.so variants/b.so ...  (each: bench.py without the CPU baseline)
for v in "$@"; do
  B2PT_GPU_LIB=$PWD/$v python bench.py --no-cpu-baseline --no-variants --frame-spp ${SPP:-256} ${BENCH_ARGS} --steps 4 --warmup 3 2>/dev/null | python -c "
import json,sys
l=json.loads(sys.stdin.readline()); r=l['roofline']
print('$v', round(l['value']), 'Mrays/s', round(l['ms_per_step'],2), 'ms/step  extend', round(r['avg_launch_ms'],3), 'ms', round(r['share_of_step'],3), 'nodes/ray', round(r['nodes_per_ray'],2), 'tris', round(r['tris_per_ray'],2), ' shadow share', round(r['shadow_kernel']['share_of_step'],3), 'nodes', round(r['shadow_kernel']['nodes_per_ray'],2))
"
done
