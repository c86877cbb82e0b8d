#!/bin/bash
# queue target at 32 light samples per vertex (library default 12 Mi rays; the 32-bit visibility-slot numbering allows up to 22 Mi)
cd /root/repo
O=gpurun_out/r02o; mkdir -p $O
for q in 6291456 12582912 16777216 20971520; do
  python bench.py --no-cpu-baseline --no-variants --steps 3 --warmup 3 --queue $q 2>/dev/null | python -c "
import json,sys
l=json.loads(sys.stdin.readline()); r=l['roofline']
print('queue $q', round(l['spp_per_s']/1e6,1), 'Mspp/s', round(l['ms_per_step'],2), 'ms/frame  launches', l['gpu_launches'], 'extend share', round(r['share_of_step'],3), 'shadow share', round(r['shadow_kernel']['share_of_step'],3))
"
done | tee $O/queue_sweep_nee32.txt
