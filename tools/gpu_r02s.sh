#!/bin/bash
# the driver's bench line on the final build of the round
cd /root/repo
O=gpurun_out/r02s; mkdir -p $O
timeout 600 python bench.py --steps 4 --warmup 3 > $O/bench_default.json 2> $O/bench_default.err; echo "bench rc=$?"
python - <<'PY'
import json
l=json.loads(open('gpurun_out/r02s/bench_default.json').readline())
print(round(l['ms_per_step'],1), round(l['spp_per_s']/1e6,1), round(l['value']), round(l['e2e']['value']), l['gpu_launches'], (l.get('variants') or {}).get('nee4',{}).get('seconds_per_2048spp_frame'), round(l['roofline']['frac'],3), round(l['roofline']['shadow_kernel']['frac'],3), l['clocks']['reasons'], round(l['traced_rays_per_s_M']))
PY
