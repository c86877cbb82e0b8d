#!/bin/bash
# last numbers of the round on the committed build: smoke, the driver's bench line, the reference arm
cd /root/repo
O=gpurun_out/r02q; mkdir -p $O
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/smoke.log; tail -2 $O/smoke.log
timeout 900 python bench.py --steps 4 --warmup 3 > $O/bench_default.json 2> $O/bench_default.err; echo "bench rc=$?"
timeout 300 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-variants --scene cornell --width 512 --height 512 --frame-spp 2048 --ndir 4 > $O/bench_c1_cornell_512_spp2048.json 2>/dev/null
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02q/bench_*.json')):
    l=json.loads(open(f).readline()); print(f, round(l['ms_per_step'],1), round(l['spp_per_s']/1e6,1), l['gpu_launches'], l.get('variants',{}).get('nee4',{}).get('seconds_per_2048spp_frame'), l['roofline']['frac'], l['e2e']['value'], l['cpu_baseline'] and l['cpu_baseline'].get('spp_per_s'))
PY
