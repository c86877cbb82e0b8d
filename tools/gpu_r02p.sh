#!/bin/bash
# after the default queue at many light samples went from 12 Mi to ~20 Mi rays: GPU tests, the driver's bench line, the gem scene (largest memory use)
cd /root/repo
O=gpurun_out/r02p; mkdir -p $O
( time timeout 900 python -m pytest tests -m gpu -x -q ) > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
timeout 900 python bench.py --steps 4 --warmup 3 > $O/bench_default.json 2> $O/bench_default.err; echo "bench rc=$?" >> $O/bench_default.err
timeout 300 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-variants --quality high --gem > $O/bench_c4_gem.json 2>/dev/null
timeout 300 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-variants --no-dof > $O/bench_c3_nodof.json 2>/dev/null
nvidia-smi --query-gpu=memory.used,memory.total --format=csv > $O/mem.txt
tail -3 $O/pytest_gpu.log; python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02p/bench_*.json')):
    l=json.loads(open(f).readline()); print(f, round(l['ms_per_step'],1), round(l['spp_per_s']/1e6,1), l['gpu_launches'], l.get('variants',{}).get('nee4',{}).get('seconds_per_2048spp_frame'))
PY
