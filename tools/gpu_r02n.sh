#!/bin/bash
# Candidate lists per vertex (shaft_kernel + list_visible): parity tests, then on/off A/B (B2PT_SHAFT_MIN_NDIR = n_dir from which lists are built)
cd /root/repo
O=gpurun_out/r02n; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_configs.py tests/test_gpu_parity.py tests/test_statistical.py -m gpu -x -q > $O/pytest.log 2>&1; tail -3 $O/pytest.log
line() { python -c "
import json,sys
l=json.loads(sys.stdin.readline()); r=l['roofline']
print('$1', round(l['spp_per_s']/1e6,1), 'Mspp/s', round(l['ms_per_step'],2), 'ms/step  extend share', round(r['share_of_step'],3), ' shadow share', round(r['shadow_kernel']['share_of_step'],3), 'traced Mrays/s', round(l['traced_rays_per_s_M']))
"; }
B="python bench.py --no-cpu-baseline --no-variants --steps 3 --warmup 3"
for m in 1000 8; do B2PT_SHAFT_MIN_NDIR=$m $B --frame-spp 256 2>/dev/null | line "chess nee32 min_ndir=$m"; done
for m in 1000 8; do B2PT_SHAFT_MIN_NDIR=$m $B --frame-spp 256 --quality high --gem 2>/dev/null | line "chess gem high nee32 min_ndir=$m"; done
for m in 1000 4; do B2PT_SHAFT_MIN_NDIR=$m $B --frame-spp 256 --ndir 4 2>/dev/null | line "chess nee4 min_ndir=$m"; done
for m in 1000 4; do B2PT_SHAFT_MIN_NDIR=$m $B --scene cornell --width 1024 --height 1024 --frame-spp 256 --ndir 4 2>/dev/null | line "cornell 1024 nee4 min_ndir=$m"; done
for m in 1000 4; do B2PT_SHAFT_MIN_NDIR=$m $B --scene sweep:clear_rough_plastic --frame-spp 256 --ndir 4 2>/dev/null | line "sweep plastic nee4 min_ndir=$m"; done
for m in 1000 8; do B2PT_SHAFT_MIN_NDIR=$m $B --scene cornell --width 1024 --height 1024 --frame-spp 128 --ndir 32 2>/dev/null | line "cornell 1024 nee32 min_ndir=$m"; done
