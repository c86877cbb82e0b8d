#!/bin/bash
# Candidate lists per vertex (shaft_kernel + list evaluation in nee_kernel): bit-identity test, then A/B of variants (step cap, list size)
# against lists off (B2PT_SHAFT_MIN_NDIR=1000)
cd /root/repo
O=gpurun_out/r02n; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_configs.py -m gpu -x -q -k candidate > $O/pytest.log 2>&1; tail -2 $O/pytest.log
line() { python -c "
import json,sys
l=json.loads(sys.stdin.readline()); r=l['roofline']
print('$1', round(l['spp_per_s']/1e6,1), 'Mspp/s', round(l['ms_per_step'],2), 'ms/step  extend share', round(r['share_of_step'],3), ' shadow share', round(r['shadow_kernel']['share_of_step'],3), 'traced Mrays/s', round(l['traced_rays_per_s_M']))
"; }
B="python bench.py --no-cpu-baseline --no-variants --steps 3 --warmup 3"
B2PT_SHAFT_MIN_NDIR=1000 $B --frame-spp 256 2>/dev/null | line "chess nee32 lists off"
for v in lists192 lists96 lists64 lists64_k16; do B2PT_GPU_LIB=$PWD/variants/$v.so $B --frame-spp 256 2>/dev/null | line "chess nee32 $v"; done
B2PT_SHAFT_MIN_NDIR=1000 $B --frame-spp 256 --quality high --gem 2>/dev/null | line "chess gem high nee32 lists off"
for v in lists192 lists64; do B2PT_GPU_LIB=$PWD/variants/$v.so $B --frame-spp 256 --quality high --gem 2>/dev/null | line "chess gem high nee32 $v"; done
B2PT_SHAFT_MIN_NDIR=1000 $B --frame-spp 256 --no-dof 2>/dev/null | line "chess nodof nee32 lists off"
for v in lists192 lists64; do B2PT_GPU_LIB=$PWD/variants/$v.so $B --frame-spp 256 --no-dof 2>/dev/null | line "chess nodof nee32 $v"; done
B2PT_SHAFT_MIN_NDIR=1000 $B --scene cornell --width 1024 --height 1024 --frame-spp 256 --ndir 4 2>/dev/null | line "cornell 1024 nee4 lists off"
for v in lists192; do B2PT_SHAFT_MIN_NDIR=4 B2PT_GPU_LIB=$PWD/variants/$v.so $B --scene cornell --width 1024 --height 1024 --frame-spp 256 --ndir 4 2>/dev/null | line "cornell 1024 nee4 $v"; done
B2PT_SHAFT_MIN_NDIR=1000 $B --scene cornell --width 1024 --height 1024 --frame-spp 128 --ndir 32 2>/dev/null | line "cornell 1024 nee32 lists off"
for v in lists192; do B2PT_GPU_LIB=$PWD/variants/$v.so $B --scene cornell --width 1024 --height 1024 --frame-spp 128 --ndir 32 2>/dev/null | line "cornell 1024 nee32 $v"; done
P="python bench.py --frame-spp 64 --steps 1 --warmup 1 --no-cpu-baseline --no-variants --queue 2097152"
$P > $O/plain.log 2>&1 &&
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 1500 --csv --log-file $O/launches_chess_nee32_lists.csv $P > $O/ncu_launches.log 2>&1
