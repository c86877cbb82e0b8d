#!/bin/bash
# Round-2 session d: four lanes per ray (extend + shadow) — tests, then A/B against the one-lane walks and tuning variants.
cd /root/repo
O=gpurun_out/r02d; mkdir -p $O
( time timeout 900 python -m pytest tests -m gpu -x -q ) > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
grep -q "rc=0" $O/pytest_gpu.log || { tail -40 $O/pytest_gpu.log; }
B="--steps 3 --warmup 2 --no-cpu-baseline --no-variants"
ab() { # name, bench args, variants...
  name=$1; args=$2; shift 2
  for v in "$@"; do
    tag=$(basename $v .so)
    if [ "$v" = default ]; then unset B2PT_GPU_LIB; else export B2PT_GPU_LIB=$PWD/variants/$v; fi
    timeout 300 python bench.py $B $args > $O/ab_${name}_$tag.json 2>/dev/null
  done
  unset B2PT_GPU_LIB
}
ab default "--frame-spp 256" default libb2pt_lane.so libb2pt_coopext.so libb2pt_coopsh.so libb2pt_refill3.so libb2pt_refill7.so libb2pt_b8.so libb2pt_b12.so
ab nee4 "--frame-spp 512 --ndir 4" default libb2pt_lane.so libb2pt_b8.so libb2pt_b12.so
ab cornell "--scene cornell --width 1024 --height 1024 --frame-spp 256 --ndir 4" default libb2pt_lane.so libb2pt_coopext.so libb2pt_coopsh.so libb2pt_b12.so
ab gem "--quality high --gem --frame-spp 128" default libb2pt_lane.so
ls $O
