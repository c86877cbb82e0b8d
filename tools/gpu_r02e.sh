#!/bin/bash
# Round-2 session e: compressed (8-bit) four-wide quads for extend and shadow — tests, A/B, queue sizes at 32 light samples.
cd /root/repo
O=gpurun_out/r02e; mkdir -p $O
( time timeout 900 python -m pytest tests -m gpu -x -q ) > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
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
ab default "--frame-spp 256" default libb2pt_shbin.so libb2pt_b9.so libb2pt_b10.so libb2pt_b6.so
ab nee4 "--frame-spp 512 --ndir 4" default libb2pt_b9.so libb2pt_b6.so
ab cornell "--scene cornell --width 1024 --height 1024 --frame-spp 256 --ndir 4" default libb2pt_shbin.so libb2pt_b9.so libb2pt_b6.so
ab gem "--quality high --gem --frame-spp 128" default libb2pt_b9.so
timeout 300 python bench.py $B --frame-spp 256 --queue 16777216 > $O/ab_default_q16.json 2>/dev/null
timeout 300 python bench.py $B --frame-spp 256 --queue 25165824 > $O/ab_default_q24.json 2>/dev/null
timeout 300 python bench.py $B --frame-spp 256 --queue 8388608 > $O/ab_default_q8.json 2>/dev/null
ls $O
