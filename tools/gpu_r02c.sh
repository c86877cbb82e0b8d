#!/bin/bash
# Round-2 session c: padded four-wide walk (FMA plane rows) for extend and shadow, shared-wavelength eval, vertex culling.
cd /root/repo
O=gpurun_out/r02c; mkdir -p $O
( time timeout 900 python -m pytest tests -m gpu -x -q ) > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
B="--steps 3 --warmup 2 --no-cpu-baseline --no-variants"
ab() { # name, bench args
  for v in default variants/libb2pt_shbin.so variants/libb2pt_b8.so; do
    tag=$(basename $v .so)
    if [ "$v" = default ]; then unset B2PT_GPU_LIB; else export B2PT_GPU_LIB=$PWD/$v; fi
    timeout 300 python bench.py $B $2 > $O/ab_$1_$tag.json 2>/dev/null
  done
  unset B2PT_GPU_LIB
}
ab default "--frame-spp 256"
ab nee4 "--frame-spp 512 --ndir 4"
ab cornell "--scene cornell --width 1024 --height 1024 --frame-spp 256 --ndir 4"
ab gem "--quality high --gem --frame-spp 128"
ls -la $O
