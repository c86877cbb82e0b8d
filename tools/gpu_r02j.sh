#!/bin/bash
# Round-2 session j: treeless walk for small scenes (Cornell box) — tests and A/B (B2PT_NO_FLAT=1 disables it at run time).
cd /root/repo
O=gpurun_out/r02j; mkdir -p $O
( time timeout 900 python -m pytest tests -m gpu -x -q ) > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
B="--steps 3 --warmup 2 --no-cpu-baseline --no-variants"
for nf in 0 1; do
  if [ $nf = 1 ]; then export B2PT_NO_FLAT=1; else unset B2PT_NO_FLAT; fi
  timeout 300 python bench.py $B --scene cornell --width 1024 --height 1024 --frame-spp 256 --ndir 4 > $O/ab_cornell1024_noflat$nf.json 2>/dev/null
  timeout 300 python bench.py $B --scene cornell --width 512 --height 512 --frame-spp 32 --ndir 4 > $O/ab_cornell512spp32_noflat$nf.json 2>/dev/null
  timeout 300 python bench.py $B --scene sweep:clear_rough_plastic --frame-spp 256 --ndir 4 > $O/ab_plastic_noflat$nf.json 2>/dev/null
  timeout 300 python bench.py $B --scene sweep:gold_conductor --frame-spp 256 --ndir 4 > $O/ab_gold_noflat$nf.json 2>/dev/null
done
unset B2PT_NO_FLAT
timeout 300 python tools/find_nan.py 2048 32 > $O/find_nan_2048.txt 2>&1
ls $O
