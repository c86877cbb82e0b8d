#!/bin/bash
# Source-level counters of the shading-side kernels (reduced queue, like tools/gpu_r02k.sh)
cd /root/repo
O=gpurun_out/r02m; mkdir -p $O
P="python bench.py --width 480 --height 270 --frame-spp 64 --steps 1 --warmup 1 --no-cpu-baseline --no-variants --queue 262144"
$P > $O/plain.log 2>&1 || exit 1
for k in nee_kernel light_kernel generate_kernel terminal_kernel nee_eval_kernel; do
  timeout 300 ncu --section SourceCounters --section InstructionStats --import-source on --clock-control none -k regex:"$k" -s 3 -c 1 -o $O/src_$k $P > $O/ncu_$k.log 2>&1
  echo "$k rc=$?"
done
timeout 300 ncu --section SourceCounters --section InstructionStats --import-source on --clock-control none -k regex:"shade_kernel<1, 1>|shade_kernel<2, 0>" -s 6 -c 2 -o $O/src_shade $P > $O/ncu_shade.log 2>&1
echo "shade rc=$?"
ls -la $O
