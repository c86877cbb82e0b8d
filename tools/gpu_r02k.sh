#!/bin/bash
# Source-level counters of the extend kernel on a SMALL problem (the SASS-patched pass does not finish on full-size queues).
cd /root/repo
O=gpurun_out/r02k; mkdir -p $O
P="python bench.py --width 480 --height 270 --frame-spp 64 --steps 1 --warmup 1 --no-cpu-baseline --no-variants --queue 262144"
$P > $O/plain.log 2>&1 &&
timeout 420 ncu --section SourceCounters --section InstructionStats --import-source on --clock-control none -k regex:'extend_kernel' -s 3 -c 1 -o $O/prof_extend_src $P > $O/ncu_src.log 2>&1
echo "ncu rc=$?" >> $O/ncu_src.log
$P > $O/plain2.log 2>&1 &&
timeout 420 ncu --section SourceCounters --section InstructionStats --import-source on --clock-control none -k regex:'shadow_kernel' -s 3 -c 1 -o $O/prof_shadow_src $P > $O/ncu_src2.log 2>&1
echo "ncu rc=$?" >> $O/ncu_src2.log
ls -la $O
