#!/bin/bash
cd /root/repo
O=gpurun_out/r02n; mkdir -p $O
P="python bench.py --frame-spp 64 --steps 1 --warmup 1 --no-cpu-baseline --no-variants --queue 2097152"
$P > $O/plain.log 2>&1 &&
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 1500 --csv --log-file $O/launches_chess_nee32_lists.csv $P > $O/ncu_launches.log 2>&1
P="python bench.py --scene cornell --width 1024 --height 1024 --ndir 4 --frame-spp 64 --steps 1 --warmup 1 --no-cpu-baseline --no-variants --queue 2097152"
B2PT_SHAFT_MIN_NDIR=4 $P > $O/plain2.log 2>&1 &&
B2PT_SHAFT_MIN_NDIR=4 timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 1500 -c 1000 --csv --log-file $O/launches_cornell_nee4_lists.csv $P > $O/ncu_launches2.log 2>&1
ls -la $O
