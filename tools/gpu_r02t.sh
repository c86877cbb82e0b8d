#!/bin/bash
# launch list of the final build (ncu, cold-cache and serialised: compare shares with bench.py's roofline.share_of_step)
cd /root/repo
O=gpurun_out/r02t; mkdir -p $O
P="python bench.py --frame-spp 64 --steps 1 --warmup 1 --no-cpu-baseline --no-variants --queue 2097152"
$P > $O/plain.log 2>&1 &&
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -s 6000 -c 2500 --csv --log-file $O/launches_default.csv $P > $O/ncu_launches.log 2>&1
echo "ncu rc=$?"
