#!/bin/bash
# Round-2 final validation on one B200: tests, smoke, the driver's bench line and reference arm, the other BASELINE configurations,
# the drop-in program's wall time, NaN samples against the reference, launch list.
cd /root/repo
O=gpurun_out/r02z; mkdir -p $O
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv -lms 500 > $O/clocks.csv &
SMI=$!
( time timeout 900 python -m pytest tests -m gpu -x -q ) > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/smoke.log
timeout 900 python bench.py --steps 4 --warmup 3 > $O/bench_default.json 2> $O/bench_default.err; echo "bench rc=$?" >> $O/bench_default.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference_arm.json 2>/dev/null
B="--steps 2 --warmup 1 --no-cpu-baseline --no-variants"
timeout 300 python bench.py $B --ndir 4 > $O/bench_nee4.json 2>/dev/null
timeout 300 python bench.py $B --no-dof > $O/bench_c3_nodof.json 2>/dev/null
timeout 300 python bench.py $B --quality high --gem > $O/bench_c4_gem.json 2>/dev/null
timeout 300 python bench.py --steps 2 --warmup 1 --no-variants --scene cornell --width 512 --height 512 --frame-spp 32 --ndir 4 > $O/bench_c1_cornell_512_spp32.json 2>/dev/null
timeout 300 python bench.py $B --scene cornell --width 512 --height 512 --frame-spp 2048 --ndir 4 > $O/bench_c1_cornell_512_spp2048.json 2>/dev/null
timeout 300 python bench.py $B --scene sweep:clear_rough_plastic --frame-spp 1024 --ndir 4 > $O/bench_c5_plastic.json 2>/dev/null
timeout 300 python bench.py $B --scene sweep:gold_conductor --frame-spp 1024 --ndir 4 > $O/bench_c5_gold.json 2>/dev/null
kill $SMI
timeout 300 bash tools/time_program.sh > $O/program_wall.txt 2>&1
timeout 300 python tools/find_nan.py 1024 32 > $O/find_nan.txt 2>&1
P="python bench.py --frame-spp 64 --steps 1 --warmup 1 --no-cpu-baseline --no-variants --queue 2097152"
$P > $O/plain.log 2>&1 &&
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 6000 -c 2500 --csv --log-file $O/launches_default.csv $P > $O/ncu_launches.log 2>&1
ls -la $O
