#!/bin/bash
# Round-2 first GPU session: tests, smoke, bench (default + the other BASELINE configurations), launch list, ncu captures.
cd /root/repo
O=gpurun_out/r02a; mkdir -p $O
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv -lms 500 > $O/clocks.csv &
SMI=$!
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/smoke.log
timeout 600 python bench.py --steps 3 --warmup 2 > $O/bench_default.json 2> $O/bench_default.err; echo "bench rc=$?" >> $O/bench_default.err
B="--steps 2 --warmup 1 --no-cpu-baseline --no-variants"
timeout 300 python bench.py $B --ndir 4 --frame-spp 512 > $O/bench_nee4.json 2>/dev/null
timeout 300 python bench.py $B --no-dof --frame-spp 512 > $O/bench_c3_nodof.json 2>/dev/null
timeout 300 python bench.py $B --quality high --gem --frame-spp 256 > $O/bench_c4_gem.json 2>/dev/null
timeout 300 python bench.py $B --scene cornell --width 512 --height 512 --frame-spp 32 --ndir 4 > $O/bench_c1_cornell.json 2>/dev/null
timeout 300 python bench.py $B --scene cornell --width 512 --height 512 --frame-spp 2048 --ndir 4 > $O/bench_c1_cornell_2048.json 2>/dev/null
timeout 300 python bench.py $B --scene sweep:clear_rough_plastic --frame-spp 512 --ndir 4 > $O/bench_c5_plastic.json 2>/dev/null
timeout 300 python bench.py $B --scene sweep:gold_conductor --frame-spp 512 --ndir 4 > $O/bench_c5_gold.json 2>/dev/null
kill $SMI
# launch list + full captures of the traversal / NEE / shading kernels on the default workload (short frame, small queue)
P="python bench.py --frame-spp 64 --steps 1 --warmup 1 --no-cpu-baseline --no-variants --queue 2097152"
$P > $O/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 2500 --csv --log-file $O/launches.csv $P > $O/ncu_launches.log 2>&1
$P > $O/plain2.log 2>&1 &&
ncu --set full --clock-control none -k regex:'extend_kernel|light_kernel|nee_kernel|shadow_kernel|lit_kernel|nee_eval_kernel|terminal_kernel' -s 140 -c 7 -o $O/prof_chain $P > $O/ncu_chain.log 2>&1
$P > $O/plain3.log 2>&1 &&
ncu --set full --clock-control none -k regex:shade_kernel -s 160 -c 8 -o $O/prof_shade $P > $O/ncu_shade.log 2>&1
C="python bench.py --scene cornell --width 1024 --height 1024 --ndir 4 --frame-spp 64 --steps 1 --warmup 1 --no-cpu-baseline --no-variants --queue 2097152"
$C > $O/plain4.log 2>&1 &&
ncu --set full --clock-control none -k regex:'extend_kernel|shadow_kernel|nee_kernel' -s 60 -c 3 -o $O/prof_cornell $C > $O/ncu_cornell.log 2>&1
ls -la $O
