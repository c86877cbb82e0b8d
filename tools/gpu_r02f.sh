#!/bin/bash
# Round-2 session f: current build (exact quads, one lane per ray, 8 blocks/SM; culling + shared eval) — launch lists and counters.
cd /root/repo
O=gpurun_out/r02f; mkdir -p $O
( time timeout 900 python -m pytest tests -m gpu -x -q ) > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
B="--steps 3 --warmup 2 --no-cpu-baseline --no-variants"
timeout 300 python bench.py $B --frame-spp 256 > $O/bench_default_256.json 2>/dev/null
timeout 300 python bench.py $B --frame-spp 512 --ndir 4 > $O/bench_nee4_512.json 2>/dev/null
timeout 300 python bench.py $B --scene cornell --width 1024 --height 1024 --frame-spp 256 --ndir 4 > $O/bench_cornell_256.json 2>/dev/null
P="python bench.py --frame-spp 64 --steps 1 --warmup 1 --no-cpu-baseline --no-variants --queue 2097152"
C="python bench.py --scene cornell --width 1024 --height 1024 --ndir 4 --frame-spp 64 --steps 1 --warmup 1 --no-cpu-baseline --no-variants --queue 2097152"
$P > $O/plain.log 2>&1 &&
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 6000 -c 2500 --csv --log-file $O/launches_default.csv $P > $O/ncu_launches.log 2>&1
$C > $O/plain_c.log 2>&1 &&
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 2500 --csv --log-file $O/launches_cornell.csv $C > $O/ncu_launches_c.log 2>&1
SEC="--section ComputeWorkloadAnalysis --section LaunchStats --section MemoryWorkloadAnalysis --section MemoryWorkloadAnalysis_Tables --section Occupancy --section SchedulerStats --section SpeedOfLight --section WarpStateStats --section WorkloadDistribution --metrics smsp__thread_inst_executed_per_inst_executed.ratio,smsp__inst_executed.sum,smsp__inst_executed_op_global_ld.sum,smsp__inst_executed_op_local_ld.sum,smsp__inst_executed_op_local_st.sum,smsp__inst_executed_pipe_fp64.sum,smsp__inst_executed_pipe_fma.sum,smsp__inst_executed_pipe_alu.sum,smsp__inst_executed_pipe_lsu.sum,l1tex__t_bytes.sum,lts__t_bytes.sum,dram__bytes_read.sum,dram__bytes_write.sum"
$P > $O/plain2.log 2>&1 &&
timeout 420 ncu $SEC --clock-control none -k regex:'generate_kernel|extend_kernel|light_kernel|nee_kernel|shadow_kernel|lit_kernel|nee_eval_kernel|terminal_kernel' -s 800 -c 8 -o $O/prof_chain $P > $O/ncu_chain.log 2>&1
$P > $O/plain4.log 2>&1 &&
timeout 420 ncu $SEC --clock-control none -k regex:'shade_kernel' -s 800 -c 8 -o $O/prof_shade $P > $O/ncu_shade.log 2>&1
$C > $O/plain3.log 2>&1 &&
timeout 420 ncu $SEC --clock-control none -k regex:'extend_kernel|nee_kernel|shadow_kernel|nee_eval_kernel|light_kernel' -s 300 -c 5 -o $O/prof_cornell $C > $O/ncu_cornell.log 2>&1
ls -la $O
