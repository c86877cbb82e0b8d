#!/bin/bash
# Round-2 session b: packet walks — tests, A/B against the per-ray walks, launch lists, counter captures (selected sections).
cd /root/repo
O=gpurun_out/r02b; mkdir -p $O
( time timeout 900 python -m pytest tests -m gpu -x -q ) > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
B="--steps 3 --warmup 2 --no-cpu-baseline --no-variants"
ab() { # name, bench args
  for pw in 1 0; do for ps in 1 0; do
    B2PT_PACKET_WALK=$pw B2PT_PACKET_SHADOW=$ps timeout 300 python bench.py $B $2 > $O/ab_$1_w${pw}s${ps}.json 2>/dev/null
  done; done
}
ab default "--frame-spp 256"
ab nee4 "--frame-spp 512 --ndir 4"
ab cornell "--scene cornell --width 1024 --height 1024 --frame-spp 256 --ndir 4"
ab gem "--quality high --gem --frame-spp 128"
P="python bench.py --frame-spp 64 --steps 1 --warmup 1 --no-cpu-baseline --no-variants --queue 2097152"
C="python bench.py --scene cornell --width 1024 --height 1024 --ndir 4 --frame-spp 64 --steps 1 --warmup 1 --no-cpu-baseline --no-variants --queue 2097152"
$P > $O/plain.log 2>&1 &&
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 2500 --csv --log-file $O/launches_default.csv $P > $O/ncu_launches.log 2>&1
$C > $O/plain_c.log 2>&1 &&
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 1500 -c 2500 --csv --log-file $O/launches_cornell.csv $C > $O/ncu_launches_c.log 2>&1
SEC="--section ComputeWorkloadAnalysis --section LaunchStats --section MemoryWorkloadAnalysis --section MemoryWorkloadAnalysis_Tables --section Occupancy --section SchedulerStats --section SpeedOfLight --section WarpStateStats --section WorkloadDistribution --metrics smsp__thread_inst_executed_per_inst_executed.ratio,smsp__inst_executed.sum,smsp__inst_executed_op_global_ld.sum,smsp__inst_executed_op_local_ld.sum,smsp__inst_executed_op_local_st.sum,smsp__inst_executed_pipe_fp64.sum,smsp__inst_executed_pipe_fma.sum,smsp__inst_executed_pipe_alu.sum,smsp__inst_executed_pipe_lsu.sum,l1tex__t_bytes.sum,lts__t_bytes.sum,dram__bytes_read.sum,dram__bytes_write.sum"
$P > $O/plain2.log 2>&1 &&
timeout 420 ncu $SEC --clock-control none -k regex:'extend_kernel|light_kernel|nee_kernel|shadow_kernel|nee_eval_kernel|terminal_kernel|generate_kernel' -s 140 -c 7 -o $O/prof_chain $P > $O/ncu_chain.log 2>&1
$C > $O/plain3.log 2>&1 &&
timeout 420 ncu $SEC --clock-control none -k regex:'extend_kernel|nee_kernel|shadow_kernel|nee_eval_kernel|shade_kernel' -s 120 -c 12 -o $O/prof_cornell $C > $O/ncu_cornell.log 2>&1
ls -la $O
