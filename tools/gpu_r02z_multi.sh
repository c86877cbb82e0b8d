#!/bin/bash
# Final multi-GPU check of the round on N = $1 GPUs: multi-GPU tests, smoke (b2pt_group_render), torchrun bench at N and N/2 (strong
# scaling of the 2048-spp frame), the single-process bench, the reference arm under torchrun (rank 0 only).
N=${1:-8}
cd /root/repo
O=gpurun_out/r02z_${N}gpu; mkdir -p $O
( time timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q ) > $O/pytest_multi.log 2>&1; echo "pytest rc=$?" >> $O/pytest_multi.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/smoke.log
for n in $N $((N/2)); do
  TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511"
  timeout 600 $TR bench.py --gpus $n --steps 4 --warmup 3 > $O/bench_default_${n}gpu.json 2> $O/bench_default_${n}.err; echo "rc=$?" >> $O/bench_default_${n}.err
done
timeout 300 python bench.py --single-process --gpus $N --steps 3 --warmup 2 > $O/bench_single_process_${N}gpu.json 2> $O/bench_single.err; echo "rc=$?" >> $O/bench_single.err
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512"
timeout 600 $TR bench.py --impl reference --gpus $N --steps 2 --warmup 1 > $O/bench_reference_arm_${N}gpu.json 2> $O/bench_ref.err; echo "rc=$?" >> $O/bench_ref.err
ls -la $O
