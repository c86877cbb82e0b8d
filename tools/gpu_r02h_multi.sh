#!/bin/bash
# Round-2 multi-GPU session: N = $1 GPUs of one box.  tests + smoke (b2pt_group_render), torchrun bench (strong scaling, one reduce per
# frame, frame-diff check), the single-process bench (b2pt_group_render).
N=${1:-2}
cd /root/repo
O=gpurun_out/r02h_${N}gpu; mkdir -p $O
if [ "$N" = 2 ]; then
  ( time timeout 600 python -m pytest tests/test_gpu_multi.py tests/test_statistical.py -m gpu -x -q ) > $O/pytest_multi.log 2>&1; echo "pytest rc=$?" >> $O/pytest_multi.log
fi
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/smoke.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR bench.py --gpus $N --steps 3 --warmup 2 > $O/bench_default_${N}gpu.json 2> $O/bench_default.err; echo "rc=$?" >> $O/bench_default.err
timeout 300 python bench.py --single-process --gpus $N --steps 3 --warmup 2 > $O/bench_single_process_${N}gpu.json 2> $O/bench_single.err; echo "rc=$?" >> $O/bench_single.err
if [ "$N" = 8 ]; then
  timeout 600 $TR bench.py --gpus $N --steps 2 --warmup 1 --ndir 4 --no-variants > $O/bench_nee4_${N}gpu.json 2>/dev/null
  timeout 600 $TR bench.py --gpus $N --steps 2 --warmup 1 --scene sweep:clear_rough_plastic --frame-spp 4096 --ndir 4 > $O/bench_c5_plastic_${N}gpu.json 2>/dev/null
fi
ls -la $O
