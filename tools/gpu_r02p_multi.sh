#!/bin/bash
# 8 GPUs after the default-queue change: torchrun bench (strong scaling of the 2048-spp frame) and the single-process bench
N=${1:-8}
cd /root/repo
O=gpurun_out/r02p_${N}gpu; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 400 $TR bench.py --gpus $N --steps 4 --warmup 3 > $O/bench_default_${N}gpu.json 2> $O/bench_default_${N}.err; echo "rc=$?" >> $O/bench_default_${N}.err
timeout 300 python bench.py --single-process --gpus $N --steps 3 --warmup 2 > $O/bench_single_process_${N}gpu.json 2> $O/bench_single.err; echo "rc=$?" >> $O/bench_single.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02p_*gpu/bench_*.json')):
    l=json.loads(open(f).readline()); print(f, l['n_gpus'], round(l['ms_per_step'],1), round(l['spp_per_s']/1e6,1), l.get('multi_gpu_max_abs_diff'))
PY
