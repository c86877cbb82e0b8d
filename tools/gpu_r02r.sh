#!/bin/bash
# refill threshold 0 (a warp finishes its 32 consecutive rays, then takes the next 32) against the former 12 on the other configurations; GPU tests on the new default
cd /root/repo
O=gpurun_out/r02r; mkdir -p $O
( timeout 900 python -m pytest tests -m gpu -x -q ) > $O/pytest_gpu.log 2>&1; tail -2 $O/pytest_gpu.log
SPP=512 BENCH_ARGS="--ndir 4" bash tools/ab.sh variants/base.so variants/refill0.so 2>&1 | tee $O/ab_refill0_nee4.txt
SPP=512 BENCH_ARGS="--quality high --gem" bash tools/ab.sh variants/base.so variants/refill0.so 2>&1 | tee $O/ab_refill0_gem.txt
SPP=512 BENCH_ARGS="--no-dof" bash tools/ab.sh variants/base.so variants/refill0.so 2>&1 | tee $O/ab_refill0_nodof.txt
