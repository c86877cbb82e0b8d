#!/bin/bash
# nee_kernel reads the vertex's normal / wo / material from a record written by light_kernel instead of redoing hit_point per sample
cd /root/repo
O=gpurun_out/r02q; mkdir -p $O
( timeout 900 python -m pytest tests -m gpu -x -q ) > $O/pytest_gpu.log 2>&1; tail -2 $O/pytest_gpu.log
SPP=512 bash tools/ab.sh variants/base.so variants/vtxgeo.so variants/base.so variants/vtxgeo.so 2>&1 | tee $O/ab_chess_nee32.txt
SPP=512 BENCH_ARGS="--ndir 4" bash tools/ab.sh variants/base.so variants/vtxgeo.so 2>&1 | tee $O/ab_chess_nee4.txt
SPP=256 BENCH_ARGS="--scene cornell --width 1024 --height 1024 --ndir 4" bash tools/ab.sh variants/base.so variants/vtxgeo.so 2>&1 | tee $O/ab_cornell.txt
