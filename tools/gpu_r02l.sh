#!/bin/bash
# A/B: prefetch of the pushed quads into L1, resident blocks per SM of the extend kernel
cd /root/repo
O=gpurun_out/r02l; mkdir -p $O
SPP=256 bash tools/ab.sh variants/base.so variants/pf.so variants/ext6.so variants/ext7.so variants/ext10.so variants/pf_ext6.so variants/base.so > $O/ab2_chess_nee32.txt 2>&1
SPP=256 BENCH_ARGS="--ndir 4" bash tools/ab.sh variants/base.so variants/pf.so variants/ext6.so variants/pf_ext6.so > $O/ab2_chess_nee4.txt 2>&1
SPP=256 BENCH_ARGS="--quality high" bash tools/ab.sh variants/base.so variants/pf.so variants/ext6.so > $O/ab2_chess_high.txt 2>&1
cat $O/ab2_*.txt
