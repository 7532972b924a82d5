#!/bin/bash
# ncu --set full of one large K2 GEMM+filter block (staging-only epilogue) at 1M x 32768: tensor-pipe utilisation of the
# current kernel.  Launch 21 = the last (488K-row) block of the first timed repetition.
mkdir -p gpurun_out
O=gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:knn_gemm_filter -s 21 -c 1 -o $O/k2_large_block -f \
    python tools/microbench_knn.py --n 1000000 --nq 32768 --reps 1 > $O/ncu_k2_large.log 2>&1; tail -c 200 $O/ncu_k2_large.log
