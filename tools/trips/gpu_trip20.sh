#!/bin/bash
# Full GPU suite (new K2s select, zero-copy host path), K2 block-growth sweep at 10M x 131072, ncu --set full of one small
# early GEMM block and one select launch (what bounds the ~0.44 ms floor of the early blocks?)
mkdir -p gpurun_out
O=gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q ) > $O/gpu_tests.log 2>&1; echo "gpu tests exit $?"; tail -5 $O/gpu_tests.log
for g in 2 3 4; do
  RG_KNN_GROWTH=$g timeout 300 python tools/microbench_knn.py --n 10000000 --nq 131072 --reps 2 2>&1 | tail -1
done | tee $O/knn_growth_10M.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_knn_10Mx32K_v3.csv \
    python tools/microbench_knn.py --n 10000000 --nq 32768 --reps 1 > $O/ncu_k2_list.log 2>&1; tail -c 300 $O/ncu_k2_list.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:knn_gemm_filter -s 14 -c 1 -o $O/k2_small_block -f \
    python tools/microbench_knn.py --n 1000000 --nq 32768 --reps 1 > $O/ncu_k2_small.log 2>&1; tail -c 200 $O/ncu_k2_small.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:knn_select -s 14 -c 1 -o $O/k2_select -f \
    python tools/microbench_knn.py --n 1000000 --nq 32768 --reps 1 > $O/ncu_k2_select.log 2>&1; tail -c 200 $O/ncu_k2_select.log
ls -la $O | head -30
