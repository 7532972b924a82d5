#!/bin/bash
# K2 pair kernel (cta_group::2): parity, throughput probe, ncu capture of the largest GEMM launch
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_knn_gpu.py -x -q > $O/knn_tests.log 2>&1; echo "knn tests exit $?"; tail -5 $O/knn_tests.log
timeout 300 python tools/microbench_knn.py --n 1000000 --nq 32768 > $O/knn_mb_1M.log 2>&1; tail -2 $O/knn_mb_1M.log
timeout 300 python tools/microbench_knn.py --n 4000000 --nq 65536 --reps 2 > $O/knn_mb_4M.log 2>&1; tail -2 $O/knn_mb_4M.log
timeout 300 python tools/microbench_knn.py --n 10000000 --nq 131072 --reps 1 > $O/knn_mb_10M.log 2>&1; tail -2 $O/knn_mb_10M.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:knn_gemm_filter -s 10 -c 1 \
    -o $O/r01_k2_pair_1M -f python tools/microbench_knn.py --n 1000000 --nq 32768 --reps 1 > $O/ncu_k2.log 2>&1
tail -2 $O/ncu_k2.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $O/r01_launches_knn_pair_1Mx32K.csv \
    python tools/microbench_knn.py --n 1000000 --nq 32768 --reps 1 > $O/ncu_k2_list.log 2>&1
ls -la $O | head
