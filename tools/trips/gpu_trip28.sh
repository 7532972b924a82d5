#!/bin/bash
# K2 epilogue variants per block: dense-hit handling (bulk reservation) only for blocks that start before RG_KNN_DENSE_ROWS
# rows, sparse staging for the rest.  kNN parity tests, throughput at 10M x 131072 for several switch points, launch list.
mkdir -p gpurun_out
O=gpurun_out
( timeout 600 python -m pytest tests/test_knn_gpu.py tests/test_cli_gpu.py tests/test_build_gpu.py -q ) > $O/knn_tests.log 2>&1; echo "tests exit $?"; tail -3 $O/knn_tests.log
for d in 4096 0 1099511627776 16384 2048 4096; do
  echo -n "dense_rows=$d "; RG_KNN_DENSE_ROWS=$d timeout 300 python tools/microbench_knn.py --n 10000000 --nq 131072 --reps 2 2>&1 | tail -1
done | tee $O/knn_dense_rows_10M.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_knn_10Mx32K_v4.csv \
    python tools/microbench_knn.py --n 10000000 --nq 32768 --reps 1 > $O/ncu_k2_list.log 2>&1; tail -c 200 $O/ncu_k2_list.log
