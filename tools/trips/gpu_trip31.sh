#!/bin/bash
# K2 with the register-tied tcgen05.wait::ld: kNN parity tests and throughput at 10M x 131072 (must match 489 ms)
mkdir -p gpurun_out
O=gpurun_out
( timeout 600 python -m pytest tests/test_knn_gpu.py tests/test_cli_gpu.py tests/test_build_gpu.py -q ) > $O/knn_tests.log 2>&1; echo "tests exit $?"; tail -3 $O/knn_tests.log
timeout 300 python tools/microbench_knn.py --n 10000000 --nq 131072 --reps 3 2>&1 | tail -1 | tee $O/knn_mb_10M_waitdep.txt
