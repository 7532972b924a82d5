#!/bin/bash
# Re-entry evidence: full GPU test suite, K2 pair-kernel ncu capture + launch list, K2 throughput at 10M, bench both arms.
mkdir -p gpurun_out
O=gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > $O/gpu_tests.log 2>&1; echo "gpu tests exit $?"; tail -6 $O/gpu_tests.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:knn_gemm_filter -s 10 -c 1 \
    -o $O/r01_k2_pair_1M -f python tools/microbench_knn.py --n 1000000 --nq 32768 --reps 1 > $O/ncu_k2.log 2>&1
tail -2 $O/ncu_k2.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file $O/r01_launches_knn_pair_1Mx32K.csv \
    python tools/microbench_knn.py --n 1000000 --nq 32768 --reps 1 > $O/ncu_k2_list.log 2>&1
timeout 300 python tools/microbench_knn.py --n 10000000 --nq 131072 --reps 2 > $O/knn_mb_10M.log 2>&1; tail -2 $O/knn_mb_10M.log
( time timeout 900 python -X faulthandler bench.py --impl reference --steps 3 --warmup 1 ) > $O/r01_bench_ref_10m.txt 2>&1; tail -c 600 $O/r01_bench_ref_10m.txt
( time timeout 1500 python -X faulthandler bench.py ) > $O/r01_bench_10m.txt 2>&1; tail -c 1500 $O/r01_bench_10m.txt
ls -la $O | head -40
