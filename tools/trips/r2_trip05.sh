#!/bin/bash
# Round 2, trip 5 (1 GPU): K1 register-cap A/B after removing the gather ring (64 / 72 / 80 registers vs round 1's kernel),
# search + build parity suites, K2 with the rank ladder (256-row first block) at 1.25M / 10M-row shards.
mkdir -p gpurun_out
O=gpurun_out
( time timeout 1500 python -m pytest tests/test_search_gpu.py tests/test_knn_gpu.py tests/test_build_gpu.py tests/test_cli_gpu.py -x -q ) > $O/r2t5_tests.log 2>&1; echo "tests exit $?"; tail -5 $O/r2t5_tests.log
for lib in r1 default r72 r64 r1 default; do
  if [ $lib == default ]; then export RG_B200_LIB=; else export RG_B200_LIB=$PWD/mysteryann_b200/variants/$lib.so; fi
  cfgs="w=2 w=0 w=0"
  if [ $lib == r1 ]; then cfgs="w=2 w=2 w=0"; fi
  ( timeout 600 python tools/k1_sweep.py --Ls 55 100 200 500 --reps 8 --configs $cfgs ) 2>&1 | grep '^{' | sed "s/^{/{\"lib\": \"$lib\", /" | tee -a $O/r2t5_ab.txt | cut -c1-150
done
export RG_B200_LIB=
for n in 1250000 10000000; do
  timeout 300 python tools/microbench_knn.py --n $n --nq 262144 --reps 3 2>&1 | grep '^{' | tee -a $O/r2t5_knn.txt
done
RG_KNN_RMIN=16 timeout 300 python tools/microbench_knn.py --n 1250000 --nq 262144 --reps 3 2>&1 | grep '^{' | tee -a $O/r2t5_knn.txt
RG_KNN_FIRST_ROWS=1024 timeout 300 python tools/microbench_knn.py --n 1250000 --nq 262144 --reps 3 2>&1 | grep '^{' | tee -a $O/r2t5_knn.txt
RG_KNN_GROWTH=8 timeout 300 python tools/microbench_knn.py --n 1250000 --nq 262144 --reps 3 2>&1 | grep '^{' | tee -a $O/r2t5_knn.txt
RG_KNN_GROWTH=3 timeout 300 python tools/microbench_knn.py --n 1250000 --nq 262144 --reps 3 2>&1 | grep '^{' | tee -a $O/r2t5_knn.txt
