#!/bin/bash
# Round 2, trip 2 (1 GPU): first run of the reworked K1 (in-place merge, 16-bit quotient hash, next-hop speculation, gather
# ring) and K2 orchestration (optimistic thresholds, scratch cache, q_batch 131072): parity suites, K1 variant sweep on
# the 10M bench index, K2 throughput at 1.25M / 10M-row shards under both schedules.
mkdir -p gpurun_out
O=gpurun_out
( time timeout 1200 python -m pytest tests/test_search_gpu.py tests/test_knn_gpu.py -x -q ) > $O/r2t2_tests.log 2>&1; echo "tests exit $?"; tail -15 $O/r2t2_tests.log
( timeout 900 python tools/k1_sweep.py --Ls 55 100 200 500 --configs w=0 w=2 w=4 w=2,hs=2 w=4,hs=2 w=2,sb=2 w=4,sb=2 w=2,pf=2 w=2,hs=2,pf=2 w=3 --out $O/r2t2_k1_sweep.json ) > $O/r2t2_k1_sweep.txt 2>&1; grep '^{' $O/r2t2_k1_sweep.txt | cut -c1-200
for n in 1250000 10000000; do
  for opt in 1 0; do
    echo "== knn n=$n optimistic=$opt"
    RG_KNN_OPTIMISTIC=$opt timeout 300 python tools/microbench_knn.py --n $n --nq 262144 --reps 3 2>&1 | grep '^{' | tee -a $O/r2t2_knn.txt
  done
done
RG_KNN_OPTIMISTIC=1 RG_KNN_QBATCH=32768 timeout 300 python tools/microbench_knn.py --n 1250000 --nq 262144 --reps 3 2>&1 | grep '^{' | tee -a $O/r2t2_knn.txt
RG_KNN_TRACE=1 timeout 300 python tools/microbench_knn.py --n 1250000 --nq 262144 --reps 1 > $O/r2t2_knn_trace.txt 2>&1
