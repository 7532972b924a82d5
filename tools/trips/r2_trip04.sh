#!/bin/bash
# Round 2, trip 4 (1 GPU): full GPU suite on the reworked code (80-register K1, auto hash rule, pipelined K3, NCCL C ABI at
# world 1), K1 sweep with one warp per query and other staging shapes at large L, K2 at 1.25M / 10M-row shards + launch list,
# and the new bench.py line (parity block, roofline_knn).
mkdir -p gpurun_out
O=gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > $O/r2t4_tests.log 2>&1; echo "tests exit $?"; tail -5 $O/r2t4_tests.log
( timeout 900 python tools/k1_sweep.py --Ls 55 100 200 500 --reps 6 --configs w=0 w=0 w=1 w=1,sr=16 w=2,sr=16 w=1,hs=3 w=2,hs=3 w=2,hs=2 w=3,hs=3 w=4,hs=3 --out $O/r2t4_k1_sweep.json ) > $O/r2t4_k1_sweep.txt 2>&1; grep '^{' $O/r2t4_k1_sweep.txt | cut -c1-175
for n in 1250000 10000000; do
  timeout 300 python tools/microbench_knn.py --n $n --nq 262144 --reps 3 2>&1 | grep '^{' | tee -a $O/r2t4_knn.txt
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2t4_launches_knn_1p25M.csv \
    python tools/microbench_knn.py --n 1250000 --nq 131072 --reps 1 > $O/r2t4_ncu_knn.log 2>&1; tail -c 300 $O/r2t4_ncu_knn.log
( time timeout 900 python bench.py ) > $O/r2t4_bench.txt 2>&1; tail -c 3000 $O/r2t4_bench.txt
( time timeout 900 python bench.py --impl reference --steps 3 --warmup 1 ) > $O/r2t4_bench_ref.txt 2>&1; tail -c 800 $O/r2t4_bench_ref.txt
