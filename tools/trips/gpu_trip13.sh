#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
RG_KNN_TRACE=1 timeout 300 python tools/microbench_knn.py --n 1000000 --nq 32768 --reps 3 > $O/knn_tr_1M.log 2>&1; tail -60 $O/knn_tr_1M.log
RG_KNN_TRACE=1 timeout 300 python tools/microbench_knn.py --n 4000000 --nq 65536 --reps 2 > $O/knn_tr_4M.log 2>&1; tail -30 $O/knn_tr_4M.log
nvidia-smi --query-gpu=clocks.sm,power.draw,clocks_event_reasons.active --format=csv -lms 100 > $O/clocks_knn.csv &
SMI=$!
timeout 300 python tools/microbench_knn.py --n 10000000 --nq 131072 --reps 2 > $O/knn_mb_10M.log 2>&1; tail -2 $O/knn_mb_10M.log
kill $SMI
sort $O/clocks_knn.csv | uniq -c | sort -rn | head -8
