#!/bin/bash
# Measurements that were planned but not run in round 1 (the GPU budget was spent); nothing here changes the product.
#   gpurun --gpus 8 --timeout 900 -- 'bash tools/trips/next_round_checks.sh 8'
#   gpurun          --timeout 900 -- 'bash tools/trips/next_round_checks.sh 1'
mkdir -p gpurun_out
O=gpurun_out
N=${1:-1}
if [ "$N" -gt 1 ]; then
  TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
  # build kNN, full C4 shape: one base shard per rank (config C4) vs base shards x query groups (sharded_knn.Grid)
  for B in 0 2 1; do
    ( timeout 300 $TR --master-port 2951$B tools/bench_knn_sharded.py --rows 10000000 --queries 10000000 --base-shards $B ) 2>&1 | grep '^{' | tee -a $O/knn_grid_${N}gpu.txt
  done
else
  # C3 with k = 100 (recall@100 >= 0.9 selects L_pq)
  ( timeout 600 python bench.py --n 2500000 --dim 512 --normalize --k 100 --no-cpu-baseline ) 2>&1 | grep '^{' | tee $O/bench_c3_k100.txt
  # three back-to-back kNN repetitions: are the 0.25-0.7 s outliers the per-call scratch cudaMalloc/cudaFree?  
  RG_KNN_TRACE=1 timeout 300 python tools/microbench_knn.py --n 10000000 --nq 131072 --reps 5 > $O/knn_trace.txt 2>&1; tail -1 $O/knn_trace.txt
fi
