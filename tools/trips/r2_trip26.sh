#!/bin/bash
# Round 2, trip 26 (8 GPUs): full C4 (10M training queries x 10M base, K=100) in the canonical layout (8 base shards) and in
# the 2 x 4 grid of rg_knn_exact_grid; grid / sharded tests at world 2 and 4.
mkdir -p gpurun_out
O=gpurun_out
( time timeout 600 python -m pytest tests/test_knn_gpu.py -x -q -k "grid or sharded" ) > $O/r2t26_tests.log 2>&1; echo "tests exit $?"; tail -3 $O/r2t26_tests.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
for bs in 2 0; do
  ( timeout 600 $TR --master-port 2956$bs tools/bench_knn_sharded.py --rows 10000000 --queries 10000000 --base-shards $bs ) 2>&1 | grep '^{' | tee -a $O/r2t26_knn_8gpu_c4.txt
done
