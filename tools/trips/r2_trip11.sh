#!/bin/bash
# Round 2, trip 11 (4 GPUs): sharded-kNN tests after the single-chunk fast path (both paths), bench.py at N=4 (same L_pq as the
# other N), the kNN slice tool at 4 GPUs.
mkdir -p gpurun_out
O=gpurun_out
( time timeout 900 python -m pytest tests/test_knn_gpu.py tests/test_cli_gpu.py -x -q -k "sharded or two_gpus" ) > $O/r2t11_tests.log 2>&1; echo "tests exit $?"; tail -4 $O/r2t11_tests.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
( time timeout 900 $TR --master-port 29531 bench.py --gpus 4 ) > $O/r2t11_bench_4gpu.txt 2>&1; grep '^{' $O/r2t11_bench_4gpu.txt | cut -c1-600
( timeout 600 $TR --master-port 29532 tools/bench_knn_sharded.py --rows 10000000 --queries 1048576 --exchange capi ) 2>&1 | grep '^{' | tee -a $O/r2t11_knn_4gpu.txt
( timeout 600 $TR --master-port 29533 tools/bench_knn_sharded.py --rows 10000000 --queries 262144 --exchange capi ) 2>&1 | grep '^{' | tee -a $O/r2t11_knn_4gpu.txt
