#!/bin/bash
# Round 2, trip 22 (4 GPUs): rg_knn_exact_grid (base shards x query groups over one communicator): tests at world 2 and 4,
# the kNN slice tool in the layouts 4x1 (canonical), 2x2 and 1x4, bench.py at N=4.
mkdir -p gpurun_out
O=gpurun_out
( time timeout 900 python -m pytest tests/test_knn_gpu.py -x -q -k "grid or sharded" ) > $O/r2t22_tests.log 2>&1; echo "tests exit $?"; tail -4 $O/r2t22_tests.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
for bs in 0 2 1; do
  ( timeout 600 $TR --master-port 2954$bs tools/bench_knn_sharded.py --rows 10000000 --queries 1048576 --base-shards $bs ) 2>&1 | grep '^{' | tee -a $O/r2t22_knn_4gpu.txt
done
( time timeout 900 $TR --master-port 29551 bench.py --gpus 4 ) > $O/r2t22_bench_4gpu.txt 2>&1; grep '^{' $O/r2t22_bench_4gpu.txt | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(json.dumps({k:d.get(k) for k in ('value','n_gpus','ms_per_step','e2e','roofline','roofline_knn')})[:2500])"
