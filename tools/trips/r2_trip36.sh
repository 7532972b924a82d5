#!/bin/bash
# Round 2, trip 36 (2 GPUs): the torchrun path of the final bench (search sharded by queries, build-kNN slice base-sharded through
# rg_knn_exact_sharded) after the last K1 / prune changes.
mkdir -p gpurun_out
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
( time timeout 360 $TR --master-port 29581 bench.py --gpus 2 ) > $O/r2t36_bench_2gpu.txt 2>&1; grep '^{' $O/r2t36_bench_2gpu.txt | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(json.dumps({k:d.get(k) for k in ('value','n_gpus','ms_per_step','e2e','roofline','roofline_knn','clocks')})[:2500])"
tail -3 $O/r2t36_bench_2gpu.txt | cut -c1-200
