#!/bin/bash
# Round 2, trip 28 (4 GPUs): the multi-GPU tests (CLI grid layouts, rg_knn_exact_grid, sharded), bench.py at N=4 with the
# build-kNN slice timed as best of three in both layouts.
mkdir -p gpurun_out
O=gpurun_out
( time timeout 900 python -m pytest tests/test_cli_gpu.py tests/test_knn_gpu.py -x -q -k "two_gpus or grid or sharded or devices" ) > $O/r2t28_tests.log 2>&1; echo "tests exit $?"; tail -4 $O/r2t28_tests.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
( time timeout 900 $TR --master-port 29571 bench.py --gpus 4 ) > $O/r2t28_bench_4gpu.txt 2>&1; grep '^{' $O/r2t28_bench_4gpu.txt | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(json.dumps({k:d.get(k) for k in ('value','n_gpus','ms_per_step','e2e','roofline','roofline_knn')})[:3000])"
