#!/bin/bash
# Round 2, trip 7 (2 GPUs): the NCCL paths - rg_knn_exact_sharded from threads (tests) and from torchrun ranks (bench tool,
# C ABI exchange vs torch all_to_all), compute_groundtruth --devices 2, and bench.py at N=2 with the sharded kNN slice.
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi topo -m > $O/r2t7_topo.txt 2>&1
( time timeout 900 python -m pytest tests/test_knn_gpu.py tests/test_cli_gpu.py -x -q -k "sharded or two_gpus" ) > $O/r2t7_tests.log 2>&1; echo "tests exit $?"; tail -5 $O/r2t7_tests.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
for ex in capi torch; do
  ( timeout 600 $TR --master-port 29511 tools/bench_knn_sharded.py --rows 10000000 --queries 524288 --exchange $ex ) 2>&1 | grep '^{' | tee -a $O/r2t7_knn_sharded_2gpu.txt
done
( timeout 600 $TR --master-port 29512 tools/bench_knn_sharded.py --rows 2500000 --queries 524288 --exchange capi ) 2>&1 | grep '^{' | tee -a $O/r2t7_knn_sharded_2gpu.txt
( time timeout 900 python bench.py --impl reference --gpus 2 --steps 3 --warmup 1 ) > $O/r2t7_bench_ref.txt 2>&1; tail -c 300 $O/r2t7_bench_ref.txt
( time timeout 900 $TR --master-port 29513 bench.py --gpus 2 ) > $O/r2t7_bench_2gpu.txt 2>&1; grep '^{' $O/r2t7_bench_2gpu.txt | tail -c 3500
