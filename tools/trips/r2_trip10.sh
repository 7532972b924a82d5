#!/bin/bash
# Round 2, trip 10 (8 GPUs): the full C4 shape (10M training queries x 10M x 200 base, K=100) base-sharded over 8 GPUs through
# rg_knn_exact_sharded, then bench.py at N=8 (sharded build kNN for the index, sharded kNN slice, search) and the reference arm.
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi topo -m > $O/r2t10_topo.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
( timeout 600 $TR --master-port 29521 tools/bench_knn_sharded.py --rows 10000000 --queries 10000000 --exchange capi ) 2>&1 | grep '^{' | tee -a $O/r2t10_knn_c4_8gpu.txt
( time timeout 900 $TR --master-port 29522 bench.py --gpus 8 ) > $O/r2t10_bench_8gpu.txt 2>&1; grep '^{' $O/r2t10_bench_8gpu.txt | cut -c1-3500
( time timeout 600 python bench.py --impl reference --gpus 8 --steps 3 --warmup 1 ) > $O/r2t10_bench_ref.txt 2>&1; grep '^{' $O/r2t10_bench_ref.txt | cut -c1-300
