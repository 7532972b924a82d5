#!/bin/bash
# 8-GPU evidence: the full C4 build kNN (10M train x 10M base x 200, K=100) base-sharded over N GPUs, then bench.py under
# torchrun as the driver launches it.  usage: gpurun --gpus 8 -- 'bash tools/trips/gpu_trip23.sh 8'
mkdir -p gpurun_out
O=gpurun_out
N=${1:-8}
nvidia-smi topo -m > $O/host_${N}gpu.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
( time timeout 420 $TR --master-port 29513 tools/bench_knn_sharded.py --rows 10000000 --queries 10000000 ) > $O/knn_sharded_${N}gpu_c4.txt 2>&1; echo "knn exit $?"; grep '^{' $O/knn_sharded_${N}gpu_c4.txt || tail -5 $O/knn_sharded_${N}gpu_c4.txt
( time timeout 600 $TR --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 ) > $O/bench_${N}gpu.txt 2>&1; echo "bench exit $?"; grep '^{' $O/bench_${N}gpu.txt | cut -c1-400 || tail -5 $O/bench_${N}gpu.txt
tail -4 $O/bench_${N}gpu.txt | cut -c1-300
