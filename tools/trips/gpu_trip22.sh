#!/bin/bash
# N-GPU build kNN (C4 shape slice): base-sharded K2/K3 + NCCL all-to-all + K4, timed after a warm-up, checked against the
# unsharded kernels.  usage: gpurun --gpus N -- 'bash tools/trips/gpu_trip22.sh N'
mkdir -p gpurun_out
O=gpurun_out
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
for nq in 262144 1048576; do
  ( time timeout 600 $TR --master-port 29513 tools/bench_knn_sharded.py --rows 10000000 --queries $nq ) > $O/knn_sharded_${N}gpu_$nq.txt 2>&1; echo "knn exit $?"; grep '^{' $O/knn_sharded_${N}gpu_$nq.txt || tail -5 $O/knn_sharded_${N}gpu_$nq.txt
done
