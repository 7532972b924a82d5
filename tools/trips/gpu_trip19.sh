#!/bin/bash
# Two-GPU validation (gpurun --gpus 2): bench.py under torchrun exactly as the driver launches it (fresh box: the index
# build runs the base-sharded NCCL kNN), the reference arm under torchrun (rank 0 works, rank 1 exits 0), and the
# sharded build-kNN timing + self-check against the unsharded kernels.
mkdir -p gpurun_out
O=gpurun_out
N=${1:-2}
nvidia-smi --query-gpu=index,name --format=csv > $O/host_${N}gpu.txt; nvidia-smi topo -m >> $O/host_${N}gpu.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
( time timeout 1200 $TR --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 ) > $O/bench_${N}gpu.txt 2>&1; echo "bench exit $?"; grep '^{' $O/bench_${N}gpu.txt | tail -c 2500
( time timeout 900 $TR --master-port 29512 bench.py --impl reference --gpus $N --steps 3 --warmup 1 ) > $O/bench_ref_${N}gpu.txt 2>&1; echo "ref exit $?"; grep '^{' $O/bench_ref_${N}gpu.txt | tail -c 600
( time timeout 600 $TR --master-port 29513 tools/bench_knn_sharded.py --rows 10000000 --queries 262144 ) > $O/knn_sharded_${N}gpu.txt 2>&1; echo "knn exit $?"; grep '^{' $O/knn_sharded_${N}gpu.txt
( time timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline ) > $O/bench_1gpu_samebox.txt 2>&1; grep '^{' $O/bench_1gpu_samebox.txt | tail -c 900
tail -5 $O/bench_${N}gpu.txt
