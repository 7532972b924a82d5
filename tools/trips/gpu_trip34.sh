#!/bin/bash
# bench.py on N GPUs under torchrun, as the driver launches it (fills the N=4 row of the 1/2/4/8 table)
mkdir -p gpurun_out
O=gpurun_out
N=${1:-4}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
( time timeout 600 $TR --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 ) > $O/bench_${N}gpu.txt 2>&1; echo "bench exit $?"; grep '^{' $O/bench_${N}gpu.txt | cut -c1-300 || tail -5 $O/bench_${N}gpu.txt
