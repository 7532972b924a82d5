#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
N=${1:-500000}
( time timeout 600 python bench.py --n $N --no-cpu-baseline --warps 2 ) > $O/b8_w2.txt 2>&1; grep -o '"value": [0-9.]*' $O/b8_w2.txt | head -2
for W in 1 4 3; do
timeout 300 python bench.py --n $N --L 35 --no-cpu-baseline --warps $W > $O/b8_w$W.txt 2>&1; echo "W=$W"; grep -o '"value": [0-9.]*' $O/b8_w$W.txt | head -2
done
timeout 300 python bench.py --n $N --L 35 --no-cpu-baseline --warps 2 --hash-space 1 > $O/b8_w2s.txt 2>&1; echo "W=2 smem hash"; grep -o '"value": [0-9.]*' $O/b8_w2s.txt | head -2
timeout 300 python bench.py --n $N --L 35 --no-cpu-baseline --warps 2 --queries 100000 > $O/b8_w2_100k.txt 2>&1; echo "W=2 100K queries"; grep -o '"value": [0-9.]*' $O/b8_w2_100k.txt | head -2
