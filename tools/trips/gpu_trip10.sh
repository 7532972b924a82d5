#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
N=${1:-10000000}
( time timeout 1500 python -X faulthandler bench.py --n $N ) > $O/bench10m_a.txt 2>&1; tail -c 2500 $O/bench10m_a.txt
for W in 1 4; do
timeout 300 python bench.py --n $N --L 60 --no-cpu-baseline --warps $W > $O/b10_w$W.txt 2>&1; echo "W=$W"; grep -o '"value": [0-9.]*' $O/b10_w$W.txt | head -2
done
timeout 300 python bench.py --n $N --L 60 --no-cpu-baseline --queries 100000 > $O/b10_100k.txt 2>&1; echo "100K queries"; grep -o '"value": [0-9.]*' $O/b10_100k.txt | head -2
