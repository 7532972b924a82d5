#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
for NQ in 10000 100000; do
timeout 600 python tools/microbench_search.py --n 10000000 --nq $NQ --Ls 35 --configs 2:0:0:8:0:2,2:0:0:8:1:1,2:4:4:8:1:1,2:4:3:8:1:1,2:4:2:8:1:1,2:8:4:8:1:1 > $O/mb6_$NQ.txt 2>&1
echo "nq=$NQ"; grep gather $O/mb6_$NQ.txt
done
