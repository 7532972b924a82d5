#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
( time timeout 300 python -m pytest tests/test_search_gpu.py -x -q -m gpu ) 2>&1 | tail -30 > $O/pytest_k1v2.txt
tail -12 $O/pytest_k1v2.txt
grep -q " passed" $O/pytest_k1v2.txt && ! grep -q failed $O/pytest_k1v2.txt || exit 1
for NQ in 10000 100000; do
timeout 150 python tools/microbench_search.py --n 10000000 --nq $NQ --Ls 35 100 --configs 2:4:0:8:2,2:2:0:8:2,2:1:0:8:2,2:8:0:8:2,2:4:0:16:2,2:4:0:8:1,1:4:0:8:2,2:3:0:8:2 > $O/mb7_$NQ.txt 2>&1
echo "nq=$NQ"; grep gather $O/mb7_$NQ.txt || tail -5 $O/mb7_$NQ.txt
done
