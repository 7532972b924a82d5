#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
for N in 50000 2000000 10000000; do
timeout 600 python tools/microbench_search.py --n $N --Ls 35 --configs 2:0:0:8:0:2,2:0:0:8:1:1,1:0:0:8:1:1,3:0:0:8:1:1 > $O/mb5_$N.txt 2>&1
echo "n=$N"; grep gather $O/mb5_$N.txt
done
RG_X=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:rg_search_kernel -s 2 -c 1 \
    -o $O/k1_ghash -f python tools/microbench_search.py --n 10000000 --Ls 35 --configs 2:0:0:8:1:1 > $O/ncu_k1_ghash.log 2>&1
tail -3 $O/ncu_k1_ghash.log
