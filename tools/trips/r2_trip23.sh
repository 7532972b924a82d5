#!/bin/bash
# Round 2, trip 23 (1 GPU): ncu --set full of the search kernel at L_pq = 500 / 200 / 55 with the current K1 (bench workload).
mkdir -p gpurun_out
O=gpurun_out
( timeout 600 python bench.py --no-cpu-baseline --knn-slice 0 --steps 5 ) > $O/r2t23_warm.txt 2>&1; tail -c 300 $O/r2t23_warm.txt
for L in 500 200 55; do
  timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:rg_search_kernelILb.ELi.ELi.ELb0 -s 4 -c 1 -o $O/r2t23_k1_L$L -f \
      python bench.py --L $L --steps 2 --warmup 3 --no-cpu-baseline --knn-slice 0 > $O/r2t23_ncu_L$L.log 2>&1; tail -c 150 $O/r2t23_ncu_L$L.log
done
