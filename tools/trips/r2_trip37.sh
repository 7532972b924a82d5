#!/bin/bash
# Round 2, trip 37 (1 GPU): C3 k = 100 and ncu --set full of K1 at L_pq = 200 / 500 with the final kernel.
mkdir -p gpurun_out
O=gpurun_out
( time timeout 200 python bench.py --config C3k100 ) > $O/r2t37_bench_c3k100.txt 2>&1; grep '^{' $O/r2t37_bench_c3k100.txt | cut -c1-200
timeout 200 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --knn-slice 0 > $O/r2t37_prep.log 2>&1; tail -c 100 $O/r2t37_prep.log
NCU="ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:rg_search_kernelILb.ELi.ELi.ELb0 -s 4 -c 1 -f"
for L in 200 500; do
  timeout 150 $NCU -o $O/r2t37_k1_L$L python bench.py --L $L --steps 2 --warmup 3 --no-cpu-baseline --knn-slice 0 > $O/r2t37_ncu_L$L.log 2>&1; tail -c 100 $O/r2t37_ncu_L$L.log
done
