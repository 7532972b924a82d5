#!/bin/bash
# Round 2, trip 6 (1 GPU): ncu --set full of the reworked K1 at L_pq = 200 / 500 (stall split after the rework), and the
# C1 graph-build quality comparison against two multi-threaded builds of the compiled reference.
mkdir -p gpurun_out
O=gpurun_out
( time timeout 900 python bench.py --no-cpu-baseline --knn-slice 0 ) > $O/r2t6_bench.txt 2>&1; tail -c 300 $O/r2t6_bench.txt
for L in 200 500; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:rg_search_kernel -s 4 -c 1 -o $O/r2t6_k1_L$L -f \
      python bench.py --L $L --steps 2 --warmup 3 --no-cpu-baseline --knn-slice 0 > $O/r2t6_ncu_L$L.log 2>&1; tail -c 200 $O/r2t6_ncu_L$L.log
done
( time timeout 1500 python tools/build_quality_c1.py --out $O/r2t6_build_quality_c1.txt ) > $O/r2t6_build_quality.log 2>&1; tail -60 $O/r2t6_build_quality.log
