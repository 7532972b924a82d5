#!/bin/bash
# Round 2, trip 14 (1 GPU): ncu --set full of K1 with the bucketed visited set at L_pq = 100 / 500; warps / stage-rows sweep with it.
mkdir -p gpurun_out
O=gpurun_out
for L in 500 100; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:rg_search_kernel -s 4 -c 1 -o $O/r2t14_k1_L$L -f \
      python bench.py --L $L --steps 2 --warmup 3 --no-cpu-baseline --knn-slice 0 > $O/r2t14_ncu_L$L.log 2>&1; tail -c 200 $O/r2t14_ncu_L$L.log
done
( timeout 1200 python tools/k1_sweep.py --Ls 55 100 200 500 --reps 6 --configs hs=0 hs=0,w=1 hs=0,w=3 hs=0,w=4 hs=0,sr=12 hs=0,pf=1 hs=0,pf=0 hs=0,l2=1 --out $O/r2t14_k1_sweep.json ) > $O/r2t14_k1_sweep.txt 2>&1; grep '^{' $O/r2t14_k1_sweep.txt | cut -c1-175
