#!/bin/bash
# Round 2, trip 21 (1 GPU): bucket table sizing (target load 0.4 / 0.5 / 0.6: smaller slabs fit L2 at larger L_pq), one fresh
# process each.
mkdir -p gpurun_out
O=gpurun_out
for load in 0.4 0.5 0.6 0.4; do
  echo "== target load $load"
  ( RG_K1_TARGET_LOAD=$load timeout 900 python tools/k1_sweep.py --Ls 200 300 500 150 --reps 6 --configs hs=0 ) 2>&1 | grep '^{' | cut -c1-250 | tee -a $O/r2t21_k1_load_$load.txt
done
