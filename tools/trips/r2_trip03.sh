#!/bin/bash
# Round 2, trip 3 (1 GPU): same-box A/B of K1 builds - round 1's kernel (r1), the reworked kernel at 64 / 80 / 128
# registers per thread (default, mb3, mb2) - over hash flavour and speculation, L = 55..500.
mkdir -p gpurun_out
O=gpurun_out
for lib in r1 default mb3 mb2; do
  if [ $lib == default ]; then export RG_B200_LIB=; else export RG_B200_LIB=$PWD/mysteryann_b200/variants/$lib.so; fi
  cfgs="w=2 w=2 w=2,hs=2 w=2,pf=2 w=2,hs=2,pf=2"
  if [ $lib == r1 ]; then cfgs="w=2 w=2 w=2,pf=2 w=4"; fi
  ( timeout 600 python tools/k1_sweep.py --Ls 55 100 200 500 --reps 8 --configs $cfgs ) 2>&1 | grep '^{' | sed "s/^{/{\"lib\": \"$lib\", /" | tee -a $O/r2t3_ab.txt | cut -c1-150
done
