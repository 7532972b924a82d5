#!/bin/bash
# Round 2, trip 15 (1 GPU): CTA-wide gather list with dynamic batches (batch_mode 2), persisting-L2 set-aside given back
# when no window is used, partial pinning (l2_hint 4): parity suite, sweep, ncu --set full of the search kernel at L_pq = 500 / 200.
mkdir -p gpurun_out
O=gpurun_out
( time timeout 1200 python -m pytest tests/test_search_gpu.py tests/test_build_gpu.py -x -q ) > $O/r2t15_tests.log 2>&1; echo "tests exit $?"; tail -5 $O/r2t15_tests.log
( timeout 1500 python tools/k1_sweep.py --Ls 55 100 200 500 --reps 6 --configs hs=0 hs=4 hs=4,bm=1 hs=4,w=3 hs=4,w=4 hs=3,bm=2 hs=3 hs=4,l2=7 hs=4,sr=12 hs=0 --out $O/r2t15_k1_sweep.json ) > $O/r2t15_k1_sweep.txt 2>&1; grep '^{' $O/r2t15_k1_sweep.txt | cut -c1-175
for L in 500 200; do
  timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:rg_search_kernelILb.ELi.ELi.ELb0 -s 4 -c 1 -o $O/r2t15_k1_L$L -f \
      python bench.py --L $L --steps 2 --warmup 3 --no-cpu-baseline --knn-slice 0 > $O/r2t15_ncu_L$L.log 2>&1; tail -c 200 $O/r2t15_ncu_L$L.log
done
