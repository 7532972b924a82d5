#!/bin/bash
# Round 2, trip 24 (1 GPU): early issue of the next hop's filter + first gather before the merge (adj_prefetch bit 2), shared
# memory shaved to keep 11 CTAs per SM at L_pq = 500: parity suite, sweep pf=3 vs pf=7.
mkdir -p gpurun_out
O=gpurun_out
( time timeout 1500 python -m pytest tests/test_search_gpu.py tests/test_build_gpu.py -q ) > $O/r2t24_tests.log 2>&1; echo "tests exit $?"; tail -8 $O/r2t24_tests.log
( timeout 1500 python tools/k1_sweep.py --Ls 55 100 150 200 300 500 --reps 6 --configs hs=0 hs=0,pf=7 hs=4,pf=7 hs=0 hs=0,pf=7 --out $O/r2t24_k1_sweep.json ) > $O/r2t24_k1_sweep.txt 2>&1; grep '^{' $O/r2t24_k1_sweep.txt | cut -c1-250
