#!/bin/bash
# Round 2, trip 19 (1 GPU): bucket tables filled to 90 % before a query is handed to the big-table pass; parity suite, sweep
# (max cmps printed), C1 differential run against the compiled reference over its whole L_pq sweep (10 ... 2000).
mkdir -p gpurun_out
O=gpurun_out
( time timeout 1200 python -m pytest tests/test_search_gpu.py tests/test_build_gpu.py -x -q ) > $O/r2t19_tests.log 2>&1; echo "tests exit $?"; tail -5 $O/r2t19_tests.log
( timeout 1500 python tools/k1_sweep.py --Ls 55 100 150 200 300 500 --reps 6 --configs hs=0 hs=0,bm=2 hs=0,pf=1 hs=0 --out $O/r2t19_k1_sweep.json ) > $O/r2t19_k1_sweep.txt 2>&1; grep '^{' $O/r2t19_k1_sweep.txt | cut -c1-250
( time timeout 1500 python tools/c1_differential.py --out $O/r2t19_c1_differential.txt ) > $O/r2t19_c1.log 2>&1; tail -70 $O/r2t19_c1.log
