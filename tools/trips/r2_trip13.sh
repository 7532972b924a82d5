#!/bin/bash
# Round 2, trip 13 (1 GPU): bucketed visited set without atomics (hash_space 4/5, the new auto) - parity suite, then the K1 sweep
# against the atomicCAS tables (hs=3 / hs=2) on the same box; variant with L2 prefetch instead of register-held buckets.
mkdir -p gpurun_out
O=gpurun_out
( time timeout 1200 python -m pytest tests/test_search_gpu.py -x -q ) > $O/r2t13_tests.log 2>&1; echo "tests exit $?"; tail -15 $O/r2t13_tests.log
( timeout 1200 python tools/k1_sweep.py --Ls 55 100 200 500 --reps 6 --configs hs=3 hs=0 hs=5 hs=2 hs=0,sr=16 hs=0 --out $O/r2t13_k1_sweep.json ) > $O/r2t13_k1_sweep.txt 2>&1; grep '^{' $O/r2t13_k1_sweep.txt | cut -c1-175
( RG_B200_LIB=$PWD/mysteryann_b200/variants/nospecregs.so timeout 900 python tools/k1_sweep.py --Ls 55 100 200 500 --reps 6 --configs hs=0 ) > $O/r2t13_k1_sweep_nospecregs.txt 2>&1; grep '^{' $O/r2t13_k1_sweep_nospecregs.txt | cut -c1-175
