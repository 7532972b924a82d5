#!/bin/bash
# Round 2, trip 20 (1 GPU): exception list for exhausted probe windows (no big-table reruns for 1e-8 events), batch stealing
# (batch_mode 3), parity test of the build-time searches against the oracle; parity suite + sweep.
mkdir -p gpurun_out
O=gpurun_out
( time timeout 1500 python -m pytest tests/test_search_gpu.py tests/test_build_gpu.py -q ) > $O/r2t20_tests.log 2>&1; echo "tests exit $?"; tail -8 $O/r2t20_tests.log
( timeout 1500 python tools/k1_sweep.py --Ls 55 100 150 200 300 500 --reps 6 --configs hs=0 hs=0,bm=3 hs=0 hs=0,bm=3 --out $O/r2t20_k1_sweep.json ) > $O/r2t20_k1_sweep.txt 2>&1; grep '^{' $O/r2t20_k1_sweep.txt | cut -c1-250
