#!/bin/bash
# Round 2, trip 16 (1 GPU): packed shared-memory layout (+1 CTA per SM at several L_pq), merge with four entries per thread,
# one match.any per filter round; state effects of the persisting-L2 set-aside (order of L, slack).
mkdir -p gpurun_out
O=gpurun_out
( time timeout 1200 python -m pytest tests/test_search_gpu.py tests/test_build_gpu.py -x -q ) > $O/r2t16_tests.log 2>&1; echo "tests exit $?"; tail -5 $O/r2t16_tests.log
( timeout 1500 python tools/k1_sweep.py --Ls 55 100 200 500 --reps 6 --configs hs=0 hs=4 hs=4 hs=3 hs=2 hs=4,bm=2 hs=4,sr=12 hs=4,w=3 hs=4,sr=4,w=4 --out $O/r2t16_k1_sweep.json ) > $O/r2t16_k1_sweep.txt 2>&1; grep '^{' $O/r2t16_k1_sweep.txt | cut -c1-175
echo "== fresh process, large L first"
( timeout 900 python tools/k1_sweep.py --Ls 500 200 100 55 100 200 --reps 6 --configs hs=4 ) 2>&1 | grep '^{' | cut -c1-175 | tee $O/r2t16_k1_order.txt
echo "== fresh process, 25 % slack on the set-aside"
( RG_K1_PERSIST_SLACK_PCT=25 timeout 900 python tools/k1_sweep.py --Ls 55 100 200 500 --reps 6 --configs hs=0 hs=4 ) 2>&1 | grep '^{' | cut -c1-175 | tee $O/r2t16_k1_slack.txt
