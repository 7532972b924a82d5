#!/bin/bash
# Round 2, trip 9 (1 GPU): n/256-node build waves - build tests with the tightened margins, C1 quality table, and the 10M
# bench with a rebuilt index (build seconds, L_pq at recall 0.9, QPS), then both arms + the launch list + ncu of K1.
mkdir -p gpurun_out
O=gpurun_out
( time timeout 1500 python -m pytest tests/test_build_gpu.py -x -q -s ) > $O/r2t9_tests.log 2>&1; echo "tests exit $?"; tail -4 $O/r2t9_tests.log; grep "recall gpu-built" $O/r2t9_tests.log | head -20
( time timeout 1500 python tools/build_quality_c1.py --out $O/r2t9_build_quality_c1.txt ) > $O/r2t9_build_quality.log 2>&1; grep "^# " $O/r2t9_build_quality.log
( time timeout 900 python bench.py --impl reference --steps 3 --warmup 1 ) > $O/r2t9_bench_ref.txt 2>&1; grep '^{' $O/r2t9_bench_ref.txt | cut -c1-400
( time timeout 900 python bench.py ) > $O/r2t9_bench.txt 2>&1; grep '^{' $O/r2t9_bench.txt | cut -c1-3000
