#!/bin/bash
# Final round-1 validation and evidence refresh: GPU suite, smoke, both bench arms at the 10M default, and the C4-shape
# build kNN on one GPU (bench --train 10000000: 10M training queries x 10M base x 200, K=100).
mkdir -p gpurun_out
O=gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q ) > $O/gpu_tests.log 2>&1; echo "gpu tests exit $?"; tail -3 $O/gpu_tests.log
( time timeout 300 python __graft_entry__.py --smoke ) > $O/smoke.log 2>&1; tail -2 $O/smoke.log
( time timeout 900 python bench.py --impl reference --steps 3 --warmup 1 ) > $O/bench_ref.txt 2>&1; grep '^{' $O/bench_ref.txt | cut -c1-160
( time timeout 900 python bench.py ) > $O/bench.txt 2>&1; grep '^{' $O/bench.txt | cut -c1-200
( time timeout 1200 python bench.py --train 10000000 --no-cpu-baseline ) > $O/bench_train10m.txt 2>&1; grep '^{' $O/bench_train10m.txt | cut -c1-200
