#!/bin/bash
# Validation of the automatic warps-per-query rule: GPU suite, CLI pipeline at 10M (full L sweep), both bench arms.
mkdir -p gpurun_out
O=gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q ) > $O/gpu_tests.log 2>&1; echo "gpu tests exit $?"; tail -3 $O/gpu_tests.log
( time timeout 900 python bench.py --impl reference --steps 3 --warmup 1 ) > $O/bench_ref.txt 2>&1; tail -c 300 $O/bench_ref.txt
( time timeout 900 python bench.py ) > $O/bench.txt 2>&1; grep '^{' $O/bench.txt | cut -c1-200
for L in 80 120 160 180; do ( timeout 600 python bench.py --L $L --no-cpu-baseline --steps 10 ) 2>&1 | grep '^{' | python -c "
import json,sys
j=json.loads(sys.stdin.read()); print('L',j['config']['L_pq'],'value',j['value'],'frac',j['roofline']['frac'])"; done | tee $O/k1_auto_warps.txt
timeout 900 python tools/cli_pipeline_10m.py --out $O/cli_pipeline_10m.txt > $O/cli_pipeline.log 2>&1; tail -45 $O/cli_pipeline_10m.txt | cut -c1-120
