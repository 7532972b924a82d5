#!/bin/bash
# Round 2, trip 27 (1 GPU): BASELINE.json configs[4] with the final K1 - 100M x 200 base, graph built on the GPU from 5M
# training queries, 100 000-query search batches (bucketed visited set with 32-bit ids at this id range).
mkdir -p gpurun_out
O=gpurun_out
( time timeout 2400 python bench.py --config C5 --no-cpu-baseline --steps 5 --warmup 3 ) > $O/r2t27_bench_c5.txt 2>&1; tail -c 4500 $O/r2t27_bench_c5.txt
