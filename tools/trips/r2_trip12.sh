#!/bin/bash
# Round 2, trip 12 (1 GPU): BASELINE.json configs[4] for real - 100M x 200 base, graph built on the GPU from 5M training
# queries (exact kNN 5M x 100M, connectivity enhancement of 100M nodes), 100 000-query search batches.
mkdir -p gpurun_out
O=gpurun_out
( time timeout 2400 python bench.py --config C5 --no-cpu-baseline --steps 5 --warmup 3 ) > $O/r2t12_bench_c5.txt 2>&1; tail -c 4000 $O/r2t12_bench_c5.txt
nvidia-smi --query-gpu=memory.used,memory.total --format=csv
