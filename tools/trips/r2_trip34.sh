#!/bin/bash
# Round 2, trip 34 (1 GPU): C3 (2.5M x 512 unit rows, k = 10 and k = 100) through bench.py with the final K1: parity block against
# the compiled reference on all 10 000 queries at a second row length (32 packed steps per row, no 8-wide tail).
mkdir -p gpurun_out
O=gpurun_out
( time timeout 900 python bench.py --config C3 ) > $O/r2t34_bench_c3.txt 2>&1; grep '^{' $O/r2t34_bench_c3.txt | cut -c1-600
( time timeout 900 python bench.py --config C3k100 ) > $O/r2t34_bench_c3k100.txt 2>&1; grep '^{' $O/r2t34_bench_c3k100.txt | cut -c1-600
