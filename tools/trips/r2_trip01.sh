#!/bin/bash
# Round 2, trip 1 (1 GPU): baseline of the round-1 K1 at large beam widths before any change - `ncu --set full` at
# L_pq = 200 and 500 on the 10M bench index (stall reasons, lts hit rate, DRAM bytes vs cmps x 800), plus timing lines.
mkdir -p gpurun_out
O=gpurun_out
nproc > $O/r2_host.txt; free -g >> $O/r2_host.txt; nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv >> $O/r2_host.txt
( time timeout 900 python bench.py --no-cpu-baseline ) > $O/r2t1_bench.txt 2>&1; tail -c 600 $O/r2t1_bench.txt
for L in 100 200 500; do
  ( timeout 300 python bench.py --L $L --steps 5 --warmup 3 --no-cpu-baseline ) 2>&1 | grep '^{' > $O/r2t1_bench_L$L.txt; head -c 400 $O/r2t1_bench_L$L.txt; echo
done
for L in 200 500; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:rg_search_kernel -s 4 -c 1 -o $O/r2t1_k1_L$L -f \
      python bench.py --L $L --steps 2 --warmup 3 --no-cpu-baseline > $O/r2t1_ncu_L$L.log 2>&1; tail -c 200 $O/r2t1_ncu_L$L.log
done
ls -la $O/r2t1*
