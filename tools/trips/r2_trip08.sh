#!/bin/bash
# Round 2, trip 8 (1 GPU): packed-CAS visited hash - parity suite + sweep; graph-build wave-size sensitivity at C1;
# C3 with k=10 and k=100 through bench.py.
mkdir -p gpurun_out
O=gpurun_out
( time timeout 1500 python -m pytest tests/test_search_gpu.py tests/test_build_gpu.py -x -q ) > $O/r2t8_tests.log 2>&1; echo "tests exit $?"; tail -4 $O/r2t8_tests.log
( timeout 900 python tools/k1_sweep.py --Ls 55 100 200 500 --reps 8 --configs w=0 w=0 hs=2 hs=3 hs=3,pf=2 w=0 --out $O/r2t8_k1_sweep.json ) > $O/r2t8_k1_sweep.txt 2>&1; grep '^{' $O/r2t8_k1_sweep.txt | cut -c1-175
for wave in 512 1024 16384; do
  echo "== wave $wave"; RG_BUILD_WAVE=$wave timeout 900 python tools/build_quality_c1.py --l2-n 0 --out $O/r2t8_build_quality_wave$wave.txt > $O/r2t8_bq_$wave.log 2>&1; tail -1 $O/r2t8_bq_$wave.log; grep "GPU build:" $O/r2t8_bq_$wave.log
done
( time timeout 900 python bench.py --config C3 ) > $O/r2t8_bench_c3.txt 2>&1; grep '^{' $O/r2t8_bench_c3.txt | cut -c1-1500
( time timeout 900 python bench.py --config C3k100 ) > $O/r2t8_bench_c3k100.txt 2>&1; grep '^{' $O/r2t8_bench_c3k100.txt | cut -c1-1500
