#!/bin/bash
# K1 variants on the real 10M index: L2 eviction hints / persisting hash window / speculative adjacency prefetch
mkdir -p gpurun_out
O=gpurun_out
( timeout 600 python -m pytest tests/test_search_gpu.py -x -q ) > $O/k1_tests.log 2>&1; echo "k1 tests exit $?"; tail -4 $O/k1_tests.log
timeout 900 python bench.py --L 60 --no-cpu-baseline --steps 5 > $O/b15_build.txt 2>&1; tail -c 300 $O/b15_build.txt
for v in 0:0 1:0 2:0 3:0 0:1 0:2 0:3 1:3 3:3; do
  l2=${v%%:*}; pf=${v##*:}
  timeout 300 python bench.py --L 60 --no-cpu-baseline --l2-hint $l2 --adj-prefetch $pf > $O/b15_$l2_$pf.txt 2>&1
  echo "l2=$l2 pf=$pf $(grep -o '"value": [0-9.]*' $O/b15_$l2_$pf.txt | head -2 | tr '\n' ' ') $(grep -o '"frac": [0-9.]*' $O/b15_$l2_$pf.txt)"
done | tee $O/k1_variants_10m.txt
for v in 0:0 3:3; do
  l2=${v%%:*}; pf=${v##*:}
  timeout 300 python bench.py --L 60 --no-cpu-baseline --l2-hint $l2 --adj-prefetch $pf --queries 100000 > $O/b15q_$l2_$pf.txt 2>&1
  echo "100K queries l2=$l2 pf=$pf $(grep -o '"value": [0-9.]*' $O/b15q_$l2_$pf.txt | head -2 | tr '\n' ' ') $(grep -o '"frac": [0-9.]*' $O/b15q_$l2_$pf.txt)"
done | tee -a $O/k1_variants_10m.txt
