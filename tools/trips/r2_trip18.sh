#!/bin/bash
# Round 2, trip 18 (1 GPU): bucket slabs of any size (no power-of-two rounding: 35 KB instead of 64 KB per query at L_pq = 200,
# so the slabs of all resident queries fit L2): parity suite, sweep, ncu at L_pq = 200.
mkdir -p gpurun_out
O=gpurun_out
python - <<'PY'
from cuda import cudart
for name in ("cudaDevAttrL2CacheSize", "cudaDevAttrMaxPersistingL2CacheSize", "cudaDevAttrMaxAccessPolicyWindowSize"):
    print(name, cudart.cudaDeviceGetAttribute(getattr(cudart.cudaDeviceAttr, name), 0))
PY
( time timeout 1200 python -m pytest tests/test_search_gpu.py tests/test_build_gpu.py -x -q ) > $O/r2t18_tests.log 2>&1; echo "tests exit $?"; tail -5 $O/r2t18_tests.log
( timeout 1500 python tools/k1_sweep.py --Ls 55 100 150 200 300 500 --reps 6 --configs hs=0 hs=4 hs=3 hs=0 --out $O/r2t18_k1_sweep.json ) > $O/r2t18_k1_sweep.txt 2>&1; grep '^{' $O/r2t18_k1_sweep.txt | cut -c1-230
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:rg_search_kernelILb.ELi.ELi.ELb0 -s 4 -c 1 -o $O/r2t18_k1_L200 -f \
      python bench.py --L 200 --steps 2 --warmup 3 --no-cpu-baseline --knn-slice 0 > $O/r2t18_ncu_L200.log 2>&1; tail -c 200 $O/r2t18_ncu_L200.log
