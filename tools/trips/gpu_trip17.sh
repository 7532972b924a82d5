#!/bin/bash
# K2 dense-hit bulk append + prefix-merge select: parity tests, throughput at 10M, launch shares; build-search cache-hint sweep at 2M
mkdir -p gpurun_out
O=gpurun_out
( timeout 600 python -m pytest tests/test_knn_gpu.py tests/test_cli_gpu.py tests/test_build_gpu.py -x -q ) > $O/t17_tests.log 2>&1; echo "tests exit $?"; tail -4 $O/t17_tests.log
python - <<'PY'
import torch
p = torch.cuda.get_device_properties(0)
print("L2", p.L2_cache_size, "persisting max", getattr(p, "persisting_l2_cache_max_size", None), "window max", getattr(p, "access_policy_max_window_size", None))
PY
timeout 300 python tools/microbench_knn.py --n 10000000 --nq 131072 --reps 2 > $O/knn_mb_10M.log 2>&1; tail -1 $O/knn_mb_10M.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_knn_10Mx32K_v2.csv \
    python tools/microbench_knn.py --n 10000000 --nq 32768 --reps 1 > $O/ncu_k2_list.log 2>&1; tail -c 300 $O/ncu_k2_list.log
for v in 3:3 0:0 1:0 0:3 1:3; do
  RG_BUILD_L2_HINT=${v%%:*} RG_BUILD_ADJ_PREFETCH=${v##*:} timeout 300 python tools/microbench_build.py --n 2000000 2>&1 | tail -1
done | tee $O/build_hints_2M.txt
for w in 1 4; do
  RG_BUILD_L2_HINT=0 RG_BUILD_ADJ_PREFETCH=0 RG_BUILD_WARPS=$w timeout 300 python tools/microbench_build.py --n 2000000 2>&1 | tail -1
done | tee -a $O/build_hints_2M.txt
