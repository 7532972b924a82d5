#!/bin/bash
# Re-entry validation: full GPU suite, smoke, both bench arms at the 10M default, launch list + ncu --set full of K1 inside
# bench.py (roofline.traffic), K2 launch shares at 10M, and the full-train-set (10M) index experiment.
mkdir -p gpurun_out
O=gpurun_out
nproc > $O/host.txt; nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv >> $O/host.txt
( time timeout 900 python -m pytest tests -m gpu -x -q ) > $O/gpu_tests.log 2>&1; echo "gpu tests exit $?"; tail -5 $O/gpu_tests.log
( time timeout 300 python __graft_entry__.py --smoke ) > $O/smoke.log 2>&1; tail -2 $O/smoke.log
( time timeout 900 python bench.py --impl reference --steps 3 --warmup 1 ) > $O/bench_ref.txt 2>&1; tail -c 400 $O/bench_ref.txt
( time timeout 900 python bench.py ) > $O/bench.txt 2>&1; tail -c 1800 $O/bench.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $O/launches_bench.csv \
    python bench.py --L 60 --steps 3 --warmup 3 --no-cpu-baseline > $O/ncu_list.log 2>&1; tail -c 300 $O/ncu_list.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:rg_search_kernel -s 10 -c 1 -o $O/k1_bench10m -f \
    python bench.py --L 60 --steps 3 --warmup 3 --no-cpu-baseline > $O/ncu_full.log 2>&1; tail -c 300 $O/ncu_full.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_knn_10Mx64K.csv \
    python tools/microbench_knn.py --n 10000000 --nq 65536 --reps 1 > $O/ncu_k2_list.log 2>&1; tail -c 300 $O/ncu_k2_list.log
( time timeout 1200 python bench.py --train 10000000 --no-cpu-baseline ) > $O/bench_train10m.txt 2>&1; tail -c 1800 $O/bench_train10m.txt
ls -la $O | head -40
