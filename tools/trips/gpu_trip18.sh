#!/bin/bash
# Session-8 validation on one GPU: full GPU suite (incl. the zero-copy host path), both bench arms at the 10M default with
# the reference's L sweep, e2e A/B (zero-copy vs staged), launch list + ncu --set full of K1 at the selected L,
# the C3 shape (2.5M x 512, unit rows) and the C5-scale bit-exactness check (100M x 200 on one GPU).
mkdir -p gpurun_out
O=gpurun_out
nproc > $O/host.txt; free -g >> $O/host.txt; nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv >> $O/host.txt
( time timeout 900 python -m pytest tests -m gpu -x -q ) > $O/gpu_tests.log 2>&1; echo "gpu tests exit $?"; tail -5 $O/gpu_tests.log
( time timeout 300 python __graft_entry__.py --smoke ) > $O/smoke.log 2>&1; tail -2 $O/smoke.log
( time timeout 900 python bench.py --impl reference --steps 3 --warmup 1 ) > $O/bench_ref.txt 2>&1; tail -c 400 $O/bench_ref.txt
( time timeout 900 python bench.py ) > $O/bench.txt 2>&1; tail -c 1800 $O/bench.txt
L=$(python - <<'PY'
import json
for line in open("gpurun_out/bench.txt"):
    if line.startswith("{"):
        print(json.loads(line)["config"]["L_pq"]); break
else:
    print(55)
PY
)
echo "selected L = $L"
( timeout 600 python bench.py --L $L --zero-copy 0 --no-cpu-baseline ) > $O/bench_staged.txt 2>&1; tail -c 700 $O/bench_staged.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $O/launches_bench.csv \
    python bench.py --L $L --steps 3 --warmup 3 --no-cpu-baseline > $O/ncu_list.log 2>&1; tail -c 300 $O/ncu_list.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:rg_search_kernel -s 10 -c 1 -o $O/k1_bench10m -f \
    python bench.py --L $L --steps 3 --warmup 3 --no-cpu-baseline > $O/ncu_full.log 2>&1; tail -c 300 $O/ncu_full.log
( time timeout 900 python bench.py --n 2500000 --dim 512 --normalize ) > $O/bench_c3.txt 2>&1; tail -c 1500 $O/bench_c3.txt
( time timeout 900 python tools/check_large_index.py ) > $O/check_100m.txt 2>&1; tail -c 1500 $O/check_100m.txt
ls -la $O | head -40
