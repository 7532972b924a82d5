#!/bin/bash
# GPU-box visit: GPU tests, smoke, default bench, ncu launch list + full captures of K1 (search) and K2 (kNN GEMM).
mkdir -p gpurun_out
O=gpurun_out
{
  nvidia-smi -L; nproc; grep -m1 "model name" /proc/cpuinfo; grep -o -m1 'avx512f' /proc/cpuinfo; free -g | head -2
} > $O/env.txt 2>&1
( time python -m pytest tests -x -q -m gpu ) 2>&1 | tail -40 > $O/pytest_gpu.txt
tail -5 $O/pytest_gpu.txt
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.txt 2>&1; tail -2 $O/smoke.txt
N=${1:-500000}
( time timeout 1500 python -X faulthandler bench.py --n $N ) > $O/bench_main.txt 2>&1; tail -4 $O/bench_main.txt
LSEL=$(python - <<EOF
import json
for l in open("$O/bench_main.txt"):
    if l.startswith("{"):
        print(json.loads(l)["config"]["L_pq"]); break
else:
    print(35)
EOF
)
echo "L_sel=$LSEL"
# launch list of the same command (index cached in /tmp on this box); our kernels only
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'rg_|knn_' -c 400 --csv \
    --log-file $O/launches_bench.csv python bench.py --n $N --L $LSEL --steps 5 --warmup 3 --no-cpu-baseline > $O/launches_bench.log 2>&1
tail -3 $O/launches_bench.log
# full capture of the dominant kernel inside the bench workload
timeout 900 ncu --set full --clock-control none --import-source on -k regex:rg_search_kernel -s 2 -c 1 \
    -o $O/k1_bench -f python bench.py --n $N --L $LSEL --steps 2 --warmup 3 --no-cpu-baseline > $O/ncu_k1.log 2>&1
tail -3 $O/ncu_k1.log
# K1 tuning sweep on a random graph
timeout 600 python tools/microbench_search.py --Ls 20 35 100 --configs 2:0:0:8,2:0:0:16,1:0:0:8,3:0:0:8 --out $O/microbench.json > $O/microbench.txt 2>&1
tail -14 $O/microbench.txt
# kNN throughput + captures
timeout 600 python tools/microbench_knn.py --n 1000000 --nq 131072 > $O/knn_bench.txt 2>&1; tail -2 $O/knn_bench.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'knn_|absmax|to_half|compact|fill_f32' -c 200 --csv \
    --log-file $O/launches_knn.csv python tools/microbench_knn.py --n 1000000 --nq 32768 --reps 1 > $O/launches_knn.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:knn_gemm_filter_kernel -s 12 -c 1 \
    -o $O/k2_gemm -f python tools/microbench_knn.py --n 1000000 --nq 32768 --reps 1 > $O/ncu_k2.log 2>&1
tail -3 $O/ncu_k2.log
ls -la $O
