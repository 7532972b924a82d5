#!/bin/bash
# Round evidence at the named config: bench line, reference arm, ncu launch list of the same command, full K1 capture.
mkdir -p gpurun_out
O=gpurun_out
N=${1:-10000000}
( time timeout 1500 python -X faulthandler bench.py --n $N ) > $O/r01_bench_10m.txt 2>&1; tail -c 600 $O/r01_bench_10m.txt
( time timeout 900 python -X faulthandler bench.py --impl reference --n $N --steps 3 --warmup 1 ) > $O/r01_bench_ref_10m.txt 2>&1; tail -c 900 $O/r01_bench_ref_10m.txt
LSEL=$(python - <<PY
import json
for l in open("$O/r01_bench_10m.txt"):
    if l.startswith("{"):
        print(json.loads(l)["config"]["L_pq"]); break
else:
    print(60)
PY
)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'rg_search' -c 200 --csv \
    --log-file $O/r01_launches_bench_10m.csv python bench.py --n $N --L $LSEL --steps 5 --warmup 3 --no-cpu-baseline > $O/launches_bench.log 2>&1
tail -2 $O/launches_bench.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:rg_search_kernel -s 6 -c 1 \
    -o $O/r01_k1_bench_10m -f python bench.py --n $N --L $LSEL --steps 2 --warmup 3 --no-cpu-baseline > $O/ncu_k1.log 2>&1
tail -2 $O/ncu_k1.log
ls -la $O | head -40
