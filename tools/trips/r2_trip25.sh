#!/bin/bash
# Round 2, trip 25 (1 GPU): final evidence - whole GPU suite, smoke, bench (both arms), ncu --set full of the final K1 at
# L_pq = 55 / 200 / 500 (traffic profile for bench.py), launch list, drop-in CLI pipeline over the reference's L_pq sweep.
mkdir -p gpurun_out
O=gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q ) > $O/r2t25_tests.log 2>&1; echo "tests exit $?"; tail -5 $O/r2t25_tests.log
( timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" ) 2>&1 | tail -2
( time timeout 1200 python bench.py ) > $O/r2t25_bench_first.txt 2>&1; grep '^{' $O/r2t25_bench_first.txt > $O/r2t25_bench_line.json; cut -c1-400 $O/r2t25_bench_line.json
NCU="ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:rg_search_kernelILb.ELi.ELi.ELb0 -s 4 -c 1 -f"
timeout 900 $NCU --cache-control none -o $O/r2t25_k1_L55 python bench.py --L 55 --steps 2 --warmup 3 --no-cpu-baseline --knn-slice 0 > $O/r2t25_ncu_L55.log 2>&1; tail -c 120 $O/r2t25_ncu_L55.log
python tools/make_k1_traffic.py $O/r2t25_k1_L55.ncu-rep $O/r2t25_bench_line.json 2>&1 | tail -2; cp profiles/k1_traffic.json $O/r2t25_k1_traffic.json
for L in 200 500; do
  timeout 900 $NCU -o $O/r2t25_k1_L$L python bench.py --L $L --steps 2 --warmup 3 --no-cpu-baseline --knn-slice 0 > $O/r2t25_ncu_L$L.log 2>&1; tail -c 120 $O/r2t25_ncu_L$L.log
done
( time timeout 900 python bench.py --impl reference --steps 3 --warmup 1 ) > $O/r2t25_bench_ref.txt 2>&1; grep '^{' $O/r2t25_bench_ref.txt | cut -c1-500
( time timeout 1200 python bench.py ) > $O/r2t25_bench.txt 2>&1; grep '^{' $O/r2t25_bench.txt | cut -c1-3500
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2t25_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $O/r2t25_launches_run.log 2>&1; tail -c 200 $O/r2t25_launches_run.log
( time timeout 1500 python tools/cli_pipeline_10m.py --n 10000000 --train 2000000 --out $O/r2t25_cli_pipeline_10m.txt ) > $O/r2t25_cli.log 2>&1; tail -50 $O/r2t25_cli_pipeline_10m.txt
