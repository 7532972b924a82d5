#!/bin/bash
# Round 2, trip 31 (1 GPU): final evidence with the packed-FP32 K1 - whole GPU suite, smoke, bench (both arms), ncu --set full
# of K1 at L_pq = 55 (traffic profile for bench.py) and 500, launch list, beam-width sweep at 10 000- and 100 000-query batches.
mkdir -p gpurun_out
O=gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q ) > $O/r2t31_tests.log 2>&1; echo "tests exit $?"; tail -4 $O/r2t31_tests.log
( timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" ) 2>&1 | tail -2
( time timeout 900 python bench.py ) > $O/r2t31_bench_first.txt 2>&1; grep '^{' $O/r2t31_bench_first.txt > $O/r2t31_bench_line.json; cut -c1-300 $O/r2t31_bench_line.json
NCU="ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:rg_search_kernelILb.ELi.ELi.ELb0 -s 4 -c 1 -f"
timeout 600 $NCU --cache-control none -o $O/r2t31_k1_L55 python bench.py --L 55 --steps 2 --warmup 3 --no-cpu-baseline --knn-slice 0 > $O/r2t31_ncu_L55.log 2>&1; tail -c 120 $O/r2t31_ncu_L55.log
python tools/make_k1_traffic.py $O/r2t31_k1_L55.ncu-rep $O/r2t31_bench_line.json 2>&1 | tail -2; cp profiles/k1_traffic.json $O/r2t31_k1_traffic.json
timeout 600 $NCU -o $O/r2t31_k1_L500 python bench.py --L 500 --steps 2 --warmup 3 --no-cpu-baseline --knn-slice 0 > $O/r2t31_ncu_L500.log 2>&1; tail -c 120 $O/r2t31_ncu_L500.log
( time timeout 600 python bench.py --impl reference --steps 3 --warmup 1 ) > $O/r2t31_bench_ref.txt 2>&1; grep '^{' $O/r2t31_bench_ref.txt | cut -c1-400
( time timeout 900 python bench.py ) > $O/r2t31_bench.txt 2>&1; grep '^{' $O/r2t31_bench.txt | cut -c1-3500
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2t31_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $O/r2t31_launches_run.log 2>&1; tail -c 200 $O/r2t31_launches_run.log
( timeout 600 python tools/k1_sweep.py --Ls 55 100 150 200 300 500 --reps 6 --configs w=0 w=0 ) > $O/r2t31_sweep_10k.txt 2>&1; grep '^{' $O/r2t31_sweep_10k.txt | cut -c1-200
( timeout 900 python tools/k1_sweep.py --queries 100000 --Ls 55 100 200 300 500 --reps 3 --configs w=0 ) > $O/r2t31_sweep_100k.txt 2>&1; grep '^{' $O/r2t31_sweep_100k.txt | cut -c1-200; tail -3 $O/r2t31_sweep_100k.txt | cut -c1-300
