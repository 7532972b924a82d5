#!/bin/bash
# Round 2, trip 35 (1 GPU): packed-FP32 distance with direct 64-bit pair loads in K1 and in the occlusion prunes of the graph
# build (scalar main loop removed): whole GPU suite (search parity, prune list-level parity, build quality), bench (first run
# builds the index: graph_build_phases_s), ncu --set full at L_pq = 55 for the traffic profile, bench again, C3 k = 10.
mkdir -p gpurun_out
O=gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q ) > $O/r2t35_tests.log 2>&1; echo "tests exit $?"; tail -4 $O/r2t35_tests.log
( timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" ) 2>&1 | tail -1
( time timeout 900 python bench.py ) > $O/r2t35_bench_first.txt 2>&1; grep '^{' $O/r2t35_bench_first.txt > $O/r2t35_bench_line.json; cut -c1-200 $O/r2t35_bench_line.json
NCU="ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:rg_search_kernelILb.ELi.ELi.ELb0 -s 4 -c 1 -f"
timeout 600 $NCU --cache-control none -o $O/r2t35_k1_L55 python bench.py --L 55 --steps 2 --warmup 3 --no-cpu-baseline --knn-slice 0 > $O/r2t35_ncu_L55.log 2>&1; tail -c 100 $O/r2t35_ncu_L55.log
python tools/make_k1_traffic.py $O/r2t35_k1_L55.ncu-rep $O/r2t35_bench_line.json 2>&1 | tail -1 | cut -c1-300; cp profiles/k1_traffic.json $O/r2t35_k1_traffic.json
( time timeout 900 python bench.py ) > $O/r2t35_bench.txt 2>&1; grep '^{' $O/r2t35_bench.txt | cut -c1-300
( time timeout 900 python bench.py --config C3 ) > $O/r2t35_bench_c3.txt 2>&1; grep '^{' $O/r2t35_bench_c3.txt | cut -c1-300
( timeout 600 python tools/k1_sweep.py --Ls 55 100 200 500 --reps 6 --configs w=0 w=0 ) > $O/r2t35_sweep.txt 2>&1; grep '^{' $O/r2t35_sweep.txt | cut -c1-200
