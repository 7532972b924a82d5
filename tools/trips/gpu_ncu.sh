#!/bin/bash
# ncu capture of the primary K1 kernel on the synthetic random-graph microbench (one launch, full set + source)
mkdir -p gpurun_out
CFG=${1:-2:0:0:8}
L=${2:-20}
timeout 900 ncu --set full --clock-control none --import-source on -k regex:rg_search_kernel -s 2 -c 1 \
    -o gpurun_out/k1_prof -f python tools/microbench_search.py --configs $CFG --Ls $L > gpurun_out/ncu_log.txt 2>&1
tail -5 gpurun_out/ncu_log.txt
ls -la gpurun_out/
