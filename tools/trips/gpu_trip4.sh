#!/bin/bash
# K1 occupancy experiments: global (L2-resident) visited hash, single/double staging buffers, register landing
mkdir -p gpurun_out
timeout 900 python tools/microbench_search.py --Ls 35 100 --configs \
2:0:0:8:0:2,2:0:0:8:1:2,2:0:0:8:1:1,2:0:0:16:1:1,2:0:0:16:1:2,1:0:0:8:1:1,1:0:0:8:1:2,3:0:0:8:1:1,3:0:0:8:0:1,2:0:0:8:0:1,2:4:0:8:1:1,2:8:0:8:1:1 \
  --out gpurun_out/microbench4.json > gpurun_out/microbench4.txt 2>&1
cat gpurun_out/microbench4.txt | grep -v "^$" | tail -30
