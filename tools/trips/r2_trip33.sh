#!/bin/bash
# Round 2, trip 33 (1 GPU): launch list of the final bench command with the timed search steps inside the window: a first plain
# run fills the index cache, the second runs under ncu --metrics gpu__time_duration.sum.
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --knn-slice 0 > $O/r2t33_prep.log 2>&1; tail -c 200 $O/r2t33_prep.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $O/r2t33_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $O/r2t33_launches_run.log 2>&1; tail -c 300 $O/r2t33_launches_run.log
python tools/launch_summary.py $O/r2t33_launches.csv 2>/dev/null | head -14
