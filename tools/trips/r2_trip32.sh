#!/bin/bash
# Round 2, trip 32 (1 GPU): launch list of the final bench command (the build-kNN slice, whose three repetitions alone are
# ~300 launches, switched off so that the timed search steps are inside the captured window).
mkdir -p gpurun_out
O=gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $O/r2t32_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --knn-slice 0 > $O/r2t32_launches_run.log 2>&1; tail -c 300 $O/r2t32_launches_run.log
python tools/launch_summary.py $O/r2t32_launches.csv | head -20
