#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.txt 2>&1; tail -2 gpurun_out/smoke.txt
( time timeout 900 python -X faulthandler bench.py --n 100000 --steps 5 ) > gpurun_out/bench_100k.txt 2>&1; tail -4 gpurun_out/bench_100k.txt
( time timeout 600 python -X faulthandler bench.py --impl reference --n 100000 --steps 3 --warmup 1 ) > gpurun_out/bench_ref_100k.txt 2>&1; tail -4 gpurun_out/bench_ref_100k.txt
( time timeout 1500 python -X faulthandler bench.py ) > gpurun_out/bench_500k.txt 2>&1; tail -4 gpurun_out/bench_500k.txt
