#!/bin/bash
# One GPU-box visit: environment facts, GPU tests, micro-benchmark.  Output lands in gpurun_out/.
mkdir -p gpurun_out
{
  nvidia-smi -L; nproc; grep -m1 "model name" /proc/cpuinfo; grep -o -m1 'avx512f' /proc/cpuinfo; free -g | head -2
} > gpurun_out/env.txt 2>&1
python -m pytest tests -x -q -m gpu 2>&1 | tail -40 > gpurun_out/pytest_gpu.txt
tail -5 gpurun_out/pytest_gpu.txt
timeout 600 python tools/microbench_search.py --out gpurun_out/microbench.json > gpurun_out/microbench.txt 2>&1
tail -30 gpurun_out/microbench.txt
