#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
( time timeout 600 python -m pytest tests/test_build_gpu.py -q -m gpu -s ) > $O/pytest_build.txt 2>&1
grep -E "^metric=|^degrees|passed|failed|Error" $O/pytest_build.txt
