#!/bin/bash
# Drain pass (PDL-launched 8-warp grid for the batch tail): parity tests, then bench at 10M with drain = 0 / 30 / 60 / 100 / 150 %
mkdir -p gpurun_out
O=gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q ) > $O/gpu_tests.log 2>&1; echo "gpu tests exit $?"; tail -5 $O/gpu_tests.log
for d in 0 60 30 100 150 0 60; do
  ( timeout 900 python bench.py --L 55 --drain $d --no-cpu-baseline ) > $O/bench_drain_$d.txt 2>&1
  python - $d <<'PY'
import json, sys
d = sys.argv[1]
for line in open(f"gpurun_out/bench_drain_{d}.txt"):
    if line.startswith("{"):
        j = json.loads(line)
        print("drain", d, "value", j["value"], "ms", j["ms_per_step"], "e2e", j["e2e"]["value"], "frac", j["roofline"]["frac"], "launches", j["gpu_launches"], "clk", j["clocks"]["sm_mhz"])
        break
else:
    print("drain", d, "FAILED"); print(open(f"gpurun_out/bench_drain_{d}.txt").read()[-1500:])
PY
done | tee $O/drain_sweep.txt
