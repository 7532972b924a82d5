#!/bin/bash
# K1 at larger beam widths: once the visited-hash slabs of the resident CTAs outgrow L2 (L_pq >= 80: 64 KB per CTA x 1924),
# does running fewer CTAs per SM (slabs back in L2) beat full occupancy?  bench.py --L {100,200} x --ctas
mkdir -p gpurun_out
O=gpurun_out
python - <<'PY'
from cuda import cudart
for name in ("cudaDevAttrMaxPersistingL2CacheSize", "cudaDevAttrMaxAccessPolicyWindowSize", "cudaDevAttrL2CacheSize"):
    print(name, cudart.cudaDeviceGetAttribute(getattr(cudart.cudaDeviceAttr, name), 0))
PY
for cfg in 100:0:0 100:10:0 100:8:0 100:6:0 100:0:13 200:0:0 200:8:0 200:6:0 200:4:0 80:0:0 80:10:0 80:8:0; do
  L=${cfg%%:*}; rest=${cfg#*:}; C=${rest%%:*}; H=${rest##*:}
  ( timeout 900 python bench.py --L $L --ctas $C --hash-log2 $H --no-cpu-baseline --steps 10 ) > $O/bench_L${L}_c${C}_h${H}.txt 2>&1
  python - $L $C $H <<'PY'
import json, sys
L, C, H = sys.argv[1:4]
fn = f"gpurun_out/bench_L{L}_c{C}_h{H}.txt"
for line in open(fn):
    if line.startswith("{"):
        j = json.loads(line)
        print("L", L, "ctas", C, "hash_log2", H, "value", j["value"], "ms", j["ms_per_step"], "frac", j["roofline"]["frac"], "overflow", j["config"]["visited_overflow_queries"], "recall", j["config"]["recall_at_10"])
        break
else:
    print("L", L, "ctas", C, "FAILED"); print(open(fn).read()[-800:])
PY
done | tee $O/k1_large_L_occupancy.txt
