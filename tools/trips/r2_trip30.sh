#!/bin/bash
# Round 2, trip 30 (1 GPU): K1 at large beam widths with fewer resident queries (slabs of the bucketed visited set inside L2):
# warps per query x CTAs per SM at L_pq = 200 / 300 / 500, run-time options only, one process on the default library.
mkdir -p gpurun_out
O=gpurun_out
( timeout 900 python tools/k1_sweep.py --Ls 200 300 500 --reps 6 --configs w=2 w=4 w=4,c=5 w=4,c=4 w=3 w=3,c=6 w=2,c=10 w=2,c=9 w=2,c=8 w=2,c=7 w=2 w=4 ) > $O/r2t30_sweep.txt 2>&1
grep '^{' $O/r2t30_sweep.txt | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['cfg'].ljust(10), d['L'], d['ms'], d['frac'], d['same_as_first_cfg'], d['overflow'], d['smi'])"
tail -3 $O/r2t30_sweep.txt | cut -c1-300
