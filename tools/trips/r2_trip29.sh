#!/bin/bash
# Round 2, trip 29 (1 GPU): K1 with the packed-FP32 distance loop (FFMA2 / FADD2), the pair-wise pool shift and the L2
# prefetch of a warp's next gather batch (adj_prefetch bit 3): whole GPU suite on the new default library, then same-box
# sweeps of the A/B builds (tools/build_variant.sh: base = scalar distance + per-entry shift, x2 = packed distance only,
# pin = new default + SR_TID.X kept in a register) with pf=3 vs pf=11.
mkdir -p gpurun_out
O=gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q -x ) > $O/r2t29_tests.log 2>&1; echo "tests exit $?"; tail -4 $O/r2t29_tests.log
run() {  # name lib
  ( RG_B200_LIB=$2 timeout 600 python tools/k1_sweep.py --Ls 55 100 200 500 --reps 8 --configs pf=3 pf=11 pf=3 pf=11 ) > $O/r2t29_sweep_$1.txt 2>&1
  echo "== $1"; grep '^{' $O/r2t29_sweep_$1.txt | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['cfg'].ljust(6), d['L'], d['ms'], d['frac'], d['same_as_first_cfg'], d['crc'], d['smi'])"
}
run base1 mysteryann_b200/variants/base.so
run main1 mysteryann_b200/libroargraph_b200.so
run pin mysteryann_b200/variants/pin.so
run x2 mysteryann_b200/variants/x2.so
run base2 mysteryann_b200/variants/base.so
run main2 mysteryann_b200/libroargraph_b200.so
