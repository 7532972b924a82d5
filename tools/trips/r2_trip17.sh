#!/bin/bash
# Round 2, trip 17 (1 GPU): headline bench with the reworked K1, drift of one configuration inside a long sweep (SM clock
# sampled under load), ncu --set full at L_pq = 200 and 55.
mkdir -p gpurun_out
O=gpurun_out
python - <<'PY'
import torch
p = torch.cuda.get_device_properties(0)
print("L2", p.L2_cache_size, "persisting max", getattr(p, "persisting_l2_cache_max_size", None), "window max", getattr(p, "access_policy_max_window_size", None))
PY
( time timeout 1200 python bench.py ) > $O/r2t17_bench.txt 2>&1; grep '^{' $O/r2t17_bench.txt | cut -c1-1200
( timeout 1500 python tools/k1_sweep.py --Ls 55 100 200 500 --reps 6 --configs hs=4 hs=4 hs=4 hs=0 hs=0 hs=4 --out $O/r2t17_k1_sweep.json ) > $O/r2t17_k1_sweep.txt 2>&1; grep '^{' $O/r2t17_k1_sweep.txt | cut -c1-230
for L in 200 55; do
  timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:rg_search_kernelILb.ELi.ELi.ELb0 -s 4 -c 1 -o $O/r2t17_k1_L$L -f \
      python bench.py --L $L --steps 2 --warmup 3 --no-cpu-baseline --knn-slice 0 > $O/r2t17_ncu_L$L.log 2>&1; tail -c 200 $O/r2t17_ncu_L$L.log
done
