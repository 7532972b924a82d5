#!/bin/bash
# K1 at larger beam widths, part 2: warps per query (2 vs 4 vs 3), denser visited hash (hash_log2 one step smaller), partial
# persisting window when the slabs outgrow the set-aside.  bench.py at L = 100 / 200 / 500.
mkdir -p gpurun_out
O=gpurun_out
run() {  # L warps hash_log2 partial
  tag="L$1_w$2_h$3_p$4"
  if [ "$4" != "0" ]; then export RG_SEARCH_PERSIST_PARTIAL=$4; else unset RG_SEARCH_PERSIST_PARTIAL; fi
  ( timeout 900 python bench.py --L $1 --warps $2 --hash-log2 $3 --no-cpu-baseline --steps 10 ) > $O/bench_$tag.txt 2>&1
  python - $tag <<'PY'
import json, sys
tag = sys.argv[1]
fn = f"gpurun_out/bench_{tag}.txt"
for line in open(fn):
    if line.startswith("{"):
        j = json.loads(line)
        print(tag, "value", j["value"], "ms", j["ms_per_step"], "frac", j["roofline"]["frac"], "overflow", j["config"]["visited_overflow_queries"])
        break
else:
    print(tag, "FAILED"); print(open(fn).read()[-800:])
PY
}
{
run 100 0 0 0; run 100 4 0 0; run 100 3 0 0; run 100 0 0 0.75; run 100 0 0 0.5; run 100 4 13 0
run 200 0 0 0; run 200 4 0 0; run 200 0 14 0; run 200 0 0 0.75; run 200 4 14 0
run 500 0 0 0; run 500 4 0 0; run 500 8 0 0; run 500 0 15 0
run 55 0 0 0; run 55 3 0 0; run 30 0 0 0; run 30 0 12 0
} | tee $O/k1_large_L_variants.txt
