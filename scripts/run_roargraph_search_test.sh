#!/bin/bash
# Search evaluation with the drop-in driver - the workflow of the reference's run_roargraph_search_test.sh: same flags,
# same L_pq sweep, same CSV (L_pq,qps,avg_cmps,mean_latency_ms,recall,avg_hops).  Each L_pq is one batched GPU search
# over all queries; -T is accepted for compatibility (there are no host search threads).
# usage: scripts/run_roargraph_search_test.sh [data dir, default data/t2i-10M]
set -e
root=$(cd "$(dirname "$0")/.." && pwd)
python "$root/__graft_entry__.py"
bin=$root/mysteryann_b200/host/bin
prefix=${1:-data/t2i-10M}
num_threads=16
topk=10
[ -f ${prefix}/groundtruth.base.10M.query.10k.ibin ] || \
$bin/compute_groundtruth --data_type float --dist_fn mips \
  --base_file ${prefix}/base.10M.fbin --query_file ${prefix}/query.public.10k.fbin \
  --gt_file ${prefix}/groundtruth.base.10M.query.10k.ibin --K 100
$bin/test_search_roargraph --data_type float --dist ip \
  --base_data_path ${prefix}/base.10M.fbin \
  --projection_index_save_path ${prefix}/t2i_10M_roar.index \
  --gt_path ${prefix}/groundtruth.base.10M.query.10k.ibin \
  --query_path ${prefix}/query.public.10k.fbin \
  --L_pq 10 15 20 25 30 35 40 45 50 55 60 65 70 75 80 85 90 95 100 110 120 130 140 150 160 170 180 190 200 220 240 260 280 \
         300 350 400 450 500 550 600 650 700 750 800 900 1000 1100 1200 1300 1400 1500 1600 1700 1800 1900 2000 \
  --k ${topk} -T ${num_threads} \
  --evaluation_save_path ${prefix}/test_search_t2i_10M_top${topk}_T${num_threads}.csv
