#!/bin/bash
# Build a RoarGraph index with the drop-in drivers - the workflow of the reference's run_roargraph_test.sh
# (same flags and canonical parameters M_sq=100, M_pjbp=35, L_pjpq=500) with the two GPU steps made explicit:
#   1. learn->base exact kNN  (replaces DiskANN's compute_groundtruth; tcgen05 GEMM + fused top-k + FP32 re-rank)
#   2. graph construction     (--gpu_build 1: all phases on the GPU; drop the flag for the edge-exact host build)
# usage: scripts/run_roargraph_test.sh [data dir, default data/t2i-10M]
set -e
root=$(cd "$(dirname "$0")/.." && pwd)
python "$root/__graft_entry__.py"            # nvcc sm_100a + host C++ layer + CLI drivers (no cmake needed)
bin=$root/mysteryann_b200/host/bin
prefix=${1:-data/t2i-10M}
[ -f ${prefix}/t2i.train.in.base.nn.dist.10M.ibin ] || \
$bin/compute_groundtruth --data_type float --dist_fn mips \
  --base_file ${prefix}/base.10M.fbin --query_file ${prefix}/query.learn.10M.fbin \
  --gt_file ${prefix}/t2i.train.in.base.nn.dist.10M.ibin --K 100
$bin/test_build_roargraph --data_type float --dist ip \
  --base_data_path ${prefix}/base.10M.fbin \
  --sampled_query_data_path ${prefix}/query.learn.10M.fbin \
  --projection_index_save_path ${prefix}/t2i_10M_roar.index \
  --learn_base_nn_path ${prefix}/t2i.train.in.base.nn.dist.10M.ibin \
  --M_sq 100 --M_pjbp 35 --L_pjpq 500 -T 64 --gpu_build 1
