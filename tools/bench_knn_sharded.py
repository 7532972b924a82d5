"""Multi-GPU build kNN (BASELINE.json config C4 shape): base sharded over the ranks, K2/K3 per shard, NCCL all-to-all of
the per-shard lists, K4 merge (mysteryann_b200/sharded_knn.py).  Checks the merged answer of a query sample against a
single-GPU run of the same kernels on the whole base (when it fits) and prints one JSON line (rank 0).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tools/bench_knn_sharded.py --rows 10000000 --queries 262144
"""
import argparse
import json
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from mysteryann_b200 import build, capi, sharded_knn, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", "--rows", dest="n", type=int, default=10_000_000)  # under torchrun use --rows (--n is an ambiguous prefix of its own options)
    ap.add_argument("--nq", "--queries", dest="nq", type=int, default=262_144)
    ap.add_argument("--dim", type=int, default=200)
    ap.add_argument("--K", type=int, default=100)
    ap.add_argument("--check", type=int, default=4096, help="queries verified against the unsharded answer (0 = none)")
    ap.add_argument("--base-shards", type=int, default=0, help="2-D layout: base shards x query groups (sharded_knn.Grid); "
                                                             "0 = one base shard per rank, the scheme of config C4")
    ap.add_argument("--exchange", default="capi", choices=["capi", "torch"],
                    help="capi = rg_knn_exact_sharded (grouped ncclSend/ncclRecv inside the library), torch = all_to_all_single")
    ap.add_argument("--warm", type=int, default=65536, help="queries of the untimed warm-up call (NCCL channels, scratch allocation)")
    a = ap.parse_args()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    if rank == 0:
        build.build()
    dist.barrier()
    base, train, _ = synth.make_torch(a.n, a.nq, 1, a.dim, device=dev)   # same seed on every rank: identical data
    st = torch.cuda.current_stream().cuda_stream
    knn = sharded_knn.knn_sharded if a.exchange == "capi" else sharded_knn.knn_sharded_torch
    if a.base_shards in (0, world):
        b = sharded_knn.shard_bounds(a.n, world)
        shard = base[b[rank]:b[rank + 1]]

        def run(q=train):
            return knn(shard, b[rank], q, a.K, metric=capi.METRIC_IP, gather=False, stream=st)
    else:
        # base shards x query groups: this rank answers its group's queries against its base shard; the exchange and the
        # merge run inside the group.  `qb` below are the bounds of the merged slices in rank order.
        grid = sharded_knn.Grid(a.base_shards)
        if a.exchange != "capi":
            grid.make_groups()
        lo, hi = grid.base_bounds(a.n)
        shard = base[lo:hi]

        def run(q=train):
            q0, q1 = grid.query_bounds(q.shape[0])
            if a.exchange == "capi":   # rg_knn_exact_grid: one communicator, the exchange runs inside each query group
                ids, d = sharded_knn.knn_grid(shard, lo, q[q0:q1].contiguous(), a.K, a.base_shards, metric=capi.METRIC_IP, stream=st)
            else:
                ids, d, _ = knn(shard, lo, q[q0:q1].contiguous(), a.K, metric=capi.METRIC_IP, group=grid.group, gather=False, stream=st)
            return ids, d, grid.result_bounds(q.shape[0])

    run(train[:min(a.warm, a.nq)].contiguous())  # warm-up (NCCL channels, scratch)
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    ids, d, qb = run()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = t.item()
    ok = None
    if a.check:
        m = min(a.check, qb[rank + 1] - qb[rank])
        want_i = torch.empty((m, a.K), dtype=torch.int32, device=dev)
        want_d = torch.empty((m, a.K), dtype=torch.float32, device=dev)
        capi.knn_exact_device(base, train[qb[rank]:qb[rank] + m].contiguous(), a.K, want_i, want_d, metric=capi.METRIC_IP, stream=st)
        torch.cuda.synchronize()
        same = torch.tensor([int(torch.equal(want_i, ids[:m]) and torch.equal(want_d, d[:m]))], device=dev)
        dist.all_reduce(same, op=dist.ReduceOp.MIN)
        ok = bool(same.item())
    if rank == 0:
        flops = 2.0 * a.n * a.nq * a.dim
        print(json.dumps(dict(n=a.n, nq=a.nq, dim=a.dim, K=a.K, n_gpus=world, ms=round(ms, 2),
                              tflops_algorithmic_total=round(flops / (ms * 1e-3) / 1e12, 1),
                              tflops_per_gpu=round(flops / (ms * 1e-3) / 1e12 / world, 1),
                              c4_extrapolated_s=round(ms * 1e-3 * (10_000_000 / a.nq) * (10_000_000 / a.n), 1),
                              sharded_equals_unsharded=ok, base_shards=a.base_shards or world,
                              exchange=("rg_knn_exact_sharded" if a.base_shards in (0, world) else "rg_knn_exact_grid") + ": grouped ncclSend/ncclRecv + K4 merge" if a.exchange == "capi"
                              else "torch all_to_all_single + K4 merge", knn_stats=capi.knn_last_stats())), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
