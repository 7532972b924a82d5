"""world_size-2 gloo test of the multi-GPU kNN host logic (mysteryann_b200/sharded_knn.py) on CPU.

The CUDA kernels cannot run here, so the two compute steps (per-shard exact kNN, K4 merge) are played by the CPU
oracle / a numpy merge - the shard bounds, the all-to-all layout and the gather are the code under test.  The GPU
version of the same flow is tests/test_knn_gpu.py::test_knn_device_api_and_merge (one process) and
tools/bench_knn_sharded.py (torchrun, NCCL)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from mysteryann_b200 import sharded_knn  # noqa: E402


def test_shard_bounds():
    assert sharded_knn.shard_bounds(10, 3) == [0, 4, 7, 10]
    assert sharded_knn.shard_bounds(2, 4) == [0, 1, 2, 2, 2]
    assert sharded_knn.shard_bounds(0, 2) == [0, 0, 0]
    b = sharded_knn.shard_bounds(10_000_000, 8)
    assert b[0] == 0 and b[-1] == 10_000_000 and all(b[i + 1] - b[i] == 1_250_000 for i in range(8))
    with pytest.raises(ValueError):
        sharded_knn.shard_bounds(5, 0)


def numpy_merge(metric):
    def merge(pid, pd):
        G, m, K = pid.shape
        ids = pid.numpy().view(np.uint32).transpose(1, 0, 2).reshape(m, G * K)
        d = pd.numpy().transpose(1, 0, 2).reshape(m, G * K)
        score = np.where(ids == 0xFFFFFFFF, np.inf, -d if metric == 1 else d)
        order = np.lexsort((ids, score), axis=1)[:, :K]          # (score, id) ascending, like K4
        oid = np.take_along_axis(ids, order, axis=1)
        od = np.take_along_axis(d, order, axis=1)
        return torch.from_numpy(oid.view(np.int32).copy()), torch.from_numpy(od.copy())
    return merge


def _worker(rank, world, port, n, nq, dim, K, metric, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from mysteryann_b200 import synth
        from oracle.binding import Oracle

        o = Oracle()
        base, q, _ = synth.make_numpy(n, nq, 1, dim, seed=7)
        b = sharded_knn.shard_bounds(n, world)
        shard = base[b[rank]:b[rank + 1]]

        def local_knn(queries):
            ids, d, _ = o.exact_knn(shard, queries.numpy(), K, metric=metric)
            ids = np.where(ids == 0xFFFFFFFF, ids, ids + np.uint32(b[rank]))      # id_base
            return torch.from_numpy(ids.view(np.int32).copy()), torch.from_numpy(d.copy())

        ids, d, qb = sharded_knn.knn_sharded_with(local_knn, numpy_merge(metric), torch.from_numpy(q), K, gather=True)
        assert qb == sharded_knn.shard_bounds(nq, world)
        want_ids, want_d, _ = o.exact_knn(base, q, K, metric=metric)
        assert np.array_equal(ids.numpy().view(np.uint32), want_ids), f"rank {rank}: ids differ"
        assert np.array_equal(d.numpy().view(np.uint32), want_d.view(np.uint32)), f"rank {rank}: dists differ"
        # slice-only variant: this rank's merged slice is the matching rows of the full answer
        sid, sd, qb = sharded_knn.knn_sharded_with(local_knn, numpy_merge(metric), torch.from_numpy(q), K, gather=False)
        assert np.array_equal(sid.numpy().view(np.uint32), want_ids[qb[rank]:qb[rank + 1]])
        open(os.path.join(out_dir, f"ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n,nq,dim,K,metric", [(1001, 37, 24, 10, 1), (640, 5, 16, 20, 0), (3, 9, 8, 4, 1)])
def test_sharded_knn_gloo_world2(tmp_path, n, nq, dim, K, metric):
    import socket

    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_worker, args=(2, port, n, nq, dim, K, metric, str(tmp_path)), nprocs=2, join=True)
    assert os.path.exists(tmp_path / "ok0") and os.path.exists(tmp_path / "ok1")


def test_product_path_refuses_cpu_tensors():
    from mysteryann_b200 import capi

    with pytest.raises(capi.RoarGraphError):
        sharded_knn.knn_sharded(torch.zeros(4, 8), 0, torch.zeros(2, 8), 2)


def test_grid_layout():
    g = sharded_knn.Grid(2, world=8, rank=5)
    assert (g.base_shards, g.query_groups, g.shard, g.qgroup) == (2, 4, 1, 2)
    assert g.base_bounds(10) == (5, 10) and g.query_bounds(103) == (52, 78)
    assert sharded_knn.Grid(8, world=8, rank=3).query_bounds(100) == (0, 100)      # plain base sharding
    assert sharded_knn.Grid(1, world=8, rank=3).base_bounds(100) == (0, 100)       # plain query sharding
    rb = sharded_knn.Grid(2, world=4, rank=0).result_bounds(10)
    assert rb == [0, 3, 5, 8, 10]
    with pytest.raises(ValueError):
        sharded_knn.Grid(3, world=8, rank=0)
    # grid_layout (the row ranges rg_knn_exact_grid works on) agrees with Grid for every rank
    for world, bs, n, nq in ((8, 2, 1001, 103), (4, 4, 10, 7), (4, 1, 9, 11)):
        for r in range(world):
            g = sharded_knn.Grid(bs, world=world, rank=r)
            rb = g.result_bounds(nq)
            assert sharded_knn.grid_layout(r, world, bs, n, nq) == (g.base_bounds(n), g.query_bounds(nq), (rb[r], rb[r + 1]))


def _grid_worker(rank, world, port, base_shards, n, nq, dim, K, metric, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from mysteryann_b200 import synth
        from oracle.binding import Oracle

        o = Oracle()
        base, q, _ = synth.make_numpy(n, nq, 1, dim, seed=9)
        grid = sharded_knn.Grid(base_shards).make_groups()
        lo, hi = grid.base_bounds(n)
        q0, q1 = grid.query_bounds(nq)
        shard = base[lo:hi]

        def local_knn(queries):
            ids, d, _ = o.exact_knn(shard, queries.numpy(), K, metric=metric, threads=2)
            ids = np.where(ids == 0xFFFFFFFF, ids, ids + np.uint32(lo))
            return torch.from_numpy(ids.view(np.int32).copy()), torch.from_numpy(d.copy())

        sid, sd, _ = sharded_knn.knn_sharded_with(local_knn, numpy_merge(metric), torch.from_numpy(q[q0:q1]), K,
                                                  group=grid.group, gather=False)
        rb = grid.result_bounds(nq)
        assert sid.shape[0] == rb[rank + 1] - rb[rank]
        ids = sharded_knn.gather_rows(sid, rb)          # over the whole world
        d = sharded_knn.gather_rows(sd, rb)
        want_ids, want_d, _ = o.exact_knn(base, q, K, metric=metric, threads=2)
        assert np.array_equal(ids.numpy().view(np.uint32), want_ids), f"rank {rank}: ids differ"
        assert np.array_equal(d.numpy().view(np.uint32), want_d.view(np.uint32)), f"rank {rank}: dists differ"
        open(os.path.join(out_dir, f"ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,base_shards", [(4, 2), (4, 1), (4, 4), (2, 1)])
def test_grid_knn_gloo(tmp_path, world, base_shards):
    """2-D decomposition (base shards x query groups) on CPU: 2x2, query-only, base-only layouts give the full answer."""
    import socket

    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_grid_worker, args=(world, port, base_shards, 403, 29, 16, 7, 1, str(tmp_path)), nprocs=world, join=True)
    assert all(os.path.exists(tmp_path / f"ok{r}") for r in range(world))
