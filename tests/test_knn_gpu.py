"""GPU parity tests of K2/K3/K4 (exact kNN) through the C ABI against the CPU oracle / FP64 brute force.
ids must match the oracle except where FP32 scores tie within 1e-6 relative; distances within 1e-5 relative."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def capi():
    from mysteryann_b200 import build, capi

    build.build()
    assert capi.device_count() > 0
    return capi


def check_knn(ids, dists, want_ids, want_dists, tag):
    assert np.allclose(dists, want_dists, rtol=1e-5, atol=1e-5), f"{tag}: distances differ"
    bad = np.argwhere(ids != want_ids)
    for q, j in bad:   # only near-ties may swap
        d0, d1 = float(want_dists[q, j]), float(dists[q, j])
        assert abs(d0 - d1) <= 1e-6 * max(1.0, abs(d0)), f"{tag}: id mismatch q={q} rank={j}: {ids[q, j]} vs {want_ids[q, j]}"
    return len(bad)


@pytest.mark.parametrize("metric", (1, 0))
@pytest.mark.parametrize("n,nq,dim,K", [(5000, 300, 200, 100), (20000, 1000, 200, 100), (3000, 130, 64, 10),
                                        (700, 50, 8, 5), (40000, 257, 512, 32)])
def test_knn_vs_oracle(capi, oracle, metric, n, nq, dim, K):
    from mysteryann_b200 import synth

    base, train, _ = synth.make_numpy(n, nq, 1, dim, seed=n + dim)
    want_ids, want_d, _ = oracle.exact_knn(base, train, K, metric=metric)
    ids, d = capi.knn_exact(base, train, K, metric=metric)
    check_knn(ids, d, want_ids, want_d, f"n={n} dim={dim} metric={metric}")
    st = capi.knn_last_stats()
    assert st["launches"] > 0
    assert st["exact_scans"] <= nq // 10, st     # the certificate passes for nearly every query


def test_knn_small_base_and_id_offset(capi, oracle):
    rng = np.random.default_rng(0)
    base = rng.standard_normal((50, 16)).astype(np.float32)     # fewer base rows than the candidate list
    q = rng.standard_normal((9, 16)).astype(np.float32)
    want_ids, want_d, _ = oracle.exact_knn(base, q, 20, metric=1)
    ids, d = capi.knn_exact(base, q, 20, metric=1, id_base=1000)
    check_knn(ids - 1000, d, want_ids, want_d, "small")
    ids, d = capi.knn_exact(base[:7], q, 10, metric=0)          # K > n: tail filled with 0xFFFFFFFF
    assert (ids[:, 7:] == 0xFFFFFFFF).all() and (np.sort(ids[:, :7], axis=1) == np.arange(7)).all()


def test_knn_adversarial_order_uses_exact_scan(capi, oracle):
    """Base sorted by decreasing score for every query: each later block beats the threshold -> lists overflow ->
    the exact FP32 scan must take over and still return the exact answer."""
    rng = np.random.default_rng(1)
    dim = 32
    direction = rng.standard_normal(dim).astype(np.float32)
    scale = np.linspace(0.01, 4.0, 30000, dtype=np.float32)[:, None]       # increasing <q,b> with the row index
    base = (scale * direction[None, :] + 0.01 * rng.standard_normal((30000, dim))).astype(np.float32)
    q = (direction[None, :] + 0.01 * rng.standard_normal((40, dim))).astype(np.float32)
    want_ids, want_d, _ = oracle.exact_knn(base, q, 50, metric=1)
    ids, d = capi.knn_exact(base, q, 50, metric=1)
    check_knn(ids, d, want_ids, want_d, "adversarial")
    assert capi.knn_last_stats()["exact_scans"] > 0


def test_knn_device_api_and_merge(capi, oracle):
    """Base sharded in 3 parts (as on 3 GPUs), per-shard top-K merged by K4 == single-shot answer."""
    import torch
    from mysteryann_b200 import synth

    base, q, _ = synth.make_numpy(9000, 200, 1, 200, seed=4)
    K = 40
    want_ids, want_d, _ = oracle.exact_knn(base, q, K, metric=1)
    db, dq = torch.from_numpy(base).cuda(), torch.from_numpy(q).cuda()
    bounds = [0, 2500, 6100, 9000]
    part_ids = torch.empty((3, 200, K), dtype=torch.int32, device="cuda")
    part_d = torch.empty((3, 200, K), dtype=torch.float32, device="cuda")
    for g in range(3):
        shard = db[bounds[g]:bounds[g + 1]].contiguous()
        capi.knn_exact_device(shard, dq, K, part_ids[g], part_d[g], metric=1, id_base=bounds[g])
    out_ids = torch.empty((200, K), dtype=torch.int32, device="cuda")
    out_d = torch.empty((200, K), dtype=torch.float32, device="cuda")
    capi.knn_merge_device(part_ids, part_d, out_ids, out_d, metric=1)
    torch.cuda.synchronize()
    check_knn(out_ids.cpu().numpy().view(np.uint32), out_d.cpu().numpy(), want_ids, want_d, "merge")


@pytest.mark.parametrize("metric", (1, 0))
def test_knn_query_batches(capi, oracle, metric, monkeypatch):
    """More queries than one K2 batch: the batch loop (per-batch FP16 conversion, candidate lists, flag offsets) against the
    oracle - once with the batch size forced down to 4096 (three batches, the last one ragged), once with the default
    131072 crossed by 140000 queries."""
    from mysteryann_b200 import synth

    base, train, _ = synth.make_numpy(20000, 10000, 1, 64, seed=77)
    want_ids, want_d, _ = oracle.exact_knn(base, train, 10, metric=metric)
    monkeypatch.setenv("RG_KNN_QBATCH", "4096")
    ids, d = capi.knn_exact(base, train, 10, metric=metric)
    check_knn(ids, d, want_ids, want_d, f"q_batch=4096 metric={metric}")
    monkeypatch.delenv("RG_KNN_QBATCH")
    base, train, _ = synth.make_numpy(3000, 140000, 1, 64, seed=78)
    want_ids, want_d, _ = oracle.exact_knn(base, train, 10, metric=metric)
    ids, d = capi.knn_exact(base, train, 10, metric=metric)
    check_knn(ids, d, want_ids, want_d, f"nq=140000 metric={metric}")


@pytest.mark.parametrize("metric,optimistic", ((1, 1), (0, 1), (1, 0)))
def test_knn_million_row_base_sample(capi, oracle, metric, optimistic, monkeypatch):
    """1.2M base rows x 20000 queries on the GPU (block schedule far past 64K rows, sparse epilogue at depth, the
    optimistic threshold ranks r < k' in play), 256 sampled queries against the oracle; both threshold schedules."""
    import torch

    monkeypatch.setenv("RG_KNN_OPTIMISTIC", str(optimistic))
    rng = np.random.default_rng(123 + metric)
    n, nq, dim, K = 1_200_000, 20000, 64, 100
    base = rng.standard_normal((n, dim), dtype=np.float32)
    q = rng.standard_normal((nq, dim), dtype=np.float32) + 0.25
    db, dq = torch.from_numpy(base).cuda(), torch.from_numpy(q).cuda()
    ids = torch.empty((nq, K), dtype=torch.int32, device="cuda")
    d = torch.empty((nq, K), dtype=torch.float32, device="cuda")
    capi.knn_exact_device(db, dq, K, ids, d, metric=metric)
    st = capi.knn_last_stats()
    assert st["exact_scans"] <= 2 and st["second_pass"] <= nq // 100, st
    sample = rng.choice(nq, 256, replace=False)
    want_ids, want_d, _ = oracle.exact_knn(base, q[sample], K, metric=metric)
    check_knn(ids.cpu().numpy().view(np.uint32)[sample], d.cpu().numpy()[sample], want_ids, want_d,
              f"1.2M rows metric={metric} optimistic={optimistic}")
    capi.knn_release_scratch()


def test_knn_best_rows_first_takes_second_pass(capi, oracle):
    """Rows stored best-first: the optimistic schedule takes its threshold from the first block, nothing later beats it,
    the lists end with fewer than K survivors, the certificate fails - and the conservative second pass must return the
    exact answer without resorting to the FP32 scan."""
    rng = np.random.default_rng(2)
    dim, n, K = 32, 60_000, 50
    direction = rng.standard_normal(dim).astype(np.float32)
    # <q,b> decreases with the row index, ~1.5 % between ranks K and k' (well above the FP16 rounding bound of the certificate)
    scale = (4.0 * np.exp(-np.arange(n, dtype=np.float64) / 2000.0)).astype(np.float32)[:, None]
    base = (scale * direction[None, :] + 0.0002 * rng.standard_normal((n, dim))).astype(np.float32)
    q = (direction[None, :] + 0.01 * rng.standard_normal((64, dim))).astype(np.float32)
    want_ids, want_d, _ = oracle.exact_knn(base, q, K, metric=1)
    ids, d = capi.knn_exact(base, q, K, metric=1)
    check_knn(ids, d, want_ids, want_d, "best-first")
    st = capi.knn_last_stats()
    assert st["second_pass"] > 0 and st["exact_scans"] == 0, st


def test_knn_sharded_capi_world1(capi, oracle):
    """rg_knn_exact_sharded with world = 1 (no communicator): the slice is the whole query set."""
    import torch
    from mysteryann_b200 import synth

    base, q, _ = synth.make_numpy(6000, 300, 1, 200, seed=9)
    K = 20
    want_ids, want_d, _ = oracle.exact_knn(base, q, K, metric=1)
    db, dq = torch.from_numpy(base).cuda(), torch.from_numpy(q).cuda()
    assert capi.knn_sharded_slice(300, 0, 1) == (0, 300)
    assert capi.knn_sharded_slice(10, 1, 3) == (4, 7)
    ids = torch.empty((300, K), dtype=torch.int32, device="cuda")
    d = torch.empty((300, K), dtype=torch.float32, device="cuda")
    capi.knn_exact_sharded(db, 0, dq, K, ids, d, None, 0, 1, metric=1)
    check_knn(ids.cpu().numpy().view(np.uint32), d.cpu().numpy(), want_ids, want_d, "sharded world=1")
    with pytest.raises(capi.RoarGraphError):   # world > 1 needs a communicator
        capi.knn_exact_sharded(db, 0, dq, K, ids, d, None, 0, 2, metric=1)


@pytest.mark.parametrize("seg", (0, 256))
def test_knn_sharded_capi_two_gpus(capi, oracle, seg, monkeypatch):
    """Two base shards on two GPUs, one host thread per GPU (the layout of compute_groundtruth --devices 2): communicators
    from rg_nccl_comm_init_all, grouped ncclSend/ncclRecv exchange and K4 merge inside rg_knn_exact_sharded - once as a single
    chunk (one K2 call over all queries), once with 256-row segments (the chunked path of C4-size query sets).  Skipped on
    a one-GPU box."""
    import ctypes as C
    import threading

    import torch
    from mysteryann_b200 import synth

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    if seg:
        monkeypatch.setenv("RG_KNN_SHARD_SEG", str(seg))
    assert capi.lib().rg_nccl_version() > 0
    n, nq, dim, K = 30000, 1001, 200, 30
    base, q, _ = synth.make_numpy(n, nq, 1, dim, seed=12)
    want_ids, want_d, _ = oracle.exact_knn(base, q, K, metric=1)
    comms = (C.c_void_p * 2)()
    assert capi.lib().rg_nccl_comm_init_all(comms, 2, None) == 0, capi.lib().rg_last_error_string()
    bounds = [0, 17000, n]
    out, err = [None, None], [None, None]

    def work(r):
        try:
            dev = torch.device("cuda", r)
            with torch.cuda.device(dev):
                shard = torch.from_numpy(base[bounds[r]:bounds[r + 1]]).to(dev)
                dq = torch.from_numpy(q).to(dev)
                lo, hi = capi.knn_sharded_slice(nq, r, 2)
                ids = torch.empty((hi - lo, K), dtype=torch.int32, device=dev)
                d = torch.empty((hi - lo, K), dtype=torch.float32, device=dev)
                capi.knn_exact_sharded(shard, bounds[r], dq, K, ids, d, C.c_void_p(comms[r]), r, 2, metric=1)
                out[r] = (lo, hi, ids.cpu().numpy().view(np.uint32), d.cpu().numpy())
        except Exception as e:  # noqa: BLE001
            err[r] = e

    th = [threading.Thread(target=work, args=(r,)) for r in range(2)]
    [t.start() for t in th]
    [t.join(timeout=300) for t in th]
    assert err == [None, None], err
    for r in range(2):
        lo, hi, ids, d = out[r]
        check_knn(ids, d, want_ids[lo:hi], want_d[lo:hi], f"sharded rank {r}")
        capi.nccl_comm_destroy(C.c_void_p(comms[r]))


@pytest.mark.parametrize("world,base_shards", [(2, 1), (2, 2), (4, 2), (4, 1), (4, 4)])
def test_knn_grid_capi(capi, oracle, world, base_shards):
    """rg_knn_exact_grid: base_shards base shards x world / base_shards query groups over one communicator (threads, one per
    GPU).  Every rank holds shard rank % base_shards and ITS group's queries only, the exchange runs inside a group
    (peer = group * base_shards + p); base_shards = 1 is plain query sharding, base_shards = world the sharded call.
    Skipped when the box has fewer GPUs than ranks."""
    import ctypes as C
    import threading

    import torch
    from mysteryann_b200 import sharded_knn, synth

    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    n, nq, dim, K = 30011, 1003, 200, 30
    base, q, _ = synth.make_numpy(n, nq, 1, dim, seed=13)
    want_ids, want_d, _ = oracle.exact_knn(base, q, K, metric=1)
    comms = (C.c_void_p * world)()
    assert capi.lib().rg_nccl_comm_init_all(comms, world, None) == 0, capi.lib().rg_last_error_string()
    out, err = [None] * world, [None] * world

    def work(r):
        try:
            dev = torch.device("cuda", r)
            with torch.cuda.device(dev):
                (b0, b1), (g0, g1), (o0, o1) = sharded_knn.grid_layout(r, world, base_shards, n, nq)
                shard = torch.from_numpy(base[b0:b1]).to(dev)
                dq = torch.from_numpy(q[g0:g1]).to(dev)
                ids = torch.empty((o1 - o0, K), dtype=torch.int32, device=dev)
                d = torch.empty((o1 - o0, K), dtype=torch.float32, device=dev)
                capi.knn_exact_grid(shard, b0, dq, K, ids, d, C.c_void_p(comms[r]), r, world, base_shards, metric=1)
                out[r] = (o0, o1, ids.cpu().numpy().view(np.uint32), d.cpu().numpy())
        except Exception as e:  # noqa: BLE001
            err[r] = e

    th = [threading.Thread(target=work, args=(r,)) for r in range(world)]
    [t.start() for t in th]
    [t.join(timeout=300) for t in th]
    assert err == [None] * world, err
    covered = np.zeros(nq, bool)
    for r in range(world):
        lo, hi, ids, d = out[r]
        check_knn(ids, d, want_ids[lo:hi], want_d[lo:hi], f"grid {base_shards}x{world // base_shards} rank {r}")
        assert not covered[lo:hi].any()
        covered[lo:hi] = True
        capi.nccl_comm_destroy(C.c_void_p(comms[r]))
    assert covered.all()
    with pytest.raises(capi.RoarGraphError):   # base_shards must divide world
        capi.knn_exact_grid(torch.zeros((8, dim), device="cuda"), 0, torch.zeros((8, dim), device="cuda"), 4,
                            torch.zeros((8, 4), dtype=torch.int32, device="cuda"), torch.zeros((8, 4), device="cuda"),
                            None, 0, 4, 3, metric=1)
