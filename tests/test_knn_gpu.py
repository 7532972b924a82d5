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
