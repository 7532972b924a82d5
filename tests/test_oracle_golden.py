"""CPU tests: the plain-C oracle against (a) the committed golden vectors produced by the compiled
reference (tests/golden/make_golden.py) and (b) the compiled reference itself where it is available."""
import numpy as np
import pytest

from conftest import CASES, GOLDEN, load_case

DIMS = (1, 2, 3, 4, 5, 7, 8, 9, 12, 15, 16, 17, 23, 24, 31, 32, 40, 48, 100, 128, 200, 203, 512)


def bits(x):
    return np.ascontiguousarray(x, np.float32).view(np.uint32)


def test_distance_golden(oracle):
    z = np.load(GOLDEN + "/distance.npz")
    for d in DIMS:
        a, b = z[f"a_{d}"].astype(np.float32), z[f"b_{d}"].astype(np.float32)
        assert (bits(oracle.distance_batch(0, a, b)) == bits(z[f"l2_{d}"])).all(), d
        assert (bits(oracle.distance_batch(1, a, b)) == bits(z[f"ip_{d}"])).all(), d
        # COSINE uses DistanceInnerProduct (src/index.cpp:14-17)
        assert (bits(oracle.distance_batch(4, a, b)) == bits(z[f"ip_{d}"])).all(), d


def test_pool_golden(oracle):
    z = np.load(GOLDEN + "/pool.npz")
    for i in range(5):
        oi, od, of, pop = oracle.pool_script(int(z[f"cap_{i}"]), z[f"kind_{i}"], z[f"ids_{i}"], z[f"dists_{i}"])
        assert (oi == z[f"out_ids_{i}"]).all() and (bits(od) == bits(z[f"out_dists_{i}"])).all()
        assert (of == z[f"out_flags_{i}"]).all() and (pop == z[f"pop_{i}"]).all()


@pytest.mark.parametrize("name", CASES)
def test_search_golden(oracle, name):
    c = load_case(name)
    for L in c["Ls"]:
        r = oracle.search(c["base"], c["offsets"], c["adj"], c["ep"], c["test"], 10, int(L), metric=c["metric"])
        assert r["rc"] == 0
        assert (r["ids"] == c[f"ids_{L}"]).all()
        assert (bits(r["dists"]) == bits(c[f"dists_{L}"])).all()
        assert (r["cmps"] == c[f"cmps_{L}"]).all() and (r["hops"] == c[f"hops_{L}"]).all()


def test_search_not_enough_results(oracle):
    # a 3-node graph cannot return k=10 (src/index_bipartite.cpp:2408-2412)
    base = np.eye(3, 8, dtype=np.float32)
    off = np.array([0, 2, 3, 4], np.uint64)
    adj = np.array([1, 2, 0, 0], np.uint32)
    r = oracle.search(base, off, adj, 0, base[:1], 10, 16, metric=1)
    assert r["rc"] == 2
    r = oracle.search(base, off, adj, 0, base[:1], 3, 16, metric=1)
    assert r["rc"] == 0 and sorted(r["ids"][0]) == [0, 1, 2] and r["cmps"][0] == 3  # ep re-scored once


@pytest.mark.parametrize("metric", (0, 1))
def test_knn_matches_fp64(oracle, metric):
    rng = np.random.default_rng(3)
    base = rng.standard_normal((3000, 40)).astype(np.float32)
    q = rng.standard_normal((50, 40)).astype(np.float32)
    ids, dists, _ = oracle.exact_knn(base, q, 20, metric=metric, part_size=1000)
    ids1, dists1, _ = oracle.exact_knn(base, q, 20, metric=metric)  # single part: same answer
    assert (ids == ids1).all() and (bits(dists) == bits(dists1)).all()
    s = q.astype(np.float64) @ base.astype(np.float64).T
    if metric == 0:
        s = (q.astype(np.float64) ** 2).sum(1)[:, None] + (base.astype(np.float64) ** 2).sum(1)[None] - 2 * s
        want = np.sort(s, axis=1)[:, :20]
    else:
        want = -np.sort(-s, axis=1)[:, :20]   # stored as +ip, descending (compute_groundtruth.cpp:438-441)
    assert np.allclose(dists, want, rtol=1e-5, atol=1e-5)
    got_set = [set(r) for r in ids]
    ref_ids = np.argsort(s if metric == 0 else -s, axis=1)[:, :20]
    assert np.mean([len(g & set(r)) / 20 for g, r in zip(got_set, ref_ids)]) > 0.999


@pytest.mark.parametrize("metric", (0, 1))
def test_build_search_restatement(oracle, metric):
    """rgo_search_projection_internal (SearchProjectionGraphInternal, src/index_bipartite.cpp:1279-1350) against the pinned
    search restatement: for a target row that nobody links to, the two beam searches differ only in that the build search
    marks the entry point visited (:1311), which cannot change the pool (a re-scored entry point is dropped as a duplicate,
    neighbor.h:161) - so the number of expanded nodes equals the hops of rgo_search_roargraph with that row as the query,
    the first expanded node is the entry point, no node is expanded twice and the target never shows up."""
    rng = np.random.default_rng(41 + metric)
    n, dim, L = 3000, 24, 40
    base = rng.standard_normal((n, dim)).astype(np.float32)
    deg = rng.integers(3, 14, n)
    off = np.zeros(n + 1, np.uint64)
    np.cumsum(deg, out=off[1:])
    node_lo, count = 100, 60
    adj = rng.integers(0, n - count, int(off[-1])).astype(np.uint32)
    adj[adj >= node_lo] += count                    # nobody links to the targets [node_lo, node_lo + count)
    ep = 7
    ids, dists, cnt = oracle.search_expanded(base, off, adj, ep, node_lo, count, L, 4 * L, metric=metric)
    ref = oracle.search(base, off, adj, ep, base[node_lo:node_lo + count], 10, L, metric=metric, threads=2)
    assert (cnt == ref["hops"]).all()
    assert (ids[:, 0] == ep).all()
    for t in range(count):
        row = ids[t, :cnt[t]]
        assert len(set(row.tolist())) == len(row) and node_lo + t not in row
        assert dists[t, 0] == oracle.distance_batch(metric, base[ep:ep + 1], base[node_lo + t:node_lo + t + 1])[0]
    # a target that IS linked is skipped as a neighbour (:1328): it is never expanded either
    adj2 = adj.copy()
    adj2[off[ep]:off[ep + 1]] = node_lo
    ids2, _, cnt2 = oracle.search_expanded(base, off, adj2, ep, node_lo, 1, L, 4 * L, metric=metric)
    assert node_lo not in ids2[0, :cnt2[0]] and cnt2[0] == 1   # the entry point's only neighbour is the target itself


def test_recall(oracle):
    gt = np.array([[1, 2, 3, 9], [4, 5, 6, 9]], np.uint32)
    res = np.array([[3, 1, 7], [8, 8, 8]], np.uint32)
    assert oracle.recall(res, gt, 3) == pytest.approx(2 / 6)


# ---- live differential checks against the compiled reference (only where oracle/_ref exists) ----
def test_distance_vs_ref(oracle, ref):
    rng = np.random.default_rng(9)
    for d in (8, 24, 200, 203, 512, 960):
        a = (rng.standard_normal((4000, d)) * 5).astype(np.float32)
        b = rng.standard_normal((4000, d)).astype(np.float32)
        for m in (0, 1):
            assert (bits(oracle.distance_batch(m, a, b)) == bits(ref.distance_batch(m, a, b))).all()


def test_pool_vs_ref(oracle, ref):
    rng = np.random.default_rng(10)
    for cap in (1, 2, 7, 64, 500):
        nops = 4000
        kind = (rng.random(nops) < 0.2).astype(np.uint8)
        ids = rng.integers(0, 900, nops).astype(np.uint32)
        dists = np.round(rng.standard_normal(nops), 1).astype(np.float32)
        a, b = oracle.pool_script(cap, kind, ids, dists), ref.pool_script(cap, kind, ids, dists)
        assert all((x == y).all() for x, y in zip(a, b))


@pytest.mark.parametrize("metric,dim,n", [(1, 200, 6000), (0, 48, 4000), (1, 104, 3000)])
def test_search_vs_ref_live(oracle, ref, tmp_path, metric, dim, n):
    """Whole searches, oracle vs the compiled reference (IndexBipartite::SearchRoarGraph through its own loaders), on fresh
    seeded random graphs incl. zero-degree nodes, duplicate neighbours and heavy ties: ids / dists / cmps / hops bit-equal."""
    from mysteryann_b200 import io

    rng = np.random.default_rng(1000 + dim)
    base = np.round(rng.standard_normal((n, dim)) * 2).astype(np.float32) / 2   # half-integers: many exact ties
    queries = rng.standard_normal((150, dim)).astype(np.float32)
    deg = rng.integers(0, 40, n)
    off = np.zeros(n + 1, np.uint64)
    np.cumsum(deg, out=off[1:])
    adj = rng.integers(0, n, int(off[-1])).astype(np.uint32)
    ep = int(np.argmax(deg))
    io.write_fbin(tmp_path / "b.fbin", base)
    io.write_index(tmp_path / "g.index", ep, off, adj)
    h = ref.open(str(tmp_path / "b.fbin"), str(tmp_path / "g.index"), metric=metric, threads=4)
    try:
        for L, k in ((1, 1), (10, 10), (33, 10), (200, 100)):
            want = ref.search(h, queries, k, L, threads=4)
            got = oracle.search(base, off, adj, ep, queries, k, L, metric=metric)
            for key in ("ids", "cmps", "hops"):
                assert (got[key] == want[key]).all(), (key, L)
            assert (bits(got["dists"]) == bits(want["dists"])).all(), L
    finally:
        ref.close(h)
