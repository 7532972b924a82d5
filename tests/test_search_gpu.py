"""GPU parity tests of K1 (beam search) through the C ABI: bit-exact ids / dists / cmps / hops against
(a) the golden vectors produced by the compiled reference and (b) the CPU oracle on seeded random graphs."""
import numpy as np
import pytest

from conftest import CASES, load_case

pytestmark = pytest.mark.gpu


def bits(x):
    return np.ascontiguousarray(x, np.float32).view(np.uint32)


def report(tag, got, want):
    """assert equality with a useful message"""
    for key in ("hops", "cmps", "ids"):
        g, w = got[key], want[key]
        if not (g == w).all():
            bad = np.argwhere(g != w)
            q = int(bad[0][0])
            raise AssertionError(f"{tag}: {key} differs in {len(set(bad[:, 0]))} queries; first q={q}: got "
                                 f"{got[key][q]} want {want[key][q]} | hops {got['hops'][q]}/{want['hops'][q]} "
                                 f"cmps {got['cmps'][q]}/{want['cmps'][q]}")
    gb, wb = bits(got["dists"]), bits(want["dists"])
    if not (gb == wb).all():
        bad = np.argwhere(gb != wb)
        q, j = bad[0]
        raise AssertionError(f"{tag}: dists differ at {len(bad)} places; first q={q} j={j}: "
                             f"{got['dists'][q, j]!r} vs {want['dists'][q, j]!r}")


@pytest.fixture(scope="module")
def capi():
    from mysteryann_b200 import build, capi

    build.build()
    assert capi.device_count() > 0, "no CUDA device"
    return capi


# (gather, warps per query, visited-hash space, L2 hints, adjacency prefetch): cp.async / TMA bulk gathers, 1..8 warps per
# query, shared / global hash, evict_first rows + persisting hash window, speculative adjacency prefetch (bit 2 = 4: early
# issue of the next hop's filter and first gather before the merge)
# hash space: 1 shared memory, 2 / 3 global atomicCAS tables (32-bit keys / 16-bit quotient entries), 4 / 5 global buckets
# without atomics (16-bit entries where the id range allows / 32-bit ids), 0 auto (= 4)
CONFIGS = ((2, 4, 2, 0, 0), (1, 4, 3, 0, 0), (2, 1, 1, 0, 0), (2, 2, 3, 3, 3), (1, 3, 1, 3, 3), (2, 8, 5, 1, 2),
           (2, 2, 0, 2, 1), (2, 2, 2, 0, 0), (2, 3, 4, 3, 1), (2, 1, 4, 3, 3), (2, 2, 5, 3, 3), (2, 4, 4, 0, 0),
           (2, 8, 0, 3, 3), (2, 2, 4, 3, 3, 1), (2, 4, 2, 3, 3, 2), (1, 3, 3, 3, 1, 2), (2, 2, 4, 3, 3, 3), (2, 5, 2, 3, 3, 3), (2, 2, 4, 3, 7), (2, 3, 5, 3, 5), (2, 8, 4, 0, 7), (2, 1, 4, 3, 7), (2, 4, 0, 3, 7))


def configure(ix, cfg, **kw):
    gather, warps, space, l2, pf = cfg[:5]
    bm = cfg[5] if len(cfg) > 5 else 0   # batch_mode: 0 auto, 1 per-warp gather lists, 2 one list per query, 3 per-warp + stealing
    ix.configure(gather=gather, warps_per_query=warps, hash_space=space, l2_hint=l2, adj_prefetch=pf, batch_mode=bm, **kw)


@pytest.mark.parametrize("cfg", CONFIGS)
@pytest.mark.parametrize("name", CASES)
def test_golden(capi, name, cfg):
    c = load_case(name)
    ix = capi.Index(c["base"], c["offsets"], c["adj"], c["ep"], metric=c["metric"])
    configure(ix, cfg)
    gather, warps, space, l2, pf = cfg[:5]
    for L in c["Ls"]:
        L = int(L)
        got = ix.search(c["test"], 10, L)
        assert got["rc"] == 0
        want = {k: c[f"{k}_{L}"] for k in ("ids", "dists", "cmps", "hops")}
        report(f"{name} L={L} gather={gather} warps={warps} space={space} l2={l2} pf={pf}", got, want)
    ix.close()


def random_graph(rng, n, dmin, dmax, zero_frac=0.0):
    deg = rng.integers(dmin, dmax + 1, n)
    if zero_frac:
        deg[rng.random(n) < zero_frac] = 0
    off = np.zeros(n + 1, np.uint64)
    np.cumsum(deg, out=off[1:])
    adj = rng.integers(0, n, int(off[-1])).astype(np.uint32)   # may contain duplicates and self loops
    return off, adj


@pytest.mark.parametrize("metric", (0, 1))
@pytest.mark.parametrize("dim,dmin,dmax", [(200, 1, 70), (512, 8, 40), (64, 0, 95), (8, 3, 12), (104, 30, 130)])
def test_random_graph_vs_oracle(capi, oracle, metric, dim, dmin, dmax):
    rng = np.random.default_rng(dim * 7 + metric)
    n, nq = 20000, 300
    base = rng.standard_normal((n, dim)).astype(np.float32)
    q = rng.standard_normal((nq, dim)).astype(np.float32)
    off, adj = random_graph(rng, n, dmin, dmax, zero_frac=0.01)
    ep = int(rng.integers(0, n))
    if off[ep + 1] == off[ep]:
        ep = int(np.argmax(np.diff(off)))
    ix = capi.Index(base, off, adj, ep, metric=metric)
    for cfg in CONFIGS[:5] + CONFIGS[8:]:  # incl. every bucket configuration
        configure(ix, cfg)
        gather, warps, space = cfg[:3]
        for L, k in ((1, 1), (10, 10), (37, 10), (64, 20), (200, 100)):
            want = oracle.search(base, off, adj, ep, q, k, L, metric=metric)
            got = ix.search(q, k, L)
            assert got["rc"] == want["rc"] == 0
            report(f"dim={dim} metric={metric} gather={gather} warps={warps} space={space} L={L}", got, want)
    ix.close()


@pytest.mark.parametrize("metric", (0, 1))
@pytest.mark.parametrize("warps,space,hash_log2,pf", [(0, 0, 0, 3), (2, 4, 0, 3), (4, 5, 0, 3), (2, 2, 0, 3), (3, 3, 0, 3), (1, 4, 0, 3),
                                                      (2, 4, 9, 3), (8, 0, 0, 3), (2, 4, 0, 7), (3, 5, 0, 7), (2, 4, 9, 7)])
def test_build_search_expanded_vs_oracle(capi, oracle, metric, warps, space, hash_log2, pf):
    """The connectivity-enhancement searches (SearchProjectionGraphInternal, src/index_bipartite.cpp:1279-1350; K1's build
    variant): expanded nodes of base rows used as queries, in expansion order, ids and distance bit patterns, against the
    oracle - every visited-set flavour, incl. a table small enough to send queries through the big-table pass."""
    rng = np.random.default_rng(31 + metric)
    n, dim = 20000, 64
    base = rng.standard_normal((n, dim)).astype(np.float32)
    off, adj = random_graph(rng, n, 4, 40, zero_frac=0.01)   # duplicates and self loops included
    ep = int(np.argmax(np.diff(off)))
    ix = capi.Index(base, off, adj, ep, metric=metric)
    ix.configure(warps_per_query=warps, hash_space=space, hash_log2=hash_log2, adj_prefetch=pf)
    node_lo, count = 1234, 400
    for L, cap in ((20, 32), (100, 100), (300, 64)):
        want_ids, want_d, want_n = oracle.search_expanded(base, off, adj, ep, node_lo, count, L, cap, metric=metric)
        got_ids, got_d, got_n = ix.search_expanded(node_lo, count, L, cap)
        assert (got_n == want_n).all(), (L, np.argwhere(got_n != want_n)[:5].ravel(), got_n[:8], want_n[:8])
        mask = np.arange(cap)[None, :] < want_n[:, None]
        assert (got_ids[mask] == want_ids[mask]).all(), f"L={L} warps={warps} space={space}"
        assert (got_d.view(np.uint32)[mask] == want_d.view(np.uint32)[mask]).all(), f"L={L} warps={warps} space={space}"
    if hash_log2:
        assert ix.last_overflow > 0
    ix.close()


def test_visited_overflow_takes_exact_fallback(capi, oracle):
    """A tiny visited hash forces most queries through the big-table pass; results stay exact."""
    rng = np.random.default_rng(5)
    n, dim = 30000, 40
    base = rng.standard_normal((n, dim)).astype(np.float32)
    q = rng.standard_normal((200, dim)).astype(np.float32)
    off, adj = random_graph(rng, n, 20, 60)
    ix = capi.Index(base, off, adj, 3, metric=1)
    want = oracle.search(base, off, adj, 3, q, 10, 50, metric=1)
    assert want["cmps"].max() > 256
    for space in (1, 2, 3, 4, 5):
        ix.configure(hash_log2=8, hash_space=space)
        report(f"overflow space={space}", ix.search(q, 10, 50), want)
        assert ix.last_overflow > 100
    ix.configure(hash_log2=16, hash_space=1)   # > 15: too big for shared memory, global tables are used anyway
    report("global-primary", ix.search(q, 10, 50), want)
    ix.close()


def test_hash16_displacement_exhausted(capi, oracle):
    """16-bit quotient visited set on a 2^20-id range with a 1024-slot table: 10 remainder bits leave 6 displacement bits,
    so long probe runs exhaust the field before the load limit and those queries take the exact big-table pass."""
    rng = np.random.default_rng(11)
    n, dim = 1 << 20, 8
    base = rng.standard_normal((n, dim)).astype(np.float32)
    q = rng.standard_normal((300, dim)).astype(np.float32)
    off, adj = random_graph(rng, n, 4, 12)
    ix = capi.Index(base, off, adj, 5, metric=0)
    want = oracle.search(base, off, adj, 5, q, 10, 80, metric=0)
    for hl, lo, hi in ((10, 1, 300), (12, 0, 299)):
        ix.configure(hash_log2=hl, hash_space=3)
        report(f"hash16 hl={hl}", ix.search(q, 10, 80), want)
        assert lo <= ix.last_overflow <= hi
    ix.close()


def test_bucket16_displacement_exhausted(capi, oracle):
    """Bucketed 16-bit visited set on a 2^20-id range with 128 buckets per query: 13 remainder bits leave 3 displacement bits
    (a probe window of 7 buckets).  Ids whose window is full go to the per-warp exception list in shared memory (and are
    found there again); when that list is full too, or the table reaches 90 % load, the query takes the exact big-table pass.
    With 1024 buckets nothing of the kind happens.  Eight warps split 128 buckets into 16-bucket ranges (wrap-around probing)."""
    rng = np.random.default_rng(12)
    n, dim = 1 << 20, 8
    base = rng.standard_normal((n, dim)).astype(np.float32)
    q = rng.standard_normal((300, dim)).astype(np.float32)
    off, adj = random_graph(rng, n, 4, 12)
    ix = capi.Index(base, off, adj, 5, metric=0)
    want = oracle.search(base, off, adj, 5, q, 10, 80, metric=0)
    seen_exceptions = 0
    for hl, warps, tight in ((10, 2, True), (10, 8, True), (10, 1, True), (13, 2, False), (13, 1, False)):
        ix.configure(hash_log2=hl, hash_space=4, warps_per_query=warps)
        report(f"bucket16 hl={hl} warps={warps}", ix.search(q, 10, 80), want)
        if tight:
            seen_exceptions += ix.last_exceptions
        else:
            assert ix.last_overflow == 0 and ix.last_exceptions == 0
    assert seen_exceptions > 0, "the exception list was never exercised"
    ix.close()


def test_large_L_and_ties(capi, oracle):
    """Large beam widths (pool in shared memory up to L=2000) and heavy distance ties (grid-valued vectors)."""
    rng = np.random.default_rng(8)
    n, dim = 8000, 16
    base = rng.integers(-2, 3, (n, dim)).astype(np.float32)     # many exact ties, incl. duplicated rows
    q = rng.integers(-2, 3, (100, dim)).astype(np.float32)
    off, adj = random_graph(rng, n, 10, 30)
    for metric in (0, 1):
        ix = capi.Index(base, off, adj, 1, metric=metric)
        for L in (10, 100, 500, 2000):
            want = oracle.search(base, off, adj, 1, q, 10, L, metric=metric)
            for warps in (4, 1):
                ix.configure(warps_per_query=warps)
                report(f"ties metric={metric} L={L} warps={warps}", ix.search(q, 10, L), want)
        ix.close()


def test_not_enough_results(capi):
    base = np.eye(3, 8, dtype=np.float32)
    off = np.array([0, 2, 3, 4], np.uint64)
    adj = np.array([1, 2, 0, 0], np.uint32)
    ix = capi.Index(base, off, adj, 0, metric=1)
    r = ix.search(base[:1], 10, 16)
    assert r["rc"] == capi.RG_ERR_NOT_ENOUGH_RESULTS and (r["ids"] == 0xFFFFFFFF).all()
    assert b"not enough results" in capi.lib().rg_last_error_string()
    r = ix.search(base[:1], 3, 16)
    assert r["rc"] == 0 and sorted(r["ids"][0]) == [0, 1, 2] and r["cmps"][0] == 3 and r["hops"][0] == 3
    with pytest.raises(capi.RoarGraphError):   # k > L (tests/test_search_roargraph.cpp:192-195)
        ix.search(base[:1], 10, 5)
    ix.close()


def test_device_api_and_empty_batch(capi, oracle):
    import torch

    c = load_case("ip_d200")
    ix = capi.Index(torch.from_numpy(c["base"]).cuda(), c["offsets"], c["adj"], c["ep"], metric=1)
    dq = torch.from_numpy(c["test"]).cuda()
    nq, k, L = dq.shape[0], 10, 32
    ids = torch.empty((nq, k), dtype=torch.int32, device="cuda")
    dists = torch.empty((nq, k), dtype=torch.float32, device="cuda")
    cmps = torch.empty(nq, dtype=torch.int32, device="cuda")
    hops = torch.empty(nq, dtype=torch.int32, device="cuda")
    status = torch.full((2,), 7, dtype=torch.int32, device="cuda")
    ix.search_device(dq, k, L, ids, dists, cmps, hops, status, stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    got = dict(ids=ids.cpu().numpy().view(np.uint32), dists=dists.cpu().numpy(),
               cmps=cmps.cpu().numpy().view(np.uint32), hops=hops.cpu().numpy().view(np.uint32))
    report("device api", got, {k_: c[f"{k_}_{L}"] for k_ in ("ids", "dists", "cmps", "hops")})
    assert status.cpu().tolist() == [0, 0]
    assert ix.launches >= 3
    assert ix.search(np.zeros((0, 200), np.float32), 10, 32)["ids"].shape == (0, 10)
    ix.close()


def test_pinned_host_buffers_zero_copy(capi):
    """rg_search_batch on page-locked caller buffers (the kernel reads queries / writes results through the mapped
    pointers), with and without the optional cmps/hops outputs, against the staged path and the golden vectors."""
    import torch

    c = load_case("ip_d200")
    ix = capi.Index(c["base"], c["offsets"], c["adj"], c["ep"], metric=1)
    hq = torch.from_numpy(c["test"]).pin_memory()
    nq, k = hq.shape[0], 10
    for L in (10, 100):
        want = {k_: c[f"{k_}_{L}"] for k_ in ("ids", "dists", "cmps", "hops")}
        for zero_copy in (1, 0):
            ix.set_option("zero_copy", zero_copy)
            ids = torch.full((nq, k), -1, dtype=torch.int32).pin_memory()
            dists = torch.full((nq, k), 7.0, dtype=torch.float32).pin_memory()
            cmps = torch.zeros(nq, dtype=torch.int32).pin_memory()
            hops = torch.zeros(nq, dtype=torch.int32).pin_memory()
            ix.search_raw(hq.data_ptr(), nq, k, L, ids.data_ptr(), dists.data_ptr(), cmps.data_ptr(), hops.data_ptr())
            got = dict(ids=ids.numpy().view(np.uint32), dists=dists.numpy(), cmps=cmps.numpy().view(np.uint32),
                       hops=hops.numpy().view(np.uint32))
            report(f"pinned zero_copy={zero_copy} L={L}", got, want)
            ids.fill_(-1)
            ix.search_raw(hq.data_ptr(), nq, k, L, ids.data_ptr(), dists.data_ptr())  # no stats requested
            assert (ids.numpy().view(np.uint32) == want["ids"]).all()
    # mixed: pinned queries, pageable results -> staged path, same answer
    ix.set_option("zero_copy", 1)
    r_ids = np.empty((nq, k), np.uint32)
    r_d = np.empty((nq, k), np.float32)
    ix.search_raw(hq.data_ptr(), nq, k, 10, r_ids.ctypes.data, r_d.ctypes.data)
    assert (r_ids == c["ids_10"]).all() and (bits(r_d) == bits(c["dists_10"])).all()
    with pytest.raises(capi.RoarGraphError):
        ix.set_option("zero_copy", 2)
    ix.close()


def test_auto_warps_rule_large_batch(capi, oracle):
    """With enough queries to fill the GPU and L_pq in ~[80, 180] the library switches to four warps per query so that the
    visited-hash slabs of the resident CTAs fit the persisting part of L2 (rg_search.cu, search_device_impl); the
    results must not change.  3000 queries at L_pq = 100 take that branch on a B200, L_pq = 40 and 300 do not."""
    rng = np.random.default_rng(21)
    n, dim, nq = 30000, 200, 3000
    base = rng.standard_normal((n, dim)).astype(np.float32)
    off, adj = random_graph(rng, n, 8, 40)
    q = rng.standard_normal((nq, dim)).astype(np.float32)
    ix = capi.Index(base, off, adj, 7, metric=1)
    for L in (100, 40, 300):
        want = oracle.search(base, off, adj, 7, q, 10, L, metric=1)
        report(f"auto warps L={L}", ix.search(q, 10, L), want)
    ix.close()
