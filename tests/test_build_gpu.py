"""GPU graph construction (rg_build_roargraph_device) against the CPU restatement of the reference's BuildRoarGraph:
same degree bounds and list invariants, same entry point, and - searched with the same GPU beam search - recall@10
within a small margin at every beam width (the GPU build applies the reference's pruning rules phase by phase to all
nodes at once, so like a multi-threaded reference build it is not edge-identical to the one-thread build)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def capi():
    from mysteryann_b200 import build, capi, hostlib

    build.build()
    hostlib.build()
    assert capi.device_count() > 0
    return capi


def recall(ids, gt, k):
    return float(np.mean([len(set(a[:k].tolist()) & set(b[:k].tolist())) / k for a, b in zip(ids, gt)]))


@pytest.mark.parametrize("metric,n,dim,M_sq,M,L_build", [(1, 20000, 200, 100, 35, 500), (0, 8000, 104, 40, 14, 60),
                                                        (1, 6000, 512, 64, 24, 100)])
def test_gpu_build_matches_cpu_build_quality(capi, metric, n, dim, M_sq, M, L_build, tmp_path):
    import torch
    from mysteryann_b200 import hostlib, io, synth

    base, train, test = synth.make_numpy(n, n, 2000, dim, seed=n + dim, normalize=(dim == 512))
    knn, _ = capi.knn_exact(base, train, M_sq, metric=metric)
    gt, _ = capi.knn_exact(base, test, 10, metric=metric)

    d_base = torch.from_numpy(base).cuda()
    d_knn = torch.from_numpy(knn.view(np.int32)).cuda()
    g = capi.Graph(d_base, d_knn, M_sq=M_sq, M_pjbp=M, L_pjpq=L_build, metric=metric)
    ep, off, adj = g.download()
    deg = np.diff(off.astype(np.int64))
    assert len(deg) == n and deg.max() == g.max_degree <= 2 * M and int(off[-1]) == g.nnz == len(adj)
    assert adj.max() < n
    for v in np.random.default_rng(0).integers(0, n, 400):       # list invariants: no self loop, no repeated id
        row = adj[off[v]:off[v + 1]]
        assert v not in row and len(set(row.tolist())) == len(row)
    assert deg.mean() > 0.5 * M, deg.mean()

    cpu_index = str(tmp_path / "cpu.index")
    hostlib.build_index(base, train, knn, cpu_index, metric=metric, M_sq=M_sq, M_pjbp=M, L_pjpq=L_build, threads=8)
    cep, coff, cadj = io.read_index(cpu_index)
    assert ep == cep                                              # same centroid-nearest entry point
    cdeg = np.diff(coff.astype(np.int64))
    assert abs(deg.mean() - cdeg.mean()) < 0.25 * cdeg.mean(), (deg.mean(), cdeg.mean())

    ix_gpu = capi.Index.from_graph(d_base, g, metric=metric)
    ix_cpu = capi.Index(d_base, coff, cadj, cep, metric=metric)
    ix_dl = capi.Index(d_base, off, adj, ep, metric=metric)      # the downloaded CSR describes the same graph
    gaps = []
    for L in (10, 20, 50, 100, 200):
        a = ix_gpu.search(test, 10, L)
        b = ix_cpu.search(test, 10, L)
        c = ix_dl.search(test, 10, L)
        assert (a["ids"] == c["ids"]).all() and (a["cmps"] == c["cmps"]).all()
        ra, rb = recall(a["ids"], gt, 10), recall(b["ids"], gt, 10)
        print(f"metric={metric} n={n} L={L}: recall gpu-built {ra:.4f} cpu-built {rb:.4f} "
              f"cmps {a['cmps'].mean():.0f}/{b['cmps'].mean():.0f}")
        gaps.append((L, ra, rb))
    print("degrees gpu/cpu", deg.mean(), cdeg.mean(), deg.max(), cdeg.max(), "phases", g.phase_seconds)
    # Both builds are order-dependent (threads on the CPU, waves of n/256 nodes + atomics on the GPU).  Measured at C1
    # against two 16-thread builds of the compiled reference (profiles/r02_build_quality_c1*.txt): the GPU build ends
    # <= 0.002 below the reference curve at every L_pq of the sweep, the two reference builds differ by up to 0.0018.
    # Here: 2000 test queries (~0.003 sampling noise on a recall DIFFERENCE of two graphs over the same queries), graphs
    # of 6-20K nodes where a wave is 0.6-2 % of the nodes; the small L2 graph (M_pjbp = 14, L_pjpq = 60) sits 0.002-0.009
    # below its CPU build (mean -0.005), the other two within +-0.001 on average.
    # One run in seven of this session put the small L2 graph at -0.016 (L=10), -0.008, -0.006, -0.004, -0.002 (mean -0.007):
    # the margins cover that observed worst case; the C1-scale comparison against the compiled reference is the claim.
    for L, ra, rb in gaps:
        assert ra >= rb - (0.02 if L <= 20 else 0.008), gaps
    assert np.mean([ra - rb for _, ra, rb in gaps]) > -0.009, gaps
    for x in (ix_gpu, ix_cpu, ix_dl):
        x.close()
    g.close()


@pytest.mark.parametrize("metric,n,dim,M_sq,M", [(1, 20000, 200, 100, 35), (0, 6000, 104, 40, 14), (1, 3000, 512, 64, 24)])
def test_projection_lists_equal_host(capi, metric, n, dim, M_sq, M):
    """P1 of the build (pivot projection + PruneBiSearchBaseGetBase, src/index_bipartite.cpp:1059-1097, 1612-1694) is
    deterministic given the kNN file: the GPU prune kernel's list of EVERY training query must equal the host restatement's
    (which is byte-identical to the reference at -T 1) - same ids, same order, same length."""
    import torch
    from mysteryann_b200 import hostlib, synth

    base, train, _ = synth.make_numpy(n, n, 10, dim, seed=3 * n + dim, normalize=(dim == 512))
    knn, _ = capi.knn_exact(base, train, M_sq, metric=metric)
    want = hostlib.projection_lists(base, knn, metric=metric, M_sq=M_sq, M_pjbp=M)
    got = capi.projection_lists_device(torch.from_numpy(base).cuda(), torch.from_numpy(knn.view(np.int32)).cuda(), M_sq=M_sq,
                                       M_pjbp=M, metric=metric).cpu().numpy().view(np.uint32)
    assert (got[:, 0] == want[:, 0]).all(), np.argwhere(got[:, 0] != want[:, 0])[:5]
    assert want[:, 0].max() <= M and want[:, 0].min() >= 1
    mask = np.arange(1, M + 1)[None, :] <= want[:, :1]
    bad = np.argwhere((got[:, 1:] != want[:, 1:]) & mask)
    assert len(bad) == 0, (len(bad), bad[:5], got[bad[0][0]], want[bad[0][0]])
