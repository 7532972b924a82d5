"""GPU tests of the three drop-in CLI drivers (mysteryann_b200/host/apps): compute_groundtruth, test_build_roargraph
(--gpu_build) and test_search_roargraph, run as a user of the reference would run them, with the CPU oracle as checker.
Reference drivers: thirdparty/DiskANN/tests/utils/compute_groundtruth.cpp, tests/test_build_roargraph.cpp,
tests/test_search_roargraph.cpp."""
import os
import subprocess

import numpy as np
import pytest

from test_knn_gpu import check_knn

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def bins():
    from mysteryann_b200 import build, hostlib

    build.build()
    hostlib.build()
    return hostlib.BIN_DIR


def run(cmd):
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, f"{' '.join(cmd)}\n{p.stdout[-2000:]}\n{p.stderr[-2000:]}"
    return p.stdout


def pad8(x):
    from mysteryann_b200 import io

    return io.pad_rows(x, 8)


@pytest.mark.parametrize("dist_fn,metric", [("mips", 1), ("l2", 0)])
def test_compute_groundtruth_parts_and_format(bins, oracle, tmp_path, dist_fn, metric):
    """3 base parts (--part_size), unpadded D=100 rows, K=20: file = i32 n, i32 K, u32 ids, f32 dists (IP as +ip)."""
    from mysteryann_b200 import io

    rng = np.random.default_rng(7)
    base = rng.standard_normal((6000, 100)).astype(np.float32)
    q = rng.standard_normal((150, 100)).astype(np.float32)
    io.write_fbin(tmp_path / "base.fbin", base)
    io.write_fbin(tmp_path / "q.fbin", q)
    out = run([os.path.join(bins, "compute_groundtruth"), "--data_type", "float", "--dist_fn", dist_fn, "--base_file",
               str(tmp_path / "base.fbin"), "--query_file", str(tmp_path / "q.fbin"), "--gt_file", str(tmp_path / "gt.bin"),
               "--K", "20", "--part_size", "2500"])
    assert "Number of parts: 3" in out
    raw = np.fromfile(tmp_path / "gt.bin", dtype=np.uint8)
    assert raw.size == 8 + 150 * 20 * 8
    assert tuple(raw[:8].view(np.int32)) == (150, 20)
    ids, dists = io.read_ibin(tmp_path / "gt.bin")
    want_ids, want_d, _ = oracle.exact_knn(pad8(base), pad8(q), 20, metric=metric)
    check_knn(ids, dists, want_ids, want_d, f"cli {dist_fn}")


def test_compute_groundtruth_two_gpus_nccl(bins, oracle, tmp_path):
    """--devices 2 without --part_size: one base shard per GPU, rg_knn_exact_sharded_host from one thread per GPU (NCCL
    exchange + K4 merge on the devices).  Skipped on a one-GPU box."""
    import torch

    from mysteryann_b200 import io

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    rng = np.random.default_rng(31)
    base = rng.standard_normal((9001, 96)).astype(np.float32)
    q = rng.standard_normal((333, 96)).astype(np.float32)
    io.write_fbin(tmp_path / "base.fbin", base)
    io.write_fbin(tmp_path / "q.fbin", q)
    out = run([os.path.join(bins, "compute_groundtruth"), "--data_type", "float", "--dist_fn", "l2", "--base_file",
               str(tmp_path / "base.fbin"), "--query_file", str(tmp_path / "q.fbin"), "--gt_file", str(tmp_path / "gt.bin"),
               "--K", "25", "--devices", "2"])
    assert "Base sharded over 2 GPUs x 1 query group" in out
    ids, dists = io.read_ibin(tmp_path / "gt.bin")
    want_ids, want_d, _ = oracle.exact_knn(base, q, 25, metric=0)
    check_knn(ids, dists, want_ids, want_d, "cli 2 gpus")
    # grid layouts (rg_knn_exact_grid_host): base shards x query groups over the same communicator
    layouts = [(2, 1)] + ([(4, 2), (4, 1), (4, 4)] if torch.cuda.device_count() >= 4 else [])
    for ndev, bs in layouts:
        out = run([os.path.join(bins, "compute_groundtruth"), "--data_type", "float", "--dist_fn", "l2", "--base_file",
                   str(tmp_path / "base.fbin"), "--query_file", str(tmp_path / "q.fbin"), "--gt_file", str(tmp_path / "gt2.bin"),
                   "--K", "25", "--devices", str(ndev), "--base_shards", str(bs)])
        assert f"Base sharded over {bs} GPUs x {ndev // bs} query group" in out
        ids, dists = io.read_ibin(tmp_path / "gt2.bin")
        check_knn(ids, dists, want_ids, want_d, f"cli {ndev} gpus, {bs} base shards")


def test_compute_groundtruth_cosine_and_uint8(bins, tmp_path):
    """cosine = L2 on normalised rows (compute_groundtruth.cpp:146-175); uint8 input is converted like load_bin_as_float."""
    from mysteryann_b200 import io

    rng = np.random.default_rng(8)
    base = rng.integers(0, 255, (3000, 64)).astype(np.uint8)
    q = rng.integers(0, 255, (64, 64)).astype(np.uint8)
    for name, x in (("base.u8bin", base), ("q.u8bin", q)):
        with open(tmp_path / name, "wb") as f:
            np.array(x.shape, np.int32).tofile(f)
            x.tofile(f)
    run([os.path.join(bins, "compute_groundtruth"), "--data_type", "uint8", "--dist_fn", "cosine", "--base_file",
         str(tmp_path / "base.u8bin"), "--query_file", str(tmp_path / "q.u8bin"), "--gt_file", str(tmp_path / "gt.bin"), "--K", "10"])
    ids, dists = io.read_ibin(tmp_path / "gt.bin")
    b = base.astype(np.float64)
    b /= np.linalg.norm(b, axis=1, keepdims=True)
    qq = q.astype(np.float64)
    qq /= np.linalg.norm(qq, axis=1, keepdims=True)
    d = ((qq[:, None, :] - b[None, :, :]) ** 2).sum(-1)
    order = np.argsort(d, axis=1, kind="stable")[:, :10]
    want_d = np.take_along_axis(d, order, 1)
    assert np.allclose(dists, want_d, rtol=1e-4, atol=2e-6)
    agree = np.mean([len(set(a) & set(b_)) / 10 for a, b_ in zip(ids.tolist(), order.tolist())])
    assert agree > 0.995   # only FP32 near-ties may differ from the FP64 ranking


def test_pipeline_knn_build_search(bins, oracle, tmp_path):
    """The reference's workflow end to end on the GPU: learn->base kNN, RoarGraph build (--gpu_build), test ground truth,
    L sweep; the CSV row of every L must equal what the CPU oracle computes on the index file the build wrote."""
    from mysteryann_b200 import io, synth

    base, train, test = synth.make_numpy(20000, 20000, 500, 200, seed=11)
    for name, x in (("base.fbin", base), ("train.fbin", train), ("test.fbin", test)):
        io.write_fbin(tmp_path / name, x)
    gt_tool = os.path.join(bins, "compute_groundtruth")
    run([gt_tool, "--data_type", "float", "--dist_fn", "mips", "--base_file", str(tmp_path / "base.fbin"), "--query_file",
         str(tmp_path / "train.fbin"), "--gt_file", str(tmp_path / "train.gt.bin"), "--K", "100"])
    run([gt_tool, "--data_type", "float", "--dist_fn", "mips", "--base_file", str(tmp_path / "base.fbin"), "--query_file",
         str(tmp_path / "test.fbin"), "--gt_file", str(tmp_path / "test.gt.bin"), "--K", "100"])
    out = run([os.path.join(bins, "test_build_roargraph"), "--data_type", "float", "--dist", "ip", "--base_data_path",
               str(tmp_path / "base.fbin"), "--sampled_query_data_path", str(tmp_path / "train.fbin"),
               "--projection_index_save_path", str(tmp_path / "rg.index"), "--learn_base_nn_path", str(tmp_path / "train.gt.bin"),
               "--M_sq", "100", "--M_pjbp", "35", "--L_pjpq", "500", "--gpu_build", "1"])
    assert "GPU build phases" in out
    Ls = [10, 20, 50, 100]
    run([os.path.join(bins, "test_search_roargraph"), "--data_type", "float", "--dist", "ip", "--base_data_path",
         str(tmp_path / "base.fbin"), "--query_path", str(tmp_path / "test.fbin"), "--gt_path", str(tmp_path / "test.gt.bin"),
         "--projection_index_save_path", str(tmp_path / "rg.index"), "--k", "10", "--evaluation_save_path", str(tmp_path / "eval.csv"),
         "--L_pq"] + [str(L) for L in Ls])
    rows = [line.strip().split(",") for line in open(tmp_path / "eval.csv") if line.strip()]
    assert [int(r[0]) for r in rows] == Ls and all(len(r) == 6 for r in rows)
    ep, off, adj = io.read_index(tmp_path / "rg.index")
    gt_ids, _ = io.read_ibin(tmp_path / "test.gt.bin")
    for r in rows:
        L = int(r[0])
        want = oracle.search(base, off, adj, ep, test, 10, L, metric=1)
        recall = oracle.recall(want["ids"], gt_ids, 10)
        assert abs(float(r[4]) - recall) < 1e-5, (L, r, recall)
        assert abs(float(r[2]) - want["cmps"].mean()) < 1e-2 * max(1.0, want["cmps"].mean() * 1e-3), (L, r)
        assert abs(float(r[5]) - want["hops"].mean()) < 1e-2, (L, r)
    assert float(rows[-1][4]) > 0.9   # the GPU-built graph is a working RoarGraph


def test_search_driver_multi_gpu_flag(bins, oracle, tmp_path):
    """--devices N replicates the index on N GPUs and shards the queries over one host thread per GPU: the CSV (recall,
    avg cmps, avg hops) must equal the single-GPU run's; asking for more GPUs than are visible is an error."""
    from mysteryann_b200 import capi, hostlib, io, synth

    base, train, test = synth.make_numpy(6000, 6000, 301, 200, seed=5)
    knn, _, _ = oracle.exact_knn(base, train, 100, metric=1)
    for name, x in (("base.fbin", base), ("test.fbin", test)):
        io.write_fbin(tmp_path / name, x)
    hostlib.build_index(base, train, knn, str(tmp_path / "rg.index"), metric=1, M_sq=100, M_pjbp=35, L_pjpq=100, threads=4)
    gt, gd, _ = oracle.exact_knn(base, test, 100, metric=1)
    io.write_ibin(tmp_path / "gt.bin", gt, gd)
    cmd = [os.path.join(bins, "test_search_roargraph"), "--data_type", "float", "--dist", "ip", "--base_data_path",
           str(tmp_path / "base.fbin"), "--query_path", str(tmp_path / "test.fbin"), "--gt_path", str(tmp_path / "gt.bin"),
           "--projection_index_save_path", str(tmp_path / "rg.index"), "--k", "10", "--L_pq", "10", "40", "100"]

    def rows(path):
        return [[float(v) for i, v in enumerate(line.strip().split(",")) if i in (0, 2, 4, 5)] for line in open(path) if line.strip()]

    run(cmd + ["--evaluation_save_path", str(tmp_path / "one.csv")])
    n_dev = capi.device_count()
    if n_dev >= 2:
        out = run(cmd + ["--evaluation_save_path", str(tmp_path / "two.csv"), "--devices", "2"])
        assert "queries sharded" in out
        assert rows(tmp_path / "one.csv") == rows(tmp_path / "two.csv")
    p = subprocess.run(cmd + ["--devices", str(n_dev + 1)], capture_output=True, text=True, timeout=600)
    assert p.returncode != 0 and "CUDA device(s) visible" in (p.stdout + p.stderr)


def test_search_driver_l2_and_cosine(bins, oracle, tmp_path):
    """test_search_roargraph with --dist l2 (CSV must equal the oracle's numbers on the same index file) and --dist cosine
    (the driver normalises base rows and queries like the reference, src/index_bipartite.cpp:2675-2680 and
    tests/test_search_roargraph.cpp:167-172, then searches with the inner product: it must agree with an inner-product
    search over rows normalised beforehand, up to the rounding of the two normalisations)."""
    from mysteryann_b200 import io

    rng = np.random.default_rng(17)
    n, dim, nq = 3000, 200, 200
    base = (rng.standard_normal((n, dim)) * rng.uniform(0.5, 2.0, (n, 1))).astype(np.float32)
    test = (rng.standard_normal((nq, dim)) + 0.3).astype(np.float32)
    deg = rng.integers(8, 25, n)
    off = np.zeros(n + 1, np.uint64)
    np.cumsum(deg, out=off[1:])
    adj = rng.integers(0, n, int(off[-1])).astype(np.uint32)
    ep = 11
    io.write_fbin(tmp_path / "base.fbin", base)
    io.write_fbin(tmp_path / "test.fbin", test)
    io.write_index(tmp_path / "rnd.index", ep, off, adj)
    Ls = [10, 30, 60]

    def drive(dist, gt_name, csv_name):
        run([os.path.join(bins, "test_search_roargraph"), "--data_type", "float", "--dist", dist, "--base_data_path",
             str(tmp_path / "base.fbin"), "--query_path", str(tmp_path / "test.fbin"), "--gt_path", str(tmp_path / gt_name),
             "--projection_index_save_path", str(tmp_path / "rnd.index"), "--k", "10", "--evaluation_save_path",
             str(tmp_path / csv_name), "--L_pq"] + [str(L) for L in Ls])
        rows = [[float(v) for v in line.strip().split(",")] for line in open(tmp_path / csv_name) if line.strip()]
        assert [int(r[0]) for r in rows] == Ls
        return rows

    # L2: exact agreement with the oracle
    gt, gd, _ = oracle.exact_knn(base, test, 100, metric=0)
    io.write_ibin(tmp_path / "gt_l2.bin", gt, gd)
    for r in drive("l2", "gt_l2.bin", "l2.csv"):
        want = oracle.search(base, off, adj, ep, test, 10, int(r[0]), metric=0)
        assert abs(r[4] - oracle.recall(want["ids"], gt, 10)) < 1e-5, r
        assert abs(r[2] - want["cmps"].mean()) < 1e-2 and abs(r[5] - want["hops"].mean()) < 1e-2, r

    # cosine: the same as inner product over pre-normalised rows
    nb = (base / np.linalg.norm(base, axis=1, keepdims=True)).astype(np.float32)
    nt = (test / np.linalg.norm(test, axis=1, keepdims=True)).astype(np.float32)
    gt, gd, _ = oracle.exact_knn(nb, nt, 100, metric=1)
    io.write_ibin(tmp_path / "gt_cos.bin", gt, gd)
    for r in drive("cosine", "gt_cos.bin", "cos.csv"):
        want = oracle.search(nb, off, adj, ep, nt, 10, int(r[0]), metric=1)
        assert abs(r[4] - oracle.recall(want["ids"], gt, 10)) < 5e-3, r
        assert abs(r[2] - want["cmps"].mean()) < 5e-3 * want["cmps"].mean() and abs(r[5] - want["hops"].mean()) < 0.5, r


def test_search_driver_unpadded_dim(bins, oracle, tmp_path):
    """dim % 8 != 0 (D = 100): the loader pads rows to 104 floats once, for base and queries alike (the reference re-aligns
    the already padded query buffer with the unpadded stride, tests/test_search_roargraph.cpp:166 - a latent bug every one of
    its datasets avoids by having dim % 8 == 0).  The CSV must equal the oracle's numbers on zero-padded rows."""
    from mysteryann_b200 import io

    rng = np.random.default_rng(23)
    n, dim, nq = 2500, 100, 120
    base = rng.standard_normal((n, dim)).astype(np.float32)
    test = (rng.standard_normal((nq, dim)) + 0.2).astype(np.float32)
    deg = rng.integers(6, 20, n)
    off = np.zeros(n + 1, np.uint64)
    np.cumsum(deg, out=off[1:])
    adj = rng.integers(0, n, int(off[-1])).astype(np.uint32)
    io.write_fbin(tmp_path / "base.fbin", base)
    io.write_fbin(tmp_path / "test.fbin", test)
    io.write_index(tmp_path / "rnd.index", 3, off, adj)
    gt, gd, _ = oracle.exact_knn(pad8(base), pad8(test), 50, metric=1)
    io.write_ibin(tmp_path / "gt.bin", gt, gd)
    run([os.path.join(bins, "test_search_roargraph"), "--data_type", "float", "--dist", "ip", "--base_data_path",
         str(tmp_path / "base.fbin"), "--query_path", str(tmp_path / "test.fbin"), "--gt_path", str(tmp_path / "gt.bin"),
         "--projection_index_save_path", str(tmp_path / "rnd.index"), "--k", "10", "--evaluation_save_path",
         str(tmp_path / "out.csv"), "--L_pq", "10", "40"])
    rows = [[float(v) for v in line.strip().split(",")] for line in open(tmp_path / "out.csv") if line.strip()]
    assert [int(r[0]) for r in rows] == [10, 40]
    for r in rows:
        want = oracle.search(pad8(base), off, adj, 3, pad8(test), 10, int(r[0]), metric=1)
        assert abs(r[4] - oracle.recall(want["ids"], gt, 10)) < 1e-5, r
        assert abs(r[2] - want["cmps"].mean()) < 1e-2 and abs(r[5] - want["hops"].mean()) < 1e-2, r


def test_search_driver_rejects_foreign_index(bins, tmp_path):
    """An index file whose neighbour ids or degree words do not fit the base must be refused before anything reaches the GPU."""
    from mysteryann_b200 import io

    rng = np.random.default_rng(29)
    base = rng.standard_normal((200, 16)).astype(np.float32)
    io.write_fbin(tmp_path / "base.fbin", base)
    io.write_fbin(tmp_path / "test.fbin", base[:10])
    io.write_ibin(tmp_path / "gt.bin", np.zeros((10, 10), np.uint32), np.zeros((10, 10), np.float32))
    off = np.arange(0, 201 * 4, 4, dtype=np.uint64)
    adj = rng.integers(0, 200, 800).astype(np.uint32)
    adj[123] = 5000                                        # out of range
    io.write_index(tmp_path / "bad_id.index", 0, off, adj)
    raw = np.fromfile(tmp_path / "bad_id.index", dtype=np.uint32).copy()
    raw[2] = 0x7FFFFFF0                                    # absurd degree word of node 0
    raw.tofile(tmp_path / "bad_deg.index")
    for name in ("bad_id.index", "bad_deg.index"):
        p = subprocess.run([os.path.join(bins, "test_search_roargraph"), "--data_type", "float", "--dist", "ip", "--base_data_path",
                            str(tmp_path / "base.fbin"), "--query_path", str(tmp_path / "test.fbin"), "--gt_path", str(tmp_path / "gt.bin"),
                            "--projection_index_save_path", str(tmp_path / name), "--k", "10", "--L_pq", "10"],
                           capture_output=True, text=True, timeout=300)
        assert p.returncode != 0, name
        assert "out of range" in (p.stdout + p.stderr) or "truncated" in (p.stdout + p.stderr), p.stdout + p.stderr


def test_per_query_api_from_openmp_threads(tmp_path):
    """Existing callers of the reference call IndexBipartite::SearchRoarGraph once per query from OpenMP threads
    (tests/test_search_roargraph.cpp:203-209).  The drop-in class must give the reference's answers that way too
    (concurrent callers are micro-batched into one GPU launch inside the class; RG_MICROBATCH_US=0 makes every call its own
    batch of one)."""
    from conftest import load_case
    from mysteryann_b200 import hostlib, io

    hostlib.build()
    c = load_case("ip_d200")
    io.write_fbin(tmp_path / "base.fbin", c["base"])
    io.write_index(tmp_path / "g.index", c["ep"], c["offsets"], c["adj"])
    for window in ("50", "0"):
        os.environ["RG_MICROBATCH_US"] = window
        for L in (10, 32):
            got = hostlib.search_per_query(tmp_path / "base.fbin", tmp_path / "g.index", c["test"], 10, L, metric=1, threads=8)
            for key in ("ids", "cmps", "hops"):
                assert (got[key] == c[f"{key}_{L}"]).all(), (key, L, window)
            assert (got["dists"].view(np.uint32) == c[f"dists_{L}"].view(np.uint32)).all()
            print(f"per-query API, 8 threads, window {window} us, L={L}: {len(c['test']) / got['seconds']:.0f} queries/s")
    del os.environ["RG_MICROBATCH_US"]
