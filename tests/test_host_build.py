"""CPU tests of the host C++ layer: BuildRoarGraph reproduces the reference's index file byte for byte at one
thread (golden fixtures; and live against the compiled reference on a larger seeded set), file formats round-trip,
the CLI drivers parse the reference's flags."""
import hashlib
import os
import subprocess

import numpy as np
import pytest

from conftest import CASES, GOLDEN, load_case
from mysteryann_b200 import hostlib, io, synth


@pytest.fixture(scope="module", autouse=True)
def _built():
    from mysteryann_b200 import build

    build.build()
    hostlib.build()


@pytest.mark.parametrize("name", CASES)
def test_build_matches_golden_index(name, tmp_path):
    c = load_case(name)
    out = str(tmp_path / "index")
    base = io.pad_rows(c["base"])
    train = io.pad_rows(c["train"])
    hostlib.build_index(base, train, c["knn_ids"], out, metric=c["metric"], M_sq=c["M_sq"], M_pjbp=c["M_pjbp"],
                        L_pjpq=c["L_pjpq"], threads=1)
    got = np.fromfile(out, dtype=np.uint8)
    assert got.size == c["index"].size and (got == c["index"]).all()


def test_build_matches_reference_live(ref, oracle, tmp_path):
    """5K x 104-d L2 set (exercises the phantom-entry quirk of the internal reverse prune under L2)."""
    base, train, _ = synth.make_numpy(5000, 3000, 10, 104, seed=77)
    knn, knn_d, _ = oracle.exact_knn(base, train, 40, metric=0)
    p = lambda f: str(tmp_path / f)
    io.write_fbin(p("b"), base); io.write_fbin(p("t"), train); io.write_ibin(p("nn"), knn, knn_d)
    ref.build_index(p("b"), p("t"), p("nn"), p("ref_index"), metric=0, M_sq=40, M_pjbp=14, L_pjpq=60, threads=1)
    hostlib.build_index(base, train, knn, p("our_index"), metric=0, M_sq=40, M_pjbp=14, L_pjpq=60, threads=1)
    a, b = open(p("ref_index"), "rb").read(), open(p("our_index"), "rb").read()
    assert hashlib.md5(a).hexdigest() == hashlib.md5(b).hexdigest()


def test_multithreaded_build_is_valid(tmp_path):
    c = load_case("ip_d200")
    out = str(tmp_path / "index_mt")
    hostlib.build_index(c["base"], c["train"], c["knn_ids"], out, metric=1, M_sq=c["M_sq"], M_pjbp=c["M_pjbp"],
                        L_pjpq=c["L_pjpq"], threads=4)
    ep, off, adj = io.read_index(out)
    deg = np.diff(off)
    assert ep == c["ep"] and len(deg) == len(c["base"]) and deg.max() <= 2 * c["M_pjbp"] and adj.max() < len(deg)


def test_index_and_ibin_roundtrip(tmp_path):
    c = load_case("l2_d48")
    p = str(tmp_path / "idx")
    io.write_index(p, c["ep"], c["offsets"], c["adj"])
    assert (np.fromfile(p, dtype=np.uint8) == c["index"]).all()
    ids = np.arange(12, dtype=np.uint32).reshape(3, 4)
    io.write_ibin(p + ".ibin", ids, ids.astype(np.float32))
    a, b = io.read_ibin(p + ".ibin")
    assert (a == ids).all() and (b == ids).all()


def test_cli_flags(tmp_path):
    """The drivers accept the reference's flag set (run_roargraph_test.sh:5-10) and build the same index."""
    c = load_case("ip_d24_norm")
    p = lambda f: str(tmp_path / f)
    io.write_fbin(p("base.fbin"), c["base"]); io.write_fbin(p("train.fbin"), c["train"])
    io.write_ibin(p("nn.ibin"), c["knn_ids"], np.zeros(c["knn_ids"].shape, np.float32))
    exe = os.path.join(hostlib.BIN_DIR, "test_build_roargraph")
    r = subprocess.run([exe, "--data_type", "float", "--dist", "ip", "--base_data_path", p("base.fbin"),
                        "--sampled_query_data_path", p("train.fbin"), "--projection_index_save_path", p("out.index"),
                        "--learn_base_nn_path", p("nn.ibin"), "--M_sq", str(c["M_sq"]), "--M_pjbp", str(c["M_pjbp"]),
                        "--L_pjpq", str(c["L_pjpq"]), "-T", "1"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert "indexing time" in r.stdout
    assert (np.fromfile(p("out.index"), dtype=np.uint8) == c["index"]).all()
    r = subprocess.run([exe, "--dist", "ip"], capture_output=True, text=True)   # missing required flags
    assert r.returncode != 0 and "required" in r.stderr


def test_search_without_gpu_fails_loudly(tmp_path):
    """The host class has no CPU search path: without a CUDA device InitVisitedListPool / SearchRoarGraph raise."""
    from conftest import load_case
    from mysteryann_b200 import capi, hostlib, io

    if capi.device_count() > 0:
        pytest.skip("a GPU is present")
    c = load_case("ip_d200")
    io.write_fbin(tmp_path / "base.fbin", c["base"])
    io.write_index(tmp_path / "g.index", c["ep"], c["offsets"], c["adj"])
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        hostlib.search_per_query(tmp_path / "base.fbin", tmp_path / "g.index", c["test"], 10, 32)


def test_error_behaviour_matches_reference(tmp_path):
    """Error paths of the loaders / driver as the reference has them (SURVEY.md 8b "Errors"): a truncated fbin raises
    "Data file size wrong!" (util.h:124), an unopenable input file exits with status 255 (exit(-1), util.h:87-90), a kNN
    file whose row count differs from the training set is rejected (src/index_bipartite.cpp:2635-2637), an unwritable
    save path fails (:2608-2610)."""
    c = load_case("ip_d24_norm")
    p = lambda f: str(tmp_path / f)
    io.write_fbin(p("base.fbin"), c["base"]); io.write_fbin(p("train.fbin"), c["train"])
    io.write_ibin(p("nn.ibin"), c["knn_ids"], np.zeros(c["knn_ids"].shape, np.float32))
    exe = os.path.join(hostlib.BIN_DIR, "test_build_roargraph")

    def build(**over):
        a = {"--base_data_path": p("base.fbin"), "--sampled_query_data_path": p("train.fbin"),
             "--projection_index_save_path": p("out.index"), "--learn_base_nn_path": p("nn.ibin")}
        a.update(over)
        cmd = [exe, "--data_type", "float", "--dist", "ip", "--M_sq", str(c["M_sq"]), "--M_pjbp", str(c["M_pjbp"]),
               "--L_pjpq", str(c["L_pjpq"]), "-T", "1"]
        for k_, v in a.items():
            cmd += [k_, v]
        return subprocess.run(cmd, capture_output=True, text=True)

    raw = open(p("base.fbin"), "rb").read()
    open(p("short.fbin"), "wb").write(raw[: len(raw) // 2])
    r = build(**{"--base_data_path": p("short.fbin")})
    assert r.returncode != 0 and "Data file size wrong!" in (r.stderr + r.stdout)
    r = build(**{"--base_data_path": p("does_not_exist.fbin")})
    assert r.returncode == 255
    io.write_ibin(p("bad_nn.ibin"), c["knn_ids"][:-3], np.zeros(c["knn_ids"][:-3].shape, np.float32))
    r = build(**{"--learn_base_nn_path": p("bad_nn.ibin")})
    assert r.returncode != 0
    r = build(**{"--projection_index_save_path": p("no_such_dir/out.index")})
    assert r.returncode != 0 and not os.path.exists(p("no_such_dir/out.index"))
    assert "cannot open file" in (r.stderr + r.stdout)          # same what() as the reference's SaveProjectionGraph
