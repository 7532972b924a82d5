"""CPU tests of bench.py's host-side contract pieces that a GPU run depends on but cannot check itself: the committed
traffic profile belongs to the committed K1 sources (otherwise the driver's bench line would carry "traffic": null), the
canonical BASELINE.json configurations resolve to the shapes DESIGN.md section 9 names, and the reference arm without a GPU
arm's library still refuses to run silently."""
import json
import os
import sys
import types

from conftest import ROOT

sys.path.insert(0, ROOT)
import bench  # noqa: E402


def test_traffic_profile_matches_committed_kernel():
    t = json.load(open(os.path.join(ROOT, "profiles", "k1_traffic.json")))
    assert t["k1_source_hash"] == bench.k1_source_hash(), \
        "profiles/k1_traffic.json was captured from other K1 sources: re-run tools/make_k1_traffic.py on a fresh ncu capture"
    w = t["workload"]
    assert (w["n_base"], w["dim"], w["queries"], w["k"]) == (10_000_000, 200, 10_000, 10)   # BASELINE.json configs[1]
    # the capture's DRAM bytes stay close to the gathered rows: re-reads would be the first thing to fix
    assert 1.0 <= t["dram_bytes_per_launch"] / t["algorithmic_bytes_per_launch"] < 1.10


def test_traffic_lookup_is_keyed_on_workload_and_kernel():
    t = json.load(open(os.path.join(ROOT, "profiles", "k1_traffic.json")))
    w = t["workload"]
    args = types.SimpleNamespace(n=w["n_base"], dim=w["dim"], queries=w["queries"], k=w["k"], train=0)
    got, src = bench.load_traffic(args, w["L_pq"])
    assert got == t["dram_bytes_per_launch"] and src.startswith("from_profile:")
    got, src = bench.load_traffic(args, w["L_pq"] + 5)           # another beam width: not quoted
    assert got is None and "workload" in src
    args.n //= 2
    got, src = bench.load_traffic(args, w["L_pq"])
    assert got is None and "workload" in src


def test_canonical_configs(monkeypatch):
    want = {"C2": (10_000_000, 200, 10_000, 10, False), "C3": (2_500_000, 512, 10_000, 10, True),
            "C3k100": (2_500_000, 512, 10_000, 100, True), "C5": (100_000_000, 200, 100_000, 10, False)}
    for name, (n, dim, nq, k, norm) in want.items():
        monkeypatch.setattr(sys, "argv", ["bench.py", "--config", name])
        a = bench.parse_args()
        assert (a.n, a.dim, a.queries, a.k, a.normalize) == (n, dim, nq, k, norm), name
        assert a.cpu_sample == nq                                  # the CPU arm times the whole batch (same_config)
    monkeypatch.setattr(sys, "argv", ["bench.py"])
    a = bench.parse_args()
    assert (a.gpus, a.n, a.dim, a.queries, a.k, a.recall) == (1, 10_000_000, 200, 10_000, 10, 0.9)
    assert a.warmup >= 3 and a.steps >= 1


def test_l_sweep_is_the_references():
    # run_roargraph_search_test.sh:13 of the reference, cut at 500
    Ls = bench.L_SWEEP
    assert list(Ls) == sorted(Ls) and Ls[0] == 10 and Ls[-1] == 500 and 55 in Ls
