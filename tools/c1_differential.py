"""BASELINE.json configs[0] (C1) as a full differential run: 100K x 200 IP base, 100K training queries, 10K OOD test queries,
index built by the COMPILED REFERENCE (oracle/_ref, -T threads), then every L_pq of the reference's own sweep
(run_roargraph_search_test.sh:13: 10 ... 2000) searched by the reference's SearchRoarGraph (OpenMP loop of
tests/test_search_roargraph.cpp:196-214) and by K1 on the GPU, on the same index file; ids, distance bit patterns, cmps and
hops are compared for every query.  Needs /root/reference-built oracle/_ref (travels to the GPU box as a prebuilt .so).

    python tools/c1_differential.py --out profiles/r02_c1_differential.txt
"""
import argparse
import json
import os
import sys
import tempfile
import time

import numpy as np
import torch

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from mysteryann_b200 import build, capi, io, synth  # noqa: E402
from oracle.binding import Ref, ref_available  # noqa: E402

L_SWEEP = [10, 15, 20, 25, 30, 35, 40, 45, 50, 55, 60, 65, 70, 75, 80, 85, 90, 95, 100, 110, 120, 130, 140, 150, 160, 170, 180,
           190, 200, 220, 240, 260, 280, 300, 350, 400, 450, 500, 550, 600, 650, 700, 750, 800, 900, 1000, 1100, 1200, 1300,
           1400, 1500, 1600, 1700, 1800, 1900, 2000]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=100_000)
    ap.add_argument("--train", type=int, default=100_000)
    ap.add_argument("--test", type=int, default=10_000)
    ap.add_argument("--dim", type=int, default=200)
    ap.add_argument("--threads", type=int, default=os.cpu_count() or 8)
    ap.add_argument("--big-L-queries", type=int, default=2000, help="queries compared at L_pq > 500 (CPU time)")
    ap.add_argument("--out", default="")
    a = ap.parse_args()
    assert ref_available(), "oracle/_ref is not built"
    build.build()
    lines = []

    def log(s):
        print(s, flush=True)
        lines.append(s)

    base, train, test = synth.make_numpy(a.n, a.train, a.test, a.dim)
    knn, knn_d = capi.knn_exact(base, train, 100, metric=1)
    gt, _ = capi.knn_exact(base, test, 10, metric=1)
    r = Ref()
    bad = 0
    with tempfile.TemporaryDirectory() as tmp:
        fb, ft, fk, fi = (os.path.join(tmp, x) for x in ("base.fbin", "train.fbin", "knn.ibin", "ref.index"))
        io.write_fbin(fb, base)
        io.write_fbin(ft, train)
        io.write_ibin(fk, knn, knn_d)
        sec = r.build_index(fb, ft, fk, fi, metric=1, M_sq=100, M_pjbp=35, L_pjpq=500, threads=a.threads)
        ep, off, adj = io.read_index(fi)
        log(f"# C1 differential on {torch.cuda.get_device_name(0)}, {a.threads} host threads: {a.n} x {a.dim} IP, {a.train} training, "
            f"{a.test} test queries; index built by the compiled reference in {sec:.1f} s (avg degree {len(adj) / a.n:.2f}, ep {ep})")
        h = r.open(fb, fi, metric=1, threads=a.threads)
        ix = capi.Index(torch.from_numpy(base).cuda(), off, adj, ep, metric=1)
        log("L_pq   queries  recall@10  mean cmps  ids  dists(bits)  cmps  hops   reference QPS      GPU QPS (host buffers)")
        for L in L_SWEEP:
            nq = a.test if L <= 500 else min(a.test, a.big_L_queries)
            q = test[:nq]
            want = r.search(h, q, 10, L, threads=a.threads, warmup=(L == L_SWEEP[0]))
            ix.search(q[:64], 10, L)
            t0 = time.time()
            got = ix.search(q, 10, L)
            gsec = time.time() - t0
            par = bench.compare_results(got, want)
            bad += 0 if par["ok"] else 1
            rec = bench.recall_at_k(got["ids"], gt[:nq], 10)
            log(f"{L:<6d} {nq:<8d} {rec:<10.4f} {got['cmps'].mean():<10.1f} {str(par['ids']):<5s}{str(par['dists_bits']):<13s}{str(par['cmps']):<6s}"
                f"{str(par['hops']):<7s}{nq / want['seconds']:>12.0f} {nq / gsec:>18.0f}")
        r.close(h)
        ix.close()
    log("# " + json.dumps({"L_values": len(L_SWEEP), "L_values_with_any_difference": bad}))
    if a.out:
        open(a.out, "w").write("\n".join(lines) + "\n")
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
