"""Graph-build quality at BASELINE.json's config C1 (100K x 200 IP base, 100K training queries, 10K OOD test queries, k=10,
M_sq=100 M_pjbp=35 L_pjpq=500): the GPU build (rg_build_roargraph_device) against the REFERENCE's own multi-threaded build
(oracle/_ref = the unmodified src/index_bipartite.cpp, LinkProjection :1043-1277, at -T = all host threads, run twice - two
multi-threaded reference builds differ from each other), searched with the same bit-exact beam search at every L of the
reference's sweep (run_roargraph_search_test.sh:13, cut at 500).  Reports recall@10 and mean cmps per L and the GPU
build's worst deficit against the worse / the better of the two reference builds.

Second part (--l2-n > 0): what dropping the value-initialised "phantom" queue entries of src/index_bipartite.cpp:1438 does.
Under L2 a phantom {id 0, distance 0} sorts FIRST, so the reference's supply re-prune keeps node 0 as a neighbour of every
node whose supply list overflows; the GPU build does not reproduce that.  Measured: in-degree of node 0 and the recall curves.

    python tools/build_quality_c1.py --out profiles/r02_build_quality_c1.txt
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
from bench import L_SWEEP  # noqa: E402
from mysteryann_b200 import build, capi, hostlib, io, synth  # noqa: E402
from oracle.binding import Ref, ref_available  # noqa: E402


def recall(ids, gt, k):
    return float(np.mean([len(set(a[:k].tolist()) & set(b[:k].tolist())) / k for a, b in zip(ids, gt)]))


def curve(ix, test, gt, k, Ls):
    out = []
    for L in Ls:
        r = ix.search(test, k, L)
        out.append((L, recall(r["ids"], gt, k), float(r["cmps"].mean())))
    return out


def run(metric, n, n_train, n_test, dim, M_sq, M, L_build, threads, tmp, log, ref_builds=2, host_build=False):
    base, train, test = synth.make_numpy(n, n_train, n_test, dim)
    t0 = time.time()
    knn, knn_d = capi.knn_exact(base, train, M_sq, metric=metric)
    gt, _ = capi.knn_exact(base, test, 10, metric=metric)
    log(f"# data {n} x {dim}, {n_train} training / {n_test} test queries, metric={'ip' if metric == 1 else 'l2'}; "
        f"GPU exact kNN (train K={M_sq}, test K=10) {time.time() - t0:.1f} s")
    io.write_fbin(os.path.join(tmp, "base.fbin"), base)
    io.write_fbin(os.path.join(tmp, "train.fbin"), train)
    io.write_ibin(os.path.join(tmp, "knn.ibin"), knn, knn_d)
    d_base = torch.from_numpy(base).cuda()
    graphs = {}
    if ref_available():
        r = Ref()
        for i in range(ref_builds):
            path = os.path.join(tmp, f"ref{i}.index")
            sec = r.build_index(os.path.join(tmp, "base.fbin"), os.path.join(tmp, "train.fbin"), os.path.join(tmp, "knn.ibin"), path,
                                metric=metric, M_sq=M_sq, M_pjbp=M, L_pjpq=L_build, threads=threads)
            graphs[f"reference -T {threads} #{i + 1}"] = (io.read_index(path), sec)
    if host_build or not graphs:
        path = os.path.join(tmp, "host.index")
        sec = hostlib.build_index(base, train, knn, path, metric=metric, M_sq=M_sq, M_pjbp=M, L_pjpq=L_build, threads=threads)
        graphs[f"host restatement -T {threads}"] = (io.read_index(path), sec)
    t0 = time.time()
    g = capi.Graph(d_base, torch.from_numpy(knn.view(np.int32)).cuda(), M_sq=M_sq, M_pjbp=M, L_pjpq=L_build, metric=metric)
    graphs["GPU build"] = (g.download(), time.time() - t0)
    Ls = [L for L in L_SWEEP if L >= 10]
    curves = {}
    for name, ((ep, off, adj), sec) in graphs.items():
        ix = capi.Index(d_base, off, adj, ep, metric=metric)
        curves[name] = curve(ix, test, gt, 10, Ls)
        ix.close()
        deg = np.diff(off.astype(np.int64))
        indeg0 = int((adj == 0).sum())
        log(f"# {name}: {sec:.1f} s, ep {ep}, avg degree {deg.mean():.2f}, max {deg.max()}, edges into node 0: {indeg0}")
    names = list(curves)
    log("L_pq  " + "  ".join(f"{nm[:24]:>24s}" for nm in names) + "   (recall@10 / mean cmps)")
    for i, L in enumerate(Ls):
        log(f"{L:<5d} " + "  ".join(f"{curves[nm][i][1]:>14.4f} /{curves[nm][i][2]:>8.0f}" for nm in names))
    refs = [nm for nm in names if nm != "GPU build"]
    summary = {}
    if refs:
        worst_vs_min = min(curves["GPU build"][i][1] - min(curves[nm][i][1] for nm in refs) for i in range(len(Ls)))
        worst_vs_max = min(curves["GPU build"][i][1] - max(curves[nm][i][1] for nm in refs) for i in range(len(Ls)))
        spread = max(abs(curves[refs[0]][i][1] - curves[refs[-1]][i][1]) for i in range(len(Ls))) if len(refs) > 1 else 0.0
        summary = dict(worst_gpu_minus_worse_reference=round(worst_vs_min, 4), worst_gpu_minus_better_reference=round(worst_vs_max, 4),
                       max_spread_between_reference_builds=round(spread, 4))
        log("# " + json.dumps(summary))
    g.close()
    return summary


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=100_000)
    ap.add_argument("--train", type=int, default=100_000)
    ap.add_argument("--test", type=int, default=10_000)
    ap.add_argument("--dim", type=int, default=200)
    ap.add_argument("--threads", type=int, default=os.cpu_count() or 8)
    ap.add_argument("--l2-n", type=int, default=30_000, help="size of the L2 phantom-entry experiment (0 = skip)")
    ap.add_argument("--out", default="")
    a = ap.parse_args()
    build.build()
    hostlib.build()
    lines = []

    def log(s):
        print(s, flush=True)
        lines.append(s)

    log(f"# tools/build_quality_c1.py on {torch.cuda.get_device_name(0)}, {a.threads} host threads")
    with tempfile.TemporaryDirectory() as tmp:
        run(1, a.n, a.train, a.test, a.dim, 100, 35, 500, a.threads, tmp, log)
    if a.l2_n:
        log("")
        log("# ---- L2 metric: the :1438 phantom entries (reference + host restatement reproduce them, the GPU build drops them)")
        with tempfile.TemporaryDirectory() as tmp:
            run(0, a.l2_n, a.l2_n, 2000, a.dim, 100, 35, 500, a.threads, tmp, log, ref_builds=1, host_build=True)
    if a.out:
        open(a.out, "w").write("\n".join(lines) + "\n")


if __name__ == "__main__":
    main()
