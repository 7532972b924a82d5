"""Bit-exact parity of K1 at the LARGEST shape of BASELINE.json (C5: 100M x 200 fp32 = 80 GB of base rows + a 28.8 GB
fixed-stride adjacency on one B200), i.e. with 64-bit row offsets, > 2^32 base elements and the full memory plan of
DESIGN.md section 3, against the CPU oracle - without an 80 GB host copy of the base:

  * the graph only ever points into a random subset S of the rows (S spread over the whole id range, entry point in S),
    so a search can only touch rows of S;
  * the host mirror of the base is a lazily-mapped n x dim array in which only the rows of S are filled (a few GB of
    resident pages); the oracle indexes it with the same global ids.

One node carries 70 neighbours so that the adjacency stride is the canonical 72 words (M_pjbp = 35).  Prints one JSON line.

    python tools/check_large_index.py --n 100000000 --subset 1000000 --queries 2000
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from mysteryann_b200 import build, capi  # noqa: E402
from oracle.binding import Oracle  # noqa: E402  (checker only)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=100_000_000)
    ap.add_argument("--dim", type=int, default=200)
    ap.add_argument("--subset", type=int, default=1_000_000)
    ap.add_argument("--queries", type=int, default=2000)
    ap.add_argument("--k", type=int, default=10)
    ap.add_argument("--L", type=int, nargs="+", default=[60, 200])
    ap.add_argument("--seed", type=int, default=7)
    a = ap.parse_args()
    build.build()
    dev = torch.device("cuda", 0)
    n, dim, m = a.n, a.dim, a.subset
    rng = np.random.default_rng(a.seed)
    t0 = time.time()

    # base rows, generated on the device in slices (no host copy of the whole array exists anywhere)
    g = torch.Generator(device=dev).manual_seed(a.seed)
    base = torch.empty((n, dim), dtype=torch.float32, device=dev)
    step = 4_000_000
    for lo in range(0, n, step):
        base[lo:lo + step].normal_(generator=g)
    torch.cuda.synchronize()
    t_base = time.time() - t0

    # subset S (sorted global ids, always holding the first and the last row) and a random graph inside S
    S = np.unique(np.concatenate([rng.integers(0, n, size=m, dtype=np.int64), np.array([0, n - 1])]))
    m = len(S)
    deg = rng.integers(8, 25, size=m).astype(np.int64)
    deg[rng.integers(0, m)] = 70  # canonical maximum out-degree (2 * M_pjbp): adjacency stride 72
    nbr_local = rng.integers(0, m, size=int(deg.sum()), dtype=np.int64)
    adj = S[nbr_local].astype(np.uint32)
    offsets = np.zeros(n + 1, np.uint64)
    per_node = np.zeros(n, np.uint64)
    per_node[S] = deg.astype(np.uint64)
    np.cumsum(per_node, out=offsets[1:])
    del per_node
    ep = int(S[rng.integers(0, m)])
    above_32bit = int((S.astype(np.uint64) * np.uint64(dim) >= np.uint64(1 << 32)).sum())

    t1 = time.time()
    ix = capi.Index(base, offsets, adj, ep, metric=capi.METRIC_IP, device=0)
    torch.cuda.synchronize()
    t_index = time.time() - t1
    free_b, total_b = torch.cuda.mem_get_info()

    # host mirror: lazily mapped, only rows of S resident
    try:
        mirror = np.zeros((n, dim), np.float32)  # calloc -> untouched pages are never materialised
    except MemoryError:  # overcommit refused the reservation: a sparse file does the same job
        import tempfile

        mirror = np.memmap(tempfile.NamedTemporaryFile(dir="/tmp", suffix=".mirror"), np.float32, "w+", shape=(n, dim))
    idx = torch.from_numpy(S).to(dev)
    chunk = 100_000
    for lo in range(0, m, chunk):
        mirror[S[lo:lo + chunk]] = base[idx[lo:lo + chunk]].cpu().numpy()

    q = torch.randn((a.queries, dim), generator=g, device=dev, dtype=torch.float32)
    hq = q.cpu().numpy()
    o = Oracle()
    out = dict(n=n, dim=dim, subset=m, rows_beyond_2p32_elements=above_32bit, queries=a.queries, k=a.k, ep=ep,
               max_degree=70, adj_stride=72, base_gb=round(n * dim * 4 / 1e9, 1), adj_gb=round(n * 72 * 4 / 1e9, 1),
               hbm_used_gb=round((total_b - free_b) / 1e9, 1), hbm_total_gb=round(total_b / 1e9, 1),
               seconds=dict(base=round(t_base, 1), index_upload=round(t_index, 1)), checks=[])
    ok_all = True
    for L in a.L:
        got = ix.search(hq, a.k, L)
        want = o.search(mirror, offsets, adj, ep, hq, a.k, L, metric=1)
        same = {key: bool((got[key] == want[key]).all()) for key in ("ids", "cmps", "hops")}
        same["dists"] = bool((got["dists"].view(np.uint32) == want["dists"].view(np.uint32)).all())
        # device timing of the same batch (queries resident), for the record: rows spread over the whole 80 GB
        ids = torch.empty((a.queries, a.k), dtype=torch.int32, device=dev)
        d = torch.empty((a.queries, a.k), dtype=torch.float32, device=dev)
        st = torch.cuda.current_stream().cuda_stream
        for _ in range(2):
            ix.search_device(q, a.k, L, ids, d, stream=st)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ix.search_device(q, a.k, L, ids, d, stream=st)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        out["checks"].append(dict(L=L, bit_identical=same, mean_cmps=round(float(got["cmps"].mean()), 1),
                                  max_id_returned=int(got["ids"].max()), overflow_queries=ix.last_overflow,
                                  ms=round(ms, 3), gathered_gbs=round(float(got["cmps"].sum()) * dim * 4 / (ms * 1e-3) / 1e9, 1)))
        ok_all &= all(same.values())
    out["ok"] = bool(ok_all)
    ix.close()
    print(json.dumps(out), flush=True)
    sys.exit(0 if ok_all else 1)


if __name__ == "__main__":
    main()
