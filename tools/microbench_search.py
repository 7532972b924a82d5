"""Quick K1 throughput probe on a synthetic random graph (no index build needed): sweeps gather mode /
stage rows / warps and prints achieved gathered-GB/s.  Not the bench - a tuning aid."""
import argparse
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from mysteryann_b200 import build, capi  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=2_000_000)
    ap.add_argument("--dim", type=int, default=200)
    ap.add_argument("--nq", type=int, default=10000)
    ap.add_argument("--deg", type=int, default=48)
    ap.add_argument("--Ls", type=int, nargs="+", default=[20, 100])
    ap.add_argument("--configs", type=str, default="2:0:0:8,2:0:0:9,2:0:0:17,1:0:0:9")
    ap.add_argument("--out", type=str, default="")
    a = ap.parse_args()
    build.build()
    g = torch.Generator(device="cuda"); g.manual_seed(1)
    base = torch.randn(a.n, a.dim, device="cuda", generator=g)
    q = torch.randn(a.nq, a.dim, device="cuda", generator=g)
    rng = np.random.default_rng(0)
    deg = rng.integers(a.deg // 2, a.deg + a.deg // 2, a.n)
    off = np.zeros(a.n + 1, np.uint64); np.cumsum(deg, out=off[1:])
    adj = rng.integers(0, a.n, int(off[-1])).astype(np.uint32)
    ix = capi.Index(base, off, adj, 0, metric=1)
    k = 10
    ids = torch.empty((a.nq, k), dtype=torch.int32, device="cuda"); dists = torch.empty((a.nq, k), device="cuda")
    cmps = torch.empty(a.nq, dtype=torch.int32, device="cuda"); hops = torch.empty_like(cmps)
    st = torch.cuda.current_stream().cuda_stream
    rows = []
    for cfg in a.configs.split(","):
        f = [int(v) for v in cfg.split(":")] + [0] * 7
        gather, warps, ctas, stage, ghash, l2, pf = f[:7]
        ix.configure(gather=gather, warps_per_query=warps, ctas_per_sm=ctas, stage_rows=stage, hash_space=ghash,
                     l2_hint=l2, adj_prefetch=pf)
        for L in a.Ls:
            for _ in range(2):
                ix.search_device(q, k, L, ids, dists, cmps, hops, None, st)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 5
            e0.record()
            for _ in range(reps):
                ix.search_device(q, k, L, ids, dists, cmps, hops, None, st)
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / reps
            c = float(cmps.sum().item()); h = float(hops.sum().item())
            gbs = c * a.dim * 4 / (ms * 1e-3) / 1e9
            row = dict(gather=gather, warps=warps, ctas=ctas, stage=stage, hash_space=ghash, l2_hint=l2, adj_prefetch=pf, L=L, ms=round(ms, 3),
                       qps=round(a.nq / ms * 1e3), mean_cmps=c / a.nq, mean_hops=h / a.nq, gathered_GBs=round(gbs, 1))
            rows.append(row)
            print(json.dumps(row), flush=True)
    if a.out:
        json.dump(rows, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
