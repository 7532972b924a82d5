"""K1 variant sweep on the bench index (10M x 200 by default): one data/index preparation, then every kernel configuration
x beam width, CUDA-event timed with an L2 flush between repetitions.  Prints one JSON line per (config, L) with the
gathered-row bandwidth (cmps x dim x 4 / time) and its fraction of the measured HBM peak.  A tuning aid, not the bench.

    python tools/k1_sweep.py --Ls 55 100 200 500 --configs w=2 w=4 w=2,hs=2 w=2,sb=2 ...
config keys: w warps/query, hs hash_space, sr stage_rows, c ctas/SM, hl hash_log2, l2 l2_hint, pf adj_prefetch, bm batch_mode
"""
import argparse
import json
import os
import subprocess
import sys
import zlib

import torch

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--Ls", type=int, nargs="+", default=[55, 100, 200, 500])
    ap.add_argument("--configs", nargs="+", default=["w=0"])
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--n", type=int, default=10_000_000)
    ap.add_argument("--queries", type=int, default=10_000)
    ap.add_argument("--out", default="")
    a, rest = ap.parse_known_args()
    sys.argv = [sys.argv[0], "--n", str(a.n), "--queries", str(a.queries)] + rest
    args = bench.parse_args()
    from mysteryann_b200 import build

    build.build()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    d = bench.prepare(args, 0, 1, dev)
    ix, q = d["index"], d["queries"]
    nq, k = args.queries, args.k
    ids = torch.empty((nq, k), dtype=torch.int32, device=dev)
    dists = torch.empty((nq, k), dtype=torch.float32, device=dev)
    cmps = torch.empty(nq, dtype=torch.int32, device=dev)
    hops = torch.empty(nq, dtype=torch.int32, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    peak, _ = bench.load_peaks()
    ref = {}
    rows = []
    for cfg in a.configs:
        kv = dict(x.split("=") for x in cfg.split(",") if x)
        g = lambda key, dflt=0: int(kv.get(key, dflt))
        ix.configure(gather=g("g"), warps_per_query=g("w"), ctas_per_sm=g("c"), stage_rows=g("sr"), hash_log2=g("hl"),
                     hash_space=g("hs"), l2_hint=g("l2", 3), adj_prefetch=g("pf", 3), batch_mode=g("bm"))
        for L in a.Ls:
            for _ in range(2):
                ix.search_device(q, k, L, ids, dists, cmps, hops, None, st)
            torch.cuda.synchronize()
            key = (ids.cpu(), cmps.cpu(), hops.cpu())
            if L not in ref:
                ref[L] = key
            same = all(torch.equal(x, y) for x, y in zip(ref[L], key))
            ms = 0.0
            sampler = bench.ClockSampler(0)  # SM clock under load: a long sweep runs into the power cap
            sampler.start()
            for _ in range(a.reps):
                flush.fill_(1)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                ix.search_device(q, k, L, ids, dists, cmps, hops, None, st)
                e1.record()
                torch.cuda.synchronize()
                ms += e0.elapsed_time(e1) / a.reps
            sampler.stop_flag = True
            sampler.join()
            clk = sampler.summary()
            smi = f"{clk['sm_mhz']} MHz {','.join(clk['reasons'])}"
            c = float(cmps.sum().item())
            gbs = c * args.dim * 4 / (ms * 1e-3) / 1e9
            row = dict(cfg=cfg, L=L, ms=round(ms, 3), qps=round(nq / ms * 1e3), mean_cmps=round(c / nq, 1), max_cmps=int(cmps.max().item()),
                       gathered_GBs=round(gbs, 1), frac=round(gbs / peak, 4), overflow=ix.last_overflow, same_as_first_cfg=same, smi=smi,
                       crc=zlib.crc32(key[0].numpy().tobytes() + key[1].numpy().tobytes() + key[2].numpy().tobytes()))  # compare across A/B libraries
            rows.append(row)
            print(json.dumps(row), flush=True)
    if a.out:
        json.dump(rows, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
