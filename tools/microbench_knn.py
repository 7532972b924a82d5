"""K2-K4 throughput probe: exact kNN of NQ synthetic queries against N base rows on one GPU."""
import argparse
import json
import sys

import torch

sys.path.insert(0, ".")
from mysteryann_b200 import build, capi, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=1_000_000)
    ap.add_argument("--nq", type=int, default=131072)
    ap.add_argument("--dim", type=int, default=200)
    ap.add_argument("--K", type=int, default=100)
    ap.add_argument("--reps", type=int, default=3)
    a = ap.parse_args()
    build.build()
    base, train, _ = synth.make_torch(a.n, a.nq, 1, a.dim, device="cuda")
    ids = torch.empty((a.nq, a.K), dtype=torch.int32, device="cuda")
    d = torch.empty((a.nq, a.K), dtype=torch.float32, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    capi.knn_exact_device(base, train, a.K, ids, d, stream=st)
    torch.cuda.synchronize()
    per_rep = []
    for _ in range(a.reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        capi.knn_exact_device(base, train, a.K, ids, d, stream=st)
        e1.record()
        torch.cuda.synchronize()
        per_rep.append(round(e0.elapsed_time(e1), 2))
    ms = sum(per_rep) / a.reps
    flops = 2.0 * a.n * a.nq * a.dim
    peaks = json.load(open("MEASURED_PEAKS.json")) if __import__("os").path.exists("MEASURED_PEAKS.json") else {}
    tf = flops / (ms * 1e-3) / 1e12
    print(json.dumps(dict(n=a.n, nq=a.nq, dim=a.dim, K=a.K, ms=round(ms, 2), tflops_algorithmic=round(tf, 1),
                          frac_of_bf16_burst=round(tf / peaks.get("bf16_tflops", 1640.9), 4),
                          frac_of_bf16_sustained=round(tf / peaks.get("bf16_tflops_sustained", 1378.6), 4),
                          per_rep_ms=per_rep, stats=capi.knn_last_stats())))


if __name__ == "__main__":
    main()
