"""GPU graph-build timing probe: exact kNN of the training queries + rg_build_roargraph_device, phases printed.
K1 cache hints of the build searches are taken from RG_BUILD_L2_HINT / RG_BUILD_ADJ_PREFETCH / RG_BUILD_WARPS."""
import argparse
import json
import os
import sys
import time

import torch

sys.path.insert(0, ".")
from mysteryann_b200 import build, capi, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=2_000_000)
    ap.add_argument("--train", type=int, default=0)
    ap.add_argument("--dim", type=int, default=200)
    a = ap.parse_args()
    build.build()
    n_train = a.train or a.n // 5
    base, train, _ = synth.make_torch(a.n, n_train, 1, a.dim, device="cuda")
    ids = torch.empty((n_train, 100), dtype=torch.int32, device="cuda")
    d = torch.empty((n_train, 100), dtype=torch.float32, device="cuda")
    capi.knn_exact_device(base, train, 100, ids, d, stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    del d, train
    t0 = time.time()
    g = capi.Graph(base, ids, M_sq=100, M_pjbp=35, L_pjpq=500, metric=capi.METRIC_IP)
    out = dict(n=a.n, n_train=n_train, build_s=round(time.time() - t0, 2), phases={k: round(v, 2) for k, v in g.phase_seconds.items()},
               avg_degree=round(g.nnz / a.n, 2), env={k: os.environ.get(k) for k in ("RG_BUILD_L2_HINT", "RG_BUILD_ADJ_PREFETCH", "RG_BUILD_WARPS")})
    print(json.dumps(out))


if __name__ == "__main__":
    main()
