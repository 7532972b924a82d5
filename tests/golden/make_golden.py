"""Generates the committed golden fixtures under tests/golden/ by running the UNMODIFIED reference
(oracle/_ref/libroargraph_ref.so, compiled from /root/reference by oracle/Makefile) in the CPU container.
Run:  python tests/golden/make_golden.py      (needs /root/reference; the fixtures then travel with the repo)

All vectors are rounded to fp16-representable values and stored as fp16 so the fixtures stay small;
they are converted back to fp32 exactly on load, so every machine sees bit-identical inputs.

Per case (npz):  base, train, test (fp16) | knn_ids (oracle exact kNN, K = M_sq; FP64-verified) |
                 index (raw bytes of the reference's SaveProjectionGraph after a -T 1 BuildRoarGraph) |
                 for each L: ids_L, dists_L, cmps_L, hops_L  from the reference's SearchRoarGraph
distance.npz  :  pairs of vectors of many lengths + Distance{L2,InnerProduct}::compare outputs
pool.npz      :  random insert / closest_unexpanded scripts + NeighborPriorityQueue final states
"""
import hashlib
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from mysteryann_b200 import io, synth  # noqa: E402
from oracle.binding import Oracle, Ref  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
CASES = {
    # name: (N, NT, NQ, D, metric, M_sq, M_pjbp, L_pjpq, Ls, seed)
    "ip_d200": (1000, 600, 64, 200, 1, 32, 12, 48, (10, 16, 32, 64, 100), 11),
    "l2_d48": (800, 500, 64, 48, 0, 24, 8, 32, (10, 20, 40, 128), 12),
    "ip_d24_norm": (600, 400, 32, 24, 1, 16, 6, 24, (10, 33), 13),
}


def fp16_round(x):
    return x.astype(np.float16).astype(np.float32)


def make_case(name, o, r, tmp):
    N, NT, NQ, D, metric, M_sq, M_pjbp, L_pjpq, Ls, seed = CASES[name]
    base, train, test = synth.make_numpy(N, NT, NQ, D, seed=seed, normalize=name.endswith("norm"))
    base, train, test = fp16_round(base), fp16_round(train), fp16_round(test)
    knn_ids, knn_d, _ = o.exact_knn(base, train, M_sq, metric=metric)
    # FP64 cross-check of the pivot and of the neighbour set (what the build consumes)
    s = train.astype(np.float64) @ base.astype(np.float64).T
    if metric == 0:
        s = (train.astype(np.float64) ** 2).sum(1)[:, None] + (base.astype(np.float64) ** 2).sum(1)[None, :] - 2 * s
    else:
        s = -s
    ref64 = np.argsort(s, axis=1, kind="stable")[:, :M_sq]
    assert (ref64[:, 0] == knn_ids[:, 0]).mean() > 0.995
    p = lambda f: os.path.join(tmp, name + "_" + f)
    io.write_fbin(p("base.fbin"), base)
    io.write_fbin(p("train.fbin"), train)
    io.write_ibin(p("nn.ibin"), knn_ids, knn_d)
    r.build_index(p("base.fbin"), p("train.fbin"), p("nn.ibin"), p("index"), metric=metric, M_sq=M_sq,
                  M_pjbp=M_pjbp, L_pjpq=L_pjpq, threads=1)
    index_bytes = np.fromfile(p("index"), dtype=np.uint8)
    out = dict(base=base.astype(np.float16), train=train.astype(np.float16), test=test.astype(np.float16),
               knn_ids=knn_ids, index=index_bytes,
               params=np.array([metric, M_sq, M_pjbp, L_pjpq], np.uint32), Ls=np.array(Ls, np.uint32))
    h = r.open(p("base.fbin"), p("index"), metric=metric, threads=2)
    for L in Ls:
        res = r.search(h, test, 10, L, threads=2)
        out[f"ids_{L}"] = res["ids"]
        out[f"dists_{L}"] = res["dists"]
        out[f"cmps_{L}"] = res["cmps"]
        out[f"hops_{L}"] = res["hops"]
    r.close(h)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, "index md5", hashlib.md5(index_bytes.tobytes()).hexdigest(), "bytes", index_bytes.size)


def make_distance(r):
    rng = np.random.default_rng(5)
    out = {}
    for d in (1, 2, 3, 4, 5, 7, 8, 9, 12, 15, 16, 17, 23, 24, 31, 32, 40, 48, 100, 128, 200, 203, 512):
        a = fp16_round(rng.standard_normal((32, d)) * 3)
        b = fp16_round(rng.standard_normal((32, d)))
        out[f"a_{d}"] = a.astype(np.float16)
        out[f"b_{d}"] = b.astype(np.float16)
        out[f"l2_{d}"] = r.distance_batch(0, a, b)
        out[f"ip_{d}"] = r.distance_batch(1, a, b)
    np.savez_compressed(os.path.join(HERE, "distance.npz"), **out)


def make_pool(r):
    rng = np.random.default_rng(6)
    out = {}
    for i, (cap, nops) in enumerate([(1, 20), (4, 60), (10, 300), (33, 500), (100, 2000)]):
        kind = (rng.random(nops) < 0.25).astype(np.uint8)
        ids = rng.integers(0, max(8, nops // 3), nops).astype(np.uint32)
        # few distinct distance values -> many (distance, id) ties and duplicate ids
        vals = fp16_round(rng.standard_normal(max(4, nops // 5)))
        dists = vals[ids % len(vals)] if i % 2 == 0 else fp16_round(rng.standard_normal(nops))
        oi, od, of, pop = r.pool_script(cap, kind, ids, dists)
        out.update({f"cap_{i}": np.uint32(cap), f"kind_{i}": kind, f"ids_{i}": ids, f"dists_{i}": dists,
                    f"out_ids_{i}": oi, f"out_dists_{i}": od, f"out_flags_{i}": of, f"pop_{i}": pop})
    np.savez_compressed(os.path.join(HERE, "pool.npz"), **out)


if __name__ == "__main__":
    o, r = Oracle(), Ref()
    with tempfile.TemporaryDirectory() as tmp:
        for name in CASES:
            make_case(name, o, r, tmp)
    make_distance(r)
    make_pool(r)
    print("golden fixtures written to", HERE)
