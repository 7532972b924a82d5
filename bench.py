#!/usr/bin/env python
"""bench.py - RoarGraph search hot path on B200: QPS at recall@10 >= 0.9 on a synthetic cross-modal (OOD) IP set.

    python bench.py --gpus N --steps K --warmup W            # our CUDA path (one rank per GPU under torchrun)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's own CPU search on the host cores

One "step" = one pass of the search hot path over one batch of --queries synthetic OOD queries at the smallest
beam width L whose recall@10 reaches --recall (the BASELINE.json metric).  Prints ONE JSON line (rank 0).
See DESIGN.md "Measurement" for the definition of every field.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# the reference's own beam-width sweep (run_roargraph_search_test.sh:13), cut at 500 like BASELINE.json's "L sweep 10-500"
L_SWEEP = [10, 15, 20, 25, 30, 35, 40, 45, 50, 55, 60, 65, 70, 75, 80, 85, 90, 95, 100, 110, 120, 130, 140, 150, 160, 170,
           180, 190, 200, 220, 240, 260, 280, 300, 350, 400, 450, 500]


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n", type=int, default=int(os.environ.get("RG_BENCH_N", 10_000_000)), help="base vectors")
    ap.add_argument("--train", type=int, default=0, help="training queries for the build (0 = n/5, min 50K)")
    ap.add_argument("--queries", type=int, default=10_000)
    ap.add_argument("--dim", type=int, default=200)
    ap.add_argument("--k", type=int, default=10)
    ap.add_argument("--recall", type=float, default=0.9)
    ap.add_argument("--L", type=int, default=0, help="fixed beam width (0 = smallest L reaching --recall)")
    ap.add_argument("--M_sq", type=int, default=100)
    ap.add_argument("--M_pjbp", type=int, default=35)
    ap.add_argument("--L_pjpq", type=int, default=500)
    ap.add_argument("--seed", type=int, default=20240430)
    ap.add_argument("--cache", default=os.environ.get("RG_BENCH_CACHE", "/tmp/rg_bench_cache"))
    ap.add_argument("--cpu-sample", type=int, default=0, help="queries timed on the CPU baseline (0 = the whole batch)")
    ap.add_argument("--config", default="", choices=["", "C2", "C3", "C3k100", "C5"],
                    help="canonical BASELINE.json workloads: C2 = 10M x 200 (default), C3 = 2.5M x 512 unit rows k=10, "
                         "C3k100 = the same with k=100, C5 = 100M x 200 with 100K-query batches")
    ap.add_argument("--knn-slice", type=int, default=262_144,
                    help="training queries of the timed build-kNN slice (roofline_knn); 0 = skip")
    ap.add_argument("--gt-check", type=int, default=128, help="ground-truth rows cross-checked against the CPU oracle at N=1")
    ap.add_argument("--gather", type=int, default=0)
    ap.add_argument("--stage-rows", type=int, default=0)
    ap.add_argument("--warps", type=int, default=0, help="warps per query (0 = library default)")
    ap.add_argument("--hash-space", type=int, default=0)
    ap.add_argument("--ctas", type=int, default=0, help="resident K1 CTAs per SM (0 = library default)")
    ap.add_argument("--hash-log2", type=int, default=0, help="visited-hash slots per query, log2 (0 = library default)")
    ap.add_argument("--l2-hint", type=int, default=None, help="K1 L2 policy bit mask (None = library default)")
    ap.add_argument("--adj-prefetch", type=int, default=None, help="K1 adjacency prefetch bit mask (None = library default)")
    ap.add_argument("--zero-copy", type=int, default=1, help="e2e: 1 = rg_search_batch works in place on the pinned host buffers, 0 = staged copies")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--normalize", action="store_true", help="L2-normalise all rows (CLIP-like config C3: --n 2500000 --dim 512)")
    a = ap.parse_args()
    canon = {"C2": dict(n=10_000_000, dim=200, queries=10_000, k=10, normalize=False),
             "C3": dict(n=2_500_000, dim=512, queries=10_000, k=10, normalize=True),
             "C3k100": dict(n=2_500_000, dim=512, queries=10_000, k=100, normalize=True),
             "C5": dict(n=100_000_000, dim=200, queries=100_000, k=10, normalize=False)}
    for key, val in canon.get(a.config, {}).items():
        setattr(a, key, val)
    if a.config == "C5" and not a.train:
        a.train = 5_000_000  # the build kNN at 100M rows costs 0.2 s per 1000 training queries on one GPU
    if a.cpu_sample <= 0:
        a.cpu_sample = a.queries
    return a


# ---------------------------------------------------------------------------------------------------------
# data + index preparation (not timed)
# ---------------------------------------------------------------------------------------------------------
def gpu_exact_knn(base, queries, K):
    """Exact inner-product kNN through the C ABI (K2-K4: tcgen05 GEMM + filter, FP32 re-rank, certificate)."""
    import torch

    from mysteryann_b200 import capi

    ids = torch.empty((queries.shape[0], K), dtype=torch.int32, device=base.device)
    dist = torch.empty((queries.shape[0], K), dtype=torch.float32, device=base.device)
    capi.knn_exact_device(base, queries, K, ids, dist, metric=capi.METRIC_IP, stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    return ids, dist


def _sync(t):
    if t.is_cuda:
        import torch

        torch.cuda.synchronize()


def prepare(args, rank, world, device):
    """Returns dict(base (cuda), queries (cuda, this rank's batch), gt (numpy), index (capi.Index), info).
    Rank 0 builds the index on its GPU (exact kNN of the training queries -> rg_build_roargraph_device) and caches the
    index file in --cache; other ranks (and the --impl reference arm) load it."""
    import torch
    import torch.distributed as dist

    from mysteryann_b200 import capi, io, synth

    n_train = args.train or max(50_000, args.n // 5)
    base, train, test = synth.make_torch(args.n, n_train, args.queries * world, args.dim, seed=args.seed, device=device,
                                         normalize=args.normalize)
    tag = f"n{args.n}_t{n_train}_d{args.dim}_s{args.seed}_M{args.M_sq}_{args.M_pjbp}_{args.L_pjpq}_gpu2" + ("_unit" if args.normalize else "")  # gpu2: builder with n/256-node waves
    os.makedirs(args.cache, exist_ok=True)
    index_path = os.path.join(args.cache, tag + ".index")
    info = {"n_train": n_train, "index_cached": os.path.exists(index_path), "builder": "rg_build_roargraph_device (GPU)"}
    index = None
    need_build = not os.path.exists(index_path)
    if world > 1:  # every rank must agree (the file may appear while a slow rank is still generating data)
        t = torch.tensor([int(need_build)], device=device)
        dist.broadcast(t, src=0)
        need_build = bool(t.item())
    knn_ids = None
    if need_build and world > 1:
        # build kNN, base-sharded over the ranks: K2/K3 per shard, NCCL all-to-all of the per-shard lists, K4 merge,
        # all-gather of the merged slices (mysteryann_b200/sharded_knn.py); rank 0 keeps the ids for the graph build
        from mysteryann_b200 import sharded_knn

        b = sharded_knn.shard_bounds(args.n, world)
        torch.cuda.synchronize()
        dist.barrier()
        t0 = time.time()
        knn_ids, _, _ = sharded_knn.knn_sharded(base[b[rank]:b[rank + 1]], b[rank], train, args.M_sq, metric=capi.METRIC_IP,
                                                stream=torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        dist.barrier()
        info["knn_s"] = round(time.time() - t0, 2)
        info["knn_tflops"] = round(2.0 * args.n * n_train * args.dim / (time.time() - t0) / 1e12, 1)
        info["knn_parallelism"] = (f"base sharded over {world} GPUs, rg_knn_exact_sharded + all-gather of the merged lists; first call "
                                   "(includes NCCL communicator set-up and scratch allocation; wall clock) - roofline_knn is the timed figure")
        if rank != 0:
            knn_ids = None
    if rank == 0 and need_build:
        if knn_ids is None:
            torch.cuda.synchronize()  # the data generation kernels are not part of the kNN time
            t0 = time.time()
            knn_ids, _ = gpu_exact_knn(base, train, args.M_sq)
            info["knn_s"] = round(time.time() - t0, 2)
            info["knn_tflops"] = round(2.0 * args.n * n_train * args.dim / (time.time() - t0) / 1e12, 1)
            info["knn_parallelism"] = "1 GPU"
        st = capi.knn_last_stats()
        info["knn_exact_scans"] = st["exact_scans"]
        capi.knn_release_scratch()  # the FP16 base copy (42 GB at 100M rows) must not sit next to the build's two adjacency arrays
        t0 = time.time()
        g = capi.Graph(base, knn_ids, M_sq=args.M_sq, M_pjbp=args.M_pjbp, L_pjpq=args.L_pjpq, metric=capi.METRIC_IP)
        info["graph_build_s"] = round(time.time() - t0, 2)
        info["graph_build_phases_s"] = {k: round(v, 2) for k, v in g.phase_seconds.items()}
        del knn_ids
        if args.n <= 50_000_000:  # the 100M-row index (13 GB of files) is used in place, single GPU, no CPU arm
            ep, offsets, adj = g.download()
            io.write_index(index_path + ".tmp", ep, offsets, adj)
            np.savez(index_path + ".csr.tmp.npz", ep=np.uint32(ep), offsets=offsets, adj=adj)  # fast reload for the other ranks
            os.replace(index_path + ".csr.tmp.npz", index_path + ".csr.npz")
            os.replace(index_path + ".tmp", index_path)
        else:
            ep, offsets, adj = g.ep, None, None
            info["index_file"] = "not written (used in place)"
        index = capi.Index.from_graph(base, g, metric=capi.METRIC_IP)
        info.update(avg_degree=round(g.nnz / args.n, 2), max_degree=int(g.max_degree), ep=int(ep))
        g.close()
        del offsets, adj
        if args.n <= 50_000_000:
            json.dump(info, open(index_path + ".info.json", "w"))  # build timings travel with the cached index
    elif rank == 0 and os.path.exists(index_path + ".info.json"):
        cached = json.load(open(index_path + ".info.json"))
        cached.update(index_cached=True)
        info = cached
    if world > 1:
        dist.barrier()
    knn_slice = train[:min(args.knn_slice, n_train)].clone() if args.knn_slice else None
    del train
    capi.knn_release_scratch()
    torch.cuda.empty_cache()
    if index is None:
        if os.path.exists(index_path + ".csr.npz"):
            z = np.load(index_path + ".csr.npz")
            ep, offsets, adj = int(z["ep"]), z["offsets"], z["adj"]
        else:
            ep, offsets, adj = io.read_index(index_path)
        index = capi.Index(base, offsets, adj, ep, metric=capi.METRIC_IP, device=device.index or 0)
        info.update(avg_degree=round(len(adj) / args.n, 2), max_degree=int(np.diff(offsets).max()), ep=int(ep))
        del offsets, adj
    q = test[rank * args.queries:(rank + 1) * args.queries].contiguous()
    gt, _ = gpu_exact_knn(base, q, args.k)
    capi.knn_release_scratch()
    return dict(base=base, queries=q, gt=gt.cpu().numpy().astype(np.uint32), index=index, info=info, index_path=index_path,
                knn_slice=knn_slice)


def metric_name(args):
    return f"QPS at recall@{args.k}={args.recall:g} (IP, OOD queries)"


def recall_at_k(ids, gt, k):
    hit = 0
    for a, b in zip(ids[:, :k], gt[:, :k]):
        hit += len(set(a.tolist()) & set(b.tolist()))
    return hit / (k * len(ids))


class ClockSampler(threading.Thread):
    """nvidia-smi style clock / throttle-reason samples DURING the timed region (pynvml)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.samples, self.reasons, self.stop_flag, self.max_mhz = [], set(), False, None
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if not self.nv:
            return
        nv = self.nv
        names = {nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
                 nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap"}
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.02)

    def summary(self):
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


def load_peaks():
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(p["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


def load_tensor_peaks():
    """(burst, sustained) dense bf16/fp16 TFLOP/s: MEASURED_PEAKS.json, else the profiling guide's fallback."""
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(p["bf16_tflops"]), float(p.get("bf16_tflops_sustained", p["bf16_tflops"])), "measured"
    except Exception:
        return 1640.9, 1378.6, "fallback"


def k1_source_hash():
    """Fingerprint of the K1 sources: a traffic figure taken from a committed ncu capture is only quoted while the kernel
    it was captured from is the kernel that runs."""
    import hashlib

    h = hashlib.sha256()
    for f in ("rg_search.cu", "rg_distance.cuh", "rg_common.cuh"):
        h.update(open(os.path.join(ROOT, "mysteryann_b200", "csrc", f), "rb").read())
    return h.hexdigest()[:16]


def load_traffic(args, L):
    """DRAM bytes per K1 launch (dram__bytes_read.sum + dram__bytes_write.sum of one rg_search_kernel launch) from the
    committed `ncu --set full` capture of this same workload, profiles/k1_traffic.json (written by tools/make_k1_traffic.py).
    ncu cannot run inside the timed bench, so the figure is FROM A PROFILE: it is quoted only when the workload matches and
    the profile's K1 source fingerprint equals the running kernel's; otherwise (None, reason)."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "k1_traffic.json")))
        w = t["workload"]
        n_train = args.train or max(50_000, args.n // 5)
        if (w["n_base"], w["dim"], w["queries"], w["L_pq"], w["k"], w.get("n_train", n_train)) != \
                (args.n, args.dim, args.queries, L, args.k, n_train):
            return None, "profile is of another workload"
        if t.get("k1_source_hash") != k1_source_hash():
            return None, "profile is of another kernel version"
        return int(t["dram_bytes_per_launch"]), "from_profile:profiles/k1_traffic.json"
    except Exception as e:  # noqa: BLE001
        return None, f"no profile ({type(e).__name__})"


def time_knn_slice(args, d, rank, world, device):
    """Build kNN (BASELINE.json's second metric) on a fixed slice: the first --knn-slice training queries against the whole
    base, K = M_sq.  N = 1: rg_knn_exact_device; N > 1: the base sharded over the ranks, rg_knn_exact_sharded (K2/K3 per
    shard, grouped ncclSend/ncclRecv, K4 merge).  CUDA events on the launching stream, max over ranks; one untimed
    warm-up call (scratch allocation, NCCL channels)."""
    import torch
    import torch.distributed as dist

    from mysteryann_b200 import capi, sharded_knn

    q = d["knn_slice"]
    if q is None or q.shape[0] == 0:
        return None
    base, K = d["base"], args.M_sq
    st = torch.cuda.current_stream().cuda_stream
    b = sharded_knn.shard_bounds(args.n, world)

    def run(qq):
        if world == 1:
            ids = torch.empty((qq.shape[0], K), dtype=torch.int32, device=device)
            dd = torch.empty((qq.shape[0], K), dtype=torch.float32, device=device)
            capi.knn_exact_device(base, qq, K, ids, dd, metric=capi.METRIC_IP, stream=st)
            return ids, dd
        ids, dd, _ = sharded_knn.knn_sharded(base[b[rank]:b[rank + 1]], b[rank], qq, K, metric=capi.METRIC_IP, gather=False,
                                             stream=st)
        return ids, dd

    run(q[:min(65536, q.shape[0])].contiguous())
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    # best of three repetitions: on the shared boxes of this pool a single repetition occasionally takes 1.5-2x (host-side
    # hiccups between the ~100 launches of a call); every repetition is listed in the line
    reps_ms = []
    for _ in range(3):
        if world > 1:
            dist.barrier()
        e0.record()
        ids, _ = run(q)
        e1.record()
        torch.cuda.synchronize()
        r_ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([r_ms], device=device, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)   # a repetition takes as long as its slowest rank
            r_ms = t.item()
        reps_ms.append(r_ms)
    stats = capi.knn_last_stats()
    ms = min(reps_ms)
    # the merged slices must equal the unsharded kernels' answer (first rows of this rank's slice)
    same = 1
    if world > 1:
        lo, hi = capi.knn_sharded_slice(q.shape[0], rank, world)
        m = min(512, hi - lo)
        want = torch.empty((m, K), dtype=torch.int32, device=device)
        wd = torch.empty((m, K), dtype=torch.float32, device=device)
        capi.knn_exact_device(base, q[lo:lo + m].contiguous(), K, want, wd, metric=capi.METRIC_IP, stream=st)
        torch.cuda.synchronize()
        same = int(torch.equal(want, ids[:m]))
        t = torch.tensor([ms, float(1 - same), float(stats["second_pass"]), float(stats["exact_scans"])], device=device,
                         dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        same = int(t[1].item() == 0)
        stats = dict(stats, second_pass_max_rank=int(t[2].item()), exact_scans_max_rank=int(t[3].item()))
    burst, sustained, kind = load_tensor_peaks()
    nq = q.shape[0]
    flops = 2.0 * nq * args.n * args.dim
    # Same slice in the grid layout of rg_knn_exact_grid: 2 base shards x N/2 query groups.  The base still is sharded and
    # the per-shard lists still are exchanged over NCCL and merged, but a shard keeps 1/2 of the rows instead of 1/N, which
    # keeps K2 in its efficient regime (DESIGN.md "K2").  Reported next to the canonical one-shard-per-GPU figure.
    grid = None
    if world >= 4 and world % 2 == 0:
        (b0, b1), (g0, g1), (o0, o1) = sharded_knn.grid_layout(rank, world, 2, args.n, nq)
        shard = base[b0:b1]

        def run_grid(qq):
            (_, _), (h0, h1), _ = sharded_knn.grid_layout(rank, world, 2, args.n, qq.shape[0])
            return sharded_knn.knn_grid(shard, b0, qq[h0:h1].contiguous(), K, 2, metric=capi.METRIC_IP, stream=st)

        run_grid(q[:min(65536, nq)].contiguous())
        torch.cuda.synchronize()
        g_reps = []
        for _ in range(3):
            dist.barrier()
            e0.record()
            gids, _ = run_grid(q)
            e1.record()
            torch.cuda.synchronize()
            tt = torch.tensor([e0.elapsed_time(e1)], device=device, dtype=torch.float64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            g_reps.append(tt.item())
        m = min(512, o1 - o0)
        want = torch.empty((m, K), dtype=torch.int32, device=device)
        wd = torch.empty((m, K), dtype=torch.float32, device=device)
        capi.knn_exact_device(base, q[o0:o0 + m].contiguous(), K, want, wd, metric=capi.METRIC_IP, stream=st)
        torch.cuda.synchronize()
        t = torch.tensor([0.0, float(1 - int(torch.equal(want, gids[:m])))], device=device, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        gms = min(g_reps)
        gper = flops / (gms * 1e-3) / 1e12 / world
        grid = {"layout": f"2 base shards x {world // 2} query groups, rg_knn_exact_grid", "shard_rows": b1 - b0,
                "knn_s": round(gms * 1e-3, 4), "repetitions_s": [round(x * 1e-3, 4) for x in g_reps], "achieved": round(gper, 1), "frac": round(gper / burst, 4),
                "c4_extrapolated_s": round(gms * 1e-3 * 10_000_000 / nq * (10_000_000 / args.n), 2),
                "equals_unsharded": bool(t[1].item() == 0)}
    capi.knn_release_scratch()
    torch.cuda.empty_cache()
    per_gpu = flops / (ms * 1e-3) / 1e12 / world
    return {"bound": "tensor", "achieved": round(per_gpu, 1), "peak": burst, "unit": "TFLOP/s per GPU",
            "frac": round(per_gpu / burst, 4), "frac_of_sustained_peak": round(per_gpu / sustained, 4), "peak_kind": kind,
            "kernel": "knn_gemm_filter_kernel (tcgen05 kind::f16) + select + FP32 re-rank" + (" + NCCL exchange + K4 merge" if world > 1 else ""),
            "n_ranks": world, "shard_rows": b[1] - b[0], "queries": nq, "K": K, "knn_s": round(ms * 1e-3, 4),
            "timing": "best of 3 repetitions, CUDA events, max over ranks per repetition", "repetitions_s": [round(x * 1e-3, 4) for x in reps_ms],
            "algorithmic_flops": flops, "c4_extrapolated_s": round(ms * 1e-3 * 10_000_000 / nq * (10_000_000 / args.n), 2),
            "parallelism": "1 GPU" if world == 1 else f"base sharded over {world} GPUs, rg_knn_exact_sharded (grouped ncclSend/ncclRecv + K4 merge)",
            "sharded_equals_unsharded": bool(same) if world > 1 else None, "knn_stats": stats,
            **({"grid": grid} if grid else {})}


# ---------------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist

    from mysteryann_b200 import build, capi

    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    if rank == 0:
        build.build()
    if world > 1:
        dist.barrier()
    d = prepare(args, rank, world, device)
    roofline_knn = time_knn_slice(args, d, rank, world, device)  # BASELINE.json's "build kNN s", same slice at every N
    d["knn_slice"] = None
    nq, k, dim = args.queries, args.k, args.dim
    ix = d["index"]
    ix.configure(gather=args.gather, warps_per_query=args.warps, ctas_per_sm=args.ctas, stage_rows=args.stage_rows,
                 hash_log2=args.hash_log2, hash_space=args.hash_space,
                 l2_hint=args.l2_hint, adj_prefetch=args.adj_prefetch)
    ix.set_option("zero_copy", args.zero_copy)
    q = d["queries"]
    ids = torch.empty((nq, k), dtype=torch.int32, device=device)
    dists = torch.empty((nq, k), dtype=torch.float32, device=device)
    cmps = torch.empty(nq, dtype=torch.int32, device=device)
    hops = torch.empty(nq, dtype=torch.int32, device=device)
    status = torch.zeros(2, dtype=torch.int32, device=device)
    stream = torch.cuda.current_stream().cuda_stream

    def search(L):
        ix.search_device(q, k, L, ids, dists, cmps, hops, status, stream)

    # ---- beam width: the smallest L of the sweep whose recall@k reaches the target, decided ONCE on rank 0's query set and
    # used by every rank (round 1 took the max over per-rank choices, so N=4 timed another L than N=1/2/8)
    sweep = []
    L_sel = args.L
    if rank == 0:
        for L in ([args.L] if args.L else L_SWEEP):
            if L < k:
                continue
            search(L)
            torch.cuda.synchronize()
            r = recall_at_k(ids.cpu().numpy().view(np.uint32), d["gt"], k)
            sweep.append((L, round(r, 4)))
            if not args.L and r >= args.recall:
                L_sel = L
                break
        if not L_sel:
            L_sel = L_SWEEP[-1]
    if world > 1:
        t = torch.tensor([L_sel], device=device)
        dist.broadcast(t, src=0)
        L_sel = int(t.item())
    search(L_sel)
    torch.cuda.synchronize()
    recall = recall_at_k(ids.cpu().numpy().view(np.uint32), d["gt"], k)
    n_overflow = ix.last_overflow
    sum_cmps = float(cmps.sum().item())
    mean_hops = float(hops.float().mean().item())
    assert status.cpu().tolist() == [0, 0], "search reported short/overflowed queries"

    flush = torch.empty(256 << 20, dtype=torch.uint8, device=device)  # > 126 MB L2

    def timed_loop(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        for e0, e1 in evs:
            flush.fill_(1)  # L2 flush between timed iterations (outside the event pair)
            e0.record()
            fn()
            e1.record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        return sum(e0.elapsed_time(e1) for e0, e1 in evs)  # ms of the K steps on this rank

    sampler = ClockSampler(local)
    sampler.start()
    launches0 = ix.launches
    dev_ms = timed_loop(lambda: search(L_sel), args.steps, args.warmup)
    launches = (ix.launches - launches0) * args.steps // (args.steps + args.warmup)

    # ---- e2e: the C-ABI host-buffer call (pinned host queries in, pinned host results out) -----------
    hq = q.cpu().pin_memory()
    h_ids = torch.empty((nq, k), dtype=torch.int32).pin_memory()
    h_dists = torch.empty((nq, k), dtype=torch.float32).pin_memory()

    def e2e_step():
        ix.search_raw(hq.data_ptr(), nq, k, L_sel, h_ids.data_ptr(), h_dists.data_ptr())

    for _ in range(args.warmup):
        e2e_step()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()  # synchronous: returns when the results are in host memory
    e2e_ms = (time.perf_counter() - t0) * 1e3
    sampler.stop_flag = True
    sampler.join()
    assert (h_ids.numpy() == ids.cpu().numpy()).all(), "e2e path and device path disagree"

    if world > 1:
        t = torch.tensor([dev_ms, e2e_ms], device=device, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms, e2e_ms = t.tolist()
        s = torch.tensor([sum_cmps, recall], device=device, dtype=torch.float64)
        dist.all_reduce(s, op=dist.ReduceOp.SUM)
        sum_cmps_all, recall = s[0].item(), s[1].item() / world
    else:
        sum_cmps_all = sum_cmps

    if rank == 0:
        peak, peak_kind = load_peaks()
        ms_per_step = dev_ms / args.steps
        value = nq * world / (ms_per_step * 1e-3)
        alg_bytes = sum_cmps * dim * 4  # this rank's launch: gathered vector bytes
        traffic, traffic_src = load_traffic(args, L_sel)
        achieved = alg_bytes / (ms_per_step * 1e-3) / 1e9
        out = {
            "metric": metric_name(args), "value": round(value, 1), "unit": "queries/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms_per_step, 4),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{args.n}x{dim} fp32 IP base, {nq} OOD queries per GPU, k={k}, L_pq={L_sel}, "
                                   f"RoarGraph M_sq={args.M_sq} M_pjbp={args.M_pjbp} L_pjpq={args.L_pjpq}",
                       "n_base": args.n, "dim": dim, "queries_per_gpu": nq, "k": k, "L_pq": L_sel,
                       "recall_at_10" if k == 10 else f"recall_at_{k}": round(recall, 4), "recall_sweep": sweep, "mean_cmps": round(sum_cmps / nq, 1),
                       "mean_hops": round(mean_hops, 1), "visited_overflow_queries": n_overflow, "parallelism": f"queries sharded over {world} GPU(s), index replicated",
                       "l2": "256 MiB flush write between timed iterations", "index": d["info"],
                       "ground_truth": "rg_knn_exact_device (exact, FP32 re-ranked)"},
            "e2e": {"value": round(nq * world / (e2e_ms / args.steps * 1e-3), 1), "unit": "queries/s",
                    "h2d_bytes_per_step": nq * dim * 4, "d2h_bytes_per_step": nq * k * 8 + 8,
                    "api": "rg_search_batch (C ABI, pinned host buffers; " + ("kernel reads queries / writes results in place over PCIe"
                                                                                if args.zero_copy else "staged H2D + D2H copies") + ")"},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
                         "frac": round(achieved / peak, 4), "traffic": traffic, "traffic_source": traffic_src,
                         "peak_kind": peak_kind, "frac_of_nominal_8TBs": round(achieved / 8000.0, 4),
                         "kernel": "rg_search_kernel", "algorithmic_bytes_per_launch": int(alg_bytes)},
            "clocks": sampler.summary(),
        }
        if achieved > 0.95 * peak:
            # the denominator is the measured COPY bandwidth (half reads, half writes, bus turnarounds included); K1's traffic is
            # > 95 % reads, and hub rows next to the entry point are shared by the queries of a batch through L2
            out["roofline"]["note"] = ("peak is the measured copy bandwidth (reads + writes); K1 is read-dominated and a few hub rows "
                                       "are served from L2, so the gathered-row rate can come close to or pass it; "
                                       "frac_of_nominal_8TBs is the fraction of the HBM3e nominal rate")
        if args.config:
            out["config"]["canonical"] = args.config
        if roofline_knn is not None:
            out["roofline_knn"] = roofline_knn
            out["config"]["knn_s"] = roofline_knn["knn_s"]
            out["config"]["knn_parallelism"] = roofline_knn["parallelism"]
        if not args.no_cpu_baseline and world == 1:  # reported at N=1 only; --impl reference times it at every N
            res_gpu = dict(ids=ids.cpu().numpy().view(np.uint32), dists=dists.cpu().numpy(),
                           cmps=cmps.cpu().numpy().view(np.uint32), hops=hops.cpu().numpy().view(np.uint32))
            out["cpu_baseline"], out["parity"] = cpu_baseline(args, d, L_sel, res_gpu)
        print(json.dumps(out), flush=True)
        if "parity" in out and not out["parity"]["ok"]:
            raise SystemExit("PARITY FAILURE: the GPU results differ from the reference's on the same inputs: "
                             + json.dumps(out["parity"]))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


class CpuReference:
    """The reference's CPU search (oracle/_ref when it is loadable, else the C port) on the host cores.  Opened once:
    the base is written as an fbin next to the cached index and loaded by the reference's own loader."""

    def __init__(self, args, d):
        from mysteryann_b200 import io
        from oracle.binding import Oracle, Ref, ref_available

        self.args, self.d = args, d
        self.queries = d["queries"].cpu().numpy()
        if ref_available():
            self.kind, self.r = "reference", Ref()
            self.threads = self.r.num_procs()
            fb = os.path.join(args.cache, f"base_n{args.n}_d{args.dim}_s{args.seed}{'_unit' if args.normalize else ''}.fbin")
            if not os.path.exists(fb):
                io.write_fbin(fb + ".tmp", d["base"].cpu().numpy())
                os.replace(fb + ".tmp", fb)
            self.h = self.r.open(fb, d["index_path"], metric=1, threads=self.threads)
        else:
            self.kind, self.o = "port", Oracle()
            self.threads = self.o.num_procs()
            self.base = d["base"].cpu().numpy()
            self.ep, self.offsets, self.adj = io.read_index(d["index_path"])

    def search(self, L, sample):
        q = self.queries[:sample]
        if self.kind == "reference":
            return self.r.search(self.h, q, self.args.k, L, threads=self.threads, warmup=True)
        return self.o.search(self.base, self.offsets, self.adj, self.ep, q, self.args.k, L, metric=1, threads=self.threads)

    def report(self, L, sample, res):
        return {"value": round(sample / res["seconds"], 1), "unit": "queries/s", "cores": self.threads, "kind": self.kind,
                "sample": (f"all {sample} queries" if sample == self.args.queries else f"first {sample} of the {self.args.queries} queries")
                          + f", L_pq={L}, one cold pass, OpenMP schedule(dynamic,1) like tests/test_search_roargraph.cpp:203"}

    def close(self):
        if self.kind == "reference":
            self.r.close(self.h)


def compare_results(got, want):
    """Bit-for-bit comparison of two result sets (ids, dists as bit patterns, cmps, hops) -> parity dict."""
    n = len(want["ids"])
    par = {"n": int(n)}
    ok = True
    for key in ("ids", "cmps", "hops"):
        eq = bool((got[key][:n] == want[key][:n]).all())
        par[key] = eq
        ok = ok and eq
    gb = np.ascontiguousarray(got["dists"][:n], np.float32).view(np.uint32)
    wb = np.ascontiguousarray(want["dists"][:n], np.float32).view(np.uint32)
    par["dists_bits"] = bool((gb == wb).all())
    par["ok"] = ok and par["dists_bits"]
    if not par["ok"]:
        bad = np.argwhere((got["ids"][:n] != want["ids"][:n]).any(axis=1)).ravel()
        par["first_bad_queries"] = [int(b) for b in bad[:5]]
    return par


def check_ground_truth(args, d, rows):
    """The bench's recall is computed against ground truth from the repo's own K2-K4 kernels; cross-check its first `rows`
    lists against the CPU oracle's exact_knn (FP32 brute force, compute_groundtruth.cpp:126-248 restated) on the full base.
    ids must agree except where the two FP32 scores are within 1e-6 relative (the tie rule of BASELINE.json)."""
    from oracle.binding import Oracle

    o = Oracle()
    base = d["base"].cpu().numpy()
    q = d["queries"][:rows].cpu().numpy()
    want_ids, want_d, sec = o.exact_knn(base, q, args.k, metric=1)
    gt = d["gt"][:rows, :args.k]
    diff = np.argwhere(gt != want_ids)
    # a differing position is acceptable only on a near-tie: compare the oracle's score at that rank with the exact score
    # of the id the GPU put there
    worst = 0.0
    for qi, j in diff:
        mine = float(np.dot(base[gt[qi, j]].astype(np.float64), q[qi].astype(np.float64)))
        ref = float(want_d[qi, j])
        worst = max(worst, abs(mine - ref) / max(1.0, abs(ref)))
    return {"rows": int(rows), "id_mismatches": int(len(diff)), "max_rel_score_gap_at_mismatch": worst,
            "ok": bool(worst <= 1e-6), "oracle_seconds": round(sec, 2)}


def cpu_baseline(args, d, L, res_gpu=None):
    ref = CpuReference(args, d)
    sample = min(args.cpu_sample, args.queries)
    res = ref.search(L, sample)
    out = ref.report(L, sample, res)
    parity = None
    if res_gpu is not None:
        # the protocol of tests/test_search_roargraph.cpp:196-214 on the same index file and queries: the reference's
        # ids / distances / (cmps, hops) against ours, every query of the sample, bit for bit
        parity = compare_results(res_gpu, res)
        parity["against"] = f"{ref.kind} SearchRoarGraph, first {sample} queries, L_pq={L}"
        if args.gt_check:
            parity["ground_truth_vs_oracle"] = check_ground_truth(args, d, min(args.gt_check, args.queries))
            parity["ok"] = parity["ok"] and parity["ground_truth_vs_oracle"]["ok"]
    ref.close()
    return out, parity


def run_reference(args):
    """--impl reference: the reference's own CPU search loop on the host cores, same config/metric."""
    import torch

    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    if rank != 0:
        return
    assert torch.cuda.is_available(), "the index for the reference arm is prepared on the GPU (data prep, untimed)"
    from mysteryann_b200 import build

    build.build()
    device = torch.device("cuda", 0)
    args.knn_slice = 0
    d = prepare(args, 0, 1, device)
    ref = CpuReference(args, d)
    # same beam width rule as our arm: smallest L of the sweep reaching the recall target on the same query batch
    L_sel = args.L
    sample = min(args.cpu_sample, args.queries)
    if not L_sel:
        for L in L_SWEEP:
            if L < args.k:
                continue
            res = ref.search(L, sample)
            if recall_at_k(res["ids"], d["gt"][:len(res["ids"])], args.k) >= args.recall:
                L_sel = L
                break
        L_sel = L_sel or L_SWEEP[-1]
    times = []
    for _ in range(args.warmup + args.steps):
        res = ref.search(L_sel, sample)
        times.append(res["seconds"])
    t = float(np.mean(times[args.warmup:]))
    value = sample / t
    cb = ref.report(L_sel, sample, res)
    cb["value"] = round(value, 1)
    ref.close()
    out = {"impl": "reference", "metric": metric_name(args), "value": round(value, 1),
           "unit": "queries/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": round(t * 1e3, 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": "f32", "data": "synthetic",
           "config": {"workload": f"{args.n}x{args.dim} fp32 IP base, {args.queries} OOD queries, k={args.k}, "
                                  f"L_pq={L_sel}; each step = " + ("the whole batch" if sample == args.queries else f"a {sample}-query sample")
                                  + f" on {cb['cores']} host threads",
                      "same_config": sample == args.queries,
                      "n_base": args.n, "dim": args.dim, "k": args.k, "L_pq": L_sel, "index": d["info"]},
           "cpu_baseline": cb,
           "e2e": {"value": round(value, 1), "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    a = parse_args()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
