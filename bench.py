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
    ap.add_argument("--cpu-sample", type=int, default=2000, help="queries timed on the CPU baseline")
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
    return ap.parse_args()


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
    tag = f"n{args.n}_t{n_train}_d{args.dim}_s{args.seed}_M{args.M_sq}_{args.M_pjbp}_{args.L_pjpq}_gpu" + ("_unit" if args.normalize else "")
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
        info["knn_parallelism"] = f"base sharded over {world} GPUs, NCCL all-to-all + K4 merge"
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
        t0 = time.time()
        g = capi.Graph(base, knn_ids, M_sq=args.M_sq, M_pjbp=args.M_pjbp, L_pjpq=args.L_pjpq, metric=capi.METRIC_IP)
        info["graph_build_s"] = round(time.time() - t0, 2)
        info["graph_build_phases_s"] = {k: round(v, 2) for k, v in g.phase_seconds.items()}
        del knn_ids
        ep, offsets, adj = g.download()
        io.write_index(index_path + ".tmp", ep, offsets, adj)
        np.savez(index_path + ".csr.tmp.npz", ep=np.uint32(ep), offsets=offsets, adj=adj)  # fast reload for the other ranks
        os.replace(index_path + ".csr.tmp.npz", index_path + ".csr.npz")
        os.replace(index_path + ".tmp", index_path)
        index = capi.Index.from_graph(base, g, metric=capi.METRIC_IP)
        info.update(avg_degree=round(g.nnz / args.n, 2), max_degree=int(g.max_degree), ep=int(ep))
        g.close()
        del offsets, adj
        json.dump(info, open(index_path + ".info.json", "w"))  # build timings travel with the cached index
    elif rank == 0 and os.path.exists(index_path + ".info.json"):
        cached = json.load(open(index_path + ".info.json"))
        cached.update(index_cached=True)
        info = cached
    if world > 1:
        dist.barrier()
    del train
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
    return dict(base=base, queries=q, gt=gt.cpu().numpy().astype(np.uint32), index=index, info=info, index_path=index_path)


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


def load_traffic(args, L):
    """DRAM bytes per K1 launch from the committed `ncu --set full` capture of this same workload (profiles/k1_traffic.json:
    dram__bytes_read.sum + dram__bytes_write.sum of one rg_search_kernel launch); None when the workload differs."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "k1_traffic.json")))
        w = t["workload"]
        n_train = args.train or max(50_000, args.n // 5)
        if (w["n_base"], w["dim"], w["queries"], w["L_pq"], w["k"], w.get("n_train", n_train)) == \
                (args.n, args.dim, args.queries, L, args.k, n_train):
            return int(t["dram_bytes_per_launch"])
    except Exception:
        pass
    return None


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

    # ---- beam width: the smallest L of the sweep whose recall@k reaches the target --------------------
    sweep = []
    L_sel = args.L
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
    if world > 1:  # all ranks time the same L: take the max over ranks
        t = torch.tensor([L_sel], device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
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
        achieved = alg_bytes / (ms_per_step * 1e-3) / 1e9
        out = {
            "metric": "QPS at recall@10=0.9 (IP, OOD queries)", "value": round(value, 1), "unit": "queries/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms_per_step, 4),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{args.n}x{dim} fp32 IP base, {nq} OOD queries per GPU, k={k}, L_pq={L_sel}, "
                                   f"RoarGraph M_sq={args.M_sq} M_pjbp={args.M_pjbp} L_pjpq={args.L_pjpq}",
                       "n_base": args.n, "dim": dim, "queries_per_gpu": nq, "k": k, "L_pq": L_sel,
                       "recall_at_10": round(recall, 4), "recall_sweep": sweep, "mean_cmps": round(sum_cmps / nq, 1),
                       "mean_hops": round(mean_hops, 1), "visited_overflow_queries": n_overflow, "parallelism": f"queries sharded over {world} GPU(s), index replicated",
                       "l2": "256 MiB flush write between timed iterations", "index": d["info"],
                       "ground_truth": "rg_knn_exact_device (exact, FP32 re-ranked)"},
            "e2e": {"value": round(nq * world / (e2e_ms / args.steps * 1e-3), 1), "unit": "queries/s",
                    "h2d_bytes_per_step": nq * dim * 4, "d2h_bytes_per_step": nq * k * 8 + 8,
                    "api": "rg_search_batch (C ABI, pinned host buffers; " + ("kernel reads queries / writes results in place over PCIe"
                                                                                if args.zero_copy else "staged H2D + D2H copies") + ")"},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
                         "frac": round(achieved / peak, 4), "traffic": load_traffic(args, L_sel), "peak_kind": peak_kind,
                         "kernel": "rg_search_kernel", "algorithmic_bytes_per_launch": int(alg_bytes)},
            "clocks": sampler.summary(),
        }
        if not args.no_cpu_baseline and world == 1:  # reported at N=1 only; --impl reference times it at every N
            out["cpu_baseline"] = cpu_baseline(args, d, L_sel)
        print(json.dumps(out), flush=True)
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
                "sample": f"first {sample} of the {self.args.queries} queries, L_pq={L}, OpenMP schedule(dynamic,1) like "
                          "tests/test_search_roargraph.cpp:203"}

    def close(self):
        if self.kind == "reference":
            self.r.close(self.h)


def cpu_baseline(args, d, L):
    ref = CpuReference(args, d)
    sample = min(args.cpu_sample, args.queries)
    res = ref.search(L, sample)
    out = ref.report(L, sample, res)
    ref.close()
    return out


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
    d = prepare(args, 0, 1, device)
    ref = CpuReference(args, d)
    # same beam width rule as our arm: smallest L reaching the recall target (on the CPU sample)
    L_sel = args.L
    sample = min(args.cpu_sample, args.queries)
    if not L_sel:
        for L in L_SWEEP:
            res = ref.search(L, min(sample, 1000))
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
    out = {"impl": "reference", "metric": "QPS at recall@10=0.9 (IP, OOD queries)", "value": round(value, 1),
           "unit": "queries/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": round(t * 1e3, 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": "f32", "data": "synthetic",
           "config": {"workload": f"{args.n}x{args.dim} fp32 IP base, {args.queries} OOD queries, k={args.k}, "
                                  f"L_pq={L_sel}; each step = {sample}-query sample on {cb['cores']} host threads",
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
