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

L_SWEEP = [10, 12, 14, 16, 18, 20, 22, 24, 26, 28, 30, 35, 40, 45, 50, 60, 70, 80, 90, 100, 120, 150, 200, 300, 500]


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n", type=int, default=int(os.environ.get("RG_BENCH_N", 500_000)), help="base vectors")
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
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


# ---------------------------------------------------------------------------------------------------------
# data + index preparation (not timed)
# ---------------------------------------------------------------------------------------------------------
def torch_exact_knn(base, queries, K, chunk=8192):
    """Bootstrap exact kNN (inner product) for ground truth / the learn->base file until K2 (tcgen05) lands:
    FP32 matmul (TF32 off) + top-k on the GPU.  Data preparation, never inside a timed region."""
    import torch

    torch.backends.cuda.matmul.allow_tf32 = False
    ids = torch.empty((queries.shape[0], K), dtype=torch.int64, device=base.device)
    dist = torch.empty((queries.shape[0], K), dtype=torch.float32, device=base.device)
    bchunk = 1 << 20
    for s in range(0, queries.shape[0], chunk):
        q = queries[s:s + chunk]
        best_v = best_i = None
        for b0 in range(0, base.shape[0], bchunk):
            sc = q @ base[b0:b0 + bchunk].T
            v, i = sc.topk(min(K, sc.shape[1]), dim=1)
            i += b0
            if best_v is None:
                best_v, best_i = v, i
            else:
                v2 = torch.cat([best_v, v], 1)
                i2 = torch.cat([best_i, i], 1)
                best_v, sel = v2.topk(K, dim=1)
                best_i = i2.gather(1, sel)
        ids[s:s + chunk], dist[s:s + chunk] = best_i, best_v
    return ids, dist


def _sync(t):
    if t.is_cuda:
        import torch

        torch.cuda.synchronize()


def prepare(args, rank, world, device):
    """Returns dict(base (cuda), queries (cuda, this rank's batch), gt (numpy), offsets, adj, ep).  Rank 0 builds the
    index once (host CPU BuildRoarGraph on all cores) and caches it in --cache; other ranks load it."""
    import torch
    import torch.distributed as dist

    from mysteryann_b200 import hostlib, io, synth

    n_train = args.train or max(50_000, args.n // 5)
    base, train, test = synth.make_torch(args.n, n_train, args.queries * world, args.dim, seed=args.seed, device=device)
    tag = f"n{args.n}_t{n_train}_d{args.dim}_s{args.seed}_M{args.M_sq}_{args.M_pjbp}_{args.L_pjpq}"
    os.makedirs(args.cache, exist_ok=True)
    index_path = os.path.join(args.cache, tag + ".index")
    info = {"n_train": n_train, "index_cached": os.path.exists(index_path)}
    if rank == 0 and not os.path.exists(index_path):
        t0 = time.time()
        knn_ids, _ = torch_exact_knn(base, train, args.M_sq)
        _sync(base)
        info["knn_s"] = round(time.time() - t0, 2)
        t0 = time.time()
        hostlib.build()
        hostlib.build_index(base.cpu().numpy(), train.cpu().numpy(), knn_ids.cpu().numpy().astype(np.uint32),
                            index_path + ".tmp", metric=1, M_sq=args.M_sq, M_pjbp=args.M_pjbp, L_pjpq=args.L_pjpq,
                            threads=os.cpu_count() or 1)
        os.replace(index_path + ".tmp", index_path)
        info["graph_build_s"] = round(time.time() - t0, 2)
    if world > 1:
        dist.barrier()
    del train
    ep, offsets, adj = io.read_index(index_path)
    q = test[rank * args.queries:(rank + 1) * args.queries].contiguous()
    gt, _ = torch_exact_knn(base, q, args.k)
    _sync(base)
    info.update(avg_degree=round(len(adj) / args.n, 2), max_degree=int(np.diff(offsets).max()), ep=int(ep))
    return dict(base=base, queries=q, gt=gt.cpu().numpy().astype(np.uint32), offsets=offsets, adj=adj, ep=ep,
                info=info, index_path=index_path)


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
    ix = capi.Index(d["base"], d["offsets"], d["adj"], d["ep"], metric=capi.METRIC_IP, device=local)
    ix.configure(gather=args.gather, warps_per_query=args.warps, stage_rows=args.stage_rows, hash_space=args.hash_space)
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
                       "mean_hops": round(mean_hops, 1), "parallelism": f"queries sharded over {world} GPU(s), index replicated",
                       "l2": "256 MiB flush write between timed iterations", "index": d["info"],
                       "knn_bootstrap": "torch fp32 matmul+topk (data prep, untimed)"},
            "e2e": {"value": round(nq * world / (e2e_ms / args.steps * 1e-3), 1), "unit": "queries/s",
                    "h2d_bytes_per_step": nq * dim * 4, "d2h_bytes_per_step": nq * k * 8 + 8,
                    "api": "rg_search_batch (C ABI, pinned host buffers)"},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
                         "frac": round(achieved / peak, 4), "traffic": None, "peak_kind": peak_kind,
                         "kernel": "rg_search_kernel", "algorithmic_bytes_per_launch": int(alg_bytes)},
            "clocks": sampler.summary(),
        }
        if not args.no_cpu_baseline:
            cb = cpu_baseline(args, d, L_sel)
            cb.pop("_res")
            out["cpu_baseline"] = cb
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def cpu_baseline(args, d, L, threads=None, sample=None):
    """The reference's CPU search (oracle/_ref when it is loadable, else the C port) on a bounded query sample."""
    from mysteryann_b200 import io
    from oracle.binding import Oracle, Ref, ref_available

    sample = min(sample or args.cpu_sample, args.queries)
    q = d["queries"][:sample].cpu().numpy()
    base = d["base"].cpu().numpy()
    if ref_available():
        r = Ref()
        threads = threads or r.num_procs()
        with tempfile.TemporaryDirectory() as tmp:
            fb = os.path.join(tmp, "base.fbin")
            io.write_fbin(fb, base)
            h = r.open(fb, d["index_path"], metric=1, threads=threads)
            res = r.search(h, q, args.k, L, threads=threads, warmup=True)
            r.close(h)
        kind = "reference"
    else:
        o = Oracle()
        threads = threads or o.num_procs()
        res = o.search(base, d["offsets"], d["adj"], d["ep"], q, args.k, L, metric=1, threads=threads)
        kind = "port"
    return {"value": round(sample / res["seconds"], 1), "unit": "queries/s", "cores": threads, "kind": kind,
            "sample": f"first {sample} of the {args.queries} queries, L_pq={L}, OpenMP schedule(dynamic,1) like "
                      "tests/test_search_roargraph.cpp:203", "_res": res}


def run_reference(args):
    """--impl reference: the reference's own CPU search loop on the host cores, same config/metric."""
    import torch

    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    if rank != 0:
        return
    device = torch.device("cuda", 0) if torch.cuda.is_available() else torch.device("cpu")
    d = prepare(args, 0, 1, device)
    # same beam width rule as our arm: smallest L reaching the recall target (on the CPU sample)
    L_sel = args.L
    sample = min(args.cpu_sample, args.queries)
    if not L_sel:
        for L in L_SWEEP:
            cb = cpu_baseline(args, d, L, sample=min(sample, 1000))
            if recall_at_k(cb["_res"]["ids"], d["gt"][:len(cb["_res"]["ids"])], args.k) >= args.recall:
                L_sel = L
                break
        L_sel = L_sel or L_SWEEP[-1]
    times = []
    for _ in range(args.warmup + args.steps):
        cb = cpu_baseline(args, d, L_sel, sample=sample)
        times.append(sample / cb["value"])
    t = float(np.mean(times[args.warmup:]))
    value = sample / t
    cb.pop("_res")
    cb["value"] = round(value, 1)
    out = {"impl": "reference", "metric": "QPS at recall@10=0.9 (IP, OOD queries)", "value": round(value, 1),
           "unit": "queries/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": round(t * 1e3, 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": "f32", "data": "synthetic",
           "config": {"workload": f"{args.n}x{args.dim} fp32 IP base, {args.queries} OOD queries, k={args.k}, "
                                  f"L_pq={L_sel}; each step = {sample}-query sample on {cb['cores']} host threads",
                      "n_base": args.n, "dim": args.dim, "k": args.k, "L_pq": L_sel},
           "cpu_baseline": cb,
           "e2e": {"value": round(value, 1), "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    a = parse_args()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
