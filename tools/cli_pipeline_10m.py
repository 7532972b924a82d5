"""The reference's whole workflow at the BASELINE.json C2 shape, run through the three drop-in CLI drivers exactly as a
user of the reference would run them (run_roargraph_test.sh / run_roargraph_search_test.sh), on files:

    compute_groundtruth (train -> base, K=100)   -> train.gt.bin      [learn_base_nn_path]
    compute_groundtruth (test  -> base, K=100)   -> test.gt.bin       [gt_path]
    test_build_roargraph --gpu_build 1           -> rg.index
    test_search_roargraph, the reference's L_pq sweep cut at 500 -> eval.csv (L_pq, qps, avg_cmps, latency, recall, hops)

Synthetic data as in bench.py (same generator and seed).  Prints per-stage wall-clock seconds and the evaluation table,
plus the gathered-row bandwidth each row implies (avg_cmps x dim x 4 B x QPS).

    python tools/cli_pipeline_10m.py --n 10000000 --train 2000000 --out gpurun_out/cli_pipeline_10m.txt
"""
import argparse
import json
import os
import subprocess
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from mysteryann_b200 import build, hostlib, io, synth  # noqa: E402

L_SWEEP = [10, 15, 20, 25, 30, 35, 40, 45, 50, 55, 60, 65, 70, 75, 80, 85, 90, 95, 100, 110, 120, 130, 140, 150, 160, 170,
           180, 190, 200, 220, 240, 260, 280, 300, 350, 400, 450, 500]  # run_roargraph_search_test.sh:13, cut at 500


def run(cmd, log):
    t0 = time.time()
    p = subprocess.run(cmd, capture_output=True, text=True)
    dt = time.time() - t0
    log.write(f"$ {' '.join(cmd)}\n{p.stdout[-4000:]}\n{p.stderr[-2000:]}\n")
    if p.returncode != 0:
        raise SystemExit(f"FAILED ({p.returncode}): {' '.join(cmd)}\n{p.stdout[-2000:]}\n{p.stderr[-2000:]}")
    return dt


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=10_000_000)
    ap.add_argument("--train", type=int, default=2_000_000)
    ap.add_argument("--queries", type=int, default=10_000)
    ap.add_argument("--dim", type=int, default=200)
    ap.add_argument("--dir", default="/tmp/rg_cli_pipeline")
    ap.add_argument("--out", default="")
    a = ap.parse_args()
    build.build()
    hostlib.build()
    os.makedirs(a.dir, exist_ok=True)
    f = lambda name: os.path.join(a.dir, name)
    log = open(f("commands.log"), "w")
    sec = {}

    t0 = time.time()
    base, train, test = synth.make_torch(a.n, a.train, a.queries, a.dim, device="cuda")
    for name, x in (("base.fbin", base), ("train.fbin", train), ("test.fbin", test)):
        io.write_fbin(f(name), x.cpu().numpy())
    del base, train, test
    import torch

    torch.cuda.empty_cache()
    sec["write_inputs"] = round(time.time() - t0, 1)

    gt_tool, b = os.path.join(hostlib.BIN_DIR, "compute_groundtruth"), hostlib.BIN_DIR
    sec["compute_groundtruth_train"] = round(run([gt_tool, "--data_type", "float", "--dist_fn", "mips", "--base_file", f("base.fbin"),
                                                  "--query_file", f("train.fbin"), "--gt_file", f("train.gt.bin"), "--K", "100"], log), 1)
    sec["compute_groundtruth_test"] = round(run([gt_tool, "--data_type", "float", "--dist_fn", "mips", "--base_file", f("base.fbin"),
                                                 "--query_file", f("test.fbin"), "--gt_file", f("test.gt.bin"), "--K", "100"], log), 1)
    sec["test_build_roargraph"] = round(run([os.path.join(b, "test_build_roargraph"), "--data_type", "float", "--dist", "ip",
                                             "--base_data_path", f("base.fbin"), "--sampled_query_data_path", f("train.fbin"),
                                             "--projection_index_save_path", f("rg.index"), "--learn_base_nn_path", f("train.gt.bin"),
                                             "--M_sq", "100", "--M_pjbp", "35", "--L_pjpq", "500", "--gpu_build", "1"], log), 1)
    sec["test_search_roargraph"] = round(run([os.path.join(b, "test_search_roargraph"), "--data_type", "float", "--dist", "ip",
                                              "--base_data_path", f("base.fbin"), "--query_path", f("test.fbin"), "--gt_path",
                                              f("test.gt.bin"), "--projection_index_save_path", f("rg.index"), "--k", "10",
                                              "--evaluation_save_path", f("eval.csv"), "--L_pq"] + [str(L) for L in L_SWEEP], log), 1)
    log.close()
    peak = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
    lines = [f"# drop-in CLI pipeline, {a.n} x {a.dim} fp32 IP base, {a.train} training queries, {a.queries} OOD test queries, k=10",
             f"# stage seconds (process wall clock, file I/O and uploads included): {json.dumps(sec)}",
             "# eval.csv of test_search_roargraph (one rg_search_batch call per L_pq on page-locked host arrays; QPS = queries / wall clock)",
             "L_pq,qps,avg_cmps,mean_latency_ms,recall@10,avg_hops,gathered_GB/s,frac_of_measured_hbm_peak"]
    for row in open(f("eval.csv")):
        r = row.strip().split(",")
        if len(r) == 6:
            gbs = float(r[1]) * float(r[2]) * a.dim * 4 / 1e9
            lines.append(",".join(r) + f",{gbs:.0f},{gbs / peak:.3f}")
    text = "\n".join(lines) + "\n"
    print(text)
    if a.out:
        open(a.out, "w").write(text)


if __name__ == "__main__":
    main()
