"""profiles/k1_traffic.json from an `ncu --set full` capture of ONE K1 launch of the bench workload:

    ncu --set full --clock-control none --import-source on -k regex:rg_search_kernel -s 6 -c 1 -o gpurun_out/k1 -f \\
        python bench.py --L <L> --steps 3 --warmup 3 --no-cpu-baseline --knn-slice 0
    python tools/make_k1_traffic.py gpurun_out/k1.ncu-rep gpurun_out/bench_line.json [summary.txt]

bench_line.json is the JSON line of a plain bench run of the same command (workload, algorithmic bytes).  The file records
the fingerprint of the K1 sources it was captured from; bench.py only quotes it while that fingerprint matches."""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    rep, line = sys.argv[1], sys.argv[2]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]
    m = {h: (u, v) for h, u, v in zip(hdr, units, vals)}

    def to_bytes(key):
        u, v = m[key]
        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[u]
        return int(float(v.replace(",", "")) * scale)

    rd, wr = to_bytes("dram__bytes_read.sum"), to_bytes("dram__bytes_write.sum")
    b = json.loads(open(line).read().strip().splitlines()[-1])
    cfg = b["config"]
    alg = b["roofline"]["algorithmic_bytes_per_launch"]
    doc = {"source": f"{os.path.basename(rep)} (ncu --set full --clock-control none, one rg_search_kernel launch of bench.py --L {cfg['L_pq']})",
           "kernel": m["Kernel Name"][1], "k1_source_hash": bench.k1_source_hash(),
           "workload": {"n_base": cfg["n_base"], "dim": cfg["dim"], "queries": cfg["queries_per_gpu"], "L_pq": cfg["L_pq"],
                        "k": cfg["k"], "n_train": cfg["index"]["n_train"]},
           "dram_bytes_read": rd, "dram_bytes_write": wr, "dram_bytes_per_launch": rd + wr,
           "algorithmic_bytes_per_launch": alg, "ratio_traffic_over_algorithmic": round((rd + wr) / alg, 4),
           "gpu_time_ms_under_ncu": float(m["gpu__time_duration.sum"][1].replace(",", "")) * {"ms": 1, "us": 1e-3, "s": 1e3, "usecond": 1e-3, "msecond": 1, "second": 1e3}.get(m["gpu__time_duration.sum"][0], 1)}
    json.dump(doc, open(os.path.join(ROOT, "profiles", "k1_traffic.json"), "w"), indent=1)
    print(json.dumps(doc))


if __name__ == "__main__":
    main()
