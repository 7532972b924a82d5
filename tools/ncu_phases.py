"""Share of executed warp instructions and of stall samples per phase of K1, from an .ncu-rep with source info.
The phases are found by marker comments / lambda names in the CURRENT rg_search.cu, so run it on a capture of the same source.
usage: python tools/ncu_phases.py gpurun_out/k1.ncu-rep"""
import collections
import csv
import os
import subprocess
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
src = open(os.path.join(ROOT, "mysteryann_b200/csrc/rg_search.cu")).read().splitlines()
MARKS = [("visited_test_and_set32(uint32_t *table", "cas hash"), ("ld_cg_v4(const void *ptr)", "bucket helpers"),
         ("uint32_t bucket_test_and_set(", "bucket test_and_set"), ("uint32_t lower_bound_key(", "lower_bound (merge)"),
         ("__global__ void __launch_bounds__", "prologue"), ("auto compact_mine", "compact_mine"), ("auto spec_block", "spec_block"),
         ("auto gather_and_score", "gather issue"), ("for (uint32_t r0 = 0; r0 < rows; r0 += 8)", "score loop + candidate append"),
         ("// ---- next query", "query setup"), ("__syncthreads();  // the hop's candidates", "hop top (barrier wait)"),
         ("// (a) position of every", "merge"), ("first unexpanded entry at or after", "cursor scan"),
         ("// ---- expand P[cur]", "expand setup + adjacency"), ("// speculation for the NEXT hop", "spec scan + adjacency read-ahead"),
         ("uint32_t n_w = 0;", "filter rounds"), ("// a re-scored entry point lands", "batch loop / hop tail"),
         ("if (overflow) {", "results")]
bounds = []
for text, name in MARKS:
    for i, l in enumerate(src):
        if text in l:
            bounds.append((i + 1, name))
            break
bounds.sort()


def phase(f, line):
    if f == "rg_distance.cuh":
        return "distance"
    if f == "rg_common.cuh":
        return "mbarrier wait" if 100 <= line <= 112 else "rg_common (keys, TMA issue)"
    if f != "rg_search.cu":
        return "intrinsics (" + f + ")"
    name = "head"
    for b, n in bounds:
        if line >= b:
            name = n
    return name


out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--print-source", "cuda,sass", "--csv"],
                     capture_output=True, text=True).stdout
agg = collections.defaultdict(lambda: [0, 0])
fname = ""
for r in csv.reader(out.splitlines()):
    if len(r) == 2 and r[0] == "File Path":
        fname = r[1].split("/")[-1]
    elif len(r) > 8 and r[0].isdigit() and r[7].isdigit():
        a = agg[phase(fname, int(r[0]))]
        a[0] += int(r[6]) if r[6].isdigit() else 0
        a[1] += int(r[7])
ts = sum(a[0] for a in agg.values()) or 1
ti = sum(a[1] for a in agg.values()) or 1
print(f"total warp instructions {ti}, samples {ts}")
for k, (s_, i_) in sorted(agg.items(), key=lambda x: -x[1][0]):
    print(f"{k:40s} {s_ / ts * 100:5.1f}% of stall samples {i_ / ti * 100:5.1f}% of instructions")
