"""Summarise an .ncu-rep per CUDA source line: share of executed warp instructions and of stall samples.
usage: python tools/ncu_lines.py gpurun_out/k1_prof.ncu-rep [top_n]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
items, fname = [], ""
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        fname = r[1].split("/")[-1]
    elif len(r) > 8 and r[0].isdigit() and r[7].isdigit():
        stalls = {}
        items.append((int(r[7]), int(r[6]) if r[6].isdigit() else 0, fname, int(r[0]), r[1].strip()[:100]))
tot_i = sum(i[0] for i in items) or 1
tot_s = sum(i[1] for i in items) or 1
print(f"total warp instructions {tot_i}, samples {tot_s}")
items.sort(reverse=True)
for inst, smp, f, line, src in items[:top]:
    print(f"{inst / tot_i * 100:5.1f}% inst {smp / tot_s * 100:5.1f}% smp  {f}:{line:<4} {src}")
