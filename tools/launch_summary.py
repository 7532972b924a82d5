"""Condenses an ncu launch list (--metrics gpu__time_duration.sum --csv) into per-kernel count / total / share.
usage: python tools/launch_summary.py launches.csv [first_id]"""
import csv
import re
import sys
from collections import OrderedDict

rows = [r for r in csv.reader(open(sys.argv[1], errors="replace")) if len(r) > 14 and r[0].isdigit()]
first = int(sys.argv[2]) if len(sys.argv) > 2 else 0
agg = OrderedDict()
for r in rows:
    if int(r[0]) < first:
        continue
    name = re.sub(r"\(.*", "", r[4])[:90]
    v = float(r[14].replace(",", ""))
    unit = r[13]
    us = v / 1e3 if unit in ("ns", "nsecond") else v * 1e3 if unit in ("ms", "msecond") else v
    c = agg.setdefault(name, [0, 0.0])
    c[0] += 1
    c[1] += us
tot = sum(c[1] for c in agg.values())
print(f"# {sys.argv[1]}: launches with ID >= {first}; total {tot / 1e3:.3f} ms")
for name, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{us / tot * 100:6.2f}%  {us / 1e3:10.3f} ms  x{n:<4d} {name}")
