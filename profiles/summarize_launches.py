"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel (template arguments kept) the number of
launches, the summed duration and its share.  usage: python profiles/summarize_launches.py launches.csv > summary.txt"""
import csv
import re
import sys
from collections import defaultdict

rows = []
with open(sys.argv[1], newline="") as f:
    lines = [l for l in f if l.startswith('"')]
for r in csv.DictReader(lines):
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = r["Kernel Name"]
    name = re.sub(r"\(anonymous namespace\)::", "", name)
    name = re.sub(r"\(.*$", "", name)
    unit = r.get("Metric Unit", "ns")
    val = float(r["Metric Value"].replace(",", ""))
    scale = {"ns": 1e-6, "nsecond": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0, "s": 1e3, "second": 1e3}.get(unit, 1e-6)
    rows.append((name, val * scale))
tot = sum(v for _, v in rows)
agg = defaultdict(lambda: [0, 0.0])
for n, v in rows:
    agg[n][0] += 1
    agg[n][1] += v
print(f"{len(rows)} launches, {tot:.3f} ms in kernels (cold-cache, serialised: compare SHARES)")
print(f"{'kernel':90s} {'launches':>8s} {'ms':>10s} {'share':>7s} {'avg us':>9s}")
for n, (k, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{n[:90]:90s} {k:8d} {v:10.3f} {100 * v / tot:6.2f}% {1e3 * v / k:9.2f}")
