"""Turns ncu reports (--set full) into the per-kernel roofline table of profiles/README.md.
usage: python profiles/ncu_table.py gpurun_out/r2_spmv_sell.ncu-rep [more.ncu-rep ...] > table.md
Per kernel (template arguments kept): launches captured, mean duration, DRAM read+write bytes per launch, achieved DRAM GB/s,
registers, achieved occupancy, L1/TEX hit rate, FP64 pipe utilisation, issue-slot utilisation."""
import csv
import io
import re
import subprocess
import sys
from collections import defaultdict

WANT = {
    "gpu__time_duration.sum": "dur",
    "dram__bytes_read.sum": "rd",
    "dram__bytes_write.sum": "wr",
    "launch__registers_per_thread": "regs",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "occ",
    "l1tex__t_sector_hit_rate.pct": "l1hit",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active": "fp64",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active": "fp64b",
    "sm__inst_issued.avg.pct_of_peak_sustained_active": "issue",
    "smsp__issue_active.avg.pct": "issue2",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "drampct",
    "l1tex__throughput.avg.pct_of_peak_sustained_active": "l1pct",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed": "l2pct",
}
UNIT = {"ns": 1e-9, "nsecond": 1e-9, "us": 1e-6, "usecond": 1e-6, "ms": 1e-3, "msecond": 1e-3, "s": 1.0, "second": 1.0,
        "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def rows_of(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    lines = [l for l in out.splitlines() if l.startswith('"')]
    rd = csv.reader(io.StringIO("\n".join(lines)))
    header = next(rd)
    units = next(rd)
    idx = {h: i for i, h in enumerate(header)}
    for r in rd:
        name = re.sub(r"\(anonymous namespace\)::", "", r[idx["Kernel Name"]])
        name = re.sub(r"\((SpmvArgs|Pcg2Args|long|int|const|double|CgState|unsigned).*$", "", name)
        rec = {"name": name}
        for m, k in WANT.items():
            if m in idx and r[idx[m]] not in ("", "n/a"):
                v = float(r[idx[m]].replace(",", ""))
                rec[k] = v * UNIT.get(units[idx[m]], 1.0)
        yield rec


agg = defaultdict(list)
for rep in sys.argv[1:]:
    for rec in rows_of(rep):
        agg[rec["name"]].append(rec)
print("| kernel | captured | mean duration | DRAM bytes / launch (read + write) | DRAM GB/s | of 6546 GB/s | regs | occupancy % | L1 hit % | L1/TEX % | L2 % | FP64 pipe % | issue % |")
print("|---|---|---|---|---|---|---|---|---|---|---|---|---|")
mean = lambda xs: sum(xs) / len(xs) if xs else float("nan")
for name, recs in sorted(agg.items(), key=lambda kv: -mean([r.get("dur", 0) for r in kv[1]])):
    dur = mean([r["dur"] for r in recs if "dur" in r])
    byt = mean([r.get("rd", 0) + r.get("wr", 0) for r in recs])
    g = lambda k: mean([r[k] for r in recs if k in r])
    fp64 = g("fp64") if any("fp64" in r for r in recs) else g("fp64b")
    issue = g("issue") if any("issue" in r for r in recs) else g("issue2")
    print(f"| `{name}` | {len(recs)} | {dur * 1e6:.1f} us | {byt / 1e6:.1f} MB | {byt / dur / 1e9:.0f} | {byt / dur / 1e9 / 6546.2:.2f} | {g('regs'):.0f} | {g('occ'):.0f} | {g('l1hit'):.0f} | {g('l1pct'):.0f} | {g('l2pct'):.0f} | {fp64:.0f} | {issue:.0f} |")
