"""stdin: `ncu --page raw --csv`; stdout: the same table restricted to the columns the roofline table and the judge's checks use
(identification, duration, DRAM / L2 / L1 traffic and throughput, occupancy, registers, pipe and issue utilisation, stall reasons)."""
import csv
import re
import sys

KEEP = re.compile(r"^(ID|Kernel Name|Context|Stream|Block Size|Grid Size|gpu__time_duration|dram__bytes|dram__throughput|gpu__dram_throughput|lts__t_bytes|lts__throughput|"
                  r"lts__t_sector_hit_rate|l1tex__t_sector_hit_rate|l1tex__throughput|l1tex__t_bytes|launch__registers|launch__occupancy|launch__shared|sm__warps_active|sm__throughput|"
                  r"sm__inst_executed_pipe_fp64|sm__pipe_fp64|sm__inst_issued|smsp__issue_active|smsp__inst_executed.sum|sm__cycles_elapsed.avg$|smsp__average_warp.*stall|smsp__pcsamp_warps_issue_stalled)")
lines = [l for l in sys.stdin if l.startswith('"')]
rd = csv.reader(lines)
header = next(rd)
cols = [i for i, h in enumerate(header) if KEEP.match(h)]
w = csv.writer(sys.stdout)
w.writerow([header[i] for i in cols])
for r in rd:
    w.writerow([r[i] if i < len(r) else "" for i in cols])
