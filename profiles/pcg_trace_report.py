"""Reads the per-warp trace of the persistent CG kernel (PE_PCG_TRACE=<prefix>, kernels_pcg2.cuh) and prints, for the first inner
pass and the CG pass of the last iteration of the last displacement solve: when the warps finished their own stream relative to
the first warp's start, the wait each warp then spent at the grid barrier, and the slices per warp.
usage: python profiles/pcg_trace_report.py <prefix>_rank0.bin"""
import sys

import numpy as np

t = np.fromfile(sys.argv[1], dtype=np.uint64).reshape(-1, 8)
for name, o in (("first inner pass", 0), ("CG pass", 4)):
    start, end, released, meta = (t[:, o + k].astype(np.int64) for k in range(4))
    live = start > 0
    if not live.any():
        print(name, ": not recorded")
        continue
    start, end, released, meta = start[live], end[live], released[live], meta[live]
    smid, slices = (meta >> 32).astype(int), (meta & 0xffffffff).astype(int)
    t0 = start.min()
    e = (end - t0) / 1e3
    q = lambda a, p: float(np.percentile(a, p))
    print(f"{name}: {live.sum()} warps on {len(set(smid))} SMs, {slices.sum()} slices ({slices.min()}..{slices.max()} per warp, mean {slices.mean():.2f})")
    print(f"  start spread            : {q((start - t0) / 1e3, 50):7.2f} us median, {q((start - t0) / 1e3, 100):7.2f} max")
    print(f"  own stream finished at  : min {e.min():7.2f}  p10 {q(e, 10):7.2f}  median {q(e, 50):7.2f}  p90 {q(e, 90):7.2f}  p99 {q(e, 99):7.2f}  max {e.max():7.2f} us")
    print(f"  barrier released at     : median {q((released - t0) / 1e3, 50):7.2f} us  (last warp -> release: {q((released - t0) / 1e3, 50) - e.max():6.2f} us)")
    w = (released - end) / 1e3
    print(f"  wait at the barrier     : mean {w.mean():6.2f}  median {q(w, 50):6.2f}  max {w.max():6.2f} us  = {100 * w.mean() / q((released - t0) / 1e3, 50):.1f} % of the pass")
    per_sm = {}
    for s, x in zip(smid, e):
        per_sm.setdefault(s, []).append(x)
    last = sorted(((max(v), s) for s, v in per_sm.items()), reverse=True)[:5]
    print("  SMs that finished last  : " + ", ".join(f"SM {s} at {x:.1f} us" for x, s in last))
    late = e > q(e, 99)
    print(f"  slices of the latest 1% : {slices[late].tolist()[:16]}")
