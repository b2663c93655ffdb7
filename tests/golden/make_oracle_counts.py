"""Runs the CPU oracle (SSOR-CG, the reference's solver settings) on the 3D Q1/Q1 benchmark configs and
records iteration counts, phase timings and field checksums to tests/golden/oracle_counts_r<refine>.json.
These records are what bench.py's bounded CPU sample extrapolates with (it cannot afford a full oracle
step at 128^3 inside a default run).  Usage:  python tests/golden/make_oracle_counts.py <refine> [steps] [max_its]"""
import json
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import helpers as H  # noqa: E402

refine = int(sys.argv[1])
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
max_its = int(sys.argv[3]) if len(sys.argv) > 3 else 1000
threads = H.load_oracle().po_set_threads(0)
inp = H.capi.InputData(text=H.make_input(dim=3, refine=refine, degree_u=1))
mesh = H.fss.make_mesh(inp)
prm = inp.params()
prm.cg_max_iterations = max_its
b = H.create_oracle_backend()
t0 = time.time()
H.fss.upload_problem(b, inp, mesh, prm)
t_setup = time.time() - t0
t0 = time.time()
init = H.fss.initialize(b, inp)
t_init = time.time() - t0
out = {"refine": refine, "threads": threads, "cg_max_iterations": max_its, "setup_s": t_setup, "init_s": t_init, "init": init,
       "stats": b.stats(), "steps": []}
print(out, flush=True)
for s in range(steps):
    t0 = time.time()
    rep = H.fss.time_step(b, inp)
    rep["wall_s"] = time.time() - t0
    p, u = b.get_vector(H.capi.VEC_P), b.get_vector(H.capi.VEC_U)
    rep["p_l2"], rep["u_l2"], rep["p_sum"] = float(np.linalg.norm(p)), float(np.linalg.norm(u)), float(p.sum())
    out["steps"].append(rep)
    print(rep, flush=True)
    Path(__file__).with_name(f"oracle_counts_r{refine}.json").write_text(json.dumps(out, indent=1))
