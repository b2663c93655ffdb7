"""Timing probe of the hanging-node path on one GPU (not a parity test — tests/test_zz_gpu_amr.py is):

    python profiles/amr_probe.py [base_refine=5] [rounds=2] [degree_u=1]

Builds a 3D box of 2^base cells per axis, refines `rounds` times around the well axis (r < 2.5, then r < 1.5, ...), uploads
it with hanging-node constraints and reports pe_setup time (pattern from cell lists, M/K assembly + condense), the first
displacement_assemble (elasticity assembly + condense + block-CSR copy) and three time steps; prints one JSON line."""
import json
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "tests"))
import helpers as H  # noqa: E402

capi, fss = H.capi, H.fss


def main():
    base = int(sys.argv[1]) if len(sys.argv) > 1 else 5
    rounds = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    deg = int(sys.argv[3]) if len(sys.argv) > 3 else 1
    inp = capi.InputData(text=H.make_input(dim=3, refine=base, degree_u=deg, extra_gpu="  set Preconditioner = 0\n  set CG max iterations = 20000\n"))
    t0 = time.time()
    mesh0 = capi.mesh_rectangle(3, [10.0] * 3, base)
    F = capi.Forest(mesh0, base)
    for r in range(rounds):
        m = F.active_mesh().arrays
        ctr = m.xyz[m.cell_vertices].mean(axis=1)
        F.set_flags(refine=(np.hypot(ctr[:, 0], ctr[:, 1]) < 2.5 / (r + 1)).astype(np.int8))
        F.execute()
    am = F.active_mesh()
    t_host_mesh = time.time() - t0
    dev = capi.create_device_backend(0)
    t0 = time.time()
    dp, du, (Lp, Lu) = fss.upload_problem(dev, inp, am, forest=F)
    t_upload_setup = time.time() - t0
    st = dev.stats()
    t0 = time.time()
    dev.pressure_set_uniform(inp.p_init)
    dev.displacement_assemble()
    dev.lib.pe_synchronize(dev.ctx)
    t_first_assemble = time.time() - t0
    fss.initialize(dev, inp)
    steps = []
    for _ in range(3):
        t0 = time.time()
        rep = fss.time_step(dev, inp)
        dev.lib.pe_synchronize(dev.ctx)
        steps.append({"s": time.time() - t0, "cg_u": rep["cg_its_displacement"], "cg_p": rep["cg_its_pressure"], "inner": rep["inner_counts"]})
    out = {"cells": am.arrays.n_cells, "levels": np.bincount(F.levels()).tolist(), "n_dofs_p": dp.n_dofs, "n_dofs_u": du.n_dofs,
           "hanging_p": Lp.n_lines, "hanging_u_lines": int((np.diff(Lu.entry_ptr) > 0).sum()), "nnz_u": st["nnz_u"], "nnz_p": st["nnz_p"],
           "host_forest_s": t_host_mesh, "upload_plus_pe_setup_s": t_upload_setup, "pe_setup_ms": st["setup_ms"],
           "first_displacement_assemble_s": t_first_assemble, "bsr_block_size": dev.stats()["bsr_block_size"], "steps": steps}
    print(json.dumps(out))
    dev.close()


if __name__ == "__main__":
    main()
