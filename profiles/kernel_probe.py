"""One call of every cell-loop / vector kernel family of a time step outside the matrix stream, on the 3D Q1/Q1 config at the
given refinement — the command profiled with ncu for the per-kernel table of profiles/README.md (row g1).  Run with
PE_PCG2=0 PE_PCG=0 so that the CG vector kernels are separate launches.
usage: python profiles/kernel_probe.py <refine>"""
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import helpers as H  # noqa: E402

refine = int(sys.argv[1]) if len(sys.argv) > 1 else 6
capi, fss = H.capi, H.fss
t0 = time.time()
inp = capi.InputData(text=H.make_input(dim=3, refine=refine, degree_u=1))
mesh = fss.make_mesh(inp)
prm = inp.params()
prm.cg_max_iterations = 3
dev = capi.create_device_backend(0)
fss.upload_problem(dev, inp, mesh, prm)
dev.pressure_set_uniform(inp.p_init)
dev.displacement_assemble()          # k_elasticity (8 colours), k_u_rhs (8 colours)
try:
    dev.displacement_solve()         # 3 iterations: k_cheb_first, k_cg_update (+ the matrix passes)
except capi.BackendError as e:
    assert e.status == capi.PE_ERR_NO_CONVERGENCE
dev.project_assemble_matrix()
dev.project_assemble_rhs(fss.VOLUMETRIC_COMPONENTS[3])   # k_projection_rhs (8 colours)
dev.pressure_begin_step()
dev.pressure_zero_update()
dev.update_volumetric_strain()
r = dev.assemble_residual(inp.time_step)                 # k_residual_t1, k_pressure_residual
st = dev.stats()
print({"refine": refine, "n_cells": st["n_cells"], "residual": r, "launches": st["kernel_launches"], "wall_s": round(time.time() - t0, 2)})
dev.close()
