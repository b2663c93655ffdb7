"""Golden vectors produced by THE REFERENCE'S OWN CODE: runs oracle/_ref/fss_ref — the unmodified sources of
/root/reference/lib/include compiled against the deal.II API shim of oracle/dealii_shim (NOT deal.II; oracle/Makefile, target
`ref`) with oracle/ref_main.cpp as the missing Runner.cpp — on a handful of parameter files, and records what
`PoroElasticProblem<dim>::run()` prints and writes:

  reference_run_<case>.json   the parameter file, the loop's own prints per time step ("pressure converged; iterations", "Solution
                              limits", "Error", FSS:368-405) and the iteration count + final residual of every CG solve in call
                              order (the shim's solver log; the reference's own prints of them are commented out)
  reference_run_<case>.npz    every vector handed to DataOut at FSS:411 (p, u, projected strains, stresses) after every time
                              step, dof by dof, with the support point of each dof

Every run ends before the reference's first mesh refinement (time step 5, FSS:333), which the shim does not provide.
/root/reference exists only in the build container, so the records are committed; tests/test_reference_run.py compares the CPU
oracle with them and tests/test_zzz_gpu_reference_run.py the CUDA path.   usage: python tests/golden/make_reference_run.py"""
import json
import re
import subprocess
import sys
import tempfile
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
ROOT = HERE.parents[1]

SHIPPED = (Path("/root/reference/input.data").read_text() if Path("/root/reference/input.data").exists() else "")

PROPERTIES = """
subsection Properties
  set Young modulus         = {E}
  set Biot coefficient      = {biot}
  set Bulk density          = 2700
  set Fluid compressibility = 5.8e-10
  set Permeability          = {perm}
  set Poisson ratio         = {nu}
  set Porosity              = 0.3
  set Viscosity             = 1e-3
  set Well radius           = {rw}
  set Flow rate             = {q}
end
"""


def text(dim, size, refine, dirichlet, neumann=("", "", ""), dt=60, steps=3, E="1.4e10", biot="0.9", perm="10", nu="0.3", rw="1", q="1e-5", solver=""):
    return f"""
subsection Mesh
  set Dimensions               = {dim}
  set Domain size              = {size}
  set Initial refinement level = {refine}
  set Max refinement level     = 6
end
subsection In situ
  set Displacement boundary labels     = {dirichlet[0]}
  set Displacement boundary components = {dirichlet[1]}
  set Displacement boundary values     = {dirichlet[2]}
  set Initial pressure                 = 10e6
  set Stress boundary labels           = {neumann[0]}
  set Stress boundary components       = {neumann[1]}
  set Stress boundary values           = {neumann[2]}
end
{PROPERTIES.format(E=E, biot=biot, perm=perm, nu=nu, rw=rw, q=q)}
subsection Solver
  set Time step  = {dt}
  set Time max   = {dt * steps}
{solver}end
"""


CASES = {
    # the reference's input.data exactly as shipped, cut to the four time steps before its first refinement
    "shipped_4steps": lambda: SHIPPED.replace("set Time max   = 1e3", "set Time max   = 240"),
    # 3D (the reference's well source is written for 2D but compiles for 3D in release mode; Q2 displacement, 512 cells)
    "box3d_r3": lambda: text(3, "10, 10, 10", 3, ("0, 1, 2, 3, 4, 5", "0, 0, 1, 1, 2, 2", "0, -1e-5, 0, -1e-5, 0, -1e-5"), steps=3),
    # BASELINE configs[1] in small: traction on the top face (DS:249-277), rollers elsewhere
    "neumann2d_r4": lambda: text(2, "10, 10", 4, ("0, 1, 2", "0, 0, 1", "0, 0, 0"), ("3", "1", "-1e6"), steps=3),
    # anisotropic cells, other material, a condition list in which two conditions claim the corner dofs (first one wins)
    "rect2d_r3": lambda: text(2, "10, 6", 3, ("3, 2, 0, 1, 0", "1, 1, 0, 0, 1", "-2e-5, 0, 0, 1e-5, 0"), dt=30, steps=4, E="2.1e10", biot="0.8",
                              perm="25", nu="0.25", rw="1.5", q="2e-5"),
    # traction on the top face of a 3D box (3D face quadrature and normals in DS:249-277), rollers on the other five faces
    "neumann3d_r2": lambda: text(3, "10, 8, 6", 2, ("0, 1, 2, 3, 4", "0, 0, 1, 1, 2", "0, 0, 0, 0, 0"), ("5", "2", "-2e6"), steps=2),
    # the shipped case two levels finer (64^2 cells, 33,282 displacement dofs): the same agreement at a size where SSOR-CG needs
    # several hundred iterations
    "shipped_r6": lambda: SHIPPED.replace("set Time max   = 1e3", "set Time max   = 180").replace("set Initial refinement level = 4", "set Initial refinement level = 6"),
    # BASELINE.json's benchmark configurations are Q1/Q1; the reference hard-codes FE_Q(2) at DS:67.  The q1_ cases run the same
    # unmodified code with the shim's run-time override DEALII_SHIM_FESYSTEM_DEGREE=1 (configs[2]/[3] in small: 16^3 cells; configs[1]
    # in small: 32^2 cells with the top traction)
    "q1_box3d_r4": lambda: text(3, "10, 10, 10", 4, ("0, 1, 2, 3, 4, 5", "0, 0, 1, 1, 2, 2", "0, -1e-5, 0, -1e-5, 0, -1e-5"), steps=4),
    "q1_neumann2d_r5": lambda: text(2, "10, 10", 5, ("0, 1, 2", "0, 0, 1", "0, 0, 0"), ("3", "1", "-1e6"), steps=4),
    # the loop's other exits (FSS:349-381): an FSS tolerance that is never met, so the coupling loop runs to its cap, and a pressure
    # loop that hits its cap of 2 passes before the residual is below tolerance
    "caps2d_r3": lambda: text(2, "10, 10", 3, ("0, 1, 2, 3", "0, 0, 1, 1", "0, -1e-5, 0, -1e-5"), steps=2,
                              solver="  set FSS tolerance = 1e-20\n  set Max FSS iterations = 3\n  set Max pressure iterations = 2\n"),
}


def parse_log(stdout):
    """run()'s prints per time step (FSS:328-406)"""
    steps, cur = [], None
    for line in stdout.splitlines():
        line = line.strip()
        if line.startswith("Time: "):
            cur = {"time": float(line.split()[1]), "coupling_iterations": 0, "pressure_converged_iterations": [], "solution_limits": [], "error": []}
            steps.append(cur)
        elif line.startswith("Coupling iteration:"):
            cur["coupling_iterations"] += 1
        elif line.startswith("pressure converged; iterations:"):
            cur["pressure_converged_iterations"].append(int(line.split()[-1]))
        elif line.startswith("Solution limits:"):
            cur["solution_limits"].append(float(line.split()[2]))
        elif line.startswith("Error:"):
            cur["error"].append(float(line.split()[1]))
    return steps


def parse_dump(path, dim):
    fields, cur = {}, None
    for line in Path(path).read_text().splitlines():
        if line.startswith("#"):
            continue
        if line.startswith("field "):
            _, name, _, ncomp, _, ndofs = line.split()
            while name in fields:   # FSS:262 adds stresses[0] twice (as sigma_xx and as sigma_yy); names are kept as printed
                name += "'"
            cur = fields[name] = {"n_comp": int(ncomp), "rows": []}
            continue
        cur["rows"].append([float(x) for x in line.split()])
    out = {}
    for name, f in fields.items():
        a = np.array(f["rows"])
        out[name] = {"x": a[:, :dim], "comp": a[:, dim].astype(np.int32), "v": a[:, dim + 1]}
    return out


# Full-size cases: too big to commit dof by dof.  They are run by hand (minutes to an hour) and reduced to what the oracle's own
# full-size record holds (tests/golden/oracle_counts_r<refine>.json, oracle_fields_r<refine>.npz): norms and the values at the same
# 4096 sample dofs — the numbering is the same on both sides (tests/test_reference_run.py checks that on the small cases).
BIG_CASES = {
    # BASELINE.json configs[2] (C3): 3D, 64^3 cells, Q1/Q1 — 823,875 displacement + 274,625 pressure dofs
    "q1_c3_r6": (lambda: text(3, "10, 10, 10", 6, ("0, 1, 2, 3, 4, 5", "0, 0, 1, 1, 2, 2", "0, -1e-5, 0, -1e-5, 0, -1e-5"), steps=4), "r6"),
    # BASELINE.json configs[1] (C2): 2D consolidation, 512^2 cells, Q1/Q1, traction on the top face — 526,338 + 263,169 dofs.  The
    # reference cannot run it: its first displacement solve hits SolverControl(1000, 1e-12) (DS:298-299) and run() throws
    # NoConvergence after 1000 iterations at a residual of 0.052 (recorded by hand in reference_run_q1_c2_r9_noconvergence.json;
    # the oracle needs 1779 iterations for that solve).  Kept here as the parameter file of that record.
    "q1_c2_r9_noconvergence": (lambda: text(2, "10, 10", 9, ("0, 1, 2", "0, 0, 1", "0.0, 0.0, 0.0"), ("3", "1", "-1000000.0"), steps=3), "c2_r9"),
    # the same workload one level coarser (256^2 cells): the largest the reference's CG cap lets it run (894 iterations)
    "q1_c2_r8": (lambda: text(2, "10, 10", 8, ("0, 1, 2", "0, 0, 1", "0.0, 0.0, 0.0"), ("3", "1", "-1000000.0"), steps=4), "c2_r8"),
    # BASELINE.json configs[3] (C4, the headline configuration): 128^3 cells, Q1/Q1 — 6,440,067 + 2,146,689 dofs; hours on one core
    "q1_c4_r7_2steps": (lambda: text(3, "10, 10, 10", 7, ("0, 1, 2, 3, 4, 5", "0, 0, 1, 1, 2, 2", "0, -1e-5, 0, -1e-5, 0, -1e-5"), steps=2), "r7"),
}


def reduce_big(name, workdir):
    """workdir: where `DEALII_SHIM_FESYSTEM_DEGREE=1 DEALII_SHIM_SOLVER_LOG=solver.log fss_ref input.data > run.log` has finished"""
    inp_fn, tag = BIG_CASES[name]
    workdir = Path(workdir)
    inp = (workdir / "input.data").read_text()
    assert inp == inp_fn(), "the run was taken with another parameter file"
    dim = int(re.search(r"set Dimensions\s*=\s*(\d)", inp).group(1))
    sample = np.load(HERE / f"oracle_fields_{tag}.npz")
    stdout = (workdir / "run.log").read_text()
    cg = [{"n": int(m.group(1)), "its": int(m.group(2)), "res": float(m.group(3))}
          for m in re.finditer(r"cg n=(\d+) its=(\d+) res=(\S+)", (workdir / "solver.log").read_text())]
    steps = parse_log(stdout)
    for k in range(1, len(steps) + 1):
        d = parse_dump(workdir / "solution" / f"solution-{k:04d}.vtk", dim)
        p, u = d["p"]["v"], d["u"]["v"]
        steps[k - 1].update({"p_l2": float(np.linalg.norm(p)), "p_sum": float(p.sum()), "u_l2": float(np.linalg.norm(u)),
                      "p_samples": p[sample["p_dof"]].tolist(), "u_samples": np.stack([u[sample["u_dof"] + a] for a in range(dim)], axis=1).tolist()})
    rec = {"case": name, "dim": dim, "degree_u": 1, "input": inp, "n_steps": len(steps), "cg_solves": cg, "steps": steps,
           "time_loop_stdout": stdout[stdout.index("starting time loop"):], "samples_of": f"oracle_fields_{tag}.npz (p_dof, u_dof)",
           "produced_by": "oracle/_ref/fss_ref (reference sources, unmodified) with DEALII_SHIM_FESYSTEM_DEGREE=1; reduced by make_reference_run.py reduce_big"}
    (HERE / f"reference_run_{name}.json").write_text(json.dumps(rec))
    return rec


def run_case(name, exe):
    inp = CASES[name]()
    dim = int(re.search(r"set Dimensions\s*=\s*(\d)", inp).group(1))
    with tempfile.TemporaryDirectory() as tmp:
        tmp = Path(tmp)
        (tmp / "solution").mkdir()
        (tmp / "input.data").write_text(inp)
        env = {"DEALII_SHIM_SOLVER_LOG": str(tmp / "solver.log"), "PATH": "/usr/bin:/bin"}
        if name.startswith("q1_"):
            env["DEALII_SHIM_FESYSTEM_DEGREE"] = "1"
        res = subprocess.run([str(exe), "input.data"], cwd=tmp, env=env, capture_output=True, text=True, timeout=3600)
        if res.returncode != 0:
            raise RuntimeError(f"{name}: fss_ref failed: {res.stderr[-2000:]}")
        steps = parse_log(res.stdout)
        cg = [{"n": int(m.group(1)), "its": int(m.group(2)), "res": float(m.group(3))}
              for m in re.finditer(r"cg n=(\d+) its=(\d+) res=(\S+)", (tmp / "solver.log").read_text())]
        dumps = [parse_dump(tmp / "solution" / f"solution-{k + 1:04d}.vtk", dim) for k in range(len(steps))]
    arrays = {}
    for fname, f in dumps[0].items():
        arrays[f"{fname}__x"] = f["x"]
        arrays[f"{fname}__comp"] = f["comp"]
        arrays[f"{fname}__v"] = np.stack([d[fname]["v"] for d in dumps])
    np.savez_compressed(HERE / f"reference_run_{name}.npz", **arrays)
    loop_log = res.stdout[res.stdout.index("starting time loop"):]  # run()'s own prints, FSS:325-406
    rec = {"case": name, "dim": dim, "degree_u": 1 if name.startswith("q1_") else 2, "input": inp, "n_steps": len(steps), "steps": steps, "cg_solves": cg,
           "fields": sorted(dumps[0].keys()),
           "time_loop_stdout": loop_log,
           "produced_by": "oracle/_ref/fss_ref = /root/reference/lib/include/*.h (unmodified) + oracle/dealii_shim (deal.II API shim, NOT deal.II) + oracle/ref_main.cpp"}
    (HERE / f"reference_run_{name}.json").write_text(json.dumps(rec, indent=1))
    return rec


if __name__ == "__main__" and len(sys.argv) == 4 and sys.argv[1] == "reduce":
    r = reduce_big(sys.argv[2], sys.argv[3])
    print(r["case"], "steps", r["n_steps"], "cg", [c["its"] for c in r["cg_solves"]])
    sys.exit(0)

if __name__ == "__main__":
    subprocess.check_call(["make", "-s", "-C", str(ROOT / "oracle"), "ref"])
    exe = ROOT / "oracle" / "_ref" / "fss_ref"
    for name in (sys.argv[1:] or CASES):
        rec = run_case(name, exe)
        print(name, "steps", rec["n_steps"], "cg solves", len(rec["cg_solves"]), "its", [c["its"] for c in rec["cg_solves"]][:12], "...",
              "limits", [s["solution_limits"] for s in rec["steps"]])
