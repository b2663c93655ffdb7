"""Shared by the CPU and GPU tests that compare a backend with the committed records of the REFERENCE'S OWN CODE
(tests/golden/reference_run_*.{json,npz}, produced by tests/golden/make_reference_run.py from oracle/_ref/fss_ref: the unmodified
sources of /root/reference/lib/include compiled against the deal.II API shim of oracle/dealii_shim — NOT deal.II)."""
import json

import numpy as np

import helpers as H

capi, fss = H.capi, H.fss
GOLD = H.ROOT / "tests" / "golden"
CASES = ["shipped_4steps", "box3d_r3", "neumann2d_r4", "rect2d_r3", "caps2d_r3", "neumann3d_r2", "shipped_r6", "q1_box3d_r4", "q1_neumann2d_r5"]
# what the reference leaves to its defaults / hard-codes: FE_Q(2) displacement (DS:67), uniform mesh until time step 5 (FSS:333)
GPU_SECTION = "\nsubsection GPU\n  set Displacement FE degree = {degree}\n  set Refine every = 0\n{extra}end\n"


def load(case):
    return json.loads((GOLD / f"reference_run_{case}.json").read_text()), np.load(GOLD / f"reference_run_{case}.npz")


def problem(rec, backend, extra=""):
    inp = capi.InputData(text=rec["input"] + GPU_SECTION.format(extra=extra, degree=rec.get("degree_u", 2)))
    mesh = fss.make_mesh(inp)
    dofs_p, dofs_u, _ = fss.upload_problem(backend, inp, mesh)
    return inp, dofs_p, dofs_u


def dof_order(ref_x, ref_comp, support_points, n_comp):
    """for every dof of the record: the backend dof with the same support point and component"""
    key = lambda x, c: tuple(np.round(np.asarray(x) * 1e6).astype(np.int64)) + (int(c),)
    mine = {key(support_points[i], i % n_comp): i for i in range(support_points.shape[0])}
    assert len(mine) == support_points.shape[0]
    return np.array([mine[key(ref_x[k], ref_comp[k])] for k in range(ref_x.shape[0])])


def split_cg_log(rec, n_p, n_u):
    """The shim's solver log in call order -> (initialisation, [per time step]) with the solves named by their place in
    run(): FSS:310-317 is one displacement solve and dim projections; a coupling iteration is k pressure solves, one displacement
    solve and dim projections of the normal strains; a time step ends with the (zero right-hand side) shear projections of FSS:409."""
    dim, log = rec["dim"], list(rec["cg_solves"])
    n_shear = 1 if dim == 2 else 3
    take = lambda k: [log.pop(0) for _ in range(k)]
    first = take(1 + dim)
    assert first[0]["n"] == n_u and all(c["n"] == n_p for c in first[1:])
    init = {"displacement": first[0]["its"], "projection": sum(c["its"] for c in first[1:])}
    steps = []
    for s in rec["steps"]:
        step = {"pressure": [], "pressure_solves_per_coupling_iteration": [], "displacement": [], "displacement_res": [], "projection": []}
        for _ in range(s["coupling_iterations"]):
            k = 0
            while log[0]["n"] == n_p and n_p != n_u:  # the pressure solves of this coupling iteration end at the displacement solve
                step["pressure"].append(log.pop(0)["its"])
                k += 1
            step["pressure_solves_per_coupling_iteration"].append(k)
            disp = take(1)[0]
            assert disp["n"] == n_u
            step["displacement"].append(disp["its"])
            step["displacement_res"].append(disp["res"])
            proj = take(dim)
            assert all(c["n"] == n_p for c in proj)
            step["projection"] += [c["its"] for c in proj]
        shear = take(n_shear)
        assert all(c["n"] == n_p and c["its"] == 0 and c["res"] == 0.0 for c in shear)  # FSS:167-176 never assembles these right-hand sides
        steps.append(step)
    assert not log
    return init, steps


def expected_prints(rep, pressure_tol):
    """FSS:349-381 from a mirror report: per coupling iteration the "pressure converged; iterations: passes - 1" print (a pass that
    finds the residual below tolerance leaves without solving; a loop that runs into `Max pressure iterations` prints nothing and has
    solved in every pass) -> (printed numbers, pressure solves per coupling iteration)"""
    history, at, printed, solves = rep["residual_history"], 0, [], []
    for passes in rep["inner_counts"]:
        converged = history[at + passes - 1] < pressure_tol
        printed += [passes - 1] if converged else []
        solves.append(passes - 1 if converged else passes)
        at += passes
    return printed, solves
