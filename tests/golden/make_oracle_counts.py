"""Runs the CPU oracle (SSOR-CG, the reference's solver settings) on the 3D Q1/Q1 benchmark configs and records, per time
step, iteration counts, phase timings, field checksums (tests/golden/oracle_counts_r<refine>.json) and the values of p and u
at 4096 fixed lattice nodes (tests/golden/oracle_fields_r<refine>.npz; keyed by lattice index (i, j, k), i.e. by support-point
coordinate x = -L/2 + L i/n, plus the global dof numbers of the host library's first-touch numbering for a fast look-up).
These records are what bench.py's bounded CPU sample extrapolates with (it cannot afford full oracle steps at 128^3 inside a
default run), what bench.py checks the fields of every run against (`parity`), and what tests/test_gpu_golden.py compares.
Usage:  python tests/golden/make_oracle_counts.py <refine> [steps] [max_its] [threads]
        python tests/golden/make_oracle_counts.py c2 [steps] [max_its] [threads] [refine]     # 2D consolidation (BASELINE configs[1]):
                                                  # top traction, rollers elsewhere, refine 9 -> oracle_counts_c2_r9.json

The record is rewritten after every step, so an interrupted run leaves a valid (shorter) record."""
import json
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import helpers as H  # noqa: E402

N_SAMPLES = 4096
SEED = 20261018


def sample_lattice(refine, dim=3):
    """The fixed sample nodes: N_SAMPLES distinct lattice indices (i, j[, k]) in [0, 2^refine]^dim."""
    n = 2 ** refine
    rng = np.random.RandomState(SEED)
    flat = rng.choice((n + 1) ** dim, size=min(N_SAMPLES, (n + 1) ** dim), replace=False)
    flat.sort()
    return np.stack([(flat // (n + 1) ** a) % (n + 1) for a in range(dim)], axis=1).astype(np.int32)


def lattice_of(points, refine, L=10.0):
    n = 2 ** refine
    return np.rint((np.asarray(points) + L / 2) / L * n).astype(np.int64)


def dofs_at(ijk, support_points, refine, n_comp):
    """Global dof number of (lattice node, component 0) for every sample, from the dofs' support points."""
    n = 2 ** refine
    lat = lattice_of(support_points[::n_comp], refine)
    flat = lambda a: sum(a[:, c].astype(np.int64) * (n + 1) ** c for c in range(a.shape[1]))
    key = flat(lat)
    order = np.argsort(key)
    want = flat(ijk)
    pos = np.searchsorted(key[order], want)
    assert (key[order][pos] == want).all()
    return (order[pos] * n_comp).astype(np.int64)


if __name__ == "__main__":
    c2 = sys.argv[1] == "c2"
    refine = (int(sys.argv[5]) if len(sys.argv) > 5 else 9) if c2 else int(sys.argv[1])
    tag = f"c2_r{refine}" if c2 else f"r{refine}"
    dim = 2 if c2 else 3
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    max_its = int(sys.argv[3]) if len(sys.argv) > 3 else 1000
    threads = H.load_oracle().po_set_threads(int(sys.argv[4]) if len(sys.argv) > 4 else 0)
    if c2:  # the same text as bench.py --workload c2 (undrained top load DS:249-277, rollers elsewhere)
        text = H.make_input(dim=2, refine=refine, degree_u=1, dirichlet=([0, 1, 2], [0, 0, 1], [0.0, 0.0, 0.0]), neumann=([3], [1], [-1e6]))
    else:
        text = H.make_input(dim=3, refine=refine, degree_u=1)
    inp = H.capi.InputData(text=text)
    mesh = H.fss.make_mesh(inp)
    prm = inp.params()
    prm.cg_max_iterations = max_its
    b = H.create_oracle_backend()
    t0 = time.time()
    dofs_p, dofs_u, _ = H.fss.upload_problem(b, inp, mesh, prm)
    t_setup = time.time() - t0
    ijk = sample_lattice(refine, dim)
    pid = dofs_at(ijk, dofs_p.support_points(), refine, 1)
    uid = dofs_at(ijk, dofs_u.support_points(), refine, dim)
    t0 = time.time()
    init = H.fss.initialize(b, inp)
    t_init = time.time() - t0
    out = {"refine": refine, "threads": threads, "cg_max_iterations": max_its, "setup_s": t_setup, "init_s": t_init, "init": init,
           "stats": b.stats(), "fields": f"oracle_fields_{tag}.npz", "workload": "c2" if c2 else "cube", "steps": []}
    print(out, flush=True)
    here = Path(__file__).parent
    P, U = [], []

    def grab():
        p, u = b.get_vector(H.capi.VEC_P), b.get_vector(H.capi.VEC_U)
        P.append(p[pid].copy())
        U.append(np.stack([u[uid + a] for a in range(dim)], axis=1))
        return p, u

    grab()  # entry 0 = the state after initialisation (FSS:310-317)
    for s in range(steps):
        t0 = time.time()
        rep = H.fss.time_step(b, inp)
        rep["wall_s"] = time.time() - t0
        p, u = grab()
        rep["p_l2"], rep["u_l2"], rep["p_sum"] = float(np.linalg.norm(p)), float(np.linalg.norm(u)), float(p.sum())
        out["steps"].append(rep)
        print(rep, flush=True)
        (here / f"oracle_counts_{tag}.json").write_text(json.dumps(out, indent=1))
        np.savez(here / f"oracle_fields_{tag}.npz", ijk=ijk, p_dof=pid, u_dof=uid, p=np.array(P), u=np.array(U))
