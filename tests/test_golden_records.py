"""The committed oracle records under tests/golden (what bench.py's `parity` and tests/test_gpu_golden.py compare the CUDA path
with at full size) checked on the CPU: internal consistency, reproducibility of the sample keys, a replay of the small
record with today's oracle (a record that went stale against oracle.cpp would otherwise only show up on a GPU box), and a
closed form for the 2D consolidation workload (BASELINE configs[1])."""
import json

import numpy as np
import pytest

import helpers as H
import golden.make_oracle_counts as G

capi, fss = H.capi, H.fss
GOLD = H.ROOT / "tests" / "golden"
C2_TEXT = dict(dim=2, degree_u=1, dirichlet=([0, 1, 2], [0, 0, 1], [0.0, 0.0, 0.0]), neumann=([3], [1], [-1e6]))


def load(tag):
    return json.loads((GOLD / f"oracle_counts_{tag}.json").read_text()), np.load(GOLD / f"oracle_fields_{tag}.npz")


@pytest.mark.parametrize("tag,dim,refine", [("r5", 3, 5), ("r6", 3, 6), ("r7", 3, 7), ("c2_r9", 2, 9)])
def test_record_is_consistent(tag, dim, refine):
    rec, f = load(tag)
    n_steps = len(rec["steps"])
    assert n_steps >= 4 and rec["refine"] == refine
    assert f["p"].shape == (n_steps + 1, f["ijk"].shape[0]) and f["u"].shape == (n_steps + 1, f["ijk"].shape[0], dim)
    assert np.array_equal(G.sample_lattice(refine, dim), f["ijk"])  # the sample is a function of (seed, refine, dim) only
    assert np.isfinite(f["p"]).all() and np.isfinite(f["u"]).all()
    n = 2 ** refine
    assert rec["stats"]["n_dofs_p"] == (n + 1) ** dim and rec["stats"]["n_dofs_u"] == dim * (n + 1) ** dim
    for s in rec["steps"]:
        assert s["fss_iterations"] == 1 and s["pressure_error"] < 1e-8  # as-is control flow (FSS:345-405), PS tolerance reached
        assert all(s[k] >= 0 for k in ("cg_its_pressure", "cg_its_displacement", "cg_its_projection"))
    # the transient decays: no later step needs more displacement-CG iterations than the first
    its = [s["cg_its_displacement"] for s in rec["steps"]]
    assert max(its) == its[0]


@pytest.mark.parametrize("tag,dim,refine", [("r5", 3, 5), ("c2_r9", 2, 9)])
def test_sample_dofs_are_where_the_record_says(tag, dim, refine):
    """coordinate key -> dof number of the host library's first-touch numbering (what bench.py looks the samples up by)"""
    _, f = load(tag)
    text = H.make_input(refine=refine, **C2_TEXT) if tag.startswith("c2") else H.make_input(dim=3, refine=refine, degree_u=1)
    mesh = fss.make_mesh(capi.InputData(text=text))
    sp_p = capi.HostDofs(mesh, 1, 1).support_points()
    sp_u = capi.HostDofs(mesh, 1, dim).support_points()
    assert np.array_equal(G.lattice_of(sp_p[f["p_dof"]], refine), f["ijk"])
    assert np.array_equal(G.lattice_of(sp_u[f["u_dof"]], refine), f["ijk"])
    assert np.array_equal(G.lattice_of(sp_u[f["u_dof"] + dim - 1], refine), f["ijk"])  # components of a node are consecutive dofs


def test_small_record_replays_with_todays_oracle():
    rec, f = load("r5")
    inp = capi.InputData(text=H.make_input(dim=3, refine=5, degree_u=1))
    mesh = fss.make_mesh(inp)
    prm = inp.params()
    prm.cg_max_iterations = rec["cg_max_iterations"]
    b = H.create_oracle_backend()
    try:
        fss.upload_problem(b, inp, mesh, prm)
        fss.initialize(b, inp)
        for k in range(2):
            rep = fss.time_step(b, inp)
            gold = rec["steps"][k]
            for key in ("pressure_iterations", "cg_its_pressure", "cg_its_displacement", "cg_its_projection"):
                assert rep[key] == gold[key]
            p, u = b.get_vector(capi.VEC_P), b.get_vector(capi.VEC_U)
            assert fss.rel_l2(p[f["p_dof"]], f["p"][k + 1]) <= 1e-12
            assert fss.rel_l2(np.stack([u[f["u_dof"] + a] for a in range(3)], axis=1), f["u"][k + 1]) <= 1e-12
            assert float(np.linalg.norm(p)) == pytest.approx(gold["p_l2"], rel=1e-12)
    finally:
        b.close()


def test_c2_record_starts_from_the_closed_form_consolidation_state():
    """Undrained top load on a laterally confined column (rollers on the sides and the bottom, traction t on the top, uniform
    p = p_init): uniaxial strain, sigma_yy = (lambda + 2G) eps_yy - alpha p = t, so u_y = (y + L/2)(t + alpha p)/(lambda + 2G) and
    u_x = 0 — a linear field the Q1 space holds exactly (DS:249-277 face term against DS:206-247 cell terms)."""
    rec, f = load("c2_r9")
    inp = capi.InputData(text=H.make_input(refine=9, **C2_TEXT))
    prm = inp.params()
    L, n = 10.0, 2 ** 9
    y = -L / 2 + L * f["ijk"][:, 1] / n
    exact = (y + L / 2) * (-1e6 + prm.biot_coef * inp.p_init) / (prm.lame_lambda + 2 * prm.shear_modulus)
    assert np.abs(f["u"][0][:, 1] - exact).max() <= 1e-9 * np.abs(exact).max()
    assert np.abs(f["u"][0][:, 0]).max() <= 1e-9 * np.abs(exact).max()
    assert np.array_equal(f["p"][0], np.full(f["ijk"].shape[0], inp.p_init))
    # the well source then moves the pressure away from p_init around the axis; the lateral rollers keep u_x(x = +-L/2) = 0
    assert np.abs(f["p"][-1] - inp.p_init).max() > 1e3
    side = (f["ijk"][:, 0] == 0) | (f["ijk"][:, 0] == n)
    assert side.any() and np.abs(f["u"][-1][side, 0]).max() == 0.0
