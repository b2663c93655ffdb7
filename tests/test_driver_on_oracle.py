"""The product's C++ driver (csrc/host/host_capi.cpp + problem.hpp) executed on the CPU.

libporoel_host.so links the CUDA library, so PoroElasticProblem::initialize / step / run / refine_mesh / output_results
cannot run without a GPU.  tests/driver_on_oracle.cpp forwards the pe_* C-ABI to the CPU oracle; the tests build
host_capi.cpp + that shim into a private library and compare the C++ driver with the Python mirror of the same loop
(fss.py) on the same oracle.  Both issue the same operator calls in the same order, so the fields must agree bit for bit:
any difference is a host-side bug (mesh, numbering, constraint tables, upload order, control flow of the time loop).
"""
import ctypes as C
import subprocess

import numpy as np
import pytest

import helpers as H

capi, fss = H.capi, H.fss


@pytest.fixture(scope="module")
def oracle_driver(tmp_path_factory):
    H.load_oracle()  # makes sure oracle/liboracle.so is built
    out = tmp_path_factory.mktemp("drv") / "libdriver_on_oracle.so"
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-Wall", "-Wl,-Bsymbolic", "-o", str(out),
                           str(H.ROOT / "poroelasticity-dealii_b200" / "csrc" / "host" / "host_capi.cpp"), str(H.ROOT / "tests" / "driver_on_oracle.cpp"),
                           "-L", str(H.ROOT / "oracle"), "-loracle", f"-Wl,-rpath,{H.ROOT / 'oracle'}"])
    lib = C.CDLL(str(out))
    capi._declare_host_api(lib)
    capi._declare_operator_api(lib, "pe_")
    assert lib.pe_version() == -1
    return lib


@pytest.fixture()
def on_oracle(oracle_driver, monkeypatch):
    """capi.InputData / capi.Problem / OperatorBackend("pe_") all go through the oracle-backed build inside the test."""
    monkeypatch.setattr(capi, "_host", oracle_driver)
    monkeypatch.setattr(capi, "_dev", oracle_driver)
    return oracle_driver


def mirror_run(text, n_steps):
    inp = capi.InputData(text=text)
    b = H.create_oracle_backend()
    mesh = fss.make_mesh(inp)
    dp, du, _ = fss.upload_problem(b, inp, mesh)
    fss.initialize(b, inp)
    reps = [fss.time_step(b, inp) for _ in range(n_steps)]
    return b, reps


@pytest.mark.parametrize("kw", [dict(dim=2, refine=3, degree_u=2), dict(dim=3, refine=2, degree_u=1), dict(dim=3, refine=2, degree_u=1, cells=[3, 2, 4]),
                                dict(dim=2, refine=3, degree_u=1, neumann=([3], [1], [-2e6]), dirichlet=([0, 1, 2], [0, 0, 1], [0, 0, 0]))])
def test_cpp_driver_equals_the_python_mirror_on_uniform_meshes(on_oracle, kw):
    text = H.make_input(**kw)
    prob = capi.Problem(capi.InputData(text=text), device=0)
    prob.initialize()
    reps = [prob.step() for _ in range(3)]
    b, mreps = mirror_run(text, 3)
    for r, m in zip(reps, mreps):
        assert r["fss_iterations"] == m["fss_iterations"] and r["pressure_iterations"] == m["pressure_iterations"]
        assert (r["cg_its_pressure"], r["cg_its_displacement"], r["cg_its_projection"]) == (m["cg_its_pressure"], m["cg_its_displacement"], m["cg_its_projection"])
        assert r["pressure_error"] == m["pressure_error"]
    for which in (capi.VEC_P, capi.VEC_U, capi.VEC_VOL_STRAIN, capi.VEC_VOL_STRAIN0):
        assert np.array_equal(prob.backend.get_vector(which), b.get_vector(which))
    assert np.array_equal(prob.global_ids(capi.FIELD_PRESSURE), np.arange(prob.backend.n_p))
    prob.close()
    b.close()


def test_cpp_driver_adaptive_run_equals_the_python_mirror(on_oracle):
    """Refine every = 5 on the shipped input: PoroElasticProblem::refine_mesh (problem.hpp) against fss.refine_mesh."""
    text = H.SHIPPED_INPUT + "\nsubsection GPU\n  set Refine every = 5\nend\n"
    inp = capi.InputData(text=text)
    ora = H.create_oracle_backend()
    snaps = {}

    def on_step(step, rep, mesh, dp, du):
        snaps[step] = (ora.get_vector(capi.VEC_P), ora.get_vector(capi.VEC_U), ora.get_vector(capi.VEC_VOL_STRAIN), mesh.arrays.n_cells, rep)

    fss.run_adaptive(ora, inp, 17, inp.refine_every, on_step)
    prob = capi.Problem(capi.InputData(text=text), device=0)
    prob.initialize()
    cells = []
    for step in range(1, 18):
        rep = prob.step()
        st = prob.backend.stats()
        p, u, ev, n_cells, mrep = snaps[step]
        cells.append(int(st["n_cells"]))
        assert st["n_cells"] == n_cells, step
        assert np.array_equal(prob.backend.get_vector(capi.VEC_P), p), step
        assert np.array_equal(prob.backend.get_vector(capi.VEC_U), u), step
        assert np.array_equal(prob.backend.get_vector(capi.VEC_VOL_STRAIN), ev), step
        assert rep["pressure_iterations"] == mrep["pressure_iterations"] and rep["cg_its_displacement"] == mrep["cg_its_displacement"]
    assert cells[3] == 256 and cells[4] > 256 and cells[9] > cells[4] and cells[14] > cells[9]
    prob.close()
    ora.close()


def test_cpp_driver_surfaces_solver_failures(on_oracle):
    """SolverControl::NoConvergence (PS:175, DS:299) is never swallowed: it comes back as an error of the call."""
    text = H.make_input(dim=2, refine=3, degree_u=2, extra_gpu="  set CG max iterations = 2\n")
    prob = capi.Problem(capi.InputData(text=text), device=0)
    with pytest.raises(capi.HostError, match="NoConvergence"):
        prob.initialize()
    prob.close()


@pytest.mark.parametrize("degree_u,refine_every", [(1, 0), (2, 0), (2, 2)])
def test_vtk_writer(on_oracle, tmp_path, monkeypatch, degree_u, refine_every):
    """FSS:227-291: ./solution/solution-NNNN.vtk with the vertex values of u (FE_Q(1) and FE_Q(2)), p, strains and stresses;
    on adaptive meshes only the vertices in use are written."""
    monkeypatch.chdir(tmp_path)
    (tmp_path / "solution").mkdir()
    text = H.make_input(dim=2, refine=3, degree_u=degree_u, extra_gpu=f"  set Write VTK = 1\n  set Max time steps = 3\n  set Refine every = {refine_every}\n")
    prob = capi.Problem(capi.InputData(text=text), device=0)
    prob.run(verbose=False)
    st = prob.backend.stats()
    files = sorted((tmp_path / "solution").glob("solution-*.vtk"))
    assert [f.name for f in files] == ["solution-0001.vtk", "solution-0002.vtk", "solution-0003.vtk"]
    tok = files[-1].read_text().split()
    n_pts = int(tok[tok.index("POINTS") + 1])
    n_cells = int(tok[tok.index("CELLS") + 1])
    assert n_cells == st["n_cells"] and n_pts == st["n_dofs_p"]  # Q1 pressure: one dof per vertex in use
    if refine_every:
        assert n_cells > 64
    pts = np.array(tok[tok.index("POINTS") + 3: tok.index("POINTS") + 3 + 3 * n_pts], dtype=float).reshape(-1, 3)
    i0 = tok.index("CELLS") + 3
    conn = np.array(tok[i0: i0 + 5 * n_cells], dtype=int).reshape(-1, 5)
    assert (conn[:, 0] == 4).all() and conn[:, 1:].min() == 0 and conn[:, 1:].max() == n_pts - 1
    # VTK_QUAD is counter-clockwise: positive area for every cell, areas tile the box
    q = pts[conn[:, 1:], :2]
    d1, d2 = q[:, 2] - q[:, 0], q[:, 3] - q[:, 1]
    area = 0.5 * np.abs(d1[:, 0] * d2[:, 1] - d1[:, 1] * d2[:, 0])
    assert np.isclose(area.sum(), 100.0)
    k = tok.index("p") - 1  # "SCALARS p double 1 LOOKUP_TABLE default"
    assert tok[k] == "SCALARS"
    pv = np.array(tok[k + 6: k + 6 + n_pts], dtype=float)
    p = prob.backend.get_vector(capi.VEC_P)
    assert np.allclose(np.sort(pv), np.sort(p), rtol=1e-11)
    k = tok.index("VECTORS")
    uv = np.array(tok[k + 3: k + 3 + 3 * n_pts], dtype=float).reshape(-1, 3)
    # the roller boundary data: u_x = 0 on x = -5, u_x = -1e-5 on x = +5 (input.data:14-16)
    assert np.allclose(uv[np.isclose(pts[:, 0], -5.0), 0], 0.0, atol=1e-18) and np.allclose(uv[np.isclose(pts[:, 0], 5.0), 0], -1e-5, rtol=1e-11)
    assert np.all(uv[:, 2] == 0)
    for name in ("eps_xx", "eps_xy", "eps_yy", "sigma_xx", "sigma_xy", "sigma_yy"):
        assert name in tok
    prob.close()


def test_fss_poroel_executable_runs_the_shipped_input_with_refinement(tmp_path):
    """csrc/host/main.cpp (`fss-poroel <input.data>`, PCL:5-27) built against the oracle shim: the reference's input.data
    as shipped (no GPU subsection; `Refine every` defaults to the reference's 5) runs all 17 steps, refines at steps 5, 10 and 15 (FSS:333-340) and prints the reference's log."""
    H.load_oracle()
    exe = tmp_path / "fss-poroel-on-oracle"
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-Wall", "-o", str(exe), str(H.ROOT / "poroelasticity-dealii_b200" / "csrc" / "host" / "main.cpp"),
                           str(H.ROOT / "tests" / "driver_on_oracle.cpp"), "-L", str(H.ROOT / "oracle"), "-loracle", f"-Wl,-rpath,{H.ROOT / 'oracle'}"])
    f = tmp_path / "input.data"
    f.write_text(H.SHIPPED_INPUT)  # exactly the reference's file: the every-5th-step refinement is the default
    out = subprocess.run([str(exe), str(f)], capture_output=True, text=True, timeout=600, cwd=tmp_path)
    assert out.returncode == 0, out.stderr[-2000:]
    log = out.stdout
    assert log.count("Time: ") == 17 and log.count("Refining mesh") == 3
    assert log.count("Coupling iteration: 1") == 17 and "Coupling iteration: 2" not in log  # as-is: one coupling iteration (FSS:399)
    lines = log.splitlines()
    for step in (5, 10, 15):
        i = lines.index(f"Time: {60 * step}")
        assert lines[i + 1] == "Refining mesh" and "active cells" in lines[i + 2]
    no_file = subprocess.run([str(exe)], capture_output=True, text=True, timeout=60)
    assert no_file.returncode == 1 and "specify the file name" in no_file.stdout


def test_cpp_driver_reads_domain_msh_and_refines_it(on_oracle, tmp_path, monkeypatch):
    """read_mesh() (FSS:438-445: "domain.msh" in the working directory) as the coarse mesh of the adaptive loop; the Gmsh
    file uses the boundary ids of the reference's domain.geo:22-25 (0 bottom, 1 right, 2 top, 3 left)."""
    import shutil
    monkeypatch.chdir(tmp_path)
    shutil.copy(H.ROOT / "tests" / "golden" / "square10.msh", tmp_path / "domain.msh")
    text = H.make_input(dim=2, refine=2, degree_u=2, dirichlet=([3, 1, 0, 2], [0, 0, 1, 1], [0, -1e-5, 0, -1e-5]),
                        extra_gpu="  set Read mesh file = 1\n  set Refine every = 2\n")
    inp = capi.InputData(text=text)
    ora = H.create_oracle_backend()
    snaps = {}
    fss.run_adaptive(ora, inp, 4, inp.refine_every, lambda s, r, m, dp, du: snaps.__setitem__(s, (ora.get_vector(capi.VEC_P), m.arrays.n_cells)))
    prob = capi.Problem(capi.InputData(text=text), device=0)
    prob.initialize()
    for step in range(1, 5):
        prob.step()
        p, n_cells = snaps[step]
        assert prob.backend.stats()["n_cells"] == n_cells
        assert np.array_equal(prob.backend.get_vector(capi.VEC_P), p)
    assert snaps[1][1] == 100 and snaps[2][1] > 100 and snaps[4][1] > snaps[2][1]
    prob.close()
    ora.close()
