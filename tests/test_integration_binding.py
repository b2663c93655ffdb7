"""INTEGRATION.md's reference-side binding as a program: integration/GpuBackend.h (upload_from_dealii) + integration/fss_gpu.cpp
(PoroElasticProblem::run() with every solver call replaced by its pe_* entry point), compiled against the deal.II API shim and
the reference's own InputDataPoroel.h / TensorIndexer.h.  Here it is linked with the oracle-backed pe_* of
tests/driver_on_oracle.cpp, so the C-ABI is driven exactly as the reference would drive it and answered on the CPU: the
program has to print the reference's time-loop log and reproduce the recorded fields of the reference's own run
(tests/golden/reference_run_*).  tests/test_zzz_gpu_reference_run.py runs the same program linked with libporoel.so."""
import subprocess
from pathlib import Path

import numpy as np
import pytest

import reference_run as R
from reference_run import H

REFERENCE = Path("/root/reference/lib/include")


def read_fields(path):
    tok = Path(path).read_text().split()
    n_p = int(tok[1])
    p = np.array(tok[2:2 + n_p], dtype=float)
    n_u = int(tok[3 + n_p])
    return p, np.array(tok[4 + n_p:4 + n_p + n_u], dtype=float)


@pytest.fixture(scope="module")
def fss_gpu_on_oracle(tmp_path_factory):
    if not REFERENCE.is_dir():
        pytest.skip("the reference's headers are not on this host")
    H.load_oracle()
    exe = tmp_path_factory.mktemp("integration") / "fss_gpu_on_oracle"
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-w", "-I", str(H.ROOT / "oracle" / "dealii_shim"), "-I", str(REFERENCE), "-I", str(H.ROOT / "include"),
                           "-I", str(H.ROOT / "integration"), "-o", str(exe), str(H.ROOT / "integration" / "fss_gpu.cpp"),
                           str(H.ROOT / "tests" / "driver_on_oracle.cpp"), "-L", str(H.ROOT / "oracle"), "-loracle", f"-Wl,-rpath,{H.ROOT / 'oracle'}"])
    return exe


@pytest.mark.parametrize("case", R.CASES)
def test_binding_reproduces_the_reference_run_through_the_c_abi(fss_gpu_on_oracle, case, tmp_path):
    rec, gold = R.load(case)
    (tmp_path / "solution").mkdir()
    (tmp_path / "input.data").write_text(rec["input"])
    env = {"PATH": "/usr/bin:/bin"}
    if rec.get("degree_u", 2) == 1:  # the q1_ records: the shim's run-time override of the hard-coded FE_Q(2) (shim_fe.h, FESystem)
        env["DEALII_SHIM_FESYSTEM_DEGREE"] = "1"
    out = subprocess.run([str(fss_gpu_on_oracle), "input.data"], cwd=tmp_path, capture_output=True, text=True, timeout=900, env=env)
    assert out.returncode == 0, out.stderr[-1000:]
    mine, theirs = out.stdout[out.stdout.index("starting time loop"):].splitlines(), rec["time_loop_stdout"].splitlines()
    assert len(mine) == len(theirs)
    for a, b in zip(mine, theirs):
        if a != b:  # a residual at the rounding floor (the capped case) may differ in its last digits, nothing else
            assert a.split()[0] == b.split()[0] == "Error:" and float(b.split()[1]) < 1e-12 and float(a.split()[1]) < 1e-12, (a, b)
    for k in range(rec["n_steps"]):
        p, u = read_fields(tmp_path / "solution" / f"fields-{k + 1:04d}.txt")
        # deal.II's numbering on both sides (the same DoFHandler code): dof by dof, no coordinate matching needed
        assert np.linalg.norm(p - gold["p__v"][k]) <= 1e-12 * np.linalg.norm(gold["p__v"][k])
        assert np.linalg.norm(u - gold["u__v"][k]) <= 1e-10 * np.linalg.norm(gold["u__v"][k])


def test_integration_md_quotes_the_binding_that_is_compiled():
    """the upload function printed in INTEGRATION.md is the one in integration/GpuBackend.h (statement by statement, whitespace aside)"""
    import re
    md = (H.ROOT / "INTEGRATION.md").read_text()
    src = (H.ROOT / "integration" / "GpuBackend.h").read_text()
    squeeze = lambda s: re.sub(r"\s+", "", s)
    for stmt in ["pe_upload_mesh(ctx, dim, tria.n_vertices(), xyz.data(), tria.n_active_cells(), cells.data(),",
                 "pe_upload_dofs(ctx, PE_FIELD_PRESSURE, p_dh.n_dofs(), cd_p.data())",
                 "pe_upload_dofs(ctx, PE_FIELD_DISPLACEMENT, u_dh.n_dofs(), cd_u.data())",
                 "g.push_back(u_constraints.get_inhomogeneity(i));",
                 "for (unsigned v = 0; v < GeometryInfo<dim>::vertices_per_cell; ++v) cells.push_back(cell->vertex_index(v));",
                 "prm.perm_over_visc = data.perm / data.visc;"]:
        assert squeeze(stmt) in squeeze(src), stmt
        assert squeeze(stmt) in squeeze(md), stmt
