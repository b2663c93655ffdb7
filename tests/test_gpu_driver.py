"""The C++ driver executable (the main() the reference never shipped) on the reference's shipped input."""
import re
import subprocess

import pytest

import helpers as H

pytestmark = pytest.mark.gpu
BIN = H.ROOT / "poroelasticity-dealii_b200" / "bin" / "fss-poroel"


def test_fss_poroel_runs_the_shipped_case(tmp_path):
    # input.data as shipped (2D, refine 4, Q2/Q1, dt 60, t_max 1e3 => 17 steps) on the uniform mesh (the adaptive schedule: test_zz_gpu_amr.py)
    text = H.SHIPPED_INPUT + "\nsubsection GPU\n  set Refine every = 0\n  set CG max iterations = 5000\nend\n"
    f = tmp_path / "input.data"
    f.write_text(text)
    out = subprocess.run([str(BIN), str(f)], capture_output=True, text=True, timeout=300, cwd=tmp_path)
    assert out.returncode == 0, out.stderr[-2000:]
    s = out.stdout
    assert "# Listing of Parameters" in s and "set Initial refinement level" in s   # prm.print_parameters (ID:82)
    assert "starting time loop" in s and "time max 1000" in s                       # FSS:325-326
    times = [float(x) for x in re.findall(r"^Time: (\S+)", s, flags=re.M)]
    assert times == [60.0 * k for k in range(1, 18)]                                # while (time < t_max), FSS:327
    assert len(re.findall(r"Coupling iteration: 1$", s, flags=re.M)) == 17          # one FSS iteration per step (FSS:399 off)
    assert "Coupling iteration: 2" not in s
    conv = [int(x) for x in re.findall(r"pressure converged; iterations: (\d+)", s)]
    assert len(conv) == 17 and conv[0] == 5                                          # oracle: 6 residual evaluations in step 1
    errs = [float(x) for x in re.findall(r"Error: (\S+)", s)]
    assert len(errs) == 17 and all(e < 1e-8 for e in errs)


def test_fss_poroel_reports_errors(tmp_path):
    f = tmp_path / "bad.data"
    f.write_text("subsection Mesh\n  set Dimensions = 7\nend\n")
    out = subprocess.run([str(BIN), str(f)], capture_output=True, text=True, timeout=60, cwd=tmp_path)
    assert out.returncode == 1 and "Exception on processing" in out.stderr and "does not match" in out.stderr
    out = subprocess.run([str(BIN)], capture_output=True, text=True, timeout=60)
    assert out.returncode == 1 and "specify the file name" in out.stdout              # PCL:7-10


def test_fss_poroel_writes_vtk(tmp_path):
    """the minimal legacy-VTK writer behind `Write VTK = 1` (FSS:227-291 writes ./solution/solution-NNNN.vtk)."""
    text = H.make_input(dim=3, refine=2, degree_u=1, extra_gpu="  set Refine every = 0\n  set Write VTK = 1\n  set Max time steps = 2\n")
    f = tmp_path / "input.data"
    f.write_text(text)
    (tmp_path / "solution").mkdir()
    out = subprocess.run([str(BIN), str(f)], capture_output=True, text=True, timeout=300, cwd=tmp_path)
    assert out.returncode == 0, out.stderr[-2000:]
    files = sorted((tmp_path / "solution").glob("solution-*.vtk"))
    assert [p.name for p in files] == ["solution-0001.vtk", "solution-0002.vtk"]
    body = files[-1].read_text()
    assert "DATASET UNSTRUCTURED_GRID" in body and "POINTS 125 double" in body and "CELLS 64 576" in body
    for name in ("VECTORS u double", "SCALARS p double 1", "SCALARS eps_xx double 1", "SCALARS sigma_zz double 1"):
        assert name in body


def test_fss_poroel_default_iteration_cap(tmp_path):
    """The shipped case also runs with the reference's own 1000-iteration CG cap (PS:175, DS:299)."""
    f = tmp_path / "input.data"
    f.write_text(H.SHIPPED_INPUT + "\nsubsection GPU\n  set Refine every = 0\n  set Max time steps = 3\nend\n")
    out = subprocess.run([str(BIN), str(f)], capture_output=True, text=True, timeout=300, cwd=tmp_path)
    assert out.returncode == 0, out.stderr[-2000:]
    assert out.stdout.count("Coupling iteration: 1") == 3


def test_fss_poroel_runs_the_shipped_input_as_is_with_refinement(tmp_path):
    """`fss-poroel input.data` on the reference's file exactly as shipped (no GPU subsection): the as-is schedule of FSS:333-340
    is the default, so the run refines at steps 5, 10 and 15 and completes all 17 steps on the hanging-node path."""
    f = tmp_path / "input.data"
    f.write_text(H.SHIPPED_INPUT)
    out = subprocess.run([str(BIN), str(f)], capture_output=True, text=True, timeout=600, cwd=tmp_path)
    assert out.returncode == 0, out.stderr[-2000:]
    log = out.stdout
    assert log.count("Time: ") == 17 and log.count("Refining mesh") == 3
    assert log.count("Coupling iteration: 1") == 17 and "Coupling iteration: 2" not in log
    lines = log.splitlines()
    for step in (5, 10, 15):
        i = lines.index(f"Time: {60 * step}")
        assert lines[i + 1] == "Refining mesh" and "active cells" in lines[i + 2]
    cells = [int(x) for x in re.findall(r"active cells: (\d+)", log)] or [int(x) for x in re.findall(r"(\d+) active cells", log)]
    assert cells and max(cells) > 256
