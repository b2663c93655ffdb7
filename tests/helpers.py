"""Shared test plumbing: package import, oracle binding, canned inputs."""
import ctypes as C
import importlib
import subprocess
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))
pkg = importlib.import_module("poroelasticity-dealii_b200")
capi = pkg.capi
fss = pkg.fss

_oracle = None


def load_oracle():
    """The CPU oracle (oracle/liboracle.so) — test infrastructure, never used by the product."""
    global _oracle
    if _oracle is None:
        path = ROOT / "oracle" / "liboracle.so"
        if not path.exists():
            subprocess.check_call(["make", "-C", str(ROOT / "oracle")])
        lib = C.CDLL(str(path))
        capi._declare_operator_api(lib, "po_")
        lib.po_create.argtypes = [C.POINTER(C.c_void_p)]
        lib.po_set_threads.argtypes = [C.c_int]
        lib.po_set_preconditioner.argtypes = [C.c_void_p, C.c_int]
        lib.po_time_vmult.argtypes = [C.c_void_p, C.c_int, C.c_int]
        lib.po_time_ssor.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_int]
        _oracle = lib
    return _oracle


def create_oracle_backend(precond=-1):
    lib = load_oracle()
    ctx = C.c_void_p()
    assert lib.po_create(C.byref(ctx)) == 0
    lib.po_set_preconditioner(ctx, precond)
    return capi.OperatorBackend(lib, "po_", ctx)


SHIPPED_INPUT = """
subsection Mesh
  set Dimensions               = 2
  set Domain size              = 10, 10
  set Initial refinement level = 4
  set Max refinement level     = 6
end
subsection In situ
  set Displacement boundary labels     = 0, 1, 2, 3
  set Displacement boundary components = 0, 0, 1, 1
  set Displacement boundary values     = 0, -1e-5, 0, -1e-5
  set Initial pressure                 = 10e6
  set Stress boundary components       =
  set Stress boundary labels           =
  set Stress boundary values           =
end
subsection Properties
  set Young modulus         = 1.4e10
  set Biot coefficient      = 0.9
  set Bulk density          = 2700
  set Fluid compressibility = 5.8e-10   # 1.45e-10 for water
  set Permeability          = 10          # in mDa
  set Poisson ratio         = 0.3
  set Porosity              = 0.3
  set Viscosity             = 1e-3
  set Well radius           = 1
  set Flow rate             = 1e-5
end
subsection Solver
  set Time step  = 60
  set Time max   = 1e3
end
"""


def make_input(dim=2, refine=4, degree_u=2, extra_gpu="", cells=None, neumann=None, dirichlet=None):
    """input.data text with the shipped properties (input.data:24-35), extended to 3D as SURVEY §8d says."""
    size = ", ".join(["10"] * dim)
    if dirichlet is None:
        labels = ", ".join(str(i) for i in range(2 * dim))
        comps = ", ".join(str(i // 2) for i in range(2 * dim))
        vals = ", ".join("0" if i % 2 == 0 else "-1e-5" for i in range(2 * dim))
    else:
        labels, comps, vals = (", ".join(str(x) for x in col) for col in dirichlet)
    nl, nc, nv = (", ".join(str(x) for x in col) for col in neumann) if neumann else ("", "", "")
    cells_line = f"  set Cells per axis = {', '.join(str(c) for c in cells)}\n" if cells else ""
    return f"""
subsection Mesh
  set Dimensions               = {dim}
  set Domain size              = {size}
  set Initial refinement level = {refine}
end
subsection In situ
  set Displacement boundary labels     = {labels}
  set Displacement boundary components = {comps}
  set Displacement boundary values     = {vals}
  set Initial pressure                 = 10e6
  set Stress boundary components       = {nc}
  set Stress boundary labels           = {nl}
  set Stress boundary values           = {nv}
end
subsection Properties
  set Young modulus         = 1.4e10
  set Biot coefficient      = 0.9
  set Bulk density          = 2700
  set Fluid compressibility = 5.8e-10
  set Permeability          = 10
  set Poisson ratio         = 0.3
  set Porosity              = 0.3
  set Viscosity             = 1e-3
  set Well radius           = 1
  set Flow rate             = 1e-5
end
subsection Solver
  set Time step  = 60
  set Time max   = 1e3
end
subsection GPU
  set Displacement FE degree = {degree_u}
{cells_line}{extra_gpu}
end
"""
