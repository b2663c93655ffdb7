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
        # no-op when up to date; rebuilds (-march=native) when the library was compiled on another host CPU (oracle/Makefile)
        subprocess.check_call(["make", "-s", "-C", str(ROOT / "oracle")], stdout=subprocess.DEVNULL)
        lib = C.CDLL(str(path))
        capi._declare_operator_api(lib, "po_")
        lib.po_create.argtypes = [C.POINTER(C.c_void_p)]
        lib.po_set_threads.argtypes = [C.c_int]
        lib.po_set_preconditioner.argtypes = [C.c_void_p, C.c_int]
        lib.po_cg_csr.argtypes = [C.c_int64, C.POINTER(C.c_int64), C.POINTER(C.c_int32), C.POINTER(C.c_double), C.POINTER(C.c_double),
                                  C.POINTER(C.c_double), C.c_double, C.c_int, C.c_double, C.POINTER(C.c_int), C.POINTER(C.c_double)]
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


SHIPPED_INPUT = pkg.inputs.SHIPPED_INPUT
make_input = pkg.inputs.make_input


def oracle_cg(A, b, x0=None, omega=-1.0, max_steps=1000, tol=1e-12):
    """SolverCG of the oracle (cg_solve + ssor_apply in oracle.cpp) on a scipy matrix; omega > 0 SSOR, 0 Jacobi, < 0 identity."""
    lib = load_oracle()
    A = A.tocsr()
    A.sort_indices()
    n = A.shape[0]
    rp = np.ascontiguousarray(A.indptr, dtype=np.int64)
    col = np.ascontiguousarray(A.indices, dtype=np.int32)
    val = np.ascontiguousarray(A.data, dtype=np.float64)
    x = np.zeros(n) if x0 is None else np.ascontiguousarray(x0, dtype=np.float64).copy()
    bb = np.ascontiguousarray(b, dtype=np.float64)
    its, res = C.c_int(), C.c_double()
    P = lambda a, t: a.ctypes.data_as(C.POINTER(t))
    rc = lib.po_cg_csr(n, P(rp, C.c_int64), P(col, C.c_int32), P(val, C.c_double), P(x, C.c_double), P(bb, C.c_double), omega, max_steps, tol,
                       C.byref(its), C.byref(res))
    return rc, x, its.value, res.value
