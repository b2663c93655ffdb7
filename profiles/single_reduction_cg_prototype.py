"""Numerics prototype (CPU, numpy/scipy on the oracle's assembled elasticity matrix): Jacobi-preconditioned CG in the
standard three-phase form of dealii::SolverCG against the single-reduction (Chronopoulos-Gear) recurrence, which needs one
global sum per iteration instead of two (candidates for the 8-GPU step).

    python profiles/single_reduction_cg_prototype.py <refine>

Both stop on the recursively updated residual norm <= 1e-12 (DS:298-299).  Printed: iterations, TRUE residual ||b - A x|| at
exit, and the distance between the two solutions relative to ||x||."""
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "tests"))
import helpers as H  # noqa: E402

capi, fss = H.capi, H.fss
refine = int(sys.argv[1]) if len(sys.argv) > 1 else 5
H.load_oracle().po_set_threads(8)
inp = capi.InputData(text=H.make_input(dim=3, refine=refine, degree_u=1))
mesh = fss.make_mesh(inp)
b = H.create_oracle_backend()
fss.upload_problem(b, inp, mesh)
b.pressure_set_uniform(inp.p_init)
b.displacement_assemble()
A = b.get_matrix(capi.MAT_ELASTICITY).tocsr()
rhs = b.get_vector(capi.VEC_U_RHS)
n = A.shape[0]
Dinv = 1.0 / A.diagonal()
tol, maxit = 1e-12, 20000


def standard():
    x = np.zeros(n); g = -rhs.copy(); h = Dinv * g; d = -h; gh = g @ h; it = 0
    while True:
        it += 1
        h = A @ d; alpha = gh / (d @ h); g += alpha * h; x += alpha * d
        if np.linalg.norm(g) <= tol or it >= maxit:
            break
        h = Dinv * g; beta = gh; gh = g @ h; beta = gh / beta; d = beta * d - h
    return x, it


def chronopoulos_gear():
    # r = b - A x ; u = M^-1 r ; w = A u ; one reduction: gamma = r.u, delta = w.u, rho = r.r
    x = np.zeros(n); r = rhs.copy(); p = np.zeros(n); s = np.zeros(n)
    gamma_old = alpha_old = 1.0
    it = 0
    while True:
        u = Dinv * r
        w = A @ u
        gamma, delta, rho = r @ u, w @ u, r @ r     # ONE global sum of three numbers
        if np.sqrt(rho) <= tol or it >= maxit:
            break
        it += 1
        if it == 1:
            beta, alpha = 0.0, gamma / delta
        else:
            beta = gamma / gamma_old
            alpha = gamma / (delta - beta * gamma / alpha_old)
        p = u + beta * p
        s = w + beta * s
        x += alpha * p
        r -= alpha * s
        gamma_old, alpha_old = gamma, alpha
    return x, it


xs, its = standard()
xc, itc = chronopoulos_gear()
print(f"n {n}: standard CG {its} iterations, true residual {np.linalg.norm(rhs - A @ xs):.2e}; "
      f"single-reduction CG {itc} iterations (+1 matrix pass for the final check), true residual {np.linalg.norm(rhs - A @ xc):.2e}; "
      f"|x_cg1 - x_std| / |x| = {np.linalg.norm(xc - xs) / np.linalg.norm(xs):.2e}; |b| = {np.linalg.norm(rhs):.2e}")
