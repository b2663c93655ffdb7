"""Manufactured-solution convergence rates of the assembled operators — a pin that does not come from reading deal.II.

The reference ships no golden vectors, so the oracle's operators are pinned here by ANALYSIS instead: for a
smooth u* that vanishes on the boundary, the Galerkin solution of  A u_h = F,  F_i = int f* . phi_i,  f* = -div sigma(u*)
(sigma = lambda tr(eps) I + 2 mu eps: constitutive_model CM:9-57 as used by DS:237-242) must converge to u* with order
k+1 in L2 for FE_Q(k) (k = 1: h^2, k = 2: h^3).  A wrong Lame weighting, a wrong component coupling, a wrong quadrature or a
wrong Dirichlet elimination makes the error stagnate.  Likewise  J p_h = F  with  J = M/(M_b dt) + (k/mu) K  (PS:158-169) and
p* = prod sin(pi x_a / L) (natural boundary conditions hold exactly) must converge with h^2.

Only the MATRICES come from the code under test (oracle: `-m "not gpu"`; CUDA library: `-m gpu`); f* is derived with sympy,
the load vector is integrated here with 4-point Gauss rules and Lagrange bases inferred from the dofs' support points (no
assumption on the local numbering), and the systems are solved by scipy's sparse direct solver."""
import numpy as np
import pytest
import scipy.sparse.linalg as spla
import sympy as sp

import helpers as H

capi, fss = H.capi, H.fss
L = 10.0


def lagrange_1d(degree, node, t):
    """value at t of the 1-D Lagrange polynomial of FE_Q(degree) that is 1 at reference coordinate `node`"""
    nodes = np.linspace(0.0, 1.0, degree + 1)
    out = np.ones(np.broadcast(node, t).shape)
    for xn in nodes:
        away = np.abs(node - xn) > 1e-9
        out = out * np.where(away, (t - xn) / np.where(away, node - xn, 1.0), 1.0)
    return out


def load_vector(mesh, dofs, n_comp, degree, f_of_x):
    """F_i = int f . phi_i with a 4-point Gauss rule per axis; f_of_x(x) -> (n_points, n_comp)"""
    m = mesh.arrays
    dim = m.dim
    cv = m.xyz[m.cell_vertices]                      # cells x vpc x dim
    x0, x1 = cv.min(axis=1), cv.max(axis=1)          # axis-aligned cells of the box meshes
    h = x1 - x0
    spts = dofs.support_points()
    cd = dofs.cell_dofs                              # cells x n_loc
    xi_dof = (spts[cd] - x0[:, None, :]) / h[:, None, :]   # reference coordinates of every local dof
    comp = np.arange(cd.shape[1]) % n_comp           # FESystem of identical bases: local index = scalar * n_comp + component
    g, w = np.polynomial.legendre.leggauss(4)
    g, w = 0.5 * (g + 1.0), 0.5 * w
    grids = np.meshgrid(*([g] * dim), indexing="ij")
    xq = np.stack([a.ravel() for a in grids], axis=1)                   # nq x dim
    wq = np.prod(np.stack(np.meshgrid(*([w] * dim), indexing="ij")), axis=0).ravel()
    F = np.zeros(dofs.n_dofs)
    vol = np.prod(h, axis=1)
    for q in range(xq.shape[0]):
        x = x0 + h * xq[q]                           # cells x dim
        fx = f_of_x(x)                               # cells x n_comp
        N = np.ones(cd.shape)
        for a in range(dim):
            N = N * lagrange_1d(degree, xi_dof[:, :, a], xq[q, a])
        np.add.at(F, cd, N * fx[:, comp] * (wq[q] * vol)[:, None])
    return F


def elasticity_case(dim):
    """u*, f* = -div sigma(u*) as numpy callables; u* = 0 on the boundary of [-L/2, L/2]^dim"""
    X = sp.symbols("x0:%d" % dim)
    lam, mu = sp.symbols("lam mu")
    freq = [[1, 1, 3], [3, 1, 1], [1, 3, 1]]
    amp = [1.0e-3, -0.5e-3, 0.7e-3]
    u = [amp[c] * sp.prod([sp.cos(freq[c][a] * sp.pi * X[a] / L) for a in range(dim)]) for c in range(dim)]
    eps = [[(sp.diff(u[i], X[j]) + sp.diff(u[j], X[i])) / 2 for j in range(dim)] for i in range(dim)]
    tr = sum(eps[i][i] for i in range(dim))
    sig = [[lam * tr * (1 if i == j else 0) + 2 * mu * eps[i][j] for j in range(dim)] for i in range(dim)]
    f = [-sum(sp.diff(sig[i][j], X[j]) for j in range(dim)) for i in range(dim)]
    uf = sp.lambdify(X, u, "numpy")
    ff = sp.lambdify((*X, lam, mu), f, "numpy")
    return (lambda x: np.stack(np.broadcast_arrays(*uf(*x.T)), axis=1)), (lambda x, l, m: np.stack(np.broadcast_arrays(*ff(*x.T, l, m)), axis=1))


def all_faces_clamped(dim):
    labels = [f for f in range(2 * dim) for _ in range(dim)]
    comps = [c for _ in range(2 * dim) for c in range(dim)]
    return labels, comps, [0.0] * len(labels)


def elasticity_errors(make_backend, dim, degree, levels):
    u_star, f_star = elasticity_case(dim)
    errs = []
    for refine in levels:
        inp = capi.InputData(text=H.make_input(dim=dim, refine=refine, degree_u=degree, dirichlet=all_faces_clamped(dim)))
        mesh = fss.make_mesh(inp)
        b = make_backend()
        try:
            dofs_p, dofs_u, (line_dof, _) = fss.upload_problem(b, inp, mesh)
            b.pressure_set_uniform(0.0)
            b.displacement_assemble()
            A = b.get_matrix(capi.MAT_ELASTICITY).tocsc()
            prm = inp.params()
            F = load_vector(mesh, dofs_u, dim, degree, lambda x: f_star(x, prm.lame_lambda, prm.shear_modulus))
            F[line_dof] = 0.0                      # clamped rows: identity-like rows of the eliminated matrix, u = 0
            uh = spla.spsolve(A[:, :A.shape[0]], F)
            spts = dofs_u.support_points()
            ue = u_star(spts)[np.arange(dofs_u.n_dofs), np.arange(dofs_u.n_dofs) % dim]
            errs.append(np.linalg.norm(uh - ue) / np.linalg.norm(ue))
        finally:
            b.close()
    return np.array(errs)


def pressure_errors(make_backend, dim, levels, dt=60.0):
    errs = []
    for refine in levels:
        inp = capi.InputData(text=H.make_input(dim=dim, refine=refine, degree_u=1))
        mesh = fss.make_mesh(inp)
        b = make_backend()
        try:
            dofs_p, _, _ = fss.upload_problem(b, inp, mesh)
            b.pressure_set_uniform(0.0)
            b.assemble_jacobian(dt)
            J = b.get_matrix(capi.MAT_JACOBIAN).tocsc()
            prm = inp.params()
            c0, kappa = 1.0 / (prm.m_modulus * dt), prm.perm_over_visc
            p_star = lambda x: np.prod(np.sin(np.pi * x / L), axis=1)
            coef = c0 + kappa * dim * (np.pi / L) ** 2            # c p* - kappa laplace(p*) = coef p*
            F = load_vector(mesh, dofs_p, 1, 1, lambda x: (coef * p_star(x))[:, None])
            ph = spla.spsolve(J[:, :J.shape[0]], F)
            pe = p_star(dofs_p.support_points())
            errs.append(np.linalg.norm(ph - pe) / np.linalg.norm(pe))
        finally:
            b.close()
    return np.array(errs)


def rates(errs):
    return np.log2(errs[:-1] / errs[1:])


CASES = [(2, 1, [3, 4, 5, 6]), (2, 2, [2, 3, 4, 5]), (3, 1, [2, 3, 4]), (3, 2, [2, 3])]


def check_elasticity(make_backend, dim, degree, levels):
    e = elasticity_errors(make_backend, dim, degree, levels)
    r = rates(e)
    # asymptotic order k+1 (nodal values of FE_Q(2) on uniform meshes are superconvergent: order 4 is fine, less than 3 is not)
    assert r[-1] >= degree + 1 - 0.25, (e, r)
    assert np.all(r >= degree + 1 - 0.6), (e, r)
    assert e[-1] < (2e-2 if degree == 1 else 2e-3), e
    return e, r


def check_pressure(make_backend, dim, levels):
    e = pressure_errors(make_backend, dim, levels)
    r = rates(e)
    assert np.all(np.abs(r - 2.0) <= 0.25), (e, r)
    return e, r


@pytest.mark.parametrize("dim,degree,levels", CASES)
def test_oracle_elasticity_converges_with_order_k_plus_1(dim, degree, levels):
    check_elasticity(H.create_oracle_backend, dim, degree, levels)


@pytest.mark.parametrize("dim,levels", [(2, [3, 4, 5, 6]), (3, [2, 3, 4])])
def test_oracle_pressure_operator_converges_with_order_2(dim, levels):
    check_pressure(H.create_oracle_backend, dim, levels)


@pytest.mark.gpu
@pytest.mark.parametrize("dim,degree,levels", CASES)
def test_gpu_elasticity_converges_with_order_k_plus_1(dim, degree, levels):
    e_gpu, _ = check_elasticity(lambda: capi.create_device_backend(0), dim, degree, levels)
    e_ora = elasticity_errors(H.create_oracle_backend, dim, degree, levels[-1:])
    assert e_gpu[-1] == pytest.approx(e_ora[-1], rel=1e-6)   # the same discrete solution, hence the same error


@pytest.mark.gpu
@pytest.mark.parametrize("dim,levels", [(2, [3, 4, 5, 6]), (3, [2, 3, 4])])
def test_gpu_pressure_operator_converges_with_order_2(dim, levels):
    check_pressure(lambda: capi.create_device_backend(0), dim, levels)
