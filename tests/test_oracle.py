"""Pins for the CPU oracle.  The reference ships no golden vectors and deal.II is not
available, so the oracle is pinned by closed forms, conservation properties, the patch test, the
contraction factor of the inner loop, and an independent numpy/scipy restatement (oracle/oracle_np.py)."""
import sys

import numpy as np
import pytest

import helpers as H

sys.path.insert(0, str(H.ROOT / "oracle"))
from oracle_np import PoroNP  # noqa: E402

capi, fss = H.capi, H.fss


def oracle_problem(**kw):
    inp = capi.InputData(text=H.make_input(**kw))
    mesh = fss.make_mesh(inp)
    b = H.create_oracle_backend()
    dp, du, cons = fss.upload_problem(b, inp, mesh)
    return inp, mesh, b, dp, du, cons


def test_t3_tensor_index_maps():
    # TI:25-30 and FSS:100-110
    assert fss.TENSOR_TO_ENTRY[2] == [0, 1, 1, 2] and fss.TENSOR_TO_ENTRY[3] == [0, 1, 2, 1, 3, 4, 2, 4, 5]
    assert [fss.TENSOR_TO_ENTRY[2][c] for c in fss.VOLUMETRIC_COMPONENTS[2]] == [0, 2]
    assert [fss.TENSOR_TO_ENTRY[3][c] for c in fss.VOLUMETRIC_COMPONENTS[3]] == [0, 3, 5]
    assert [fss.TENSOR_TO_ENTRY[3][c] for c in fss.SHEAR_COMPONENTS[3]] == [1, 2, 4]


def test_t4_well_source():
    # RHS:106-110 with r = 1, q = 1e-5: f = -q / (3.1415926 r^2) inside the well, 0 outside
    inp, mesh, b, dp, du, _ = oracle_problem(dim=2, refine=5, degree_u=1)
    f = b.get_vector(capi.VEC_WELL_RHS)
    sp = dp.support_points()
    far = np.hypot(sp[:, 0], sp[:, 1]) > 1.0 + 10 / 32 * np.sqrt(2)
    assert np.all(f[far] == 0.0) and np.all(f <= 0)
    # cells entirely inside the well integrate the constant exactly: h^2 * f0 per cell
    f0 = -1e-5 / (3.1415926 * 1.0)
    assert f0 == pytest.approx(-3.1830989161357205e-06, rel=1e-15)
    centre = np.argmin(np.hypot(sp[:, 0], sp[:, 1]))
    assert f[centre] == pytest.approx(f0 * (10 / 32) ** 2, rel=1e-12)  # 4 cells x h^2/4 each
    b.close()


def test_t6_element_matrices_closed_form():
    # one-cell-wide strips are awkward with refine>=2; use the assembled 2x2 / 2x2x2 meshes and closed forms of
    # the interior node instead: M_ii = (4/9) h^2 * 4 cells /4 ..., verified through global identities below and
    # entrywise against the closed-form bilinear element matrices via the numpy restatement on a 1-cell patch.
    h = 2.5
    inp, mesh, b, dp, du, _ = oracle_problem(dim=2, refine=2, degree_u=1)
    M, K = b.get_matrix(capi.MAT_MASS), b.get_matrix(capi.MAT_LAPLACE)
    # closed forms for bilinear squares: M_e = h^2/36 [[4,2,2,1],...], K_e diag 2/3, edge -1/6, diagonal -1/3
    cd = dp.cell_dofs[0]
    corner = cd[0]  # vertex (-5,-5) belongs to one cell only
    assert M[corner, corner] == pytest.approx(4 * h * h / 36, rel=1e-14)
    assert M[corner, cd[1]] == pytest.approx(2 * h * h / 36, rel=1e-14)
    assert M[corner, cd[3]] == pytest.approx(1 * h * h / 36, rel=1e-14)
    assert K[corner, corner] == pytest.approx(2 / 3, rel=1e-14)
    assert K[corner, cd[1]] == pytest.approx(-1 / 6, rel=1e-13)
    assert K[corner, cd[3]] == pytest.approx(-1 / 3, rel=1e-14)
    b.close()
    inp, mesh, b, dp, du, _ = oracle_problem(dim=3, refine=2, degree_u=1)
    M, K = b.get_matrix(capi.MAT_MASS), b.get_matrix(capi.MAT_LAPLACE)
    cd = dp.cell_dofs[0]
    assert M[cd[0], cd[0]] == pytest.approx(8 * h ** 3 / 216, rel=1e-14)   # trilinear: h^3/216 * 8 on the diagonal
    assert M[cd[0], cd[7]] == pytest.approx(1 * h ** 3 / 216, rel=1e-14)
    assert K[cd[0], cd[0]] == pytest.approx(h / 3, rel=1e-14)
    assert K[cd[0], cd[7]] == pytest.approx(-h / 12, rel=1e-14)
    assert abs(K[cd[0], cd[1]]) <= 1e-15 * h  # edge neighbours cancel for trilinear cubes
    b.close()


@pytest.mark.parametrize("dim,refine,deg", [(2, 4, 2), (2, 4, 1), (3, 3, 1)])
def test_t5_t8_patch_test_and_matrix_properties(dim, refine, deg):
    inp, mesh, b, dp, du, (line_dof, g) = oracle_problem(dim=dim, refine=refine, degree_u=deg)
    fss.initialize(b, inp)
    M, K, A = b.get_matrix(capi.MAT_MASS), b.get_matrix(capi.MAT_LAPLACE), b.get_matrix(capi.MAT_ELASTICITY)
    assert M.sum() == pytest.approx(10.0 ** dim, rel=1e-12)
    assert abs(K.sum(axis=1)).max() <= 1e-12 * abs(K).max()
    for X in (M, K, A):
        assert abs(X - X.T).max() <= 1e-14 * abs(X).max()
    # A is SPD after elimination: constrained rows are diagonal-only with positive diagonal
    Ad = A.toarray() if A.shape[0] < 3000 else None
    if Ad is not None:
        assert np.linalg.eigvalsh(Ad).min() > 0
        off = Ad[line_dof].copy()
        off[np.arange(len(line_dof)), line_dof] = 0
        assert np.all(off == 0) and np.all(Ad[line_dof, line_dof] > 0)
    # exact solution of the initial state is linear: u_a = -1e-5 (x_a + 5)/10 -> strains -1e-6 (T5)
    u = b.get_vector(capi.VEC_U)
    sp = du.support_points()
    comp = np.zeros(du.n_dofs, int)
    for c in range(dim):
        comp[du.cell_dofs[:, c::dim].ravel()] = c
    exact = -1e-5 * (sp[np.arange(du.n_dofs), comp] + 5.0) / 10.0
    assert np.abs(u - exact).max() <= 1e-15
    ev0 = b.get_vector(capi.VEC_VOL_STRAIN0)
    assert np.allclose(ev0, -1e-6 * dim, rtol=2e-7)
    for c in fss.VOLUMETRIC_COMPONENTS[dim]:
        assert np.allclose(b.get_vector(capi.VEC_STRAIN0 + fss.TENSOR_TO_ENTRY[dim][c]), -1e-6, rtol=2e-7)
    b.close()


def test_t7_inner_loop_contraction_and_as_is_control_flow():
    inp, mesh, b, *_ = oracle_problem(dim=2, refine=4, degree_u=2)
    fss.initialize(b, inp)
    prm = inp.params()
    rho = prm.biot_coef ** 2 * prm.m_modulus / prm.bulk_modulus
    assert rho == pytest.approx(0.38756, abs(1e-5))
    for _ in range(3):
        rep = fss.time_step(b, inp)
        h = np.array(rep["residual_history"])
        assert rep["fss_iterations"] == 1                 # FSS:399 commented out => one coupling iteration
        assert h[-1] < 1e-8 and np.all(h[:-1] >= 1e-8)
        assert rep["pressure_error"] == pytest.approx(h[-1], rel=1e-12)  # post-mechanics "Error" == last residual
        r = h[2:] / h[1:-1]
        assert np.all(r <= rho * 1.001) and np.all(r > 0.25)
    b.close()


def key(x):
    return tuple(np.round(np.asarray(x) * 1e6).astype(np.int64))


def perm_by_coords(src, dst):
    d = {key(x): i for i, x in enumerate(src)}
    return np.array([d[key(x)] for x in dst])


@pytest.mark.parametrize("dim,n,deg", [(2, 8, 1), (2, 4, 2), (3, 4, 1), (3, 2, 2)])
def test_t9_independent_numpy_restatement(dim, n, deg):
    inp = capi.InputData(text=H.make_input(dim=dim, refine=2, degree_u=deg, cells=[n] * dim))
    mesh = fss.make_mesh(inp)
    b = H.create_oracle_backend()
    dp, du, _ = fss.upload_problem(b, inp, mesh)
    prm = inp.params()
    P = {k: getattr(prm, k) for k in ("lame_lambda", "shear_modulus", "bulk_modulus", "biot_coef", "m_modulus", "perm_over_visc", "well_radius", "flow_rate")}
    R = PoroNP(dim, [10] * dim, [n] * dim, deg, P)
    R.assemble_displacement([(int(l), int(c), float(v)) for l, c, v in zip(inp.displacement_boundary_labels, inp.displacement_boundary_components,
                                                                         inp.displacement_boundary_values)])
    pp = perm_by_coords(R.xp, dp.support_points())
    comp = np.zeros(du.n_dofs, int)
    for c in range(dim):
        comp[du.cell_dofs[:, c::dim].ravel()] = c
    pu = perm_by_coords(R.xu, du.support_points()) * dim + comp
    fss.initialize(b, inp)
    R.initialize(inp.p_init)
    rel = lambda X, Y: abs(X - Y).max() / abs(Y).max()
    assert rel(b.get_matrix(capi.MAT_MASS), R.M[pp][:, pp]) <= 1e-14
    assert rel(b.get_matrix(capi.MAT_LAPLACE), R.K[pp][:, pp]) <= 1e-14
    assert rel(b.get_matrix(capi.MAT_ELASTICITY), R.eliminated_matrix()[pu][:, pu]) <= 1e-14
    assert np.abs(b.get_vector(capi.VEC_WELL_RHS) - R.f[pp]).max() <= 1e-14 * np.abs(R.f).max()
    assert fss.rel_l2(b.get_vector(capi.VEC_U_RHS), R.rhs_displacement(R.p)[pu]) <= 1e-12
    assert fss.rel_l2(b.get_vector(capi.VEC_U), R.u[pu]) <= 1e-10
    assert fss.rel_l2(b.get_vector(capi.VEC_VOL_STRAIN0), R.ev0[pp]) <= 1e-6  # projection CG stops at 1e-8 relative residual
    # projection right-hand sides (SP:159-196) entry by entry, before any solve tolerance enters
    vol = fss.VOLUMETRIC_COMPONENTS[dim]
    _, rhs_np = R.project_strains(R.u, vol)
    for comp_t in vol:
        got = b.get_vector(capi.VEC_PROJ_RHS0 + fss.TENSOR_TO_ENTRY[dim][comp_t])
        assert np.abs(got - rhs_np[comp_t][pp]).max() <= 1e-9 * np.abs(rhs_np[comp_t]).max()
    for _ in range(2):
        rep = fss.time_step(b, inp)
        hist = R.time_step(inp.time_step)
        assert rep["inner_counts"] == [len(hist)]
        assert fss.rel_l2(b.get_vector(capi.VEC_P), R.p[pp]) <= 1e-10
        assert fss.rel_l2(b.get_vector(capi.VEC_U), R.u[pu]) <= 1e-9
    b.close()


def test_cg_and_ssor_follow_dealii_algorithms():
    """SolverCG / PreconditionSSOR restated in plain numpy on the assembled Jacobian: identical iteration count and iterate."""
    inp, mesh, b, *_ = oracle_problem(dim=2, refine=3, degree_u=1)
    fss.initialize(b, inp)
    b.pressure_begin_step(); b.pressure_zero_update(); b.update_volumetric_strain()
    b.assemble_residual(inp.time_step); b.assemble_jacobian(inp.time_step)
    J = b.get_matrix(capi.MAT_JACOBIAN).toarray()
    r = b.get_vector(capi.VEC_P_RESIDUAL)
    its, res = b.pressure_solve()
    x_oracle = b.get_vector(capi.VEC_P_UPDATE)
    n = len(r)
    D = np.diag(J).copy()

    def ssor(src, om=1.0):
        dst = np.zeros(n)
        for i in range(n):
            dst[i] = (src[i] - om * (J[i, :i] @ dst[:i])) / D[i]
        dst *= om * (2 - om) * D
        for i in range(n - 1, -1, -1):
            dst[i] = (dst[i] - om * (J[i, i + 1:] @ dst[i + 1:])) / D[i]
        return dst

    tol = 1e-8 * np.linalg.norm(r)
    x = np.zeros(n)
    g = -r.copy()
    h = ssor(g)
    d = -h
    gh = g @ h
    it = 0
    while True:
        it += 1
        h = J @ d
        alpha = gh / (d @ h)
        g = g + alpha * h
        x = x + alpha * d
        if np.linalg.norm(g) <= tol:
            break
        h = ssor(g)
        beta = gh
        gh = g @ h
        beta = gh / beta
        d = beta * d - h
    assert it == its
    assert np.linalg.norm(x - x_oracle) <= 1e-9 * np.linalg.norm(x)
    b.close()


def test_neumann_face_term_closed_form():
    # uniform traction t on the top face of the square, component y: sum of rhs over the top dofs = t * n_y * length
    labels = [0, 1, 2]
    text = H.make_input(dim=2, refine=2, degree_u=1, dirichlet=(labels, [0, 0, 1], [0.0] * 3), neumann=([3], [1], [-1e6]))
    inp = capi.InputData(text=text)
    mesh = fss.make_mesh(inp)
    b = H.create_oracle_backend()
    dp, du, _ = fss.upload_problem(b, inp, mesh)
    b.pressure_set_uniform(0.0)
    b.displacement_assemble()
    rhs = b.get_vector(capi.VEC_U_RHS)
    assert rhs.sum() == pytest.approx(-1e6 * 10.0, rel=1e-12)
    sp = du.support_points()
    assert np.all(rhs[sp[:, 1] < 5 - 1e-9] == 0)
    b.close()


@pytest.mark.parametrize("dim,fname,deg", [(2, "distorted_quad8.msh", 1), (2, "distorted_quad8.msh", 2), (3, "distorted_hex4.msh", 1), (3, "distorted_hex4.msh", 2)])
def test_patch_test_on_distorted_gmsh_meshes(dim, fname, deg):
    """Non-affine cells (interior nodes moved by up to 20 % of h): the isoparametric spaces still contain the linear
    field u_a = -1e-5 (x_a + 5)/10 and a uniform pressure loads only constrained rows, so the initial state must be
    that field exactly — a check of the general J^-T grad(N), det J quadrature and of the Gmsh quad / hex readers."""
    inp = capi.InputData(text=H.make_input(dim=dim, refine=2, degree_u=deg))
    mesh = capi.mesh_read_msh(H.ROOT / "tests" / "golden" / fname, dim)
    vol = 0.0
    b = H.create_oracle_backend()
    dp, du, _ = fss.upload_problem(b, inp, mesh)
    fss.initialize(b, inp)
    M = b.get_matrix(capi.MAT_MASS)
    assert M.sum() == pytest.approx(10.0 ** dim, rel=1e-12)  # sum of det J * w over distorted cells
    u = b.get_vector(capi.VEC_U)
    sp = du.support_points()
    comp = np.zeros(du.n_dofs, int)
    for c in range(dim):
        comp[du.cell_dofs[:, c::dim].ravel()] = c
    exact = -1e-5 * (sp[np.arange(du.n_dofs), comp] + 5.0) / 10.0
    assert np.abs(u - exact).max() <= 1e-14
    assert np.allclose(b.get_vector(capi.VEC_VOL_STRAIN0), -1e-6 * dim, rtol=1e-6)
    rep = fss.time_step(b, inp)
    assert rep["fss_iterations"] == 1
    b.close()


@pytest.mark.parametrize("dim,deg", [(2, 1), (2, 2), (3, 1)])
def test_t6_elasticity_element_matrix_against_sympy(dim, deg):
    """the single-cell elasticity matrix a(phi_i, phi_j) = int (C : eps(phi_i)) : eps(phi_j) (DS:237-242 with
    CM:45-57) integrated EXACTLY with sympy from the Lagrange basis definition, against the oracle's assembled matrix
    (no Dirichlet data, so nothing is eliminated).  Pins FE_Q(deg), the local dof order, QGauss(deg+1) and the tensor
    contraction independently of any quadrature code."""
    import sympy as sp
    size = 10
    inp = capi.InputData(text=H.make_input(dim=dim, refine=2, degree_u=deg, cells=[1] * dim, dirichlet=([], [], [])))
    mesh = fss.make_mesh(inp)
    b = H.create_oracle_backend()
    dp, du, (line_dof, _) = fss.upload_problem(b, inp, mesh)
    assert len(line_dof) == 0
    b.pressure_set_uniform(0.0)
    b.displacement_assemble()
    A = b.get_matrix(capi.MAT_ELASTICITY).toarray()
    prm = inp.params()
    lam, mu = sp.nsimplify(prm.lame_lambda), sp.nsimplify(prm.shear_modulus)
    xs = sp.symbols("x0:%d" % dim)
    nodes1d = [sp.Rational(k, deg) for k in range(deg + 1)]

    def lagr(i, x):
        e = sp.Integer(1)
        for k, t in enumerate(nodes1d):
            if k != i:
                e *= (x - t) / (nodes1d[i] - t)
        return e

    spts = du.support_points()
    comp = np.zeros(du.n_dofs, int)
    for c in range(dim):
        comp[du.cell_dofs[:, c::dim].ravel()] = c
    shapes = []
    for d in range(du.n_dofs):  # physical cell [-5,5]^dim: unit coordinate xi = (x + 5)/10
        xi = [(spts[d][a] + size / 2) / size for a in range(dim)]
        idx = [int(round(v * deg)) for v in xi]
        N = sp.Integer(1)
        for a in range(dim):
            N *= lagr(idx[a], (xs[a] + sp.Rational(size, 2)) / size)
        shapes.append((N, int(comp[d])))

    def strain(N, c):
        g = [sp.diff(N, x) for x in xs]
        E = sp.zeros(dim, dim)
        for a in range(dim):
            E[c, a] += g[a] / 2
            E[a, c] += g[a] / 2
        return E

    eps = [strain(N, c) for N, c in shapes]
    lims = [(x, -sp.Rational(size, 2), sp.Rational(size, 2)) for x in xs]
    worst = 0.0
    n = du.n_dofs
    pairs = [(i, j) for i in range(n) for j in range(i, n)]
    if n > 20:  # keep the symbolic work bounded: a deterministic subset of the upper triangle
        pairs = pairs[::7]
    for i, j in pairs:
        sig = lam * eps[i].trace() * sp.eye(dim) + 2 * mu * eps[i]
        integrand = sum(sig[a, c] * eps[j][a, c] for a in range(dim) for c in range(dim))
        exact = float(sp.integrate(sp.expand(integrand), *lims))
        scale = max(abs(exact), abs(A).max() * 1e-3)
        worst = max(worst, abs(A[i, j] - exact) / scale, abs(A[j, i] - exact) / scale)
    assert worst <= 1e-12, worst
    b.close()


@pytest.mark.parametrize("dim,n_cells,n_dofs,cg_iterations", [(2, 256, 289, 26), (3, 4096, 4913, 30)])
def test_t10_dealii_step4_published_iteration_counts(dim, n_cells, n_dofs, cg_iterations):
    """EXTERNAL PIN.  The "Results" section of deal.II's tutorial step-4 prints, for -Laplace u = 4 sum_a x_a^4 on [-1,1]^dim
    (hyper_cube + refine_global(4), FE_Q(1), QGauss(2), u = |x|^2 on the boundary through interpolate_boundary_values +
    MatrixTools::apply_boundary_values, SolverCG with PreconditionIdentity and SolverControl(1000, 1e-12)):

        2D: 256 active cells, 289 degrees of freedom, "26 CG iterations needed to obtain convergence."
        3D: 4096 active cells, 4913 degrees of freedom, "30 CG iterations needed to obtain convergence."

    (quoted from the published tutorial output; deal.II itself cannot be installed here).  Reproducing both numbers pins
    the oracle's refine_global mesh, its create_laplace_matrix restatement (PS:99-101), the QGauss(2) tables, and the
    SolverCG recurrence / SolverControl stopping rule used at PS:175-179, DS:299-305, SP:209-214."""
    import itertools
    inp = capi.InputData(text=H.make_input(dim=dim, refine=4, degree_u=1))
    mesh = capi.mesh_rectangle(dim, [2.0] * dim, 4)
    b = H.create_oracle_backend()
    dp, _, _ = fss.upload_problem(b, inp, mesh)
    assert mesh.arrays.n_cells == n_cells and dp.n_dofs == n_dofs
    K = b.get_matrix(capi.MAT_LAPLACE).tocsr()
    x = dp.support_points()
    # VectorTools::create_right_hand_side with QGauss(2) on the affine cells
    m = mesh.arrays
    a = 0.5 / np.sqrt(3.0)
    g1 = np.array([0.5 - a, 0.5 + a])
    f = np.zeros(n_dofs)
    X = m.xyz[m.cell_vertices]
    lo, h = X.min(axis=1), X.max(axis=1) - X.min(axis=1)
    for q in itertools.product(range(2), repeat=dim):
        xi = g1[list(q)]
        pts = lo + xi * h
        fq = 4.0 * (pts ** 4).sum(axis=1) * np.prod(h, axis=1) / 2 ** dim
        for k in range(1 << dim):
            N = np.prod([xi[d] if (k >> d) & 1 else 1 - xi[d] for d in range(dim)])
            np.add.at(f, dp.cell_dofs[:, k], fq * N)
    # MatrixTools::apply_boundary_values: symmetric elimination, rhs_i = a_ii g_i, solution_i = g_i
    bnd = np.isclose(np.abs(x).max(axis=1), 1.0)
    g = np.where(bnd, (x ** 2).sum(axis=1), 0.0)
    rhs = f - K @ g
    D = K.diagonal()
    keep = sp_diag(~bnd)
    A = keep @ K @ keep + sp_diag(bnd) @ sp_diag(D)
    rhs[bnd] = D[bnd] * g[bnd]
    rc, u, its, res = H.oracle_cg(A, rhs, x0=g, omega=-1.0, max_steps=1000, tol=1e-12)
    assert rc == 0 and res <= 1e-12
    assert its == cg_iterations
    assert np.linalg.norm(A @ u - rhs) <= 1e-11 and np.array_equal(u[bnd], g[bnd])
    b.close()


def sp_diag(v):
    import scipy.sparse as sp
    return sp.diags(np.asarray(v, dtype=float)).tocsr()
