"""oracle_np.py — an INDEPENDENT numpy/scipy restatement of the reference's discrete problem.

TEST INFRASTRUCTURE ONLY (guards the C++ oracle against restatement bugs; deal.II itself
is not available — the reference's own sources run against an API shim, tests/test_reference_run.py, and this file
is the check that does not share deal.II's algorithms with either).  Deliberately written differently from
oracle/oracle.cpp and from the CUDA kernels:

* structured box meshes only, nodes keyed by their coordinates (no deal.II numbering at all);
* elasticity through the Voigt B-matrix / D-matrix form  K_e = sum_q B^T D B JxW  instead of the
  4th-order tensor contraction of DS:237-242 / CM:45-57;
* sparse direct solves (scipy ``spsolve``) instead of CG, so fields are "exact" discrete solutions.

Forms restated (paths relative to /root/reference/lib/include):
  mass / Laplace ............ PS:96-101
  well source ............... PS:142-147, right_hand_side.h:99-116 (pi = 3.1415926)
  residual / Jacobian ....... PS:113-169
  elasticity + coupling rhs . DS:216-246, Dirichlet elimination of DS:279-286
  strain projection ......... SP:109-232
  time step ................. PoroelasticityFSS.h:328-407 (as-is: FSS:399 commented out)
"""
from __future__ import annotations

import itertools

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla


def gauss(n):
    if n == 2:
        a = 0.5 / np.sqrt(3.0)
        return np.array([0.5 - a, 0.5 + a]), np.array([0.5, 0.5])
    a = 0.5 * np.sqrt(0.6)
    return np.array([0.5 - a, 0.5, 0.5 + a]), np.array([5 / 18, 8 / 18, 5 / 18])


def lagrange_basis(degree, x):
    """values and derivatives of the 1D Lagrange basis on equidistant nodes of [0,1]; shape (degree+1,)"""
    nodes = np.linspace(0, 1, degree + 1)
    v = np.ones(degree + 1)
    d = np.zeros(degree + 1)
    for i in range(degree + 1):
        others = [nodes[j] for j in range(degree + 1) if j != i]
        v[i] = np.prod([(x - t) / (nodes[i] - t) for t in others])
        d[i] = sum(np.prod([(x - t) / (nodes[i] - t) for t in others if t is not s]) / (nodes[i] - s) for s in others)
    return v, d


class Box:
    """n[a] cells per axis on [-L/2, L/2]^dim; scalar Lagrange nodes of a given degree keyed by integer lattice index."""

    def __init__(self, dim, size, n):
        self.dim, self.size, self.n = dim, np.asarray(size, float), np.asarray(n, int)
        self.h = self.size / self.n

    def cells(self):
        return itertools.product(*[range(k) for k in self.n[::-1]])  # z,y,x order; order is irrelevant here

    def lattice(self, degree):
        return self.n * degree + 1

    def node_id(self, degree, idx):
        L = self.lattice(degree)
        out = 0
        for a in reversed(range(self.dim)):
            out = out * L[a] + idx[a]
        return out

    def n_nodes(self, degree):
        return int(np.prod(self.lattice(degree)))

    def node_coords(self, degree):
        L = self.lattice(degree)
        grids = np.meshgrid(*[np.arange(L[a]) for a in range(self.dim)], indexing="ij")
        ids = sum(grids[a] * int(np.prod(L[:a])) for a in range(self.dim))
        xyz = np.zeros((self.n_nodes(degree), self.dim))
        for a in range(self.dim):
            xyz[ids.ravel(), a] = (-0.5 * self.size[a] + grids[a] * self.h[a] / degree).ravel()
        return xyz

    def cell_nodes(self, degree, cell_xyz):
        """global node ids of the cell's (degree+1)^dim nodes, local order lexicographic (x fastest)"""
        loc = itertools.product(*[range(degree + 1) for _ in range(self.dim)][::-1])
        out = []
        for l in loc:
            l = l[::-1]
            out.append(self.node_id(degree, [cell_xyz[a] * degree + l[a] for a in range(self.dim)]))
        return np.array(out)


def shape_tables(dim, degree, nq1d):
    """N[q, s], dN[q, s, a] on the unit cell, local node order lexicographic, q-points x fastest"""
    x, w = gauss(nq1d)
    qs = list(itertools.product(*[range(nq1d)] * dim))
    qs = [q[::-1] for q in qs]
    ls = [l[::-1] for l in itertools.product(*[range(degree + 1)] * dim)]
    N = np.zeros((len(qs), len(ls)))
    dN = np.zeros((len(qs), len(ls), dim))
    W = np.zeros(len(qs))
    pts = np.zeros((len(qs), dim))
    for iq, q in enumerate(qs):
        vals = [lagrange_basis(degree, x[q[a]]) for a in range(dim)]
        W[iq] = np.prod([w[q[a]] for a in range(dim)])
        pts[iq] = [x[q[a]] for a in range(dim)]
        for il, l in enumerate(ls):
            N[iq, il] = np.prod([vals[a][0][l[a]] for a in range(dim)])
            for a in range(dim):
                dN[iq, il, a] = np.prod([vals[b][1][l[b]] if b == a else vals[b][0][l[b]] for b in range(dim)])
    return N, dN, W, pts


class PoroNP:
    def __init__(self, dim, size, n, degree_u, prm):
        """prm: dict with lame_lambda, shear_modulus, bulk_modulus, biot_coef, m_modulus, perm_over_visc, well_radius, flow_rate"""
        self.box = Box(dim, size, n)
        self.dim, self.degree_u, self.prm = dim, degree_u, prm
        self.np_ = self.box.n_nodes(1)
        self.nus = self.box.n_nodes(degree_u)
        self.nu = self.nus * dim
        self.xp = self.box.node_coords(1)
        self.xu = self.box.node_coords(degree_u)
        self._assemble_pressure()

    # ---- pressure: M, K, f (affine box cells: J = diag(h))
    def _assemble_pressure(self):
        dim, box = self.dim, self.box
        N, dN, W, pts = shape_tables(dim, 1, 2)
        detJ = np.prod(box.h)
        G = dN / box.h  # physical gradients
        Me = np.einsum("qi,qj,q->ij", N, N, W) * detJ
        Ke = np.einsum("qia,qja,q->ij", G, G, W) * detJ
        rows, cols, mv, kv = [], [], [], []
        f = np.zeros(self.np_)
        rw, rate = self.prm["well_radius"], self.prm["flow_rate"]
        for cell in box.cells():
            c = cell[::-1]
            ids = box.cell_nodes(1, c)
            rows.append(np.repeat(ids, len(ids)))
            cols.append(np.tile(ids, len(ids)))
            mv.append(Me.ravel())
            kv.append(Ke.ravel())
            x0 = -0.5 * box.size + np.array(c) * box.h
            xq = x0 + pts * box.h
            fq = np.where(xq[:, 0] ** 2 + xq[:, 1] ** 2 <= rw * rw, -rate / (3.1415926 * rw * rw), 0.0)
            f[ids] += (N * (fq * W)[:, None]).sum(axis=0) * detJ
        rows, cols = np.concatenate(rows), np.concatenate(cols)
        self.M = sp.csr_matrix((np.concatenate(mv), (rows, cols)), shape=(self.np_, self.np_))
        self.K = sp.csr_matrix((np.concatenate(kv), (rows, cols)), shape=(self.np_, self.np_))
        self.f = f

    # ---- displacement: A (Voigt), coupling matrix Gc with b = alpha * Gc p
    def assemble_displacement(self, dirichlet):
        """dirichlet: list of (axis_face_label, component, value) in the colorize convention"""
        dim, box, du = self.dim, self.box, self.degree_u
        lam, mu = self.prm["lame_lambda"], self.prm["shear_modulus"]
        Nu, dNu, W, pts = shape_tables(dim, du, du + 1)
        Np, _, _, _ = shape_tables(dim, 1, du + 1)  # pressure shape at the displacement q-points (DS:167-168)
        detJ = np.prod(box.h)
        Gu = dNu / box.h
        ns = Nu.shape[1]
        nv = 3 if dim == 2 else 6
        D = np.zeros((nv, nv))
        D[:dim, :dim] = lam
        D[np.arange(dim), np.arange(dim)] += 2 * mu
        D[np.arange(dim, nv), np.arange(dim, nv)] = mu  # engineering shear strains
        shear_pairs = [(0, 1)] if dim == 2 else [(0, 1), (0, 2), (1, 2)]
        Ae = np.zeros((ns * dim, ns * dim))
        Ce = np.zeros((ns * dim, Np.shape[1]))
        for q in range(len(W)):
            B = np.zeros((nv, ns * dim))
            for s in range(ns):
                for a in range(dim):
                    B[a, s * dim + a] = Gu[q, s, a]
                for k, (a, b) in enumerate(shear_pairs):
                    B[dim + k, s * dim + a] = Gu[q, s, b]
                    B[dim + k, s * dim + b] = Gu[q, s, a]
            Ae += B.T @ D @ B * W[q] * detJ
            div = B[:dim].sum(axis=0)  # trace of the strain of each shape function
            Ce += np.outer(div, Np[q]) * W[q] * detJ
        rows, cols, av, crow, ccol, cv = [], [], [], [], [], []
        for cell in box.cells():
            c = cell[::-1]
            su = box.cell_nodes(du, c)
            idu = (su[:, None] * dim + np.arange(dim)[None, :]).ravel()
            idp = box.cell_nodes(1, c)
            rows.append(np.repeat(idu, len(idu)))
            cols.append(np.tile(idu, len(idu)))
            av.append(Ae.ravel())
            crow.append(np.repeat(idu, len(idp)))
            ccol.append(np.tile(idp, len(idu)))
            cv.append(Ce.ravel())
        self.A_full = sp.csr_matrix((np.concatenate(av), (np.concatenate(rows), np.concatenate(cols))), shape=(self.nu, self.nu))
        self.Gc = sp.csr_matrix((np.concatenate(cv), (np.concatenate(crow), np.concatenate(ccol))), shape=(self.nu, self.np_))
        # Dirichlet: first condition in list order wins (DS:123-134)
        g = np.full(self.nu, np.nan)
        tol = 1e-9 * box.size.max()
        for label, comp, value in dirichlet:
            axis, side = label // 2, label % 2
            coord = -0.5 * box.size[axis] if side == 0 else 0.5 * box.size[axis]
            on = np.where(np.abs(self.xu[:, axis] - coord) < tol)[0]
            ids = on * dim + comp
            free = np.isnan(g[ids])
            g[ids[free]] = value
        self.cons = ~np.isnan(g)
        self.g = np.where(self.cons, g, 0.0)
        free = ~self.cons
        self.free = free
        self.A_ff = self.A_full[free][:, free].tocsc()
        self.A_fc = self.A_full[free][:, self.cons]
        self.lu = spla.splu(self.A_ff)

    def eliminated_matrix(self):
        """what distribute_local_to_global leaves in system_matrix: free-free block, zero couplings to constrained
        dofs, constrained diagonal = sum over cells of |a_ii^cell| (= the assembled diagonal here, all positive)."""
        A = self.A_full.tolil()
        c = np.where(self.cons)[0]
        d = self.A_full.diagonal()
        A[c, :] = 0
        A[:, c] = 0
        A = A.tocsr()
        A = A + sp.csr_matrix((d[c], (c, c)), shape=A.shape)
        A.eliminate_zeros()
        return A.tocsr()

    def solve_displacement(self, p):
        b = self.prm["biot_coef"] * (self.Gc @ p)
        rhs = b[self.free] - self.A_fc @ self.g[self.cons]
        u = self.g.copy()
        u[self.free] = self.lu.solve(rhs)
        return u

    def rhs_displacement(self, p):
        """rhs_vector as the reference leaves it: zero on constrained rows"""
        b = self.prm["biot_coef"] * (self.Gc @ p)
        out = np.zeros(self.nu)
        out[self.free] = b[self.free] - self.A_fc @ self.g[self.cons]
        return out

    # ---- strain projection (SP:109-232), exact mass solve
    def project_strains(self, u, comps):
        dim, box, du = self.dim, self.box, self.degree_u
        Np, _, W, _ = shape_tables(dim, 1, 2)
        _, dNu, _, _ = shape_tables(dim, du, 2)
        Gu = dNu / box.h
        detJ = np.prod(box.h)
        rhs = {c: np.zeros(self.np_) for c in comps}
        U = u.reshape(-1, dim)
        for cell in box.cells():
            c = cell[::-1]
            su = box.cell_nodes(du, c)
            idp = box.cell_nodes(1, c)
            grad = np.einsum("sc,qsa->qca", U[su], Gu)  # du_c/dx_a at q
            eps = 0.5 * (grad + grad.transpose(0, 2, 1))
            for comp in comps:
                i, j = comp // dim, comp % dim
                rhs[comp][idp] += (Np * (eps[:, i, j] * W)[:, None]).sum(axis=0) * detJ
        lu = spla.splu(self.M.tocsc())
        return {c: lu.solve(rhs[c]) for c in comps}, rhs

    # ---- one time step, as-is semantics (FSS:328-407)
    def initialize(self, p_init):
        self.p = np.full(self.np_, float(p_init))
        self.u = self.solve_displacement(self.p)
        vol = [0, 3] if self.dim == 2 else [0, 4, 8]
        strains, _ = self.project_strains(self.u, vol)
        self.ev = sum(strains[c] for c in vol)
        self.ev0 = self.ev.copy()

    def residual(self, dt, p_old):
        P = self.prm
        t1 = (self.ev - self.ev0) * (P["biot_coef"] / dt) + (self.p - p_old) * (1.0 / P["m_modulus"] / dt)
        return -(self.M @ t1 + P["perm_over_visc"] * (self.K @ self.p) + self.f)

    def time_step(self, dt, tol=1e-8, max_inner=50):
        P = self.prm
        p_old = self.p.copy()
        J = (self.M * (1.0 / P["m_modulus"] / dt) + P["perm_over_visc"] * self.K).tocsc()
        lu = spla.splu(J)
        dp = np.zeros(self.np_)
        hist = []
        for it in range(max_inner):
            self.ev = self.ev + (P["biot_coef"] / P["bulk_modulus"]) * dp
            r = self.residual(dt, p_old)
            hist.append(np.linalg.norm(r))
            if hist[-1] < tol:
                break
            dp = lu.solve(r)
            self.p = self.p + dp
        self.u = self.solve_displacement(self.p)
        return hist


# ======================================================================================================================
# Adaptive (hanging-node) meshes: an independent restatement of what deal.II's ConstraintMatrix does for the
# reference on refined meshes (PS:71-78, 153, 168, 180; DS:109-146, 279-286, 306; SP:104-105, 193-194, 215).
# Written differently from oracle.cpp and csrc/host/amr.hpp on purpose:
#   * cells are axis-aligned boxes given only by their corner coordinates; nodes are keyed by coordinates;
#   * a node is "hanging" when it lies in the closed box of a cell without being one of that cell's lattice nodes; its
#     weights are that cell's Lagrange basis functions evaluated at the node (geometry, no refinement tree);
#   * constraints are applied globally by sparse triple products  A^ = E^T A E,  b^ = E^T (f - A g)  and direct solves,
#     not cell by cell.
# ======================================================================================================================
class AdaptiveNP:
    def __init__(self, dim, cell_lo, cell_hi, degree_u, prm):
        self.dim, self.degree_u, self.prm = dim, degree_u, prm
        self.lo, self.hi = np.asarray(cell_lo, float), np.asarray(cell_hi, float)
        self.nc = len(self.lo)
        self.xp, self.cell_p = self._nodes(1)
        self.xu, self.cell_u = self._nodes(degree_u)
        self.np_, self.nus = len(self.xp), len(self.xu)
        self.nu = self.nus * dim
        self.lines_p = self._hanging_lines(1, self.xp, self.cell_p)
        self._assemble_pressure()

    @staticmethod
    def _key(x):
        return tuple(np.round(np.asarray(x) * 1e7).astype(np.int64))

    def _nodes(self, degree):
        ids, coords, cells = {}, [], []
        ls = [l[::-1] for l in itertools.product(*[range(degree + 1)] * self.dim)]
        for c in range(self.nc):
            row = []
            for l in ls:
                x = self.lo[c] + np.array(l) / degree * (self.hi[c] - self.lo[c])
                k = self._key(x)
                if k not in ids:
                    ids[k] = len(coords)
                    coords.append(x)
                row.append(ids[k])
            cells.append(row)
        return np.array(coords), np.array(cells)

    def _hanging_lines(self, degree, X, cell_nodes):
        """{node: {master: weight}} for scalar nodes, chains resolved."""
        ls = [l[::-1] for l in itertools.product(*[range(degree + 1)] * self.dim)]
        lines = {}
        eps = 1e-9
        for c in range(self.nc):
            inside = np.all((X >= self.lo[c] - eps) & (X <= self.hi[c] + eps), axis=1)
            own = set(cell_nodes[c].tolist())
            for n in np.nonzero(inside)[0]:
                if n in own or n in lines:
                    continue
                xi = (X[n] - self.lo[c]) / (self.hi[c] - self.lo[c])
                vals = [lagrange_basis(degree, xi[a])[0] for a in range(self.dim)]
                w = {}
                for il, l in enumerate(ls):
                    v = np.prod([vals[a][l[a]] for a in range(self.dim)])
                    if abs(v) > 1e-14:
                        w[int(cell_nodes[c][il])] = float(v)
                lines[int(n)] = w
        changed = True
        while changed:  # masters that are themselves hanging
            changed = False
            for n, w in lines.items():
                for m in list(w):
                    if m in lines:
                        wm = w.pop(m)
                        for mm, v in lines[m].items():
                            w[mm] = w.get(mm, 0.0) + wm * v
                        changed = True
        return lines

    def _expansion(self, lines, n, ncomp=1):
        rows, cols, vals = [], [], []
        cons = set()
        for node, w in lines.items():
            for k in range(ncomp):
                cons.add(node * ncomp + k)
                for m, v in w.items():
                    rows.append(node * ncomp + k)
                    cols.append(m * ncomp + k)
                    vals.append(v)
        for i in range(n):
            if i not in cons:
                rows.append(i)
                cols.append(i)
                vals.append(1.0)
        return sp.csr_matrix((vals, (rows, cols)), shape=(n, n))

    # ---- pressure
    def _assemble_pressure(self):
        dim = self.dim
        N, dN, W, pts = shape_tables(dim, 1, 2)
        rows, cols, mv, kv = [], [], [], []
        f = np.zeros(self.np_)
        rw, rate = self.prm["well_radius"], self.prm["flow_rate"]
        for c in range(self.nc):
            h = self.hi[c] - self.lo[c]
            detJ = np.prod(h)
            G = dN / h
            ids = self.cell_p[c]
            rows.append(np.repeat(ids, len(ids)))
            cols.append(np.tile(ids, len(ids)))
            mv.append((np.einsum("qi,qj,q->ij", N, N, W) * detJ).ravel())
            kv.append((np.einsum("qia,qja,q->ij", G, G, W) * detJ).ravel())
            xq = self.lo[c] + pts * h
            fq = np.where(xq[:, 0] ** 2 + xq[:, 1] ** 2 <= rw * rw, -rate / (3.1415926 * rw * rw), 0.0)
            f[ids] += (N * (fq * W)[:, None]).sum(axis=0) * detJ
        rows, cols = np.concatenate(rows), np.concatenate(cols)
        self.M = sp.csr_matrix((np.concatenate(mv), (rows, cols)), shape=(self.np_, self.np_))
        self.K = sp.csr_matrix((np.concatenate(kv), (rows, cols)), shape=(self.np_, self.np_))
        self.f = f
        self.Ep = self._expansion(self.lines_p, self.np_)
        self.cons_p = np.zeros(self.np_, bool)
        self.cons_p[list(self.lines_p)] = True

    def condensed(self, X):
        """ConstraintMatrix::condense(SparseMatrix): E^T X E, constrained diagonal = mean |diagonal| of X."""
        Xc = (self.Ep.T @ X @ self.Ep).tolil()
        avg = np.abs(X.diagonal()).mean()
        for i in np.nonzero(self.cons_p)[0]:
            Xc[i, i] = avg
        return Xc.tocsr()

    def solve_condensed(self, Xc, rhs_condensed):
        free = ~self.cons_p
        x = np.zeros(self.np_)
        x[free] = spla.spsolve(Xc[free][:, free].tocsc(), rhs_condensed[free])
        return self.Ep @ x  # distribute (homogeneous lines)

    # ---- displacement
    def assemble_displacement(self, dirichlet):
        dim, du = self.dim, self.degree_u
        lam, mu = self.prm["lame_lambda"], self.prm["shear_modulus"]
        Nu, dNu, W, pts = shape_tables(dim, du, du + 1)
        Np, _, _, _ = shape_tables(dim, 1, du + 1)
        ns = Nu.shape[1]
        nv = 3 if dim == 2 else 6
        D = np.zeros((nv, nv))
        D[:dim, :dim] = lam
        D[np.arange(dim), np.arange(dim)] += 2 * mu
        D[np.arange(dim, nv), np.arange(dim, nv)] = mu
        shear_pairs = [(0, 1)] if dim == 2 else [(0, 1), (0, 2), (1, 2)]
        rows, cols, av, crow, ccol, cv = [], [], [], [], [], []
        for c in range(self.nc):
            h = self.hi[c] - self.lo[c]
            detJ = np.prod(h)
            Gu = dNu / h
            Ae = np.zeros((ns * dim, ns * dim))
            Ce = np.zeros((ns * dim, Np.shape[1]))
            for q in range(len(W)):
                B = np.zeros((nv, ns * dim))
                for s in range(ns):
                    for a in range(dim):
                        B[a, s * dim + a] = Gu[q, s, a]
                    for k, (a, b) in enumerate(shear_pairs):
                        B[dim + k, s * dim + a] = Gu[q, s, b]
                        B[dim + k, s * dim + b] = Gu[q, s, a]
                Ae += B.T @ D @ B * W[q] * detJ
                Ce += np.outer(B[:dim].sum(axis=0), Np[q]) * W[q] * detJ
            idu = (self.cell_u[c][:, None] * dim + np.arange(dim)[None, :]).ravel()
            idp = self.cell_p[c]
            rows.append(np.repeat(idu, len(idu)))
            cols.append(np.tile(idu, len(idu)))
            av.append(Ae.ravel())
            crow.append(np.repeat(idu, len(idp)))
            ccol.append(np.tile(idp, len(idu)))
            cv.append(Ce.ravel())
        self.A_full = sp.csr_matrix((np.concatenate(av), (np.concatenate(rows), np.concatenate(cols))), shape=(self.nu, self.nu))
        self.Gc = sp.csr_matrix((np.concatenate(cv), (np.concatenate(crow), np.concatenate(ccol))), shape=(self.nu, self.np_))
        # constraints: hanging lines first, then Dirichlet values on dofs that are not constrained yet (DS:112-134), then close()
        hang = self._hanging_lines(du, self.xu, self.cell_u)
        lines = {}
        for node, w in hang.items():
            for k in range(dim):
                lines[node * dim + k] = ({m * dim + k: v for m, v in w.items()}, 0.0)
        box_lo, box_hi = self.lo.min(axis=0), self.hi.max(axis=0)
        tol = 1e-9 * (box_hi - box_lo).max()
        for label, comp, value in dirichlet:
            axis, side = label // 2, label % 2
            coord = box_lo[axis] if side == 0 else box_hi[axis]
            for node in np.nonzero(np.abs(self.xu[:, axis] - coord) < tol)[0]:
                d = int(node) * dim + comp
                if d not in lines:
                    lines[d] = ({}, float(value))
        changed = True
        while changed:
            changed = False
            for d, (w, g) in list(lines.items()):
                for m in list(w):
                    if m in lines:
                        wm = w.pop(m)
                        wm2, g2 = lines[m]
                        for mm, v in wm2.items():
                            w[mm] = w.get(mm, 0.0) + wm * v
                        g += wm * g2
                        changed = True
                lines[d] = (w, g)
        self.lines_u = lines
        n = self.nu
        rows, cols, vals = [], [], []
        self.g = np.zeros(n)
        self.cons = np.zeros(n, bool)
        for d, (w, g) in lines.items():
            self.cons[d] = True
            self.g[d] = g
            for m, v in w.items():
                rows.append(d)
                cols.append(m)
                vals.append(v)
        free = np.nonzero(~self.cons)[0]
        rows += free.tolist()
        cols += free.tolist()
        vals += [1.0] * len(free)
        self.Eu = sp.csr_matrix((vals, (rows, cols)), shape=(n, n))
        self.free = ~self.cons
        Ah = self.Eu.T @ self.A_full @ self.Eu
        self.A_ff = Ah[self.free][:, self.free].tocsc()
        self.lu = spla.splu(self.A_ff)

    def condensed_elasticity(self):
        """system_matrix after distribute_local_to_global: E^T A E on the free dofs, constrained diagonal = sum of the
        cells' |a_ii| (= the assembled diagonal, all positive)."""
        Ah = (self.Eu.T @ self.A_full @ self.Eu).tolil()
        d = self.A_full.diagonal()
        for i in np.nonzero(self.cons)[0]:
            Ah[i, i] = d[i]
        Ah = Ah.tocsr()
        Ah.eliminate_zeros()
        return Ah

    def rhs_displacement(self, p):
        f = self.prm["biot_coef"] * (self.Gc @ p)
        out = self.Eu.T @ (f - self.A_full @ self.g)
        out[self.cons] = 0.0
        return out

    def solve_displacement(self, p):
        rhs = self.rhs_displacement(p)
        x = np.zeros(self.nu)
        x[self.free] = self.lu.solve(rhs[self.free])
        return self.Eu @ x + self.g

    def project_strains(self, u, comps):
        dim, du = self.dim, self.degree_u
        Np, _, W, _ = shape_tables(dim, 1, 2)
        _, dNu, _, _ = shape_tables(dim, du, 2)
        rhs = {c: np.zeros(self.np_) for c in comps}
        U = u.reshape(-1, dim)
        for c in range(self.nc):
            h = self.hi[c] - self.lo[c]
            Gu = dNu / h
            grad = np.einsum("sc,qsa->qca", U[self.cell_u[c]], Gu)
            eps = 0.5 * (grad + grad.transpose(0, 2, 1))
            for comp in comps:
                i, j = comp // dim, comp % dim
                rhs[comp][self.cell_p[c]] += (Np * (eps[:, i, j] * W)[:, None]).sum(axis=0) * np.prod(h)
        Mc = self.condensed(self.M)
        out = {}
        for c in comps:
            rc = self.Ep.T @ rhs[c]
            rc[self.cons_p] = 0.0
            rhs[c] = rc
            out[c] = self.solve_condensed(Mc, rc)
        return out, rhs

    def initialize(self, p_init):
        self.p = np.full(self.np_, float(p_init))
        self.u = self.solve_displacement(self.p)
        vol = [0, 3] if self.dim == 2 else [0, 4, 8]
        strains, _ = self.project_strains(self.u, vol)
        self.ev = sum(strains[c] for c in vol)
        self.ev0 = self.ev.copy()

    def residual(self, dt, p_old):
        P = self.prm
        t1 = (self.ev - self.ev0) * (P["biot_coef"] / dt) + (self.p - p_old) * (1.0 / P["m_modulus"] / dt)
        r = -(self.M @ t1 + P["perm_over_visc"] * (self.K @ self.p) + self.f)
        r = self.Ep.T @ r  # condense(residual), PS:153
        r[self.cons_p] = 0.0
        return r

    def time_step(self, dt, tol=1e-8, max_inner=50):
        P = self.prm
        p_old = self.p.copy()
        Jc = self.condensed(self.M * (1.0 / P["m_modulus"] / dt) + P["perm_over_visc"] * self.K)
        dp = np.zeros(self.np_)
        hist = []
        for it in range(max_inner):
            self.ev = self.ev + (P["biot_coef"] / P["bulk_modulus"]) * dp
            r = self.residual(dt, p_old)
            hist.append(np.linalg.norm(r))
            if hist[-1] < tol:
                break
            dp = self.solve_condensed(Jc, r)
            self.p = self.p + dp
        self.u = self.solve_displacement(self.p)
        return hist
