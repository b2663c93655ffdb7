"""Adaptive-refinement host surface (csrc/host/amr.hpp) against independent checks.

The reference delegates all of this to deal.II (PoroelasticityFSS.h:333-340, 447-498: Triangulation,
make_hanging_node_constraints, KellyErrorEstimator, refine_and_coarsen_fixed_fraction, SolutionTransfer); none of it
can be run here, so every property is checked against something independent of the C++ code: geometry-based
neighbour searches in numpy, closed forms, and hand-computed threshold examples.
"""
import itertools

import numpy as np
import pytest

import helpers as H

capi = H.capi


def refine_where(F, pred, rounds=1):
    """Flag the active cells whose centre satisfies pred(centre, level) and execute; returns the new active mesh."""
    for _ in range(rounds):
        m = F.active_mesh().arrays
        ctr = m.xyz[m.cell_vertices].mean(axis=1)
        lv = F.levels()
        F.set_flags(refine=np.array([pred(c, l) for c, l in zip(ctr, lv)], dtype=np.int8))
        F.execute()
    return F.active_mesh()


def corner_refined_forest(dim, rounds=3, base=1):
    """Box refined towards the corner (-5,..,-5): a staircase of levels with hanging nodes on faces and edges."""
    m0 = capi.mesh_rectangle(dim, [10.0] * dim, base)
    F = capi.Forest(m0, base)
    target = -5.0 * np.ones(dim)
    for r in range(rounds):
        refine_where(F, lambda c, l, r=r: np.linalg.norm(c - target) < 6.0 / (r + 1))
    return F


def cell_boxes(m):
    x = m.xyz[m.cell_vertices]
    return x.min(axis=1), x.max(axis=1)


def touching_pairs(m, min_shared_dim):
    """Pairs of axis-aligned cells whose closed boxes share a set of dimension >= min_shared_dim (geometry only)."""
    lo, hi = cell_boxes(m)
    n, dim = lo.shape
    pairs = []
    for i in range(n):
        ov_lo = np.maximum(lo[i], lo[i + 1:])
        ov_hi = np.minimum(hi[i], hi[i + 1:])
        ext = ov_hi - ov_lo
        touch = (ext > -1e-12).all(axis=1)
        shared_dim = (ext > 1e-12).sum(axis=1)
        for j in np.nonzero(touch & (shared_dim >= min_shared_dim))[0]:
            pairs.append((i, i + 1 + j))
    return pairs


@pytest.mark.parametrize("dim", [2, 3])
def test_refinement_keeps_two_to_one_balance_across_lines(dim):
    F = corner_refined_forest(dim, rounds=4 if dim == 2 else 3)
    m = F.active_mesh().arrays
    lv = F.levels()
    assert lv.max() - lv.min() >= 2  # a real staircase
    # cells sharing (part of) a line: 2D faces, 3D faces and edges
    for i, j in touching_pairs(m, 1):
        assert abs(int(lv[i]) - int(lv[j])) <= 1, (i, j, lv[i], lv[j])
    # the active cells tile the box exactly
    lo, hi = cell_boxes(m)
    assert np.isclose(np.prod(hi - lo, axis=1).sum(), 10.0 ** dim)
    # boundary faces tile the boundary and carry the colorized ids (FSS:429-432)
    area = {}
    for c, f, bid in zip(m.bface_cell, m.bface_local, m.bface_id):
        assert bid == f
        ext = np.delete(hi[c] - lo[c], f // 2)
        area[bid] = area.get(bid, 0.0) + np.prod(ext)
    assert len(area) == 2 * dim and all(np.isclose(a, 10.0 ** (dim - 1)) for a in area.values())


def test_coarsening_needs_the_whole_family_and_respects_balance():
    m0 = capi.mesh_rectangle(2, [10.0, 10.0], 1)
    F = capi.Forest(m0, 1)
    refine_where(F, lambda c, l: c[0] < 0 and c[1] < 0)            # cell 0 -> 4 children
    refine_where(F, lambda c, l: l == 2 and c[0] > -2.5 and c[1] > -2.5)  # its inner child again (neighbours follow)
    lv = F.levels()
    n = len(lv)
    # three of four siblings flagged: nothing is coarsened
    co = np.zeros(n, dtype=np.int8)
    idx3 = np.nonzero(lv == 3)[0]
    co[idx3[:3]] = 1
    F.set_flags(coarsen=co)
    F.prepare()
    assert not F.get_flags()[1].any()
    # the level-3 family may go; level-2 cells next to remaining level-3 cells may not
    co = np.ones(n, dtype=np.int8)
    F.set_flags(coarsen=co)
    nc, nr = F.execute()
    assert nr == 0 and nc >= 1
    m = F.active_mesh().arrays
    lv = F.levels()
    for i, j in touching_pairs(m, 1):
        assert abs(int(lv[i]) - int(lv[j])) <= 1
    # roots (level == base level) are never coarsened
    for _ in range(4):
        F.set_flags(coarsen=np.ones(len(F.levels()), dtype=np.int8))
        F.execute()
    assert (F.levels() == 1).all() and len(F.levels()) == 4


def test_refinement_wins_over_coarsening():
    m0 = capi.mesh_rectangle(2, [10.0, 10.0], 2)
    F = capi.Forest(m0, 1)  # pretend the 4x4 grid sits one level above its roots: nothing to coarsen below it
    refine_where(F, lambda c, l: c[0] < 0)  # left half -> level 2
    m = F.active_mesh().arrays
    ctr = m.xyz[m.cell_vertices].mean(axis=1)
    lv = F.levels()
    # refine one fine cell at the interface and ask every other fine cell to coarsen
    pick = np.argmin(np.abs(ctr[:, 0] + 0.625) + np.abs(ctr[:, 1] - 0.625) + 100 * (lv != 2))
    re = np.zeros(len(lv), dtype=np.int8)
    re[pick] = 1
    co = (lv == 2).astype(np.int8)
    co[pick] = 0
    F.set_flags(refine=re, coarsen=co)
    F.prepare()
    r, c = F.get_flags()
    assert r[pick] and not (r & c).any()
    # the coarse neighbour across the interface must refine, and the picked cell's own family must stay
    sib = [i for i in range(len(lv)) if lv[i] == 2 and np.abs(ctr[i] - ctr[pick]).max() < 1.26 and i != pick]
    assert not c[sib].any()
    F.execute()
    m = F.active_mesh().arrays
    lv = F.levels()
    for i, j in touching_pairs(m, 1):
        assert abs(int(lv[i]) - int(lv[j])) <= 1


def fe_eval_on_lattice(m, dofs, values, degree, npts=5):
    """Evaluate an FE_Q(degree) function cell by cell on an npts^dim lattice of every (axis-aligned) cell."""
    dim = m.dim
    lo, hi = cell_boxes(m)
    sup = dofs.support_points()
    out = {}
    t = np.linspace(0, 1, npts)

    def lag(deg, node, x):
        nodes = np.linspace(0, 1, deg + 1)
        v = 1.0
        for k, nk in enumerate(nodes):
            if k != node:
                v = v * (x - nk) / (nodes[node] - nk)
        return v

    for c in range(m.n_cells):
        cd = dofs.cell_dofs[c][:: dofs.n_comp]
        unit = (sup[cd] - lo[c]) / (hi[c] - lo[c])
        node = np.rint(unit * degree).astype(int)
        for idx in itertools.product(range(npts), repeat=dim):
            xi = t[list(idx)]
            val = 0.0
            for s in range(len(cd)):
                w = 1.0
                for a in range(dim):
                    w *= lag(degree, node[s, a], xi[a])
                val += w * values[cd[s]]
            key = tuple(np.round(lo[c] + xi * (hi[c] - lo[c]), 9))
            out.setdefault(key, []).append(val)
    return out


@pytest.mark.parametrize("dim,degree", [(2, 1), (2, 2), (3, 1), (3, 2)])
def test_hanging_node_constraints_make_the_space_continuous(dim, degree):
    F = corner_refined_forest(dim, rounds=3 if dim == 2 else 2)
    am = F.active_mesh()
    dofs = capi.HostDofs(am, degree, 1)
    L = capi.make_constraints(F, am, dofs)
    assert L.n_lines > 0
    constrained = set(L.line_dof.tolist())
    # masters are unconstrained after close(); weights of a line sum to one (constants are reproduced)
    assert not (set(L.entry_dof.tolist()) & constrained)
    for i in range(L.n_lines):
        _, ed, ew, g = L.line(i)
        assert len(ed) > 0 and np.isclose(ew.sum(), 1.0) and g == 0.0
    # random values on the free dofs, hanging dofs by their lines -> the function is single-valued on every shared
    # face / edge (independent check: lattice points are matched by coordinates, not by connectivity)
    rng = np.random.default_rng(5)
    v = rng.standard_normal(dofs.n_dofs)
    for i in range(L.n_lines):
        d, ed, ew, g = L.line(i)
        v[d] = v[ed] @ ew + g
    samples = fe_eval_on_lattice(am.arrays, dofs, v, degree)
    worst = max(max(vals) - min(vals) for vals in samples.values())
    assert worst < 1e-12
    # and without the constraints it is not (the test has teeth)
    v2 = v.copy()
    v2[L.line_dof[0]] += 1.0
    samples = fe_eval_on_lattice(am.arrays, dofs, v2, degree)
    assert max(max(vals) - min(vals) for vals in samples.values()) > 0.1


@pytest.mark.parametrize("dim,degree", [(2, 2), (3, 1)])
def test_vector_constraints_repeat_the_scalar_weights_and_dirichlet_never_overwrites(dim, degree):
    F = corner_refined_forest(dim, rounds=2)
    am = F.active_mesh()
    ds = capi.HostDofs(am, degree, 1)
    dv = capi.HostDofs(am, degree, dim)
    Ls = capi.make_constraints(F, am, ds)
    Lv = capi.make_constraints(F, am, dv)
    assert Lv.n_lines == dim * Ls.n_lines
    for i in range(Ls.n_lines):
        d, ed, ew, _ = Ls.line(i)
        for k in range(dim):
            dk, edk, ewk, _ = Lv.line(dim * i + k)
            assert dk == dim * d + k and (edk == dim * ed + k).all() and np.allclose(ewk, ew)
    # Dirichlet on the refined corner's faces: x-min fixes component 0 to 0.25, y-min fixes component 1 to -1
    Ld = capi.make_constraints(F, am, dv, labels=[0, 2], comps=[0, 1], values=[0.25, -1.0])
    sp = dv.support_points()
    lines = {Ld.line(i)[0]: Ld.line(i) for i in range(Ld.n_lines)}
    hanging = set(Lv.line_dof.tolist())
    n_resolved = 0
    for d, (_, ed, ew, g) in lines.items():
        comp = d % dim
        on0 = np.isclose(sp[d, 0], -5.0) and comp == 0
        on2 = np.isclose(sp[d, 1], -5.0) and comp == 1
        # masters are free dofs
        assert not any(int(e) in lines for e in ed)
        if on0 or on2:
            # boundary dof of the constrained component: fully resolved to its value, hanging or not
            assert len(ed) == 0 and np.isclose(g, 0.25 if on0 else -1.0)
            n_resolved += d in hanging
        else:
            assert d in hanging
            # a hanging dof with some parents on the Dirichlet face keeps the free parents and an inhomogeneity
            assert np.isclose(ew.sum() + (g / (0.25 if comp == 0 else -1.0) if g != 0 else 0.0), 1.0)
    assert n_resolved > 0 or dim == 2  # in 2D hanging nodes sit on interior lines only
    assert any(len(l[1]) > 0 and l[3] != 0 for l in lines.values())  # the mixed case occurs on this mesh


def numpy_kelly(m, p_vertex):
    """Independent estimator for axis-aligned cells: neighbours by geometry, jumps of the Q1 normal derivative."""
    dim = m.dim
    lo, hi = cell_boxes(m)
    ga = 0.5 - 0.5 / np.sqrt(3.0)
    gp = [ga, 1 - ga]

    def grad(c, x):
        xi = (x - lo[c]) / (hi[c] - lo[c])
        g = np.zeros(dim)
        for k in range(1 << dim):
            vtx = m.cell_vertices[c, k]
            # which corner is it geometrically?
            bits = np.rint((m.xyz[vtx] - lo[c]) / (hi[c] - lo[c])).astype(int)
            for a in range(dim):
                w = (1.0 if bits[a] else -1.0) / (hi[c, a] - lo[c, a])
                for b in range(dim):
                    if b != a:
                        w *= xi[b] if bits[b] else 1 - xi[b]
                g[a] += p_vertex[vtx] * w
        return g

    acc = np.zeros(m.n_cells)
    for i, j in touching_pairs(m, dim - 1):
        ov_lo, ov_hi = np.maximum(lo[i], lo[j]), np.minimum(hi[i], hi[j])
        ext = ov_hi - ov_lo
        axis = int(np.argmin(ext))
        others = [a for a in range(dim) if a != axis]
        area = np.prod(ext[others])
        integral = 0.0
        for q in itertools.product(gp, repeat=dim - 1):
            x = ov_lo.copy()
            for a, t in zip(others, q):
                x[a] = ov_lo[a] + t * ext[a]
            jump = grad(i, x)[axis] - grad(j, x)[axis]
            integral += jump * jump * area / len(gp) ** (dim - 1)
        acc[i] += integral
        acc[j] += integral
    h = np.linalg.norm(hi - lo, axis=1)
    return np.sqrt(acc * h / 24.0)


@pytest.mark.parametrize("dim", [2, 3])
def test_kelly_estimator_closed_form_on_a_uniform_mesh(dim):
    n = 4
    m0 = capi.mesh_rectangle(dim, [10.0] * dim, 2)
    F = capi.Forest(m0, 2)
    am = F.active_mesh()
    dp = capi.HostDofs(am, 1, 1)
    x = dp.support_points()
    eta = F.kelly(am, dp, x[:, 0] ** 2)
    h = 10.0 / n
    # the Q1 interpolant of x^2 has slope jumps of 2h across the faces normal to x and none elsewhere;
    # eta^2 = diam/24 * (#interior x-faces) * (2h)^2 * h^(dim-1)
    m = am.arrays
    ctr = m.xyz[m.cell_vertices].mean(axis=1)
    n_faces = np.where((ctr[:, 0] < -5 + h) | (ctr[:, 0] > 5 - h), 1, 2)
    expect = np.sqrt(np.sqrt(dim) * h / 24.0 * n_faces * (2 * h) ** 2 * h ** (dim - 1))
    assert eta.dtype == np.float32  # Vector<float>, FSS:452
    assert np.allclose(eta, expect.astype(np.float32), rtol=1e-6)
    # linear fields have no jumps at all
    assert np.abs(F.kelly(am, dp, 3 * x[:, 0] - 2 * x[:, dim - 1] + 1)).max() < 1e-6


@pytest.mark.parametrize("dim", [2, 3])
def test_kelly_estimator_on_hanging_node_meshes_matches_a_geometric_restatement(dim):
    F = corner_refined_forest(dim, rounds=3 if dim == 2 else 2)
    am = F.active_mesh()
    dp = capi.HostDofs(am, 1, 1)
    L = capi.make_constraints(F, am, dp)
    x = dp.support_points()
    p = np.sin(0.3 * x[:, 0]) * np.cos(0.2 * x[:, 1]) + 0.05 * x[:, dim - 1] ** 2
    for i in range(L.n_lines):  # a conforming field, like every pressure the time loop produces
        d, ed, ew, g = L.line(i)
        p[d] = p[ed] @ ew
    eta = F.kelly(am, dp, p)
    m = am.arrays
    pv = np.zeros(m.n_vertices)
    pv[m.cell_vertices.ravel()] = p[dp.cell_dofs.ravel()]
    ref = numpy_kelly(m, pv)
    assert np.allclose(eta, ref.astype(np.float32), rtol=2e-6, atol=1e-9)
    assert eta.max() > 1e-3


def test_fixed_fraction_marking_follows_grid_refinement():
    """Hand-computed thresholds of GridRefinement::refine_and_coarsen_fixed_fraction(0.6, 0.4) (FSS:460-462)."""
    m0 = capi.mesh_rectangle(2, [10.0, 10.0], 2)  # 16 cells
    F = capi.Forest(m0, 1)  # all 16 cells sit at level 1
    crit = np.arange(1, 17, dtype=np.float32)  # total 136
    # (0.3, 0.1): descending sums 16, 31, 45 >= 40.8 -> 3 cells, threshold (13 + 14) / 2; ascending sums 1, 3, 6, 10, 15 >= 13.6
    # -> the iterator stops on 6, threshold (6 + 5) / 2
    F.mark_fixed_fraction(crit, 0.3, 0.1, min_level=0, max_level=9)
    r, c = F.get_flags()
    assert (r == (crit >= 13.5)).all() and r.sum() == 3
    assert (c == (crit <= 5.5)).all()
    # (0.6, 0.4) as in the reference: descending sums ... 81 < 81.6, 91 -> 7 cells, threshold (9 + 10) / 2 = 9.5; ascending sums
    # 1, 3, ..., 45, 55 >= 54.4 -> threshold (11 + 10) / 2 = 10.5 >= 9.5, so it is lowered to 0.999 * 9.5
    F.mark_fixed_fraction(crit, 0.6, 0.4, min_level=0, max_level=9)
    r, c = F.get_flags()
    assert (r == (crit >= 9.5)).all() and r.sum() == 7
    assert (c == (crit <= 0.999 * 9.5)).all() and c.sum() == 9
    # level limits of FSS:463-472: cells at min_level keep their size, cells at max_level are not refined further
    F.mark_fixed_fraction(crit, 0.6, 0.4, min_level=1, max_level=9)
    r, c = F.get_flags()
    assert r.sum() == 7 and not c.any()
    F.mark_fixed_fraction(crit, 0.6, 0.4, min_level=0, max_level=2)
    r, c = F.get_flags()
    assert r.sum() == 7  # n_levels() == 2 is not > 2: the level-1 cells may still refine (FSS:463)
    F.mark_fixed_fraction(crit, 0.6, 0.4, min_level=0, max_level=1)
    r, c = F.get_flags()
    assert not r.any() and c.sum() == 9  # cells already at max_grid_level lose their refine flag (FSS:464-467)
    # all indicators equal: thresholds coincide, the top one is lowered by a permille -> everything refines, nothing coarsens
    F.mark_fixed_fraction(np.ones(16, dtype=np.float32), 0.6, 0.4, min_level=0, max_level=9)
    r, c = F.get_flags()
    assert not c.any()
    # all zero: no refinement (GridRefinement::refine returns early), no coarsening (threshold not above the minimum)
    F.mark_fixed_fraction(np.zeros(16, dtype=np.float32), 0.6, 0.4, min_level=0, max_level=9)
    r, c = F.get_flags()
    assert not r.any() and not c.any()


@pytest.mark.parametrize("dim", [2, 3])
def test_solution_transfer_keeps_old_values_and_interpolates_new_vertices(dim):
    F = corner_refined_forest(dim, rounds=2)
    am = F.active_mesh()
    dp = capi.HostDofs(am, 1, 1)
    x = dp.support_points()
    lin = 2.0 + x @ np.arange(1, dim + 1)
    rng = np.random.default_rng(3)
    rnd = rng.standard_normal(dp.n_dofs)
    L = capi.make_constraints(F, am, dp)
    for i in range(L.n_lines):
        d, ed, ew, _ = L.line(i)
        rnd[d] = rnd[ed] @ ew
    F.store(am, dp, [lin, rnd])
    # refine near one corner, coarsen near the opposite one
    m = am.arrays
    ctr = m.xyz[m.cell_vertices].mean(axis=1)
    F.set_flags(refine=(ctr.sum(axis=1) < -2.0 * dim).astype(np.int8), coarsen=(ctr.sum(axis=1) > 0).astype(np.int8))
    nc, nr = F.execute()
    assert nr > 0
    am2 = F.active_mesh()
    dp2 = capi.HostDofs(am2, 1, 1)
    out = F.fetch(am2, dp2, 2)
    x2 = dp2.support_points()
    assert np.allclose(out[0], 2.0 + x2 @ np.arange(1, dim + 1), atol=1e-12)  # Q1 interpolation is exact for linears
    # dofs that existed before keep their value exactly
    old = {tuple(np.round(p, 9)): v for p, v in zip(x, rnd)}
    kept = [(i, old[tuple(np.round(p, 9))]) for i, p in enumerate(x2) if tuple(np.round(p, 9)) in old]
    assert len(kept) >= 15
    assert all(out[1][i] == v for i, v in kept)
    # new vertices carry the parent's interpolant: the transferred field still satisfies the new mesh's constraints
    # wherever the constraining line lies inside a refined (not a coarsened) region -> check on the refined corner
    L2 = capi.make_constraints(F, am2, dp2)
    for i in range(L2.n_lines):
        d, ed, ew, _ = L2.line(i)
        if x2[d].sum() < -2.0 * dim - 2.5:
            assert np.isclose(out[1][d], out[1][ed] @ ew, atol=1e-12)


# ---- unstructured coarse meshes with mixed cell orientations (read_mesh(), FSS:438-445) --------------------------------
def _hex_rotations():
    """node permutations of a Gmsh hexahedron under the 24 proper rotations of the cube"""
    pos = np.array([[0, 0, 0], [1, 0, 0], [1, 1, 0], [0, 1, 0], [0, 0, 1], [1, 0, 1], [1, 1, 1], [0, 1, 1]])
    perms = set()
    for axes in itertools.permutations(range(3)):
        for flips in itertools.product([0, 1], repeat=3):
            R = np.zeros((3, 3))
            for a in range(3):
                R[a, axes[a]] = -1.0 if flips[a] else 1.0
            if np.linalg.det(R) < 0:
                continue
            img = (pos - 0.5) @ R.T + 0.5
            perms.add(tuple(int(np.argmin(np.abs(pos - p).sum(axis=1))) for p in img))
    return sorted(perms)


def rotated_msh(src, dst, dim, seed):
    """Copy of a Gmsh 2.2 file in which every cell's node list is replaced by a randomly rotated, equally oriented one."""
    rng = np.random.default_rng(seed)
    rots = _hex_rotations()
    assert len(rots) == 24
    out, in_elems, count_line = [], False, False
    for line in open(src).read().splitlines():
        if line.startswith("$Elements"):
            in_elems, count_line = True, True
        elif line.startswith("$EndElements"):
            in_elems = False
        elif in_elems and count_line:
            count_line = False
        elif in_elems:
            t = line.split()
            etype, ntags = int(t[1]), int(t[2])
            head, nodes = t[: 3 + ntags], t[3 + ntags:]
            if dim == 2 and etype == 3:
                k = int(rng.integers(4))
                nodes = nodes[k:] + nodes[:k]
            elif dim == 3 and etype == 5:
                perm = rots[int(rng.integers(24))]
                nodes = [nodes[i] for i in perm]
            line = " ".join(head + nodes)
        out.append(line)
    open(dst, "w").write("\n".join(out) + "\n")


@pytest.mark.parametrize("dim,fname,degree", [(2, "distorted_quad8.msh", 1), (2, "distorted_quad8.msh", 2), (3, "distorted_hex4.msh", 1), (3, "distorted_hex4.msh", 2)])
def test_refinement_of_unstructured_meshes_does_not_depend_on_cell_orientation(dim, fname, degree, tmp_path):
    """Lines, quads and hanging nodes are identified by vertex numbers, never by local indices of a neighbour, so a mesh
    whose cells are rotated at random (what a mesh generator may produce) must give the same refined mesh, the same
    constraint lines and the same error indicators as the consistently oriented one."""
    src = H.ROOT / "tests" / "golden" / fname
    rot = tmp_path / "rotated.msh"
    rotated_msh(src, rot, dim, seed=11)
    results = []
    for path in (src, rot):
        mesh = capi.mesh_read_msh(path, dim)
        F = capi.Forest(mesh, 0)
        for rnd in range(2):
            m = F.active_mesh().arrays
            ctr = m.xyz[m.cell_vertices].mean(axis=1)
            # flags from the geometry only, so both meshes refine the same cells
            F.set_flags(refine=(np.sin(3.1 * ctr[:, 0] + rnd) * np.cos(2.3 * ctr[:, dim - 1]) > 0.35).astype(np.int8))
            F.execute()
        am = F.active_mesh()
        d = capi.HostDofs(am, degree, 1)
        L = capi.make_constraints(F, am, d)
        sp = d.support_points()
        kk = lambda x: tuple(np.round(x, 8))
        lines = {}
        for i in range(L.n_lines):
            dof, ed, ew, g = L.line(i)
            if len(ed) == 1 and kk(sp[ed[0]]) == kk(sp[dof]):
                continue  # FE_Q(2) identity line of a duplicated midpoint dof: both dofs share the coordinates
            lines[kk(sp[dof])] = sorted((kk(sp[e]), round(float(w), 12)) for e, w in zip(ed, ew))
        dp = capi.HostDofs(am, 1, 1)
        xp = dp.support_points()
        p = np.sin(0.4 * xp[:, 0]) + 0.03 * xp[:, dim - 1] ** 2 + 0.1 * xp[:, 0] * xp[:, 1]
        Lp = capi.make_constraints(F, am, dp)
        for i in range(Lp.n_lines):
            dof, ed, ew, _ = Lp.line(i)
            p[dof] = p[ed] @ ew
        eta = F.kelly(am, dp, p)
        a = am.arrays
        ctr = a.xyz[a.cell_vertices].mean(axis=1)
        results.append((a.n_cells, lines, {kk(c): float(e) for c, e in zip(ctr, eta)}, sorted(F.levels().tolist())))
    (n0, l0, e0, lv0), (n1, l1, e1, lv1) = results
    assert n0 == n1 and lv0 == lv1 and max(lv0) == 2 and len(l0) > 0
    assert l0 == l1
    assert set(e0) == set(e1)
    assert max(abs(e0[k] - e1[k]) for k in e0) <= 1e-6 * max(e0.values())


@pytest.mark.parametrize("dim,seed", [(2, 0), (2, 1), (3, 2)])
def test_random_refine_coarsen_sequences_keep_every_invariant(dim, seed):
    """Six rounds of random refine / coarsen flags: after every round the active cells tile the box, levels across shared
    lines differ by at most one, the constrained FE_Q(2) space is continuous, the transfer reproduces a linear field, and
    flags requested on cells that may legally change are honoured (refinement always, coarsening for whole families whose
    neighbours allow it)."""
    rng = np.random.default_rng(seed)
    m0 = capi.mesh_rectangle(dim, [10.0] * dim, 1)
    F = capi.Forest(m0, 1)
    lin = lambda x: 1.5 + x @ np.arange(1, dim + 1)
    for rnd in range(6 if dim == 2 else 3):
        am = F.active_mesh()
        dp = capi.HostDofs(am, 1, 1)
        F.store(am, dp, [lin(dp.support_points())])
        lv = F.levels()
        re = (rng.random(len(lv)) < 0.2) & (lv < 4)
        co = (rng.random(len(lv)) < 0.5) & ~re
        F.set_flags(refine=re.astype(np.int8), coarsen=co.astype(np.int8))
        F.prepare()
        r2, c2 = F.get_flags()
        assert (r2 | ~re).all()          # prepare never drops a refine flag
        assert not (c2 & ~co).any()      # and never invents a coarsen flag
        n_before = len(lv)
        nc, nr = F.execute()
        assert nr == r2.sum()
        am = F.active_mesh()
        m = am.arrays
        lv = F.levels()
        assert len(lv) == n_before + ((1 << dim) - 1) * (nr - nc)
        lo, hi = cell_boxes(m)
        assert np.isclose(np.prod(hi - lo, axis=1).sum(), 10.0 ** dim)
        for i, j in touching_pairs(m, 1):
            assert abs(int(lv[i]) - int(lv[j])) <= 1
        dp = capi.HostDofs(am, 1, 1)
        assert np.allclose(F.fetch(am, dp, 1)[0], lin(dp.support_points()), atol=1e-12)
        if len(lv) > (600 if dim == 2 else 80):
            continue  # the lattice check below is a python loop over cells x lattice points x nodes
        d2 = capi.HostDofs(am, 2, 1)
        L = capi.make_constraints(F, am, d2)
        v = rng.standard_normal(d2.n_dofs)
        for i in range(L.n_lines):
            d, ed, ew, g = L.line(i)
            v[d] = v[ed] @ ew
        samples = fe_eval_on_lattice(m, d2, v, 2)
        assert max(max(vals) - min(vals) for vals in samples.values()) < 1e-12


def test_refinement_pass_does_not_depend_on_the_thread_count(tmp_path):
    """The estimator's face table and the 2:1 closure's line table are built by all host threads (hash buckets, one thread per
    bucket).  Everything a pass produces — indicators, line numbering, the executed mesh, transferred values, constraint lines —
    has to be the same bytes for 1, 3 and 8 threads (tests/amr_dump.cpp)."""
    import subprocess
    exe = tmp_path / "amr_dump"
    subprocess.check_call(["g++", "-O2", "-fopenmp", "-std=c++17", "-o", str(exe), str(H.ROOT / "tests" / "amr_dump.cpp")])
    for dim, base in ((2, 4), (3, 3)):
        dumps = []
        for threads in (1, 3, 8):
            out = tmp_path / f"dump_{dim}_{threads}.bin"
            subprocess.check_call([str(exe), str(dim), str(base), str(out)], env={"OMP_NUM_THREADS": str(threads), "PATH": "/usr/bin:/bin"},
                                  stdout=subprocess.DEVNULL)
            dumps.append(out.read_bytes())
        assert len(dumps[0]) > 100000 and dumps[0] == dumps[1] == dumps[2]
