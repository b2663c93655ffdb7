"""Host surface: input.data parser (ID:77-222), mesh (FSS:418-445), dof numbering (PS:73, DS:110-135)."""
from pathlib import Path

import numpy as np
import pytest

import helpers as H

capi, fss = H.capi, H.fss
REF = Path("/root/reference")


def _cross2(a, b):
    return a[0] * b[1] - a[1] * b[0]


def test_t1_t2_derived_moduli_and_units():
    d = capi.InputData(text=H.SHIPPED_INPUT)
    # SURVEY §4 T1 / T2 (ID:162, ID:213-222 on input.data:24-35)
    assert d.lame_constant == pytest.approx(8.076923076923077e9, rel=1e-15)
    assert d.shear_modulus == pytest.approx(5.384615384615384e9, rel=1e-15)
    assert d.bulk_modulus == pytest.approx(1.1666666666666666e10, rel=1e-15)
    assert d.grain_bulk_modulus == pytest.approx(1.1666666666666669e11, rel=1e-15)
    assert d.n_modulus == pytest.approx(1.9444444444444446e11, rel=1e-15)
    assert d.m_modulus == pytest.approx(5.58213716108453e9, rel=1e-14)
    assert d.perm == pytest.approx(10 * 9.869233e-16, rel=1e-15)
    assert d.perm / d.visc == pytest.approx(9.869233e-12, rel=1e-15)
    assert (d.dim, d.initial_refinement_level, d.max_refinement_level) == (2, 4, 6)
    assert (d.time_step, d.t_max, d.p_init) == (60.0, 1e3, 10e6)
    assert list(d.displacement_boundary_labels) == [0, 1, 2, 3]
    assert list(d.displacement_boundary_components) == [0, 0, 1, 1]
    assert list(d.displacement_boundary_values) == [0.0, -1e-5, 0.0, -1e-5]
    assert len(d.stress_boundary_labels) == 0
    # defaults of keys the shipped file does not set (ID:136-141)
    assert (d.max_fss_iterations, d.max_pressure_iterations, d.fss_tol, d.pressure_tol) == (50, 50, 1e-8, 1e-8)


def test_defaults_match_declare_parameters():
    d = capi.InputData(text="")
    assert d.dim == 2 and d.domain_size[:2] == [10.0, 10.0] and d.initial_refinement_level == 3
    assert d.youngs_modulus == 7e9 and d.poisson_ratio == 0.3 and d.biot_coef == 0.9
    assert d.f_comp == pytest.approx(45.8e-11) and d.r_well == 0.1 and d.flow_rate == 1e-6
    assert list(d.displacement_boundary_labels) == [0, 2, 3, 1]
    assert list(d.displacement_boundary_values) == [0, 0, 0, -0.1]
    assert d.displacement_degree == 2  # DS:67


@pytest.mark.skipif(not (REF / "input.data").exists(), reason="reference tree not mounted")
def test_shipped_input_file_parses_unchanged():
    a = capi.InputData(path=REF / "input.data")
    b = capi.InputData(text=H.SHIPPED_INPUT)
    for k in ("dim", "perm", "m_modulus", "lame_constant", "time_step", "t_max", "r_well", "flow_rate", "initial_refinement_level"):
        assert getattr(a, k) == getattr(b, k)


@pytest.mark.parametrize("bad,msg", [
    ("subsection Mesh\n set Dimensions = 4\nend\n", "does not match"),          # Integer(1,3)
    ("subsection Properties\n set Poisson ratio = 0.6\nend\n", "does not match"),  # Double(0,0.5)
    ("subsection Mesh\n set Nonsense = 1\nend\n", "No such entry"),
    ("subsection Nope\nend\n", "no such subsection"),
    ("subsection Mesh\n set Dimensions = 2\n", "Unbalanced"),
    ("subsection Mesh\n set Dimensions = 3\n set Domain size = 10, 10\nend\n", "Domain size"),
    ("subsection In situ\n set Displacement boundary labels = 0, 1\nend\n", "differ in length"),
])
def test_parser_errors(bad, msg):
    with pytest.raises(capi.HostError) as e:
        capi.InputData(text=bad)
    assert msg.lower() in str(e.value).lower()


def test_create_mesh_is_morton_ordered_and_colorized():
    m = capi.mesh_rectangle(2, [10, 10], 2).arrays
    assert m.n_cells == 16 and m.n_vertices == 25 and m.morton
    cent = m.xyz[m.cell_vertices].mean(axis=1)
    # children of the first coarse child come first, x is the least-significant bit
    exp = [(-3.75, -3.75), (-1.25, -3.75), (-3.75, -1.25), (-1.25, -1.25), (1.25, -3.75), (3.75, -3.75), (1.25, -1.25), (3.75, -1.25)]
    assert np.allclose(cent[:8], exp)
    # lexicographic vertex order inside a cell
    x = m.xyz[m.cell_vertices[5]]
    assert x[0][0] < x[1][0] and x[0][1] == x[1][1] and x[2][1] > x[0][1] and x[2][0] == x[0][0]
    # boundary ids = face numbers; 4 faces per side
    for f in range(4):
        sel = m.bface_id == f
        assert sel.sum() == 4 and np.all(m.bface_local[sel] == f)
        axis, side = f // 2, f % 2
        fc = cent[m.bface_cell[sel]][:, axis]
        assert np.all(fc < 0) if side == 0 else np.all(fc > 0)
    m3 = capi.mesh_rectangle(3, [10, 10, 10], 2).arrays
    assert m3.n_cells == 64 and len(m3.bface_cell) == 6 * 16
    c3 = m3.xyz[m3.cell_vertices].mean(axis=1)
    assert np.allclose(c3[1] - c3[0], [2.5, 0, 0]) and np.allclose(c3[2] - c3[0], [0, 2.5, 0]) and np.allclose(c3[4] - c3[0], [0, 0, 2.5])


def test_subdivided_box_any_cell_count():
    m = capi.mesh_subdivided(3, [10, 10, 10], [3, 5, 2]).arrays
    assert m.n_cells == 30 and m.n_vertices == 4 * 6 * 3 and not m.morton
    vol = 0.0
    for cv in m.cell_vertices:
        x = m.xyz[cv]
        vol += np.prod(x.max(axis=0) - x.min(axis=0))
    assert vol == pytest.approx(1000.0)


def test_first_touch_numbering_q1_and_q2():
    mesh = capi.mesh_rectangle(2, [10, 10], 2)
    d1 = capi.HostDofs(mesh, 1, 1)
    # cell 0 touches 0..3, cell 1 shares its left edge (vertices 1,3 of cell 0) and adds two, ...
    assert d1.n_dofs == 25
    assert list(d1.cell_dofs[0]) == [0, 1, 2, 3]
    assert list(d1.cell_dofs[1]) == [1, 4, 3, 5]
    assert list(d1.cell_dofs[2]) == [2, 3, 6, 7]
    assert list(d1.cell_dofs[3]) == [3, 5, 7, 8]
    dv = capi.HostDofs(mesh, 1, 2)  # FESystem(FE_Q(1), 2): component minor
    assert dv.n_dofs == 50 and list(dv.cell_dofs[0]) == [0, 1, 2, 3, 4, 5, 6, 7]
    assert list(dv.cell_dofs[1]) == [2, 3, 8, 9, 6, 7, 10, 11]
    d2 = capi.HostDofs(mesh, 2, 1)  # vertices, then lines (x=0, x=1, y=0, y=1), then interior
    assert d2.n_dofs == 81
    assert list(d2.cell_dofs[0]) == [0, 1, 2, 3, 4, 5, 6, 7, 8]
    # cell 1: vertices (1, new 9, 3, new 10); lines: left = cell 0's right line (5), then three new; interior new
    assert list(d2.cell_dofs[1]) == [1, 9, 3, 10, 5, 11, 12, 13, 14]
    sp = d2.support_points()
    assert np.allclose(sp[4], [-5.0, -3.75]) and np.allclose(sp[6], [-3.75, -5.0]) and np.allclose(sp[8], [-3.75, -3.75])


@pytest.mark.parametrize("dim,refine,deg,n_u,n_p", [(2, 4, 2, 2178, 289), (2, 4, 1, 578, 289), (3, 3, 1, 2187, 729), (3, 2, 2, 2187, 125)])
def test_dof_counts(dim, refine, deg, n_u, n_p):
    mesh = capi.mesh_rectangle(dim, [10] * dim, refine)
    assert capi.HostDofs(mesh, deg, dim).n_dofs == n_u
    assert capi.HostDofs(mesh, 1, 1).n_dofs == n_p


def test_dirichlet_first_condition_wins():
    mesh = capi.mesh_rectangle(2, [10, 10], 2)
    du = capi.HostDofs(mesh, 1, 2)
    sp = du.support_points()
    # same component on two adjacent faces: the corner keeps the value of the first condition (DS:123-134)
    ld, g = capi.make_dirichlet(mesh, du, [0, 2], [0, 0], [1.0, 2.0])
    comp = np.zeros(du.n_dofs, int)
    comp[du.cell_dofs[:, 1::2].ravel()] = 1
    assert np.all(comp[ld] == 0)
    corner = [i for i, d in enumerate(ld) if np.allclose(sp[d], [-5, -5])]
    assert len(corner) == 1 and g[corner[0]] == 1.0
    on_y = [i for i, d in enumerate(ld) if sp[d][1] == -5 and sp[d][0] > -5]
    assert len(on_y) == 4 and np.all(g[on_y] == 2.0)
    assert np.all(np.diff(ld) > 0)  # ConstraintMatrix::close sorts the lines


def test_read_msh_fixture():
    m = capi.mesh_read_msh(H.ROOT / "tests" / "golden" / "square10.msh", 2).arrays
    assert m.n_cells == 100 and m.n_vertices == 121
    # counter-clockwise gmsh quads become lexicographic cells with positive Jacobian
    for cv in m.cell_vertices[:10]:
        x = m.xyz[cv]
        assert _cross2(x[1] - x[0], x[2] - x[0]) > 0 and np.allclose(x[3] - x[2], x[1] - x[0], atol=1e-9)
    assert len(m.bface_cell) == 40
    cent = m.xyz[m.cell_vertices].mean(axis=1)
    for label, (axis, sign) in {0: (1, -1), 1: (0, 1), 2: (1, 1), 3: (0, -1)}.items():  # domain.geo:22-25
        sel = m.bface_id == label
        assert sel.sum() == 10
        assert np.all(np.sign(cent[m.bface_cell[sel]][:, axis]) == sign)


@pytest.mark.skipif(not (REF / "domain.msh").exists(), reason="reference tree not mounted")
def test_read_shipped_domain_msh():
    m = capi.mesh_read_msh(REF / "domain.msh", 2).arrays
    assert m.n_cells == 100 and m.n_vertices == 121 and len(m.bface_cell) == 40
    assert sorted(set(m.bface_id.tolist())) == [0, 1, 2, 3]
    area = 0.0
    for cv in m.cell_vertices:
        x = m.xyz[cv]
        area += 0.5 * abs(_cross2(x[3] - x[0], x[2] - x[1]))
    assert area == pytest.approx(100.0, rel=1e-10)
    du = capi.HostDofs(capi.mesh_read_msh(REF / "domain.msh", 2), 2, 2)
    assert du.n_dofs == 882  # SURVEY §8 C1'
