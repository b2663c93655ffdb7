"""Host-side mirror of the reference's solver interface in Python.

``setup_problem`` performs what ``PoroElasticProblem::run`` does before the time loop
(lib/include/PoroelasticityFSS.h:297-317) and ``time_step`` is one pass of FSS:328-407, both written
against :class:`capi.OperatorBackend` so the same call sequence can drive the CUDA library (``pe_*``)
and, in the tests, the CPU oracle (``po_*``).  The production driver is the C++
``PoroElasticProblem`` in csrc/host/problem.hpp; this mirror exists so parity tests read like the
reference's own loop.
"""
from __future__ import annotations

import numpy as np

from . import capi

TENSOR_TO_ENTRY = {2: [0, 1, 1, 2], 3: [0, 1, 2, 1, 3, 4, 2, 4, 5]}  # TensorIndexer.h:25-30
VOLUMETRIC_COMPONENTS = {2: [0, 3], 3: [0, 4, 8]}  # FSS:100-110
SHEAR_COMPONENTS = {2: [1], 3: [1, 2, 5]}


def make_mesh(inp: capi.InputData, mesh_file=None):
    """create_mesh() (FSS:418-435) or read_mesh() (FSS:438-445)."""
    if mesh_file is not None or inp.mesh_from_file:
        return capi.mesh_read_msh(mesh_file or "domain.msh", inp.dim)
    if inp.cells_per_axis[0] > 0:
        return capi.mesh_subdivided(inp.dim, inp.domain_size[: inp.dim], inp.cells_per_axis[: inp.dim])
    return capi.mesh_rectangle(inp.dim, inp.domain_size[: inp.dim], inp.initial_refinement_level)


def upload_problem(backend: capi.OperatorBackend, inp: capi.InputData, mesh: capi.HostMesh, prm: capi.PeParams | None = None,
                   forest: capi.Forest | None = None):
    """setup_dofs() (FSS:131-151) + set_boundary_conditions (FSS:300-306) for a single-rank backend.

    With ``forest`` the mesh is the forest's active mesh and both handlers get their hanging-node constraints
    (PS:74-77, DS:112-114); the displacement table also holds the Dirichlet lines and is closed (DS:117-137)."""
    prm = prm or inp.params()
    dofs_p = capi.HostDofs(mesh, 1, 1)
    dofs_u = capi.HostDofs(mesh, prm.degree_u, inp.dim)
    if forest is not None:
        Lp = capi.make_constraints(forest, mesh, dofs_p)
        Lu = capi.make_constraints(forest, mesh, dofs_u, inp.displacement_boundary_labels, inp.displacement_boundary_components,
                                   inp.displacement_boundary_values)
        backend.set_params(prm)
        backend.upload_mesh(mesh.arrays)
        backend.upload_dofs(capi.FIELD_PRESSURE, dofs_p.n_dofs, dofs_p.cell_dofs)
        backend.upload_dofs(capi.FIELD_DISPLACEMENT, dofs_u.n_dofs, dofs_u.cell_dofs)
        backend.upload_constraint_lines(capi.FIELD_PRESSURE, Lp)
        backend.upload_constraint_lines(capi.FIELD_DISPLACEMENT, Lu)
        backend.upload_neumann(inp.stress_boundary_labels, inp.stress_boundary_components, inp.stress_boundary_values)
        backend.setup()
        return dofs_p, dofs_u, (Lp, Lu)
    line_dof, g = capi.make_dirichlet(mesh, dofs_u, inp.displacement_boundary_labels, inp.displacement_boundary_components,
                                      inp.displacement_boundary_values)
    backend.set_params(prm)
    backend.upload_mesh(mesh.arrays)
    backend.upload_dofs(capi.FIELD_PRESSURE, dofs_p.n_dofs, dofs_p.cell_dofs)
    backend.upload_dofs(capi.FIELD_DISPLACEMENT, dofs_u.n_dofs, dofs_u.cell_dofs)
    backend.upload_constraints(capi.FIELD_DISPLACEMENT, line_dof, g)
    backend.upload_neumann(inp.stress_boundary_labels, inp.stress_boundary_components, inp.stress_boundary_values)
    backend.setup()
    return dofs_p, dofs_u, (line_dof, g)


def get_normal_strain_components(b: capi.OperatorBackend, dim):  # FSS:153-164
    comps = VOLUMETRIC_COMPONENTS[dim]
    b.project_assemble_rhs(comps)
    return sum(b.project_solve(TENSOR_TO_ENTRY[dim][c]) for c in comps)


def get_volumetric_strain(b: capi.OperatorBackend, dim, as_initial=False):  # FSS:179-186 (+317)
    b.volumetric_strain_from_projection([TENSOR_TO_ENTRY[dim][c] for c in VOLUMETRIC_COMPONENTS[dim]], as_initial)


def initialize(b: capi.OperatorBackend, inp: capi.InputData):
    """FSS:310-317."""
    b.pressure_set_uniform(inp.p_init)
    b.displacement_assemble()
    its, res = b.displacement_solve()
    b.project_assemble_matrix()
    pits = get_normal_strain_components(b, inp.dim)
    get_volumetric_strain(b, inp.dim, as_initial=True)
    return {"cg_its_displacement": its, "residual": res, "cg_its_projection": pits}


def time_step(b: capi.OperatorBackend, inp: capi.InputData, log=None):
    """One pass of FSS:328-407 (as-is semantics)."""
    dt = inp.time_step
    rep = {"fss_iterations": 0, "pressure_iterations": 0, "cg_its_pressure": 0, "cg_its_displacement": 0, "cg_its_projection": 0,
           "inner_counts": [], "residual_history": []}
    b.pressure_begin_step()  # FSS:342
    pressure_error = inp.pressure_tol * 2  # FSS:345
    fss_iteration = 0
    while fss_iteration < inp.max_fss_iterations and pressure_error > inp.fss_tol:
        fss_iteration += 1
        pressure_iteration = 0
        b.pressure_zero_update()  # FSS:356
        while pressure_iteration < inp.max_pressure_iterations:
            pressure_iteration += 1
            rep["pressure_iterations"] += 1
            b.update_volumetric_strain()
            pressure_error = b.assemble_residual(dt)
            rep["residual_history"].append(pressure_error)
            if pressure_error < inp.pressure_tol:
                break
            b.assemble_jacobian(dt)
            its, _ = b.pressure_solve()
            rep["cg_its_pressure"] += its
            b.pressure_add_update()  # FSS:379
        rep["inner_counts"].append(pressure_iteration)
        rep["pressure_linfty"] = b.pressure_linfty()
        b.displacement_assemble()  # FSS:395
        its, _ = b.displacement_solve()
        rep["cg_its_displacement"] += its
        rep["cg_its_projection"] += get_normal_strain_components(b, inp.dim)  # FSS:398
        if inp.couple_volumetric_strain:
            get_volumetric_strain(b, inp.dim)  # FSS:399, commented out in the reference
        pressure_error = b.assemble_residual(dt)  # FSS:402-405
        if log:
            log(f"        Error: {pressure_error:g}")
    rep["fss_iterations"] = fss_iteration
    rep["pressure_error"] = pressure_error
    return rep


def rel_l2(a, b):
    a = np.asarray(a)
    b = np.asarray(b)
    nb = np.linalg.norm(b)
    return float(np.linalg.norm(a - b) / (nb if nb > 0 else 1.0))
