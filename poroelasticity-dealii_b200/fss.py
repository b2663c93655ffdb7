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


def get_normal_strain_components(b: capi.OperatorBackend, dim, each=None):  # FSS:153-164
    """returns the CG iterations of the dim projection solves summed; `each` (a list) also receives them one by one"""
    comps = VOLUMETRIC_COMPONENTS[dim]
    b.project_assemble_rhs(comps)
    its = [b.project_solve(TENSOR_TO_ENTRY[dim][c]) for c in comps]
    if each is not None:
        each.extend(its)
    return sum(its)


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
           "inner_counts": [], "residual_history": [], "cg_each": {"pressure": [], "displacement": [], "projection": []}}
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
            rep["cg_each"]["pressure"].append(its)
            b.pressure_add_update()  # FSS:379
        rep["inner_counts"].append(pressure_iteration)
        rep["pressure_linfty"] = b.pressure_linfty()
        b.displacement_assemble()  # FSS:395
        its, res = b.displacement_solve()
        rep["cg_its_displacement"] += its
        rep["cg_each"]["displacement"].append(its)
        rep["displacement_residual"] = res
        rep["cg_its_projection"] += get_normal_strain_components(b, inp.dim, rep["cg_each"]["projection"])  # FSS:398
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


# ---- adaptive time loop (FSS:333-340, refine_mesh FSS:447-498) ---------------------------------------------------------
def make_forest(inp: capi.InputData, mesh: capi.HostMesh):
    """The initial mesh's cells become the roots of the refinement forest (csrc/host/problem.hpp::initialize)."""
    refined_box = not inp.mesh_from_file and inp.cells_per_axis[0] <= 0
    return capi.Forest(mesh, inp.initial_refinement_level if refined_box else 0)


def refine_mesh(b: capi.OperatorBackend, inp: capi.InputData, forest: capi.Forest, mesh: capi.HostMesh, dofs_p: capi.HostDofs,
                top_fraction=0.6, bottom_fraction=0.4):
    """refine_mesh(initial, initial + max) of FSS:335-337 followed by setup_dofs and the solution transfer; returns the new
    (mesh, dofs_p, dofs_u, report).  The caller re-assembles (FSS:338-339)."""
    p, ev, ev0 = b.get_vector(capi.VEC_P), b.get_vector(capi.VEC_VOL_STRAIN), b.get_vector(capi.VEC_VOL_STRAIN0)
    eta = forest.kelly(mesh, dofs_p, p)  # FSS:452-458
    forest.mark_fixed_fraction(eta, top_fraction, bottom_fraction, inp.initial_refinement_level,
                               inp.initial_refinement_level + inp.max_refinement_level)  # FSS:460-472
    forest.store(mesh, dofs_p, [p, ev, ev0])  # FSS:475-479
    n_coarsened, n_refined = forest.execute()  # FSS:481-483
    new_mesh = forest.active_mesh()
    new_p, new_u, (Lp, Lu) = upload_problem(b, inp, new_mesh, forest=forest)  # setup_dofs(), FSS:485
    vals = forest.fetch(new_mesh, new_p, 3)  # FSS:488-497
    b.set_vector(capi.VEC_P, vals[0])
    b.set_vector(capi.VEC_VOL_STRAIN, vals[1])
    b.set_vector(capi.VEC_VOL_STRAIN0, vals[2])
    report = {"n_cells": new_mesh.arrays.n_cells, "n_dofs_p": new_p.n_dofs, "n_dofs_u": new_u.n_dofs, "n_hanging_p": Lp.n_lines,
              "n_coarsened_families": n_coarsened, "n_refined_cells": n_refined, "eta_max": float(eta.max()), "levels": forest.levels()}
    return new_mesh, new_p, new_u, report


def run_adaptive(b: capi.OperatorBackend, inp: capi.InputData, n_steps, refine_every, on_step=None):
    """PoroElasticProblem::run (FSS:294-415) with the reference's every-n-th-step refinement; returns the final
    (forest, mesh, dofs_p, dofs_u) and the per-step reports."""
    mesh0 = make_mesh(inp)
    forest = make_forest(inp, mesh0)
    mesh = forest.active_mesh()
    dofs_p, dofs_u, _ = upload_problem(b, inp, mesh, forest=forest)
    initialize(b, inp)
    reports = []
    for step in range(1, n_steps + 1):
        amr = None
        if refine_every > 0 and step % refine_every == 0:  # FSS:333
            mesh, dofs_p, dofs_u, amr = refine_mesh(b, inp, forest, mesh, dofs_p)
            b.displacement_assemble()    # FSS:338
            b.project_assemble_matrix()  # FSS:339
        rep = time_step(b, inp)
        rep["amr"] = amr
        reports.append(rep)
        if on_step:
            on_step(step, rep, mesh, dofs_p, dofs_u)
    return forest, mesh, dofs_p, dofs_u, reports
