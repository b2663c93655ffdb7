"""ctypes bindings of the two in-tree shared libraries.

* ``libporoel.so``      — the CUDA (sm_100a) device library, ABI in ``include/poroel.h``
* ``libporoel_host.so`` — the host C++ surface (input parser, mesh, dofs, partition, the
  ``PoroElasticProblem`` driver), ABI in ``include/poroel_host.h``

Nothing here computes: it only marshals numpy arrays to the C-ABI.  If the CUDA library is missing
the import of :func:`load_device` raises — there is no CPU fallback in the product.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
LIB_DIR = ROOT / "lib"

# ---- enums mirrored from include/poroel.h -------------------------------------------------------
PE_OK, PE_ERR_CUDA, PE_ERR_NCCL, PE_ERR_NO_CONVERGENCE, PE_ERR_NAN, PE_ERR_BAD_INPUT, PE_ERR_UNSUPPORTED, PE_ERR_STATE = 0, -1, -2, -3, -4, -5, -6, -7
FIELD_PRESSURE, FIELD_DISPLACEMENT = 0, 1
PRECOND_JACOBI, PRECOND_CHEBYSHEV = 0, 1
VEC_P, VEC_P_OLD, VEC_P_UPDATE, VEC_P_RESIDUAL, VEC_VOL_STRAIN, VEC_VOL_STRAIN0, VEC_WELL_RHS, VEC_U, VEC_U_RHS = range(9)
VEC_STRAIN0, VEC_PROJ_RHS0, VEC_STRESS0 = 16, 32, 48
MAT_MASS, MAT_LAPLACE, MAT_JACOBIAN, MAT_ELASTICITY, MAT_PROJECTION = 0, 1, 2, 3, 4

i32p = C.POINTER(C.c_int32)
i64p = C.POINTER(C.c_int64)
i8p = C.POINTER(C.c_int8)
f64p = C.POINTER(C.c_double)


class PeParams(C.Structure):
    _fields_ = [
        ("dim", C.c_int32), ("degree_u", C.c_int32), ("degree_p", C.c_int32), ("preconditioner", C.c_int32),
        ("chebyshev_degree", C.c_int32), ("cg_max_iterations", C.c_int32), ("cg_check_interval", C.c_int32), ("reserved0", C.c_int32),
        ("lame_lambda", C.c_double), ("shear_modulus", C.c_double), ("bulk_modulus", C.c_double), ("biot_coef", C.c_double),
        ("m_modulus", C.c_double), ("perm_over_visc", C.c_double), ("well_radius", C.c_double), ("flow_rate", C.c_double),
        ("cg_rel_tol_pressure", C.c_double), ("cg_abs_tol_displacement", C.c_double), ("cg_rel_tol_projection", C.c_double),
        ("chebyshev_eig_ratio", C.c_double),
    ]


class PeStats(C.Structure):
    _fields_ = [
        ("n_cells", C.c_int64), ("n_dofs_p", C.c_int64), ("n_dofs_u", C.c_int64), ("nnz_p", C.c_int64), ("nnz_u", C.c_int64),
        ("cg_iterations_pressure", C.c_int64), ("cg_iterations_displacement", C.c_int64), ("cg_iterations_projection", C.c_int64),
        ("cg_solves_pressure", C.c_int64), ("cg_solves_displacement", C.c_int64), ("cg_solves_projection", C.c_int64),
        ("spmv_launches_p", C.c_int64), ("spmv_launches_u", C.c_int64), ("kernel_launches", C.c_int64),
        ("spmv_bytes_p", C.c_double), ("spmv_bytes_u", C.c_double),
        ("eig_max_p", C.c_double), ("eig_max_u", C.c_double), ("eig_max_m", C.c_double), ("setup_ms", C.c_double),
        ("spmv_ms_p", C.c_double), ("spmv_ms_u", C.c_double), ("spmv_timed_p", C.c_int64), ("spmv_timed_u", C.c_int64),
        ("pcg_ms_p", C.c_double), ("pcg_ms_u", C.c_double), ("pcg_iterations_p", C.c_int64), ("pcg_iterations_u", C.c_int64),
        ("bsr_block_size", C.c_int64),
        ("inner_ms_u", C.c_double), ("inner_passes_u", C.c_int64), ("inner_bytes_u", C.c_double), ("update_ms_u", C.c_double),
        ("reduce_ms_u", C.c_double), ("sell_format_u", C.c_int64),
        ("wait_inner_ms_u", C.c_double), ("wait_cg_ms_u", C.c_double), ("wait_peer_ms_u", C.c_double), ("wait_update_ms_u", C.c_double),
        ("phase_ms_p", C.c_double * 10),
    ]

    def as_dict(self):
        return {k: (list(getattr(self, k)) if k == "phase_ms_p" else getattr(self, k)) for k, _ in self._fields_}


class InputView(C.Structure):
    _fields_ = [
        ("dim", C.c_int32), ("initial_refinement_level", C.c_int32), ("max_refinement_level", C.c_int32),
        ("max_fss_iterations", C.c_int32), ("max_pressure_iterations", C.c_int32),
        ("displacement_degree", C.c_int32), ("preconditioner", C.c_int32), ("chebyshev_degree", C.c_int32), ("cg_max_iterations", C.c_int32),
        ("mesh_from_file", C.c_int32), ("refine_every", C.c_int32), ("couple_volumetric_strain", C.c_int32), ("write_vtk", C.c_int32),
        ("max_time_steps", C.c_int32), ("cells_per_axis", C.c_int32 * 3), ("n_dirichlet", C.c_int32), ("n_neumann", C.c_int32),
        ("domain_size", C.c_double * 3),
        ("perm", C.c_double), ("poro", C.c_double), ("visc", C.c_double), ("f_comp", C.c_double), ("youngs_modulus", C.c_double),
        ("poisson_ratio", C.c_double), ("biot_coef", C.c_double), ("bulk_density", C.c_double), ("r_well", C.c_double), ("flow_rate", C.c_double),
        ("time_step", C.c_double), ("t_max", C.c_double), ("fss_tol", C.c_double), ("pressure_tol", C.c_double), ("p_init", C.c_double),
        ("lame_constant", C.c_double), ("shear_modulus", C.c_double), ("bulk_modulus", C.c_double), ("grain_bulk_modulus", C.c_double),
        ("n_modulus", C.c_double), ("m_modulus", C.c_double), ("chebyshev_eig_ratio", C.c_double),
        ("dirichlet_labels", i32p), ("dirichlet_components", i32p), ("dirichlet_values", f64p),
        ("neumann_labels", i32p), ("neumann_components", i32p), ("neumann_values", f64p),
    ]


class MeshView(C.Structure):
    _fields_ = [
        ("dim", C.c_int32), ("morton", C.c_int32), ("n_vertices", C.c_int64), ("n_cells", C.c_int64), ("n_bfaces", C.c_int64),
        ("xyz", f64p), ("cell_vertices", i32p), ("bface_cell", i32p), ("bface_local", i8p), ("bface_id", i32p),
    ]


class DofsView(C.Structure):
    _fields_ = [("degree", C.c_int32), ("n_comp", C.c_int32), ("n_loc", C.c_int32), ("reserved", C.c_int32), ("n_dofs", C.c_int64), ("cell_dofs", i32p)]


class ConstraintsView(C.Structure):
    _fields_ = [("n_lines", C.c_int64), ("n_entries", C.c_int64), ("line_dof", i32p), ("entry_ptr", i64p), ("entry_dof", i32p),
                ("entry_w", f64p), ("inhomogeneity", f64p)]


class PartFieldView(C.Structure):
    _fields_ = [
        ("n_owned", C.c_int64), ("n_local", C.c_int64), ("n_neighbors", C.c_int32), ("reserved", C.c_int32),
        ("cell_dofs", i32p), ("local_to_global", i64p), ("neighbor_rank", i32p), ("send_ptr", i64p), ("send_idx", i32p), ("recv_ptr", i64p),
    ]


class PartView(C.Structure):
    _fields_ = [("mesh", MeshView), ("cell_global", i64p), ("n_owned_cells", C.c_int64), ("field", PartFieldView * 2)]


class StepReport(C.Structure):
    _fields_ = [
        ("time", C.c_double), ("time_step_number", C.c_int32), ("fss_iterations", C.c_int32), ("pressure_iterations", C.c_int32),
        ("cg_its_pressure", C.c_int32), ("cg_its_displacement", C.c_int32), ("cg_its_projection", C.c_int32), ("status", C.c_int32),
        ("pressure_error", C.c_double), ("pressure_linfty", C.c_double),
    ]

    def as_dict(self):
        return {k: (list(getattr(self, k)) if k == "phase_ms_p" else getattr(self, k)) for k, _ in self._fields_}


# symbols declared in include/poroel.h (checked by tests/test_abi.py)
DEVICE_SYMBOLS = [
    "pe_nccl_unique_id", "pe_create", "pe_destroy", "pe_last_error", "pe_version", "pe_set_params", "pe_upload_mesh", "pe_upload_dofs",
    "pe_upload_constraints", "pe_upload_neumann", "pe_upload_partition", "pe_setup", "pe_pressure_set_uniform", "pe_pressure_begin_step",
    "pe_pressure_zero_update", "pe_pressure_update_volumetric_strain", "pe_pressure_assemble_residual", "pe_pressure_assemble_jacobian",
    "pe_pressure_solve", "pe_pressure_add_update", "pe_pressure_linfty", "pe_displacement_assemble", "pe_displacement_solve",
    "pe_project_assemble_matrix", "pe_project_assemble_rhs", "pe_project_solve", "pe_volumetric_strain_from_projection",
    "pe_effective_stresses", "pe_spmv", "pe_get_vector", "pe_set_vector", "pe_get_matrix_size", "pe_get_matrix", "pe_get_stats",
    "pe_reset_stats", "pe_synchronize", "pe_stream", "pe_set_profiling",
]
HOST_SYMBOLS = [
    "peh_last_error", "peh_input_create", "peh_input_destroy", "peh_input_read_file", "peh_input_read_string", "peh_input_view_get",
    "peh_input_to_params", "peh_mesh_create_rectangle", "peh_mesh_create_subdivided", "peh_mesh_read_msh", "peh_mesh_reorder_sfc", "peh_mesh_permute_cells", "peh_mesh_destroy",
    "peh_mesh_view_get", "peh_dofs_distribute", "peh_dofs_destroy", "peh_dofs_view_get", "peh_dofs_support_points", "peh_make_dirichlet",
    "peh_forest_create", "peh_forest_destroy", "peh_forest_active_mesh", "peh_forest_active_levels", "peh_forest_set_flags",
    "peh_forest_get_flags", "peh_forest_prepare", "peh_forest_execute", "peh_forest_kelly", "peh_forest_mark_fixed_fraction",
    "peh_forest_store", "peh_forest_fetch", "peh_constraints_make", "peh_constraints_destroy", "peh_constraints_view_get",
    "peh_partition", "peh_partition_structured", "peh_part_destroy", "peh_part_view_get", "peh_problem_create", "peh_problem_destroy", "peh_problem_initialize",
    "peh_problem_step", "peh_problem_run", "peh_problem_ctx", "peh_problem_mesh", "peh_problem_global_ids",
]

_dev = None
_host = None


def _declare_operator_api(lib, prefix):
    """Signatures shared by the device library (pe_*) and the oracle (po_*)."""
    P = C.c_void_p
    g = lambda n: getattr(lib, prefix + n)
    g("set_params").argtypes = [P, C.POINTER(PeParams)]
    g("upload_mesh").argtypes = [P, C.c_int, C.c_int64, f64p, C.c_int64, i32p, C.c_int64, i32p, i8p, i32p]
    g("upload_dofs").argtypes = [P, C.c_int, C.c_int64, i32p]
    g("upload_constraints").argtypes = [P, C.c_int, C.c_int64, i32p, i64p, i32p, f64p, f64p]
    g("upload_neumann").argtypes = [P, C.c_int, i32p, i32p, f64p]
    g("setup").argtypes = [P]
    g("pressure_set_uniform").argtypes = [P, C.c_double]
    for n in ("pressure_begin_step", "pressure_zero_update", "pressure_update_volumetric_strain", "pressure_add_update",
              "displacement_assemble", "project_assemble_matrix", "effective_stresses"):
        g(n).argtypes = [P]
    g("pressure_assemble_residual").argtypes = [P, C.c_double, f64p]
    g("pressure_assemble_jacobian").argtypes = [P, C.c_double]
    g("pressure_solve").argtypes = [P, C.POINTER(C.c_int), f64p]
    g("pressure_linfty").argtypes = [P, f64p]
    g("displacement_solve").argtypes = [P, C.POINTER(C.c_int), f64p]
    g("project_assemble_rhs").argtypes = [P, C.c_int, i32p]
    g("project_solve").argtypes = [P, C.c_int, C.POINTER(C.c_int)]
    g("volumetric_strain_from_projection").argtypes = [P, C.c_int, i32p, C.c_int]
    g("get_vector").argtypes = [P, C.c_int, f64p, C.c_int64]
    g("set_vector").argtypes = [P, C.c_int, f64p, C.c_int64]
    g("get_matrix_size").argtypes = [P, C.c_int, i64p, i64p]
    g("get_matrix").argtypes = [P, C.c_int, i64p, i32p, f64p]
    g("get_stats").argtypes = [P, C.POINTER(PeStats)]
    g("reset_stats").argtypes = [P]
    g("last_error").argtypes = [P]
    g("last_error").restype = C.c_char_p
    g("destroy").argtypes = [P]
    g("destroy").restype = None


def _preload_nccl():
    """libporoel.so needs `libnccl.so.2`.  PyTorch bundles a newer NCCL under the same soname; whichever is
    loaded first wins for the whole process, so load torch's copy first (if there is one) to keep a later
    `import torch` working.  Without the wheel the system library is used."""
    import importlib.util
    try:
        spec = importlib.util.find_spec("nvidia.nccl")
    except (ImportError, ValueError):
        spec = None
    if spec and spec.submodule_search_locations:
        cand = Path(list(spec.submodule_search_locations)[0]) / "lib" / "libnccl.so.2"
        if cand.exists():
            C.CDLL(str(cand), mode=C.RTLD_GLOBAL)


def load_device():
    """dlopen libporoel.so; raises (loudly) when the CUDA extension has not been built."""
    global _dev
    if _dev is not None:
        return _dev
    path = LIB_DIR / "libporoel.so"
    if not path.exists():
        raise RuntimeError(f"{path} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(the product has no CPU fallback)")
    _preload_nccl()
    lib = C.CDLL(str(path))
    _declare_operator_api(lib, "pe_")
    P = C.c_void_p
    lib.pe_create.argtypes = [C.POINTER(P), C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_size_t]
    lib.pe_nccl_unique_id.argtypes = [C.c_void_p, C.POINTER(C.c_size_t)]
    lib.pe_upload_partition.argtypes = [P, C.c_int, C.c_int64, C.c_int, i32p, i64p, i32p, i64p]
    lib.pe_spmv.argtypes = [P, C.c_int, f64p, f64p, C.c_int, C.POINTER(C.c_float)]
    lib.pe_synchronize.argtypes = [P]
    lib.pe_set_profiling.argtypes = [P, C.c_int]
    lib.pe_stream.argtypes = [P]
    lib.pe_stream.restype = C.c_void_p
    _dev = lib
    return lib


def _declare_host_api(lib):
    """ctypes signatures of include/poroel_host.h (also applied to the oracle-backed driver build of the tests)."""
    P = C.c_void_p
    lib.peh_last_error.restype = C.c_char_p
    lib.peh_input_create.restype = P
    lib.peh_input_destroy.argtypes = [P]
    lib.peh_input_read_file.argtypes = [P, C.c_char_p, C.c_int]
    lib.peh_input_read_string.argtypes = [P, C.c_char_p]
    lib.peh_input_view_get.argtypes = [P, C.POINTER(InputView)]
    lib.peh_input_to_params.argtypes = [P, C.POINTER(PeParams)]
    lib.peh_mesh_create_rectangle.argtypes = [C.c_int, f64p, C.c_int]
    lib.peh_mesh_create_rectangle.restype = P
    lib.peh_mesh_create_subdivided.argtypes = [C.c_int, f64p, i32p]
    lib.peh_mesh_create_subdivided.restype = P
    lib.peh_mesh_read_msh.argtypes = [C.c_char_p, C.c_int]
    lib.peh_mesh_read_msh.restype = P
    lib.peh_mesh_reorder_sfc.argtypes = [P, i64p]
    lib.peh_mesh_permute_cells.argtypes = [P, i64p]
    lib.peh_mesh_destroy.argtypes = [P]
    lib.peh_mesh_view_get.argtypes = [P, C.POINTER(MeshView)]
    lib.peh_dofs_distribute.argtypes = [P, C.c_int, C.c_int]
    lib.peh_dofs_distribute.restype = P
    lib.peh_dofs_destroy.argtypes = [P]
    lib.peh_dofs_view_get.argtypes = [P, C.POINTER(DofsView)]
    lib.peh_dofs_support_points.argtypes = [P, P, f64p]
    lib.peh_make_dirichlet.argtypes = [P, P, C.c_int, i32p, i32p, f64p, i32p, f64p]
    lib.peh_make_dirichlet.restype = C.c_int64
    f32p = C.POINTER(C.c_float)
    lib.peh_forest_create.argtypes = [P, C.c_int]
    lib.peh_forest_create.restype = P
    lib.peh_forest_destroy.argtypes = [P]
    lib.peh_forest_active_mesh.argtypes = [P]
    lib.peh_forest_active_mesh.restype = P
    lib.peh_forest_active_levels.argtypes = [P, i32p]
    lib.peh_forest_active_levels.restype = C.c_int64
    lib.peh_forest_set_flags.argtypes = [P, C.c_int64, i8p, i8p]
    lib.peh_forest_get_flags.argtypes = [P, C.c_int64, i8p, i8p]
    lib.peh_forest_prepare.argtypes = [P]
    lib.peh_forest_execute.argtypes = [P, i32p, i32p]
    lib.peh_forest_kelly.argtypes = [P, P, P, f64p, f32p]
    lib.peh_forest_mark_fixed_fraction.argtypes = [P, C.c_int64, f32p, C.c_double, C.c_double, C.c_int, C.c_int]
    lib.peh_forest_store.argtypes = [P, P, P, C.c_int, f64p]
    lib.peh_forest_fetch.argtypes = [P, P, P, C.c_int, f64p]
    lib.peh_constraints_make.argtypes = [P, P, P, C.c_int, i32p, i32p, f64p]
    lib.peh_constraints_make.restype = P
    lib.peh_constraints_destroy.argtypes = [P]
    lib.peh_constraints_view_get.argtypes = [P, C.POINTER(ConstraintsView)]
    lib.peh_partition.argtypes = [P, P, P, C.c_int, C.c_int]
    lib.peh_partition.restype = P
    lib.peh_partition_structured.argtypes = [C.c_int, f64p, i32p, C.c_int, C.c_int, C.c_int]
    lib.peh_partition_structured.restype = P
    lib.peh_part_destroy.argtypes = [P]
    lib.peh_part_view_get.argtypes = [P, C.POINTER(PartView)]
    lib.peh_problem_create.argtypes = [P, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_size_t]
    lib.peh_problem_create.restype = P
    lib.peh_problem_destroy.argtypes = [P]
    lib.peh_problem_initialize.argtypes = [P, C.c_int]
    lib.peh_problem_step.argtypes = [P, C.c_int, C.POINTER(StepReport)]
    lib.peh_problem_run.argtypes = [P, C.c_int]
    lib.peh_problem_ctx.argtypes = [P]
    lib.peh_problem_ctx.restype = P
    lib.peh_problem_mesh.argtypes = [P]
    lib.peh_problem_mesh.restype = P
    lib.peh_problem_global_ids.argtypes = [P, C.c_int, i64p]


def load_host():
    global _host
    if _host is not None:
        return _host
    load_device()  # libporoel_host.so links against libporoel.so
    path = LIB_DIR / "libporoel_host.so"
    if not path.exists():
        raise RuntimeError(f"{path} is missing: run __graft_entry__.build()")
    lib = C.CDLL(str(path))
    _declare_host_api(lib)
    _host = lib
    return lib


# ---- numpy helpers -------------------------------------------------------------------------------
def _p(a, ct):
    return a.ctypes.data_as(C.POINTER(ct)) if a is not None and a.size else C.cast(None, C.POINTER(ct))


def _np(ptr, n, dtype):
    if n == 0:
        return np.zeros(0, dtype=dtype)
    return np.ctypeslib.as_array(ptr, shape=(n,)).astype(dtype, copy=True)


class HostError(RuntimeError):
    pass


class Mesh:
    """Plain arrays of a mesh produced by the host library (FSS:418-445)."""

    def __init__(self, view: MeshView):
        d = view.dim
        self.dim = d
        self.morton = bool(view.morton)
        self.xyz = _np(view.xyz, view.n_vertices * d, np.float64).reshape(-1, d)
        self.cell_vertices = _np(view.cell_vertices, view.n_cells * (1 << d), np.int32).reshape(-1, 1 << d)
        self.bface_cell = _np(view.bface_cell, view.n_bfaces, np.int32)
        self.bface_local = _np(view.bface_local, view.n_bfaces, np.int8)
        self.bface_id = _np(view.bface_id, view.n_bfaces, np.int32)

    @property
    def n_cells(self):
        return self.cell_vertices.shape[0]

    @property
    def n_vertices(self):
        return self.xyz.shape[0]


class HostMesh:
    """Owning handle + numpy copy."""

    def __init__(self, handle):
        if not handle:
            raise HostError(load_host().peh_last_error().decode())
        self.h = handle
        v = MeshView()
        load_host().peh_mesh_view_get(handle, C.byref(v))
        self.arrays = Mesh(v)

    def _refresh(self):
        v = MeshView()
        load_host().peh_mesh_view_get(self.h, C.byref(v))
        self.arrays = Mesh(v)

    def reorder_sfc(self):
        """Morton order of the cell centroids (mesh.hpp::reorder_cells_sfc); returns the permutation new -> old."""
        perm = np.zeros(self.arrays.n_cells, dtype=np.int64)
        if load_host().peh_mesh_reorder_sfc(self.h, _p(perm, C.c_int64)) != 0:
            raise HostError(load_host().peh_last_error().decode())
        self._refresh()
        return perm

    def permute_cells(self, perm_new_to_old):
        perm = np.ascontiguousarray(perm_new_to_old, dtype=np.int64)
        if load_host().peh_mesh_permute_cells(self.h, _p(perm, C.c_int64)) != 0:
            raise HostError(load_host().peh_last_error().decode())
        self._refresh()

    def __del__(self):
        if getattr(self, "h", None):
            load_host().peh_mesh_destroy(self.h)
            self.h = None


def mesh_rectangle(dim, size, refine):
    s = np.asarray(size, dtype=np.float64)
    return HostMesh(load_host().peh_mesh_create_rectangle(dim, _p(s, C.c_double), refine))


def mesh_subdivided(dim, size, n):
    s = np.asarray(size, dtype=np.float64)
    nn = np.asarray(list(n) + [1] * (3 - len(n)), dtype=np.int32)
    return HostMesh(load_host().peh_mesh_create_subdivided(dim, _p(s, C.c_double), _p(nn, C.c_int32)))


def mesh_read_msh(path, dim):
    return HostMesh(load_host().peh_mesh_read_msh(str(path).encode(), dim))


class HostDofs:
    def __init__(self, mesh: HostMesh, degree, n_comp):
        lib = load_host()
        self.h = lib.peh_dofs_distribute(mesh.h, degree, n_comp)
        if not self.h:
            raise HostError(lib.peh_last_error().decode())
        v = DofsView()
        lib.peh_dofs_view_get(self.h, C.byref(v))
        self.degree, self.n_comp, self.n_loc, self.n_dofs = v.degree, v.n_comp, v.n_loc, v.n_dofs
        self.cell_dofs = _np(v.cell_dofs, mesh.arrays.n_cells * v.n_loc, np.int32).reshape(-1, v.n_loc)
        self._mesh = mesh

    def support_points(self):
        out = np.zeros((self.n_dofs, self._mesh.arrays.dim))
        load_host().peh_dofs_support_points(self._mesh.h, self.h, _p(out, C.c_double))
        return out

    def __del__(self):
        if getattr(self, "h", None):
            load_host().peh_dofs_destroy(self.h)
            self.h = None


def make_dirichlet(mesh: HostMesh, dofs: HostDofs, labels, comps, values):
    lib = load_host()
    l = np.asarray(labels, dtype=np.int32)
    c = np.asarray(comps, dtype=np.int32)
    v = np.asarray(values, dtype=np.float64)
    n = lib.peh_make_dirichlet(mesh.h, dofs.h, len(l), _p(l, C.c_int32), _p(c, C.c_int32), _p(v, C.c_double), None, None)
    if n < 0:
        raise HostError(lib.peh_last_error().decode())
    ld = np.zeros(n, dtype=np.int32)
    g = np.zeros(n, dtype=np.float64)
    lib.peh_make_dirichlet(mesh.h, dofs.h, len(l), _p(l, C.c_int32), _p(c, C.c_int32), _p(v, C.c_double), _p(ld, C.c_int32), _p(g, C.c_double))
    return ld, g


class ConstraintLines:
    """Closed constraint table x_i = sum_j w_ij x_j + g_i in the flat form of pe_upload_constraints."""

    def __init__(self, line_dof, entry_ptr, entry_dof, entry_w, inhomogeneity):
        self.line_dof = np.asarray(line_dof, dtype=np.int32)
        self.entry_ptr = np.asarray(entry_ptr, dtype=np.int64)
        self.entry_dof = np.asarray(entry_dof, dtype=np.int32)
        self.entry_w = np.asarray(entry_w, dtype=np.float64)
        self.inhomogeneity = np.asarray(inhomogeneity, dtype=np.float64)

    @property
    def n_lines(self):
        return len(self.line_dof)

    def line(self, i):
        a, b = self.entry_ptr[i], self.entry_ptr[i + 1]
        return self.line_dof[i], self.entry_dof[a:b], self.entry_w[a:b], self.inhomogeneity[i]


def make_constraints(forest, mesh: "HostMesh", dofs: "HostDofs", labels=(), comps=(), values=()) -> ConstraintLines:
    """make_hanging_node_constraints (when `forest` is given) + Dirichlet lines + close (PS:71-78, DS:109-137)."""
    lib = load_host()
    l = np.asarray(labels, dtype=np.int32)
    c = np.asarray(comps, dtype=np.int32)
    v = np.asarray(values, dtype=np.float64)
    h = lib.peh_constraints_make(forest.h if forest is not None else None, mesh.h, dofs.h, len(l), _p(l, C.c_int32), _p(c, C.c_int32), _p(v, C.c_double))
    if not h:
        raise HostError(lib.peh_last_error().decode())
    view = ConstraintsView()
    lib.peh_constraints_view_get(h, C.byref(view))
    out = ConstraintLines(_np(view.line_dof, view.n_lines, np.int32), _np(view.entry_ptr, view.n_lines + 1, np.int64),
                          _np(view.entry_dof, view.n_entries, np.int32), _np(view.entry_w, view.n_entries, np.float64),
                          _np(view.inhomogeneity, view.n_lines, np.float64))
    lib.peh_constraints_destroy(h)
    return out


class Forest:
    """Refinement forest of the adaptive time loop (FSS:333-340, 447-498; csrc/host/amr.hpp)."""

    def __init__(self, mesh: "HostMesh", base_level=0):
        lib = load_host()
        self.lib = lib
        self.h = lib.peh_forest_create(mesh.h, base_level)
        if not self.h:
            raise HostError(lib.peh_last_error().decode())

    def _ck(self, rc):
        if rc != 0:
            raise HostError(self.lib.peh_last_error().decode())

    def active_mesh(self) -> "HostMesh":
        return HostMesh(self.lib.peh_forest_active_mesh(self.h))

    def levels(self):
        n = self.lib.peh_forest_active_levels(self.h, None)
        out = np.zeros(n, dtype=np.int32)
        self.lib.peh_forest_active_levels(self.h, _p(out, C.c_int32))
        return out

    def set_flags(self, refine=None, coarsen=None):
        n = self.lib.peh_forest_active_levels(self.h, None)
        r = np.zeros(n, dtype=np.int8) if refine is None else np.ascontiguousarray(refine, dtype=np.int8)
        c = np.zeros(n, dtype=np.int8) if coarsen is None else np.ascontiguousarray(coarsen, dtype=np.int8)
        self._ck(self.lib.peh_forest_set_flags(self.h, n, r.ctypes.data_as(i8p), c.ctypes.data_as(i8p)))

    def get_flags(self):
        n = self.lib.peh_forest_active_levels(self.h, None)
        r = np.zeros(n, dtype=np.int8)
        c = np.zeros(n, dtype=np.int8)
        self._ck(self.lib.peh_forest_get_flags(self.h, n, r.ctypes.data_as(i8p), c.ctypes.data_as(i8p)))
        return r.astype(bool), c.astype(bool)

    def prepare(self):
        self._ck(self.lib.peh_forest_prepare(self.h))

    def execute(self):
        nc, nr = C.c_int32(), C.c_int32()
        self._ck(self.lib.peh_forest_execute(self.h, C.byref(nc), C.byref(nr)))
        return nc.value, nr.value

    def kelly(self, mesh: "HostMesh", dofs_p: "HostDofs", p):
        pp = np.ascontiguousarray(p, dtype=np.float64)
        eta = np.zeros(mesh.arrays.n_cells, dtype=np.float32)
        self._ck(self.lib.peh_forest_kelly(self.h, mesh.h, dofs_p.h, _p(pp, C.c_double), eta.ctypes.data_as(C.POINTER(C.c_float))))
        return eta

    def mark_fixed_fraction(self, criteria, top, bottom, min_level, max_level):
        cr = np.ascontiguousarray(criteria, dtype=np.float32)
        self._ck(self.lib.peh_forest_mark_fixed_fraction(self.h, len(cr), cr.ctypes.data_as(C.POINTER(C.c_float)), top, bottom, min_level, max_level))

    def store(self, mesh: "HostMesh", dofs_p: "HostDofs", vectors):
        v = np.ascontiguousarray(np.stack([np.asarray(x, dtype=np.float64) for x in vectors]))
        self._ck(self.lib.peh_forest_store(self.h, mesh.h, dofs_p.h, v.shape[0], _p(v, C.c_double)))

    def fetch(self, mesh: "HostMesh", dofs_p: "HostDofs", n_vec):
        out = np.zeros((n_vec, dofs_p.n_dofs))
        self._ck(self.lib.peh_forest_fetch(self.h, mesh.h, dofs_p.h, n_vec, _p(out, C.c_double)))
        return out

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.peh_forest_destroy(self.h)
            self.h = None


class InputData:
    """InputDataPoroel (ID:28-72) through the host library's parser."""

    def __init__(self, path=None, text=None, echo=False):
        lib = load_host()
        self.h = lib.peh_input_create()
        rc = lib.peh_input_read_file(self.h, str(path).encode(), int(echo)) if path is not None else lib.peh_input_read_string(self.h, (text or "").encode())
        if rc != 0:
            msg = lib.peh_last_error().decode()
            lib.peh_input_destroy(self.h)
            self.h = None
            raise HostError(msg)
        v = InputView()
        lib.peh_input_view_get(self.h, C.byref(v))
        for k, _ in InputView._fields_:
            val = getattr(v, k)
            if k in ("cells_per_axis", "domain_size"):
                val = list(val)
            if k.startswith(("dirichlet_", "neumann_")):
                continue
            setattr(self, k, val)
        nd, nn = v.n_dirichlet, v.n_neumann
        self.displacement_boundary_labels = _np(v.dirichlet_labels, nd, np.int32)
        self.displacement_boundary_components = _np(v.dirichlet_components, nd, np.int32)
        self.displacement_boundary_values = _np(v.dirichlet_values, nd, np.float64)
        self.stress_boundary_labels = _np(v.neumann_labels, nn, np.int32)
        self.stress_boundary_components = _np(v.neumann_components, nn, np.int32)
        self.stress_boundary_values = _np(v.neumann_values, nn, np.float64)

    def params(self) -> PeParams:
        p = PeParams()
        load_host().peh_input_to_params(self.h, C.byref(p))
        return p

    def __del__(self):
        if getattr(self, "h", None):
            load_host().peh_input_destroy(self.h)
            self.h = None


class OperatorBackend:
    """The L2 operator surface (PS / DS / SP methods) behind a C-ABI.

    ``prefix`` is ``pe_`` for the CUDA library; the tests instantiate the same class over the CPU
    oracle's ``po_`` functions, which have identical signatures.
    """

    def __init__(self, lib, prefix, ctx):
        self.lib, self.prefix, self.ctx = lib, prefix, ctx

    def _f(self, name):
        return getattr(self.lib, self.prefix + name)

    def _ck(self, rc, what):
        if rc != 0:
            msg = self._f("last_error")(self.ctx)
            raise BackendError(rc, f"{what}: {msg.decode() if msg else ''} (status {rc})")

    def close(self):
        if self.ctx:
            self._f("destroy")(self.ctx)
            self.ctx = None

    # uploads
    def set_params(self, prm: PeParams):
        self._ck(self._f("set_params")(self.ctx, C.byref(prm)), "set_params")
        self.dim = prm.dim

    def upload_mesh(self, m: Mesh):
        xyz = np.ascontiguousarray(m.xyz, dtype=np.float64)
        cv = np.ascontiguousarray(m.cell_vertices, dtype=np.int32)
        self._ck(self._f("upload_mesh")(self.ctx, m.dim, m.n_vertices, _p(xyz, C.c_double), m.n_cells, _p(cv, C.c_int32), len(m.bface_cell),
                                        _p(m.bface_cell, C.c_int32), _p(m.bface_local, C.c_int8), _p(m.bface_id, C.c_int32)), "upload_mesh")

    def upload_dofs(self, field, n_dofs, cell_dofs):
        cd = np.ascontiguousarray(cell_dofs, dtype=np.int32)
        self._ck(self._f("upload_dofs")(self.ctx, field, n_dofs, _p(cd, C.c_int32)), "upload_dofs")
        if field == FIELD_PRESSURE:
            self.n_p = n_dofs
        else:
            self.n_u = n_dofs

    def upload_constraints(self, field, line_dof, inhom):
        ld = np.ascontiguousarray(line_dof, dtype=np.int32)
        g = np.ascontiguousarray(inhom, dtype=np.float64)
        ep = np.zeros(len(ld) + 1, dtype=np.int64)
        self._ck(self._f("upload_constraints")(self.ctx, field, len(ld), _p(ld, C.c_int32), _p(ep, C.c_int64), None, None, _p(g, C.c_double)),
                 "upload_constraints")

    def upload_constraint_lines(self, field, L: "ConstraintLines"):
        """General lines (hanging nodes + Dirichlet), PS:71-78 / DS:109-137."""
        self._ck(self._f("upload_constraints")(self.ctx, field, L.n_lines, _p(L.line_dof, C.c_int32), _p(L.entry_ptr, C.c_int64),
                                               _p(L.entry_dof, C.c_int32), _p(L.entry_w, C.c_double), _p(L.inhomogeneity, C.c_double)),
                 "upload_constraints")

    def upload_neumann(self, labels, comps, values):
        l = np.ascontiguousarray(labels, dtype=np.int32)
        c = np.ascontiguousarray(comps, dtype=np.int32)
        v = np.ascontiguousarray(values, dtype=np.float64)
        self._ck(self._f("upload_neumann")(self.ctx, len(l), _p(l, C.c_int32), _p(c, C.c_int32), _p(v, C.c_double)), "upload_neumann")

    def upload_partition(self, field, n_owned, neighbor_rank, send_ptr, send_idx, recv_ptr):
        nr = np.ascontiguousarray(neighbor_rank, dtype=np.int32)
        sp = np.ascontiguousarray(send_ptr, dtype=np.int64)
        si = np.ascontiguousarray(send_idx, dtype=np.int32)
        rp = np.ascontiguousarray(recv_ptr, dtype=np.int64)
        self._ck(self._f("upload_partition")(self.ctx, field, n_owned, len(nr), _p(nr, C.c_int32), _p(sp, C.c_int64), _p(si, C.c_int32),
                                             _p(rp, C.c_int64)), "upload_partition")
        if field == FIELD_PRESSURE:
            self.n_p = n_owned
        else:
            self.n_u = n_owned

    def setup(self):
        self._ck(self._f("setup")(self.ctx), "setup")

    # operators (names follow the reference methods)
    def pressure_set_uniform(self, v): self._ck(self._f("pressure_set_uniform")(self.ctx, v), "pressure_set_uniform")
    def pressure_begin_step(self): self._ck(self._f("pressure_begin_step")(self.ctx), "pressure_begin_step")
    def pressure_zero_update(self): self._ck(self._f("pressure_zero_update")(self.ctx), "pressure_zero_update")
    def update_volumetric_strain(self): self._ck(self._f("pressure_update_volumetric_strain")(self.ctx), "update_volumetric_strain")

    def assemble_residual(self, dt):
        out = C.c_double()
        self._ck(self._f("pressure_assemble_residual")(self.ctx, dt, C.byref(out)), "assemble_residual")
        return out.value

    def assemble_jacobian(self, dt): self._ck(self._f("pressure_assemble_jacobian")(self.ctx, dt), "assemble_jacobian")

    def pressure_solve(self):
        its, res = C.c_int(), C.c_double()
        self._ck(self._f("pressure_solve")(self.ctx, C.byref(its), C.byref(res)), "pressure_solve")
        return its.value, res.value

    def pressure_add_update(self): self._ck(self._f("pressure_add_update")(self.ctx), "pressure_add_update")

    def pressure_linfty(self):
        out = C.c_double()
        self._ck(self._f("pressure_linfty")(self.ctx, C.byref(out)), "pressure_linfty")
        return out.value

    def displacement_assemble(self): self._ck(self._f("displacement_assemble")(self.ctx), "displacement_assemble")

    def displacement_solve(self):
        its, res = C.c_int(), C.c_double()
        self._ck(self._f("displacement_solve")(self.ctx, C.byref(its), C.byref(res)), "displacement_solve")
        return its.value, res.value

    def project_assemble_matrix(self): self._ck(self._f("project_assemble_matrix")(self.ctx), "project_assemble_matrix")

    def project_assemble_rhs(self, comps):
        c = np.ascontiguousarray(comps, dtype=np.int32)
        self._ck(self._f("project_assemble_rhs")(self.ctx, len(c), _p(c, C.c_int32)), "project_assemble_rhs")

    def project_solve(self, entry):
        its = C.c_int()
        self._ck(self._f("project_solve")(self.ctx, entry, C.byref(its)), "project_solve")
        return its.value

    def volumetric_strain_from_projection(self, entries, as_initial):
        e = np.ascontiguousarray(entries, dtype=np.int32)
        self._ck(self._f("volumetric_strain_from_projection")(self.ctx, len(e), _p(e, C.c_int32), int(as_initial)), "volumetric_strain")

    def effective_stresses(self): self._ck(self._f("effective_stresses")(self.ctx), "effective_stresses")

    # inspection
    def _vec_len(self, which):
        return self.n_u if which in (VEC_U, VEC_U_RHS) else self.n_p

    def get_vector(self, which):
        out = np.zeros(self._vec_len(which))
        self._ck(self._f("get_vector")(self.ctx, which, _p(out, C.c_double), out.size), "get_vector")
        return out

    def set_vector(self, which, values):
        v = np.ascontiguousarray(values, dtype=np.float64)
        self._ck(self._f("set_vector")(self.ctx, which, _p(v, C.c_double), v.size), "set_vector")

    def get_matrix(self, which):
        """scipy CSR with sorted columns (canonical), whatever the backend's storage order."""
        import scipy.sparse as sp
        n, nnz = C.c_int64(), C.c_int64()
        self._ck(self._f("get_matrix_size")(self.ctx, which, C.byref(n), C.byref(nnz)), "get_matrix_size")
        rp = np.zeros(n.value + 1, dtype=np.int64)
        col = np.zeros(nnz.value, dtype=np.int32)
        val = np.zeros(nnz.value, dtype=np.float64)
        self._ck(self._f("get_matrix")(self.ctx, which, _p(rp, C.c_int64), _p(col, C.c_int32), _p(val, C.c_double)), "get_matrix")
        ncols = int(col.max()) + 1 if nnz.value else n.value
        A = sp.csr_matrix((val, col, rp), shape=(n.value, max(ncols, n.value)))
        A.sort_indices()
        return A

    def stats(self):
        s = PeStats()
        self._ck(self._f("get_stats")(self.ctx, C.byref(s)), "get_stats")
        return s.as_dict()

    def reset_stats(self): self._ck(self._f("reset_stats")(self.ctx), "reset_stats")


class BackendError(RuntimeError):
    def __init__(self, status, msg):
        super().__init__(msg)
        self.status = status


def create_device_backend(device=0, rank=0, nranks=1, nccl_id: bytes | None = None) -> OperatorBackend:
    lib = load_device()
    ctx = C.c_void_p()
    buf = C.create_string_buffer(nccl_id, len(nccl_id)) if nccl_id else None
    rc = lib.pe_create(C.byref(ctx), device, rank, nranks, buf, len(nccl_id) if nccl_id else 0)
    if rc != 0:
        raise BackendError(rc, f"pe_create failed: {lib.pe_last_error(None).decode()} (status {rc})")
    b = OperatorBackend(lib, "pe_", ctx)
    return b


def nccl_unique_id() -> bytes:
    lib = load_device()
    buf = C.create_string_buffer(128)
    n = C.c_size_t(128)
    rc = lib.pe_nccl_unique_id(buf, C.byref(n))
    if rc != 0:
        raise BackendError(rc, "pe_nccl_unique_id failed")
    return buf.raw[: n.value]


def device_spmv(backend: OperatorBackend, matrix, x=None, reps=1, want_y=False):
    lib = backend.lib
    ms = C.c_float()
    n = backend.n_u if matrix == MAT_ELASTICITY else backend.n_p
    xx = np.ascontiguousarray(x, dtype=np.float64) if x is not None else None
    y = np.zeros(n) if want_y else None
    rc = lib.pe_spmv(backend.ctx, matrix, _p(xx, C.c_double) if xx is not None else None, _p(y, C.c_double) if y is not None else None, reps, C.byref(ms))
    backend._ck(rc, "spmv")
    return ms.value, y


class Problem:
    """The C++ PoroElasticProblem driver (host library) — the product path bench.py times."""

    def __init__(self, inp: InputData, device=0, rank=0, nranks=1, nccl_id: bytes | None = None):
        lib = load_host()
        self.lib = lib
        buf = C.create_string_buffer(nccl_id, len(nccl_id)) if nccl_id else None
        self.h = lib.peh_problem_create(inp.h, device, rank, nranks, buf, len(nccl_id) if nccl_id else 0)
        if not self.h:
            raise HostError(lib.peh_last_error().decode())
        self._inp = inp
        self.backend = None

    def _attach_backend(self):
        dev = load_device()
        self.backend = OperatorBackend(dev, "pe_", C.c_void_p(self.lib.peh_problem_ctx(self.h)))
        self.refresh_sizes()

    def refresh_sizes(self):
        """dof counts of the current mesh (they change when the adaptive loop refines, FSS:333-340)"""
        st = self.backend.stats()
        self.backend.n_p, self.backend.n_u = st["n_dofs_p"], st["n_dofs_u"]
        return st

    def initialize(self, verbose=False):
        if self.lib.peh_problem_initialize(self.h, int(verbose)) != 0:
            raise HostError(self.lib.peh_last_error().decode())
        self._attach_backend()

    def step(self, verbose=False):
        r = StepReport()
        if self.lib.peh_problem_step(self.h, int(verbose), C.byref(r)) != 0:
            raise HostError(self.lib.peh_last_error().decode())
        if self._inp.refine_every:
            self.refresh_sizes()
        return r.as_dict()

    def run(self, verbose=True):
        if self.lib.peh_problem_run(self.h, int(verbose)) != 0:
            raise HostError(self.lib.peh_last_error().decode())
        self._attach_backend()

    def global_ids(self, field):
        n = self.backend.n_p if field == FIELD_PRESSURE else self.backend.n_u
        out = np.zeros(n, dtype=np.int64)
        self.lib.peh_problem_global_ids(self.h, field, _p(out, C.c_int64))
        return out

    def close(self):
        if self.h:
            self.lib.peh_problem_destroy(self.h)  # destroys the pe_ctx too
            self.h = None
            if self.backend:
                self.backend.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
