/* poroel_host.h — C-ABI of the host-side surface `libporoel_host.so`.
 *
 * Host C++ that keeps the reference's input surface and solver driver:
 *   - input.data parser ........... lib/include/InputDataPoroel.h:77-222
 *   - create_mesh / read_mesh ..... lib/include/PoroelasticityFSS.h:418-445
 *   - DoF numbering, Dirichlet .... PS:73, DS:110-135 (deal.II semantics restated)
 *   - PoroElasticProblem driver ... lib/include/PoroelasticityFSS.h:294-415 (run), calling
 *     the device library exclusively through include/poroel.h.
 * Exposed as a plain C-ABI so tests and bench.py can drive it through ctypes.
 */
#ifndef POROEL_HOST_H
#define POROEL_HOST_H
#include <stddef.h>
#include <stdint.h>

#include "poroel.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct peh_input peh_input;
typedef struct peh_mesh peh_mesh;
typedef struct peh_dofs peh_dofs;
typedef struct peh_problem peh_problem;

const char* peh_last_error(void);

/* ---- input.data (ID:77-222) ---- */
typedef struct peh_input_view {
  int32_t dim, initial_refinement_level, max_refinement_level;
  int32_t max_fss_iterations, max_pressure_iterations;
  int32_t displacement_degree, preconditioner, chebyshev_degree, cg_max_iterations;
  int32_t mesh_from_file, refine_every, couple_volumetric_strain, write_vtk, max_time_steps;
  int32_t cells_per_axis[3];
  int32_t n_dirichlet, n_neumann;
  double domain_size[3];
  double perm, poro, visc, f_comp, youngs_modulus, poisson_ratio, biot_coef, bulk_density, r_well, flow_rate;
  double time_step, t_max, fss_tol, pressure_tol, p_init;
  double lame_constant, shear_modulus, bulk_modulus, grain_bulk_modulus, n_modulus, m_modulus;
  double chebyshev_eig_ratio;
  const int32_t* dirichlet_labels; const int32_t* dirichlet_components; const double* dirichlet_values;
  const int32_t* neumann_labels; const int32_t* neumann_components; const double* neumann_values;
} peh_input_view;

peh_input* peh_input_create(void);
void peh_input_destroy(peh_input*);
int  peh_input_read_file(peh_input*, const char* path, int echo);   /* ID:77-86 */
int  peh_input_read_string(peh_input*, const char* text);
int  peh_input_view_get(peh_input*, peh_input_view* out);
int  peh_input_to_params(const peh_input*, pe_params* out);

/* ---- mesh (FSS:418-445) ---- */
typedef struct peh_mesh_view {
  int32_t dim; int32_t morton;
  int64_t n_vertices, n_cells, n_bfaces;
  const double* xyz; const int32_t* cell_vertices;
  const int32_t* bface_cell; const int8_t* bface_local; const int32_t* bface_id;
} peh_mesh_view;
peh_mesh* peh_mesh_create_rectangle(int dim, const double* size, int refine_level);  /* FSS:418-435 */
peh_mesh* peh_mesh_create_subdivided(int dim, const double* size, const int32_t* n);
peh_mesh* peh_mesh_read_msh(const char* path, int dim);                              /* FSS:438-445 */
/* Morton order of the cell centroids: contiguous cell ranges become compact subdomains (partitioned unstructured meshes) */
int  peh_mesh_reorder_sfc(peh_mesh*, int64_t* perm_new_to_old_or_null);
int  peh_mesh_permute_cells(peh_mesh*, const int64_t* perm_new_to_old);
void peh_mesh_destroy(peh_mesh*);
int  peh_mesh_view_get(const peh_mesh*, peh_mesh_view* out);

/* ---- dofs and constraints (PS:73, DS:110-135) ---- */
typedef struct peh_dofs_view {
  int32_t degree, n_comp, n_loc, reserved;
  int64_t n_dofs;
  const int32_t* cell_dofs;
} peh_dofs_view;
peh_dofs* peh_dofs_distribute(const peh_mesh*, int degree, int n_comp);
void peh_dofs_destroy(peh_dofs*);
int  peh_dofs_view_get(const peh_dofs*, peh_dofs_view* out);
int  peh_dofs_support_points(const peh_mesh*, const peh_dofs*, double* out /* n_dofs*dim */);
/* returns the number of lines; call with NULL outputs first to size the arrays */
int64_t peh_make_dirichlet(const peh_mesh*, const peh_dofs*, int n, const int32_t* labels, const int32_t* comps,
                           const double* values, int32_t* line_dof, double* inhomogeneity);

/* ---- adaptive refinement of the time loop (FSS:333-340, 447-498): refinement forest, hanging-node constraints,
 *      Kelly estimator, fixed-fraction marking, FE_Q(1) solution transfer (csrc/host/amr.hpp) ---- */
typedef struct peh_forest peh_forest;
typedef struct peh_constraints peh_constraints;
peh_forest* peh_forest_create(const peh_mesh* initial, int base_level);  /* cells of `initial` become roots at base_level */
void peh_forest_destroy(peh_forest*);
peh_mesh* peh_forest_active_mesh(const peh_forest*);  /* new handle; active cells level by level (deal.II order) */
int64_t peh_forest_active_levels(const peh_forest*, int32_t* level_or_null);  /* returns the number of active cells */
int  peh_forest_set_flags(peh_forest*, int64_t n_active, const int8_t* refine, const int8_t* coarsen);
int  peh_forest_get_flags(const peh_forest*, int64_t n_active, int8_t* refine, int8_t* coarsen);
int  peh_forest_prepare(peh_forest*);   /* Triangulation::prepare_coarsening_and_refinement, FSS:481 */
int  peh_forest_execute(peh_forest*, int32_t* n_coarsened_families, int32_t* n_refined_cells);  /* FSS:483 */
/* KellyErrorEstimator<dim>::estimate(pressure dof_handler, QGauss<dim-1>(2), {}, p, eta), FSS:454-458 */
int  peh_forest_kelly(const peh_forest*, const peh_mesh* active, const peh_dofs* dofs_p, const double* p, float* eta);
/* GridRefinement::refine_and_coarsen_fixed_fraction + the level limits of FSS:460-472 */
int  peh_forest_mark_fixed_fraction(peh_forest*, int64_t n_active, const float* criteria, double top_fraction, double bottom_fraction,
                                    int min_level, int max_level);
/* SolutionTransfer<dim> of FE_Q(1) vectors (FSS:475-497): store before execute, fetch on the new active mesh */
int  peh_forest_store(peh_forest*, const peh_mesh* active, const peh_dofs* dofs_p, int n_vec, const double* values /* n_vec*n_dofs */);
int  peh_forest_fetch(const peh_forest*, const peh_mesh* active, const peh_dofs* dofs_p, int n_vec, double* values);
/* make_hanging_node_constraints + interpolate_boundary_values per (label, component, value) + close (PS:71-78, DS:109-137) */
typedef struct peh_constraints_view {
  int64_t n_lines, n_entries;
  const int32_t* line_dof; const int64_t* entry_ptr; const int32_t* entry_dof; const double* entry_w; const double* inhomogeneity;
} peh_constraints_view;
peh_constraints* peh_constraints_make(const peh_forest*, const peh_mesh* active, const peh_dofs*, int n_dirichlet, const int32_t* labels,
                                      const int32_t* comps, const double* values);
void peh_constraints_destroy(peh_constraints*);
int  peh_constraints_view_get(const peh_constraints*, peh_constraints_view* out);

/* ---- cell partition of a mesh for one rank (owned + ghost cell layer, local numbering) ---- */
typedef struct peh_part peh_part;
peh_part* peh_partition(const peh_mesh*, const peh_dofs* dofs_p, const peh_dofs* dofs_u, int rank, int nranks);
/* the same part for a structured box with FE_Q(1) for both fields (Morton 2^L grid of FSS:418-435 or a lexicographic
 * n_x x n_y x n_z box), built from the lattice alone: no global mesh, no global dof maps on the rank.  Identical to
 * peh_partition() on the global mesh, array by array.                                                              */
peh_part* peh_partition_structured(int dim, const double* size, const int32_t* cells_per_axis, int morton_order, int rank, int nranks);
void peh_part_destroy(peh_part*);
typedef struct peh_part_field_view {
  int64_t n_owned, n_local; int32_t n_neighbors, reserved;
  const int32_t* cell_dofs;          /* local cells * n_loc, local ids */
  const int64_t* local_to_global;    /* n_local */
  const int32_t* neighbor_rank; const int64_t* send_ptr; const int32_t* send_idx; const int64_t* recv_ptr;
} peh_part_field_view;
typedef struct peh_part_view {
  peh_mesh_view mesh;                /* local sub-mesh */
  const int64_t* cell_global;        /* local cell -> global cell */
  int64_t n_owned_cells;
  peh_part_field_view field[2];      /* PE_FIELD_PRESSURE, PE_FIELD_DISPLACEMENT */
} peh_part_view;
int  peh_part_view_get(const peh_part*, peh_part_view* out);

/* ---- the solver driver: PoroElasticProblem<dim> (FSS:49-90) ---- */
typedef struct peh_step_report {
  double time; int32_t time_step_number;
  int32_t fss_iterations;            /* coupling iterations of this step (FSS:347-407) */
  int32_t pressure_iterations;       /* total inner pressure iterations (FSS:358-382)  */
  int32_t cg_its_pressure, cg_its_displacement, cg_its_projection;
  int32_t status;                    /* pe_status */
  double pressure_error;             /* last "Error:" value FSS:405 */
  double pressure_linfty;            /* "Solution limits:" FSS:387-389 */
} peh_step_report;
peh_problem* peh_problem_create(const peh_input*, int device, int rank, int nranks, const void* nccl_id, size_t id_bytes);
void peh_problem_destroy(peh_problem*);
int  peh_problem_initialize(peh_problem*, int verbose); /* FSS:297-317: mesh, BCs, setup_dofs, initial state */
int  peh_problem_step(peh_problem*, int verbose, peh_step_report* out); /* FSS:328-407 */
int  peh_problem_run(peh_problem*, int verbose);        /* FSS:294-415 */
pe_ctx* peh_problem_ctx(peh_problem*);
const peh_mesh* peh_problem_mesh(peh_problem*);         /* local mesh of this rank */
int  peh_problem_global_ids(peh_problem*, int field, int64_t* out /* n_owned */);

#ifdef __cplusplus
}
#endif
#endif
