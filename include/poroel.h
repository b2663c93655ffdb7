/* poroel.h — C-ABI of the B200 (sm_100a) device library `libporoel.so`.
 *
 * Drop-in boundary for the fixed-stress-split poroelastic time step of
 * ishovkun/poroelasticity-dealii.  The reference has no FFI; the seam is the set of public
 * methods/members of its three field sub-solvers that PoroElasticProblem<dim>::run()
 * touches (lib/include/PoroelasticityFSS.h:294-415).  Every entry point below names the
 * reference method it replaces.  One pe_ctx == one process == one GPU (one rank); for
 * cell-partitioned multi-GPU runs each rank uploads its LOCAL sub-mesh (owned + one ghost
 * cell layer) plus an exchange plan, see pe_upload_partition.
 *
 * Conventions
 *   - plain pointers and sizes only; host arrays are borrowed for the duration of the call;
 *   - all functions return 0 on success or a negative pe_status; no exception crosses;
 *   - local cell dof order is deal.II's: FE_Q(k) scalar order = vertices (lexicographic,
 *     x fastest), lines, quads, hex interior; vector dof = scalar_local * n_comp + comp;
 *   - all arithmetic is FP64, indices are 32-bit; there is NO CPU fallback: if no CUDA
 *     device is usable pe_create fails with PE_ERR_CUDA.
 */
#ifndef POROEL_H
#define POROEL_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct pe_ctx pe_ctx;

enum pe_status {
  PE_OK = 0,
  PE_ERR_CUDA = -1,
  PE_ERR_NCCL = -2,
  PE_ERR_NO_CONVERGENCE = -3, /* mirrors dealii::SolverControl::NoConvergence (PS:175, DS:299, SP:209) */
  PE_ERR_NAN = -4,
  PE_ERR_BAD_INPUT = -5,
  PE_ERR_UNSUPPORTED = -6,
  PE_ERR_STATE = -7
};

enum pe_field { PE_FIELD_PRESSURE = 0, PE_FIELD_DISPLACEMENT = 1 };

enum pe_precond {
  PE_PRECOND_JACOBI = 0,    /* D^-1                                                       */
  PE_PRECOND_CHEBYSHEV = 1  /* degree-k Chebyshev polynomial in D^-1 A (north_star option) */
};

/* which = vector ids for pe_get_vector / pe_set_vector */
enum pe_vector {
  PE_VEC_P = 0,            /* pressure_solver.solution          PS:38  */
  PE_VEC_P_OLD = 1,        /* pressure_solver.old_solution      PS:39  */
  PE_VEC_P_UPDATE = 2,     /* pressure_solver.solution_update   PS:38  */
  PE_VEC_P_RESIDUAL = 3,   /* pressure_solver.residual          PS:40  */
  PE_VEC_VOL_STRAIN = 4,   /* volumetric_strain                 FSS:82 */
  PE_VEC_VOL_STRAIN0 = 5,  /* initial_volumetric_strain         FSS:82 */
  PE_VEC_WELL_RHS = 6,     /* cached create_right_hand_side     PS:142-147 */
  PE_VEC_U = 7,            /* displacement_solver.solution      DS:47  */
  PE_VEC_U_RHS = 8,        /* displacement_solver.rhs_vector    DS:53  */
  PE_VEC_STRAIN0 = 16,     /* strain_projector.strains[e], e = which - 16  SP:43 */
  PE_VEC_PROJ_RHS0 = 32,   /* strain_projector.projection_rhs[e]           SP:43 */
  PE_VEC_STRESS0 = 48      /* stresses[e]  FSS:83 (effective stresses, FSS:189-224) */
};

enum pe_matrix {
  PE_MAT_MASS = 0,      /* pressure_solver.mass_matrix     PS:44 */
  PE_MAT_LAPLACE = 1,   /* pressure_solver.laplace_matrix  PS:44 */
  PE_MAT_JACOBIAN = 2,  /* pressure_solver.jacobian        PS:44 */
  PE_MAT_ELASTICITY = 3, /* displacement_solver.system_matrix DS:52 */
  PE_MAT_PROJECTION = 4  /* strain_projector.projection_matrix SP:101-106 (the condensed mass matrix) */
};

typedef struct pe_params {
  int32_t dim;             /* 2 or 3                        input.data "Dimensions"       */
  int32_t degree_u;        /* 1 or 2; reference hard-codes 2 (DS:67)                       */
  int32_t degree_p;        /* 1 (PS:20 default)                                            */
  int32_t preconditioner;  /* pe_precond                                                    */
  int32_t chebyshev_degree;      /* number of Chebyshev steps per apply (>=1)               */
  int32_t cg_max_iterations;     /* SolverControl max_steps, reference 1000                 */
  int32_t cg_check_interval;     /* host polls the device convergence flag every N its     */
  int32_t reserved0;
  double lame_lambda;      /* ID:216 */
  double shear_modulus;    /* ID:217 */
  double bulk_modulus;     /* ID:218 */
  double biot_coef;        /* ID:158 */
  double m_modulus;        /* ID:221 */
  double perm_over_visc;   /* data.perm/data.visc, PS:137,165 */
  double well_radius;      /* RHS:87-116 */
  double flow_rate;
  double cg_rel_tol_pressure;      /* 1e-8 * ||residual||  PS:175 */
  double cg_abs_tol_displacement;  /* 1e-12                DS:298 */
  double cg_rel_tol_projection;    /* 1e-8 * ||rhs||       SP:209 */
  double chebyshev_eig_ratio;      /* lambda_max / lambda_min of the smoothed interval     */
} pe_params;

typedef struct pe_stats {
  int64_t n_cells, n_dofs_p, n_dofs_u, nnz_p, nnz_u;  /* local (this rank) */
  int64_t cg_iterations_pressure;      /* cumulative since last pe_reset_stats */
  int64_t cg_iterations_displacement;
  int64_t cg_iterations_projection;
  int64_t cg_solves_pressure, cg_solves_displacement, cg_solves_projection;
  int64_t spmv_launches_p, spmv_launches_u;  /* matrix passes (CG + preconditioner)        */
  int64_t kernel_launches;                   /* all kernels launched by this ctx           */
  double  spmv_bytes_p, spmv_bytes_u;        /* algorithmic bytes per matrix pass          */
  double  eig_max_p, eig_max_u, eig_max_m;   /* power-iteration estimates of lambda_max(D^-1 A) */
  double  setup_ms;                          /* pe_setup wall time                         */
  double  spmv_ms_p, spmv_ms_u;              /* CUDA-event time summed over the matrix passes timed while
                                                profiling is on (pe_set_profiling)          */
  int64_t spmv_timed_p, spmv_timed_u;        /* number of matrix passes in those sums      */
  double  pcg_ms_p, pcg_ms_u;                /* CUDA-event time of the persistent CG kernel launches (profiling on) */
  int64_t pcg_iterations_p, pcg_iterations_u;/* CG iterations executed inside those launches */
  int64_t bsr_block_size;                    /* block size of the block-CSR copy of the displacement matrix (0 = plain CSR);
                                                spmv_bytes_u then counts 8 B per value + 4 B per BLOCK column index */
  /* persistent single-reduction CG kernel on the TMA-fed sliced copies (profiling on; displacement solves only):     */
  double  inner_ms_u;                        /* time in the passes inside the polynomial preconditioner (incl. their barrier) */
  int64_t inner_passes_u;                    /* number of those passes                                                  */
  double  inner_bytes_u;                     /* algorithmic bytes of one such pass (FP32 copy of the values when built)  */
  double  update_ms_u, reduce_ms_u;          /* vector-update phases; the one allreduce per iteration (incl. its barrier) */
  int64_t sell_format_u;                     /* 1: the displacement matrix is streamed from its sliced block-ELL copy     */
  /* time CTA 0 of that kernel spent WAITING (a subset of the phase times above): at the grid barrier behind the inner
     passes, at the one behind the CG pass, for the peers' reduction mailboxes, at the barrier behind the updates */
  double  wait_inner_ms_u, wait_cg_ms_u, wait_peer_ms_u, wait_update_ms_u;
  /* the same clock of the pressure / projection solves: [0] CG passes (own stream), [1] inner passes, [2] updates, [3] reductions,
     [4..7] the four waits in the order above; [8] number of CG passes, [9] number of inner passes */
  double  phase_ms_p[10];
} pe_stats;

/* ---- lifetime ------------------------------------------------------------------------- */
/* rank/nranks/nccl_id describe a one-process-per-GPU job; nranks == 1 needs no id.        */
int  pe_nccl_unique_id(void* out128, size_t* n_bytes);  /* rank 0: ncclGetUniqueId         */
int  pe_create(pe_ctx** out, int device, int rank, int nranks, const void* nccl_id, size_t id_bytes);
void pe_destroy(pe_ctx*);
const char* pe_last_error(const pe_ctx*);  /* also valid with NULL: last create error      */
int  pe_version(void);

/* ---- one-time upload (PS:68-111, DS:106-153, FSS:131-151) ------------------------------ */
int  pe_set_params(pe_ctx*, const pe_params*);
int  pe_upload_mesh(pe_ctx*, int dim, int64_t n_vertices, const double* xyz /* n_vertices*dim */,
                    int64_t n_cells, const int32_t* cell_vertices /* n_cells * 2^dim, lexicographic */,
                    int64_t n_bfaces, const int32_t* bface_cell, const int8_t* bface_local,
                    const int32_t* bface_id);
int  pe_upload_dofs(pe_ctx*, int field, int64_t n_dofs_local, const int32_t* cell_dofs /* n_cells * n_loc */);
/* general linear constraint table x_i = sum_j a_ij x_j + g_i (ConstraintMatrix after close(), PS:71-78,
 * DS:109-137).  Lines without entries are Dirichlet values; lines with entries are hanging nodes of
 * an adaptively refined mesh (FSS:333-340) and may also carry an inhomogeneity.  Entries must refer
 * to unconstrained dofs (closed table).  Pressure lines are homogeneous hanging-node lines only.
 * Hanging-node meshes run on one rank.                                                        */
int  pe_upload_constraints(pe_ctx*, int field, int64_t n_lines, const int32_t* line_dof,
                           const int64_t* entry_ptr, const int32_t* entry_dof, const double* entry_w,
                           const double* inhomogeneity);
int  pe_upload_neumann(pe_ctx*, int n, const int32_t* label, const int32_t* comp, const double* value); /* DS:78-94 */
/* exchange plan of a cell-partitioned run: local dofs are [owned | ghosts grouped by owner]. */
int  pe_upload_partition(pe_ctx*, int field, int64_t n_owned, int n_neighbors, const int32_t* neighbor_rank,
                         const int64_t* send_ptr, const int32_t* send_idx, const int64_t* recv_ptr);
int  pe_setup(pe_ctx*);  /* device CSR patterns, colouring, M, K, well rhs, allocations */

/* ---- hot-path operators, 1:1 with the reference methods -------------------------------- */
int  pe_pressure_set_uniform(pe_ctx*, double p_init);                     /* FSS:311 */
int  pe_pressure_begin_step(pe_ctx*);                                     /* FSS:342  old_solution = solution */
int  pe_pressure_zero_update(pe_ctx*);                                    /* FSS:356  solution_update = 0 */
int  pe_pressure_update_volumetric_strain(pe_ctx*);                       /* PS:187-194 */
int  pe_pressure_assemble_residual(pe_ctx*, double time_step, double* l2_norm); /* PS:113-155 + FSS:364/405 */
int  pe_pressure_assemble_jacobian(pe_ctx*, double time_step);            /* PS:158-169 */
int  pe_pressure_solve(pe_ctx*, int* cg_iterations, double* final_residual); /* PS:172-185 */
int  pe_pressure_add_update(pe_ctx*);                                     /* FSS:379  solution += solution_update */
int  pe_pressure_linfty(pe_ctx*, double* value);                          /* FSS:387-389 */
int  pe_displacement_assemble(pe_ctx*);                                   /* DS:155-291 (matrix on first call only) */
int  pe_displacement_solve(pe_ctx*, int* cg_iterations, double* final_residual); /* DS:294-307 */
int  pe_project_assemble_matrix(pe_ctx*);                                 /* SP:101-106 */
int  pe_project_assemble_rhs(pe_ctx*, int n_comp, const int32_t* tensor_components); /* SP:109-198 */
int  pe_project_solve(pe_ctx*, int rhs_entry, int* cg_iterations);        /* SP:201-232 */
int  pe_volumetric_strain_from_projection(pe_ctx*, int n_entries, const int32_t* rhs_entries, int as_initial); /* FSS:179-186, 317 */
int  pe_effective_stresses(pe_ctx*);                                      /* FSS:189-224 */

/* ---- micro-benchmark / inspection ------------------------------------------------------ */
int  pe_spmv(pe_ctx*, int matrix, const double* x_host_or_null, double* y_host_or_null, int repetitions, float* ms_per_rep);
int  pe_get_vector(pe_ctx*, int which, double* host, int64_t n);          /* owned entries */
int  pe_set_vector(pe_ctx*, int which, const double* host, int64_t n);
int  pe_get_matrix_size(pe_ctx*, int matrix, int64_t* n_rows, int64_t* nnz);
int  pe_get_matrix(pe_ctx*, int matrix, int64_t* rowptr, int32_t* col, double* val); /* CSR, columns ascending (local ids) */
int  pe_get_stats(pe_ctx*, pe_stats*);
int  pe_reset_stats(pe_ctx*);
int  pe_synchronize(pe_ctx*);
int  pe_set_profiling(pe_ctx*, int on); /* bracket every matrix pass with CUDA events on the launch stream */
void* pe_stream(pe_ctx*);   /* cudaStream_t the kernels are launched on (for event timing) */

#ifdef __cplusplus
}
#endif
#endif /* POROEL_H */
