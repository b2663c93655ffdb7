// driver_on_oracle.cpp — TEST INFRASTRUCTURE: the pe_* C-ABI (include/poroel.h) forwarded to the CPU oracle's po_* functions,
// so that the product's C++ driver (csrc/host/host_capi.cpp + problem.hpp: PoroElasticProblem::initialize / step / run /
// refine_mesh / output_results) can be linked WITHOUT the CUDA library and executed on a machine without a GPU.  The
// tests build  host_capi.cpp + this file + liboracle.so  into a private library and compare the C++ driver with the Python
// mirror of the same loop (poroelasticity-dealii_b200/fss.py) on the same oracle: every host-side step — mesh, dof
// numbering, constraint tables, upload order, the adaptive loop — is then exercised on the CPU.  Never shipped, never
// loaded by the product.
#include <cstddef>
#include <cstdint>

#include "../include/poroel.h"

struct Ctx;
extern "C" {
int po_create(Ctx** out);
void po_destroy(Ctx* c);
const char* po_last_error(const Ctx* c);
int po_set_params(Ctx*, const pe_params*);
int po_upload_mesh(Ctx*, int, int64_t, const double*, int64_t, const int32_t*, int64_t, const int32_t*, const int8_t*, const int32_t*);
int po_upload_dofs(Ctx*, int, int64_t, const int32_t*);
int po_upload_constraints(Ctx*, int, int64_t, const int32_t*, const int64_t*, const int32_t*, const double*, const double*);
int po_upload_neumann(Ctx*, int, const int32_t*, const int32_t*, const double*);
int po_setup(Ctx*);
int po_pressure_set_uniform(Ctx*, double);
int po_pressure_begin_step(Ctx*);
int po_pressure_zero_update(Ctx*);
int po_pressure_update_volumetric_strain(Ctx*);
int po_pressure_assemble_residual(Ctx*, double, double*);
int po_pressure_assemble_jacobian(Ctx*, double);
int po_pressure_solve(Ctx*, int*, double*);
int po_pressure_add_update(Ctx*);
int po_pressure_linfty(Ctx*, double*);
int po_displacement_assemble(Ctx*);
int po_displacement_solve(Ctx*, int*, double*);
int po_project_assemble_matrix(Ctx*);
int po_project_assemble_rhs(Ctx*, int, const int32_t*);
int po_project_solve(Ctx*, int, int*);
int po_volumetric_strain_from_projection(Ctx*, int, const int32_t*, int);
int po_effective_stresses(Ctx*);
int po_get_vector(Ctx*, int, double*, int64_t);
int po_set_vector(Ctx*, int, const double*, int64_t);
int po_get_matrix_size(Ctx*, int, int64_t*, int64_t*);
int po_get_matrix(Ctx*, int, int64_t*, int32_t*, double*);
int po_get_stats(Ctx*, pe_stats*);
int po_reset_stats(Ctx*);
}

static Ctx* O(pe_ctx* c) { return reinterpret_cast<Ctx*>(c); }

extern "C" {

int pe_version(void) { return -1; }  // marks the oracle-backed build
int pe_nccl_unique_id(void*, size_t*) { return PE_ERR_UNSUPPORTED; }
int pe_create(pe_ctx** out, int, int, int nranks, const void*, size_t) {
  if (nranks != 1) return PE_ERR_UNSUPPORTED;
  return po_create(reinterpret_cast<Ctx**>(out));
}
void pe_destroy(pe_ctx* c) { po_destroy(O(c)); }
const char* pe_last_error(const pe_ctx* c) { return c ? po_last_error(reinterpret_cast<const Ctx*>(c)) : "oracle-backed test build"; }
int pe_set_params(pe_ctx* c, const pe_params* p) { return po_set_params(O(c), p); }
int pe_upload_mesh(pe_ctx* c, int dim, int64_t nv, const double* xyz, int64_t nc, const int32_t* cv, int64_t nb, const int32_t* bc, const int8_t* bl,
                   const int32_t* bi) {
  return po_upload_mesh(O(c), dim, nv, xyz, nc, cv, nb, bc, bl, bi);
}
int pe_upload_dofs(pe_ctx* c, int f, int64_t n, const int32_t* cd) { return po_upload_dofs(O(c), f, n, cd); }
int pe_upload_constraints(pe_ctx* c, int f, int64_t n, const int32_t* ld, const int64_t* ep, const int32_t* ed, const double* ew, const double* g) {
  return po_upload_constraints(O(c), f, n, ld, ep, ed, ew, g);
}
int pe_upload_neumann(pe_ctx* c, int n, const int32_t* l, const int32_t* comp, const double* v) { return po_upload_neumann(O(c), n, l, comp, v); }
int pe_upload_partition(pe_ctx*, int, int64_t, int, const int32_t*, const int64_t*, const int32_t*, const int64_t*) { return PE_ERR_UNSUPPORTED; }
int pe_setup(pe_ctx* c) { return po_setup(O(c)); }
int pe_pressure_set_uniform(pe_ctx* c, double v) { return po_pressure_set_uniform(O(c), v); }
int pe_pressure_begin_step(pe_ctx* c) { return po_pressure_begin_step(O(c)); }
int pe_pressure_zero_update(pe_ctx* c) { return po_pressure_zero_update(O(c)); }
int pe_pressure_update_volumetric_strain(pe_ctx* c) { return po_pressure_update_volumetric_strain(O(c)); }
int pe_pressure_assemble_residual(pe_ctx* c, double dt, double* l2) { return po_pressure_assemble_residual(O(c), dt, l2); }
int pe_pressure_assemble_jacobian(pe_ctx* c, double dt) { return po_pressure_assemble_jacobian(O(c), dt); }
int pe_pressure_solve(pe_ctx* c, int* its, double* res) { return po_pressure_solve(O(c), its, res); }
int pe_pressure_add_update(pe_ctx* c) { return po_pressure_add_update(O(c)); }
int pe_pressure_linfty(pe_ctx* c, double* v) { return po_pressure_linfty(O(c), v); }
int pe_displacement_assemble(pe_ctx* c) { return po_displacement_assemble(O(c)); }
int pe_displacement_solve(pe_ctx* c, int* its, double* res) { return po_displacement_solve(O(c), its, res); }
int pe_project_assemble_matrix(pe_ctx* c) { return po_project_assemble_matrix(O(c)); }
int pe_project_assemble_rhs(pe_ctx* c, int n, const int32_t* comps) { return po_project_assemble_rhs(O(c), n, comps); }
int pe_project_solve(pe_ctx* c, int e, int* its) { return po_project_solve(O(c), e, its); }
int pe_volumetric_strain_from_projection(pe_ctx* c, int n, const int32_t* e, int init) { return po_volumetric_strain_from_projection(O(c), n, e, init); }
int pe_effective_stresses(pe_ctx* c) { return po_effective_stresses(O(c)); }
int pe_spmv(pe_ctx*, int, const double*, double*, int, float*) { return PE_ERR_UNSUPPORTED; }
int pe_get_vector(pe_ctx* c, int w, double* h, int64_t n) { return po_get_vector(O(c), w, h, n); }
int pe_set_vector(pe_ctx* c, int w, const double* h, int64_t n) { return po_set_vector(O(c), w, h, n); }
int pe_get_matrix_size(pe_ctx* c, int m, int64_t* n, int64_t* nnz) { return po_get_matrix_size(O(c), m, n, nnz); }
int pe_get_matrix(pe_ctx* c, int m, int64_t* rp, int32_t* col, double* val) { return po_get_matrix(O(c), m, rp, col, val); }
int pe_get_stats(pe_ctx* c, pe_stats* s) { return po_get_stats(O(c), s); }
int pe_reset_stats(pe_ctx* c) { return po_reset_stats(O(c)); }
int pe_synchronize(pe_ctx*) { return PE_OK; }
int pe_set_profiling(pe_ctx*, int) { return PE_OK; }
void* pe_stream(pe_ctx*) { return nullptr; }

}  // extern "C"
