// host_capi.cpp — extern "C" surface of the host library (include/poroel_host.h).
#include <cstring>
#include <string>

#include "../../../include/poroel_host.h"
#include "amr.hpp"
#include "dofs.hpp"
#include "input_data.hpp"
#include "mesh.hpp"
#include "partition.hpp"
#include "problem.hpp"

static thread_local std::string g_err;

struct peh_input {
  input_data::InputDataPoroel d;
  std::vector<int32_t> dl, dc, nl, nc;
};
struct peh_mesh { mesh::Mesh m; };
struct peh_dofs { dofs::DofMap d; dofs::NodeMaps maps; };
struct peh_forest { amr::Forest F; };
struct peh_constraints {
  dofs::ConstraintTable T;
  std::vector<int32_t> line_dof, entry_dof;
  std::vector<int64_t> entry_ptr;
  std::vector<double> entry_w, inhomogeneity;
};
struct peh_part { partition::Part p; };
struct peh_problem {
  std::unique_ptr<poro_elastisity::PoroElasticProblem> P;
  peh_mesh local;  // view holder
};

#define PEH_TRY try {
#define PEH_CATCH(ret)                 \
  }                                    \
  catch (const std::exception& e) {    \
    g_err = e.what();                  \
    return ret;                        \
  }                                    \
  catch (...) {                        \
    g_err = "unknown exception";       \
    return ret;                        \
  }

static void fill_mesh_view(const mesh::Mesh& m, peh_mesh_view* v) {
  v->dim = m.dim;
  v->morton = m.morton ? 1 : 0;
  v->n_vertices = m.n_vertices();
  v->n_cells = m.n_cells();
  v->n_bfaces = m.n_bfaces();
  v->xyz = m.xyz.data();
  v->cell_vertices = m.cell_vertices.data();
  v->bface_cell = m.bface_cell.data();
  v->bface_local = m.bface_local.data();
  v->bface_id = m.bface_id.data();
}

extern "C" {

const char* peh_last_error(void) { return g_err.c_str(); }

peh_input* peh_input_create(void) { return new peh_input(); }
void peh_input_destroy(peh_input* p) { delete p; }
int peh_input_read_file(peh_input* p, const char* path, int echo) {
  PEH_TRY
  p->d.read_input_file(path, echo != 0);
  return 0;
  PEH_CATCH(PE_ERR_BAD_INPUT)
}
int peh_input_read_string(peh_input* p, const char* text) {
  PEH_TRY
  p->d.read_input_string(text, false);
  return 0;
  PEH_CATCH(PE_ERR_BAD_INPUT)
}
int peh_input_view_get(peh_input* p, peh_input_view* v) {
  const auto& d = p->d;
  std::memset(v, 0, sizeof *v);
  v->dim = d.dim;
  v->initial_refinement_level = d.initial_refinement_level;
  v->max_refinement_level = d.max_refinement_level;
  v->max_fss_iterations = d.max_fss_iterations;
  v->max_pressure_iterations = d.max_pressure_iterations;
  v->displacement_degree = d.displacement_degree;
  v->preconditioner = d.preconditioner;
  v->chebyshev_degree = d.chebyshev_degree;
  v->cg_max_iterations = d.cg_max_iterations;
  v->mesh_from_file = d.mesh_from_file;
  v->refine_every = d.refine_every;
  v->couple_volumetric_strain = d.couple_volumetric_strain;
  v->write_vtk = d.write_vtk;
  v->max_time_steps = d.max_time_steps;
  for (int i = 0; i < 3; ++i) {
    v->cells_per_axis[i] = d.cells_per_axis[i];
    v->domain_size[i] = i < (int)d.domain_size.size() ? d.domain_size[i] : 0.0;
  }
  v->perm = d.perm; v->poro = d.poro; v->visc = d.visc; v->f_comp = d.f_comp;
  v->youngs_modulus = d.youngs_modulus; v->poisson_ratio = d.poisson_ratio; v->biot_coef = d.biot_coef;
  v->bulk_density = d.bulk_density; v->r_well = d.r_well; v->flow_rate = d.flow_rate;
  v->time_step = d.time_step; v->t_max = d.t_max; v->fss_tol = d.fss_tol; v->pressure_tol = d.pressure_tol; v->p_init = d.p_init;
  v->lame_constant = d.lame_constant; v->shear_modulus = d.shear_modulus; v->bulk_modulus = d.bulk_modulus;
  v->grain_bulk_modulus = d.grain_bulk_modulus; v->n_modulus = d.n_modulus; v->m_modulus = d.m_modulus;
  v->chebyshev_eig_ratio = d.chebyshev_eig_ratio;
  p->dl.assign(d.displacement_boundary_labels.begin(), d.displacement_boundary_labels.end());
  p->dc.assign(d.displacement_boundary_components.begin(), d.displacement_boundary_components.end());
  p->nl.assign(d.stress_boundary_labels.begin(), d.stress_boundary_labels.end());
  p->nc.assign(d.stress_boundary_components.begin(), d.stress_boundary_components.end());
  v->n_dirichlet = (int32_t)p->dl.size();
  v->n_neumann = (int32_t)p->nl.size();
  v->dirichlet_labels = p->dl.data(); v->dirichlet_components = p->dc.data(); v->dirichlet_values = d.displacement_boundary_values.data();
  v->neumann_labels = p->nl.data(); v->neumann_components = p->nc.data(); v->neumann_values = d.stress_boundary_values.data();
  return 0;
}
int peh_input_to_params(const peh_input* p, pe_params* out) { *out = poro_elastisity::params_from_input(p->d); return 0; }

peh_mesh* peh_mesh_create_rectangle(int dim, const double* size, int refine_level) {
  PEH_TRY
  auto* m = new peh_mesh();
  m->m = mesh::create_hyper_rectangle(dim, size, refine_level);
  return m;
  PEH_CATCH(nullptr)
}
peh_mesh* peh_mesh_create_subdivided(int dim, const double* size, const int32_t* n) {
  PEH_TRY
  auto* m = new peh_mesh();
  int nn[3] = {n[0], n[1], dim == 3 ? n[2] : 1};
  m->m = mesh::create_subdivided(dim, size, nn);
  return m;
  PEH_CATCH(nullptr)
}
peh_mesh* peh_mesh_read_msh(const char* path, int dim) {
  PEH_TRY
  auto* m = new peh_mesh();
  m->m = mesh::read_msh(path, dim);
  return m;
  PEH_CATCH(nullptr)
}
int peh_mesh_reorder_sfc(peh_mesh* m, int64_t* perm_new_to_old) {
  PEH_TRY
  std::vector<int64_t> p = mesh::reorder_cells_sfc(m->m);
  if (perm_new_to_old) std::memcpy(perm_new_to_old, p.data(), p.size() * sizeof(int64_t));
  return 0;
  PEH_CATCH(PE_ERR_BAD_INPUT)
}
int peh_mesh_permute_cells(peh_mesh* m, const int64_t* perm_new_to_old) {  // test hook: an arbitrary cell order
  PEH_TRY
  mesh::Mesh& M = m->m;
  const int vpc = M.vpc();
  const int64_t nc = M.n_cells();
  std::vector<int64_t> inv((size_t)nc, -1);
  for (int64_t i = 0; i < nc; ++i) {
    if (perm_new_to_old[i] < 0 || perm_new_to_old[i] >= nc || inv[perm_new_to_old[i]] >= 0) throw std::runtime_error("not a permutation");
    inv[perm_new_to_old[i]] = i;
  }
  std::vector<int32_t> cv((size_t)nc * vpc);
  for (int64_t i = 0; i < nc; ++i)
    for (int v = 0; v < vpc; ++v) cv[i * vpc + v] = M.cell_vertices[perm_new_to_old[i] * vpc + v];
  M.cell_vertices.swap(cv);
  for (auto& c : M.bface_cell) c = (int32_t)inv[c];
  M.morton = false;
  return 0;
  PEH_CATCH(PE_ERR_BAD_INPUT)
}
void peh_mesh_destroy(peh_mesh* m) { delete m; }
int peh_mesh_view_get(const peh_mesh* m, peh_mesh_view* v) { fill_mesh_view(m->m, v); return 0; }

peh_dofs* peh_dofs_distribute(const peh_mesh* m, int degree, int n_comp) {
  PEH_TRY
  if (degree < 1 || degree > 2) throw std::runtime_error("degree must be 1 or 2");
  auto* d = new peh_dofs();
  d->d = dofs::distribute_dofs(m->m, degree, n_comp, &d->maps);
  return d;
  PEH_CATCH(nullptr)
}
void peh_dofs_destroy(peh_dofs* d) { delete d; }
int peh_dofs_view_get(const peh_dofs* d, peh_dofs_view* v) {
  v->degree = d->d.degree; v->n_comp = d->d.n_comp; v->n_loc = d->d.n_loc; v->reserved = 0;
  v->n_dofs = d->d.n_dofs; v->cell_dofs = d->d.cell_dofs.data();
  return 0;
}
int peh_dofs_support_points(const peh_mesh* m, const peh_dofs* d, double* out) {
  PEH_TRY
  std::vector<double> sp = dofs::support_points(m->m, d->d);
  std::memcpy(out, sp.data(), sp.size() * sizeof(double));
  return 0;
  PEH_CATCH(PE_ERR_BAD_INPUT)
}
int64_t peh_make_dirichlet(const peh_mesh* m, const peh_dofs* d, int n, const int32_t* labels, const int32_t* comps, const double* values,
                           int32_t* line_dof, double* inhomogeneity) {
  PEH_TRY
  std::vector<int> l(labels, labels + n), c(comps, comps + n);
  std::vector<double> v(values, values + n);
  dofs::Constraints cs = dofs::make_dirichlet(m->m, d->d, l, c, v);
  if (line_dof) std::memcpy(line_dof, cs.line_dof.data(), cs.line_dof.size() * sizeof(int32_t));
  if (inhomogeneity) std::memcpy(inhomogeneity, cs.inhomogeneity.data(), cs.inhomogeneity.size() * sizeof(double));
  return (int64_t)cs.line_dof.size();
  PEH_CATCH(-1)
}

// ---- adaptive refinement
peh_forest* peh_forest_create(const peh_mesh* m, int base_level) {
  PEH_TRY
  auto* f = new peh_forest();
  f->F = amr::Forest::from_mesh(m->m, base_level);
  return f;
  PEH_CATCH(nullptr)
}
void peh_forest_destroy(peh_forest* f) { delete f; }
peh_mesh* peh_forest_active_mesh(const peh_forest* f) {
  PEH_TRY
  auto* m = new peh_mesh();
  m->m = f->F.active_mesh();
  return m;
  PEH_CATCH(nullptr)
}
int64_t peh_forest_active_levels(const peh_forest* f, int32_t* level) {
  PEH_TRY
  std::vector<int32_t> a = f->F.active_cells();
  if (level) for (size_t i = 0; i < a.size(); ++i) level[i] = f->F.cells[a[i]].level;
  return (int64_t)a.size();
  PEH_CATCH(-1)
}
int peh_forest_set_flags(peh_forest* f, int64_t n, const int8_t* refine, const int8_t* coarsen) {
  PEH_TRY
  std::vector<int32_t> a = f->F.active_cells();
  if ((int64_t)a.size() != n) throw std::runtime_error("flag arrays must have one entry per active cell");
  for (int64_t i = 0; i < n; ++i) {
    f->F.cells[a[i]].refine_flag = refine && refine[i];
    f->F.cells[a[i]].coarsen_flag = coarsen && coarsen[i] && !(refine && refine[i]);
  }
  return 0;
  PEH_CATCH(PE_ERR_BAD_INPUT)
}
int peh_forest_get_flags(const peh_forest* f, int64_t n, int8_t* refine, int8_t* coarsen) {
  PEH_TRY
  std::vector<int32_t> a = f->F.active_cells();
  if ((int64_t)a.size() != n) throw std::runtime_error("flag arrays must have one entry per active cell");
  for (int64_t i = 0; i < n; ++i) {
    if (refine) refine[i] = f->F.cells[a[i]].refine_flag;
    if (coarsen) coarsen[i] = f->F.cells[a[i]].coarsen_flag;
  }
  return 0;
  PEH_CATCH(PE_ERR_BAD_INPUT)
}
int peh_forest_prepare(peh_forest* f) {
  PEH_TRY
  f->F.prepare();
  return 0;
  PEH_CATCH(PE_ERR_STATE)
}
int peh_forest_execute(peh_forest* f, int32_t* nc, int32_t* nr) {
  PEH_TRY
  auto r = f->F.execute();
  if (nc) *nc = r.first;
  if (nr) *nr = r.second;
  return 0;
  PEH_CATCH(PE_ERR_STATE)
}
static std::vector<double> vertex_values(const amr::Forest& F, const mesh::Mesh& m, const dofs::DofMap& d, const double* p) {
  if (d.degree != 1 || d.n_comp != 1) throw std::runtime_error("the estimator runs on the FE_Q(1) pressure handler (FSS:454)");
  std::vector<double> vv(F.n_vertices(), 0.0);
  const int vpc = m.vpc();
  for (int64_t c = 0; c < m.n_cells(); ++c)
    for (int k = 0; k < vpc; ++k) vv[m.cell_vertices[c * vpc + k]] = p[d.cell_dofs[c * vpc + k]];
  return vv;
}
int peh_forest_kelly(const peh_forest* f, const peh_mesh* m, const peh_dofs* d, const double* p, float* eta) {
  PEH_TRY
  std::vector<float> e = amr::kelly_estimate(f->F, vertex_values(f->F, m->m, d->d, p));
  std::memcpy(eta, e.data(), e.size() * sizeof(float));
  return 0;
  PEH_CATCH(PE_ERR_BAD_INPUT)
}
int peh_forest_mark_fixed_fraction(peh_forest* f, int64_t n, const float* criteria, double top, double bottom, int min_level, int max_level) {
  PEH_TRY
  amr::mark_fixed_fraction(f->F, std::vector<float>(criteria, criteria + n), top, bottom, min_level, max_level);
  return 0;
  PEH_CATCH(PE_ERR_BAD_INPUT)
}
int peh_forest_store(peh_forest* f, const peh_mesh* m, const peh_dofs* d, int n_vec, const double* values) {
  PEH_TRY
  std::vector<const double*> v(n_vec);
  for (int t = 0; t < n_vec; ++t) v[t] = values + (int64_t)t * d->d.n_dofs;
  f->F.store_vertex_values(m->m, d->d, n_vec, v.data());
  return 0;
  PEH_CATCH(PE_ERR_BAD_INPUT)
}
int peh_forest_fetch(const peh_forest* f, const peh_mesh* m, const peh_dofs* d, int n_vec, double* values) {
  PEH_TRY
  std::vector<double*> v(n_vec);
  for (int t = 0; t < n_vec; ++t) v[t] = values + (int64_t)t * d->d.n_dofs;
  f->F.fetch_vertex_values(m->m, d->d, n_vec, v.data());
  return 0;
  PEH_CATCH(PE_ERR_BAD_INPUT)
}
peh_constraints* peh_constraints_make(const peh_forest* f, const peh_mesh* m, const peh_dofs* d, int n, const int32_t* labels, const int32_t* comps,
                                      const double* values) {
  PEH_TRY
  auto* c = new peh_constraints();
  c->T.init(d->d.n_dofs);
  if (f) amr::hanging_node_constraints(f->F, m->m, d->d, d->maps, c->T);
  if (n > 0) {
    std::vector<int> l(labels, labels + n), cc(comps, comps + n);
    std::vector<double> v(values, values + n);
    dofs::add_dirichlet(c->T, m->m, d->d, l, cc, v);
  }
  c->T.close();
  c->T.flatten(c->line_dof, c->entry_ptr, c->entry_dof, c->entry_w, c->inhomogeneity);
  return c;
  PEH_CATCH(nullptr)
}
void peh_constraints_destroy(peh_constraints* c) { delete c; }
int peh_constraints_view_get(const peh_constraints* c, peh_constraints_view* v) {
  v->n_lines = (int64_t)c->line_dof.size();
  v->n_entries = (int64_t)c->entry_dof.size();
  v->line_dof = c->line_dof.data();
  v->entry_ptr = c->entry_ptr.data();
  v->entry_dof = c->entry_dof.data();
  v->entry_w = c->entry_w.data();
  v->inhomogeneity = c->inhomogeneity.data();
  return 0;
}

peh_part* peh_partition(const peh_mesh* m, const peh_dofs* dp, const peh_dofs* du, int rank, int nranks) {
  PEH_TRY
  auto* p = new peh_part();
  p->p = partition::make_part(m->m, dp->d, du->d, rank, nranks);
  return p;
  PEH_CATCH(nullptr)
}
peh_part* peh_partition_structured(int dim, const double* size, const int32_t* cells_per_axis, int morton_order, int rank, int nranks) {
  PEH_TRY
  auto* p = new peh_part();
  int n[3] = {cells_per_axis[0], cells_per_axis[1], dim == 3 ? cells_per_axis[2] : 1};
  p->p = partition::make_part_structured(dim, size, n, morton_order != 0, rank, nranks);
  return p;
  PEH_CATCH(nullptr)
}
void peh_part_destroy(peh_part* p) { delete p; }
int peh_part_view_get(const peh_part* p, peh_part_view* v) {
  fill_mesh_view(p->p.mesh, &v->mesh);
  v->cell_global = p->p.cell_global.data();
  v->n_owned_cells = p->p.n_owned_cells;
  for (int f = 0; f < 2; ++f) {
    const auto& F = p->p.field[f];
    auto& o = v->field[f];
    o.n_owned = F.n_owned; o.n_local = F.n_local; o.n_neighbors = (int32_t)F.neighbor_rank.size(); o.reserved = 0;
    o.cell_dofs = F.cell_dofs.data(); o.local_to_global = F.local_to_global.data();
    o.neighbor_rank = F.neighbor_rank.data(); o.send_ptr = F.send_ptr.data(); o.send_idx = F.send_idx.data(); o.recv_ptr = F.recv_ptr.data();
  }
  return 0;
}

peh_problem* peh_problem_create(const peh_input* in, int device, int rank, int nranks, const void* nccl_id, size_t id_bytes) {
  PEH_TRY
  auto* p = new peh_problem();
  p->P.reset(new poro_elastisity::PoroElasticProblem(in->d, device, rank, nranks, nccl_id, id_bytes));
  return p;
  PEH_CATCH(nullptr)
}
void peh_problem_destroy(peh_problem* p) { delete p; }
int peh_problem_initialize(peh_problem* p, int verbose) {
  PEH_TRY
  p->P->initialize(verbose != 0);
  return 0;
  PEH_CATCH(PE_ERR_STATE)
}
int peh_problem_step(peh_problem* p, int verbose, peh_step_report* out) {
  PEH_TRY
  poro_elastisity::StepReport R = p->P->step(verbose != 0);
  if (out) {
    out->time = R.time; out->time_step_number = R.time_step_number; out->fss_iterations = R.fss_iterations;
    out->pressure_iterations = R.pressure_iterations; out->cg_its_pressure = R.cg_its_pressure;
    out->cg_its_displacement = R.cg_its_displacement; out->cg_its_projection = R.cg_its_projection; out->status = 0;
    out->pressure_error = R.pressure_error; out->pressure_linfty = R.pressure_linfty;
  }
  return 0;
  PEH_CATCH(PE_ERR_STATE)
}
int peh_problem_run(peh_problem* p, int verbose) {
  PEH_TRY
  p->P->run(verbose != 0);
  return 0;
  PEH_CATCH(PE_ERR_STATE)
}
pe_ctx* peh_problem_ctx(peh_problem* p) { return p->P->ctx; }
const peh_mesh* peh_problem_mesh(peh_problem* p) {
  p->local.m = *p->P->local_mesh;
  return &p->local;
}
int peh_problem_global_ids(peh_problem* p, int field, int64_t* out) {
  const auto& g = p->P->global_ids[field];
  std::memcpy(out, g.data(), g.size() * sizeof(int64_t));
  return (int)g.size();
}

}  // extern "C"
