// GpuBackend.h — the file a maintainer of ishovkun/poroelasticity-dealii would add to lib/include to run the fixed-stress hot
// path on a B200 through libporoel.so (include/poroel.h).  The reference keeps its Triangulation / DoFHandler / ConstraintMatrix
// for mesh, numbering and boundary conditions and hands the arrays over once; after that run() calls pe_* where it called the
// three solver objects (INTEGRATION.md has the call-by-call table).  Plain extern "C": the reference is C++, no FFI layer.
//
// This is the code INTEGRATION.md quotes.  It is compiled and run by tests/test_integration_binding.py — against the deal.II API
// shim of oracle/dealii_shim (deal.II itself cannot be installed here) and the reference's own InputDataPoroel.h.
#pragma once
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include <poroel.h>

namespace gpu_backend {
using namespace dealii;

inline void check(pe_ctx* ctx, int rc, const char* what) {
  if (rc != PE_OK) throw std::runtime_error(std::string(what) + ": " + pe_last_error(ctx));
}

// setup_dofs() of the three solver classes (PS:68-111, DS:106-153, SP:82-98) -> one upload + pe_setup
template <int dim>
void upload_from_dealii(pe_ctx* ctx, const Triangulation<dim>& tria, const DoFHandler<dim>& p_dh, const DoFHandler<dim>& u_dh,
                        const ConstraintMatrix& u_constraints, const input_data::InputDataPoroel& data, int preconditioner,
                        int cg_max_iterations = 1000 /* PS:175, DS:299, SP:209; Jacobi-CG needs more iterations than SSOR-CG */) {
  // vertices and cells in active-cell order, deal.II's lexicographic vertex order
  std::vector<double> xyz;
  for (const auto& v : tria.get_vertices())
    for (int a = 0; a < dim; ++a) xyz.push_back(v[a]);
  std::vector<int32_t> cells, bcell, bid, cd_p, cd_u;
  std::vector<int8_t> bloc;
  std::vector<types::global_dof_index> idx_p(p_dh.get_fe().dofs_per_cell), idx_u(u_dh.get_fe().dofs_per_cell);
  auto pc = p_dh.begin_active();
  auto uc = u_dh.begin_active();
  int32_t c = 0;
  for (auto cell = tria.begin_active(); cell != tria.end(); ++cell, ++pc, ++uc, ++c) {
    for (unsigned v = 0; v < GeometryInfo<dim>::vertices_per_cell; ++v) cells.push_back(cell->vertex_index(v));
    for (unsigned f = 0; f < GeometryInfo<dim>::faces_per_cell; ++f)
      if (cell->face(f)->at_boundary()) { bcell.push_back(c); bloc.push_back((int8_t)f); bid.push_back(cell->face(f)->boundary_id()); }
    pc->get_dof_indices(idx_p);
    uc->get_dof_indices(idx_u);  // FE_Q / FESystem cell-local order == the ABI's
    cd_p.insert(cd_p.end(), idx_p.begin(), idx_p.end());
    cd_u.insert(cd_u.end(), idx_u.begin(), idx_u.end());
  }
  pe_params prm = {};  // ID:150-222
  prm.dim = dim;
  prm.degree_u = u_dh.get_fe().degree;
  prm.degree_p = p_dh.get_fe().degree;
  prm.lame_lambda = data.lame_constant;
  prm.shear_modulus = data.shear_modulus;
  prm.bulk_modulus = data.bulk_modulus;
  prm.biot_coef = data.biot_coef;
  prm.m_modulus = data.m_modulus;
  prm.perm_over_visc = data.perm / data.visc;
  prm.well_radius = data.r_well;
  prm.flow_rate = data.flow_rate;
  prm.cg_max_iterations = cg_max_iterations;
  prm.cg_rel_tol_pressure = 1e-8;        // PS:175
  prm.cg_abs_tol_displacement = 1e-12;   // DS:298
  prm.cg_rel_tol_projection = 1e-8;      // SP:209
  prm.preconditioner = preconditioner;   // the device replaces SSOR by Jacobi or a Chebyshev-Jacobi polynomial
  prm.chebyshev_degree = 4;
  prm.chebyshev_eig_ratio = 30.0;
  check(ctx, pe_set_params(ctx, &prm), "pe_set_params");
  check(ctx, pe_upload_mesh(ctx, dim, tria.n_vertices(), xyz.data(), tria.n_active_cells(), cells.data(), (int64_t)bcell.size(), bcell.data(),
                            bloc.data(), bid.data()), "pe_upload_mesh");
  check(ctx, pe_upload_dofs(ctx, PE_FIELD_PRESSURE, p_dh.n_dofs(), cd_p.data()), "pe_upload_dofs(p)");
  check(ctx, pe_upload_dofs(ctx, PE_FIELD_DISPLACEMENT, u_dh.n_dofs(), cd_u.data()), "pe_upload_dofs(u)");
  // the closed ConstraintMatrix as a flat table: Dirichlet lines have no entries, hanging-node lines (adaptive meshes, FSS:333-340)
  // carry (master, weight) pairs; the pressure handler's table (PS:71-78) goes through the same call
  std::vector<int32_t> line, edof;
  std::vector<double> g, ew;
  std::vector<int64_t> eptr(1, 0);
  for (types::global_dof_index i = 0; i < u_dh.n_dofs(); ++i)
    if (u_constraints.is_constrained(i)) {
      line.push_back((int32_t)i);
      g.push_back(u_constraints.get_inhomogeneity(i));
      if (const auto* entries = u_constraints.get_constraint_entries(i))
        for (const auto& e : *entries) { edof.push_back((int32_t)e.first); ew.push_back(e.second); }
      eptr.push_back((int64_t)edof.size());
    }
  check(ctx, pe_upload_constraints(ctx, PE_FIELD_DISPLACEMENT, (int64_t)line.size(), line.data(), eptr.data(), edof.data(), ew.data(), g.data()),
        "pe_upload_constraints");
  std::vector<int32_t> nl(data.stress_boundary_labels.begin(), data.stress_boundary_labels.end()),
      nc(data.stress_boundary_components.begin(), data.stress_boundary_components.end());
  check(ctx, pe_upload_neumann(ctx, (int)nl.size(), nl.data(), nc.data(), data.stress_boundary_values.data()), "pe_upload_neumann");
  check(ctx, pe_setup(ctx), "pe_setup");
}

}  // namespace gpu_backend
