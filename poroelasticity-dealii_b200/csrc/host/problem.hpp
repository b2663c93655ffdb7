// problem.hpp — the solver driver the reference lacks a main() for.
//
// poro_elastisity::PoroElasticProblem<dim> (lib/include/PoroelasticityFSS.h:49-90) re-written as
// plain host C++ whose only numerical back end is the CUDA device library behind
// include/poroel.h.  The loop structure, the order of operator calls, the convergence tests and the
// log lines are those of PoroElasticProblem::run (FSS:294-415), with the as-is quirks kept:
//   * mechanics -> flow feedback off unless `Couple volumetric strain = 1` (FSS:399 is commented out);
//   * the initial volumetric strain is the reference state for all steps (FSS:317, PS:122-124);
//   * create_mesh() is the default, read_mesh() optional (FSS:297-298).
// Adaptive refinement (FSS:333-340, refine_mesh FSS:447-498) runs every `Refine every` steps (default 5 = the reference's schedule; 0 =
// never, the configuration of the BASELINE benchmarks) on one rank: amr.hpp provides the refinement forest, the Kelly
// estimator, the marking and the solution transfer; hanging-node constraints go to the device as general lines.
#pragma once
#include <cstdio>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../../include/poroel.h"
#include "amr.hpp"
#include "dofs.hpp"
#include "input_data.hpp"
#include "mesh.hpp"
#include "partition.hpp"

namespace poro_elastisity {

struct StepReport {
  double time = 0;
  int time_step_number = 0, fss_iterations = 0, pressure_iterations = 0;
  int cg_its_pressure = 0, cg_its_displacement = 0, cg_its_projection = 0;
  int status = 0;
  double pressure_error = 0, pressure_linfty = 0;
};

inline pe_params params_from_input(const input_data::InputDataPoroel& d) {
  pe_params p{};
  p.dim = d.dim;
  p.degree_u = d.displacement_degree;
  p.degree_p = 1;  // PS:20
  p.preconditioner = d.preconditioner;
  p.chebyshev_degree = d.chebyshev_degree;
  p.cg_max_iterations = d.cg_max_iterations;
  p.cg_check_interval = 0;
  p.lame_lambda = d.lame_constant;
  p.shear_modulus = d.shear_modulus;
  p.bulk_modulus = d.bulk_modulus;
  p.biot_coef = d.biot_coef;
  p.m_modulus = d.m_modulus;
  p.perm_over_visc = d.perm / d.visc;
  p.well_radius = d.r_well;
  p.flow_rate = d.flow_rate;
  p.cg_rel_tol_pressure = 1e-8;       // PS:175
  p.cg_abs_tol_displacement = 1e-12;  // DS:298
  p.cg_rel_tol_projection = 1e-8;     // SP:209
  p.chebyshev_eig_ratio = d.chebyshev_eig_ratio;
  return p;
}

class PoroElasticProblem {
 public:
  PoroElasticProblem(const input_data::InputDataPoroel& data_, int device, int rank_, int nranks_, const void* nccl_id, size_t id_bytes)
      : data(data_), rank(rank_), nranks(nranks_) {
    dim = data.dim;
    // FSS:92-124
    n_stress_components = (dim * dim + dim) / 2;
    if (dim == 2) { strain_tensor_volumetric_components = {0, 3}; strain_tensor_shear_components = {1}; }
    else if (dim == 3) { strain_tensor_volumetric_components = {0, 4, 8}; strain_tensor_shear_components = {1, 2, 5}; }
    else throw std::runtime_error("PoroElasticProblem: dim must be 2 or 3");
    for (int c : strain_tensor_volumetric_components) strain_rhs_volumetric_entries.push_back(entryIndex(c));
    int rc = pe_create(&ctx, device, rank, nranks, nccl_id, id_bytes);
    if (rc != 0) throw std::runtime_error(std::string("pe_create: ") + pe_last_error(nullptr));
  }
  ~PoroElasticProblem() { if (ctx) pe_destroy(ctx); }
  PoroElasticProblem(const PoroElasticProblem&) = delete;

  // TensorIndexer<dim>::entryIndex (TensorIndexer.h:18-52)
  int entryIndex(int tensor_index) const {
    static const int m2[4] = {0, 1, 1, 2}, m3[9] = {0, 1, 2, 1, 3, 4, 2, 4, 5};
    return dim == 2 ? m2[tensor_index] : m3[tensor_index];
  }

  void check(int rc, const char* what) {
    if (rc != 0) throw std::runtime_error(std::string(what) + ": " + pe_last_error(ctx) + " (status " + std::to_string(rc) + ")");
  }

  // Partitioned runs on structured FE_Q(1) boxes build only this rank's part (partition.hpp::make_part_structured): no
  // global mesh, no global dof maps.  PE_STRUCTURED_PART=0 forces the general path.
  bool structured_fast_path() const {
    const char* e = std::getenv("PE_STRUCTURED_PART");
    return nranks > 1 && nranks <= 32 && data.refine_every == 0 && !data.mesh_from_file && data.displacement_degree == 1 && !(e && e[0] == '0');
  }
  void setup_dofs_structured() {
    int n[3] = {1, 1, 1};
    const bool morton = data.cells_per_axis[0] <= 0;
    for (int a = 0; a < dim; ++a) n[a] = morton ? (1 << data.initial_refinement_level) : data.cells_per_axis[a];
    part = partition::make_part_structured(dim, data.domain_size.data(), n, morton, rank, nranks);
    n_global_p = partition::structured_vertex_count(dim, n);
    n_global_u = n_global_p * dim;
    pe_params prm = params_from_input(data);
    check(pe_set_params(ctx, &prm), "pe_set_params");
    const mesh::Mesh* lm = &part.mesh;
    // Dirichlet lines (DS:117-135) straight on the local sub-mesh: a local dof on the domain boundary always lies on a
    // boundary face of a LOCAL cell of a box mesh, so this equals the global table restricted to the local dofs
    dofs::DofMap du_local;
    du_local.degree = 1;
    du_local.n_comp = dim;
    du_local.n_loc = (1 << dim) * dim;
    du_local.n_dofs = part.field[1].n_local;
    du_local.cell_dofs = part.field[1].cell_dofs;
    dofs::Constraints cons = dofs::make_dirichlet(*lm, du_local, data.displacement_boundary_labels, data.displacement_boundary_components,
                                                  data.displacement_boundary_values);
    std::vector<std::pair<int32_t, double>> ll;
    for (size_t i = 0; i < cons.line_dof.size(); ++i) ll.push_back({cons.line_dof[i], cons.inhomogeneity[i]});
    std::sort(ll.begin(), ll.end());
    std::vector<int32_t> line_dof;
    std::vector<double> line_g;
    for (auto& e : ll) { line_dof.push_back(e.first); line_g.push_back(e.second); }
    std::vector<int64_t> entry_ptr(line_dof.size() + 1, 0);
    for (int f = 0; f < 2; ++f)
      global_ids[f] = std::vector<int64_t>(part.field[f].local_to_global.begin(), part.field[f].local_to_global.begin() + part.field[f].n_owned);
    check(pe_upload_mesh(ctx, dim, lm->n_vertices(), lm->xyz.data(), lm->n_cells(), lm->cell_vertices.data(), lm->n_bfaces(),
                         lm->bface_cell.data(), lm->bface_local.data(), lm->bface_id.data()), "pe_upload_mesh");
    check(pe_upload_dofs(ctx, PE_FIELD_PRESSURE, part.field[0].n_local, part.field[0].cell_dofs.data()), "pe_upload_dofs(p)");
    check(pe_upload_dofs(ctx, PE_FIELD_DISPLACEMENT, part.field[1].n_local, part.field[1].cell_dofs.data()), "pe_upload_dofs(u)");
    check(pe_upload_constraints(ctx, PE_FIELD_DISPLACEMENT, (int64_t)line_dof.size(), line_dof.data(), entry_ptr.data(), nullptr, nullptr,
                                line_g.data()), "pe_upload_constraints(u)");
    std::vector<int32_t> nl(data.stress_boundary_labels.begin(), data.stress_boundary_labels.end()),
        ncmp(data.stress_boundary_components.begin(), data.stress_boundary_components.end());
    check(pe_upload_neumann(ctx, (int)nl.size(), nl.data(), ncmp.data(), data.stress_boundary_values.data()), "pe_upload_neumann");
    for (int f = 0; f < 2; ++f) {
      auto& F = part.field[f];
      check(pe_upload_partition(ctx, f, F.n_owned, (int)F.neighbor_rank.size(), F.neighbor_rank.data(), F.send_ptr.data(), F.send_idx.data(),
                                F.recv_ptr.data()), "pe_upload_partition");
    }
    local_mesh = lm;
    check(pe_setup(ctx), "pe_setup");
  }

  // setup_dofs() (FSS:131-151) on the current mesh + upload + pe_setup
  void setup_dofs() {
    if (structured_fast_path()) { setup_dofs_structured(); return; }
    const bool adaptive = data.refine_every != 0;
    // distribute_dofs for both handlers (PS:73, DS:110)
    dofs::NodeMaps maps_p, maps_u;
    dofs_p = dofs::distribute_dofs(global_mesh, 1, 1, adaptive ? &maps_p : nullptr);
    dofs::DofMap du = dofs::distribute_dofs(global_mesh, data.displacement_degree, dim, adaptive ? &maps_u : nullptr);
    const dofs::DofMap& dp = dofs_p;
    n_global_p = dp.n_dofs;
    n_global_u = du.n_dofs;
    pe_params prm = params_from_input(data);
    check(pe_set_params(ctx, &prm), "pe_set_params");
    const mesh::Mesh* lm = &global_mesh;
    const int32_t *cdp = dp.cell_dofs.data(), *cdu = du.cell_dofs.data();
    int64_t nlp = dp.n_dofs, nlu = du.n_dofs;
    std::vector<int32_t> line_dof, entry_dof, p_line_dof, p_entry_dof;
    std::vector<int64_t> entry_ptr, p_entry_ptr;
    std::vector<double> line_g, entry_w, p_line_g, p_entry_w;
    if (adaptive) {
      // hanging-node lines first, then the Dirichlet values on dofs that are still free, then close (PS:71-78, DS:109-137)
      dofs::ConstraintTable tp, tu;
      tp.init(dp.n_dofs);
      amr::hanging_node_constraints(forest, global_mesh, dp, maps_p, tp);
      tp.close();
      tp.flatten(p_line_dof, p_entry_ptr, p_entry_dof, p_entry_w, p_line_g);
      tu.init(du.n_dofs);
      amr::hanging_node_constraints(forest, global_mesh, du, maps_u, tu);
      dofs::add_dirichlet(tu, global_mesh, du, data.displacement_boundary_labels, data.displacement_boundary_components,
                          data.displacement_boundary_values);
      tu.close();
      tu.flatten(line_dof, entry_ptr, entry_dof, entry_w, line_g);
      n_hanging_p = (int64_t)p_line_dof.size();
    } else {
      // displacement_solver.set_boundary_conditions (FSS:300-306) -> Dirichlet lines (DS:117-135)
      dofs::Constraints cons = dofs::make_dirichlet(global_mesh, du, data.displacement_boundary_labels,
                                                    data.displacement_boundary_components, data.displacement_boundary_values);
      line_dof = cons.line_dof;
      line_g = cons.inhomogeneity;
      if (nranks > 1) {
        part = partition::make_part(global_mesh, dp, du, rank, nranks);
        lm = &part.mesh;
        cdp = part.field[0].cell_dofs.data();
        cdu = part.field[1].cell_dofs.data();
        nlp = part.field[0].n_local;
        nlu = part.field[1].n_local;
        std::vector<int32_t> g2l(du.n_dofs, -1);
        for (int64_t i = 0; i < nlu; ++i) g2l[part.field[1].local_to_global[i]] = (int32_t)i;
        line_dof.clear();
        line_g.clear();
        std::vector<std::pair<int32_t, double>> ll;
        for (size_t i = 0; i < cons.line_dof.size(); ++i)
          if (g2l[cons.line_dof[i]] >= 0) ll.push_back({g2l[cons.line_dof[i]], cons.inhomogeneity[i]});
        std::sort(ll.begin(), ll.end());
        for (auto& e : ll) { line_dof.push_back(e.first); line_g.push_back(e.second); }
      }
      entry_ptr.assign(line_dof.size() + 1, 0);
    }
    if (nranks > 1) {
      global_ids[0] = std::vector<int64_t>(part.field[0].local_to_global.begin(), part.field[0].local_to_global.begin() + part.field[0].n_owned);
      global_ids[1] = std::vector<int64_t>(part.field[1].local_to_global.begin(), part.field[1].local_to_global.begin() + part.field[1].n_owned);
    } else {
      global_ids[0].resize(nlp);
      global_ids[1].resize(nlu);
      for (int64_t i = 0; i < nlp; ++i) global_ids[0][i] = i;
      for (int64_t i = 0; i < nlu; ++i) global_ids[1][i] = i;
    }
    check(pe_upload_mesh(ctx, dim, lm->n_vertices(), lm->xyz.data(), lm->n_cells(), lm->cell_vertices.data(), lm->n_bfaces(),
                         lm->bface_cell.data(), lm->bface_local.data(), lm->bface_id.data()), "pe_upload_mesh");
    check(pe_upload_dofs(ctx, PE_FIELD_PRESSURE, nlp, cdp), "pe_upload_dofs(p)");
    check(pe_upload_dofs(ctx, PE_FIELD_DISPLACEMENT, nlu, cdu), "pe_upload_dofs(u)");
    if (adaptive)
      check(pe_upload_constraints(ctx, PE_FIELD_PRESSURE, (int64_t)p_line_dof.size(), p_line_dof.data(), p_entry_ptr.data(), p_entry_dof.data(),
                                  p_entry_w.data(), p_line_g.data()), "pe_upload_constraints(p)");
    check(pe_upload_constraints(ctx, PE_FIELD_DISPLACEMENT, (int64_t)line_dof.size(), line_dof.data(), entry_ptr.data(), entry_dof.data(),
                                entry_w.data(), line_g.data()), "pe_upload_constraints(u)");
    std::vector<int32_t> nl(data.stress_boundary_labels.begin(), data.stress_boundary_labels.end()),
        ncmp(data.stress_boundary_components.begin(), data.stress_boundary_components.end());
    check(pe_upload_neumann(ctx, (int)nl.size(), nl.data(), ncmp.data(), data.stress_boundary_values.data()), "pe_upload_neumann");
    if (nranks > 1)
      for (int f = 0; f < 2; ++f) {
        auto& F = part.field[f];
        check(pe_upload_partition(ctx, f, F.n_owned, (int)F.neighbor_rank.size(), F.neighbor_rank.data(), F.send_ptr.data(), F.send_idx.data(),
                                  F.recv_ptr.data()), "pe_upload_partition");
      }
    local_mesh = lm;
    check(pe_setup(ctx), "pe_setup");
  }

  // FSS:297-317
  void initialize(bool verbose) {
    if (data.refine_every < 0) throw std::runtime_error("'Refine every' must be >= 0");
    if (data.refine_every != 0 && nranks != 1)
      throw std::runtime_error("adaptive refinement (FSS:333-340) runs on one rank: set 'Refine every = 0' for partitioned runs");
    // create_mesh() / read_mesh()
    if (structured_fast_path()) {
      // nothing global is built: setup_dofs_structured() derives this rank's part from the lattice
    } else if (data.mesh_from_file) {
      global_mesh = mesh::read_msh(data.mesh_file, dim);
      // partitioned runs on unstructured meshes: contiguous cell ranges of a space-filling-curve order are compact
      // subdomains (every rank computes the same order); single-rank runs keep the file order of GridIn::read_msh
      if (nranks > 1) mesh::reorder_cells_sfc(global_mesh);
    } else if (data.cells_per_axis[0] > 0) global_mesh = mesh::create_subdivided(dim, data.domain_size.data(), data.cells_per_axis);
    else global_mesh = mesh::create_hyper_rectangle(dim, data.domain_size.data(), data.initial_refinement_level);
    if (data.refine_every != 0) {
      // the cells of the initial mesh are the roots of the refinement forest; for create_mesh() they sit at level
      // `Initial refinement level`, which FSS:335 also passes as the coarsest level the estimator may go back to
      const bool refined_box = !data.mesh_from_file && data.cells_per_axis[0] <= 0;
      forest = amr::Forest::from_mesh(global_mesh, refined_box ? data.initial_refinement_level : 0);
      global_mesh = forest.active_mesh();
    }
    setup_dofs();

    // Initialize reservoir (FSS:310-317)
    check(pe_pressure_set_uniform(ctx, data.p_init), "pressure_set_uniform");
    check(pe_displacement_assemble(ctx), "displacement_assemble");
    int its = 0;
    double res = 0;
    check(pe_displacement_solve(ctx, &its, &res), "displacement_solve");
    if (verbose && rank == 0) std::printf("initial displacement solve: %d CG iterations, residual %.3e\n", its, res);
    check(pe_project_assemble_matrix(ctx), "project_assemble_matrix");
    int pits = 0;
    get_normal_strain_components(&pits);
    get_volumetric_strain(/*as_initial=*/true);
    time = 0;
    time_step_number = 0;
  }

  // refine_mesh (FSS:447-498)
  void refine_mesh(int min_grid_level, int max_grid_level) {
    const int64_t n_old = n_global_p;
    std::vector<double> p(n_old), ev(n_old), ev0(n_old);
    check(pe_get_vector(ctx, PE_VEC_P, p.data(), n_old), "get p");
    check(pe_get_vector(ctx, PE_VEC_VOL_STRAIN, ev.data(), n_old), "get volumetric_strain");
    check(pe_get_vector(ctx, PE_VEC_VOL_STRAIN0, ev0.data(), n_old), "get initial_volumetric_strain");
    // KellyErrorEstimator on the pressure solution (FSS:452-458)
    std::vector<double> vertex_value(forest.n_vertices(), 0.0);
    const int vpc = global_mesh.vpc();
    for (int64_t c = 0; c < global_mesh.n_cells(); ++c)
      for (int k = 0; k < vpc; ++k) vertex_value[global_mesh.cell_vertices[c * vpc + k]] = p[dofs_p.cell_dofs[c * vpc + k]];
    std::vector<float> estimated_error_per_cell = amr::kelly_estimate(forest, vertex_value);
    amr::mark_fixed_fraction(forest, estimated_error_per_cell, 0.6, 0.4, min_grid_level, max_grid_level);  // FSS:460-472
    // SolutionTransfer of {pressure, volumetric strain, initial volumetric strain} (FSS:475-497)
    const double* in[3] = {p.data(), ev.data(), ev0.data()};
    forest.store_vertex_values(global_mesh, dofs_p, 3, in);
    auto counts = forest.execute();  // prepare_coarsening_and_refinement + execute_coarsening_and_refinement
    last_refinement = counts;
    global_mesh = forest.active_mesh();
    setup_dofs();
    std::vector<double> q(n_global_p), qv(n_global_p), qv0(n_global_p);
    double* out[3] = {q.data(), qv.data(), qv0.data()};
    forest.fetch_vertex_values(global_mesh, dofs_p, 3, out);
    check(pe_set_vector(ctx, PE_VEC_P, q.data(), n_global_p), "set p");
    check(pe_set_vector(ctx, PE_VEC_VOL_STRAIN, qv.data(), n_global_p), "set volumetric_strain");
    check(pe_set_vector(ctx, PE_VEC_VOL_STRAIN0, qv0.data(), n_global_p), "set initial_volumetric_strain");
  }

  // FSS:153-164
  void get_normal_strain_components(int* cg_its) {
    std::vector<int32_t> comps(strain_tensor_volumetric_components.begin(), strain_tensor_volumetric_components.end());
    check(pe_project_assemble_rhs(ctx, (int)comps.size(), comps.data()), "project_assemble_rhs");
    for (int comp : strain_tensor_volumetric_components) {
      int its = 0;
      check(pe_project_solve(ctx, entryIndex(comp), &its), "project_solve");
      if (cg_its) *cg_its += its;
    }
  }
  // FSS:167-176
  void get_shear_strain_components() {
    std::vector<int32_t> comps(strain_tensor_shear_components.begin(), strain_tensor_shear_components.end());
    // the reference never assembles the shear right-hand sides (projection_rhs stays zero for them);
    // here they are assembled so the optional stress output is meaningful.
    check(pe_project_assemble_rhs(ctx, (int)comps.size(), comps.data()), "project_assemble_rhs(shear)");
    for (int comp : strain_tensor_shear_components) {
      int its = 0;
      check(pe_project_solve(ctx, entryIndex(comp), &its), "project_solve(shear)");
    }
  }
  // FSS:179-186 (+ FSS:317 when as_initial)
  void get_volumetric_strain(bool as_initial) {
    std::vector<int32_t> e(strain_rhs_volumetric_entries.begin(), strain_rhs_volumetric_entries.end());
    check(pe_volumetric_strain_from_projection(ctx, (int)e.size(), e.data(), as_initial ? 1 : 0), "volumetric_strain");
  }

  // One pass of FSS:328-407
  StepReport step(bool verbose) {
    StepReport R;
    const double time_step = data.time_step;
    time += time_step;
    time_step_number++;
    R.time = time;
    R.time_step_number = time_step_number;
    const bool out = verbose && rank == 0;
    if (out) std::printf("Time: %g\n", time);
    if (data.refine_every > 0 && time_step_number % data.refine_every == 0) {  // FSS:333-340
      if (out) std::printf("Refining mesh\n");
      refine_mesh(data.initial_refinement_level, data.initial_refinement_level + data.max_refinement_level);
      check(pe_displacement_assemble(ctx), "displacement_assemble");
      check(pe_project_assemble_matrix(ctx), "project_assemble_matrix");
      if (out)
        std::printf("    %lld active cells, %lld pressure dofs (%lld hanging), %d families coarsened, %d cells refined\n",
                    (long long)global_mesh.n_cells(), (long long)n_global_p, (long long)n_hanging_p, last_refinement.first, last_refinement.second);
    }
    check(pe_pressure_begin_step(ctx), "begin_step");  // FSS:342
    double pressure_error = data.pressure_tol * 2;      // FSS:345
    int fss_iteration = 0;
    while (fss_iteration < data.max_fss_iterations && pressure_error > data.fss_tol) {
      fss_iteration++;
      if (out) std::printf("    Coupling iteration: %d\n", fss_iteration);
      int pressure_iteration = 0;
      check(pe_pressure_zero_update(ctx), "zero_update");  // FSS:356
      while (pressure_iteration < data.max_pressure_iterations) {
        pressure_iteration++;
        R.pressure_iterations++;
        check(pe_pressure_update_volumetric_strain(ctx), "update_volumetric_strain");
        check(pe_pressure_assemble_residual(ctx, time_step, &pressure_error), "assemble_residual");
        if (pressure_error < data.pressure_tol) {
          if (out) std::printf("        pressure converged; iterations: %d\n", pressure_iteration - 1);
          break;
        }
        check(pe_pressure_assemble_jacobian(ctx, time_step), "assemble_jacobian");
        int its = 0;
        double res = 0;
        check(pe_pressure_solve(ctx, &its, &res), "pressure_solve");
        R.cg_its_pressure += its;
        check(pe_pressure_add_update(ctx), "add_update");  // FSS:379
      }
      check(pe_pressure_linfty(ctx, &R.pressure_linfty), "linfty");
      if (out) std::printf("Solution limits: %g\t\n", R.pressure_linfty);
      // Solve displacement system (FSS:395-396)
      check(pe_displacement_assemble(ctx), "displacement_assemble");
      int its = 0;
      double res = 0;
      check(pe_displacement_solve(ctx, &its, &res), "displacement_solve");
      R.cg_its_displacement += its;
      get_normal_strain_components(&R.cg_its_projection);  // FSS:398
      if (data.couple_volumetric_strain) get_volumetric_strain(false);  // FSS:399 (commented out in the reference)
      check(pe_pressure_assemble_residual(ctx, time_step, &pressure_error), "assemble_residual");  // FSS:402-405
      if (out) std::printf("        Error: %g\n", pressure_error);
    }
    R.fss_iterations = fss_iteration;
    R.pressure_error = pressure_error;
    return R;
  }

  // FSS:294-415
  void run(bool verbose) {
    initialize(verbose);
    if (verbose && rank == 0) {
      std::printf("starting time loop\n");
      std::printf("time max %g\n", data.t_max);
    }
    while (time < data.t_max) {
      StepReport R = step(verbose);
      (void)R;
      if (data.write_vtk) {
        get_shear_strain_components();                      // FSS:409
        check(pe_effective_stresses(ctx), "stresses");      // FSS:410
        output_results(time_step_number);                   // FSS:411
      }
      if (data.max_time_steps > 0 && time_step_number >= data.max_time_steps) break;
    }
  }

  // FSS:227-291 — minimal legacy-VTK writer (Q1 patches of the active mesh, vertex values of every field; single rank only)
  void output_results(unsigned int n) {
    if (nranks != 1) return;
    const mesh::Mesh& m = global_mesh;
    const dofs::DofMap& dp = dofs_p;
    std::vector<double> p(n_global_p), u(n_global_u);
    check(pe_get_vector(ctx, PE_VEC_P, p.data(), n_global_p), "get p");
    check(pe_get_vector(ctx, PE_VEC_U, u.data(), n_global_u), "get u");
    std::vector<int32_t> v2d(m.n_vertices(), -1), v2u(m.n_vertices(), -1);
    const int vpc = m.vpc();
    // the first 2^dim scalar nodes of FE_Q(k) are the vertices, so the vertex values of u are dofs for k = 1 and k = 2
    dofs::DofMap du = dofs::distribute_dofs(m, data.displacement_degree, dim);
    for (int64_t c = 0; c < m.n_cells(); ++c)
      for (int v = 0; v < vpc; ++v) {
        v2d[m.cell_vertices[c * vpc + v]] = dp.cell_dofs[c * vpc + v];
        v2u[m.cell_vertices[c * vpc + v]] = du.cell_dofs[c * du.n_loc + v * dim];
      }
    // the forest keeps vertices of coarsened cells: write only the ones in use
    std::vector<int32_t> vnew(m.n_vertices(), -1);
    std::vector<int64_t> used;
    for (int64_t v = 0; v < m.n_vertices(); ++v)
      if (v2d[v] >= 0) { vnew[v] = (int32_t)used.size(); used.push_back(v); }
    char name[256];
    std::snprintf(name, sizeof name, "./solution/solution-%04u.vtk", n);
    FILE* f = std::fopen(name, "w");
    if (!f) return;
    const long long nv = (long long)used.size();
    std::fprintf(f, "# vtk DataFile Version 3.0\nporoelasticity fixed-stress\nASCII\nDATASET UNSTRUCTURED_GRID\nPOINTS %lld double\n", nv);
    for (int64_t v : used) std::fprintf(f, "%.12g %.12g %.12g\n", m.xyz[v * dim], m.xyz[v * dim + 1], dim == 3 ? m.xyz[v * dim + 2] : 0.0);
    std::fprintf(f, "CELLS %lld %lld\n", (long long)m.n_cells(), (long long)(m.n_cells() * (vpc + 1)));
    static const int vtk2[4] = {0, 1, 3, 2}, vtk3[8] = {0, 1, 3, 2, 4, 5, 7, 6};
    for (int64_t c = 0; c < m.n_cells(); ++c) {
      std::fprintf(f, "%d", vpc);
      for (int v = 0; v < vpc; ++v) std::fprintf(f, " %d", vnew[m.cell_vertices[c * vpc + (dim == 2 ? vtk2[v] : vtk3[v])]]);
      std::fprintf(f, "\n");
    }
    std::fprintf(f, "CELL_TYPES %lld\n", (long long)m.n_cells());
    for (int64_t c = 0; c < m.n_cells(); ++c) std::fprintf(f, "%d\n", dim == 2 ? 9 : 12);
    std::fprintf(f, "POINT_DATA %lld\nVECTORS u double\n", nv);
    for (int64_t v : used) {
      const int64_t d0 = v2u[v];
      std::fprintf(f, "%.12g %.12g %.12g\n", u[d0], u[d0 + 1], dim == 3 ? u[d0 + 2] : 0.0);
    }
    std::fprintf(f, "SCALARS p double 1\nLOOKUP_TABLE default\n");
    for (int64_t v : used) std::fprintf(f, "%.12g\n", p[v2d[v]]);
    static const char* names2[3] = {"xx", "xy", "yy"};
    static const char* names3[6] = {"xx", "xy", "xz", "yy", "yz", "zz"};
    std::vector<double> s(n_global_p);
    for (int pass = 0; pass < 2; ++pass)
      for (int e = 0; e < n_stress_components; ++e) {
        check(pe_get_vector(ctx, (pass == 0 ? PE_VEC_STRAIN0 : PE_VEC_STRESS0) + e, s.data(), n_global_p), "get strain/stress");
        std::fprintf(f, "SCALARS %s_%s double 1\nLOOKUP_TABLE default\n", pass == 0 ? "eps" : "sigma", dim == 2 ? names2[e] : names3[e]);
        for (int64_t v : used) std::fprintf(f, "%.12g\n", s[v2d[v]]);
      }
    std::fclose(f);
  }

  pe_ctx* ctx = nullptr;
  input_data::InputDataPoroel data;
  int dim = 2, rank = 0, nranks = 1;
  mesh::Mesh global_mesh;
  amr::Forest forest;          // refinement tree of the adaptive time loop (empty unless 'Refine every' > 0)
  dofs::DofMap dofs_p;         // pressure handler of the current mesh
  int64_t n_hanging_p = 0;
  std::pair<int, int> last_refinement{0, 0};
  const mesh::Mesh* local_mesh = nullptr;
  partition::Part part;
  std::vector<int64_t> global_ids[2];
  int64_t n_global_p = 0, n_global_u = 0;
  double time = 0;
  int time_step_number = 0;
  std::vector<int> strain_tensor_volumetric_components, strain_rhs_volumetric_entries, strain_tensor_shear_components;
  int n_stress_components = 0;
};

}  // namespace poro_elastisity
