// fss_gpu.cpp — PoroElasticProblem<dim>::run() of the reference (PoroelasticityFSS.h:294-415) with every call into the three
// solver objects replaced by the pe_* entry point INTEGRATION.md lists for it, and setup_dofs() replaced by
// gpu_backend::upload_from_dealii (GpuBackend.h).  Mesh, dof numbering, boundary conditions and the parameter file stay what
// they are in the reference: deal.II objects and the reference's own InputDataPoroel.h.  What a maintainer's patched
// PoroelasticityFSS.h would look like, as one translation unit:
//     fss_gpu <input.data> [preconditioner: 0 Jacobi | 1 Chebyshev-Jacobi (default)] [CG iteration cap, default the reference's 1000]
// prints the reference's log and writes ./solution/fields-NNNN.txt (p and u, dof by dof) after every time step.
// Built by tests/test_integration_binding.py against the deal.II API shim (oracle/dealii_shim; deal.II cannot be installed here),
// linked once with the oracle-backed pe_* of tests/driver_on_oracle.cpp (CPU) and once with libporoel.so (GPU box).  The
// adaptive branch (FSS:333-340) is the one INTEGRATION.md describes; it is not exercised here (the shim has no refinement).
#include <deal.II/base/function.h>
#include <deal.II/dofs/dof_handler.h>
#include <deal.II/fe/fe_q.h>
#include <deal.II/fe/fe_system.h>
#include <deal.II/grid/grid_generator.h>
#include <deal.II/grid/tria.h>
#include <deal.II/lac/constraint_matrix.h>
#include <deal.II/numerics/vector_tools.h>

#include <InputDataPoroel.h>   // the reference's own parameter reader, unmodified
#include <TensorIndexer.h>     // and its tensor index map

#include "GpuBackend.h"

using namespace dealii;

template <int dim>
class PoroElasticProblemGpu {
 public:
  PoroElasticProblemGpu(input_data::InputDataPoroel& data_, int preconditioner, int cg_cap)
      : data(data_), p_dh(triangulation), u_dh(triangulation), p_fe(1), u_fe(FE_Q<dim>(2), dim), preconditioner(preconditioner), cg_cap(cg_cap) {
    if (pe_create(&ctx, 0, 0, 1, nullptr, 0) != PE_OK) throw std::runtime_error(std::string("pe_create: ") + pe_last_error(nullptr));
    switch (dim) {  // FSS:100-110
      case 2: volumetric = {0, 3}; break;
      case 3: volumetric = {0, 4, 8}; break;
    }
    for (int c : volumetric) volumetric_entries.push_back(tensor_indexer.entryIndex(c));  // FSS:116-123
  }
  ~PoroElasticProblemGpu() { pe_destroy(ctx); }

  void run() {
    create_mesh();
    setup_dofs();
    // Initialize reservoir (FSS:310-317)
    ck(pe_pressure_set_uniform(ctx, data.p_init));
    ck(pe_displacement_assemble(ctx));
    solve_displacement();
    ck(pe_project_assemble_matrix(ctx));
    get_normal_strain_components();
    ck(pe_volumetric_strain_from_projection(ctx, (int)volumetric_entries.size(), volumetric_entries.data(), 1));

    double time = 0;
    const double time_step = data.time_step;
    unsigned int time_step_number = 0;
    double pressure_error;
    std::cout << "starting time loop" << std::endl;
    std::cout << "time max " << data.t_max << std::endl;
    while (time < data.t_max) {
      time += time_step;
      time_step_number++;
      std::cout << "Time: " << time << std::endl;
      if (time_step_number % 5 == 0) throw std::runtime_error("refinement (FSS:333-340) is not part of this driver");
      ck(pe_pressure_begin_step(ctx));                                         // FSS:342
      pressure_error = data.pressure_tol * 2;
      int fss_iteration = 0;
      while (fss_iteration < data.max_fss_iterations && pressure_error > data.fss_tol) {
        fss_iteration++;
        std::cout << "    Coupling iteration: " << fss_iteration << std::endl;
        int pressure_iteration = 0;
        ck(pe_pressure_zero_update(ctx));                                      // FSS:356
        while (pressure_iteration < data.max_pressure_iterations) {
          pressure_iteration++;
          ck(pe_pressure_update_volumetric_strain(ctx));                       // FSS:360
          ck(pe_pressure_assemble_residual(ctx, time_step, &pressure_error));  // FSS:361-364
          if (pressure_error < data.pressure_tol) {
            std::cout << "        pressure converged; iterations: " << pressure_iteration - 1 << std::endl;
            break;
          }
          ck(pe_pressure_assemble_jacobian(ctx, time_step));                   // FSS:377
          int its; double res;
          ck(pe_pressure_solve(ctx, &its, &res));                              // FSS:378 (PE_ERR_NO_CONVERGENCE = SolverControl::NoConvergence)
          ck(pe_pressure_add_update(ctx));                                     // FSS:379
        }
        double linfty;
        ck(pe_pressure_linfty(ctx, &linfty));
        std::cout << "Solution limits: " << linfty << "\t" << std::endl;      // FSS:387-389
        ck(pe_displacement_assemble(ctx));                                     // FSS:395
        solve_displacement();                                                  // FSS:396
        get_normal_strain_components();                                        // FSS:398
        ck(pe_pressure_assemble_residual(ctx, time_step, &pressure_error));    // FSS:402-405
        std::cout << "        Error: " << pressure_error << std::endl;
      }
      output_results(time_step_number);
    }
  }

 private:
  void ck(int rc) { gpu_backend::check(ctx, rc, "pe_*"); }
  void create_mesh() {  // FSS:418-435, unchanged
    Tensor<1, dim> point_1, point_2;
    for (int i = 0; i < dim; ++i) { point_1[i] += data.domain_size[i] / 2; point_2[i] -= data.domain_size[i] / 2; }
    Point<dim> p1(point_1), p2(point_2);
    GridGenerator::hyper_rectangle(triangulation, p1, p2, /*colorize = */ true);
    triangulation.refine_global(data.initial_refinement_level);
  }
  void setup_dofs() {
    p_dh.distribute_dofs(p_fe);  // PS:73
    u_dh.distribute_dofs(u_fe);  // DS:110
    // DS:112-137, unchanged: hanging nodes (none here), then the Dirichlet conditions in list order
    constraints.clear();
    DoFTools::make_hanging_node_constraints(u_dh, constraints);
    std::vector<ComponentMask> mask(dim);
    for (unsigned int comp = 0; comp < dim; ++comp) mask[comp] = u_fe.component_mask(FEValuesExtractors::Scalar(comp));
    for (size_t cond = 0; cond < data.displacement_boundary_labels.size(); ++cond)
      VectorTools::interpolate_boundary_values(u_dh, data.displacement_boundary_labels[cond],
                                               ConstantFunction<dim>(data.displacement_boundary_values[cond], dim), constraints,
                                               mask[data.displacement_boundary_components[cond]]);
    constraints.close();
    gpu_backend::upload_from_dealii(ctx, triangulation, p_dh, u_dh, constraints, data, preconditioner, cg_cap);
  }
  void solve_displacement() { int its; double res; ck(pe_displacement_solve(ctx, &its, &res)); }
  void get_normal_strain_components() {  // FSS:153-164
    std::vector<int32_t> comps(volumetric.begin(), volumetric.end());
    ck(pe_project_assemble_rhs(ctx, (int)comps.size(), comps.data()));
    for (int c : volumetric) { int its; ck(pe_project_solve(ctx, tensor_indexer.entryIndex(c), &its)); }
  }
  void output_results(unsigned int n) {  // stands in for FSS:227-291: the two solution vectors, dof by dof
    Vector<double> p(p_dh.n_dofs()), u(u_dh.n_dofs());
    ck(pe_get_vector(ctx, PE_VEC_P, &p(0), p.size()));
    ck(pe_get_vector(ctx, PE_VEC_U, &u(0), u.size()));
    std::ofstream out("./solution/fields-" + Utilities::int_to_string(n, 4) + ".txt");
    out.precision(17);
    out << "p " << p.size() << "\n";
    for (unsigned int i = 0; i < p.size(); ++i) out << p(i) << "\n";
    out << "u " << u.size() << "\n";
    for (unsigned int i = 0; i < u.size(); ++i) out << u(i) << "\n";
  }

  Triangulation<dim> triangulation;
  input_data::InputDataPoroel& data;
  DoFHandler<dim> p_dh, u_dh;
  FE_Q<dim> p_fe;
  FESystem<dim> u_fe;
  ConstraintMatrix constraints;
  indexing::TensorIndexer<dim> tensor_indexer;
  std::vector<int> volumetric;
  std::vector<int32_t> volumetric_entries;
  pe_ctx* ctx = nullptr;
  int preconditioner, cg_cap;
};

int main(int argc, char** argv) {
  if (argc < 2) { std::cout << "specify the file name" << std::endl; return 1; }
  try {
    input_data::InputDataPoroel data;
    data.read_input_file(argv[1]);
    const int precond = argc > 2 ? std::atoi(argv[2]) : PE_PRECOND_CHEBYSHEV;
    const int cg_cap = argc > 3 ? std::atoi(argv[3]) : 1000;
    if (data.dim == 2) { PoroElasticProblemGpu<2> problem(data, precond, cg_cap); problem.run(); }
    else if (data.dim == 3) { PoroElasticProblemGpu<3> problem(data, precond, cg_cap); problem.run(); }
    else return 2;
  } catch (std::exception& exc) {
    std::cerr << "Exception on processing: " << exc.what() << std::endl;
    return 1;
  }
  return 0;
}
