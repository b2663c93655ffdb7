// poisson_on_shim.cpp — an anchor for the deal.II API shim (oracle/dealii_shim) that does not come from this repository: the
// problem of deal.II's tutorial step-4, written against the deal.II API the way the tutorial structures it (make the grid,
// distribute dofs, assemble -Laplace u = f cell by cell with FEValues, interpolate_boundary_values + apply_boundary_values,
// SolverCG without preconditioner, SolverControl(1000, 1e-12)).  The tutorial's "Results" section publishes what deal.II prints
// for it: 2D — 256 active cells, 289 degrees of freedom, 26 CG iterations; 3D — 4096 cells, 4913 degrees of freedom, 30 CG
// iterations.  tests/test_reference_run.py compiles this file against the shim and expects exactly these numbers.
#include <deal.II/base/function.h>
#include <deal.II/dofs/dof_handler.h>
#include <deal.II/dofs/dof_tools.h>
#include <deal.II/fe/fe_q.h>
#include <deal.II/fe/fe_values.h>
#include <deal.II/grid/grid_generator.h>
#include <deal.II/grid/tria.h>
#include <deal.II/lac/dynamic_sparsity_pattern.h>
#include <deal.II/lac/precondition.h>
#include <deal.II/lac/solver_cg.h>
#include <deal.II/lac/sparse_matrix.h>
#include <deal.II/lac/vector.h>
#include <deal.II/numerics/matrix_tools.h>
#include <deal.II/numerics/vector_tools.h>

using namespace dealii;

template <int dim> struct Load : Function<dim> {  // f = 4 sum_a x_a^4
  double value(const Point<dim>& p, const unsigned int = 0) const override {
    double s = 0;
    for (int a = 0; a < dim; ++a) s += 4.0 * std::pow(p[a], 4.0);
    return s;
  }
};
template <int dim> struct BoundaryData : Function<dim> {  // u = |x|^2 on the boundary
  double value(const Point<dim>& p, const unsigned int = 0) const override { return p * p; }
};

template <int dim> void poisson() {
  Triangulation<dim> triangulation;
  GridGenerator::hyper_cube(triangulation, -1, 1);
  triangulation.refine_global(4);
  FE_Q<dim> fe(1);
  DoFHandler<dim> dof_handler(triangulation);
  dof_handler.distribute_dofs(fe);
  std::cout << "   Number of active cells: " << triangulation.n_active_cells() << std::endl;
  std::cout << "   Number of degrees of freedom: " << dof_handler.n_dofs() << std::endl;
  DynamicSparsityPattern dsp(dof_handler.n_dofs());
  DoFTools::make_sparsity_pattern(dof_handler, dsp);
  SparsityPattern sparsity_pattern;
  sparsity_pattern.copy_from(dsp);
  SparseMatrix<double> system_matrix;
  system_matrix.reinit(sparsity_pattern);
  Vector<double> solution(dof_handler.n_dofs()), system_rhs(dof_handler.n_dofs());

  QGauss<dim> quadrature(2);
  FEValues<dim> fe_values(fe, quadrature, update_values | update_gradients | update_quadrature_points | update_JxW_values);
  const unsigned int dofs_per_cell = fe.dofs_per_cell, n_q = quadrature.size();
  FullMatrix<double> cell_matrix(dofs_per_cell, dofs_per_cell);
  Vector<double> cell_rhs(dofs_per_cell);
  std::vector<types::global_dof_index> local(dofs_per_cell);
  const Load<dim> load;
  for (auto cell = dof_handler.begin_active(); cell != dof_handler.end(); ++cell) {
    fe_values.reinit(cell);
    cell_matrix = 0;
    cell_rhs = 0;
    for (unsigned int q = 0; q < n_q; ++q)
      for (unsigned int i = 0; i < dofs_per_cell; ++i) {
        for (unsigned int j = 0; j < dofs_per_cell; ++j) cell_matrix(i, j) += fe_values.shape_grad(i, q) * fe_values.shape_grad(j, q) * fe_values.JxW(q);
        cell_rhs(i) += fe_values.shape_value(i, q) * load.value(fe_values.quadrature_point(q)) * fe_values.JxW(q);
      }
    cell->get_dof_indices(local);
    for (unsigned int i = 0; i < dofs_per_cell; ++i) {
      for (unsigned int j = 0; j < dofs_per_cell; ++j) system_matrix.add(local[i], local[j], cell_matrix(i, j));
      system_rhs(local[i]) += cell_rhs(i);
    }
  }
  std::map<types::global_dof_index, double> boundary_values;
  VectorTools::interpolate_boundary_values(dof_handler, 0, BoundaryData<dim>(), boundary_values);
  MatrixTools::apply_boundary_values(boundary_values, system_matrix, solution, system_rhs);

  SolverControl solver_control(1000, 1e-12);
  SolverCG<> solver(solver_control);
  solver.solve(system_matrix, solution, system_rhs, PreconditionIdentity());
  std::cout << "   " << solver_control.last_step() << " CG iterations needed to obtain convergence." << std::endl;
}

int main() {
  std::cout << "Solving problem in 2 space dimensions." << std::endl;
  poisson<2>();
  std::cout << "Solving problem in 3 space dimensions." << std::endl;
  poisson<3>();
  return 0;
}
