// shim_numerics.h — DoFTools, MatrixCreator, VectorTools, DataOut and the refinement stubs of the deal.II API shim (see
// shim.h: NOT deal.II).
#pragma once

namespace dealii {

namespace DoFTools {
// uniform meshes have no hanging nodes
template <int dim> inline void make_hanging_node_constraints(const DoFHandler<dim>&, ConstraintMatrix&) {}
// dof_tools_sparsity.cc: every pair of dofs of a cell couples; with keep_constrained_dofs = true constrained dofs keep their entries
template <int dim> inline void make_sparsity_pattern(const DoFHandler<dim>& dh, DynamicSparsityPattern& dsp, const ConstraintMatrix&, bool keep_constrained_dofs) {
  AssertThrow(keep_constrained_dofs, ShimException("deal.II shim: make_sparsity_pattern without constrained dofs is outside the shim"));
  const unsigned int n = dh.dofs_per_cell, nc = dh.get_tria().n_active_cells();
  for (unsigned int c = 0; c < nc; ++c)
    for (unsigned int i = 0; i < n; ++i)
      for (unsigned int j = 0; j < n; ++j) dsp.add(dh.cell_dofs[(size_t)c * n + i], dh.cell_dofs[(size_t)c * n + j]);
}
template <int dim> inline void make_sparsity_pattern(const DoFHandler<dim>& dh, DynamicSparsityPattern& dsp) {
  make_sparsity_pattern(dh, dsp, ConstraintMatrix(), true);
}
}  // namespace DoFTools

namespace MatrixCreator {
// matrix_creator.cc: cell matrices int phi_i phi_j and int grad phi_i . grad phi_j, added entry by entry (scalar elements here)
template <int dim, class Kernel> inline void shim_assemble(const DoFHandler<dim>& dh, const Quadrature<dim>& q, SparseMatrix<double>& A, Kernel kernel) {
  const FiniteElement<dim>& fe = *dh.fe;
  FEValues<dim> fv(fe, q, update_values | update_gradients | update_JxW_values);
  const unsigned int n = fe.dofs_per_cell, nq = q.size();
  std::vector<types::global_dof_index> dofs(n);
  for (auto cell = dh.begin_active(); cell != dh.end(); ++cell) {
    fv.reinit(cell);
    cell->get_dof_indices(dofs);
    for (unsigned int i = 0; i < n; ++i)
      for (unsigned int j = 0; j < n; ++j) {
        double s = 0;
        for (unsigned int p = 0; p < nq; ++p) s += kernel(fv, i, j, p) * fv.JxW(p);
        A.add(dofs[i], dofs[j], s);
      }
  }
}
template <int dim> inline void create_mass_matrix(const DoFHandler<dim>& dh, const Quadrature<dim>& q, SparseMatrix<double>& A) {
  shim_assemble(dh, q, A, [](const FEValues<dim>& fv, unsigned int i, unsigned int j, unsigned int p) { return fv.shape_value(i, p) * fv.shape_value(j, p); });
}
template <int dim> inline void create_laplace_matrix(const DoFHandler<dim>& dh, const Quadrature<dim>& q, SparseMatrix<double>& A) {
  shim_assemble(dh, q, A, [](const FEValues<dim>& fv, unsigned int i, unsigned int j, unsigned int p) {
    return fv.shape_grad_component(i, p, 0) * fv.shape_grad_component(j, p, 0);
  });
}
}  // namespace MatrixCreator

namespace VectorTools {
// vector_tools.templates.h, create_right_hand_side: the vector is zeroed first, then b_i += phi_i(x_q) f(x_q) JxW
template <int dim> inline void create_right_hand_side(const DoFHandler<dim>& dh, const Quadrature<dim>& q, const Function<dim>& f, Vector<double>& b) {
  const FiniteElement<dim>& fe = *dh.fe;
  FEValues<dim> fv(fe, q, update_values | update_quadrature_points | update_JxW_values);
  const unsigned int n = fe.dofs_per_cell, nq = q.size();
  std::vector<types::global_dof_index> dofs(n);
  std::vector<double> values(nq);
  b = 0;
  for (auto cell = dh.begin_active(); cell != dh.end(); ++cell) {
    fv.reinit(cell);
    cell->get_dof_indices(dofs);
    for (unsigned int p = 0; p < nq; ++p) values[p] = f.value(fv.get_quadrature_points()[p], 0);
    for (unsigned int p = 0; p < nq; ++p)
      for (unsigned int i = 0; i < n; ++i) b(dofs[i]) += values[p] * fv.shape_value(i, p) * fv.JxW(p);
  }
}
// vector_tools.templates.h, interpolate_boundary_values(dof, boundary id, function, constraints, mask): the values of the selected
// components at the support points on faces with that id; a dof that is already constrained keeps its line (first condition wins)
template <int dim> inline void interpolate_boundary_values(const DoFHandler<dim>& dh, const types::boundary_id id, const Function<dim>& f,
                                                           ConstraintMatrix& constraints, const ComponentMask& mask = ComponentMask()) {
  const FiniteElement<dim>& fe = *dh.fe;
  std::map<types::global_dof_index, double> boundary_values;
  std::vector<types::global_dof_index> dofs(fe.dofs_per_cell);
  for (auto cell = dh.begin_active(); cell != dh.end(); ++cell)
    for (unsigned int face = 0; face < GeometryInfo<dim>::faces_per_cell; ++face) {
      if (!cell->face(face)->at_boundary() || cell->face(face)->boundary_id() != id) continue;
      cell->get_dof_indices(dofs);
      const unsigned int axis = face / 2, side = face % 2;
      for (unsigned int i = 0; i < fe.dofs_per_cell; ++i) {
        const auto& node = fe.nodes[i / fe.n_comp];
        const unsigned int comp = i % fe.n_comp;
        if (node.xi[axis] != (double)side || !mask[comp]) continue;
        Point<dim> x;  // support point through the Q1 map
        for (unsigned int v = 0; v < GeometryInfo<dim>::vertices_per_cell; ++v) {
          double N = 1;
          for (int a = 0; a < dim; ++a) N *= ((v >> a) & 1) ? node.xi[a] : 1 - node.xi[a];
          const Point<dim> X = cell->vertex(v);
          for (int a = 0; a < dim; ++a) x[a] += N * X[a];
        }
        boundary_values[dofs[i]] = f.value(x, comp);
      }
    }
  for (const auto& bv : boundary_values)
    if (constraints.can_store_line(bv.first) && !constraints.is_constrained(bv.first)) {
      constraints.add_line(bv.first);
      constraints.set_inhomogeneity(bv.first, bv.second);
    }
}
// the std::map variant (vector_tools.templates.h): later cells overwrite earlier ones with the same value
template <int dim> inline void interpolate_boundary_values(const DoFHandler<dim>& dh, const types::boundary_id id, const Function<dim>& f,
                                                           std::map<types::global_dof_index, double>& boundary_values) {
  ConstraintMatrix tmp;
  interpolate_boundary_values(dh, id, f, tmp);
  for (types::global_dof_index i = 0; i < dh.n_dofs(); ++i)
    if (tmp.is_constrained(i)) boundary_values[i] = tmp.get_inhomogeneity(i);
}
}  // namespace VectorTools

namespace MatrixTools {
// matrix_tools.cc, apply_boundary_values(boundary_values, matrix, solution, rhs, eliminate_columns = true): the row of a boundary
// dof keeps only its diagonal, rhs_i = a_ii g_i, solution_i = g_i; its column is eliminated into the right-hand side of the
// other rows (symmetric pattern assumed, as deal.II does)
inline void apply_boundary_values(const std::map<types::global_dof_index, double>& boundary_values, SparseMatrix<double>& A, Vector<double>& solution,
                                  Vector<double>& rhs, const bool eliminate_columns = true) {
  const SparsityPattern& sp = A.get_sparsity_pattern();
  double first_nonzero_diagonal_entry = 1;
  for (unsigned int i = 0; i < A.m(); ++i)
    if (A.diag_element(i) != 0) { first_nonzero_diagonal_entry = A.diag_element(i); break; }
  for (const auto& bv : boundary_values) {
    const unsigned int dof = bv.first;
    for (size_t k = sp.rowstart[dof] + 1; k < sp.rowstart[dof + 1]; ++k) A.val[k] = 0;
    double new_rhs;
    if (A.diag_element(dof) != 0) new_rhs = bv.second * A.diag_element(dof);
    else { A.val[sp.rowstart[dof]] = first_nonzero_diagonal_entry; new_rhs = bv.second * first_nonzero_diagonal_entry; }
    rhs(dof) = new_rhs;
    if (eliminate_columns) {
      const double diagonal_entry = A.diag_element(dof);
      for (size_t k = sp.rowstart[dof] + 1; k < sp.rowstart[dof + 1]; ++k) {  // rows that couple with `dof` (symmetric pattern)
        const unsigned int row = sp.colnums[k];
        const size_t pos = A.position(row, dof);
        rhs(row) -= A.val[pos] / diagonal_entry * new_rhs;
        A.val[pos] = 0;
      }
    }
    solution(dof) = bv.second;
  }
}
}  // namespace MatrixTools

// ---------------------------------------------------------------------------------------------- output
namespace DataComponentInterpretation {
enum DataComponentInterpretation { component_is_scalar, component_is_part_of_vector };
}
// Not a VTK writer: write_vtk() dumps every attached vector dof by dof with the support point of the dof, at full precision —
//     field <name> components <n> dofs <n_dofs>
//     <x> <y> [<z>] <component> <value>          (one line per dof, in dof order)
// which is what tests/golden/make_reference_run.py reads back.
template <int dim> class DataOut {
 public:
  template <class V>
  void add_data_vector(const DoFHandler<dim>& dh, const V& v, const std::vector<std::string>& names,
                       const std::vector<DataComponentInterpretation::DataComponentInterpretation>&) {
    fields.push_back(Field{&dh, std::vector<double>(v.v.begin(), v.v.end()), names[0]});
  }
  template <class V> void add_data_vector(const DoFHandler<dim>& dh, const V& v, const std::string& name) {
    fields.push_back(Field{&dh, std::vector<double>(v.v.begin(), v.v.end()), name});
  }
  void build_patches(unsigned int = 0) {}
  void write_vtk(std::ostream& out) const {
    out << "# deal.II API shim dump (not VTK)\n";
    out.precision(17);
    for (const Field& f : fields) {
      const DoFHandler<dim>& dh = *f.dh;
      const FiniteElement<dim>& fe = *dh.fe;
      out << "field " << f.name << " components " << fe.n_comp << " dofs " << dh.n_dofs() << "\n";
      std::vector<std::array<double, 3>> x(dh.n_dofs());
      std::vector<int> comp(dh.n_dofs(), -1);
      const unsigned int n = fe.dofs_per_cell, nc = dh.get_tria().n_active_cells();
      for (unsigned int c = 0; c < nc; ++c)
        for (unsigned int i = 0; i < n; ++i) {
          const types::global_dof_index g = dh.cell_dofs[(size_t)c * n + i];
          if (comp[g] >= 0) continue;
          comp[g] = (int)(i % fe.n_comp);
          const auto& node = fe.nodes[i / fe.n_comp];
          x[g] = {{0, 0, 0}};
          for (unsigned int v = 0; v < GeometryInfo<dim>::vertices_per_cell; ++v) {
            double N = 1;
            for (int a = 0; a < dim; ++a) N *= ((v >> a) & 1) ? node.xi[a] : 1 - node.xi[a];
            const Point<dim> X = dh.get_tria().vertex_of(c, v);
            for (int a = 0; a < dim; ++a) x[g][a] += N * X[a];
          }
        }
      for (unsigned int g = 0; g < dh.n_dofs(); ++g) {
        for (int a = 0; a < dim; ++a) out << x[g][a] << " ";
        out << comp[g] << " " << f.values[g] << "\n";
      }
    }
  }
 private:
  struct Field { const DoFHandler<dim>* dh; std::vector<double> values; std::string name; };
  std::vector<Field> fields;
};

// ---------------------------------------------------------------------------------------------- adaptive refinement: declared, not provided
template <int dim> class KellyErrorEstimator {
 public:
  template <class Q, class Map, class V, class E> static void estimate(const DoFHandler<dim>&, const Q&, const Map&, const V&, E&) {
    shim_unsupported("KellyErrorEstimator::estimate");
  }
};
template <int dim> class SolutionTransfer {
 public:
  explicit SolutionTransfer(const DoFHandler<dim>&) {}
  template <class V> void prepare_for_coarsening_and_refinement(const std::vector<V>&) { shim_unsupported("SolutionTransfer"); }
  template <class V> void interpolate(const std::vector<V>&, std::vector<V>&) const { shim_unsupported("SolutionTransfer"); }
};

}  // namespace dealii
