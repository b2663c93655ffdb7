// shim_lac.h — sparsity patterns, SparseMatrix, ConstraintMatrix, SolverCG, PreconditionSSOR of the deal.II API shim (see
// shim.h: NOT deal.II).  Iteration counts of every CG solve are appended to the file named by DEALII_SHIM_SOLVER_LOG (the
// reference keeps its own prints of them commented out, PS:181-184, SP:218-221).
#pragma once

namespace dealii {

// rows are kept as unsorted lists that are sorted and made unique when they have doubled since the last clean-up (a std::set
// per row costs ten times the memory at 128^3 cells)
class DynamicSparsityPattern {
 public:
  explicit DynamicSparsityPattern(unsigned int n) : rows(n), clean(n, 0) {}
  void add(unsigned int i, unsigned int j) {
    std::vector<unsigned int>& r = rows[i];
    r.push_back(j);
    if (r.size() >= 64 && r.size() >= 2 * (size_t)clean[i]) compress(i);
  }
  unsigned int n_rows() const { return (unsigned int)rows.size(); }
  const std::vector<unsigned int>& row(unsigned int i) const { const_cast<DynamicSparsityPattern*>(this)->compress(i); return rows[i]; }  // ascending, unique
 private:
  void compress(unsigned int i) {
    std::vector<unsigned int>& r = rows[i];
    std::sort(r.begin(), r.end());
    r.erase(std::unique(r.begin(), r.end()), r.end());
    clean[i] = (unsigned int)r.size();
  }
  std::vector<std::vector<unsigned int>> rows;
  std::vector<unsigned int> clean;
};
// sparsity_pattern.cc: in a square pattern the diagonal entry is stored first in its row, the others follow in ascending order
class SparsityPattern {
 public:
  void copy_from(const DynamicSparsityPattern& d) {
    const unsigned int n = d.n_rows();
    rowstart.assign(n + 1, 0);
    for (unsigned int i = 0; i < n; ++i) {
      const std::vector<unsigned int>& r = d.row(i);
      rowstart[i + 1] = rowstart[i] + r.size() + (std::binary_search(r.begin(), r.end(), i) ? 0 : 1);
    }
    colnums.resize(rowstart[n]);
    for (unsigned int i = 0; i < n; ++i) {
      size_t k = rowstart[i];
      colnums[k++] = i;
      for (unsigned int j : d.row(i)) if (j != i) colnums[k++] = j;
    }
  }
  unsigned int n_rows() const { return (unsigned int)rowstart.size() - 1; }
  std::vector<size_t> rowstart;
  std::vector<unsigned int> colnums;
};

template <typename number> class SparseMatrix {
 public:
  void reinit(const SparsityPattern& sp) { cols = &sp; val.assign(sp.colnums.size(), number(0)); }
  unsigned int m() const { return cols->n_rows(); }
  const SparsityPattern& get_sparsity_pattern() const { return *cols; }
  size_t position(unsigned int i, unsigned int j) const {
    const size_t b = cols->rowstart[i], e = cols->rowstart[i + 1];
    if (j == i) return b;
    const unsigned int* first = &cols->colnums[b + 1];
    const unsigned int* last = &cols->colnums[0] + e;
    const unsigned int* p = std::lower_bound(first, last, j);
    AssertThrow(p != last && *p == j, ShimException("SparseMatrix: entry outside the sparsity pattern"));
    return (size_t)(p - &cols->colnums[0]);
  }
  void add(unsigned int i, unsigned int j, number v) { val[position(i, j)] += v; }
  number el(unsigned int i, unsigned int j) const { return val[position(i, j)]; }
  number diag_element(unsigned int i) const { return val[cols->rowstart[i]]; }
  void copy_from(const SparseMatrix& o) { val = o.val; }  // same pattern (PS:160, SP:103)
  SparseMatrix& operator*=(number s) { for (auto& x : val) x *= s; return *this; }
  void add(number factor, const SparseMatrix& o) { for (size_t k = 0; k < val.size(); ++k) val[k] += factor * o.val[k]; }
  // sparse_matrix.templates.h, vmult: row sums in storage order (diagonal first)
  void vmult(Vector<number>& dst, const Vector<number>& src) const {
    const unsigned int n = m();
    for (unsigned int i = 0; i < n; ++i) {
      number s = 0;
      for (size_t k = cols->rowstart[i]; k < cols->rowstart[i + 1]; ++k) s += val[k] * src(cols->colnums[k]);
      dst(i) = s;
    }
  }
  // sparse_matrix.templates.h, precondition_SSOR: forward sweep over the entries left of the diagonal, scaling by
  // om (2 - om) a_ii, backward sweep over the entries right of the diagonal
  void precondition_SSOR(Vector<number>& dst, const Vector<number>& src, number om, const std::vector<size_t>& pos_right_of_diagonal) const {
    const int n = (int)m();
    for (int row = 0; row < n; ++row) {
      number s = 0;
      for (size_t j = cols->rowstart[row] + 1; j < pos_right_of_diagonal[row]; ++j) s += val[j] * dst(cols->colnums[j]);
      dst(row) = src(row) - s * om;
      dst(row) /= val[cols->rowstart[row]];
    }
    for (int row = 0; row < n; ++row) dst(row) *= om * (number(2.) - om) * val[cols->rowstart[row]];
    for (int row = n - 1; row >= 0; --row) {
      number s = 0;
      for (size_t j = pos_right_of_diagonal[row]; j < cols->rowstart[row + 1]; ++j) s += val[j] * dst(cols->colnums[j]);
      dst(row) -= s * om;
      dst(row) /= val[cols->rowstart[row]];
    }
  }
  const SparsityPattern* cols = nullptr;
  std::vector<number> val;
};

template <typename MatrixType = SparseMatrix<double>> class PreconditionSSOR {
 public:
  void initialize(const MatrixType& A_, double omega_) {  // precondition.h: remembers where the upper triangle of every row begins
    A = &A_;
    omega = omega_;
    const SparsityPattern& sp = A_.get_sparsity_pattern();
    const unsigned int n = sp.n_rows();
    pos_right_of_diagonal.resize(n);
    for (unsigned int row = 0; row < n; ++row) {
      size_t k = sp.rowstart[row] + 1;
      while (k < sp.rowstart[row + 1] && sp.colnums[k] < row) ++k;
      pos_right_of_diagonal[row] = k;
    }
  }
  void vmult(Vector<double>& dst, const Vector<double>& src) const { A->precondition_SSOR(dst, src, omega, pos_right_of_diagonal); }
 private:
  const MatrixType* A = nullptr;
  double omega = 1;
  std::vector<size_t> pos_right_of_diagonal;
};

struct PreconditionIdentity {
  template <class V> void vmult(V& dst, const V& src) const { dst = src; }
};

// solver_control.cc, SolverControl::check: success as soon as the value is <= tol, failure when the step count is exhausted
class SolverControl {
 public:
  enum State { iterate = 0, success, failure };
  struct NoConvergence : std::runtime_error {
    NoConvergence(unsigned int step, double res)
        : std::runtime_error("SolverControl::NoConvergence after " + std::to_string(step) + " steps, residual " + std::to_string(res)), last_step(step), last_residual(res) {}
    unsigned int last_step; double last_residual;
  };
  SolverControl(unsigned int n, double tol) : maxsteps(n), tol(tol) {}
  State check(unsigned int step, double value) {
    lstep = step; lvalue = value;
    if (value <= tol) return success;
    if (step >= maxsteps || std::isnan(value)) return failure;
    return iterate;
  }
  unsigned int last_step() const { return lstep; }
  double last_value() const { return lvalue; }
 private:
  unsigned int maxsteps, lstep = 0;
  double tol, lvalue = 0;
};

// solver_cg.h (8.4), SolverCG::solve: g = A x - b, d = -P g, alpha = (g.Pg)/(d.Ad), x += alpha d, g += alpha Ad,
// beta = (g.Pg)_new / (g.Pg)_old, d = beta d - P g; the residual norm is checked after every update.
template <class VectorType = Vector<double>> class SolverCG {
 public:
  explicit SolverCG(SolverControl& c) : control(c) {}
  template <class MatrixType, class Preconditioner>
  void solve(const MatrixType& A, VectorType& x, const VectorType& b, const Preconditioner& precondition) {
    VectorType g, d, h;
    g.reinit(x); d.reinit(x); h.reinit(x);
    int it = 0;
    double res, gh, alpha, beta;
    if (!x.all_zero()) { A.vmult(g, x); g.add(-1., b); }
    else g.equ(-1., b);
    res = g.l2_norm();
    SolverControl::State conv = control.check(0, res);
    if (conv == SolverControl::iterate) {
      precondition.vmult(h, g);
      d.equ(-1., h);
      gh = g * h;
      while (conv == SolverControl::iterate) {
        ++it;
        A.vmult(h, d);
        alpha = d * h;
        alpha = gh / alpha;
        g.add(alpha, h);
        x.add(alpha, d);
        res = g.l2_norm();
        conv = control.check(it, res);
        if (conv != SolverControl::iterate) break;
        precondition.vmult(h, g);
        beta = gh;
        gh = g * h;
        beta = gh / beta;
        d.sadd(beta, -1., h);
      }
    }
    if (const char* log = std::getenv("DEALII_SHIM_SOLVER_LOG")) {
      if (FILE* f = std::fopen(log, "a")) { std::fprintf(f, "cg n=%u its=%d res=%.17g\n", x.size(), it, res); std::fclose(f); }
    }
    if (conv != SolverControl::success) throw SolverControl::NoConvergence(it, res);
  }
 private:
  SolverControl& control;
};

// constraint_matrix.h / .templates.h for what the reference uses on meshes without hanging nodes: lines x_i = g_i (no entries).
class ConstraintMatrix {
 public:
  typedef types::global_dof_index size_type;
  void clear() { line_of.clear(); closed = false; }
  void close() { closed = true; }
  bool can_store_line(size_type) const { return true; }
  bool is_constrained(size_type i) const { return line_of.count(i) != 0; }
  void add_line(size_type i) { line_of.emplace(i, 0.0); }
  void set_inhomogeneity(size_type i, double g) { line_of[i] = g; }
  unsigned int n_constraints() const { return (unsigned int)line_of.size(); }
  double inhomogeneity(size_type i) const { auto it = line_of.find(i); return it == line_of.end() ? 0.0 : it->second; }
  double get_inhomogeneity(size_type i) const { return inhomogeneity(i); }
  // (master, weight) pairs of a line; Dirichlet lines have none (the only kind on meshes without hanging nodes)
  const std::vector<std::pair<size_type, double>>* get_constraint_entries(size_type i) const { return is_constrained(i) ? &no_entries : nullptr; }
  // condense(): with no entries in any line, a constrained row/column keeps only its diagonal and the vector entry is zeroed.
  // The reference calls these on the PRESSURE constraints only, which are empty on meshes without hanging nodes.
  template <class V> void condense(V& v) const { for (const auto& l : line_of) v(l.first) = 0; }
  template <class number> void condense(SparseMatrix<number>& A) const {
    AssertThrow(line_of.empty(), ShimException("deal.II shim: ConstraintMatrix::condense(matrix) with constraints is outside the shim"));
    (void)A;
  }
  template <class V> void distribute(V& v) const { for (const auto& l : line_of) v(l.first) = l.second; }
  // distribute_local_to_global(matrix, rhs) — constraint_matrix.templates.h with use_inhomogeneities_for_rhs = false: free rows
  // get their entries in free columns, b_i -= a_ij g_j for constrained columns; a constrained row gets |a_ii| on the diagonal (the
  // mean |diagonal| of the cell matrix if that is zero) and nothing in the right-hand side.
  template <class number>
  void distribute_local_to_global(const FullMatrix<double>& local_matrix, const Vector<double>& local_rhs, const std::vector<size_type>& dofs,
                                  SparseMatrix<number>& A, Vector<double>& b) const {
    const unsigned int n = (unsigned int)dofs.size();
    bool any = false;
    for (unsigned int i = 0; i < n; ++i) any = any || is_constrained(dofs[i]);
    double average_diagonal = 0;
    if (any) {
      for (unsigned int i = 0; i < n; ++i) average_diagonal += std::fabs(local_matrix(i, i));
      average_diagonal /= n;
    }
    for (unsigned int i = 0; i < n; ++i) {
      if (is_constrained(dofs[i])) {
        const double d = std::fabs(local_matrix(i, i));
        A.add(dofs[i], dofs[i], d != 0 ? d : average_diagonal);
        continue;
      }
      double bi = local_rhs(i);
      for (unsigned int j = 0; j < n; ++j) {
        if (is_constrained(dofs[j])) bi -= local_matrix(i, j) * inhomogeneity(dofs[j]);
        else A.add(dofs[i], dofs[j], local_matrix(i, j));
      }
      b(dofs[i]) += bi;
    }
  }
  // distribute_local_to_global(local_vector, dofs, global_vector, local_matrix): the vector-only variant that still uses the cell
  // matrix to eliminate inhomogeneous columns (DS:288-290)
  void distribute_local_to_global(const Vector<double>& local_rhs, const std::vector<size_type>& dofs, Vector<double>& b,
                                  const FullMatrix<double>& local_matrix) const {
    const unsigned int n = (unsigned int)dofs.size();
    for (unsigned int i = 0; i < n; ++i) {
      if (is_constrained(dofs[i])) continue;
      double bi = local_rhs(i);
      for (unsigned int j = 0; j < n; ++j)
        if (is_constrained(dofs[j])) bi -= local_matrix(i, j) * inhomogeneity(dofs[j]);
      b(dofs[i]) += bi;
    }
  }
  void distribute_local_to_global(const Vector<double>& local_rhs, const std::vector<size_type>& dofs, Vector<double>& b) const {
    for (unsigned int i = 0; i < dofs.size(); ++i)
      if (!is_constrained(dofs[i])) b(dofs[i]) += local_rhs(i);
  }
 private:
  std::map<size_type, double> line_of;
  std::vector<std::pair<size_type, double>> no_entries;
  bool closed = false;
};

}  // namespace dealii
