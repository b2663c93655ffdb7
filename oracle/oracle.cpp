// oracle.cpp — CPU restatement of the reference's fixed-stress hot path.  TEST INFRASTRUCTURE ONLY.
//
// PARITY PINNED ON THE REFERENCE'S OWN CODE, NOT ON deal.II: the reference (ishovkun/poroelasticity-dealii) ships no tests,
// golden vectors or example output and has no main(), and the library it is written against, deal.II >= 8.4
// (CMakeLists.txt:3), is neither vendored nor installed here.  Since round 2 the reference's source files are nevertheless
// executed in the build container: oracle/_ref/fss_ref is /root/reference/lib/include/*.h, unmodified, compiled against a
// deal.II API shim (oracle/dealii_shim — a second, independent restatement of the library calls the reference makes, NOT
// deal.II) with oracle/ref_main.cpp as the missing Runner.cpp.  Its PoroElasticProblem<dim>::run() produced the golden vectors
// tests/golden/reference_run_*; this file reproduces them with the same dof numbering, the same number of CG iterations in
// every solve, the printed values to all digits and the fields to 1e-13 (tests/test_reference_run.py).  What therefore remains
// a restatement — here and in the shim — is deal.II itself (FE tables, quadrature, sparsity order, SolverCG, SSOR,
// ConstraintMatrix); that reading is pinned by the known-answer tests in tests/ (element matrices vs closed forms and sympy,
// patch test, manufactured-solution rates, an independent numpy/scipy restatement in oracle/oracle_np.py) and by one external
// anchor: the CG iteration counts deal.II's tutorial step-4 publishes (26 in 2D / 30 in 3D), which the Laplace assembly +
// cg_solve below reproduce (tests/test_oracle.py T10).  Hanging-node meshes are outside the shim, hence pinned by the
// known-answer tests only.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
// load this library.  The product (libporoel.so) never links or calls it.
//
// What follows which reference lines (paths relative to /root/reference/lib/include):
//   FE tables / FEValues / QGauss / MappingQ1 ........ deal.II semantics used at PS:96-101, DS:159-173, SP:126-134
//   make_pattern (diag first, then ascending) ........ PS:80-88, DS:140-146 (SparsityPattern::copy_from)
//   assemble_mass_laplace ............................ PS:96-101 (MatrixCreator::create_mass/laplace_matrix)
//   well_rhs ......................................... PS:142-147 + right_hand_side.h:99-116 (pi = 3.1415926)
//   displacement_assemble ............................ DS:155-291 + ConstitutiveModel.h:9-57
//                                                      + ConstraintMatrix::distribute_local_to_global
//   cg_solve ......................................... SolverCG<>::solve (deal.II 8.4) PS:176-179, DS:300-305, SP:210-214
//   ssor_apply ....................................... SparseMatrix::precondition_SSOR PS:177-178, DS:302-303, SP:211-212
//   pressure_residual / jacobian / solve ............. PS:113-155, PS:158-169, PS:172-185, PS:187-194
//   projection rhs / solve ........................... SP:109-198, SP:201-232
//   hanging-node constraints (adaptive meshes) ....... ConstraintMatrix::condense (vector and SparseMatrix, in place),
//                                                      distribute, distribute_local_to_global with weights,
//                                                      DoFTools::make_sparsity_pattern(dh, dsp, constraints, true)
//                                                      as called at PS:71-88, PS:153, PS:168, PS:180, DS:109-146, DS:279-286,
//                                                      DS:306, SP:104-105, SP:193-194, SP:215
// Deliberate, documented deviations that do not change results beyond round-off:
//   * strain tensors of shape functions are computed once per (i,q), not inside the j loop (DS:238-239);
//   * in rhs-only calls (DS:285-286) the cell matrix is evaluated only on cells that own a
//     constrained dof with non-zero inhomogeneity (others cannot contribute);
//   * the well source vector is cached (f is time independent, PS:142-143).
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "../include/poroel.h"  // pe_params / pe_stats layouts and the status/vector/matrix enums

namespace {

using std::vector;
typedef vector<double> Vec;

// ------------------------------------------------------------------ FE tables
struct Quad { int n = 0; vector<double> pts, w; };  // pts: n*dim on [0,1]^dim, x fastest

Quad qgauss(int dim, int n1d) {
  vector<double> x, w;
  if (n1d == 2) {
    double a = 0.5 / std::sqrt(3.0);
    x = {0.5 - a, 0.5 + a};
    w = {0.5, 0.5};
  } else if (n1d == 3) {
    double a = 0.5 * std::sqrt(3.0 / 5.0);
    x = {0.5 - a, 0.5, 0.5 + a};
    w = {5.0 / 18.0, 8.0 / 18.0, 5.0 / 18.0};
  } else {
    x = {0.5};
    w = {1.0};
  }
  Quad q;
  int n = 1;
  for (int a = 0; a < dim; ++a) n *= n1d;
  q.n = n;
  q.pts.resize((size_t)n * (dim > 0 ? dim : 1));
  q.w.resize(n);
  for (int k = 0; k < n; ++k) {
    int r = k;
    double ww = 1;
    for (int a = 0; a < dim; ++a) {
      int i = r % n1d;
      r /= n1d;
      q.pts[(size_t)k * dim + a] = x[i];
      ww *= w[i];
    }
    q.w[k] = ww;
  }
  return q;
}

// unit support points of FE_Q(degree) in deal.II local order (vertices, lines, quads, hex)
vector<double> unit_support(int dim, int degree) {
  vector<double> s;
  int nv = 1 << dim;
  auto vc = [](int v, int a) { return (double)((v >> a) & 1); };
  for (int v = 0; v < nv; ++v)
    for (int a = 0; a < dim; ++a) s.push_back(vc(v, a));
  if (degree == 2) {
    static const int l2[4][2] = {{0, 2}, {1, 3}, {0, 1}, {2, 3}};
    static const int l3[12][2] = {{0, 2}, {1, 3}, {0, 1}, {2, 3}, {4, 6}, {5, 7}, {4, 5}, {6, 7}, {0, 4}, {1, 5}, {2, 6}, {3, 7}};
    static const int q3[6][4] = {{0, 2, 4, 6}, {1, 3, 5, 7}, {0, 1, 4, 5}, {2, 3, 6, 7}, {0, 1, 2, 3}, {4, 5, 6, 7}};
    int nl = dim == 2 ? 4 : 12;
    for (int l = 0; l < nl; ++l)
      for (int a = 0; a < dim; ++a) {
        const int* p = dim == 2 ? l2[l] : l3[l];
        s.push_back(0.5 * (vc(p[0], a) + vc(p[1], a)));
      }
    if (dim == 3)
      for (int f = 0; f < 6; ++f)
        for (int a = 0; a < dim; ++a) s.push_back(0.25 * (vc(q3[f][0], a) + vc(q3[f][1], a) + vc(q3[f][2], a) + vc(q3[f][3], a)));
    for (int a = 0; a < dim; ++a) s.push_back(0.5);
  }
  return s;
}

void lagrange1d(int degree, int node, double x, double& v, double& d) {
  if (degree == 1) {
    if (node == 0) { v = 1 - x; d = -1; } else { v = x; d = 1; }
  } else {  // nodes 0, 0.5, 1 -> node index 0,1,2
    if (node == 0) { v = 2 * x * x - 3 * x + 1; d = 4 * x - 3; }
    else if (node == 1) { v = 4 * x * (1 - x); d = 4 - 8 * x; }
    else { v = 2 * x * x - x; d = 4 * x - 1; }
  }
}

// values/gradients of the FE_Q(degree) scalar basis at given unit points
struct Shape {
  int ns = 0, nq = 0, dim = 0;
  vector<double> N, dN;  // N[q*ns+s], dN[(q*ns+s)*dim+a]
};
Shape make_shape(int dim, int degree, const vector<double>& pts, int npts) {
  vector<double> sup = unit_support(dim, degree);
  Shape sh;
  sh.dim = dim;
  sh.ns = (int)sup.size() / dim;
  sh.nq = npts;
  sh.N.assign((size_t)npts * sh.ns, 0);
  sh.dN.assign((size_t)npts * sh.ns * dim, 0);
  for (int q = 0; q < npts; ++q)
    for (int s = 0; s < sh.ns; ++s) {
      double v[3], d[3];
      for (int a = 0; a < dim; ++a) {
        double u = sup[(size_t)s * dim + a];
        int node = degree == 1 ? (u > 0.5 ? 1 : 0) : (u < 0.25 ? 0 : (u < 0.75 ? 1 : 2));
        lagrange1d(degree, node, pts[(size_t)q * dim + a], v[a], d[a]);
      }
      double val = 1;
      for (int a = 0; a < dim; ++a) val *= v[a];
      sh.N[(size_t)q * sh.ns + s] = val;
      for (int a = 0; a < dim; ++a) {
        double g = d[a];
        for (int b = 0; b < dim; ++b)
          if (b != a) g *= v[b];
        sh.dN[((size_t)q * sh.ns + s) * dim + a] = g;
      }
    }
  return sh;
}

// ------------------------------------------------------------------ CSR (deal.II layout: diagonal first)
struct Pattern {
  int64_t n = 0;
  vector<int64_t> rowptr;
  vector<int32_t> col;
  vector<int64_t> right_of_diag;  // PreconditionSSOR::initialize
  int64_t nnz() const { return (int64_t)col.size(); }
  int64_t find(int32_t r, int32_t c) const {
    if (r == c) return rowptr[r];
    const int32_t* b = &col[rowptr[r] + 1];
    const int32_t* e = &col[0] + rowptr[r + 1];
    const int32_t* p = std::lower_bound(b, e, c);
    return (p != e && *p == c) ? (int64_t)(p - &col[0]) : -1;
  }
};

int g_threads = 1;  // po_set_threads: 1 = faithful serial loops; >1 = OpenMP where the algorithm allows it

Pattern make_pattern(int64_t n_dofs, int64_t n_cells, int n_loc, const int32_t* cell_dofs) {
  // dof -> cells adjacency
  vector<int64_t> cnt(n_dofs + 1, 0);
  for (int64_t i = 0; i < n_cells * n_loc; ++i) cnt[cell_dofs[i] + 1]++;
  for (int64_t i = 0; i < n_dofs; ++i) cnt[i + 1] += cnt[i];
  vector<int32_t> adj(cnt[n_dofs]);
  {
    vector<int64_t> pos(cnt.begin(), cnt.end() - 1);
    for (int64_t c = 0; c < n_cells; ++c)
      for (int k = 0; k < n_loc; ++k) adj[pos[cell_dofs[c * n_loc + k]]++] = (int32_t)c;
  }
  Pattern P;
  P.n = n_dofs;
  P.rowptr.assign(n_dofs + 1, 0);
  // pass 1: count (rows are independent: OpenMP when po_set_threads(>1) was called)
#pragma omp parallel for num_threads(g_threads) schedule(static, 4096) if (g_threads > 1)
  for (int64_t r = 0; r < n_dofs; ++r) {
    vector<int32_t> tmp;
    for (int64_t a = cnt[r]; a < cnt[r + 1]; ++a)
      for (int k = 0; k < n_loc; ++k) tmp.push_back(cell_dofs[(int64_t)adj[a] * n_loc + k]);
    std::sort(tmp.begin(), tmp.end());
    P.rowptr[r + 1] = (int64_t)(std::unique(tmp.begin(), tmp.end()) - tmp.begin());
  }
  for (int64_t r = 0; r < n_dofs; ++r) P.rowptr[r + 1] += P.rowptr[r];
  P.col.resize(P.rowptr[n_dofs]);
  P.right_of_diag.resize(n_dofs);
#pragma omp parallel for num_threads(g_threads) schedule(static, 4096) if (g_threads > 1)
  for (int64_t r = 0; r < n_dofs; ++r) {
    vector<int32_t> tmp;
    for (int64_t a = cnt[r]; a < cnt[r + 1]; ++a)
      for (int k = 0; k < n_loc; ++k) tmp.push_back(cell_dofs[(int64_t)adj[a] * n_loc + k]);
    std::sort(tmp.begin(), tmp.end());
    tmp.erase(std::unique(tmp.begin(), tmp.end()), tmp.end());
    int64_t p = P.rowptr[r];
    P.col[p++] = (int32_t)r;
    P.right_of_diag[r] = P.rowptr[r + 1];
    bool found = false;
    for (int32_t c : tmp) {
      if (c == r) continue;
      if (!found && c > r) { P.right_of_diag[r] = p; found = true; }
      P.col[p++] = c;
    }
  }
  return P;
}

// ConstraintMatrix after close(): x_i = sum_j w_ij x_j + g_i; masters are never constrained themselves
struct Lines {
  vector<int32_t> line_of;  // dof -> line or -1
  vector<int32_t> dof;
  vector<int64_t> eptr;
  vector<int32_t> edof;
  vector<double> ew, g;
  int64_t n() const { return (int64_t)dof.size(); }
  bool any_entries() const { return !edof.empty(); }
  void set(int64_t n_dofs, int64_t nl, const int32_t* ld, const int64_t* ep, const int32_t* ed, const double* w, const double* inh) {
    dof.assign(ld, ld + nl);
    g.assign(nl, 0.0);
    if (inh) g.assign(inh, inh + nl);
    eptr.assign(nl + 1, 0);
    edof.clear();
    ew.clear();
    if (ep && ep[nl] > 0) {
      eptr.assign(ep, ep + nl + 1);
      edof.assign(ed, ed + ep[nl]);
      ew.assign(w, w + ep[nl]);
    }
    line_of.assign(n_dofs, -1);
    for (int64_t i = 0; i < nl; ++i) line_of[ld[i]] = (int32_t)i;
  }
  // ConstraintMatrix::condense(Vector&): constrained entries are added to their masters, then zeroed
  void condense(Vec& v) const {
    for (int64_t l = 0; l < n(); ++l) {
      for (int64_t e = eptr[l]; e < eptr[l + 1]; ++e) v[edof[e]] += v[dof[l]] * ew[e];
      v[dof[l]] = 0;
    }
  }
  // ConstraintMatrix::distribute
  void distribute(Vec& v) const {
    for (int64_t l = 0; l < n(); ++l) {
      double s = g[l];
      for (int64_t e = eptr[l]; e < eptr[l + 1]; ++e) s += v[edof[e]] * ew[e];
      v[dof[l]] = s;
    }
  }
};

// DoFTools::make_sparsity_pattern(dof_handler, dsp, constraints, keep_constrained_dofs = true): every pair of local dofs
// of a cell, plus every pair of the cell's resolved dofs (unconstrained local dofs and the masters of constrained ones).
Pattern make_pattern_constrained(int64_t n_dofs, int64_t n_cells, int n_loc, const int32_t* cell_dofs, const Lines& L) {
  vector<vector<int32_t>> rows(n_dofs);
  vector<int32_t> res;
  for (int64_t c = 0; c < n_cells; ++c) {
    const int32_t* cd = cell_dofs + c * n_loc;
    res.clear();
    for (int k = 0; k < n_loc; ++k) {
      const int32_t l = L.line_of[cd[k]];
      if (l < 0) res.push_back(cd[k]);
      else for (int64_t e = L.eptr[l]; e < L.eptr[l + 1]; ++e) res.push_back(L.edof[e]);
    }
    std::sort(res.begin(), res.end());
    res.erase(std::unique(res.begin(), res.end()), res.end());
    for (int i = 0; i < n_loc; ++i) rows[cd[i]].insert(rows[cd[i]].end(), cd, cd + n_loc);
    bool differs = (int)res.size() != n_loc;
    if (!differs) {
      vector<int32_t> loc(cd, cd + n_loc);
      std::sort(loc.begin(), loc.end());
      differs = loc != res;
    }
    if (differs)
      for (int32_t r : res) rows[r].insert(rows[r].end(), res.begin(), res.end());
  }
  Pattern P;
  P.n = n_dofs;
  P.rowptr.assign(n_dofs + 1, 0);
  for (int64_t r = 0; r < n_dofs; ++r) {
    auto& t = rows[r];
    t.push_back((int32_t)r);
    std::sort(t.begin(), t.end());
    t.erase(std::unique(t.begin(), t.end()), t.end());
    P.rowptr[r + 1] = P.rowptr[r] + (int64_t)t.size();
  }
  P.col.resize(P.rowptr[n_dofs]);
  P.right_of_diag.resize(n_dofs);
  for (int64_t r = 0; r < n_dofs; ++r) {
    int64_t p = P.rowptr[r];
    P.col[p++] = (int32_t)r;
    P.right_of_diag[r] = P.rowptr[r + 1];
    bool found = false;
    for (int32_t c : rows[r]) {
      if (c == r) continue;
      if (!found && c > r) { P.right_of_diag[r] = p; found = true; }
      P.col[p++] = c;
    }
  }
  return P;
}

// ConstraintMatrix::condense(SparseMatrix&), in place (deal.II 8.4): regular rows hand their entries in constrained
// columns to the master columns; constrained rows hand everything to the master rows; a constrained diagonal is set to
// the average absolute diagonal of the uncondensed matrix.
void condense_matrix(const Pattern& P, const Lines& L, Vec& A) {
  if (L.n() == 0) return;
  double average_diagonal = 0;
  for (int64_t i = 0; i < P.n; ++i) average_diagonal += std::fabs(A[P.rowptr[i]]);
  average_diagonal /= (double)P.n;
  auto add = [&](int32_t r, int32_t c, double v) {
    const int64_t pos = P.find(r, c);
    if (pos < 0) throw std::runtime_error("condense: entry missing from the sparsity pattern");
    A[pos] += v;
  };
  for (int64_t row = 0; row < P.n; ++row) {
    const int32_t lr = L.line_of[row];
    for (int64_t j = P.rowptr[row]; j < P.rowptr[row + 1]; ++j) {
      const int32_t column = P.col[j];
      const int32_t lc = L.line_of[column];
      const double v = A[j];
      if (lr < 0) {
        if (lc < 0) continue;
        for (int64_t q = L.eptr[lc]; q < L.eptr[lc + 1]; ++q) add((int32_t)row, L.edof[q], v * L.ew[q]);
        A[j] = 0;
      } else if (lc < 0) {
        for (int64_t q = L.eptr[lr]; q < L.eptr[lr + 1]; ++q) add(L.edof[q], column, v * L.ew[q]);
        A[j] = 0;
      } else {
        for (int64_t p = L.eptr[lr]; p < L.eptr[lr + 1]; ++p)
          for (int64_t q = L.eptr[lc]; q < L.eptr[lc + 1]; ++q) add(L.edof[p], L.edof[q], v * L.ew[p] * L.ew[q]);
        A[j] = (row == column) ? average_diagonal : 0.0;
      }
    }
  }
}

// SparseMatrix::vmult
void vmult(const Pattern& P, const Vec& val, const Vec& x, Vec& y) {
  const int64_t n = P.n;
#pragma omp parallel for num_threads(g_threads) schedule(static) if (g_threads > 1)
  for (int64_t r = 0; r < n; ++r) {
    double s = 0;
    for (int64_t j = P.rowptr[r]; j < P.rowptr[r + 1]; ++j) s += val[j] * x[P.col[j]];
    y[r] = s;
  }
}

double dot(const Vec& a, const Vec& b) {
  double s = 0;
  const int64_t n = (int64_t)a.size();
#pragma omp parallel for num_threads(g_threads) reduction(+ : s) schedule(static) if (g_threads > 1)
  for (int64_t i = 0; i < n; ++i) s += a[i] * b[i];
  return s;
}
double l2norm(const Vec& a) { return std::sqrt(dot(a, a)); }

// SparseMatrix::precondition_SSOR (natural row order; inherently sequential)
void ssor_apply(const Pattern& P, const Vec& val, double om, const Vec& src, Vec& dst) {
  const int64_t n = P.n;
  for (int64_t r = 0; r < n; ++r) {
    double s = 0;
    for (int64_t j = P.rowptr[r] + 1; j < P.right_of_diag[r]; ++j) s += val[j] * dst[P.col[j]];
    dst[r] = (src[r] - s * om) / val[P.rowptr[r]];
  }
  for (int64_t r = 0; r < n; ++r) dst[r] *= om * (2. - om) * val[P.rowptr[r]];
  for (int64_t r = n - 1; r >= 0; --r) {
    double s = 0;
    for (int64_t j = P.right_of_diag[r]; j < P.rowptr[r + 1]; ++j) s += val[j] * dst[P.col[j]];
    dst[r] = (dst[r] - s * om) / val[P.rowptr[r]];
  }
}

void jacobi_apply(const Pattern& P, const Vec& val, const Vec& src, Vec& dst) {
  for (int64_t r = 0; r < P.n; ++r) dst[r] = src[r] / val[P.rowptr[r]];
}

struct CgResult { int its = 0; double res = 0; bool ok = true; };

// SolverCG<>::solve (deal.II 8.4) with SolverControl(max_steps, tol)
// precond: omega > 0 -> SSOR(omega); omega == 0 -> Jacobi (debug aid, not the reference); omega < 0 -> PreconditionIdentity
CgResult cg_solve(const Pattern& P, const Vec& A, Vec& x, const Vec& b, double omega, int max_steps, double tol) {
  const int64_t n = P.n;
  Vec g(n), h(n), d(n);
  CgResult R;
  bool all_zero = true;
  for (int64_t i = 0; i < n; ++i)
    if (x[i] != 0) { all_zero = false; break; }
  if (!all_zero) {
    vmult(P, A, x, g);
    for (int64_t i = 0; i < n; ++i) g[i] -= b[i];
  } else
    for (int64_t i = 0; i < n; ++i) g[i] = -b[i];
  double res = l2norm(g);
  R.res = res;
  auto check = [&](int step, double v) {  // SolverControl::check
    if (v <= tol) return 1;
    if (step >= max_steps || std::isnan(v)) return -1;
    return 0;
  };
  int conv = check(0, res);
  if (conv != 0) { R.ok = conv > 0; return R; }
  auto precond = [&](const Vec& s, Vec& t) {
    if (omega > 0) ssor_apply(P, A, omega, s, t);
    else if (omega == 0) jacobi_apply(P, A, s, t);
    else t = s;
  };
  precond(g, h);
  for (int64_t i = 0; i < n; ++i) d[i] = -h[i];
  double gh = dot(g, h);
  int it = 0;
  while (conv == 0) {
    it++;
    vmult(P, A, d, h);
    double alpha = dot(d, h);
    alpha = gh / alpha;
#pragma omp parallel for num_threads(g_threads) schedule(static) if (g_threads > 1)
    for (int64_t i = 0; i < n; ++i) { g[i] += alpha * h[i]; x[i] += alpha * d[i]; }
    res = l2norm(g);
    conv = check(it, res);
    if (conv != 0) break;
    precond(g, h);
    double beta = gh;
    gh = dot(g, h);
    beta = gh / beta;
#pragma omp parallel for num_threads(g_threads) schedule(static) if (g_threads > 1)
    for (int64_t i = 0; i < n; ++i) d[i] = beta * d[i] - h[i];
  }
  R.its = it;
  R.res = res;
  R.ok = conv > 0;
  return R;
}

// ------------------------------------------------------------------ geometry (MappingQ1)
struct CellGeom {
  // per q: JxW, inverse-transpose Jacobian (row-major dim*dim), physical point
  vector<double> JxW, JinvT, xq;
};

inline double det_inv(int dim, const double* J, double* inv) {  // inv = J^-1
  if (dim == 2) {
    double det = J[0] * J[3] - J[1] * J[2];
    inv[0] = J[3] / det; inv[1] = -J[1] / det; inv[2] = -J[2] / det; inv[3] = J[0] / det;
    return det;
  }
  double c00 = J[4] * J[8] - J[5] * J[7], c01 = J[5] * J[6] - J[3] * J[8], c02 = J[3] * J[7] - J[4] * J[6];
  double det = J[0] * c00 + J[1] * c01 + J[2] * c02;
  inv[0] = c00 / det; inv[1] = (J[2] * J[7] - J[1] * J[8]) / det; inv[2] = (J[1] * J[5] - J[2] * J[4]) / det;
  inv[3] = c01 / det; inv[4] = (J[0] * J[8] - J[2] * J[6]) / det; inv[5] = (J[2] * J[3] - J[0] * J[5]) / det;
  inv[6] = c02 / det; inv[7] = (J[1] * J[6] - J[0] * J[7]) / det; inv[8] = (J[0] * J[4] - J[1] * J[3]) / det;
  return det;
}

struct Ctx {
  std::string err;
  pe_params prm{};
  bool have_params = false;
  double omega_p = 1.0, omega_u = 1.2, omega_m = 1.0;  // PS:178, DS:303, SP:212
  int precond_kind = -1;                                 // -1 SSOR (reference), 0 Jacobi
  // mesh
  int dim = 0;
  int64_t n_vertices = 0, n_cells = 0, n_bfaces = 0;
  vector<double> xyz;
  vector<int32_t> cell_vertices, bface_cell, bface_id;
  vector<int8_t> bface_local;
  // dofs
  int64_t np = 0, nu = 0;
  int nloc_p = 0, nloc_u = 0, ns_u = 0;
  vector<int32_t> cd_p, cd_u;
  // constraints (displacement)
  vector<int32_t> u_cline;  // dof -> line index or -1
  vector<int32_t> line_dof;
  vector<double> line_g;
  Lines cu, cp;  // full tables (hanging-node entries included); cp is empty on uniform meshes (PS:71-78)
  // neumann
  vector<int32_t> nm_label, nm_comp;
  vector<double> nm_value;
  // matrices
  Pattern Pp, Pu;
  Vec M, K, J, A, PM;
  bool rebuild_system_matrix = true;
  double jac_dt = -1;
  // vectors
  Vec p, p_old, dp, resid, tmp1, tmp2, ev, ev0, frhs, u, b;
  vector<Vec> strains, proj_rhs, stresses;
  int n_stress = 0;
  pe_stats st{};
  // FE tables
  Quad q2, qu;          // QGauss(2) (pressure, projection), QGauss(degree_u+1) (displacement)
  Shape geo_q2, geo_qu; // Q1 mapping shape at q2 / qu
  Shape p_q2, p_qu;     // pressure FE at q2 / qu
  Shape u_q2, u_qu;     // displacement scalar FE at q2 / qu
  bool setup_done = false;
};

void cell_geometry(const Ctx& C, int64_t cell, const Shape& geo, const Quad& Q, CellGeom& G) {
  const int dim = C.dim, vpc = 1 << dim;
  G.JxW.resize(Q.n);
  G.JinvT.resize((size_t)Q.n * dim * dim);
  G.xq.resize((size_t)Q.n * dim);
  const int32_t* cv = &C.cell_vertices[cell * vpc];
  for (int q = 0; q < Q.n; ++q) {
    double J[9] = {0}, inv[9];
    double x[3] = {0, 0, 0};
    for (int v = 0; v < vpc; ++v) {
      const double* X = &C.xyz[(int64_t)cv[v] * dim];
      const double* dN = &geo.dN[((size_t)q * vpc + v) * dim];
      double Nv = geo.N[(size_t)q * vpc + v];
      for (int a = 0; a < dim; ++a) {
        x[a] += Nv * X[a];
        for (int b = 0; b < dim; ++b) J[a * dim + b] += X[a] * dN[b];  // J_ab = dx_a/dxi_b
      }
    }
    double det = det_inv(dim, J, inv);
    G.JxW[q] = det * Q.w[q];
    for (int a = 0; a < dim; ++a) {
      G.xq[(size_t)q * dim + a] = x[a];
      for (int b = 0; b < dim; ++b) G.JinvT[((size_t)q * dim + a) * dim + b] = inv[b * dim + a];
    }
  }
}

// physical gradient of scalar shape s at q: grad = J^-T dN
inline void phys_grad(int dim, const double* JinvT, const double* dN, double* g) {
  for (int a = 0; a < dim; ++a) {
    double s = 0;
    for (int b = 0; b < dim; ++b) s += JinvT[a * dim + b] * dN[b];
    g[a] = s;
  }
}

// right_hand_side.h:99-116
inline double well_value(const Ctx& C, const double* x) {
  double r2 = x[0] * x[0] + x[1] * x[1];
  double rw = C.prm.well_radius;
  if (r2 <= rw * rw) return -C.prm.flow_rate / (3.1415926 * rw * rw);
  return 0;
}

// Cell schedule: one list in natural order when serial (faithful to the reference's loops); with
// po_set_threads(>1) a greedy vertex colouring so that cells of one list never share a dof and an
// OpenMP loop over a list is race free.
vector<vector<int64_t>> cell_schedule(const Ctx& C) {
  vector<vector<int64_t>> lists;
  if (g_threads <= 1) {
    lists.emplace_back(C.n_cells);
    for (int64_t c = 0; c < C.n_cells; ++c) lists[0][c] = c;
    return lists;
  }
  const int vpc = 1 << C.dim;
  vector<uint64_t> vmask(C.n_vertices, 0);
  for (int64_t c = 0; c < C.n_cells; ++c) {
    uint64_t used = 0;
    for (int v = 0; v < vpc; ++v) used |= vmask[C.cell_vertices[c * vpc + v]];
    int col = 0;
    while (col < 63 && ((used >> col) & 1)) ++col;
    if ((int)lists.size() <= col) lists.resize(col + 1);
    lists[col].push_back(c);
    for (int v = 0; v < vpc; ++v) vmask[C.cell_vertices[c * vpc + v]] |= (uint64_t)1 << col;
  }
  return lists;
}

void assemble_mass_laplace(Ctx& C) {  // PS:96-101
  const int dim = C.dim, ns = C.nloc_p;
  C.M.assign(C.Pp.nnz(), 0);
  C.K.assign(C.Pp.nnz(), 0);
  C.frhs.assign(C.np, 0);
  for (const auto& list : cell_schedule(C)) {
#pragma omp parallel num_threads(g_threads) if (g_threads > 1)
  {
  CellGeom G;
  vector<double> gr((size_t)ns * dim), cm((size_t)ns * ns), ck((size_t)ns * ns), cf(ns);
#pragma omp for schedule(static)
  for (int64_t li = 0; li < (int64_t)list.size(); ++li) {
    const int64_t c = list[li];
    cell_geometry(C, c, C.geo_q2, C.q2, G);
    std::fill(cm.begin(), cm.end(), 0);
    std::fill(ck.begin(), ck.end(), 0);
    std::fill(cf.begin(), cf.end(), 0);
    for (int q = 0; q < C.q2.n; ++q) {
      for (int s = 0; s < ns; ++s) phys_grad(dim, &G.JinvT[(size_t)q * dim * dim], &C.p_q2.dN[((size_t)q * ns + s) * dim], &gr[(size_t)s * dim]);
      double fq = well_value(C, &G.xq[(size_t)q * dim]);
      for (int i = 0; i < ns; ++i) {
        double Ni = C.p_q2.N[(size_t)q * ns + i];
        cf[i] += fq * Ni * G.JxW[q];
        for (int j = 0; j < ns; ++j) {
          double gg = 0;
          for (int a = 0; a < dim; ++a) gg += gr[(size_t)i * dim + a] * gr[(size_t)j * dim + a];
          cm[(size_t)i * ns + j] += Ni * C.p_q2.N[(size_t)q * ns + j] * G.JxW[q];
          ck[(size_t)i * ns + j] += gg * G.JxW[q];
        }
      }
    }
    const int32_t* cd = &C.cd_p[c * ns];
    for (int i = 0; i < ns; ++i) {
      C.frhs[cd[i]] += cf[i];
      for (int j = 0; j < ns; ++j) {
        int64_t pos = C.Pp.find(cd[i], cd[j]);
        C.M[pos] += cm[(size_t)i * ns + j];
        C.K[pos] += ck[(size_t)i * ns + j];
      }
    }
  }
  }  // omp parallel
  }  // schedule
}

// symmetric 2-tensor helpers; storage index e(i,j) for i<=j: TensorIndexer (TI:25-30)
inline int sym_entry(int dim, int i, int j) {
  static const int m2[4] = {0, 1, 1, 2}, m3[9] = {0, 1, 2, 1, 3, 4, 2, 4, 5};
  return dim == 2 ? m2[i * 2 + j] : m3[i * 3 + j];
}

// DS:155-291
void displacement_assemble(Ctx& C) {
  const int dim = C.dim, ns = C.ns_u, nl = C.nloc_u, nsp = C.nloc_p;
  const double lam = C.prm.lame_lambda, mu = C.prm.shear_modulus, alpha = C.prm.biot_coef;
  const bool build = C.rebuild_system_matrix;
  if (build) C.A.assign(C.Pu.nnz(), 0);
  C.b.assign(C.nu, 0);
  const Quad& Q = C.qu;
  // boundary faces per cell (for Neumann), built lazily
  vector<vector<int>> cell_bf;
  if (!C.nm_label.empty()) {
    cell_bf.resize(C.n_cells);
    for (int64_t f = 0; f < C.n_bfaces; ++f) cell_bf[C.bface_cell[f]].push_back((int)f);
  }
  for (const auto& list : cell_schedule(C)) {
#pragma omp parallel num_threads(g_threads) if (g_threads > 1)
  {
  CellGeom G;
  vector<double> eps((size_t)nl * dim * dim), sig((size_t)nl * dim * dim), cm((size_t)nl * nl), cr(nl), gr(dim);
  vector<double> pq(Q.n);
#pragma omp for schedule(static)
  for (int64_t li = 0; li < (int64_t)list.size(); ++li) {
    const int64_t c = list[li];
    const int32_t* cd = &C.cd_u[c * nl];
    const int32_t* cdp = &C.cd_p[c * nsp];
    bool need_matrix = build;
    if (!need_matrix)
      for (int i = 0; i < nl; ++i) {
        int li = C.u_cline[cd[i]];
        if (li >= 0 && C.line_g[li] != 0) { need_matrix = true; break; }
      }
    cell_geometry(C, c, C.geo_qu, Q, G);
    std::fill(cr.begin(), cr.end(), 0);
    if (need_matrix) std::fill(cm.begin(), cm.end(), 0);
    for (int q = 0; q < Q.n; ++q) {  // pressure_fe_values.get_function_values (DS:211-212)
      double s = 0;
      for (int k = 0; k < nsp; ++k) s += C.p[cdp[k]] * C.p_qu.N[(size_t)q * nsp + k];
      pq[q] = s;
    }
    for (int q = 0; q < Q.n; ++q) {
      const double jxw = G.JxW[q];
      // get_strain_tensor(fe_values, i, q) for all i (CM:9-24): eps = sym(e_c (x) grad N_s)
      for (int s = 0; s < ns; ++s) {
        phys_grad(dim, &G.JinvT[(size_t)q * dim * dim], &C.u_qu.dN[((size_t)q * ns + s) * dim], gr.data());
        for (int cc = 0; cc < dim; ++cc) {
          int i = s * dim + cc;
          double* e = &eps[(size_t)i * dim * dim];
          for (int a = 0; a < dim * dim; ++a) e[a] = 0;
          for (int a = 0; a < dim; ++a) {  // grad of component cc only: d(phi_i)_cc/dx_a = gr[a]
            e[cc * dim + a] += 0.5 * gr[a];
            e[a * dim + cc] += 0.5 * gr[a];
          }
          double tr = 0;
          for (int a = 0; a < dim; ++a) tr += e[a * dim + a];
          // gassman_tensor * eps (CM:45-57): sigma = lambda tr I + 2 mu eps
          double* sg = &sig[(size_t)i * dim * dim];
          for (int a = 0; a < dim * dim; ++a) sg[a] = 2 * mu * e[a];
          for (int a = 0; a < dim; ++a) sg[a * dim + a] += lam * tr;
          // body force is identically zero (RHS:69-71 in 2D; 3D restated as zero)
          cr[i] += (alpha * pq[q] * tr) * jxw;  // DS:232-234
        }
      }
      if (need_matrix)
        for (int i = 0; i < nl; ++i) {
          const double* sg = &sig[(size_t)i * dim * dim];
          for (int j = 0; j < nl; ++j) {
            const double* e = &eps[(size_t)j * dim * dim];
            double s = 0;
            for (int a = 0; a < dim * dim; ++a) s += sg[a] * e[a];
            cm[(size_t)i * nl + j] += s * jxw;  // DS:240-241
          }
        }
    }
    // Neumann faces DS:249-277
    if (!C.nm_label.empty())
      for (int bf : cell_bf[c]) {
        int f = C.bface_local[bf], axis = f / 2, side = f % 2;
        for (size_t l = 0; l < C.nm_label.size(); ++l) {
          if (C.bface_id[bf] != C.nm_label[l]) continue;
          Quad Qf = qgauss(dim - 1, C.prm.degree_u + 1);
          for (int qf = 0; qf < Qf.n; ++qf) {
            double xi[3];
            int k = 0;
            for (int a = 0; a < dim; ++a) xi[a] = (a == axis) ? (double)side : Qf.pts[(size_t)qf * (dim - 1) + k++];
            vector<double> pt(xi, xi + dim);
            Shape gsh = make_shape(dim, 1, pt, 1), ush = make_shape(dim, C.prm.degree_u, pt, 1);
            double Jm[9] = {0}, inv[9];
            const int vpc = 1 << dim;
            for (int v = 0; v < vpc; ++v) {
              const double* X = &C.xyz[(int64_t)C.cell_vertices[c * vpc + v] * dim];
              for (int a = 0; a < dim; ++a)
                for (int bb = 0; bb < dim; ++bb) Jm[a * dim + bb] += X[a] * gsh.dN[(size_t)v * dim + bb];
            }
            double det = det_inv(dim, Jm, inv);
            // n ~ J^-T n_ref, dS = |det| |J^-T n_ref| w
            double nv[3], nn = 0;
            for (int a = 0; a < dim; ++a) { nv[a] = inv[axis * dim + a] * (side ? 1.0 : -1.0); nn += nv[a] * nv[a]; }
            nn = std::sqrt(nn);
            double jxwf = std::fabs(det) * nn * Qf.w[qf];
            for (int a = 0; a < dim; ++a) nv[a] /= nn;
            for (int s = 0; s < ns; ++s) {
              int i = s * dim + C.nm_comp[l];
              cr[i] += ush.N[s] * (C.nm_value[l] * nv[C.nm_comp[l]]) * jxwf;
            }
          }
        }
      }
    // ConstraintMatrix::distribute_local_to_global (matrix+rhs DS:281-283; rhs only DS:285-286)
    double avg_diag = 0;
    if (build) {
      for (int i = 0; i < nl; ++i) avg_diag += std::fabs(cm[(size_t)i * nl + i]);
      avg_diag /= nl;
    }
    if (C.cu.any_entries()) {  // general lines: local row i goes to its masters with the line's weights
      const Lines& L = C.cu;
      for (int i = 0; i < nl; ++i) {
        const int32_t li = L.line_of[cd[i]];
        if (li >= 0 && build) {
          double dgl = std::fabs(cm[(size_t)i * nl + i]);
          C.A[C.Pu.find(cd[i], cd[i])] += (dgl != 0 ? dgl : avg_diag);
        }
        double t = cr[i];  // f_i - sum_l a_il g_l over the inhomogeneously constrained local dofs
        for (int j = 0; j < nl; ++j) {
          const int32_t lj = L.line_of[cd[j]];
          if (lj >= 0 && L.g[lj] != 0) t -= cm[(size_t)i * nl + j] * L.g[lj];
        }
        const int64_t i0 = li < 0 ? 0 : L.eptr[li], i1 = li < 0 ? 1 : L.eptr[li + 1];
        for (int64_t pi = i0; pi < i1; ++pi) {
          const int32_t gi = li < 0 ? cd[i] : L.edof[pi];
          const double wi = li < 0 ? 1.0 : L.ew[pi];
          C.b[gi] += wi * t;
          if (!build) continue;
          for (int j = 0; j < nl; ++j) {
            const int32_t lj = L.line_of[cd[j]];
            const int64_t j0 = lj < 0 ? 0 : L.eptr[lj], j1 = lj < 0 ? 1 : L.eptr[lj + 1];
            for (int64_t pj = j0; pj < j1; ++pj) {
              const int32_t gj = lj < 0 ? cd[j] : L.edof[pj];
              const double wj = lj < 0 ? 1.0 : L.ew[pj];
              C.A[C.Pu.find(gi, gj)] += wi * wj * cm[(size_t)i * nl + j];
            }
          }
        }
      }
      continue;
    }
    for (int i = 0; i < nl; ++i) {
      int li = C.u_cline[cd[i]];
      if (li >= 0) {
        if (build) {
          double dgl = std::fabs(cm[(size_t)i * nl + i]);
          C.A[C.Pu.find(cd[i], cd[i])] += (dgl != 0 ? dgl : avg_diag);
        }
        continue;
      }
      double r = cr[i];
      for (int j = 0; j < nl; ++j) {
        int lj = C.u_cline[cd[j]];
        if (lj >= 0) {
          if (C.line_g[lj] != 0) r -= cm[(size_t)i * nl + j] * C.line_g[lj];
        } else if (build)
          C.A[C.Pu.find(cd[i], cd[j])] += cm[(size_t)i * nl + j];
      }
      C.b[cd[i]] += r;
    }
  }
  }  // omp parallel
  }  // schedule
  C.rebuild_system_matrix = false;
}

// SP:109-198
void projection_rhs(Ctx& C, int n_comp, const int32_t* comps) {
  const int dim = C.dim, nsp = C.nloc_p, ns = C.ns_u, nl = C.nloc_u;
  vector<int> entries(n_comp);
  for (int c = 0; c < n_comp; ++c) {
    entries[c] = sym_entry(dim, comps[c] / dim, comps[c] % dim);
    C.proj_rhs[entries[c]].assign(C.np, 0);
  }
  for (const auto& list : cell_schedule(C)) {
#pragma omp parallel num_threads(g_threads) if (g_threads > 1)
  {
  CellGeom G;
  vector<double> cr((size_t)n_comp * nsp), gr(dim);
#pragma omp for schedule(static)
  for (int64_t li = 0; li < (int64_t)list.size(); ++li) {
    const int64_t cell = list[li];
    cell_geometry(C, cell, C.geo_q2, C.q2, G);
    std::fill(cr.begin(), cr.end(), 0);
    const int32_t* cd = &C.cd_u[cell * nl];
    const int32_t* cdp = &C.cd_p[cell * nsp];
    for (int q = 0; q < C.q2.n; ++q) {
      double grad[9] = {0};  // grad[c*dim+a] = d u_c / d x_a  (get_function_gradients)
      for (int s = 0; s < ns; ++s) {
        phys_grad(dim, &G.JinvT[(size_t)q * dim * dim], &C.u_q2.dN[((size_t)q * ns + s) * dim], gr.data());
        for (int cc = 0; cc < dim; ++cc) {
          double uv = C.u[cd[s * dim + cc]];
          for (int a = 0; a < dim; ++a) grad[cc * dim + a] += uv * gr[a];
        }
      }
      double strain[9];
      for (int i = 0; i < dim; ++i)
        for (int j = 0; j < dim; ++j) strain[i * dim + j] = (i == j) ? grad[i * dim + i] : (grad[i * dim + j] + grad[j * dim + i]) / 2;  // CM:27-42
      for (int i = 0; i < nsp; ++i) {
        double phi = C.p_q2.N[(size_t)q * nsp + i];
        for (int c = 0; c < n_comp; ++c) cr[(size_t)c * nsp + i] += phi * strain[comps[c]] * G.JxW[q];
      }
    }
    for (int c = 0; c < n_comp; ++c)
      for (int i = 0; i < nsp; ++i) {  // constraints.distribute_local_to_global(cell_rhs[c], ...), SP:193-194
        const int32_t li = C.cp.n() ? C.cp.line_of[cdp[i]] : -1;
        if (li < 0) { C.proj_rhs[entries[c]][cdp[i]] += cr[(size_t)c * nsp + i]; continue; }
        for (int64_t e = C.cp.eptr[li]; e < C.cp.eptr[li + 1]; ++e) C.proj_rhs[entries[c]][C.cp.edof[e]] += C.cp.ew[e] * cr[(size_t)c * nsp + i];
      }
  }
  }  // omp parallel
  }  // schedule
}

int fail(Ctx* c, int code, const std::string& m) {
  if (c) c->err = m;
  return code;
}

Vec* vec_by_id(Ctx& C, int which) {
  switch (which) {
    case PE_VEC_P: return &C.p;
    case PE_VEC_P_OLD: return &C.p_old;
    case PE_VEC_P_UPDATE: return &C.dp;
    case PE_VEC_P_RESIDUAL: return &C.resid;
    case PE_VEC_VOL_STRAIN: return &C.ev;
    case PE_VEC_VOL_STRAIN0: return &C.ev0;
    case PE_VEC_WELL_RHS: return &C.frhs;
    case PE_VEC_U: return &C.u;
    case PE_VEC_U_RHS: return &C.b;
    default: break;
  }
  if (which >= PE_VEC_STRAIN0 && which < PE_VEC_STRAIN0 + C.n_stress) return &C.strains[which - PE_VEC_STRAIN0];
  if (which >= PE_VEC_PROJ_RHS0 && which < PE_VEC_PROJ_RHS0 + C.n_stress) return &C.proj_rhs[which - PE_VEC_PROJ_RHS0];
  if (which >= PE_VEC_STRESS0 && which < PE_VEC_STRESS0 + C.n_stress) return &C.stresses[which - PE_VEC_STRESS0];
  return nullptr;
}

}  // namespace

extern "C" {

int po_create(Ctx** out) { *out = new Ctx(); return 0; }
void po_destroy(Ctx* c) { delete c; }
const char* po_last_error(const Ctx* c) { return c ? c->err.c_str() : ""; }
int po_set_threads(int n) {
#ifdef _OPENMP
  g_threads = n > 0 ? n : omp_get_max_threads();
#else
  g_threads = 1;
#endif
  return g_threads;
}
int po_set_preconditioner(Ctx* c, int kind) { c->precond_kind = kind; return 0; }

int po_set_params(Ctx* c, const pe_params* p) {
  if (p->dim < 2 || p->dim > 3 || p->degree_p != 1 || p->degree_u < 1 || p->degree_u > 2) return fail(c, PE_ERR_BAD_INPUT, "bad dim/degree");
  c->prm = *p;
  c->have_params = true;
  return 0;
}

int po_upload_mesh(Ctx* c, int dim, int64_t nv, const double* xyz, int64_t nc, const int32_t* cv, int64_t nbf,
                   const int32_t* bc, const int8_t* bl, const int32_t* bid) {
  c->dim = dim;
  c->n_vertices = nv;
  c->n_cells = nc;
  c->n_bfaces = nbf;
  c->xyz.assign(xyz, xyz + nv * dim);
  c->cell_vertices.assign(cv, cv + nc * (1 << dim));
  c->bface_cell.assign(bc, bc + nbf);
  c->bface_local.assign(bl, bl + nbf);
  c->bface_id.assign(bid, bid + nbf);
  return 0;
}

int po_upload_dofs(Ctx* c, int field, int64_t n, const int32_t* cd) {
  if (!c->have_params || !c->n_cells) return fail(c, PE_ERR_STATE, "params and mesh first");
  int dim = c->dim;
  if (field == PE_FIELD_PRESSURE) {
    c->np = n;
    c->nloc_p = 1 << dim;
    c->cd_p.assign(cd, cd + c->n_cells * c->nloc_p);
    c->cp = Lines();
  } else {
    c->nu = n;
    c->ns_u = (int)unit_support(dim, c->prm.degree_u).size() / dim;
    c->nloc_u = c->ns_u * dim;
    c->cd_u.assign(cd, cd + c->n_cells * c->nloc_u);
    c->u_cline.assign(n, -1);
    c->line_dof.clear();
    c->line_g.clear();
    c->cu = Lines();
  }
  c->setup_done = false;
  return 0;
}

int po_upload_constraints(Ctx* c, int field, int64_t nl, const int32_t* ld, const int64_t* eptr, const int32_t* edof, const double* ew,
                          const double* inh) {
  if (field == PE_FIELD_PRESSURE) {
    if (!c->np) return fail(c, PE_ERR_STATE, "upload the pressure dofs first");
    for (int64_t i = 0; i < nl; ++i)
      if (inh && inh[i] != 0) return fail(c, PE_ERR_UNSUPPORTED, "pressure constraints are homogeneous hanging-node lines (PS:71-78)");
    c->cp.set(c->np, nl, ld, eptr, edof, ew, inh);
    return 0;
  }
  if (!c->nu) return fail(c, PE_ERR_STATE, "upload the displacement dofs first");
  c->cu.set(c->nu, nl, ld, eptr, edof, ew, inh);
  c->line_dof.assign(ld, ld + nl);
  c->line_g = c->cu.g;
  c->u_cline = c->cu.line_of;
  return 0;
}

int po_upload_neumann(Ctx* c, int n, const int32_t* l, const int32_t* comp, const double* v) {
  c->nm_label.assign(l, l + n);
  c->nm_comp.assign(comp, comp + n);
  c->nm_value.assign(v, v + n);
  return 0;
}

int po_setup(Ctx* c) {
  if (!c->np || !c->nu) return fail(c, PE_ERR_STATE, "dofs not uploaded");
  const int dim = c->dim;
  c->q2 = qgauss(dim, 2);
  c->qu = qgauss(dim, c->prm.degree_u + 1);
  c->geo_q2 = make_shape(dim, 1, c->q2.pts, c->q2.n);
  c->geo_qu = make_shape(dim, 1, c->qu.pts, c->qu.n);
  c->p_q2 = c->geo_q2;
  c->p_qu = c->geo_qu;
  c->u_q2 = make_shape(dim, c->prm.degree_u, c->q2.pts, c->q2.n);
  c->u_qu = make_shape(dim, c->prm.degree_u, c->qu.pts, c->qu.n);
  if (c->cp.any_entries() || c->cu.any_entries()) g_threads = 1;  // the vertex colouring of cell_schedule ignores master dofs
  if (c->cp.line_of.empty()) c->cp.set(c->np, 0, nullptr, nullptr, nullptr, nullptr, nullptr);
  if (c->cu.line_of.empty()) c->cu.set(c->nu, 0, nullptr, nullptr, nullptr, nullptr, nullptr);
  c->Pp = c->cp.any_entries() ? make_pattern_constrained(c->np, c->n_cells, c->nloc_p, c->cd_p.data(), c->cp)
                              : make_pattern(c->np, c->n_cells, c->nloc_p, c->cd_p.data());
  c->Pu = c->cu.any_entries() ? make_pattern_constrained(c->nu, c->n_cells, c->nloc_u, c->cd_u.data(), c->cu)
                              : make_pattern(c->nu, c->n_cells, c->nloc_u, c->cd_u.data());
  assemble_mass_laplace(*c);
  c->J.assign(c->Pp.nnz(), 0);
  c->n_stress = (dim * dim + dim) / 2;
  for (Vec* v : {&c->p, &c->p_old, &c->dp, &c->resid, &c->tmp1, &c->tmp2, &c->ev, &c->ev0}) v->assign(c->np, 0);
  c->u.assign(c->nu, 0);
  c->b.assign(c->nu, 0);
  c->strains.assign(c->n_stress, Vec(c->np, 0));
  c->proj_rhs.assign(c->n_stress, Vec(c->np, 0));
  c->stresses.assign(c->n_stress, Vec(c->np, 0));
  c->rebuild_system_matrix = true;
  c->st = pe_stats{};
  c->st.n_cells = c->n_cells;
  c->st.n_dofs_p = c->np;
  c->st.n_dofs_u = c->nu;
  c->st.nnz_p = c->Pp.nnz();
  c->st.nnz_u = c->Pu.nnz();
  c->setup_done = true;
  return 0;
}

int po_pressure_set_uniform(Ctx* c, double v) { std::fill(c->p.begin(), c->p.end(), v); return 0; }
int po_pressure_begin_step(Ctx* c) { c->p_old = c->p; return 0; }
int po_pressure_zero_update(Ctx* c) { std::fill(c->dp.begin(), c->dp.end(), 0.0); return 0; }

int po_pressure_update_volumetric_strain(Ctx* c) {  // PS:187-194
  double f = c->prm.biot_coef / c->prm.bulk_modulus;
  for (int64_t i = 0; i < c->np; ++i) { c->tmp1[i] = c->dp[i] * f; c->ev[i] += c->tmp1[i]; }
  return 0;
}

int po_pressure_assemble_residual(Ctx* c, double dt, double* l2) {  // PS:113-155
  const int64_t n = c->np;
  const double a = c->prm.biot_coef / dt, m = 1. / c->prm.m_modulus / dt, kappa = c->prm.perm_over_visc;
  for (int64_t i = 0; i < n; ++i) {
    double t1 = (c->ev[i] - c->ev0[i]) * a;
    double t2 = (c->p[i] - c->p_old[i]) * m;
    c->tmp1[i] = t1 + t2;
  }
  vmult(c->Pp, c->M, c->tmp1, c->resid);
  vmult(c->Pp, c->K, c->p, c->tmp1);
  for (int64_t i = 0; i < n; ++i) {
    c->tmp1[i] *= kappa;
    c->resid[i] += c->tmp1[i];
    c->resid[i] += c->frhs[i];
    c->resid[i] *= -1;
  }
  c->cp.condense(c->resid);  // PS:153
  if (l2) *l2 = l2norm(c->resid);
  return 0;
}

int po_pressure_assemble_jacobian(Ctx* c, double dt) {  // PS:158-169
  const double m = 1. / c->prm.m_modulus / dt, f = c->prm.perm_over_visc;
  for (int64_t i = 0; i < c->Pp.nnz(); ++i) c->J[i] = c->M[i] * m + f * c->K[i];
  try { condense_matrix(c->Pp, c->cp, c->J); } catch (const std::exception& e) { return fail(c, PE_ERR_STATE, e.what()); }  // PS:168
  c->jac_dt = dt;
  return 0;
}

int po_pressure_solve(Ctx* c, int* its, double* res) {  // PS:172-185
  double tol = c->prm.cg_rel_tol_pressure * l2norm(c->resid);
  CgResult r = cg_solve(c->Pp, c->J, c->dp, c->resid, c->precond_kind < 0 ? c->omega_p : 0.0, c->prm.cg_max_iterations, tol);
  c->cp.distribute(c->dp);  // PS:180
  c->st.cg_iterations_pressure += r.its;
  c->st.cg_solves_pressure++;
  if (its) *its = r.its;
  if (res) *res = r.res;
  return r.ok ? 0 : fail(c, std::isnan(r.res) ? PE_ERR_NAN : PE_ERR_NO_CONVERGENCE, "pressure CG: SolverControl::NoConvergence");
}

int po_pressure_add_update(Ctx* c) { for (int64_t i = 0; i < c->np; ++i) c->p[i] += c->dp[i]; return 0; }
int po_pressure_linfty(Ctx* c, double* v) { double m = 0; for (double x : c->p) m = std::max(m, std::fabs(x)); *v = m; return 0; }

int po_displacement_assemble(Ctx* c) { displacement_assemble(*c); return 0; }

int po_displacement_solve(Ctx* c, int* its, double* res) {  // DS:294-307
  CgResult r = cg_solve(c->Pu, c->A, c->u, c->b, c->precond_kind < 0 ? c->omega_u : 0.0, c->prm.cg_max_iterations, c->prm.cg_abs_tol_displacement);
  c->cu.distribute(c->u);  // constraints.distribute(solution), DS:306
  c->st.cg_iterations_displacement += r.its;
  c->st.cg_solves_displacement++;
  if (its) *its = r.its;
  if (res) *res = r.res;
  return r.ok ? 0 : fail(c, std::isnan(r.res) ? PE_ERR_NAN : PE_ERR_NO_CONVERGENCE, "displacement CG: SolverControl::NoConvergence");
}

int po_project_assemble_matrix(Ctx* c) {  // SP:101-106
  c->PM = c->M;
  try { condense_matrix(c->Pp, c->cp, c->PM); } catch (const std::exception& e) { return fail(c, PE_ERR_STATE, e.what()); }
  return 0;
}

int po_project_assemble_rhs(Ctx* c, int n, const int32_t* comps) { projection_rhs(*c, n, comps); return 0; }

int po_project_solve(Ctx* c, int entry, int* its) {  // SP:201-232
  if (c->PM.empty()) return fail(c, PE_ERR_STATE, "projection matrix not assembled");
  double tol = c->prm.cg_rel_tol_projection * l2norm(c->proj_rhs[entry]);
  CgResult r = cg_solve(c->Pp, c->PM, c->strains[entry], c->proj_rhs[entry], c->precond_kind < 0 ? c->omega_m : 0.0, c->prm.cg_max_iterations, tol);
  c->cp.distribute(c->strains[entry]);  // SP:215
  c->st.cg_iterations_projection += r.its;
  c->st.cg_solves_projection++;
  if (its) *its = r.its;
  return r.ok ? 0 : fail(c, PE_ERR_NO_CONVERGENCE, "projection CG: SolverControl::NoConvergence");
}

int po_volumetric_strain_from_projection(Ctx* c, int n, const int32_t* entries, int as_initial) {  // FSS:179-186, 317
  std::fill(c->ev.begin(), c->ev.end(), 0.0);
  for (int k = 0; k < n; ++k)
    for (int64_t i = 0; i < c->np; ++i) c->ev[i] += c->strains[entries[k]][i];
  if (as_initial) c->ev0 = c->ev;
  return 0;
}

int po_effective_stresses(Ctx* c) {  // FSS:189-224
  const int dim = c->dim;
  const double lam = c->prm.lame_lambda, mu = c->prm.shear_modulus;
  for (int64_t l = 0; l < c->np; ++l) {
    double tr = 0;
    for (int i = 0; i < dim; ++i) tr += c->strains[sym_entry(dim, i, i)][l];
    for (int i = 0; i < dim; ++i)
      for (int j = i; j < dim; ++j) {
        int e = sym_entry(dim, i, j);
        c->stresses[e][l] = 2 * mu * c->strains[e][l] + (i == j ? lam * tr : 0.0);
      }
  }
  return 0;
}

int po_get_vector(Ctx* c, int which, double* host, int64_t n) {
  Vec* v = vec_by_id(*c, which);
  if (!v || (int64_t)v->size() != n) return fail(c, PE_ERR_BAD_INPUT, "bad vector id/size");
  std::memcpy(host, v->data(), n * sizeof(double));
  return 0;
}
int po_set_vector(Ctx* c, int which, const double* host, int64_t n) {
  Vec* v = vec_by_id(*c, which);
  if (!v || (int64_t)v->size() != n) return fail(c, PE_ERR_BAD_INPUT, "bad vector id/size");
  std::memcpy(v->data(), host, n * sizeof(double));
  return 0;
}

static const Pattern* mat_by_id(Ctx& C, int m, const Vec** val) {
  switch (m) {
    case PE_MAT_MASS: *val = &C.M; return &C.Pp;
    case PE_MAT_LAPLACE: *val = &C.K; return &C.Pp;
    case PE_MAT_JACOBIAN: *val = &C.J; return &C.Pp;
    case PE_MAT_ELASTICITY: *val = &C.A; return &C.Pu;
    case PE_MAT_PROJECTION: *val = &C.PM; return &C.Pp;
  }
  return nullptr;
}
int po_get_matrix_size(Ctx* c, int m, int64_t* n, int64_t* nnz) {
  const Vec* v;
  const Pattern* P = mat_by_id(*c, m, &v);
  if (!P) return PE_ERR_BAD_INPUT;
  *n = P->n;
  *nnz = P->nnz();
  return 0;
}
// returns deal.II layout (diagonal first); the tests canonicalise through scipy
int po_get_matrix(Ctx* c, int m, int64_t* rowptr, int32_t* col, double* val) {
  const Vec* v;
  const Pattern* P = mat_by_id(*c, m, &v);
  if (!P || v->empty()) return fail(c, PE_ERR_STATE, "matrix not assembled");
  std::memcpy(rowptr, P->rowptr.data(), (P->n + 1) * sizeof(int64_t));
  std::memcpy(col, P->col.data(), P->nnz() * sizeof(int32_t));
  std::memcpy(val, v->data(), P->nnz() * sizeof(double));
  return 0;
}
int po_get_stats(Ctx* c, pe_stats* s) { *s = c->st; return 0; }
int po_reset_stats(Ctx* c) {
  pe_stats k = c->st;
  c->st = pe_stats{};
  c->st.n_cells = k.n_cells; c->st.n_dofs_p = k.n_dofs_p; c->st.n_dofs_u = k.n_dofs_u; c->st.nnz_p = k.nnz_p; c->st.nnz_u = k.nnz_u;
  return 0;
}

// SolverCG on a caller-supplied CSR matrix (columns ascending): lets the tests pin cg_solve / ssor_apply against published
// deal.II results (tutorial step-4).  omega > 0 SSOR, == 0 Jacobi, < 0 identity.
int po_cg_csr(int64_t n, const int64_t* rowptr, const int32_t* col, const double* val, double* x, const double* b, double omega, int max_steps,
              double tol, int* its, double* res) {
  Pattern P;
  P.n = n;
  P.rowptr.assign(rowptr, rowptr + n + 1);
  P.col.resize(rowptr[n]);
  P.right_of_diag.resize(n);
  Vec A(rowptr[n]);
  for (int64_t r = 0; r < n; ++r) {  // deal.II layout: diagonal first, then ascending
    int64_t p = rowptr[r] + 1;
    bool have_diag = false;
    P.right_of_diag[r] = rowptr[r + 1];
    bool found = false;
    for (int64_t j = rowptr[r]; j < rowptr[r + 1]; ++j) {
      if (col[j] == r) { P.col[rowptr[r]] = (int32_t)r; A[rowptr[r]] = val[j]; have_diag = true; continue; }
      if (p >= rowptr[r + 1] && !have_diag) return PE_ERR_BAD_INPUT;
      if (!found && col[j] > r) { P.right_of_diag[r] = p; found = true; }
      P.col[p] = col[j];
      A[p] = val[j];
      ++p;
    }
    if (!have_diag) return PE_ERR_BAD_INPUT;
  }
  Vec xv(x, x + n), bv(b, b + n);
  CgResult R = cg_solve(P, A, xv, bv, omega, max_steps, tol);
  std::memcpy(x, xv.data(), n * sizeof(double));
  if (its) *its = R.its;
  if (res) *res = R.res;
  return R.ok ? 0 : PE_ERR_NO_CONVERGENCE;
}

// timing helpers for the CPU baseline: n repetitions of one operator on the assembled matrices
int po_time_vmult(Ctx* c, int matrix, int reps) {
  const Vec* v;
  const Pattern* P = mat_by_id(*c, matrix, &v);
  if (!P || v->empty()) return PE_ERR_STATE;
  Vec x(P->n, 1.0), y(P->n);
  for (int r = 0; r < reps; ++r) vmult(*P, *v, x, y);
  return 0;
}
int po_time_ssor(Ctx* c, int matrix, double omega, int reps) {
  const Vec* v;
  const Pattern* P = mat_by_id(*c, matrix, &v);
  if (!P || v->empty()) return PE_ERR_STATE;
  Vec x(P->n, 1.0), y(P->n);
  for (int r = 0; r < reps; ++r) ssor_apply(*P, *v, omega, x, y);
  return 0;
}

}  // extern "C"
