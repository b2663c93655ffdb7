// shim.h — a minimal stand-in for the slice of the deal.II 8.4 API that the reference's headers use.  NOT deal.II.
//
// TEST INFRASTRUCTURE ONLY (like everything under oracle/): it exists so that the reference's own, unmodified source files
//     /root/reference/lib/include/{PoroelasticityFSS,PoroElasticPressureSolver,PoroElasticDisplacementSolver,StrainProjector,
//                                  ConstitutiveModel,TensorIndexer,BoundaryConditions,right_hand_side,InputDataPoroel,
//                                  parse_command_line}.h
// can be compiled where they lie (oracle/Makefile, target _ref/fss_ref; driver oracle/ref_main.cpp = the Runner.cpp that
// code/CMakeLists.txt:8 names and the repository does not contain) and `PoroElasticProblem<dim>::run()` can execute here.  What
// runs is then the reference's application code — its loops, formulas, tolerances, call order, quirks — on top of this
// re-statement of the library calls it makes.  The numbers it prints pin the oracle (tests/golden/make_reference_run.py,
// tests/test_reference_run.py).  deal.II itself is not installed and cannot be fetched; where this file restates a deal.II
// algorithm the comment names the deal.II 8.4 function it follows.  Nothing in the product or in oracle/oracle.cpp includes it.
//
// Scope: meshes that GridGenerator::hyper_rectangle + refine_global produce (any level, 2D and 3D, Q1 mapping), FE_Q(1|2) and
// FESystem(FE_Q(k), n), QGauss, FEValues / FEFaceValues, DoFHandler with deal.II's cell-by-cell first-touch numbering,
// SparsityPattern (diagonal first) / SparseMatrix (vmult, precondition_SSOR), ConstraintMatrix for inhomogeneous Dirichlet
// lines, MatrixCreator mass / Laplace, VectorTools::create_right_hand_side / interpolate_boundary_values, SolverControl /
// SolverCG / PreconditionSSOR, ParameterHandler, a DataOut that dumps every attached vector with the support points of its
// dofs at full precision.  Everything that needs hanging nodes (KellyErrorEstimator, GridRefinement, SolutionTransfer,
// execute_coarsening_and_refinement) and GridIn is declared so that the reference compiles, and throws if reached: the runs
// taken with this shim stop before the reference's first refinement (time step 5, FSS:333).
#pragma once
#include <algorithm>
#include <array>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <list>
#include <map>
#include <memory>
#include <set>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace dealii {

// ---------------------------------------------------------------------------------------------- exceptions (release mode)
struct ShimException : std::runtime_error { using std::runtime_error::runtime_error; };
inline ShimException ExcInternalError() { return ShimException("ExcInternalError"); }
inline ShimException ExcNotImplemented() { return ShimException("ExcNotImplemented"); }
inline ShimException ExcDivideByZero() { return ShimException("ExcDivideByZero"); }
template <class A, class B> inline ShimException ExcDimensionMismatch(A a, B b) {
  return ShimException("ExcDimensionMismatch(" + std::to_string((long long)a) + ", " + std::to_string((long long)b) + ")");
}
#define Assert(cond, exc) do { } while (false)  /* deal.II's Assert is compiled out in release mode */
#define AssertThrow(cond, exc) do { if (!(cond)) throw (exc); } while (false)
[[noreturn]] inline void shim_unsupported(const char* what) {
  throw ShimException(std::string("deal.II shim: ") + what + " is outside the shim (uniform meshes only)");
}

namespace types { typedef unsigned int global_dof_index; typedef unsigned char boundary_id; }
namespace numbers {
const types::global_dof_index invalid_dof_index = static_cast<types::global_dof_index>(-1);
const types::boundary_id internal_face_boundary_id = static_cast<types::boundary_id>(-1);
}
struct LogStream { void depth_console(int) {} };
static LogStream deallog;
namespace Utilities {
inline std::string int_to_string(unsigned int i, unsigned int digits) {
  std::string s = std::to_string(i);
  while (s.size() < digits) s = "0" + s;
  return s;
}
}

// ---------------------------------------------------------------------------------------------- tensors and points
template <int rank, int dim> class Tensor;
template <int dim> class Tensor<1, dim> {
 public:
  Tensor() { for (int i = 0; i < dim; ++i) v[i] = 0; }
  double& operator[](unsigned int i) { return v[i]; }
  const double& operator[](unsigned int i) const { return v[i]; }
  Tensor& operator+=(const Tensor& o) { for (int i = 0; i < dim; ++i) v[i] += o.v[i]; return *this; }
  Tensor& operator-=(const Tensor& o) { for (int i = 0; i < dim; ++i) v[i] -= o.v[i]; return *this; }
  Tensor& operator*=(double s) { for (int i = 0; i < dim; ++i) v[i] *= s; return *this; }
  double norm() const { double s = 0; for (int i = 0; i < dim; ++i) s += v[i] * v[i]; return std::sqrt(s); }
 protected:
  double v[dim];
};
template <int dim> inline double operator*(const Tensor<1, dim>& a, const Tensor<1, dim>& b) {
  double s = 0;
  for (int i = 0; i < dim; ++i) s += a[i] * b[i];
  return s;
}
template <int dim> class Point : public Tensor<1, dim> {
 public:
  Point() {}
  explicit Point(const Tensor<1, dim>& t) : Tensor<1, dim>(t) {}
  double operator()(unsigned int i) const { return this->v[i]; }
  double& operator()(unsigned int i) { return this->v[i]; }
};

// SymmetricTensor<2,dim>: dim(dim+1)/2 independent entries, diagonal first (deal.II's order: 2D {00,11,01}, 3D {00,11,22,01,02,12}).
// SymmetricTensor<4,dim> is the n x n table of pairs of such entries.  operator[] chains resolve to the shared storage, so a write
// to [i][j] is a write to [j][i] — which is what ConstitutiveModel.h:50-54 relies on when it fills all dim^4 index tuples.
template <int dim> inline int sym_index(unsigned int i, unsigned int j) {
  if (i == j) return (int)i;
  if (i > j) std::swap(i, j);
  if (dim == 2) return 2;
  return i == 0 ? (j == 1 ? 3 : 4) : 5;
}
template <int rank, int dim> class SymmetricTensor;
template <int dim> class SymmetricTensor<2, dim> {
 public:
  static const int n = dim * (dim + 1) / 2;
  SymmetricTensor() { for (int k = 0; k < n; ++k) s[k] = 0; }
  SymmetricTensor& operator=(double zero) { for (int k = 0; k < n; ++k) s[k] = zero; return *this; }
  struct Row {
    SymmetricTensor* t; unsigned int i;
    double& operator[](unsigned int j) { return t->s[sym_index<dim>(i, j)]; }
  };
  struct ConstRow {
    const SymmetricTensor* t; unsigned int i;
    const double& operator[](unsigned int j) const { return t->s[sym_index<dim>(i, j)]; }
  };
  Row operator[](unsigned int i) { return Row{this, i}; }
  ConstRow operator[](unsigned int i) const { return ConstRow{this, i}; }
  double s[n];
};
template <int dim> inline double trace(const SymmetricTensor<2, dim>& t) {
  double s = 0;
  for (int i = 0; i < dim; ++i) s += t.s[i];
  return s;
}
// scalar product a:b — off-diagonal entries count twice (symmetric_tensor.h, operator* (SymmetricTensor<2>, SymmetricTensor<2>))
template <int dim> inline double operator*(const SymmetricTensor<2, dim>& a, const SymmetricTensor<2, dim>& b) {
  double s = 0;
  for (int k = 0; k < dim; ++k) s += a.s[k] * b.s[k];
  for (int k = dim; k < SymmetricTensor<2, dim>::n; ++k) s += 2 * a.s[k] * b.s[k];
  return s;
}
template <int dim> class SymmetricTensor<4, dim> {
 public:
  static const int n = dim * (dim + 1) / 2;
  SymmetricTensor() { for (int a = 0; a < n; ++a) for (int b = 0; b < n; ++b) s[a][b] = 0; }
  struct I3 { SymmetricTensor* t; int ab; unsigned int k; double& operator[](unsigned int l) { return t->s[ab][sym_index<dim>(k, l)]; } };
  struct I2 { SymmetricTensor* t; int ab; I3 operator[](unsigned int k) { return I3{t, ab, k}; } };
  struct I1 { SymmetricTensor* t; unsigned int i; I2 operator[](unsigned int j) { return I2{t, sym_index<dim>(i, j)}; } };
  I1 operator[](unsigned int i) { return I1{this, i}; }
  double s[n][n];
};
// double contraction C:e over the last index pair (symmetric_tensor.h): off-diagonal entries of e count twice
template <int dim> inline SymmetricTensor<2, dim> operator*(const SymmetricTensor<4, dim>& c, const SymmetricTensor<2, dim>& e) {
  SymmetricTensor<2, dim> r;
  const int n = SymmetricTensor<2, dim>::n;
  for (int a = 0; a < n; ++a) {
    double s = 0;
    for (int b = 0; b < dim; ++b) s += c.s[a][b] * e.s[b];
    for (int b = dim; b < n; ++b) s += 2 * c.s[a][b] * e.s[b];
    r.s[a] = s;
  }
  return r;
}

// ---------------------------------------------------------------------------------------------- Vector, FullMatrix
// One spare entry is always allocated behind the last one: BodyForces::vector_value (right_hand_side.h:76-82) writes
// values(direction) with direction = 3 into a Vector of size dim whenever dim == 3 — one past the end.  deal.II's release mode
// does not check the index either; the spare entry keeps that stray write inside owned memory (it is never read).
template <typename Number> class Vector {
 public:
  Vector() { v.reserve(1); }
  explicit Vector(unsigned int n) { v.reserve(n + 1); v.assign(n, Number(0)); }
  Vector(const Vector& o) { v.reserve(o.v.size() + 1); v = o.v; }
  Vector& operator=(const Vector& o) { v.reserve(o.v.size() + 1); v = o.v; return *this; }
  void reinit(unsigned int n) { v.reserve(n + 1); v.assign(n, Number(0)); }
  void reinit(const Vector& o) { reinit((unsigned int)o.v.size()); }  // same size, zeroed (vector.h: reinit(V, omit_zeroing = false))
  unsigned int size() const { return (unsigned int)v.size(); }
  Number& operator()(unsigned int i) { return v.data()[i]; }
  const Number& operator()(unsigned int i) const { return v.data()[i]; }
  Number& operator[](unsigned int i) { return v.data()[i]; }
  const Number& operator[](unsigned int i) const { return v.data()[i]; }
  Vector& operator=(Number s) { std::fill(v.begin(), v.end(), s); return *this; }
  Vector& operator+=(const Vector& o) { for (size_t i = 0; i < v.size(); ++i) v[i] += o.v[i]; return *this; }
  Vector& operator-=(const Vector& o) { for (size_t i = 0; i < v.size(); ++i) v[i] -= o.v[i]; return *this; }
  Vector& operator*=(Number s) { for (size_t i = 0; i < v.size(); ++i) v[i] *= s; return *this; }
  Number operator*(const Vector& o) const { Number s = 0; for (size_t i = 0; i < v.size(); ++i) s += v[i] * o.v[i]; return s; }
  double l2_norm() const { double s = 0; for (Number x : v) s += (double)x * (double)x; return std::sqrt(s); }
  double linfty_norm() const { double m = 0; for (Number x : v) m = std::max(m, std::fabs((double)x)); return m; }
  bool all_zero() const { for (Number x : v) if (x != Number(0)) return false; return true; }
  // BLAS-1 names used by SolverCG (vector.h)
  void add(Number a, const Vector& o) { for (size_t i = 0; i < v.size(); ++i) v[i] += a * o.v[i]; }
  void equ(Number a, const Vector& o) { v.reserve(o.v.size() + 1); v.resize(o.v.size()); for (size_t i = 0; i < v.size(); ++i) v[i] = a * o.v[i]; }
  void sadd(Number s, Number a, const Vector& o) { for (size_t i = 0; i < v.size(); ++i) v[i] = s * v[i] + a * o.v[i]; }
  std::vector<Number> v;
};
template <typename Number> class FullMatrix {
 public:
  FullMatrix(unsigned int m, unsigned int n) : rows(m), cols(n), a((size_t)m * n, Number(0)) {}
  Number& operator()(unsigned int i, unsigned int j) { return a[(size_t)i * cols + j]; }
  const Number& operator()(unsigned int i, unsigned int j) const { return a[(size_t)i * cols + j]; }
  FullMatrix& operator=(Number s) { std::fill(a.begin(), a.end(), s); return *this; }
  unsigned int m() const { return rows; }
  unsigned int n() const { return cols; }
 private:
  unsigned int rows, cols;
  std::vector<Number> a;
};

// ---------------------------------------------------------------------------------------------- functions
template <int dim> class Function {
 public:
  explicit Function(unsigned int n_components = 1) : n_components(n_components) {}
  virtual ~Function() {}
  virtual double value(const Point<dim>&, const unsigned int = 0) const { return 0; }
  const unsigned int n_components;
};
template <int dim> class ZeroFunction : public Function<dim> {
 public:
  explicit ZeroFunction(unsigned int n = 1) : Function<dim>(n) {}
};
template <int dim> class ConstantFunction : public Function<dim> {
 public:
  ConstantFunction(double value, unsigned int n = 1) : Function<dim>(n), c(value) {}
  virtual double value(const Point<dim>&, const unsigned int = 0) const { return c; }
 private:
  double c;
};
template <int dim> struct FunctionMap { typedef std::map<types::boundary_id, const Function<dim>*> type; };

// ---------------------------------------------------------------------------------------------- ParameterHandler
namespace Patterns {
struct PatternBase {
  virtual ~PatternBase() {}
  virtual bool match(const std::string& s) const = 0;
  virtual PatternBase* clone() const = 0;
};
inline std::string shim_trim(const std::string& s) {
  size_t a = s.find_first_not_of(" \t\r\n"), b = s.find_last_not_of(" \t\r\n");
  return a == std::string::npos ? std::string() : s.substr(a, b - a + 1);
}
struct Integer : PatternBase {
  long lo, hi;
  explicit Integer(long lo = -2147483647L, long hi = 2147483647L) : lo(lo), hi(hi) {}
  bool match(const std::string& s) const {
    std::istringstream is(s);
    long v;
    if (!(is >> v)) return false;
    std::string rest;
    if (is >> rest) return false;
    return v >= lo && v <= hi;
  }
  PatternBase* clone() const { return new Integer(*this); }
};
struct Double : PatternBase {
  double lo, hi;
  explicit Double(double lo = -1.7976931348623157e308, double hi = 1.7976931348623157e308) : lo(lo), hi(hi) {}
  bool match(const std::string& s) const {
    std::istringstream is(s);
    double v;
    if (!(is >> v)) return false;
    std::string rest;
    if (is >> rest) return false;
    return v >= lo && v <= hi;
  }
  PatternBase* clone() const { return new Double(*this); }
};
struct List : PatternBase {
  std::shared_ptr<PatternBase> item;
  explicit List(const PatternBase& p) : item(p.clone()) {}
  bool match(const std::string& s) const {  // comma-separated; an empty string is the empty list
    if (shim_trim(s).empty()) return true;
    std::string cur;
    std::istringstream is(s);
    while (std::getline(is, cur, ','))
      if (!item->match(shim_trim(cur))) return false;
    return true;
  }
  PatternBase* clone() const { return new List(*this); }
};
}  // namespace Patterns

class ParameterHandler {
 public:
  enum OutputStyle { Text = 1 };
  void enter_subsection(const std::string& s) { path.push_back(s); }
  void leave_subsection() { path.pop_back(); }
  void declare_entry(const std::string& name, const std::string& def, const Patterns::PatternBase& pattern, const std::string& = "") {
    Entry e;
    e.value = def;
    e.pattern.reset(pattern.clone());
    AssertThrow(e.pattern->match(def), ShimException("ParameterHandler: default of <" + name + "> violates its pattern"));
    entries[key(name)] = e;
  }
  // parameter_handler.cc, read_input: `subsection X` / `set Name = value` / `end`, `#` starts a comment, values are trimmed
  bool read_input(const std::string& filename) {
    std::ifstream in(filename.c_str());
    AssertThrow((bool)in, ShimException("ParameterHandler: cannot open <" + filename + ">"));
    std::string line;
    std::vector<std::string> saved = path;
    int lineno = 0;
    while (std::getline(in, line)) {
      ++lineno;
      const size_t hash = line.find('#');
      if (hash != std::string::npos) line = line.substr(0, hash);
      line = Patterns::shim_trim(line);
      if (line.empty()) continue;
      if (line.compare(0, 11, "subsection ") == 0) { path.push_back(Patterns::shim_trim(line.substr(11))); continue; }
      if (line == "end") {
        AssertThrow(path.size() > saved.size(), ShimException("ParameterHandler: unbalanced `end` in line " + std::to_string(lineno)));
        path.pop_back();
        continue;
      }
      if (line.compare(0, 4, "set ") == 0) {
        const size_t eq = line.find('=');
        AssertThrow(eq != std::string::npos, ShimException("ParameterHandler: no `=` in line " + std::to_string(lineno)));
        const std::string name = Patterns::shim_trim(line.substr(4, eq - 4)), value = Patterns::shim_trim(line.substr(eq + 1));
        auto it = entries.find(key(name));
        AssertThrow(it != entries.end(), ShimException("ParameterHandler: undeclared entry <" + name + "> in line " + std::to_string(lineno)));
        AssertThrow(it->second.pattern->match(value), ShimException("ParameterHandler: value <" + value + "> of <" + name + "> violates its pattern"));
        it->second.value = value;
        continue;
      }
      AssertThrow(false, ShimException("ParameterHandler: cannot parse line " + std::to_string(lineno) + ": " + line));
    }
    AssertThrow(path.size() == saved.size(), ShimException("ParameterHandler: unclosed subsection"));
    return true;
  }
  std::ostream& print_parameters(std::ostream& out, OutputStyle) const {
    out << "# Listing of Parameters (deal.II API shim)\n";
    for (const auto& e : entries) out << "#   " << e.first << " = " << e.second.value << "\n";
    return out;
  }
  std::string get(const std::string& name) const {
    auto it = entries.find(key(name));
    AssertThrow(it != entries.end(), ShimException("ParameterHandler: undeclared entry <" + name + ">"));
    return it->second.value;
  }
  long get_integer(const std::string& name) const { return std::strtol(get(name).c_str(), nullptr, 10); }
  double get_double(const std::string& name) const { return std::strtod(get(name).c_str(), nullptr); }
 private:
  struct Entry { std::string value; std::shared_ptr<Patterns::PatternBase> pattern; };
  std::string key(const std::string& name) const {
    std::string k;
    for (const auto& p : path) k += p + "/";
    return k + name;
  }
  std::vector<std::string> path;
  std::map<std::string, Entry> entries;
};

}  // namespace dealii

#include "shim_fe.h"
#include "shim_lac.h"
#include "shim_numerics.h"
