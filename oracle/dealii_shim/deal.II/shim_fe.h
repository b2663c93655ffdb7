// shim_fe.h — Triangulation, FE_Q / FESystem, QGauss, DoFHandler, FEValues of the deal.II API shim (see shim.h: NOT deal.II).
#pragma once

namespace dealii {

template <int dim> struct GeometryInfo {
  static const unsigned int vertices_per_cell = 1u << dim;
  static const unsigned int faces_per_cell = 2 * dim;
  static const unsigned int lines_per_cell = dim == 2 ? 4 : (dim == 3 ? 12 : 1);
  static const unsigned int quads_per_cell = dim == 3 ? 6 : (dim == 2 ? 1 : 0);
};
// line -> its two vertices, 3D face -> its four vertices: geometry_info.h (lines 0-3 bound the bottom face, 4-7 the top face,
// 8-11 are vertical; faces are ordered x-, x+, y-, y+, z-, z+).  In 2D the four lines are the faces x-, x+, y-, y+.
static const int shim_line_vertices[12][2] = {{0, 2}, {1, 3}, {0, 1}, {2, 3}, {4, 6}, {5, 7}, {4, 5}, {6, 7}, {0, 4}, {1, 5}, {2, 6}, {3, 7}};
static const int shim_quad_vertices[6][4] = {{0, 2, 4, 6}, {1, 3, 5, 7}, {0, 1, 4, 5}, {2, 3, 6, 7}, {0, 1, 2, 3}, {4, 5, 6, 7}};

// ---------------------------------------------------------------------------------------------- Triangulation
// Only what hyper_rectangle + refine_global can produce: a box of 2^L cells per axis.  Active cells are stored in deal.II's
// order — the children of a cell are consecutive, child c sits at (c & 1, c >> 1 & 1, c >> 2 & 1) of its parent
// (GeometryInfo::child_cell_on_face / tria.cc, execute_refinement), so a uniformly refined mesh is traversed in Morton order.
template <int dim> class Triangulation {
 public:
  struct FaceRef {
    bool boundary; types::boundary_id id;
    bool at_boundary() const { return boundary; }
    types::boundary_id boundary_id() const { return id; }
    const FaceRef* operator->() const { return this; }
  };
  struct CellAccessor {
    const Triangulation* tria = nullptr;
    unsigned int index = 0;
    FaceRef face(unsigned int f) const { return tria->face_of(index, f); }
    Point<dim> vertex(unsigned int v) const { return tria->vertex_of(index, v); }
    unsigned int vertex_index(unsigned int v) const { return (unsigned int)tria->vertex_number(index, v); }
    void clear_refine_flag() const { shim_unsupported("clear_refine_flag"); }
    void clear_coarsen_flag() const { shim_unsupported("clear_coarsen_flag"); }
  };
  struct active_cell_iterator {
    CellAccessor a;
    const CellAccessor* operator->() const { return &a; }
    active_cell_iterator& operator++() { ++a.index; return *this; }
    bool operator!=(const active_cell_iterator& o) const { return a.index != o.a.index; }
    bool operator==(const active_cell_iterator& o) const { return a.index == o.a.index; }
  };
  Triangulation() {}
  void create_box(const Point<dim>& lower, const Point<dim>& upper, bool colorize_) { lo = lower; hi = upper; colorize = colorize_; level = 0; have_mesh = true; }
  void refine_global(unsigned int times) { level += times; }
  unsigned int n_active_cells() const { return have_mesh ? 1u << (dim * level) : 0u; }
  unsigned int n_levels() const { return level + 1; }
  unsigned int n_vertices() const { unsigned int n = 1; for (int a = 0; a < dim; ++a) n *= cells_per_axis() + 1; return n; }
  std::vector<Point<dim>> get_vertices() const {  // indexed by vertex_index(): lexicographic over the lattice
    std::vector<Point<dim>> out(n_vertices());
    const unsigned int n1 = cells_per_axis() + 1;
    for (unsigned int id = 0; id < out.size(); ++id) {
      unsigned int r = id;
      for (int a = 0; a < dim; ++a) { out[id][a] = lo[a] + (hi[a] - lo[a]) * ((r % n1) / (double)cells_per_axis()); r /= n1; }
    }
    return out;
  }
  unsigned int cells_per_axis() const { return 1u << level; }
  active_cell_iterator begin_active(unsigned int = 0) const { return active_cell_iterator{CellAccessor{this, 0}}; }
  active_cell_iterator end() const { return active_cell_iterator{CellAccessor{this, n_active_cells()}}; }
  active_cell_iterator end_active(unsigned int) const { return end(); }
  void prepare_coarsening_and_refinement() { shim_unsupported("prepare_coarsening_and_refinement"); }
  void execute_coarsening_and_refinement() { shim_unsupported("execute_coarsening_and_refinement"); }
  // lattice position of active cell c (Morton index -> per-axis cell index)
  void cell_position(unsigned int c, unsigned int pos[3]) const {
    pos[0] = pos[1] = pos[2] = 0;
    for (unsigned int l = 0; l < level; ++l)
      for (int a = 0; a < dim; ++a) pos[a] |= ((c >> (dim * l + a)) & 1u) << l;
  }
  // a unique number for the lattice vertex v of cell c (lexicographic over the (2^L + 1)^dim lattice; its value is irrelevant to
  // the dof numbering, which follows the traversal order)
  uint64_t vertex_number(unsigned int c, unsigned int v) const {
    unsigned int pos[3];
    cell_position(c, pos);
    const uint64_t n1 = cells_per_axis() + 1;
    uint64_t id = 0, stride = 1;
    for (int a = 0; a < dim; ++a) { id += (pos[a] + ((v >> a) & 1u)) * stride; stride *= n1; }
    return id;
  }
  Point<dim> vertex_of(unsigned int c, unsigned int v) const {
    unsigned int pos[3];
    cell_position(c, pos);
    const double n = cells_per_axis();
    Point<dim> p;
    for (int a = 0; a < dim; ++a) p[a] = lo[a] + (hi[a] - lo[a]) * ((pos[a] + ((v >> a) & 1u)) / n);
    return p;
  }
  FaceRef face_of(unsigned int c, unsigned int f) const {
    unsigned int pos[3];
    cell_position(c, pos);
    const unsigned int axis = f / 2, side = f % 2;
    const bool b = side == 0 ? pos[axis] == 0 : pos[axis] + 1 == cells_per_axis();
    // colorize: boundary id = face number of the coarse cell (grid_generator.cc, colorize_hyper_rectangle); otherwise 0
    return FaceRef{b, b ? (types::boundary_id)(colorize ? f : 0) : numbers::internal_face_boundary_id};
  }
 private:
  Point<dim> lo, hi;
  bool colorize = false, have_mesh = false;
  unsigned int level = 0;
};

namespace GridGenerator {
template <int dim> inline void hyper_cube(Triangulation<dim>& tria, const double left = 0., const double right = 1., const bool colorize = false) {
  Point<dim> p1, p2;
  for (int i = 0; i < dim; ++i) { p1[i] = left; p2[i] = right; }
  tria.create_box(p1, p2, colorize);
}
// grid_generator.cc, hyper_rectangle: the two corners need not be ordered (FSS:422-425 passes the upper one first)
template <int dim> inline void hyper_rectangle(Triangulation<dim>& tria, const Point<dim>& p_1, const Point<dim>& p_2, const bool colorize = false) {
  Point<dim> p1, p2;
  for (int i = 0; i < dim; ++i) { p1[i] = std::min(p_1[i], p_2[i]); p2[i] = std::max(p_1[i], p_2[i]); }
  tria.create_box(p1, p2, colorize);
}
}
template <int dim> class GridIn {
 public:
  void attach_triangulation(Triangulation<dim>&) {}
  void read_msh(std::istream&) { shim_unsupported("GridIn::read_msh"); }
};
namespace GridRefinement {
template <int dim, class V> inline void refine_and_coarsen_fixed_fraction(Triangulation<dim>&, const V&, double, double) {
  shim_unsupported("GridRefinement::refine_and_coarsen_fixed_fraction");
}
}

// ---------------------------------------------------------------------------------------------- finite elements
namespace FEValuesExtractors { struct Scalar { unsigned int component; explicit Scalar(unsigned int c) : component(c) {} }; }
struct ComponentMask {
  std::vector<bool> mask;  // empty = all components
  bool operator[](unsigned int c) const { return mask.empty() || mask[c]; }
};

// FE_Q(k)^n on the reference cell [0,1]^dim.  Cell-local dof order as in deal.II (fe_q_base.cc / fe_system.cc): the dofs of all
// vertices, then of all lines, quads, the hex interior; within one geometric object the components of an FESystem are consecutive.
// Nodes of FE_Q(2) are the vertices, line midpoints, face centres and the cell centre (equidistant support points).
template <int dim> class FiniteElement {
 public:
  unsigned int degree = 1, dofs_per_cell = 0, n_comp = 1;
  struct Node { int type; int object; double xi[3]; int lat[3]; };  // type 0 vertex, 1 line, 2 quad, 3 hex; lat = index into the 1D node set
  std::vector<Node> nodes;       // scalar nodes in deal.II's cell-local order
  unsigned int n_components() const { return n_comp; }
  std::pair<unsigned int, unsigned int> system_to_component_index(unsigned int i) const { return {i % n_comp, i / n_comp}; }
  ComponentMask component_mask(const FEValuesExtractors::Scalar& s) const {
    ComponentMask m;
    m.mask.assign(n_comp, false);
    m.mask[s.component] = true;
    return m;
  }
  // 1D Lagrange basis on the equidistant node set {0, 1} or {0, 1/2, 1}; index n in lattice order (0, [1/2], 1)
  double basis1d(int n, double x) const {
    if (degree == 1) return n == 0 ? 1 - x : x;
    return n == 0 ? 2 * (x - 0.5) * (x - 1) : (n == 1 ? 4 * x * (1 - x) : 2 * x * (x - 0.5));
  }
  double dbasis1d(int n, double x) const {
    if (degree == 1) return n == 0 ? -1.0 : 1.0;
    return n == 0 ? 4 * x - 3 : (n == 1 ? 4 - 8 * x : 4 * x - 1);
  }
  double shape(unsigned int scalar_node, const double* xi) const {
    double v = 1;
    for (int a = 0; a < dim; ++a) v *= basis1d(nodes[scalar_node].lat[a], xi[a]);
    return v;
  }
  void shape_grad(unsigned int scalar_node, const double* xi, double* g) const {
    for (int a = 0; a < dim; ++a) {
      double v = 1;
      for (int b = 0; b < dim; ++b) v *= (a == b) ? dbasis1d(nodes[scalar_node].lat[b], xi[b]) : basis1d(nodes[scalar_node].lat[b], xi[b]);
      g[a] = v;
    }
  }
 protected:
  void build(unsigned int degree_, unsigned int n_comp_) {
    AssertThrow(degree_ == 1 || degree_ == 2, ShimException("deal.II shim: FE_Q(1) and FE_Q(2) only"));
    degree = degree_;
    n_comp = n_comp_;
    nodes.clear();
    const int top = degree;  // lattice index of coordinate 1
    auto add = [&](int type, int object, const double* xi) {
      Node n;
      n.type = type; n.object = object;
      for (int a = 0; a < 3; ++a) { n.xi[a] = a < dim ? xi[a] : 0.0; n.lat[a] = a < dim ? (int)std::lround(xi[a] * top) : 0; }
      nodes.push_back(n);
    };
    const int nv = 1 << dim;
    double X[8][3];
    for (int v = 0; v < nv; ++v) {
      for (int a = 0; a < 3; ++a) X[v][a] = (v >> a) & 1;
      add(0, v, X[v]);
    }
    if (degree == 2) {
      const int nl = dim == 2 ? 4 : 12;
      for (int l = 0; l < nl; ++l) {
        double m[3];
        for (int a = 0; a < 3; ++a) m[a] = 0.5 * (X[shim_line_vertices[l][0]][a] + X[shim_line_vertices[l][1]][a]);
        add(1, l, m);
      }
      if (dim == 2) { double c[3] = {0.5, 0.5, 0}; add(2, 0, c); }
      if (dim == 3) {
        for (int q = 0; q < 6; ++q) {
          double m[3] = {0, 0, 0};
          for (int k = 0; k < 4; ++k) for (int a = 0; a < 3; ++a) m[a] += 0.25 * X[shim_quad_vertices[q][k]][a];
          add(2, q, m);
        }
        double c[3] = {0.5, 0.5, 0.5};
        add(3, 0, c);
      }
    }
    dofs_per_cell = (unsigned int)nodes.size() * n_comp;
  }
};
template <int dim> class FE_Q : public FiniteElement<dim> {
 public:
  explicit FE_Q(unsigned int degree) { this->build(degree, 1); }
};
// DEALII_SHIM_FESYSTEM_DEGREE=<k> overrides the degree of the base element of every FESystem at run time.  The reference
// hard-codes FE_Q(2) for the displacement (DS:67: `fe(FE_Q<dim>(2), dim)`, the constructor's fe_degree argument is unused);
// BASELINE.json's benchmark configurations use Q1/Q1.  With the override the reference's unmodified code runs those
// configurations too (records whose name starts with q1_); without it nothing changes.
template <int dim> class FESystem : public FiniteElement<dim> {
 public:
  FESystem(const FE_Q<dim>& base, unsigned int n) {
    unsigned int degree = base.degree;
    if (const char* e = std::getenv("DEALII_SHIM_FESYSTEM_DEGREE")) degree = (unsigned int)std::atoi(e);
    this->build(degree, n);
  }
};

// ---------------------------------------------------------------------------------------------- quadrature
inline void shim_gauss_1d(unsigned int n, std::vector<double>& x, std::vector<double>& w) {  // Gauss-Legendre on [0,1], ascending
  x.resize(n); w.resize(n);
  const double pi = 3.14159265358979323846;
  for (unsigned int i = 0; i < n; ++i) {
    double z = std::cos(pi * (i + 0.75) / (n + 0.5)), pp = 0;
    for (int it = 0; it < 100; ++it) {
      double p1 = 1, p2 = 0;
      for (unsigned int j = 1; j <= n; ++j) { const double p3 = p2; p2 = p1; p1 = ((2.0 * j - 1) * z * p2 - (j - 1.0) * p3) / j; }
      pp = n * (z * p1 - p2) / (z * z - 1);
      const double z1 = z;
      z = z1 - p1 / pp;
      if (std::fabs(z - z1) < 1e-16) break;
    }
    x[n - 1 - i] = 0.5 * (z + 1);
    w[n - 1 - i] = 1.0 / ((1 - z * z) * pp * pp);
  }
}
template <int dim> class Quadrature {
 public:
  unsigned int size() const { return (unsigned int)weights.size(); }
  std::vector<std::array<double, 3>> points;
  std::vector<double> weights;
};
template <int dim> class QGauss : public Quadrature<dim> {  // tensor product, first coordinate fastest (quadrature.cc)
 public:
  explicit QGauss(unsigned int n) {
    std::vector<double> x, w;
    shim_gauss_1d(n, x, w);
    unsigned int total = 1;
    for (int a = 0; a < dim; ++a) total *= n;
    for (unsigned int q = 0; q < total; ++q) {
      std::array<double, 3> p = {{0, 0, 0}};
      double wt = 1;
      unsigned int r = q;
      for (int a = 0; a < dim; ++a) { p[a] = x[r % n]; wt *= w[r % n]; r /= n; }
      this->points.push_back(p);
      this->weights.push_back(wt);
    }
  }
};

// ---------------------------------------------------------------------------------------------- DoFHandler
// distribute_dofs follows dof_handler_policy.cc / dof_handler.cc (Implementation::distribute_dofs_on_cell): active cells in
// traversal order; on every cell first the vertices, then the lines, then (3D) the quads, then the interior; an object that has
// no numbers yet gets dofs_per_object consecutive ones, all components of the FESystem together.
template <int dim> class DoFHandler {
 public:
  struct CellAccessor {
    const DoFHandler* dh = nullptr;
    unsigned int index = 0;
    typename Triangulation<dim>::FaceRef face(unsigned int f) const { return dh->tria->face_of(index, f); }
    Point<dim> vertex(unsigned int v) const { return dh->tria->vertex_of(index, v); }
    void get_dof_indices(std::vector<types::global_dof_index>& out) const {
      const unsigned int n = dh->dofs_per_cell;
      for (unsigned int i = 0; i < n; ++i) out[i] = dh->cell_dofs[(size_t)index * n + i];
    }
  };
  struct active_cell_iterator {
    CellAccessor a;
    const CellAccessor* operator->() const { return &a; }
    active_cell_iterator& operator++() { ++a.index; return *this; }
    bool operator!=(const active_cell_iterator& o) const { return a.index != o.a.index; }
  };
  explicit DoFHandler(const Triangulation<dim>& t) : tria(&t) {}
  void clear() { cell_dofs.clear(); total = 0; }
  unsigned int n_dofs() const { return total; }
  const Triangulation<dim>& get_tria() const { return *tria; }
  const FiniteElement<dim>& get_fe() const { return *fe; }
  active_cell_iterator begin_active() const { return active_cell_iterator{CellAccessor{this, 0}}; }
  active_cell_iterator end() const { return active_cell_iterator{CellAccessor{this, tria->n_active_cells()}}; }
  void distribute_dofs(const FiniteElement<dim>& fe_) {
    fe = &fe_;
    dofs_per_cell = fe_.dofs_per_cell;
    const unsigned int nc = tria->n_active_cells(), ncomp = fe_.n_comp;
    cell_dofs.assign((size_t)nc * dofs_per_cell, numbers::invalid_dof_index);
    std::map<std::array<uint64_t, 4>, types::global_dof_index> first_of;  // geometric object (its sorted vertex numbers) -> first dof
    total = 0;
    for (unsigned int c = 0; c < nc; ++c) {
      uint64_t vn[8];
      for (unsigned int v = 0; v < GeometryInfo<dim>::vertices_per_cell; ++v) vn[v] = tria->vertex_number(c, v);
      for (size_t s = 0; s < fe_.nodes.size(); ++s) {
        const auto& nd = fe_.nodes[s];
        const uint64_t none = ~(uint64_t)0;
        std::array<uint64_t, 4> key = {{none, none, none, none}};
        if (nd.type == 0) key[0] = vn[nd.object];
        else if (nd.type == 1) { key[0] = vn[shim_line_vertices[nd.object][0]]; key[1] = vn[shim_line_vertices[nd.object][1]]; }
        else if (nd.type == 2 && dim == 3) for (int k = 0; k < 4; ++k) key[k] = vn[shim_quad_vertices[nd.object][k]];
        types::global_dof_index first;
        if ((nd.type == 2 && dim == 2) || nd.type == 3) { first = total; total += ncomp; }  // cell interior: never shared
        else {
          std::sort(key.begin(), key.end());
          auto it = first_of.find(key);
          if (it == first_of.end()) { first = total; first_of[key] = first; total += ncomp; }
          else first = it->second;
        }
        for (unsigned int k = 0; k < ncomp; ++k) cell_dofs[(size_t)c * dofs_per_cell + s * ncomp + k] = first + k;
      }
    }
  }
  const Triangulation<dim>* tria;
  const FiniteElement<dim>* fe = nullptr;
  unsigned int dofs_per_cell = 0, total = 0;
  std::vector<types::global_dof_index> cell_dofs;
};

// ---------------------------------------------------------------------------------------------- FEValues
enum UpdateFlags { update_default = 0, update_values = 1, update_gradients = 2, update_quadrature_points = 4, update_JxW_values = 8, update_normal_vectors = 16 };
inline UpdateFlags operator|(UpdateFlags a, UpdateFlags b) { return (UpdateFlags)((int)a | (int)b); }

// MappingQ1 (mapping_q1.cc): x(xi) = sum_v X_v N_v(xi), J = dx/dxi, JxW = det(J) w_q, grad phi = J^-T grad_xi phi.
template <int dim> class FEValuesBase {
 public:
  FEValuesBase(const FiniteElement<dim>& fe) : fe(fe) {}
  double JxW(unsigned int q) const { return jxw[q]; }
  const std::vector<Point<dim>>& get_quadrature_points() const { return qpoints; }
  // primitive elements: the value of the only non-zero component of shape function i (fe_values.h, shape_value)
  double shape_value(unsigned int i, unsigned int q) const { return values[(size_t)(i / fe.n_comp) * nq + q]; }
  Tensor<1, dim> shape_grad_component(unsigned int i, unsigned int q, unsigned int component) const {
    Tensor<1, dim> g;
    if (i % fe.n_comp != component) return g;
    for (int a = 0; a < dim; ++a) g[a] = grads[((size_t)(i / fe.n_comp) * nq + q) * dim + a];
    return g;
  }
  const Tensor<1, dim>& normal_vector(unsigned int q) const { return normals[q]; }
  Tensor<1, dim> shape_grad(unsigned int i, unsigned int q) const { return shape_grad_component(i, q, i % fe.n_comp); }
  const Point<dim>& quadrature_point(unsigned int q) const { return qpoints[q]; }
  // fe_values.cc, do_function_values: values[q] += dof_value(i) * shape_value(i, q), shape functions in the outer loop
  template <class V> void get_function_values(const V& u, std::vector<double>& out) const {
    std::fill(out.begin(), out.end(), 0.0);
    for (unsigned int i = 0; i < fe.dofs_per_cell; ++i) {
      const double ui = u(dofs[i]);
      for (unsigned int q = 0; q < nq; ++q) out[q] += ui * shape_value(i, q);
    }
  }
  // vector-valued: out[q][component] = gradient of that component of the field
  template <class V> void get_function_gradients(const V& u, std::vector<std::vector<Tensor<1, dim>>>& out) const {
    for (auto& per_q : out) for (auto& t : per_q) t = Tensor<1, dim>();
    for (unsigned int i = 0; i < fe.dofs_per_cell; ++i) {
      const double ui = u(dofs[i]);
      const unsigned int comp = i % fe.n_comp;
      for (unsigned int q = 0; q < nq; ++q)
        for (int a = 0; a < dim; ++a) out[q][comp][a] += ui * grads[((size_t)(i / fe.n_comp) * nq + q) * dim + a];
    }
  }
 protected:
  // evaluates everything at the reference points xi[q] (weights w[q]); face >= 0: surface element and outward normal of that face
  template <class Cell> void evaluate(const Cell& cell, const std::vector<std::array<double, 3>>& xi, const std::vector<double>& w, int face) {
    nq = (unsigned int)xi.size();
    const unsigned int ns = (unsigned int)fe.nodes.size(), nv = 1u << dim;
    values.assign((size_t)ns * nq, 0.0);
    grads.assign((size_t)ns * nq * dim, 0.0);
    jxw.assign(nq, 0.0);
    qpoints.assign(nq, Point<dim>());
    normals.assign(nq, Tensor<1, dim>());
    Point<dim> X[8];
    for (unsigned int v = 0; v < nv; ++v) X[v] = cell->vertex(v);
    dofs.resize(fe.dofs_per_cell);
    cell->get_dof_indices(dofs);
    for (unsigned int q = 0; q < nq; ++q) {
      double J[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
      for (unsigned int v = 0; v < nv; ++v) {
        double N = 1, dN[3];
        for (int a = 0; a < dim; ++a) N *= ((v >> a) & 1) ? xi[q][a] : 1 - xi[q][a];
        for (int a = 0; a < dim; ++a) {
          double g = ((v >> a) & 1) ? 1.0 : -1.0;
          for (int b = 0; b < dim; ++b) if (b != a) g *= ((v >> b) & 1) ? xi[q][b] : 1 - xi[q][b];
          dN[a] = g;
        }
        for (int a = 0; a < dim; ++a) {
          qpoints[q][a] += X[v][a] * N;
          for (int b = 0; b < dim; ++b) J[a][b] += X[v][a] * dN[b];
        }
      }
      double Ji[3][3], det;
      if (dim == 2) {
        det = J[0][0] * J[1][1] - J[0][1] * J[1][0];
        Ji[0][0] = J[1][1] / det; Ji[0][1] = -J[0][1] / det; Ji[1][0] = -J[1][0] / det; Ji[1][1] = J[0][0] / det;
      } else {
        const double c00 = J[1][1] * J[2][2] - J[1][2] * J[2][1], c01 = J[1][2] * J[2][0] - J[1][0] * J[2][2], c02 = J[1][0] * J[2][1] - J[1][1] * J[2][0];
        det = J[0][0] * c00 + J[0][1] * c01 + J[0][2] * c02;
        Ji[0][0] = c00 / det; Ji[1][0] = c01 / det; Ji[2][0] = c02 / det;
        Ji[0][1] = (J[0][2] * J[2][1] - J[0][1] * J[2][2]) / det; Ji[1][1] = (J[0][0] * J[2][2] - J[0][2] * J[2][0]) / det; Ji[2][1] = (J[0][1] * J[2][0] - J[0][0] * J[2][1]) / det;
        Ji[0][2] = (J[0][1] * J[1][2] - J[0][2] * J[1][1]) / det; Ji[1][2] = (J[0][2] * J[1][0] - J[0][0] * J[1][2]) / det; Ji[2][2] = (J[0][0] * J[1][1] - J[0][1] * J[1][0]) / det;
      }
      if (face < 0) jxw[q] = det * w[q];
      else {  // n dS = det(J) J^-T n_ref dS_ref
        const int axis = face / 2;
        const double sign = face % 2 ? 1.0 : -1.0;
        double nrm = 0;
        for (int a = 0; a < dim; ++a) { normals[q][a] = sign * Ji[axis][a]; nrm += normals[q][a] * normals[q][a]; }
        nrm = std::sqrt(nrm);
        for (int a = 0; a < dim; ++a) normals[q][a] /= nrm;
        jxw[q] = std::fabs(det) * nrm * w[q];
      }
      for (unsigned int s = 0; s < ns; ++s) {
        double gref[3];
        values[(size_t)s * nq + q] = fe.shape(s, xi[q].data());
        fe.shape_grad(s, xi[q].data(), gref);
        for (int a = 0; a < dim; ++a) {
          double g = 0;
          for (int b = 0; b < dim; ++b) g += Ji[b][a] * gref[b];  // (J^-T grad_xi)_a = sum_b (J^-1)_ba d/dxi_b
          grads[((size_t)s * nq + q) * dim + a] = g;
        }
      }
    }
  }
  const FiniteElement<dim>& fe;
  unsigned int nq = 0;
  std::vector<double> values, grads, jxw;
  std::vector<Point<dim>> qpoints;
  std::vector<Tensor<1, dim>> normals;
  std::vector<types::global_dof_index> dofs;
};
template <int dim> class FEValues : public FEValuesBase<dim> {
 public:
  FEValues(const FiniteElement<dim>& fe, const Quadrature<dim>& q, UpdateFlags) : FEValuesBase<dim>(fe), quad(q) {}
  template <class Iterator> void reinit(const Iterator& cell) { this->evaluate(cell, quad.points, quad.weights, -1); }
 private:
  Quadrature<dim> quad;
};
template <int dim> class FEFaceValues : public FEValuesBase<dim> {
 public:
  FEFaceValues(const FiniteElement<dim>& fe, const Quadrature<dim - 1>& q, UpdateFlags) : FEValuesBase<dim>(fe), quad(q) {}
  // the face's quadrature points: the remaining axes in ascending order carry the (dim-1)-dimensional formula
  template <class Iterator> void reinit(const Iterator& cell, unsigned int f) {
    std::vector<std::array<double, 3>> xi(quad.size());
    const unsigned int axis = f / 2, side = f % 2;
    for (unsigned int q = 0; q < quad.size(); ++q) {
      int t = 0;
      for (int a = 0; a < dim; ++a) xi[q][a] = (a == (int)axis) ? (double)side : quad.points[q][t++];
    }
    this->evaluate(cell, xi, quad.weights, (int)f);
  }
 private:
  Quadrature<dim - 1> quad;
};

}  // namespace dealii
