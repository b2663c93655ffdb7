// dofs.hpp — DoF numbering, support points and Dirichlet constraint tables without deal.II.
//
// Restates what the reference obtains from DoFHandler::distribute_dofs (PS:73, DS:110) for
// FE_Q(k) / FESystem(FE_Q(k), dim), k in {1,2}: active cells are visited in order and indices
// are handed out first-touch — the cell's vertices, then its lines, then (3D) its quads, then
// the cell interior; each entity carries n_comp consecutive numbers.  No renumbering is
// applied (the reference never calls DoFRenumbering).  Local dof (scalar s, component c) has
// cell-local index s*n_comp + c (FESystem of identical Lagrange bases, vertex/line/quad/hex
// blocks each with one scalar dof).
// Dirichlet rows follow DS:117-135: VectorTools::interpolate_boundary_values per
// (label, component, value) triple in list order; a dof constrained earlier is never
// overwritten.
#pragma once
#include <algorithm>
#include <cstdint>
#include <stdexcept>
#include <unordered_map>
#include <vector>

#include "mesh.hpp"

namespace dofs {

// Reference-cell description of FE_Q(k), k<=2, in deal.II's local ordering.
struct RefElement {
  int dim = 0, degree = 0, n_scalar = 0;
  std::vector<double> unit_support;  // n_scalar * dim
  // entity -> reference vertices (for global identification)
  std::vector<std::array<int, 2>> lines;
  std::vector<std::array<int, 4>> quads;
};

inline RefElement make_ref_element(int dim, int degree) {
  RefElement r;
  r.dim = dim;
  r.degree = degree;
  if (dim == 2)
    r.lines = {{0, 2}, {1, 3}, {0, 1}, {2, 3}};
  else {
    r.lines = {{0, 2}, {1, 3}, {0, 1}, {2, 3}, {4, 6}, {5, 7}, {4, 5}, {6, 7}, {0, 4}, {1, 5}, {2, 6}, {3, 7}};
    r.quads = {{0, 2, 4, 6}, {1, 3, 5, 7}, {0, 1, 4, 5}, {2, 3, 6, 7}, {0, 1, 2, 3}, {4, 5, 6, 7}};
  }
  auto vcoord = [&](int v, int a) { return (double)((v >> a) & 1); };
  int nv = 1 << dim;
  for (int v = 0; v < nv; ++v)
    for (int a = 0; a < dim; ++a) r.unit_support.push_back(vcoord(v, a));
  if (degree == 2) {
    for (auto& l : r.lines)
      for (int a = 0; a < dim; ++a) r.unit_support.push_back(0.5 * (vcoord(l[0], a) + vcoord(l[1], a)));
    for (auto& q : r.quads)
      for (int a = 0; a < dim; ++a)
        r.unit_support.push_back(0.25 * (vcoord(q[0], a) + vcoord(q[1], a) + vcoord(q[2], a) + vcoord(q[3], a)));
    for (int a = 0; a < dim; ++a) r.unit_support.push_back(0.5);
  }
  r.n_scalar = (int)r.unit_support.size() / dim;
  return r;
}

struct DofMap {
  int degree = 1, n_comp = 1, n_loc = 0;
  int64_t n_dofs = 0;
  std::vector<int32_t> cell_dofs;  // n_cells * n_loc
};

struct PairHash {
  size_t operator()(const std::pair<int64_t, int64_t>& p) const {
    return std::hash<int64_t>()(p.first * 1000003LL ^ (p.second + 0x9e3779b97f4a7c15LL));
  }
};
typedef std::pair<int64_t, int64_t> EntityKey;

// Open-addressing hash map for entity keys (lines, quads): no per-node allocation, linear probing, never erased.  The
// refinement forest does tens of millions of look-ups per pass (amr.hpp); std::unordered_map made those 5-10x slower.
// Interface: the subset of std::unordered_map the callers use (find / end / operator[] / at / size / swap).
template <class V>
class FlatMap {
 public:
  struct Slot { EntityKey first; V second; };
  typedef Slot* iterator;
  typedef const Slot* const_iterator;
  FlatMap() { rehash(16); }
  size_t size() const { return count_; }
  iterator end() { return nullptr; }
  const_iterator end() const { return nullptr; }
  const_iterator find(const EntityKey& k) const {
    size_t i = hash(k) & mask_;
    while (used_[i]) {
      if (slots_[i].first == k) return &slots_[i];
      i = (i + 1) & mask_;
    }
    return nullptr;
  }
  iterator find(const EntityKey& k) { return const_cast<iterator>(static_cast<const FlatMap*>(this)->find(k)); }
  V& operator[](const EntityKey& k) {
    if ((count_ + 1) * 2 > slots_.size()) rehash(slots_.size() * 2);
    size_t i = hash(k) & mask_;
    while (used_[i]) {
      if (slots_[i].first == k) return slots_[i].second;
      i = (i + 1) & mask_;
    }
    used_[i] = 1;
    slots_[i].first = k;
    slots_[i].second = V();
    ++count_;
    return slots_[i].second;
  }
  const V& at(const EntityKey& k) const {
    const_iterator it = find(k);
    if (!it) throw std::out_of_range("FlatMap::at");
    return it->second;
  }
  void reserve(size_t n) {
    size_t cap = 16;
    while (cap < 2 * n) cap <<= 1;
    if (cap > slots_.size()) rehash(cap);
  }
  void swap(FlatMap& o) {
    slots_.swap(o.slots_);
    used_.swap(o.used_);
    std::swap(mask_, o.mask_);
    std::swap(count_, o.count_);
  }
  template <class F>
  void for_each(F&& f) const {
    for (size_t i = 0; i < slots_.size(); ++i)
      if (used_[i]) f(slots_[i].first, slots_[i].second);
  }

 private:
  static size_t hash(const EntityKey& k) {
    uint64_t h = (uint64_t)k.first * 0x9e3779b97f4a7c15ull ^ ((uint64_t)k.second + 0x7f4a7c159e3779b9ull) * 0xc2b2ae3d27d4eb4full;
    h ^= h >> 32;
    return (size_t)h;
  }
  void rehash(size_t cap) {
    std::vector<Slot> old;
    std::vector<uint8_t> old_used;
    old.swap(slots_);
    old_used.swap(used_);
    slots_.resize(cap);
    used_.assign(cap, 0);
    mask_ = cap - 1;
    count_ = 0;
    for (size_t i = 0; i < old.size(); ++i)
      if (old_used[i]) (*this)[old[i].first] = old[i].second;
  }
  std::vector<Slot> slots_;
  std::vector<uint8_t> used_;
  size_t mask_ = 0, count_ = 0;
};
typedef FlatMap<int32_t> EntityMap;

// A line is identified by its two vertices, a quad by its four (orientation-free keys).
inline EntityKey edge_key(int64_t a, int64_t b) { return a < b ? EntityKey{a, b} : EntityKey{b, a}; }
inline EntityKey quad_key(int64_t a, int64_t b, int64_t c, int64_t d) {
  int64_t vs[4] = {a, b, c, d};
  std::sort(vs, vs + 4);
  return {(vs[0] << 31) | vs[1], (vs[2] << 31) | vs[3]};
}

// first dof (component 0) of every vertex / line / quad node; used by the hanging-node constraints
struct NodeMaps {
  std::vector<int32_t> vdof;
  EntityMap line_dof, quad_dof;
};

inline DofMap distribute_dofs(const mesh::Mesh& m, int degree, int n_comp, NodeMaps* maps = nullptr) {
  RefElement ref = make_ref_element(m.dim, degree);
  DofMap d;
  d.degree = degree;
  d.n_comp = n_comp;
  d.n_loc = ref.n_scalar * n_comp;
  const int vpc = m.vpc();
  const int64_t nc = m.n_cells();
  d.cell_dofs.resize(nc * d.n_loc);
  std::vector<int32_t> vdof(m.n_vertices(), -1);
  EntityMap line_dof, quad_dof;
  int64_t next = 0;
  for (int64_t c = 0; c < nc; ++c) {
    const int32_t* cv = &m.cell_vertices[c * vpc];
    int32_t* out = &d.cell_dofs[c * d.n_loc];
    int s = 0;
    for (int v = 0; v < vpc; ++v, ++s) {
      if (vdof[cv[v]] < 0) { vdof[cv[v]] = (int32_t)next; next += n_comp; }
      for (int k = 0; k < n_comp; ++k) out[s * n_comp + k] = vdof[cv[v]] + k;
    }
    if (degree == 2) {
      for (auto& l : ref.lines) {
        const EntityKey key = edge_key(cv[l[0]], cv[l[1]]);
        auto it = line_dof.find(key);
        int32_t base;
        if (it == line_dof.end()) { base = (int32_t)next; line_dof[key] = base; next += n_comp; }
        else base = it->second;
        for (int k = 0; k < n_comp; ++k) out[s * n_comp + k] = base + k;
        ++s;
      }
      for (auto& q : ref.quads) {
        const EntityKey key = quad_key(cv[q[0]], cv[q[1]], cv[q[2]], cv[q[3]]);
        auto it = quad_dof.find(key);
        int32_t base;
        if (it == quad_dof.end()) { base = (int32_t)next; quad_dof[key] = base; next += n_comp; }
        else base = it->second;
        for (int k = 0; k < n_comp; ++k) out[s * n_comp + k] = base + k;
        ++s;
      }
      for (int k = 0; k < n_comp; ++k) out[s * n_comp + k] = (int32_t)next + k;
      next += n_comp;
      ++s;
    }
  }
  d.n_dofs = next;
  if (maps) {
    maps->vdof.swap(vdof);
    maps->line_dof.swap(line_dof);
    maps->quad_dof.swap(quad_dof);
  }
  return d;
}

// physical support point of every dof (Q1 mapping of the unit support points); n_dofs * dim
inline std::vector<double> support_points(const mesh::Mesh& m, const DofMap& d) {
  RefElement ref = make_ref_element(m.dim, d.degree);
  const int dim = m.dim, vpc = m.vpc();
  std::vector<double> sp(d.n_dofs * dim, 0.0);
  for (int64_t c = 0; c < m.n_cells(); ++c)
    for (int s = 0; s < ref.n_scalar; ++s) {
      double x[3] = {0, 0, 0};
      for (int v = 0; v < vpc; ++v) {
        double N = 1;
        for (int a = 0; a < dim; ++a) {
          double xi = ref.unit_support[s * dim + a];
          N *= ((v >> a) & 1) ? xi : 1 - xi;
        }
        for (int a = 0; a < dim; ++a) x[a] += N * m.xyz[(int64_t)m.cell_vertices[c * vpc + v] * dim + a];
      }
      for (int k = 0; k < d.n_comp; ++k) {
        int32_t g = d.cell_dofs[c * d.n_loc + s * d.n_comp + k];
        for (int a = 0; a < dim; ++a) sp[(int64_t)g * dim + a] = x[a];
      }
    }
  return sp;
}

// General linear constraints x_i = sum_j w_ij x_j + g_i (deal.II ConstraintMatrix as used at PS:71-78,
// DS:109-137): hanging-node lines carry entries, Dirichlet lines only an inhomogeneity.
struct ConstraintLine {
  int32_t dof = -1;
  std::vector<std::pair<int32_t, double>> entries;
  double inhomogeneity = 0;
};
struct ConstraintTable {
  std::vector<ConstraintLine> lines;
  std::vector<int32_t> line_of;  // dof -> index into lines, or -1
  void init(int64_t n_dofs) { lines.clear(); line_of.assign(n_dofs, -1); }
  bool is_constrained(int32_t dof) const { return line_of[dof] >= 0; }
  // like add_line + add_entries + set_inhomogeneity; an existing line is never overwritten
  bool add_line(int32_t dof, const std::vector<std::pair<int32_t, double>>& entries, double g) {
    if (line_of[dof] >= 0) return false;
    line_of[dof] = (int32_t)lines.size();
    ConstraintLine l;
    l.dof = dof;
    l.entries = entries;
    l.inhomogeneity = g;
    lines.push_back(std::move(l));
    return true;
  }
  // ConstraintMatrix::close(): entries that refer to constrained dofs are replaced by those dofs' own lines
  // (chains of hanging nodes; hanging nodes whose parents carry Dirichlet values), duplicates are merged,
  // entries and lines are sorted by dof.
  void close() {
    bool again = true;
    int sweeps = 0;
    while (again) {
      again = false;
      if (++sweeps > 64) throw std::runtime_error("ConstraintTable::close: cyclic constraints");
      for (auto& l : lines) {
        std::vector<std::pair<int32_t, double>> out;
        bool touched = false;
        for (auto& e : l.entries) {
          const int32_t lj = line_of[e.first];
          if (lj < 0) { out.push_back(e); continue; }
          touched = true;
          const ConstraintLine& o = lines[lj];
          for (auto& oe : o.entries) out.push_back({oe.first, e.second * oe.second});
          l.inhomogeneity += e.second * o.inhomogeneity;
        }
        if (touched) { l.entries.swap(out); again = true; }
      }
    }
    for (auto& l : lines) {
      std::sort(l.entries.begin(), l.entries.end());
      std::vector<std::pair<int32_t, double>> merged;
      for (auto& e : l.entries) {
        if (!merged.empty() && merged.back().first == e.first) merged.back().second += e.second;
        else merged.push_back(e);
      }
      l.entries.swap(merged);
    }
    std::sort(lines.begin(), lines.end(), [](const ConstraintLine& a, const ConstraintLine& b) { return a.dof < b.dof; });
    for (size_t i = 0; i < lines.size(); ++i) line_of[lines[i].dof] = (int32_t)i;
  }
  int64_t n_entries() const {
    int64_t n = 0;
    for (auto& l : lines) n += (int64_t)l.entries.size();
    return n;
  }
  // flat form for pe_upload_constraints
  void flatten(std::vector<int32_t>& line_dof, std::vector<int64_t>& entry_ptr, std::vector<int32_t>& entry_dof, std::vector<double>& entry_w,
               std::vector<double>& inhomogeneity) const {
    line_dof.clear(); entry_ptr.assign(1, 0); entry_dof.clear(); entry_w.clear(); inhomogeneity.clear();
    for (auto& l : lines) {
      line_dof.push_back(l.dof);
      inhomogeneity.push_back(l.inhomogeneity);
      for (auto& e : l.entries) { entry_dof.push_back(e.first); entry_w.push_back(e.second); }
      entry_ptr.push_back((int64_t)entry_dof.size());
    }
  }
};

struct Constraints {  // pure-Dirichlet lines x_i = g_i, sorted by dof (ConstraintMatrix::close)
  std::vector<int32_t> line_dof;
  std::vector<double> inhomogeneity;
};

// DS:117-135 on top of existing (hanging-node) lines: VectorTools::interpolate_boundary_values never overwrites a
// dof that is already constrained.  Call ConstraintTable::close() afterwards (DS:136).
inline void add_dirichlet(ConstraintTable& T, const mesh::Mesh& m, const DofMap& d, const std::vector<int>& labels, const std::vector<int>& comps,
                          const std::vector<double>& values) {
  RefElement ref = make_ref_element(m.dim, d.degree);
  for (size_t cond = 0; cond < labels.size(); ++cond)
    for (int64_t b = 0; b < m.n_bfaces(); ++b) {
      if (m.bface_id[b] != labels[cond]) continue;
      int f = m.bface_local[b], axis = f / 2;
      double side = f % 2;
      for (int s = 0; s < ref.n_scalar; ++s) {
        if (ref.unit_support[s * m.dim + axis] != side) continue;
        int32_t dof = d.cell_dofs[(int64_t)m.bface_cell[b] * d.n_loc + s * d.n_comp + comps[cond]];
        T.add_line(dof, {}, values[cond]);
      }
    }
}

// DS:117-135
inline Constraints make_dirichlet(const mesh::Mesh& m, const DofMap& d, const std::vector<int>& labels,
                                  const std::vector<int>& comps, const std::vector<double>& values) {
  RefElement ref = make_ref_element(m.dim, d.degree);
  std::vector<uint8_t> is_c(d.n_dofs, 0);
  std::vector<double> g(d.n_dofs, 0.0);
  for (size_t cond = 0; cond < labels.size(); ++cond) {
    for (int64_t b = 0; b < m.n_bfaces(); ++b) {
      if (m.bface_id[b] != labels[cond]) continue;
      int f = m.bface_local[b], axis = f / 2;
      double side = f % 2;
      for (int s = 0; s < ref.n_scalar; ++s) {
        if (ref.unit_support[s * m.dim + axis] != side) continue;
        int32_t dof = d.cell_dofs[(int64_t)m.bface_cell[b] * d.n_loc + s * d.n_comp + comps[cond]];
        if (!is_c[dof]) { is_c[dof] = 1; g[dof] = values[cond]; }
      }
    }
  }
  Constraints c;
  for (int64_t i = 0; i < d.n_dofs; ++i)
    if (is_c[i]) { c.line_dof.push_back((int32_t)i); c.inhomogeneity.push_back(g[i]); }
  return c;
}

}  // namespace dofs
