// amr.hpp — adaptive refinement of the time loop (PoroelasticityFSS.h:333-340, 447-498) without deal.II.
//
// What the reference gets from deal.II and what stands in for it here:
//   Triangulation<dim> + execute_coarsening_and_refinement ... Forest: a quadtree/octree forest over the cells of the
//       initial mesh (isotropic refinement, children in lexicographic order, boundary ids inherited).  Vertices,
//       lines and quads are identified by their vertex numbers, so the same code serves the colorized box
//       (FSS:418-435) and Gmsh meshes (FSS:438-445) whatever the relative orientation of neighbouring cells.
//   prepare_coarsening_and_refinement (no mesh smoothing: `Triangulation<dim> triangulation;`, FSS:75) ... Forest::prepare:
//       a family is coarsened only if all of its children are active and flagged; after refinement and coarsening two
//       cells that share a line (2D: a face; 3D: a face or an edge) differ by at most one level; refinement wins over
//       coarsening.  The closure is the smallest flag set with these properties, hence unique.
//   DoFTools::make_hanging_node_constraints (PS:74-75, DS:112-113) ... hanging_node_constraints() for FE_Q(1|2)^n_comp.
//   KellyErrorEstimator<dim>::estimate (FSS:454-458) ... kelly_estimate(): eta_K^2 = h_K/24 sum_F int_F [dp/dn]^2,
//       QGauss<dim-1>(2), no Neumann function map (boundary faces contribute nothing), result stored as float.
//   GridRefinement::refine_and_coarsen_fixed_fraction(0.6, 0.4) + the level limits of FSS:463-472 ... mark_fixed_fraction().
//   SolutionTransfer<dim>::interpolate for the FE_Q(1) pressure handler (FSS:475-497) ... vertex-keyed transfer: a dof that
//       exists before and after keeps its value, a vertex created by refinement gets the parent's Q1 interpolant.
// Stated deviations: active cells are ordered by (level, creation index) — deal.II re-uses the storage of coarsened
// cells, so the order inside a level can differ after coarsening (results are compared by coordinates); new vertices of
// curved/distorted parents are placed at the mean of the parent entity's vertices.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <numeric>
#include <stdexcept>
#include <vector>

#include "dofs.hpp"
#include "mesh.hpp"

namespace amr {

struct Cell {
  int32_t level = 0, parent = -1, child0 = -1, center_v = -1;
  int8_t child_index = 0;
  bool active = true, dead = false, refine_flag = false, coarsen_flag = false;
  int32_t v[8] = {-1, -1, -1, -1, -1, -1, -1, -1};
  int32_t bid[6] = {-1, -1, -1, -1, -1, -1};  // boundary id per face, -1 = interior
};

struct Forest {
  int dim = 0;
  std::vector<double> xyz;  // every vertex ever created (unused ones are harmless)
  std::vector<Cell> cells;
  dofs::EntityMap edge_mid, face_mid;  // line / quad (3D) -> midpoint vertex
  // vertex-keyed payload of the solution transfer (FSS:475-497)
  static constexpr int MAX_TRANSFER = 8;
  int n_transfer = 0;
  std::vector<double> vval;     // n_vertices * n_transfer
  std::vector<uint8_t> vknown;  // vertex carried a dof of the old mesh

  int vpc() const { return 1 << dim; }
  int64_t n_vertices() const { return (int64_t)xyz.size() / dim; }

  static Forest from_mesh(const mesh::Mesh& m, int base_level) {
    Forest F;
    F.dim = m.dim;
    F.xyz = m.xyz;
    const int vpc = m.vpc();
    F.cells.resize(m.n_cells());
    for (int64_t c = 0; c < m.n_cells(); ++c) {
      F.cells[c].level = base_level;
      for (int k = 0; k < vpc; ++k) F.cells[c].v[k] = m.cell_vertices[c * vpc + k];
    }
    for (int64_t b = 0; b < m.n_bfaces(); ++b) F.cells[m.bface_cell[b]].bid[m.bface_local[b]] = m.bface_id[b];
    return F;
  }

  // active cells in deal.II's iteration order: level by level
  std::vector<int32_t> active_cells() const {  // a counting sort by level: stable, so creation order inside a level
    std::vector<int64_t> start(2, 0);
    for (const Cell& c : cells)
      if (c.active && !c.dead) {
        if ((size_t)c.level + 2 > start.size()) start.resize((size_t)c.level + 2, 0);
        start[c.level + 1]++;
      }
    for (size_t l = 1; l < start.size(); ++l) start[l] += start[l - 1];
    std::vector<int32_t> a((size_t)start.back());
    for (size_t i = 0; i < cells.size(); ++i)
      if (cells[i].active && !cells[i].dead) a[start[cells[i].level]++] = (int32_t)i;
    return a;
  }
  int n_levels() const {
    int l = 0;
    for (auto& c : cells)
      if (c.active && !c.dead) l = std::max(l, c.level + 1);
    return l;
  }

  mesh::Mesh active_mesh(std::vector<int32_t>* ids = nullptr) const {
    mesh::Mesh m;
    m.dim = dim;
    m.xyz = xyz;
    std::vector<int32_t> a = active_cells();
    const int nv = vpc();
    m.cell_vertices.resize(a.size() * nv);
    for (size_t i = 0; i < a.size(); ++i) {
      const Cell& c = cells[a[i]];
      for (int k = 0; k < nv; ++k) m.cell_vertices[i * nv + k] = c.v[k];
      for (int f = 0; f < 2 * dim; ++f)
        if (c.bid[f] >= 0) {
          m.bface_cell.push_back((int32_t)i);
          m.bface_local.push_back((int8_t)f);
          m.bface_id.push_back(c.bid[f]);
        }
    }
    if (ids) *ids = a;
    return m;
  }

  int32_t new_vertex(const int32_t* corners, int n) {
    const int32_t id = (int32_t)n_vertices();
    for (int a = 0; a < dim; ++a) {
      double s = 0;
      for (int k = 0; k < n; ++k) s += xyz[(int64_t)corners[k] * dim + a];
      xyz.push_back(s / n);
    }
    if (n_transfer) {
      vval.resize((size_t)(id + 1) * n_transfer, 0.0);
      vknown.resize(id + 1, 0);
    }
    return id;
  }

  // vertex of the 3^dim refinement lattice of cell c; idx[a] in {0,1,2}
  int32_t lattice_vertex(int32_t c, const int idx[3]) {
    int32_t corners[8];
    int n = 0;
    for (int k = 0; k < vpc(); ++k) {
      bool ok = true;
      for (int a = 0; a < dim; ++a) {
        const int bit = (k >> a) & 1;
        if ((idx[a] == 0 && bit) || (idx[a] == 2 && !bit)) ok = false;
      }
      if (ok) corners[n++] = cells[c].v[k];
    }
    if (n == 1) return corners[0];
    if (n == 2) {
      auto key = dofs::edge_key(corners[0], corners[1]);
      auto it = edge_mid.find(key);
      if (it != edge_mid.end()) return it->second;
      int32_t id = new_vertex(corners, 2);
      edge_mid[key] = id;
      return id;
    }
    if (n == 4 && dim == 3) {
      auto key = dofs::quad_key(corners[0], corners[1], corners[2], corners[3]);
      auto it = face_mid.find(key);
      if (it != face_mid.end()) return it->second;
      int32_t id = new_vertex(corners, 4);
      face_mid[key] = id;
      return id;
    }
    if (cells[c].center_v < 0) {
      const int32_t id = new_vertex(corners, n);
      cells[c].center_v = id;
    }
    return cells[c].center_v;
  }

  void refine_cell(int32_t c) {
    if (!cells[c].active || cells[c].dead) throw std::runtime_error("amr: refine of a non-active cell");
    const int nv = vpc();
    const int32_t first = (int32_t)cells.size();
    // every vertex of the 3^dim lattice is looked up (line / quad maps) or created once, at its first use by a child — the
    // creation order of new vertices is that of the children's corners, as it always was
    const int n_lat = dim == 2 ? 9 : 27;
    int32_t lat[27];
    for (int i = 0; i < n_lat; ++i) lat[i] = -1;
    Cell kids[8];
    for (int ci = 0; ci < nv; ++ci) {
      Cell& k = kids[ci];
      k.level = cells[c].level + 1;
      k.parent = c;
      k.child_index = (int8_t)ci;
      for (int v = 0; v < nv; ++v) {
        int idx[3] = {0, 0, 0};
        for (int a = 0; a < dim; ++a) idx[a] = ((ci >> a) & 1) + ((v >> a) & 1);
        int32_t& lv = lat[idx[0] + 3 * idx[1] + 9 * idx[2]];
        if (lv < 0) lv = lattice_vertex(c, idx);
        k.v[v] = lv;
      }
      for (int f = 0; f < 2 * dim; ++f) k.bid[f] = (((ci >> (f / 2)) & 1) == (f % 2)) ? cells[c].bid[f] : -1;
    }
    // solution transfer: vertices that carried no dof get the parent's Q1 interpolant (mean of the parent entity)
    if (n_transfer)
      for (int i = 0; i < n_lat; ++i) {
        int idx[3] = {i % 3, (i / 3) % 3, dim == 3 ? i / 9 : 0};
        const int32_t v = lat[i];
        if (vknown[v]) continue;
        int n = 0;
        double acc[MAX_TRANSFER] = {0};
        for (int k = 0; k < nv; ++k) {
          bool ok = true;
          for (int a = 0; a < dim; ++a) {
            const int bit = (k >> a) & 1;
            if ((idx[a] == 0 && bit) || (idx[a] == 2 && !bit)) ok = false;
          }
          if (!ok) continue;
          ++n;
          for (int t = 0; t < n_transfer; ++t) acc[t] += vval[(size_t)cells[c].v[k] * n_transfer + t];
        }
        for (int t = 0; t < n_transfer; ++t) vval[(size_t)v * n_transfer + t] = acc[t] / n;
        vknown[v] = 2;  // interpolated in this pass: may serve as a source for nothing else (one level per pass)
      }
    cells[c].active = false;
    cells[c].child0 = first;
    for (int ci = 0; ci < nv; ++ci) cells.push_back(kids[ci]);
  }

  void coarsen_cell(int32_t p) {
    const int nv = vpc();
    for (int k = 0; k < nv; ++k) {
      Cell& ch = cells[cells[p].child0 + k];
      ch.active = false;
      ch.dead = true;
    }
    cells[p].child0 = -1;
    cells[p].active = true;
  }

  bool has_active_children_only(int32_t p) const {
    if (cells[p].child0 < 0) return false;
    for (int k = 0; k < vpc(); ++k) {
      const Cell& ch = cells[cells[p].child0 + k];
      if (!ch.active || ch.dead) return false;
    }
    return true;
  }

  // Lines of the active mesh in CSR form: the active cells that have the line as one of their own, and, where the line's
  // midpoint exists and finer cells touch the line, its two halves (lines of the next level).  Two cells "share a line or
  // half a line" (2D: a face; 3D: a face or an edge) iff they are members of one line, or of a line and one of its halves.
  struct LineTable {
    std::vector<int32_t> ptr, members;   // line -> active cells
    std::vector<int32_t> half;           // 2 per line: id of the half line or -1
    int64_t n_lines() const { return (int64_t)ptr.size() - 1; }
  };
  // Built concurrently: the keys of all (cell, line) pairs are computed once and bucketed by a hash of the key (a counting
  // sort that keeps the cell order inside a bucket); every bucket is searched by one thread for the FIRST pair of each key.
  // Lines are then numbered in the order of their first pair, i.e. in cell order — the numbering a serial first-seen sweep
  // gives, whatever the number of threads, and the one that keeps the members of consecutive lines close in memory for the
  // closure sweeps of prepare().
  static constexpr int N_LINE_SHARDS = 64;
  static int line_shard(const dofs::EntityKey& k) {
    const uint64_t h = (uint64_t)k.first * 0xd6e8feb86659fd93ull + (uint64_t)k.second * 0xa0761d6478bd642full;
    return (int)((h >> 40) & (N_LINE_SHARDS - 1));
  }
  LineTable line_table() const {
    dofs::RefElement ref = dofs::make_ref_element(dim, 1);
    const int nl = (int)ref.lines.size();
    std::vector<int32_t> act;
    for (size_t i = 0; i < cells.size(); ++i)
      if (cells[i].active && !cells[i].dead) act.push_back((int32_t)i);
    const int64_t na = (int64_t)act.size(), nk = na * nl;
    if (nk >= (int64_t)1 << 31) throw std::runtime_error("amr: too many cells for 32-bit line slots");
    std::vector<dofs::EntityKey> key((size_t)nk);
    std::vector<uint8_t> shard((size_t)nk);
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < na; ++i) {
      const Cell& c = cells[act[i]];
      for (int l = 0; l < nl; ++l) {
        key[(size_t)i * nl + l] = dofs::edge_key(c.v[ref.lines[l][0]], c.v[ref.lines[l][1]]);
        shard[(size_t)i * nl + l] = (uint8_t)line_shard(key[(size_t)i * nl + l]);
      }
    }
    std::vector<int64_t> bptr(N_LINE_SHARDS + 1, 0);
    for (int64_t q = 0; q < nk; ++q) bptr[shard[q] + 1]++;
    for (int sh = 0; sh < N_LINE_SHARDS; ++sh) bptr[sh + 1] += bptr[sh];
    std::vector<int32_t> bucket((size_t)nk);
    {
      std::vector<int64_t> fill_pos(bptr.begin(), bptr.end() - 1);
      for (int64_t q = 0; q < nk; ++q) bucket[fill_pos[shard[q]]++] = (int32_t)q;
    }
    // pass 1: the first pair of every key (buckets keep the cell order, so the first hit in a bucket is the first overall)
    std::vector<dofs::FlatMap<int32_t>> first_of(N_LINE_SHARDS);  // key -> its first pair q
    std::vector<int32_t> slot_line((size_t)nk);                    // pair -> first pair of its key, later the line number
    std::vector<int32_t> number((size_t)nk + 1, 0);                // 1 at first pairs, then their exclusive prefix sum
#pragma omp parallel for schedule(dynamic, 1)
    for (int sh = 0; sh < N_LINE_SHARDS; ++sh) {
      dofs::FlatMap<int32_t>& M = first_of[sh];
      M.reserve((size_t)(bptr[sh + 1] - bptr[sh]) / (dim == 2 ? 2 : 4) + 64);
      for (int64_t b = bptr[sh]; b < bptr[sh + 1]; ++b) {
        const int32_t q = bucket[b];
        auto it = M.find(key[q]);
        if (it == M.end()) { M[key[q]] = q; slot_line[q] = q; number[q + 1] = 1; }
        else slot_line[q] = it->second;
      }
    }
    for (int64_t q = 0; q < nk; ++q) number[q + 1] += number[q];  // number[q] = line number of the key whose first pair is q
    const int64_t n_lines = number[nk];
    LineTable T;
    T.ptr.assign((size_t)n_lines + 1, 0);
    std::vector<dofs::EntityKey> line_key((size_t)n_lines);
    // pass 2: line numbers of all pairs and member counts (all pairs of a key sit in one bucket: no two threads touch a line)
#pragma omp parallel for schedule(dynamic, 1)
    for (int sh = 0; sh < N_LINE_SHARDS; ++sh)
      for (int64_t b = bptr[sh]; b < bptr[sh + 1]; ++b) {
        const int32_t q = bucket[b], first = slot_line[q], id = number[first];
        if (first == q) line_key[id] = key[q];
        slot_line[q] = id;
        T.ptr[id + 1]++;
      }
    for (int64_t l = 0; l < n_lines; ++l) T.ptr[l + 1] += T.ptr[l];
    T.members.resize((size_t)T.ptr.back());
    // pass 3: members, in cell order inside every line
    {
      std::vector<int32_t> fill(T.ptr.begin(), T.ptr.end() - 1);
#pragma omp parallel for schedule(dynamic, 1)
      for (int sh = 0; sh < N_LINE_SHARDS; ++sh)
        for (int64_t b = bptr[sh]; b < bptr[sh + 1]; ++b) {
          const int32_t q = bucket[b];
          T.members[fill[slot_line[q]]++] = act[q / nl];
        }
    }
    // a line can only have a used midpoint when a finer cell touches one of its ends
    std::vector<int8_t> vmax(n_vertices(), -1);
    for (int32_t ci : act)
      for (int k = 0; k < vpc(); ++k) vmax[cells[ci].v[k]] = std::max<int8_t>(vmax[cells[ci].v[k]], (int8_t)cells[ci].level);
    auto line_number = [&](const dofs::EntityKey& k) -> int32_t {
      const dofs::FlatMap<int32_t>& M = first_of[line_shard(k)];
      auto it = M.find(k);
      return it == M.end() ? -1 : number[it->second];
    };
    T.half.assign(2 * (size_t)n_lines, -1);
#pragma omp parallel for schedule(static)
    for (int64_t l = 0; l < n_lines; ++l) {
      const int lev = cells[T.members[T.ptr[l]]].level;
      const int64_t ends[2] = {line_key[l].first, line_key[l].second};
      if (vmax[ends[0]] <= lev && vmax[ends[1]] <= lev) continue;
      auto mid = edge_mid.find(line_key[l]);
      if (mid == edge_mid.end()) continue;
      for (int e = 0; e < 2; ++e) T.half[2 * l + e] = line_number(dofs::edge_key(ends[e], mid->second));
    }
    return T;
  }

  void clear_family_coarsen(int32_t c) {
    const int32_t p = cells[c].parent;
    if (p < 0 || cells[p].child0 < 0) { cells[c].coarsen_flag = false; return; }
    for (int k = 0; k < vpc(); ++k) cells[cells[p].child0 + k].coarsen_flag = false;
  }

  // Triangulation::prepare_coarsening_and_refinement (FSS:481)
  void prepare() {
    const LineTable T = line_table();
    bool changed = true;
    while (changed) {
      changed = false;
      for (size_t i = 0; i < cells.size(); ++i) {
        Cell& c = cells[i];
        if (!c.active || c.dead) { c.refine_flag = c.coarsen_flag = false; continue; }
        if (c.refine_flag && c.coarsen_flag) { c.coarsen_flag = false; changed = true; }
        if (c.coarsen_flag && c.parent < 0) { c.coarsen_flag = false; changed = true; }
      }
      // a family is coarsened as a whole or not at all
      for (size_t p = 0; p < cells.size(); ++p) {
        if (cells[p].dead || cells[p].active || cells[p].child0 < 0) continue;
        int n_flag = 0, n_ok = 0;
        for (int k = 0; k < vpc(); ++k) {
          const Cell& ch = cells[cells[p].child0 + k];
          if (ch.active && !ch.dead) ++n_ok;
          if (ch.active && ch.coarsen_flag) ++n_flag;
        }
        if (n_flag > 0 && (n_flag < vpc() || n_ok < vpc())) {
          for (int k = 0; k < vpc(); ++k) cells[cells[p].child0 + k].coarsen_flag = false;
          changed = true;
        }
      }
      // level difference <= 1 across every shared (half) line after the pass; refinement wins over coarsening
      auto future = [&](int32_t c) { return cells[c].level + (cells[c].refine_flag ? 1 : 0) - (cells[c].coarsen_flag ? 1 : 0); };
      auto raise = [&](int32_t lo) {
        if (cells[lo].coarsen_flag) clear_family_coarsen(lo);
        else cells[lo].refine_flag = true;
        changed = true;
      };
      for (int64_t l = 0; l < T.n_lines(); ++l) {
        int top_own = -1000, top_half = -1000;
        for (int32_t q = T.ptr[l]; q < T.ptr[l + 1]; ++q) top_own = std::max(top_own, future(T.members[q]));
        for (int e = 0; e < 2; ++e) {
          const int32_t h = T.half[2 * l + e];
          if (h < 0) continue;
          for (int32_t q = T.ptr[h]; q < T.ptr[h + 1]; ++q) top_half = std::max(top_half, future(T.members[q]));
        }
        const int top = std::max(top_own, top_half);
        for (int32_t q = T.ptr[l]; q < T.ptr[l + 1]; ++q)
          if (top - future(T.members[q]) > 1) raise(T.members[q]);
        for (int e = 0; e < 2; ++e) {
          const int32_t h = T.half[2 * l + e];
          if (h < 0) continue;
          for (int32_t q = T.ptr[h]; q < T.ptr[h + 1]; ++q)
            if (top_own - future(T.members[q]) > 1) raise(T.members[q]);
        }
      }
    }
  }

  // Triangulation::execute_coarsening_and_refinement (FSS:483); returns {n_coarsened_families, n_refined}
  std::pair<int, int> execute() {
    prepare();
    int nc = 0, nr = 0;
    const size_t n0 = cells.size();
    for (size_t p = 0; p < n0; ++p) {
      if (cells[p].dead || cells[p].active || !has_active_children_only((int32_t)p)) continue;
      bool all = true;
      for (int k = 0; k < vpc(); ++k) all = all && cells[cells[p].child0 + k].coarsen_flag;
      if (all) { coarsen_cell((int32_t)p); ++nc; }
    }
    std::vector<int32_t> todo;
    for (size_t i = 0; i < n0; ++i)
      if (cells[i].active && !cells[i].dead && cells[i].refine_flag) todo.push_back((int32_t)i);
    for (int32_t c : todo) { refine_cell(c); ++nr; }
    for (auto& c : cells) c.refine_flag = c.coarsen_flag = false;
    return {nc, nr};
  }

  // ---- solution transfer of FE_Q(1) fields keyed by vertex (FSS:475-497)
  void store_vertex_values(const mesh::Mesh& m, const dofs::DofMap& dp, int n_vec, const double* const* vec) {
    if (dp.degree != 1 || dp.n_comp != 1) throw std::runtime_error("amr: the transfer handles the FE_Q(1) pressure handler only (FSS:475)");
    if (n_vec < 1 || n_vec > MAX_TRANSFER) throw std::runtime_error("amr: between 1 and 8 vectors can be transferred at once");
    n_transfer = n_vec;
    vval.assign((size_t)n_vertices() * n_vec, 0.0);
    vknown.assign(n_vertices(), 0);
    const int nv = vpc();
    for (int64_t c = 0; c < m.n_cells(); ++c)
      for (int k = 0; k < nv; ++k) {
        const int32_t v = m.cell_vertices[c * nv + k], d = dp.cell_dofs[c * nv + k];
        vknown[v] = 1;
        for (int t = 0; t < n_vec; ++t) vval[(size_t)v * n_vec + t] = vec[t][d];
      }
  }
  void fetch_vertex_values(const mesh::Mesh& m, const dofs::DofMap& dp, int n_vec, double* const* vec) const {
    if (n_vec != n_transfer) throw std::runtime_error("amr: fetch without a matching store");
    const int nv = vpc();
    for (int64_t c = 0; c < m.n_cells(); ++c)
      for (int k = 0; k < nv; ++k) {
        const int32_t v = m.cell_vertices[c * nv + k], d = dp.cell_dofs[c * nv + k];
        if (!vknown[v]) throw std::runtime_error("amr: vertex without a transferred value");
        for (int t = 0; t < n_vec; ++t) vec[t][d] = vval[(size_t)v * n_vec + t];
      }
  }
};

// ------------------------------------------------------------------------------------------------
// DoFTools::make_hanging_node_constraints for FE_Q(1|2)^n_comp on the active mesh of F
// ------------------------------------------------------------------------------------------------
inline void q2_weights_1d(double x, double w[3]) {  // nodes 0, 1/2, 1
  w[0] = 2 * x * x - 3 * x + 1;
  w[1] = 4 * x * (1 - x);
  w[2] = 2 * x * x - x;
}

inline void hanging_node_constraints(const Forest& F, const mesh::Mesh& m, const dofs::DofMap& d, const dofs::NodeMaps& maps,
                                     dofs::ConstraintTable& T) {
  const int dim = F.dim, vpc = 1 << dim, nc = d.n_comp;
  dofs::RefElement ref = dofs::make_ref_element(dim, 1);
  std::vector<uint8_t> used(F.n_vertices(), 0);
  for (int64_t c = 0; c < m.n_cells(); ++c)
    for (int k = 0; k < vpc; ++k) used[m.cell_vertices[c * vpc + k]] = 1;
  // Smallest active cell around every vertex (by its longest line).  The midpoint of a line / quad of cell c can only be
  // in use when a smaller cell touches one of its corners (the refined neighbour's children do), so entities whose corners
  // see nothing smaller than c are skipped without a hash look-up — on mostly uniform meshes that is nearly all of them.
  std::vector<int32_t> cell_level(m.n_cells());
  {
    std::vector<int32_t> act = F.active_cells();
    for (int64_t c = 0; c < m.n_cells(); ++c) cell_level[c] = F.cells[act[c]].level;
  }
  std::vector<int8_t> vmax(F.n_vertices(), -1);
  for (int64_t c = 0; c < m.n_cells(); ++c)
    for (int k = 0; k < vpc; ++k) vmax[m.cell_vertices[c * vpc + k]] = std::max<int8_t>(vmax[m.cell_vertices[c * vpc + k]], (int8_t)cell_level[c]);
  auto line_dof = [&](int64_t a, int64_t b) {
    auto it = maps.line_dof.find(dofs::edge_key(a, b));
    return it == maps.line_dof.end() ? -1 : it->second;
  };
  auto add = [&](int32_t base, const std::vector<std::pair<int32_t, double>>& e) {  // same weights for every component
    if (base < 0) return;
    for (int k = 0; k < nc; ++k) {
      std::vector<std::pair<int32_t, double>> ek;
      for (auto& x : e)
        if (x.second != 0.0) ek.push_back({x.first + k, x.second});
      T.add_line(base + k, ek, 0.0);
    }
  };
  for (int64_t c = 0; c < m.n_cells(); ++c) {
    const int32_t* cv = &m.cell_vertices[c * vpc];
    // lines of the cell whose midpoint is a vertex of finer active cells
    const int lev = cell_level[c];
    for (auto& l : ref.lines) {
      const int32_t a = cv[l[0]], b = cv[l[1]];
      if (vmax[a] <= lev && vmax[b] <= lev) continue;
      auto it = F.edge_mid.find(dofs::edge_key(a, b));
      if (it == F.edge_mid.end() || !used[it->second]) continue;
      const int32_t mv = it->second;
      if (d.degree == 1) {
        add(maps.vdof[mv], {{maps.vdof[a], 0.5}, {maps.vdof[b], 0.5}});
      } else {
        const int32_t L = line_dof(a, b);
        add(maps.vdof[mv], {{L, 1.0}});
        add(line_dof(a, mv), {{maps.vdof[a], 0.375}, {L, 0.75}, {maps.vdof[b], -0.125}});
        add(line_dof(mv, b), {{maps.vdof[a], -0.125}, {L, 0.75}, {maps.vdof[b], 0.375}});
      }
    }
    if (dim != 3) continue;
    // quads of the cell that are refined on the other side
    for (auto& q : ref.quads) {
      const int32_t f[4] = {cv[q[0]], cv[q[1]], cv[q[2]], cv[q[3]]};  // lexicographic in the face frame
      if (vmax[f[0]] <= lev && vmax[f[1]] <= lev && vmax[f[2]] <= lev && vmax[f[3]] <= lev) continue;
      auto it = F.face_mid.find(dofs::quad_key(f[0], f[1], f[2], f[3]));
      if (it == F.face_mid.end() || !used[it->second]) continue;
      const int32_t fc = it->second;
      if (d.degree == 1) {
        add(maps.vdof[fc], {{maps.vdof[f[0]], 0.25}, {maps.vdof[f[1]], 0.25}, {maps.vdof[f[2]], 0.25}, {maps.vdof[f[3]], 0.25}});
        continue;
      }
      // 3 x 3 lattice of fine vertices on the coarse face and the 9 coarse face dofs, both indexed (i, j) in {0,1,2}^2
      auto emid = [&](int32_t a, int32_t b) { return F.edge_mid.at(dofs::edge_key(a, b)); };
      int32_t P[3][3];
      P[0][0] = f[0]; P[2][0] = f[1]; P[0][2] = f[2]; P[2][2] = f[3];
      P[1][0] = emid(f[0], f[1]); P[1][2] = emid(f[2], f[3]); P[0][1] = emid(f[0], f[2]); P[2][1] = emid(f[1], f[3]);
      P[1][1] = fc;
      int32_t D[3][3];
      D[0][0] = maps.vdof[f[0]]; D[2][0] = maps.vdof[f[1]]; D[0][2] = maps.vdof[f[2]]; D[2][2] = maps.vdof[f[3]];
      D[1][0] = line_dof(f[0], f[1]); D[1][2] = line_dof(f[2], f[3]); D[0][1] = line_dof(f[0], f[2]); D[2][1] = line_dof(f[1], f[3]);
      {
        auto qit = maps.quad_dof.find(dofs::quad_key(f[0], f[1], f[2], f[3]));
        D[1][1] = qit == maps.quad_dof.end() ? -1 : qit->second;
      }
      auto constrain = [&](int32_t base, double xi, double eta) {
        double wx[3], wy[3];
        q2_weights_1d(xi, wx);
        q2_weights_1d(eta, wy);
        std::vector<std::pair<int32_t, double>> e;
        for (int j = 0; j < 3; ++j)
          for (int i = 0; i < 3; ++i)
            if (wx[i] * wy[j] != 0.0) e.push_back({D[i][j], wx[i] * wy[j]});
        add(base, e);
      };
      constrain(maps.vdof[fc], 0.5, 0.5);
      // the four child lines that meet at the face centre
      constrain(line_dof(P[1][0], fc), 0.5, 0.25);
      constrain(line_dof(fc, P[1][2]), 0.5, 0.75);
      constrain(line_dof(P[0][1], fc), 0.25, 0.5);
      constrain(line_dof(fc, P[2][1]), 0.75, 0.5);
      // the four child quads
      for (int j = 0; j < 2; ++j)
        for (int i = 0; i < 2; ++i) {
          auto qit = maps.quad_dof.find(dofs::quad_key(P[i][j], P[i + 1][j], P[i][j + 1], P[i + 1][j + 1]));
          if (qit != maps.quad_dof.end()) constrain(qit->second, 0.25 + 0.5 * i, 0.25 + 0.5 * j);
        }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// KellyErrorEstimator<dim>::estimate for a scalar FE_Q(1) field given per vertex
// ------------------------------------------------------------------------------------------------
namespace detail {

inline double det_inv_T(int dim, const double* J, double* JiT) {  // J[a*dim+b] = dx_a/dxi_b
  if (dim == 2) {
    const double det = J[0] * J[3] - J[1] * J[2];
    JiT[0] = J[3] / det; JiT[1] = -J[2] / det; JiT[2] = -J[1] / det; JiT[3] = J[0] / det;
    return det;
  }
  const double c00 = J[4] * J[8] - J[5] * J[7], c01 = J[5] * J[6] - J[3] * J[8], c02 = J[3] * J[7] - J[4] * J[6];
  const double det = J[0] * c00 + J[1] * c01 + J[2] * c02;
  JiT[0] = c00 / det; JiT[1] = c01 / det; JiT[2] = c02 / det;
  JiT[3] = (J[2] * J[7] - J[1] * J[8]) / det; JiT[4] = (J[0] * J[8] - J[2] * J[6]) / det; JiT[5] = (J[1] * J[6] - J[0] * J[7]) / det;
  JiT[6] = (J[1] * J[5] - J[2] * J[4]) / det; JiT[7] = (J[2] * J[3] - J[0] * J[5]) / det; JiT[8] = (J[0] * J[4] - J[1] * J[3]) / det;
  return det;
}

// Q1 geometry and field gradient of one cell at a reference point.  The dimension is a template parameter so that the
// loops unroll (the estimator evaluates this eight times per interior face); the order of the floating-point operations
// is the same for both instantiations as in a plain loop over (vertex, axis), so the indicators do not depend on it.
template <int DIM>
inline double cell_gradient_t(const Forest& F, const Cell& c, const double* xi, const double* vertex_value, double* grad, double* JiT) {
  constexpr int vpc = 1 << DIM;
  double w[DIM][2];
  for (int b = 0; b < DIM; ++b) { w[b][0] = 1 - xi[b]; w[b][1] = xi[b]; }
  double J[9] = {0}, gref[3] = {0, 0, 0};
  for (int v = 0; v < vpc; ++v) {
    double dN[DIM];
    for (int a = 0; a < DIM; ++a) {
      double g = ((v >> a) & 1) ? 1.0 : -1.0;
      for (int b = 0; b < DIM; ++b)
        if (b != a) g *= w[b][(v >> b) & 1];
      dN[a] = g;
    }
    const double* X = &F.xyz[(int64_t)c.v[v] * DIM];
    const double val = vertex_value[c.v[v]];
    for (int a = 0; a < DIM; ++a) {
      gref[a] += val * dN[a];
      for (int b = 0; b < DIM; ++b) J[a * DIM + b] += X[a] * dN[b];
    }
  }
  const double det = det_inv_T(DIM, J, JiT);
  for (int a = 0; a < DIM; ++a) {
    double s = 0;
    for (int b = 0; b < DIM; ++b) s += JiT[a * DIM + b] * gref[b];
    grad[a] = s;
  }
  return det;
}
inline double cell_gradient(const Forest& F, const Cell& c, const double* xi, const double* vertex_value, double* grad, double* JiT) {
  return F.dim == 2 ? cell_gradient_t<2>(F, c, xi, vertex_value, grad, JiT) : cell_gradient_t<3>(F, c, xi, vertex_value, grad, JiT);
}

// parametric position of vertex v inside face f of cell N (corner, line midpoint or face centre)
inline bool face_param(const Forest& F, const Cell& N, int f, int32_t v, double out[2]) {
  const int dim = F.dim, nfv = 1 << (dim - 1);
  int fv[4];
  mesh::face_vertices(dim, f, fv);
  int32_t P[4];
  for (int k = 0; k < nfv; ++k) P[k] = N.v[fv[k]];
  for (int k = 0; k < nfv; ++k)
    if (P[k] == v) { out[0] = k & 1; out[1] = (k >> 1) & 1; return true; }
  for (int k1 = 0; k1 < nfv; ++k1)
    for (int k2 = k1 + 1; k2 < nfv; ++k2) {
      const int diff = k1 ^ k2;
      if (diff != 1 && diff != 2) continue;
      auto it = F.edge_mid.find(dofs::edge_key(P[k1], P[k2]));
      if (it != F.edge_mid.end() && it->second == v) {
        out[0] = 0.5 * ((k1 & 1) + (k2 & 1));
        out[1] = 0.5 * (((k1 >> 1) & 1) + ((k2 >> 1) & 1));
        return true;
      }
    }
  if (dim == 3) {
    auto it = F.face_mid.find(dofs::quad_key(P[0], P[1], P[2], P[3]));
    if (it != F.face_mid.end() && it->second == v) { out[0] = out[1] = 0.5; return true; }
  }
  return false;
}

inline void embed_face_point(int dim, int f, const double* s, double* xi) {
  const int axis = f / 2, side = f % 2;
  int t = 0;
  for (int a = 0; a < dim; ++a) xi[a] = (a == axis) ? (double)side : s[t++];
}

}  // namespace detail

// vertex_value: one value per forest vertex (pressure dof of that vertex).  Returns eta per active cell, in the order
// of Forest::active_cells(), rounded to float like the reference's Vector<float> (FSS:452).
inline std::vector<float> kelly_estimate(const Forest& F, const std::vector<double>& vertex_value) {
  const int dim = F.dim, nfv = 1 << (dim - 1);
  std::vector<int32_t> act = F.active_cells();
  std::vector<int32_t> pos(F.cells.size(), -1);
  for (size_t i = 0; i < act.size(); ++i) pos[act[i]] = (int32_t)i;
  auto face_key_of = [&](const Cell& c, int f) {
    int fv[4];
    mesh::face_vertices(dim, f, fv);
    return dim == 2 ? dofs::edge_key(c.v[fv[0]], c.v[fv[1]]) : dofs::quad_key(c.v[fv[0]], c.v[fv[1]], c.v[fv[2]], c.v[fv[3]]);
  };
  struct FaceSides {  // an interior face of the active mesh has two sides, a boundary or coarse/fine face one
    int32_t cell[2] = {-1, -1};
    int8_t face[2] = {0, 0};
    int8_t n = 0;
  };
  // The table is sharded by a hash of the key so that the threads can build it concurrently: the keys of all
  // (cell, face) pairs are computed once, bucketed by shard with a counting sort (which keeps the cell order inside a
  // shard, so the table does not depend on the number of threads), and every shard is filled by one thread.
  constexpr int N_SHARDS = 64;
  auto shard_of = [](const dofs::EntityKey& k) {
    uint64_t h = (uint64_t)k.first * 0xd6e8feb86659fd93ull + (uint64_t)k.second * 0xa0761d6478bd642full;
    return (int)((h >> 40) & (N_SHARDS - 1));
  };
  const int nf = 2 * dim;
  const int64_t na = (int64_t)act.size(), nkeys = na * nf;
  std::vector<dofs::EntityKey> keys((size_t)nkeys);
  std::vector<uint8_t> key_shard((size_t)nkeys);
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < na; ++i)
    for (int f = 0; f < nf; ++f) {
      keys[(size_t)i * nf + f] = face_key_of(F.cells[act[i]], f);
      key_shard[(size_t)i * nf + f] = (uint8_t)shard_of(keys[(size_t)i * nf + f]);
    }
  std::vector<int64_t> bucket_ptr(N_SHARDS + 1, 0);
  for (int64_t q = 0; q < nkeys; ++q) bucket_ptr[key_shard[q] + 1]++;
  for (int sh = 0; sh < N_SHARDS; ++sh) bucket_ptr[sh + 1] += bucket_ptr[sh];
  std::vector<int64_t> bucket((size_t)nkeys), fill_pos(bucket_ptr.begin(), bucket_ptr.end() - 1);
  for (int64_t q = 0; q < nkeys; ++q) bucket[fill_pos[key_shard[q]]++] = q;
  std::vector<dofs::FlatMap<FaceSides>> shards(N_SHARDS);
  int too_many = 0;
#pragma omp parallel for schedule(dynamic, 1)
  for (int sh = 0; sh < N_SHARDS; ++sh) {
    dofs::FlatMap<FaceSides>& M = shards[sh];
    M.reserve((size_t)(bucket_ptr[sh + 1] - bucket_ptr[sh]) / 2 + 64);
    for (int64_t b = bucket_ptr[sh]; b < bucket_ptr[sh + 1]; ++b) {
      const int64_t q = bucket[b];
      FaceSides& s = M[keys[q]];
      if (s.n >= 2) {
#pragma omp atomic write
        too_many = 1;
        continue;
      }
      s.cell[s.n] = act[q / nf];
      s.face[s.n] = (int8_t)(q % nf);
      s.n++;
    }
  }
  if (too_many) throw std::runtime_error("amr: a face with more than two cells");
  struct ShardedFaces {
    const std::vector<dofs::FlatMap<FaceSides>>& shards;
    decltype(shard_of)& shard;
    const dofs::FlatMap<FaceSides>::Slot* find(const dofs::EntityKey& k) const { return shards[shard(k)].find(k); }
    const dofs::FlatMap<FaceSides>::Slot* end() const { return nullptr; }
  } active_face{shards, shard_of};
  // QGauss<dim-1>(2)
  const double ga = 0.5 - 0.5 / std::sqrt(3.0), gb = 0.5 + 0.5 / std::sqrt(3.0);
  std::vector<std::array<double, 2>> qp;
  std::vector<double> qw;
  if (dim == 2) { qp = {{ga, 0}, {gb, 0}}; qw = {0.5, 0.5}; }
  else { qp = {{ga, ga}, {gb, ga}, {ga, gb}, {gb, gb}}; qw = {0.25, 0.25, 0.25, 0.25}; }
  // face integral of the squared jump; face f of K is (a part of) face fn of N
  auto integrate = [&](int32_t k, int f, int32_t n, int fn) -> double {
    const Cell& K = F.cells[k];
    const Cell& N = F.cells[n];
    int fv[4];
    mesh::face_vertices(dim, f, fv);
    double par[4][2];
    for (int i = 0; i < nfv; ++i)
      if (!detail::face_param(F, N, fn, K.v[fv[i]], par[i])) throw std::runtime_error("amr: faces do not match (mesh not 2:1 balanced?)");
    double integral = 0;
    for (size_t q = 0; q < qp.size(); ++q) {
      double xiK[3], xiN[3], sN[2] = {0, 0};
      detail::embed_face_point(dim, f, qp[q].data(), xiK);
      for (int i = 0; i < nfv; ++i) {
        const double w = ((i & 1) ? qp[q][0] : 1 - qp[q][0]) * (dim == 3 ? (((i >> 1) & 1) ? qp[q][1] : 1 - qp[q][1]) : 1.0);
        sN[0] += w * par[i][0];
        sN[1] += w * par[i][1];
      }
      detail::embed_face_point(dim, fn, sN, xiN);
      double gK[3], gN[3], JiT[9], JiTn[9];
      const double det = detail::cell_gradient(F, K, xiK, vertex_value.data(), gK, JiT);
      detail::cell_gradient(F, N, xiN, vertex_value.data(), gN, JiTn);
      const int axis = f / 2;
      double nv[3], nn = 0;
      for (int a = 0; a < dim; ++a) { nv[a] = JiT[a * dim + axis]; nn += nv[a] * nv[a]; }
      nn = std::sqrt(nn);
      double jump = 0;
      for (int a = 0; a < dim; ++a) jump += (gK[a] - gN[a]) * nv[a] / nn;
      integral += jump * jump * std::fabs(det) * nn * qw[q];
    }
    return integral;
  };
  // Pass 1 (threads): every interior face is integrated once — by the lower-numbered cell of a regular face, by the fine
  // cell of a coarse/fine face — into its own slot.  Pass 2 (sequential, fixed order): the slots are added to the two
  // cells, so the sums do not depend on the number of threads.
  std::vector<double> face_int((size_t)na * nf, 0.0);
  std::vector<int32_t> partner((size_t)na * nf, -1);
  int failed = 0;
#pragma omp parallel for schedule(dynamic, 256)
  for (int64_t i = 0; i < na; ++i) {
    const int32_t k = act[i];
    const Cell& K = F.cells[k];
    for (int f = 0; f < nf; ++f) {
      if (K.bid[f] >= 0) continue;  // no Neumann function map at FSS:456: boundary faces contribute nothing
      int32_t n = -1;
      int fn = 0;
      bool regular = false;
      const auto same = active_face.find(keys[(size_t)i * nf + f]);
      for (int e = 0; same != active_face.end() && e < same->second.n; ++e)
        if (same->second.cell[e] != k) { n = same->second.cell[e]; fn = same->second.face[e]; regular = true; }
      if (regular) {
        if (!(k < n)) continue;  // the other side integrates this face
      } else {
        if (K.parent < 0 || ((K.child_index >> (f / 2)) & 1) != (f % 2)) continue;  // interior face of the parent / finer neighbour
        const auto it = active_face.find(face_key_of(F.cells[K.parent], f));
        if (it == active_face.end() || it->second.n == 0) continue;  // the neighbour is finer: integrated from its side
        n = it->second.cell[0];
        fn = it->second.face[0];
      }
      try {
        face_int[(size_t)i * nf + f] = integrate(k, f, n, fn);
        partner[(size_t)i * nf + f] = n;
      } catch (const std::exception&) {
#pragma omp atomic write
        failed = 1;
      }
    }
  }
  if (failed) throw std::runtime_error("amr: faces do not match (mesh not 2:1 balanced?)");
  std::vector<double> sum(act.size(), 0.0);
  for (int64_t i = 0; i < na; ++i)
    for (int f = 0; f < nf; ++f) {
      const int32_t n = partner[(size_t)i * nf + f];
      if (n < 0) continue;
      sum[i] += face_int[(size_t)i * nf + f];
      sum[pos[n]] += face_int[(size_t)i * nf + f];
    }
  std::vector<float> eta(act.size());
  for (size_t i = 0; i < act.size(); ++i) {
    const Cell& K = F.cells[act[i]];
    double h = 0;  // cell->diameter(): longest diagonal
    for (int v = 0; v < (1 << (dim - 1)); ++v) {
      const int w = ((1 << dim) - 1) ^ v;
      double s = 0;
      for (int a = 0; a < dim; ++a) {
        const double dd = F.xyz[(int64_t)K.v[v] * dim + a] - F.xyz[(int64_t)K.v[w] * dim + a];
        s += dd * dd;
      }
      h = std::max(h, std::sqrt(s));
    }
    eta[i] = (float)std::sqrt(sum[i] * h / 24.0);
  }
  return eta;
}

// GridRefinement::refine_and_coarsen_fixed_fraction(tria, criteria, top, bottom) followed by the level limits of
// FSS:463-472.  criteria[i] belongs to active cell i of Forest::active_cells().
inline void mark_fixed_fraction(Forest& F, const std::vector<float>& criteria, double top_fraction, double bottom_fraction, int min_level,
                                int max_level) {
  std::vector<int32_t> act = F.active_cells();
  if (criteria.size() != act.size()) throw std::runtime_error("amr: one criterion per active cell expected");
  if (act.empty()) return;
  for (int32_t c : act) F.cells[c].refine_flag = F.cells[c].coarsen_flag = false;
  std::vector<float> tmp(criteria);
  double total_error = 0;
  for (float v : tmp) total_error += std::fabs((double)v);
  std::sort(tmp.begin(), tmp.end(), std::greater<float>());
  size_t pp = 0;
  for (double s = 0; s < top_fraction * total_error && pp != tmp.size() - 1; ++pp) s += tmp[pp];
  double top_threshold = pp != 0 ? ((double)tmp[pp] + (double)tmp[pp - 1]) / 2 : (double)tmp[pp];
  size_t qq = tmp.size() - 1;
  for (double s = 0; s < bottom_fraction * total_error && qq != 0; --qq) s += tmp[qq];
  double bottom_threshold = qq != tmp.size() - 1 ? ((double)tmp[qq] + (double)tmp[qq + 1]) / 2 : 0.0;
  const double cmax = *std::max_element(criteria.begin(), criteria.end()), cmin = *std::min_element(criteria.begin(), criteria.end());
  if (top_threshold == cmax && top_fraction != 1) top_threshold *= 0.999;
  if (bottom_threshold >= top_threshold) bottom_threshold = 0.999 * top_threshold;
  if (top_threshold < cmax) {  // GridRefinement::refine(tria, criteria, top_threshold, pp)
    bool all_zero = true;
    for (float v : criteria) all_zero = all_zero && v == 0;
    if (!all_zero) {
      double thr = top_threshold;
      if (thr == 0) {
        thr = criteria[0];
        for (float v : criteria)
          if (v > 0 && v < thr) thr = v;
      }
      size_t marked = 0;
      for (size_t i = 0; i < act.size(); ++i)
        if (std::fabs((double)criteria[i]) >= thr) {
          if (marked >= pp) break;
          ++marked;
          F.cells[act[i]].refine_flag = true;
        }
    }
  }
  if (bottom_threshold > cmin)  // GridRefinement::coarsen
    for (size_t i = 0; i < act.size(); ++i)
      if (std::fabs((double)criteria[i]) <= bottom_threshold && !F.cells[act[i]].refine_flag) F.cells[act[i]].coarsen_flag = true;
  // FSS:463-472
  if (F.n_levels() > max_level)
    for (int32_t c : act)
      if (F.cells[c].level >= max_level) F.cells[c].refine_flag = false;
  for (int32_t c : act)
    if (F.cells[c].level == min_level) F.cells[c].coarsen_flag = false;
}

}  // namespace amr
