// partition.hpp — cell partition of a mesh across the GPUs of one box.
//
// The reference is serial (Triangulation<dim>, Vector<double>, SparseMatrix<double>:
// PoroelasticityFSS.h:75, PoroElasticPressureSolver.h:36-44); this is the new-build domain
// decomposition.  Cells are split into `nranks` contiguous ranges of the global cell order
// (octants/boxes for Morton-ordered 2^L grids, slabs for lexicographic ones).  A dof is owned by
// the rank of the first cell that touches it (== the cell that numbered it, first-touch).  A rank
// keeps every cell that touches one of its owned dofs (its own cells plus one ghost layer), so
// owned matrix rows assemble without communication.  Local dof order: [owned interior | owned boundary |
// ghosts grouped by owner rank], each group ascending in the global id.  "Interior" dofs share no cell
// with a ghost dof, so their matrix rows need no halo value: the device runs those rows while the halo
// exchange is still in flight.  Ghosts grouped by owner make every receive land contiguously.
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <map>
#include <set>
#include <vector>

#include "dofs.hpp"
#include "mesh.hpp"

namespace partition {

struct FieldPart {
  int64_t n_owned = 0, n_local = 0;
  std::vector<int32_t> cell_dofs;         // local cells * n_loc (local ids)
  std::vector<int64_t> local_to_global;   // n_local
  std::vector<int32_t> neighbor_rank;
  std::vector<int64_t> send_ptr, recv_ptr;  // n_neighbors+1
  std::vector<int32_t> send_idx;            // local (owned) ids, grouped by neighbour
};

struct Part {
  mesh::Mesh mesh;  // local sub-mesh
  std::vector<int64_t> cell_global;
  int64_t n_owned_cells = 0;
  FieldPart field[2];
};

inline int rank_of_cell(int64_t c, int64_t n_cells, int nranks) {
  // inverse of c0(r) = n_cells*r/nranks
  int r = (int)(((__int128)(c + 1) * nranks - 1) / n_cells);
  while (r > 0 && (int64_t)((__int128)n_cells * r / nranks) > c) --r;
  while (r + 1 < nranks && (int64_t)((__int128)n_cells * (r + 1) / nranks) <= c) ++r;
  return r;
}

// Ownership of the dofs on partition interfaces.  Default: the lowest rank whose cells touch the dof (= the rank of the
// cell that numbered it).  On a 2x2x2 octant split that hands all three interface planes of an octant to the lower rank:
// rank 0 owns (65/64)^3 = 4.7 % more rows than rank 7 at 128^3 cells, and every CG iteration waits for rank 0.
// `balanced` (PE_BALANCED_OWNERSHIP=1) deals the interface nodes out among the ranks that touch them — all components of
// a node stay together (block rows) — by a hash of the node number every rank computes alike.
inline bool balanced_ownership_requested() {
  const char* e = std::getenv("PE_BALANCED_OWNERSHIP");
  return e && e[0] == '1';
}

inline std::vector<int32_t> dof_owner(const mesh::Mesh& m, const dofs::DofMap& d, int nranks, bool balanced = false) {
  std::vector<int32_t> owner(d.n_dofs, -1);
  const int64_t nc = m.n_cells();
  for (int64_t c = 0; c < nc; ++c) {
    int r = rank_of_cell(c, nc, nranks);
    for (int k = 0; k < d.n_loc; ++k) {
      int32_t g = d.cell_dofs[c * d.n_loc + k];
      if (owner[g] < 0) owner[g] = r;  // cells ascend, so the first toucher has the lowest rank
    }
  }
  if (!balanced || nranks == 1) return owner;
  // ranks touching every node (a node = n_comp consecutive dofs): at most 8 on box partitions, 16 slots to be safe
  const int nc_ = d.n_comp;
  const int64_t n_nodes = d.n_dofs / nc_;
  std::vector<int16_t> touch((size_t)n_nodes * 16, -1);
  for (int64_t c = 0; c < nc; ++c) {
    const int r = rank_of_cell(c, nc, nranks);
    for (int k = 0; k < d.n_loc; k += nc_) {
      int16_t* t = &touch[(size_t)(d.cell_dofs[c * d.n_loc + k] / nc_) * 16];
      for (int q = 0; q < 16; ++q) {
        if (t[q] == r) break;
        if (t[q] < 0) { t[q] = (int16_t)r; break; }
      }
    }
  }
  for (int64_t node = 0; node < n_nodes; ++node) {
    int16_t* t = &touch[(size_t)node * 16];
    int n = 0;
    while (n < 16 && t[n] >= 0) ++n;
    if (n <= 1) continue;
    std::sort(t, t + n);
    uint64_t h = (uint64_t)node * 0x9e3779b97f4a7c15ull;
    h ^= h >> 29;
    const int pick = t[(h >> 8) % (uint64_t)n];
    for (int k = 0; k < nc_; ++k) owner[node * nc_ + k] = pick;
  }
  return owner;
}

inline void build_field(const mesh::Mesh& m, const dofs::DofMap& d, const std::vector<int32_t>& owner,
                        const std::vector<std::vector<int>>& cell_ranks_of_local, const std::vector<int64_t>& local_cells, int rank,
                        FieldPart& F) {
  (void)m;
  // local dofs
  std::vector<int32_t> owned, ghost;
  {
    std::vector<int32_t> all;
    all.reserve(local_cells.size() * d.n_loc);
    for (int64_t c : local_cells)
      for (int k = 0; k < d.n_loc; ++k) all.push_back(d.cell_dofs[c * d.n_loc + k]);
    std::sort(all.begin(), all.end());
    all.erase(std::unique(all.begin(), all.end()), all.end());
    for (int32_t g : all) (owner[g] == rank ? owned : ghost).push_back(g);
  }
  {  // interior-first order of the owned dofs
    std::vector<uint8_t> boundary(d.n_dofs, 0);
    for (int64_t c : local_cells) {
      bool has_ghost = false;
      for (int k = 0; k < d.n_loc; ++k) has_ghost |= owner[d.cell_dofs[c * d.n_loc + k]] != rank;
      if (has_ghost)
        for (int k = 0; k < d.n_loc; ++k) boundary[d.cell_dofs[c * d.n_loc + k]] = 1;
    }
    std::stable_sort(owned.begin(), owned.end(), [&](int32_t a, int32_t b) { return boundary[a] != boundary[b] ? boundary[a] < boundary[b] : a < b; });
  }
  std::stable_sort(ghost.begin(), ghost.end(), [&](int32_t a, int32_t b) { return owner[a] != owner[b] ? owner[a] < owner[b] : a < b; });
  F.n_owned = (int64_t)owned.size();
  F.n_local = F.n_owned + (int64_t)ghost.size();
  F.local_to_global.clear();
  for (int32_t g : owned) F.local_to_global.push_back(g);
  for (int32_t g : ghost) F.local_to_global.push_back(g);
  std::vector<int32_t> g2l(d.n_dofs, -1);
  for (int64_t i = 0; i < F.n_local; ++i) g2l[F.local_to_global[i]] = (int32_t)i;
  F.cell_dofs.resize(local_cells.size() * d.n_loc);
  for (size_t lc = 0; lc < local_cells.size(); ++lc)
    for (int k = 0; k < d.n_loc; ++k) F.cell_dofs[lc * d.n_loc + k] = g2l[d.cell_dofs[local_cells[lc] * d.n_loc + k]];
  // receive plan: ghosts grouped by owner
  std::map<int, std::vector<int32_t>> send_sets;  // neighbour -> owned global dofs that it needs
  std::map<int, int64_t> recv_count;
  for (int32_t g : ghost) recv_count[owner[g]]++;
  // send plan: my owned dofs in local cells that also belong to another rank's local set
  for (size_t lc = 0; lc < local_cells.size(); ++lc) {
    const auto& ranks = cell_ranks_of_local[lc];
    if (ranks.size() <= 1) continue;
    int64_t c = local_cells[lc];
    for (int k = 0; k < d.n_loc; ++k) {
      int32_t g = d.cell_dofs[c * d.n_loc + k];
      if (owner[g] != rank) continue;
      for (int r : ranks)
        if (r != rank) send_sets[r].push_back(g);
    }
  }
  std::set<int> neigh;
  for (auto& kv : send_sets) neigh.insert(kv.first);
  for (auto& kv : recv_count) neigh.insert(kv.first);
  F.neighbor_rank.assign(neigh.begin(), neigh.end());
  F.send_ptr.assign(1, 0);
  F.recv_ptr.assign(1, 0);
  F.send_idx.clear();
  for (int r : F.neighbor_rank) {
    auto& s = send_sets[r];
    std::sort(s.begin(), s.end());
    s.erase(std::unique(s.begin(), s.end()), s.end());
    for (int32_t g : s) F.send_idx.push_back(g2l[g]);
    F.send_ptr.push_back((int64_t)F.send_idx.size());
    F.recv_ptr.push_back(F.recv_ptr.back() + recv_count[r]);
  }
}

inline Part make_part(const mesh::Mesh& m, const dofs::DofMap& dp, const dofs::DofMap& du, int rank, int nranks) {
  Part P;
  const int64_t nc = m.n_cells();
  const int vpc = m.vpc(), dim = m.dim;
  const bool balanced = balanced_ownership_requested();
  std::vector<int32_t> own_p = dof_owner(m, dp, nranks, false), own_u = dof_owner(m, du, nranks, balanced);
  if (balanced)  // a pressure dof follows the displacement dofs of its vertex, so that "cells touching an owned u dof" covers both fields
    for (int64_t c = 0; c < nc; ++c)
      for (int v = 0; v < vpc; ++v) own_p[dp.cell_dofs[c * dp.n_loc + v]] = own_u[du.cell_dofs[c * du.n_loc + v * du.n_comp]];
  // ranks that keep each cell = owners of its dofs (u dofs are a superset of the vertex dofs)
  std::vector<int64_t> local_cells;
  std::vector<std::vector<int>> cell_ranks;
  for (int64_t c = 0; c < nc; ++c) {
    int rs[128];
    int n = 0;
    bool mine = false;
    for (int k = 0; k < du.n_loc; ++k) {
      int r = own_u[du.cell_dofs[c * du.n_loc + k]];
      bool seen = false;
      for (int j = 0; j < n; ++j) if (rs[j] == r) { seen = true; break; }
      if (!seen && n < 128) rs[n++] = r;
      if (r == rank) mine = true;
    }
    if (mine) {
      local_cells.push_back(c);
      cell_ranks.emplace_back(rs, rs + n);
    }
  }
  P.cell_global = local_cells;
  P.n_owned_cells = 0;
  for (int64_t c : local_cells)
    if (rank_of_cell(c, nc, nranks) == rank) P.n_owned_cells++;
  // local sub-mesh
  P.mesh.dim = dim;
  P.mesh.morton = false;
  std::vector<int32_t> v2l(m.n_vertices(), -1);
  P.mesh.cell_vertices.resize(local_cells.size() * vpc);
  std::vector<int32_t> g2lc(nc, -1);
  for (size_t lc = 0; lc < local_cells.size(); ++lc) {
    g2lc[local_cells[lc]] = (int32_t)lc;
    for (int v = 0; v < vpc; ++v) {
      int32_t gv = m.cell_vertices[local_cells[lc] * vpc + v];
      if (v2l[gv] < 0) {
        v2l[gv] = (int32_t)(P.mesh.xyz.size() / dim);
        for (int a = 0; a < dim; ++a) P.mesh.xyz.push_back(m.xyz[(int64_t)gv * dim + a]);
      }
      P.mesh.cell_vertices[lc * vpc + v] = v2l[gv];
    }
  }
  for (int64_t b = 0; b < m.n_bfaces(); ++b)
    if (g2lc[m.bface_cell[b]] >= 0) {
      P.mesh.bface_cell.push_back(g2lc[m.bface_cell[b]]);
      P.mesh.bface_local.push_back(m.bface_local[b]);
      P.mesh.bface_id.push_back(m.bface_id[b]);
    }
  build_field(m, dp, own_p, cell_ranks, local_cells, rank, P.field[0]);
  build_field(m, du, own_u, cell_ranks, local_cells, rank, P.field[1]);
  return P;
}

}  // namespace partition
