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
// `balanced` (the default; PE_BALANCED_OWNERSHIP=0 switches it off) deals the interface nodes out among the ranks that touch them — all components of
// a node stay together (block rows) — by a hash of the node number every rank computes alike.
inline bool balanced_ownership_requested() {  // default on (measured at 8 GPUs, 128^3: 69.2 -> 67.7 ms per step); PE_BALANCED_OWNERSHIP=0: lowest rank owns
  const char* e = std::getenv("PE_BALANCED_OWNERSHIP");
  return !(e && e[0] == '0');
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

// ---- structured fast path ------------------------------------------------------------------------------------------------
// make_part() above needs the GLOBAL mesh and both GLOBAL dof maps on every rank (at 286^2 x 1144 cells: ~6 GB of host
// memory and tens of seconds per rank, all ranks at once).  For boxes with FE_Q(1) for both fields everything it computes
// is a function of the lattice, so a rank can build its part from two streaming sweeps over the global CELL ORDER that keep
// only two small arrays per lattice vertex (its first-touch number and its owner) and store nothing else global:
//   sweep 1  first-touch vertex numbers in cell order (== dofs::distribute_dofs for FE_Q(1): the pressure dof of a vertex is
//            that number, its displacement dofs are dim*number + component) and the owner = rank of the numbering cell;
//   sweep 2  the cells that touch a vertex this rank owns (own cells + one ghost layer), with the set of ranks that also
//            keep each of them.
// The result is identical, array by array, to make_part() on the global mesh (tests/test_partition.py).
inline Part make_part_structured(int dim, const double* size, const int* n_axis, bool morton_order, int rank, int nranks) {
  if (nranks > 32) throw std::runtime_error("make_part_structured: at most 32 ranks");
  Part P;
  const int vpc = 1 << dim;
  const int nn[3] = {n_axis[0], n_axis[1], dim == 3 ? n_axis[2] : 1};
  const int nv[3] = {nn[0] + 1, nn[1] + 1, dim == 3 ? nn[2] + 1 : 1};
  const int64_t n_cells = (int64_t)nn[0] * nn[1] * nn[2], n_lat = (int64_t)nv[0] * nv[1] * nv[2];
  int level = 0;
  if (morton_order) {
    while ((1 << level) < nn[0]) ++level;
    for (int a = 0; a < dim; ++a)
      if (nn[a] != (1 << level)) throw std::runtime_error("mesh: Morton order needs 2^L cells on every axis");
  }
  auto cell_ijk = [&](int64_t c, int ijk[3]) {
    if (morton_order) mesh::morton_decode((uint64_t)c, dim, level, ijk);
    else {
      ijk[0] = (int)(c % nn[0]);
      ijk[1] = (int)((c / nn[0]) % nn[1]);
      ijk[2] = (int)(c / ((int64_t)nn[0] * nn[1]));
    }
  };
  auto lattice = [&](const int ijk[3], int v) {
    const int i = ijk[0] + (v & 1), j = ijk[1] + ((v >> 1) & 1), k = dim == 3 ? ijk[2] + ((v >> 2) & 1) : 0;
    return (int64_t)i + (int64_t)nv[0] * (j + (int64_t)nv[1] * k);
  };
  std::vector<int64_t> first_cell(nranks + 1);
  for (int r = 0; r <= nranks; ++r) first_cell[r] = (int64_t)((__int128)n_cells * r / nranks);
  // sweep 1
  std::vector<int32_t> vnum((size_t)n_lat, -1);
  std::vector<int8_t> vowner((size_t)n_lat, -1);
  const bool balanced = balanced_ownership_requested() && nranks > 1;
  std::vector<uint32_t> vtouch(balanced ? (size_t)n_lat : 0, 0u);  // ranks whose cells touch the vertex
  {
    int32_t next = 0;
    int r = 0;
    for (int64_t c = 0; c < n_cells; ++c) {
      while (c >= first_cell[r + 1]) ++r;
      int ijk[3];
      cell_ijk(c, ijk);
      for (int v = 0; v < vpc; ++v) {
        const int64_t lv = lattice(ijk, v);
        if (vnum[lv] < 0) {
          vnum[lv] = next++;
          vowner[lv] = (int8_t)r;
        }
        if (balanced) vtouch[lv] |= 1u << r;
      }
    }
  }
  if (balanced) {  // interface vertices are dealt out among the ranks that touch them: the same hash as dof_owner()
    for (int64_t lv = 0; lv < n_lat; ++lv) {
      const uint32_t m = vtouch[lv];
      if ((m & (m - 1)) == 0) continue;  // one toucher
      int t[32], n = 0;
      for (int r = 0; r < nranks; ++r)
        if (m & (1u << r)) t[n++] = r;   // ascending
      uint64_t h = (uint64_t)vnum[lv] * 0x9e3779b97f4a7c15ull;
      h ^= h >> 29;
      vowner[lv] = (int8_t)t[(h >> 8) % (uint64_t)n];
    }
    std::vector<uint32_t>().swap(vtouch);
  }
  // sweep 2
  std::vector<uint32_t> cell_mask;  // ranks that keep each local cell
  {
    int r = 0;
    for (int64_t c = 0; c < n_cells; ++c) {
      while (c >= first_cell[r + 1]) ++r;
      int ijk[3];
      cell_ijk(c, ijk);
      uint32_t mask = 0;
      for (int v = 0; v < vpc; ++v) mask |= 1u << vowner[lattice(ijk, v)];
      if (mask & (1u << rank)) {
        P.cell_global.push_back(c);
        cell_mask.push_back(mask);
        if (r == rank) P.n_owned_cells++;
      }
    }
  }
  const int64_t nlc = (int64_t)P.cell_global.size();
  // local sub-mesh: vertices in order of first appearance
  mesh::Mesh& M = P.mesh;
  M.dim = dim;
  M.morton = false;
  M.cell_vertices.resize((size_t)nlc * vpc);
  std::vector<int32_t> lat2l((size_t)n_lat, -1);
  std::vector<int64_t> lverts;  // lattice id of every local mesh vertex
  for (int64_t lc = 0; lc < nlc; ++lc) {
    int ijk[3];
    cell_ijk(P.cell_global[lc], ijk);
    for (int v = 0; v < vpc; ++v) {
      const int64_t lv = lattice(ijk, v);
      if (lat2l[lv] < 0) {
        lat2l[lv] = (int32_t)lverts.size();
        lverts.push_back(lv);
        int idx[3] = {(int)(lv % nv[0]), (int)((lv / nv[0]) % nv[1]), (int)(lv / ((int64_t)nv[0] * nv[1]))};
        for (int a = 0; a < dim; ++a) M.xyz.push_back(-0.5 * size[a] + size[a] * ((double)idx[a] / (double)nn[a]));
      }
      M.cell_vertices[lc * vpc + v] = lat2l[lv];
    }
    for (int a = 0; a < dim; ++a) {
      if (ijk[a] == 0) { M.bface_cell.push_back((int32_t)lc); M.bface_local.push_back((int8_t)(2 * a)); M.bface_id.push_back(2 * a); }
      if (ijk[a] == nn[a] - 1) { M.bface_cell.push_back((int32_t)lc); M.bface_local.push_back((int8_t)(2 * a + 1)); M.bface_id.push_back(2 * a + 1); }
    }
  }
  // vertex-level plan: [owned interior | owned boundary | ghosts by owner], each ascending in the global number
  const int64_t nlv = (int64_t)lverts.size();
  std::vector<uint8_t> boundary((size_t)nlv, 0);
  for (int64_t lc = 0; lc < nlc; ++lc) {
    bool has_ghost = false;
    for (int v = 0; v < vpc; ++v) has_ghost |= vowner[lverts[M.cell_vertices[lc * vpc + v]]] != rank;
    if (has_ghost)
      for (int v = 0; v < vpc; ++v) boundary[M.cell_vertices[lc * vpc + v]] = 1;
  }
  std::vector<int32_t> owned, ghost;  // local mesh vertex ids
  for (int64_t l = 0; l < nlv; ++l) (vowner[lverts[l]] == rank ? owned : ghost).push_back((int32_t)l);
  auto gnum = [&](int32_t l) { return vnum[lverts[l]]; };
  std::sort(owned.begin(), owned.end(), [&](int32_t a, int32_t b) { return boundary[a] != boundary[b] ? boundary[a] < boundary[b] : gnum(a) < gnum(b); });
  std::sort(ghost.begin(), ghost.end(), [&](int32_t a, int32_t b) {
    const int oa = vowner[lverts[a]], ob = vowner[lverts[b]];
    return oa != ob ? oa < ob : gnum(a) < gnum(b);
  });
  std::vector<int32_t> pos((size_t)nlv);  // local mesh vertex -> position in the plan
  for (size_t i = 0; i < owned.size(); ++i) pos[owned[i]] = (int32_t)i;
  for (size_t i = 0; i < ghost.size(); ++i) pos[ghost[i]] = (int32_t)(owned.size() + i);
  // send sets per neighbour (owned vertices of local cells another rank keeps too), receive counts per owner
  std::vector<std::vector<int32_t>> send_sets(nranks);
  for (int64_t lc = 0; lc < nlc; ++lc) {
    const uint32_t others = cell_mask[lc] & ~(1u << rank);
    if (!others) continue;
    for (int v = 0; v < vpc; ++v) {
      const int32_t l = M.cell_vertices[lc * vpc + v];
      if (vowner[lverts[l]] != rank) continue;
      for (int r = 0; r < nranks; ++r)
        if (others & (1u << r)) send_sets[r].push_back(l);
    }
  }
  std::vector<int64_t> recv_count(nranks, 0);
  for (int32_t l : ghost) recv_count[vowner[lverts[l]]]++;
  for (int f = 0; f < 2; ++f) {
    FieldPart& F = P.field[f];
    const int nc = f == 0 ? 1 : dim;
    F.n_owned = (int64_t)owned.size() * nc;
    F.n_local = (int64_t)nlv * nc;
    F.local_to_global.resize((size_t)F.n_local);
    for (size_t i = 0; i < owned.size(); ++i)
      for (int c = 0; c < nc; ++c) F.local_to_global[i * nc + c] = (int64_t)gnum(owned[i]) * nc + c;
    for (size_t i = 0; i < ghost.size(); ++i)
      for (int c = 0; c < nc; ++c) F.local_to_global[(owned.size() + i) * nc + c] = (int64_t)gnum(ghost[i]) * nc + c;
    F.cell_dofs.resize((size_t)nlc * vpc * nc);
    for (int64_t lc = 0; lc < nlc; ++lc)
      for (int v = 0; v < vpc; ++v)
        for (int c = 0; c < nc; ++c) F.cell_dofs[(lc * vpc + v) * nc + c] = pos[M.cell_vertices[lc * vpc + v]] * nc + c;
    F.neighbor_rank.clear();
    F.send_ptr.assign(1, 0);
    F.recv_ptr.assign(1, 0);
    F.send_idx.clear();
    for (int r = 0; r < nranks; ++r) {
      if (r == rank || (send_sets[r].empty() && recv_count[r] == 0)) continue;
      F.neighbor_rank.push_back(r);
      std::vector<int32_t> s = send_sets[r];
      std::sort(s.begin(), s.end(), [&](int32_t a, int32_t b) { return gnum(a) < gnum(b); });
      s.erase(std::unique(s.begin(), s.end()), s.end());
      for (int32_t l : s)
        for (int c = 0; c < nc; ++c) F.send_idx.push_back(pos[l] * nc + c);
      F.send_ptr.push_back((int64_t)F.send_idx.size());
      F.recv_ptr.push_back(F.recv_ptr.back() + recv_count[r] * nc);
    }
  }
  return P;
}

// global dof counts of the FE_Q(1) box (the structured path never builds the global maps)
inline int64_t structured_vertex_count(int dim, const int* n_axis) {
  int64_t n = 1;
  for (int a = 0; a < dim; ++a) n *= n_axis[a] + 1;
  return n;
}

}  // namespace partition
