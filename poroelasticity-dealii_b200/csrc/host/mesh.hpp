// mesh.hpp — host-side mesh surface of the reference without deal.II.
//
//  * create_hyper_rectangle(): GridGenerator::hyper_rectangle(colorize=true) + refine_global(L)
//    as used by PoroElasticProblem::create_mesh (lib/include/PoroelasticityFSS.h:418-435):
//    2^L cells per axis, active cells in Morton (Z) order with x the least-significant bit,
//    boundary id of a face = its face number (0/1 x-min/x-max, 2/3 y, 4/5 z).
//  * create_subdivided(): arbitrary cells-per-axis box, lexicographic cell order (needed for the
//    non-power-of-two weak-scaling blocks; stated deviation — refine_global cannot make them).
//  * read_msh(): Gmsh 2.2 ASCII reader for the subset GridIn::read_msh needs for `domain.msh`
//    (PoroelasticityFSS.h:438-445): nodes, line/quad (2D) or quad/hex (3D) elements, boundary id
//    = first tag (physical id) of the boundary element.
// Cell vertex order is deal.II's lexicographic one (x fastest).
#pragma once
#include <algorithm>
#include <array>
#include <cstdint>
#include <fstream>
#include <map>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace mesh {

struct Mesh {
  int dim = 0;
  std::vector<double> xyz;             // n_vertices * dim
  std::vector<int32_t> cell_vertices;  // n_cells * 2^dim
  std::vector<int32_t> bface_cell;     // boundary faces
  std::vector<int8_t> bface_local;
  std::vector<int32_t> bface_id;
  // structured meta data (0 when unstructured): cells per axis and whether cells are in Morton order
  int n_axis[3] = {0, 0, 0};
  bool morton = false;
  int64_t n_vertices() const { return dim ? (int64_t)xyz.size() / dim : 0; }
  int vpc() const { return 1 << dim; }
  int64_t n_cells() const { return dim ? (int64_t)cell_vertices.size() / vpc() : 0; }
  int64_t n_bfaces() const { return (int64_t)bface_cell.size(); }
};

// vertices of face f of the reference cell, in the face's own lexicographic order (GeometryInfo)
inline void face_vertices(int dim, int f, int out[4]) {
  int axis = f / 2, side = f % 2, k = 0;
  for (int v = 0; v < (1 << dim); ++v)
    if (((v >> axis) & 1) == side) out[k++] = v;
}

inline void morton_decode(uint64_t m, int dim, int level, int ijk[3]) {
  ijk[0] = ijk[1] = ijk[2] = 0;
  for (int b = 0; b < level; ++b)
    for (int a = 0; a < dim; ++a) ijk[a] |= (int)((m >> (b * dim + a)) & 1u) << b;
}

inline Mesh create_box(int dim, const double* size, const int* n, bool morton_order) {
  if (dim < 2 || dim > 3) throw std::runtime_error("mesh: dim must be 2 or 3");
  Mesh m;
  m.dim = dim;
  m.morton = morton_order;
  int nn[3] = {n[0], n[1], dim == 3 ? n[2] : 1};
  for (int a = 0; a < 3; ++a) m.n_axis[a] = a < dim ? nn[a] : 0;
  int nv[3] = {nn[0] + 1, nn[1] + 1, dim == 3 ? nn[2] + 1 : 1};
  int64_t n_vert = (int64_t)nv[0] * nv[1] * nv[2];
  m.xyz.resize(n_vert * dim);
  for (int k = 0; k < nv[2]; ++k)
    for (int j = 0; j < nv[1]; ++j)
      for (int i = 0; i < nv[0]; ++i) {
        int64_t v = i + (int64_t)nv[0] * (j + (int64_t)nv[1] * k);
        int idx[3] = {i, j, k};
        for (int a = 0; a < dim; ++a)  // hyper_rectangle spans [-size/2, size/2] (FSS:423-428)
          m.xyz[v * dim + a] = -0.5 * size[a] + size[a] * ((double)idx[a] / (double)nn[a]);
      }
  int64_t n_cells = (int64_t)nn[0] * nn[1] * nn[2];
  int vpc = 1 << dim;
  m.cell_vertices.resize(n_cells * vpc);
  int level = 0;
  if (morton_order) {
    while ((1 << level) < nn[0]) ++level;
    for (int a = 0; a < dim; ++a)
      if (nn[a] != (1 << level)) throw std::runtime_error("mesh: Morton order needs 2^L cells on every axis");
  }
  for (int64_t c = 0; c < n_cells; ++c) {
    int ijk[3];
    if (morton_order)
      morton_decode((uint64_t)c, dim, level, ijk);
    else {
      ijk[0] = (int)(c % nn[0]);
      ijk[1] = (int)((c / nn[0]) % nn[1]);
      ijk[2] = (int)(c / ((int64_t)nn[0] * nn[1]));
    }
    for (int v = 0; v < vpc; ++v) {
      int i = ijk[0] + (v & 1), j = ijk[1] + ((v >> 1) & 1), k = dim == 3 ? ijk[2] + ((v >> 2) & 1) : 0;
      m.cell_vertices[c * vpc + v] = (int32_t)(i + (int64_t)nv[0] * (j + (int64_t)nv[1] * k));
    }
    for (int a = 0; a < dim; ++a) {
      if (ijk[a] == 0) { m.bface_cell.push_back((int32_t)c); m.bface_local.push_back((int8_t)(2 * a)); m.bface_id.push_back(2 * a); }
      if (ijk[a] == nn[a] - 1) { m.bface_cell.push_back((int32_t)c); m.bface_local.push_back((int8_t)(2 * a + 1)); m.bface_id.push_back(2 * a + 1); }
    }
  }
  return m;
}

// FSS:418-435
inline Mesh create_hyper_rectangle(int dim, const double* size, int refine_level) {
  int n[3] = {1 << refine_level, 1 << refine_level, 1 << refine_level};
  return create_box(dim, size, n, /*morton=*/true);
}
inline Mesh create_subdivided(int dim, const double* size, const int* n) { return create_box(dim, size, n, false); }

// FSS:438-445 (GridIn::read_msh, format 2.x ASCII)
inline Mesh read_msh(const std::string& path, int dim) {
  std::ifstream f(path);
  if (!f) throw std::runtime_error("cannot open mesh file " + path);
  Mesh m;
  m.dim = dim;
  std::string line;
  std::map<int64_t, int32_t> node_index;  // gmsh node number -> 0-based
  std::vector<std::array<double, 3>> nodes;
  struct Elem { int type, tag; std::vector<int32_t> v; };
  std::vector<Elem> cells, bnd;
  const int cell_type = dim == 2 ? 3 : 5, bnd_type = dim == 2 ? 1 : 3;
  const int cell_nv = 1 << dim, bnd_nv = 1 << (dim - 1);
  while (std::getline(f, line)) {
    if (line.rfind("$MeshFormat", 0) == 0) {
      double ver; int ftype, dsize;
      f >> ver >> ftype >> dsize;
      if (ver < 2.0 || ver >= 3.0 || ftype != 0) throw std::runtime_error("read_msh: only Gmsh 2.x ASCII is supported");
    } else if (line.rfind("$Nodes", 0) == 0) {
      int64_t n; f >> n;
      nodes.resize(n);
      for (int64_t i = 0; i < n; ++i) {
        int64_t id; f >> id >> nodes[i][0] >> nodes[i][1] >> nodes[i][2];
        node_index[id] = (int32_t)i;
      }
    } else if (line.rfind("$Elements", 0) == 0) {
      int64_t n; f >> n;
      for (int64_t i = 0; i < n; ++i) {
        int64_t id; int type, ntags;
        f >> id >> type >> ntags;
        int tag0 = 0;
        for (int t = 0; t < ntags; ++t) { int tg; f >> tg; if (t == 0) tag0 = tg; }
        static const int nodes_per_type[16] = {0, 2, 3, 4, 4, 8, 6, 5, 3, 6, 9, 10, 27, 18, 14, 1};
        if (type < 1 || type > 15) throw std::runtime_error("read_msh: unsupported element type");
        int nv = nodes_per_type[type];
        Elem e; e.type = type; e.tag = tag0; e.v.resize(nv);
        for (int k = 0; k < nv; ++k) { int64_t g; f >> g; e.v[k] = node_index.at(g); }
        if (type == cell_type) cells.push_back(e);
        else if (type == bnd_type) bnd.push_back(e);
      }
    }
  }
  if (cells.empty()) throw std::runtime_error("read_msh: no cells of the requested dimension");
  // keep only vertices used by cells (GridTools::delete_unused_vertices)
  std::vector<int32_t> remap(nodes.size(), -1);
  int32_t nv_used = 0;
  for (auto& e : cells) for (auto v : e.v) if (remap[v] < 0) remap[v] = 0;
  for (size_t i = 0; i < nodes.size(); ++i) if (remap[i] == 0) remap[i] = nv_used++;
  m.xyz.resize((size_t)nv_used * dim);
  for (size_t i = 0; i < nodes.size(); ++i)
    if (remap[i] >= 0) for (int a = 0; a < dim; ++a) m.xyz[(size_t)remap[i] * dim + a] = nodes[i][a];
  // gmsh (counter-clockwise / bottom-then-top) -> lexicographic
  static const int perm2[4] = {0, 1, 3, 2}, perm3[8] = {0, 1, 3, 2, 4, 5, 7, 6};
  m.cell_vertices.resize(cells.size() * cell_nv);
  for (size_t c = 0; c < cells.size(); ++c)
    for (int v = 0; v < cell_nv; ++v)
      m.cell_vertices[c * cell_nv + v] = remap[cells[c].v[dim == 2 ? perm2[v] : perm3[v]]];
  // boundary faces: faces that belong to exactly one cell; id from the boundary element (default 0)
  std::map<std::vector<int32_t>, int> bnd_id;
  for (auto& e : bnd) {
    std::vector<int32_t> key;
    for (auto v : e.v) key.push_back(remap[v]);
    std::sort(key.begin(), key.end());
    bnd_id[key] = e.tag;
  }
  std::map<std::vector<int32_t>, std::pair<int, std::pair<int32_t, int>>> faces;  // key -> (count, (cell, face))
  for (size_t c = 0; c < cells.size(); ++c)
    for (int fc = 0; fc < 2 * dim; ++fc) {
      int fv[4];
      face_vertices(dim, fc, fv);
      std::vector<int32_t> key;
      for (int k = 0; k < bnd_nv; ++k) key.push_back(m.cell_vertices[c * cell_nv + fv[k]]);
      std::sort(key.begin(), key.end());
      auto& e = faces[key];
      if (e.first++ == 0) e.second = {(int32_t)c, fc};
    }
  std::vector<std::array<int32_t, 3>> bf;
  for (auto& kv : faces)
    if (kv.second.first == 1) {
      auto it = bnd_id.find(kv.first);
      bf.push_back({kv.second.second.first, kv.second.second.second, it == bnd_id.end() ? 0 : it->second});
    }
  std::sort(bf.begin(), bf.end());
  for (auto& b : bf) { m.bface_cell.push_back(b[0]); m.bface_local.push_back((int8_t)b[1]); m.bface_id.push_back(b[2]); }
  return m;
}

// Space-filling-curve cell order for unstructured meshes: cells are sorted by the Morton key of
// their centroid inside the mesh's bounding box, so that the contiguous cell ranges partition.hpp hands to the ranks are
// compact subdomains (the same idea as the Z-order of refine_global() on the box) and neighbouring cells are close in
// memory.  Boundary faces follow their cells.  Returns the permutation new -> old.
inline std::vector<int64_t> reorder_cells_sfc(Mesh& m) {
  const int dim = m.dim, vpc = m.vpc();
  const int64_t nc = m.n_cells();
  double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
  std::vector<double> ctr((size_t)nc * dim, 0.0);
  for (int64_t c = 0; c < nc; ++c)
    for (int a = 0; a < dim; ++a) {
      double s = 0;
      for (int v = 0; v < vpc; ++v) s += m.xyz[(int64_t)m.cell_vertices[c * vpc + v] * dim + a];
      s /= vpc;
      ctr[c * dim + a] = s;
      lo[a] = std::min(lo[a], s);
      hi[a] = std::max(hi[a], s);
    }
  const int bits = dim == 2 ? 31 : 21;
  double ext = 0;
  for (int a = 0; a < dim; ++a) ext = std::max(ext, hi[a] - lo[a]);
  if (ext <= 0) ext = 1;
  std::vector<std::pair<uint64_t, int64_t>> keyed((size_t)nc);
  for (int64_t c = 0; c < nc; ++c) {
    uint64_t key = 0;
    for (int a = 0; a < dim; ++a) {
      // one cube for all axes keeps the curve's cells isotropic
      const double t = (ctr[c * dim + a] - lo[a]) / ext;
      uint64_t q = (uint64_t)(t * (double)(((uint64_t)1 << bits) - 1));
      for (int b = 0; b < bits; ++b) key |= ((q >> b) & 1u) << (b * dim + a);
    }
    keyed[c] = {key, c};
  }
  std::sort(keyed.begin(), keyed.end());
  std::vector<int64_t> perm((size_t)nc), inv((size_t)nc);
  for (int64_t i = 0; i < nc; ++i) { perm[i] = keyed[i].second; inv[keyed[i].second] = i; }
  std::vector<int32_t> cv((size_t)nc * vpc);
  for (int64_t i = 0; i < nc; ++i)
    for (int v = 0; v < vpc; ++v) cv[i * vpc + v] = m.cell_vertices[perm[i] * vpc + v];
  m.cell_vertices.swap(cv);
  std::vector<std::array<int32_t, 3>> bf((size_t)m.n_bfaces());
  for (int64_t b = 0; b < m.n_bfaces(); ++b) bf[b] = {(int32_t)inv[m.bface_cell[b]], (int32_t)m.bface_local[b], m.bface_id[b]};
  std::sort(bf.begin(), bf.end());
  for (int64_t b = 0; b < m.n_bfaces(); ++b) { m.bface_cell[b] = bf[b][0]; m.bface_local[b] = (int8_t)bf[b][1]; m.bface_id[b] = bf[b][2]; }
  m.morton = false;
  return perm;
}

}  // namespace mesh
