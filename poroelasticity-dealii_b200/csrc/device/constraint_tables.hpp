// constraint_tables.hpp — host-side preparation of the hanging-node tables the device kernels read (plain C++, no CUDA):
// the transposed constraint table (master -> hanging lines) and the per-cell dof lists of the constraint-aware sparsity
// pattern.  Shared by kernels_constraints.cu and by the CPU emulation harness tests/emu_constraints.cpp.
#pragma once
#include <algorithm>
#include <cstdint>
#include <vector>

namespace pe_constraint_tables {

struct Transposed {
  std::vector<int32_t> hline;     // dof -> hanging line or -1
  std::vector<int32_t> tline_of;  // dof -> slot in t_master or -1
  std::vector<int32_t> t_master, t_ptr, t_line;
  std::vector<double> t_w;
};

// lines: dof[l], entries ptr[l]..ptr[l+1] of (edof, w); masters ascending, their lines ascending
inline Transposed transpose(int64_t n_local, const std::vector<int32_t>& dof, const std::vector<int32_t>& ptr, const std::vector<int32_t>& edof,
                            const std::vector<double>& w) {
  Transposed T;
  const int64_t n = (int64_t)dof.size();
  T.hline.assign((size_t)n_local, -1);
  T.tline_of.assign((size_t)n_local, -1);
  for (int64_t l = 0; l < n; ++l) T.hline[dof[l]] = (int32_t)l;
  std::vector<int32_t> cnt((size_t)n_local, 0);
  for (int32_t m : edof) cnt[m]++;
  T.t_ptr.assign(1, 0);
  for (int64_t d = 0; d < n_local; ++d)
    if (cnt[d]) {
      T.tline_of[d] = (int32_t)T.t_master.size();
      T.t_master.push_back((int32_t)d);
      T.t_ptr.push_back(T.t_ptr.back() + cnt[d]);
    }
  T.t_line.resize(edof.size());
  T.t_w.resize(edof.size());
  std::vector<int32_t> pos(T.t_ptr.begin(), T.t_ptr.end() - 1);
  for (int64_t l = 0; l < n; ++l)
    for (int e = ptr[l]; e < ptr[l + 1]; ++e) {
      const int slot = T.tline_of[edof[e]];
      T.t_line[pos[slot]] = (int32_t)l;
      T.t_w[pos[slot]] = w[e];
      pos[slot]++;
    }
  return T;
}

struct PatternLists {
  std::vector<int32_t> lptr, ldofs;
  int64_t max_candidates = 0;  // largest number of list entries gathered for one owned row
  bool overflow = false;       // lists exceed 32-bit indexing
};

// per-cell lists = own dofs + masters of the cell's hanging dofs, whole nodes (all components), so that the vector-valued
// matrix keeps its ncomp x ncomp block structure
inline PatternLists pattern_lists(int64_t n_cells, int nloc, int ncomp, int64_t n_owned, int64_t n_local, const int32_t* cell_dofs,
                                  const std::vector<int32_t>& dof, const std::vector<int32_t>& ptr, const std::vector<int32_t>& edof) {
  PatternLists P;
  std::vector<int32_t> hline((size_t)n_local, -1);
  for (size_t l = 0; l < dof.size(); ++l) hline[dof[l]] = (int32_t)l;
  std::vector<int64_t> weight((size_t)n_owned, 0);
  std::vector<int32_t> tmp;
  P.lptr.assign(1, 0);
  for (int64_t cell = 0; cell < n_cells; ++cell) {
    const int32_t* cd = cell_dofs + cell * nloc;
    tmp.assign(cd, cd + nloc);
    for (int k = 0; k < nloc; ++k) {
      const int32_t l = hline[cd[k]];
      if (l < 0) continue;
      for (int e = ptr[l]; e < ptr[l + 1]; ++e) {
        const int32_t node0 = edof[e] / ncomp * ncomp;
        for (int q = 0; q < ncomp; ++q) tmp.push_back(node0 + q);
      }
    }
    if ((int)tmp.size() > nloc) {
      std::sort(tmp.begin(), tmp.end());
      tmp.erase(std::unique(tmp.begin(), tmp.end()), tmp.end());
    }
    for (int32_t d : tmp)
      if (d < n_owned) weight[d] += (int64_t)tmp.size();
    P.ldofs.insert(P.ldofs.end(), tmp.begin(), tmp.end());
    if (P.ldofs.size() >= ((size_t)1 << 31)) { P.overflow = true; return P; }
    P.lptr.push_back((int32_t)P.ldofs.size());
  }
  for (int64_t x : weight) P.max_candidates = std::max(P.max_candidates, x);
  return P;
}

}  // namespace pe_constraint_tables
