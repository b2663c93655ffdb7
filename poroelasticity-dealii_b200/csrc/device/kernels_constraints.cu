// kernels_constraints.cu — hanging-node constraints of adaptive meshes on the device.
//
// deal.II applies the reference's ConstraintMatrix cell by cell (distribute_local_to_global, DS:279-286, SP:193-194) and
// to assembled objects (condense, PS:153, PS:168, SP:105; distribute, PS:180, DS:306, SP:215).  Here the cell kernels of
// kernels_assembly.cu stay exactly as they are on uniform meshes — hanging dofs are assembled like free ones — and the
// constraints are applied to the ASSEMBLED matrix / vectors by gather kernels with a fixed summation order:
//
//     A^[r, c] = sum_{i in T(r)} sum_{k in T(c)} w_i w_k A[i, k]      (r, c not hanging),   T(r) = {(r, 1)} + {(h, w_hr)}
//     A^[h, h] = diagonal kept (elasticity: sum of the cells' |a_hh|) or the average |diagonal| (ConstraintMatrix::condense)
//     b^[r]    = b[r] + sum_{(h, w) in T(r)} w b[h] ,  b^[h] = 0 ;      x[h] = sum_m w_hm x[m] + g_h   (distribute)
//
// which equals deal.II's result in exact arithmetic (both are E^T A E with E the expansion matrix of the closed
// constraint table) and is bitwise reproducible run to run: no atomics touch floating-point data.
//
// The sparsity pattern must hold the condensed entries as well (DoFTools::make_sparsity_pattern(dh, dsp, constraints,
// keep_constrained_dofs = true), PS:80-88, DS:140-146): it is built from per-cell dof LISTS — the cell's own dofs plus the
// masters of its hanging dofs — by variable-length variants of the kernels in kernels_pattern.cu.
#include <algorithm>

#include "constraint_tables.hpp"
#include "constraints_dev.cuh"
#include "pe_internal.cuh"

namespace {

constexpr int T = 256;
// ---- all device code lives in constraints_dev.cuh (shared with the CPU emulation harness of the tests) ----------------
using namespace pe_constraints_dev;

HangView view_of(const Field& F) {
  const Field::Hanging& G = F.hang;
  return HangView{G.hline.p, G.dof.p, G.tline_of.p, G.t_ptr.p, G.t_line.p, G.t_w.p};
}

}  // namespace

void pe_hanging_upload(pe_ctx* c, Field& F) {
  Field::Hanging& G = F.hang;
  cudaStream_t s = c->stream;
  pe_constraint_tables::Transposed TT = pe_constraint_tables::transpose(F.n_local, G.h_dof, G.h_ptr, G.h_edof, G.h_w);
  G.n_masters = (int64_t)TT.t_master.size();
  G.hline.upload(TT.hline, s);
  G.tline_of.upload(TT.tline_of, s);
  G.dof.upload(G.h_dof, s);
  G.ptr.upload(G.h_ptr, s);
  G.edof.upload(G.h_edof, s);
  G.w.upload(G.h_w, s);
  G.g.upload(G.h_g, s);
  G.t_master.upload(TT.t_master, s);
  G.t_ptr.upload(TT.t_ptr, s);
  G.t_line.upload(TT.t_line, s);
  G.t_w.upload(TT.t_w, s);
  PE_CUDA(cudaStreamSynchronize(s));
}

void pe_build_pattern_lists(pe_ctx* c, Field& F) {
  const Field::Hanging& G = F.hang;
  pe_constraint_tables::PatternLists PL =
      pe_constraint_tables::pattern_lists(c->n_cells, F.nloc, F.ncomp, F.n_owned, F.n_local, F.h_cell_dofs.data(), G.h_dof, G.h_ptr, G.h_edof);
  if (PL.overflow) throw PeError(PE_ERR_UNSUPPORTED, "pattern lists exceed 32-bit indexing on one rank");
  const std::vector<int32_t>&lptr = PL.lptr, &ldofs = PL.ldofs;
  const int64_t max_cand = PL.max_candidates;
  int cap = 32;
  while (cap < max_cand) cap <<= 1;
  const size_t smem = (size_t)ROW_WARPS * cap * sizeof(int32_t);
  if (smem > 200 * 1024) throw PeError(PE_ERR_UNSUPPORTED, "a dof couples to too many others for the row-pattern kernel");
  cudaStream_t s = c->stream;
  const int64_t n_lists = c->n_cells;
  DBuf<int32_t> d_lptr, d_ldofs, adj_ptr, fill, adj;
  d_lptr.upload(lptr, s);
  d_ldofs.upload(ldofs, s);
  adj_ptr.alloc_zero((size_t)F.n_owned + 1, s);
  count_adjacency_var<<<pe_div_up(n_lists, T), T, 0, s>>>(d_lptr.p, d_ldofs.p, n_lists, F.n_owned, adj_ptr.p);
  const int64_t n_adj = pe_exclusive_scan_i32(c, adj_ptr.p, F.n_owned);
  fill.alloc_zero((size_t)F.n_owned, s);
  adj.alloc((size_t)n_adj);
  fill_adjacency_var<<<pe_div_up(n_lists, T), T, 0, s>>>(d_lptr.p, d_ldofs.p, n_lists, F.n_owned, adj_ptr.p, fill.p, adj.p);
  PE_CUDA(cudaFuncSetAttribute(row_pattern_var<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  PE_CUDA(cudaFuncSetAttribute(row_pattern_var<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  DBuf<int> overflow;
  overflow.alloc_zero(1, s);
  F.rowptr.alloc_zero((size_t)F.n_owned + 1, s);
  const int blocks = pe_div_up(F.n_owned, ROW_WARPS);
  row_pattern_var<false><<<blocks, ROW_WARPS * 32, smem, s>>>(d_lptr.p, d_ldofs.p, F.n_owned, adj_ptr.p, adj.p, cap, F.rowptr.p, nullptr, nullptr,
                                                               overflow.p);
  const int64_t nnz = pe_exclusive_scan_i32(c, F.rowptr.p, F.n_owned);
  if (nnz < 0) throw PeError(PE_ERR_UNSUPPORTED, "nnz exceeds 32-bit indexing on one rank");
  F.nnz = nnz;
  F.col.alloc((size_t)nnz);
  row_pattern_var<true><<<blocks, ROW_WARPS * 32, smem, s>>>(d_lptr.p, d_ldofs.p, F.n_owned, adj_ptr.p, adj.p, cap, nullptr, F.rowptr.p, F.col.p,
                                                              overflow.p);
  int h_over = 0;
  PE_CUDA(cudaMemcpyAsync(&h_over, overflow.p, sizeof(int), cudaMemcpyDeviceToHost, s));
  PE_CUDA(cudaStreamSynchronize(s));
  PE_CUDA(cudaGetLastError());
  if (h_over) throw PeError(PE_ERR_UNSUPPORTED, "row-pattern shared-memory capacity exceeded");
  c->st.kernel_launches += 4;
  F.n_interior = F.n_owned;  // hanging-node meshes run on one rank
}

void pe_condense_matrix(pe_ctx* c, Field& F, const double* src, double* dst, bool keep_diag, double hang_diag) {
  if (src == dst) throw PeError(PE_ERR_STATE, "condense needs distinct source and destination arrays");
  const int warps = 8;
  k_condense_matrix<<<pe_div_up(F.n_owned, warps), warps * 32, 0, c->stream>>>(F.n_owned, F.rowptr.p, F.col.p, src, dst, view_of(F),
                                                                                keep_diag ? 1 : 0, hang_diag);
  c->st.kernel_launches++;
  PE_CUDA(cudaGetLastError());
}

void pe_condense_vector(pe_ctx* c, Field& F, double* v) {
  const Field::Hanging& G = F.hang;
  if (!G.n) return;
  if (G.n_masters)
    k_condense_vector_gather<<<pe_div_up(G.n_masters, T), T, 0, c->stream>>>(G.n_masters, G.t_master.p, G.t_ptr.p, G.t_line.p, G.t_w.p, G.dof.p, v);
  k_zero_lines<<<pe_div_up(G.n, T), T, 0, c->stream>>>(G.n, G.dof.p, v);
  c->st.kernel_launches += 2;
}

void pe_distribute_hanging(pe_ctx* c, Field& F, double* v) {
  const Field::Hanging& G = F.hang;
  if (!G.n) return;
  k_distribute_hanging<<<pe_div_up(G.n, T), T, 0, c->stream>>>(G.n, G.dof.p, G.ptr.p, G.edof.p, G.w.p, G.g.p, v);
  c->st.kernel_launches++;
}

void pe_scatter_hanging_inhomogeneity(pe_ctx* c, Field& F, double* v) {
  const Field::Hanging& G = F.hang;
  PE_CUDA(cudaMemsetAsync(v, 0, F.n_local * sizeof(double), c->stream));
  if (!G.n) return;
  k_scatter_lines<<<pe_div_up(G.n, T), T, 0, c->stream>>>(G.n, G.dof.p, G.g.p, v);
  c->st.kernel_launches++;
}

double pe_avg_abs_diag(pe_ctx* c, Field& F, const double* val) {
  double* out = c->red.out.p + PE_RED_SLOTS;
  k_sum_abs_diag<<<1, 1024, 0, c->stream>>>(F.n_owned, F.rowptr.p, F.col.p, val, out);
  c->st.kernel_launches++;
  double h = 0;
  PE_CUDA(cudaMemcpyAsync(&h, out, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  PE_CUDA(cudaStreamSynchronize(c->stream));
  PE_CUDA(cudaGetLastError());
  return h / (double)F.n_owned;
}
