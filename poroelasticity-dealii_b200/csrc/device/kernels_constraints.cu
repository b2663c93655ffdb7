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

#include "pe_internal.cuh"

namespace {

constexpr int T = 256;

// ---- pattern from variable-length lists --------------------------------------------------------------------------------
__global__ void count_adjacency_var(const int32_t* __restrict__ lptr, const int32_t* __restrict__ ldofs, int64_t n_lists, int64_t n_owned,
                                    int32_t* __restrict__ cnt) {
  const int64_t l = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (l >= n_lists) return;
  for (int e = lptr[l]; e < lptr[l + 1]; ++e) {
    const int32_t d = ldofs[e];
    if (d < n_owned) atomicAdd(&cnt[d], 1);
  }
}

__global__ void fill_adjacency_var(const int32_t* __restrict__ lptr, const int32_t* __restrict__ ldofs, int64_t n_lists, int64_t n_owned,
                                   const int32_t* __restrict__ adj_ptr, int32_t* __restrict__ fill, int32_t* __restrict__ adj) {
  const int64_t l = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (l >= n_lists) return;
  for (int e = lptr[l]; e < lptr[l + 1]; ++e) {
    const int32_t d = ldofs[e];
    if (d >= n_owned) continue;
    const int32_t k = atomicAdd(&fill[d], 1);
    adj[adj_ptr[d] + k] = (int32_t)l;
  }
}

constexpr int ROW_WARPS = 4;

// one warp per row: gather the lists that contain the row's dof, bitonic sort in shared memory, unique
template <bool WRITE>
__global__ void row_pattern_var(const int32_t* __restrict__ lptr, const int32_t* __restrict__ ldofs, int64_t n_owned,
                                const int32_t* __restrict__ adj_ptr, const int32_t* __restrict__ adj, int cap, int32_t* __restrict__ rowlen,
                                const int32_t* __restrict__ rowptr, int32_t* __restrict__ col, int* __restrict__ overflow) {
  extern __shared__ int32_t smem_rows[];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int32_t* buf = smem_rows + (size_t)w * cap;
  const int64_t row = (int64_t)blockIdx.x * ROW_WARPS + w;
  if (row >= n_owned) return;
  const int a0 = adj_ptr[row], a1 = adj_ptr[row + 1];
  int n_cand = 0;
  for (int a = a0; a < a1; ++a) {  // warp-uniform loop
    const int l = adj[a];
    const int s = lptr[l], len = lptr[l + 1] - s;
    if (n_cand + len > cap) {
      if (lane == 0) atomicExch(overflow, n_cand + len);
      return;
    }
    for (int i = lane; i < len; i += 32) buf[n_cand + i] = ldofs[s + i];
    n_cand += len;
  }
  int m = 32;
  while (m < n_cand) m <<= 1;
  for (int i = n_cand + lane; i < m; i += 32) buf[i] = 0x7fffffff;
  __syncwarp();
  for (int k = 2; k <= m; k <<= 1)
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = lane; i < m; i += 32) {
        const int ixj = i ^ j;
        if (ixj > i) {
          const int32_t a = buf[i], b = buf[ixj];
          const bool up = (i & k) == 0;
          if ((a > b) == up) { buf[i] = b; buf[ixj] = a; }
        }
      }
      __syncwarp();
    }
  int base = 0;
  for (int i0 = 0; i0 < n_cand; i0 += 32) {
    const int i = i0 + lane;
    const bool keep = i < n_cand && (i == 0 || buf[i] != buf[i - 1]);
    const unsigned mask = __ballot_sync(0xffffffffu, keep);
    if (WRITE && keep) col[rowptr[row] + base + __popc(mask & ((1u << lane) - 1))] = buf[i];
    base += __popc(mask);
  }
  if (!WRITE && lane == 0) rowlen[row] = base;
}

// ---- condense / distribute -----------------------------------------------------------------------------------------------
__device__ __forceinline__ int find_or_miss(const int32_t* __restrict__ col, int lo, int hi, int32_t target) {
  const int end = hi;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (col[mid] < target) lo = mid + 1; else hi = mid;
  }
  return (lo < end && col[lo] == target) ? lo : -1;
}

struct HangView {
  const int32_t* hline;     // dof -> hanging line or -1
  const int32_t* dof;       // line -> dof
  const int32_t* tline_of;  // dof -> slot in the transposed table or -1
  const int32_t* t_ptr;
  const int32_t* t_line;
  const double* t_w;
};

// one warp per destination row
__global__ void k_condense_matrix(int64_t n_owned, const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                                  const double* __restrict__ src, double* __restrict__ dst, HangView H, int keep_diag, double hang_diag) {
  const int lane = threadIdx.x & 31;
  const int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= n_owned) return;
  const int r0 = rowptr[r], r1 = rowptr[r + 1];
  if (H.hline[r] >= 0) {  // hanging row: only the diagonal survives
    for (int j = r0 + lane; j < r1; j += 32) dst[j] = (col[j] == (int32_t)r) ? (keep_diag ? src[j] : hang_diag) : 0.0;
    return;
  }
  const int tr = H.tline_of[r];
  const int nr = tr >= 0 ? H.t_ptr[tr + 1] - H.t_ptr[tr] : 0;
  for (int j = r0 + lane; j < r1; j += 32) {
    const int32_t c = col[j];
    if (H.hline[c] >= 0) { dst[j] = 0.0; continue; }
    const int tc = H.tline_of[c];
    const int ncol = tc >= 0 ? H.t_ptr[tc + 1] - H.t_ptr[tc] : 0;
    if (nr == 0 && ncol == 0) { dst[j] = src[j]; continue; }
    double sum = 0.0;
    for (int a = -1; a < nr; ++a) {  // a == -1: the row itself with weight 1
      const int32_t i = a < 0 ? (int32_t)r : H.dof[H.t_line[H.t_ptr[tr] + a]];
      const double wi = a < 0 ? 1.0 : H.t_w[H.t_ptr[tr] + a];
      const int i0 = rowptr[i], i1 = rowptr[i + 1];
      for (int b = -1; b < ncol; ++b) {
        const int32_t k = b < 0 ? c : H.dof[H.t_line[H.t_ptr[tc] + b]];
        const double wk = b < 0 ? 1.0 : H.t_w[H.t_ptr[tc] + b];
        const int pos = (a < 0 && b < 0) ? j : find_or_miss(col, i0, i1, k);
        if (pos >= 0) sum += wi * wk * src[pos];
      }
    }
    dst[j] = sum;
  }
}

__global__ void k_condense_vector_gather(int64_t n_masters, const int32_t* __restrict__ t_master, const int32_t* __restrict__ t_ptr,
                                         const int32_t* __restrict__ t_line, const double* __restrict__ t_w, const int32_t* __restrict__ dof,
                                         double* __restrict__ v) {
  const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_masters) return;
  double acc = v[t_master[s]];
  for (int k = t_ptr[s]; k < t_ptr[s + 1]; ++k) acc += t_w[k] * v[dof[t_line[k]]];
  v[t_master[s]] = acc;
}

__global__ void k_zero_lines(int64_t n, const int32_t* __restrict__ dof, double* __restrict__ v) {
  const int64_t l = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (l < n) v[dof[l]] = 0.0;
}

__global__ void k_distribute_hanging(int64_t n, const int32_t* __restrict__ dof, const int32_t* __restrict__ ptr,
                                     const int32_t* __restrict__ edof, const double* __restrict__ w, const double* __restrict__ g,
                                     double* __restrict__ v) {
  const int64_t l = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (l >= n) return;
  double s = g[l];
  for (int e = ptr[l]; e < ptr[l + 1]; ++e) s += w[e] * v[edof[e]];
  v[dof[l]] = s;
}

__global__ void k_scatter_lines(int64_t n, const int32_t* __restrict__ dof, const double* __restrict__ g, double* __restrict__ v) {
  const int64_t l = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (l < n) v[dof[l]] = g[l];
}

// sum of |diagonal| with a fixed order: one block, thread t adds rows t, t + 1024, ..., then a tree over the threads
__global__ void k_sum_abs_diag(int64_t n, const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col, const double* __restrict__ val,
                               double* __restrict__ out) {
  __shared__ double s[1024];
  double acc = 0.0;
  for (int64_t r = threadIdx.x; r < n; r += blockDim.x) {
    const int pos = find_or_miss(col, rowptr[r], rowptr[r + 1], (int32_t)r);
    if (pos >= 0) acc += fabs(val[pos]);
  }
  s[threadIdx.x] = acc;
  __syncthreads();
  for (int o = blockDim.x >> 1; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) s[threadIdx.x] += s[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) *out = s[0];
}

HangView view_of(const Field& F) {
  const Field::Hanging& G = F.hang;
  return HangView{G.hline.p, G.dof.p, G.tline_of.p, G.t_ptr.p, G.t_line.p, G.t_w.p};
}

}  // namespace

void pe_hanging_upload(pe_ctx* c, Field& F) {
  Field::Hanging& G = F.hang;
  cudaStream_t s = c->stream;
  std::vector<int32_t> hline((size_t)F.n_local, -1), tline_of((size_t)F.n_local, -1);
  for (int64_t l = 0; l < G.n; ++l) hline[G.h_dof[l]] = (int32_t)l;
  // transposed table: masters ascending, their lines ascending
  std::vector<int32_t> cnt((size_t)F.n_local, 0);
  for (int32_t m : G.h_edof) cnt[m]++;
  std::vector<int32_t> t_master, t_ptr(1, 0);
  for (int64_t d = 0; d < F.n_local; ++d)
    if (cnt[d]) {
      tline_of[d] = (int32_t)t_master.size();
      t_master.push_back((int32_t)d);
      t_ptr.push_back(t_ptr.back() + cnt[d]);
    }
  std::vector<int32_t> t_line((size_t)G.n_entries), pos(t_ptr.begin(), t_ptr.end() - 1);
  std::vector<double> t_w((size_t)G.n_entries);
  for (int64_t l = 0; l < G.n; ++l)
    for (int e = G.h_ptr[l]; e < G.h_ptr[l + 1]; ++e) {
      const int slot = tline_of[G.h_edof[e]];
      t_line[pos[slot]] = (int32_t)l;
      t_w[pos[slot]] = G.h_w[e];
      pos[slot]++;
    }
  G.n_masters = (int64_t)t_master.size();
  G.hline.upload(hline, s);
  G.tline_of.upload(tline_of, s);
  G.dof.upload(G.h_dof, s);
  G.ptr.upload(G.h_ptr, s);
  G.edof.upload(G.h_edof, s);
  G.w.upload(G.h_w, s);
  G.g.upload(G.h_g, s);
  G.t_master.upload(t_master, s);
  G.t_ptr.upload(t_ptr, s);
  G.t_line.upload(t_line, s);
  G.t_w.upload(t_w, s);
  PE_CUDA(cudaStreamSynchronize(s));
}

void pe_build_pattern_lists(pe_ctx* c, Field& F) {
  const Field::Hanging& G = F.hang;
  // host: per-cell lists = own dofs + masters of the cell's hanging dofs, whole nodes (all components) so that the
  // vector-valued matrix keeps its ncomp x ncomp block structure
  std::vector<int32_t> hline((size_t)F.n_local, -1);
  for (int64_t l = 0; l < G.n; ++l) hline[G.h_dof[l]] = (int32_t)l;
  std::vector<int32_t> lptr(1, 0), ldofs, tmp;
  std::vector<int64_t> weight((size_t)F.n_owned, 0);  // candidates per row -> shared-memory capacity
  for (int64_t cell = 0; cell < c->n_cells; ++cell) {
    const int32_t* cd = &F.h_cell_dofs[cell * F.nloc];
    tmp.assign(cd, cd + F.nloc);
    for (int k = 0; k < F.nloc; ++k) {
      const int32_t l = hline[cd[k]];
      if (l < 0) continue;
      for (int e = G.h_ptr[l]; e < G.h_ptr[l + 1]; ++e) {
        const int32_t node0 = G.h_edof[e] / F.ncomp * F.ncomp;
        for (int q = 0; q < F.ncomp; ++q) tmp.push_back(node0 + q);
      }
    }
    if ((int)tmp.size() > F.nloc) {
      std::sort(tmp.begin(), tmp.end());
      tmp.erase(std::unique(tmp.begin(), tmp.end()), tmp.end());
    }
    for (int32_t d : tmp)
      if (d < F.n_owned) weight[d] += (int64_t)tmp.size();
    ldofs.insert(ldofs.end(), tmp.begin(), tmp.end());
    if (ldofs.size() >= ((size_t)1 << 31)) throw PeError(PE_ERR_UNSUPPORTED, "pattern lists exceed 32-bit indexing on one rank");
    lptr.push_back((int32_t)ldofs.size());
  }
  int64_t max_cand = 0;
  for (int64_t w : weight) max_cand = std::max(max_cand, w);
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
