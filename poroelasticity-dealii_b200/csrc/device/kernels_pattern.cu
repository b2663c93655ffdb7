// kernels_pattern.cu — K0: CSR sparsity pattern built once on the device.
//
// Replaces DoFTools::make_sparsity_pattern + SparsityPattern::copy_from
// (lib/include/PoroElasticPressureSolver.h:80-88, PoroElasticDisplacementSolver.h:140-146): every
// pair of local dofs of a cell couples (no coupling mask is passed at DS:143-145).  Rows are the
// dofs this rank owns; columns are local ids, ascending.
//   1. dof -> cell adjacency by counting (integer atomics) + exclusive scan + fill;
//   2. one warp per row gathers the dofs of the adjacent cells into shared memory, sorts them with a
//      warp-level bitonic network and removes duplicates; pass 1 counts, pass 2 writes.
// The result does not depend on the (non-deterministic) fill order of step 1 because every row is
// sorted.
#include "pe_internal.cuh"

namespace {

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

__global__ void scan_tile_sums(const int32_t* __restrict__ in, int64_t n, int32_t* __restrict__ tile_sums) {
  __shared__ int32_t warp_sums[SCAN_THREADS / 32];
  int64_t base = (int64_t)blockIdx.x * SCAN_TILE;
  int32_t s = 0;
  for (int k = 0; k < SCAN_ITEMS; ++k) {
    int64_t i = base + (int64_t)k * SCAN_THREADS + threadIdx.x;
    if (i < n) s += in[i];
  }
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) warp_sums[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    int32_t t = 0;
    for (int w = 0; w < SCAN_THREADS / 32; ++w) t += warp_sums[w];
    tile_sums[blockIdx.x] = t;
  }
}

// single block: exclusive scan of the tile sums (sequential over chunks of blockDim)
__global__ void scan_tile_offsets(int32_t* __restrict__ tile_sums, int n_tiles, int32_t* __restrict__ total_out) {
  __shared__ int32_t buf[1024];
  __shared__ int32_t carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < n_tiles; base += 1024) {
    int i = base + threadIdx.x;
    int32_t v = i < n_tiles ? tile_sums[i] : 0;
    buf[threadIdx.x] = v;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {  // Hillis-Steele inclusive scan
      int32_t t = threadIdx.x >= o ? buf[threadIdx.x - o] : 0;
      __syncthreads();
      buf[threadIdx.x] += t;
      __syncthreads();
    }
    int32_t excl = buf[threadIdx.x] - v + carry;
    if (i < n_tiles) tile_sums[i] = excl;
    __syncthreads();
    if (threadIdx.x == 1023) carry += buf[1023];
    __syncthreads();
  }
  if (threadIdx.x == 0) *total_out = carry;
}

// exclusive scan inside each tile + tile offset; element order inside a tile: i = base + t*ITEMS + k
__global__ void scan_apply(int32_t* __restrict__ data, int64_t n, const int32_t* __restrict__ tile_off) {
  __shared__ int32_t warp_sums[SCAN_THREADS / 32];
  int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
  int32_t v[SCAN_ITEMS];
  int32_t s = 0;
  for (int k = 0; k < SCAN_ITEMS; ++k) {
    int64_t i = base + k;
    v[k] = i < n ? data[i] : 0;
    s += v[k];
  }
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int32_t incl = s;
  for (int o = 1; o < 32; o <<= 1) {
    int32_t t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) warp_sums[w] = incl;
  __syncthreads();
  int32_t woff = 0;
  for (int k = 0; k < w; ++k) woff += warp_sums[k];
  int32_t run = tile_off[blockIdx.x] + woff + incl - s;
  for (int k = 0; k < SCAN_ITEMS; ++k) {
    int64_t i = base + k;
    if (i < n) data[i] = run;
    run += v[k];
  }
}

__global__ void count_adjacency(const int32_t* __restrict__ cell_dofs, int64_t n_entries, int64_t n_owned, int32_t* __restrict__ cnt) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_entries) return;
  int32_t d = cell_dofs[i];
  if (d < n_owned) atomicAdd(&cnt[d], 1);
}

__global__ void fill_adjacency(const int32_t* __restrict__ cell_dofs, int64_t n_entries, int nloc, int64_t n_owned,
                               const int32_t* __restrict__ adj_ptr, int32_t* __restrict__ fill, int32_t* __restrict__ adj) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_entries) return;
  int32_t d = cell_dofs[i];
  if (d >= n_owned) return;
  int32_t k = atomicAdd(&fill[d], 1);
  adj[adj_ptr[d] + k] = (int32_t)(i / nloc);
}

constexpr int ROW_WARPS = 4;

// one warp per row; CAP = shared-memory slots per warp (power of two)
template <bool WRITE>
__global__ void row_pattern(const int32_t* __restrict__ cell_dofs, int nloc, int64_t n_owned, const int32_t* __restrict__ adj_ptr,
                            const int32_t* __restrict__ adj, int cap, int32_t* __restrict__ rowlen, const int32_t* __restrict__ rowptr,
                            int32_t* __restrict__ col, int* __restrict__ overflow) {
  extern __shared__ int32_t smem_rows[];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int32_t* buf = smem_rows + (size_t)w * cap;
  int64_t row = (int64_t)blockIdx.x * ROW_WARPS + w;
  if (row >= n_owned) return;
  const int a0 = adj_ptr[row], a1 = adj_ptr[row + 1];
  const int n_cand = (a1 - a0) * nloc;
  if (n_cand > cap) {
    if (lane == 0) atomicExch(overflow, n_cand);
    return;
  }
  int m = 32;
  while (m < n_cand) m <<= 1;
  for (int i = lane; i < m; i += 32) {
    int32_t v = 0x7fffffff;
    if (i < n_cand) v = cell_dofs[(int64_t)adj[a0 + i / nloc] * nloc + i % nloc];
    buf[i] = v;
  }
  __syncwarp();
  for (int k = 2; k <= m; k <<= 1)
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = lane; i < m; i += 32) {
        int ixj = i ^ j;
        if (ixj > i) {
          int32_t a = buf[i], b = buf[ixj];
          bool up = (i & k) == 0;
          if ((a > b) == up) { buf[i] = b; buf[ixj] = a; }
        }
      }
      __syncwarp();
    }
  // unique: element i is kept when it differs from its predecessor
  int base = 0;
  for (int i0 = 0; i0 < n_cand; i0 += 32) {
    int i = i0 + lane;
    bool keep = i < n_cand && (i == 0 || buf[i] != buf[i - 1]);
    unsigned mask = __ballot_sync(0xffffffffu, keep);
    if (WRITE && keep) col[rowptr[row] + base + __popc(mask & ((1u << lane) - 1))] = buf[i];
    base += __popc(mask);
  }
  if (!WRITE && lane == 0) rowlen[row] = base;
}

// first row that references a ghost column (col >= n_owned)
__global__ void first_ghost_row(int64_t n_owned, const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col, int* __restrict__ first) {
  int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_owned) return;
  // columns ascend: the last entry of the row is its largest column
  if (rowptr[r + 1] > rowptr[r] && col[rowptr[r + 1] - 1] >= n_owned) atomicMin(first, (int)r);
}

// CSR -> block CSR, one warp per block row.  `bad` is raised when a row triple does not share one block pattern.
template <int B>
__global__ void csr_to_bsr(int64_t n_brows, const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col, const double* __restrict__ val,
                           int32_t* __restrict__ bptr, int32_t* __restrict__ bcol, double* __restrict__ bval, int* __restrict__ bad) {
  const int lane = threadIdx.x & 31;
  const int64_t I = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (I >= n_brows) return;
  const int r0 = rowptr[I * B];
  const int len = rowptr[I * B + 1] - r0;
  if (len % B != 0 || r0 % (B * B) != 0) { if (lane == 0) atomicExch(bad, 1); return; }
  const int nb = len / B, base = r0 / (B * B);
  if (lane == 0) {
    bptr[I] = base;
    if (I == n_brows - 1) bptr[n_brows] = base + nb;
  }
  for (int j = lane; j < nb; j += 32) {
    const int c0 = col[r0 + j * B];
    if (c0 % B != 0) atomicExch(bad, 1);
    if (bval == nullptr) continue;
    bcol[base + j] = c0 / B;
    for (int r = 0; r < B; ++r) {
      const int rs = rowptr[I * B + r];
      if (rowptr[I * B + r + 1] - rs != len) { atomicExch(bad, 1); continue; }
      for (int cc = 0; cc < B; ++cc) {
        if (col[rs + j * B + cc] != c0 + cc) atomicExch(bad, 1);
        bval[(size_t)base * B * B + (size_t)(r * B + cc) * nb + j] = val[rs + j * B + cc];
      }
    }
  }
}

}  // namespace

bool pe_build_bsr(pe_ctx* c, Field& F, const double* csr_val) {
  const int B = F.ncomp;
  F.bsr.B = 0;
  if (B < 2 || B > 3 || F.n_owned % B != 0 || F.n_local % B != 0 || F.nnz % (B * B) != 0) return false;
  Field::Bsr& S = F.bsr;
  S.n_brows = F.n_owned / B;
  S.nnzb = F.nnz / (B * B);
  S.bptr.alloc((size_t)S.n_brows + 1);
  S.bcol.alloc((size_t)S.nnzb);
  S.bval.alloc((size_t)F.nnz);
  DBuf<int> bad;
  bad.alloc_zero(1, c->stream);
  const int warps = 8;
  const int grid = pe_div_up(S.n_brows, warps);
  if (B == 2) csr_to_bsr<2><<<grid, warps * 32, 0, c->stream>>>(S.n_brows, F.rowptr.p, F.col.p, csr_val, S.bptr.p, S.bcol.p, S.bval.p, bad.p);
  else csr_to_bsr<3><<<grid, warps * 32, 0, c->stream>>>(S.n_brows, F.rowptr.p, F.col.p, csr_val, S.bptr.p, S.bcol.p, S.bval.p, bad.p);
  int h_bad = 0;
  PE_CUDA(cudaMemcpyAsync(&h_bad, bad.p, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  PE_CUDA(cudaStreamSynchronize(c->stream));
  PE_CUDA(cudaGetLastError());
  c->st.kernel_launches++;
  if (h_bad) {
    S.bptr.release(); S.bcol.release(); S.bval.release();
    return false;
  }
  S.B = B;
  return true;
}

int64_t pe_exclusive_scan_i32(pe_ctx* c, int32_t* data, int64_t n) {
  // scans data[0..n) in place and writes the total to data[n]
  int n_tiles = pe_div_up(n, SCAN_TILE);
  if (n_tiles == 0) {
    PE_CUDA(cudaMemsetAsync(data, 0, sizeof(int32_t), c->stream));
    return 0;
  }
  DBuf<int32_t> tiles;
  tiles.alloc((size_t)n_tiles + 1);
  scan_tile_sums<<<n_tiles, SCAN_THREADS, 0, c->stream>>>(data, n, tiles.p);
  scan_tile_offsets<<<1, 1024, 0, c->stream>>>(tiles.p, n_tiles, tiles.p + n_tiles);
  scan_apply<<<n_tiles, SCAN_THREADS, 0, c->stream>>>(data, n, tiles.p);
  PE_CUDA(cudaMemcpyAsync(data + n, tiles.p + n_tiles, sizeof(int32_t), cudaMemcpyDeviceToDevice, c->stream));
  int32_t total = 0;
  PE_CUDA(cudaMemcpyAsync(&total, tiles.p + n_tiles, sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
  PE_CUDA(cudaStreamSynchronize(c->stream));
  PE_CUDA(cudaGetLastError());
  c->st.kernel_launches += 3;
  return total;
}

void pe_build_pattern(pe_ctx* c, Field& F) {
  const int64_t n_entries = c->n_cells * F.nloc;
  if (n_entries >= (int64_t)1 << 31) throw PeError(PE_ERR_UNSUPPORTED, "cell_dofs exceeds 32-bit indexing on one rank");
  DBuf<int32_t> adj_ptr, fill, adj;
  adj_ptr.alloc_zero((size_t)F.n_owned + 1, c->stream);
  const int T = 256;
  count_adjacency<<<pe_div_up(n_entries, T), T, 0, c->stream>>>(F.cell_dofs.p, n_entries, F.n_owned, adj_ptr.p);
  int64_t n_adj = pe_exclusive_scan_i32(c, adj_ptr.p, F.n_owned);
  fill.alloc_zero((size_t)F.n_owned, c->stream);
  adj.alloc((size_t)n_adj);
  fill_adjacency<<<pe_div_up(n_entries, T), T, 0, c->stream>>>(F.cell_dofs.p, n_entries, F.nloc, F.n_owned, adj_ptr.p, fill.p, adj.p);
  // max cells around a dof -> shared-memory capacity per row
  std::vector<int32_t> h_ptr((size_t)F.n_owned + 1);
  PE_CUDA(cudaMemcpyAsync(h_ptr.data(), adj_ptr.p, h_ptr.size() * sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
  PE_CUDA(cudaStreamSynchronize(c->stream));
  int max_adj = 0;
  for (int64_t i = 0; i < F.n_owned; ++i) max_adj = std::max(max_adj, h_ptr[i + 1] - h_ptr[i]);
  int cap = 32;
  while (cap < max_adj * F.nloc) cap <<= 1;
  size_t smem = (size_t)ROW_WARPS * cap * sizeof(int32_t);
  if (smem > 200 * 1024) throw PeError(PE_ERR_UNSUPPORTED, "a dof touches too many cells for the row-pattern kernel");
  PE_CUDA(cudaFuncSetAttribute(row_pattern<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  PE_CUDA(cudaFuncSetAttribute(row_pattern<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  DBuf<int> overflow;
  overflow.alloc_zero(1, c->stream);
  F.rowptr.alloc_zero((size_t)F.n_owned + 1, c->stream);
  int blocks = pe_div_up(F.n_owned, ROW_WARPS);
  row_pattern<false><<<blocks, ROW_WARPS * 32, smem, c->stream>>>(F.cell_dofs.p, F.nloc, F.n_owned, adj_ptr.p, adj.p, cap, F.rowptr.p, nullptr,
                                                                   nullptr, overflow.p);
  int64_t nnz = pe_exclusive_scan_i32(c, F.rowptr.p, F.n_owned);
  // a wrapped 32-bit total shows up as a negative number
  if (nnz < 0) throw PeError(PE_ERR_UNSUPPORTED, "nnz exceeds 32-bit indexing on one rank");
  F.nnz = nnz;
  F.col.alloc((size_t)nnz);
  row_pattern<true><<<blocks, ROW_WARPS * 32, smem, c->stream>>>(F.cell_dofs.p, F.nloc, F.n_owned, adj_ptr.p, adj.p, cap, nullptr, F.rowptr.p,
                                                                  F.col.p, overflow.p);
  int h_over = 0;
  PE_CUDA(cudaMemcpyAsync(&h_over, overflow.p, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  PE_CUDA(cudaStreamSynchronize(c->stream));
  PE_CUDA(cudaGetLastError());
  if (h_over) throw PeError(PE_ERR_UNSUPPORTED, "row-pattern shared-memory capacity exceeded");
  c->st.kernel_launches += 4;
  // rows before the first one with a ghost column can be multiplied before the halo has arrived
  F.n_interior = F.n_owned;
  if (F.n_local > F.n_owned) {
    DBuf<int> first;
    int init = (int)F.n_owned;
    first.upload(&init, 1, c->stream);
    first_ghost_row<<<pe_div_up(F.n_owned, T), T, 0, c->stream>>>(F.n_owned, F.rowptr.p, F.col.p, first.p);
    PE_CUDA(cudaMemcpyAsync(&init, first.p, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    PE_CUDA(cudaStreamSynchronize(c->stream));
    const int unit = 32 * F.ncomp;  // whole warp blocks of 32 rows (CSR) and of 32 block rows (block CSR)
    F.n_interior = (int64_t)(init / unit) * unit;
    c->st.kernel_launches++;
  }
}
