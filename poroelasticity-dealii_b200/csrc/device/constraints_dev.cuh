// constraints_dev.cuh — device code of the hanging-node constraint kernels (see kernels_constraints.cu for the design).
//
// Kept in a header of its own, free of launch syntax, so that tests/emu_constraints.cpp can compile this very source for
// the host and run it against the CPU oracle: the condense / distribute kernels have no cross-thread communication and run
// thread after thread; the pattern kernels (warp-synchronous bitonic sort, ballots, integer atomics) and the diagonal sum
// (block-wide tree) run with one OS thread per CUDA thread and barriers standing in for __syncwarp / __syncthreads.
#pragma once
#include <cmath>
#include <cstdint>

namespace pe_constraints_dev {

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

constexpr int ROW_WARPS = 4;  // warps (rows) per CTA of row_pattern_var

// one warp per row: gather the lists that contain the row's dof, bitonic sort in shared memory, unique
template <bool WRITE>
__global__ void row_pattern_var(const int32_t* __restrict__ lptr, const int32_t* __restrict__ ldofs, int64_t n_owned,
                                const int32_t* __restrict__ adj_ptr, const int32_t* __restrict__ adj, int cap, int32_t* __restrict__ rowlen,
                                const int32_t* __restrict__ rowptr, int32_t* __restrict__ col, int* __restrict__ overflow) {
#ifdef PE_EMULATE_ON_HOST
  int32_t* smem_rows = pe_emulated_dynamic_smem();  // tests/emu_constraints.cpp
#else
  extern __shared__ int32_t smem_rows[];
#endif
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

}  // namespace pe_constraints_dev
