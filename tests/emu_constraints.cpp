// emu_constraints.cpp — TEST INFRASTRUCTURE: runs the device code of the hanging-node constraint kernels on the CPU.
//
// csrc/device/constraints_dev.cuh holds kernels without cross-lane communication, so executing their threads one after
// the other gives the same result as the GPU.  This file supplies threadIdx/blockIdx/blockDim through a sequential shim,
// includes that very header plus the host-side table builders of csrc/device/constraint_tables.hpp (the same code
// libporoel.so compiles), and exposes them with a plain C interface for tests/test_oracle_amr.py.  The launch shapes are
// those of kernels_constraints.cu.  Kernels that synchronise (row_pattern_var: __syncwarp / __ballot_sync; k_sum_abs_diag:
// __syncthreads) run with one OS thread per CUDA thread and std::barrier; the exclusive scans between the pattern kernels
// (kernels_pattern.cu, GPU-verified) are done on the host.
#include <algorithm>
#include <barrier>
#include <cstdint>
#include <cstring>
#include <memory>
#include <thread>
#include <vector>

// ---- CUDA built-ins for the host ----------------------------------------------------------------------------------------
struct EmuDim { unsigned x = 0, y = 1, z = 1; };
static thread_local EmuDim threadIdx, blockIdx, blockDim, gridDim;
struct EmuGroup {  // the threads that synchronise with each other: a warp (__syncwarp, __ballot_sync) or a block (__syncthreads)
  std::barrier<> bar;
  unsigned votes[32] = {0};
  explicit EmuGroup(int n) : bar(n) {}
};
static thread_local EmuGroup* emu_warp = nullptr;
static thread_local EmuGroup* emu_block = nullptr;
static int32_t emu_smem[1 << 22];  // dynamic shared memory of row_pattern_var (one block at a time)
static inline int32_t* pe_emulated_dynamic_smem() { return emu_smem; }
static inline void __syncwarp() { emu_warp->bar.arrive_and_wait(); }
static inline void __syncthreads() { emu_block->bar.arrive_and_wait(); }
static inline unsigned __ballot_sync(unsigned, bool pred) {
  emu_warp->votes[threadIdx.x & 31] = pred ? 1u : 0u;
  emu_warp->bar.arrive_and_wait();
  unsigned m = 0;
  for (int i = 0; i < 32; ++i) m |= emu_warp->votes[i] << i;
  emu_warp->bar.arrive_and_wait();
  return m;
}
static inline int __popc(unsigned x) { return __builtin_popcount(x); }
static inline int atomicAdd(int* p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
static inline int atomicExch(int* p, int v) { return __atomic_exchange_n(p, v, __ATOMIC_SEQ_CST); }
#define __global__ static
#define __device__ static
#define __forceinline__ inline
#define __shared__ static  // one block at a time: a function-local static is shared by the block's threads
#define PE_EMULATE_ON_HOST 1
#include "../poroelasticity-dealii_b200/csrc/device/constraint_tables.hpp"
#include "../poroelasticity-dealii_b200/csrc/device/constraints_dev.cuh"

using namespace pe_constraints_dev;

// kernels without cross-thread communication: one thread after the other
template <class Body>
static void launch(int64_t grid, unsigned block, Body&& body) {
  gridDim.x = (unsigned)grid;
  blockDim.x = block;
  for (unsigned b = 0; b < (unsigned)grid; ++b)
    for (unsigned t = 0; t < block; ++t) {
      blockIdx.x = b;
      threadIdx.x = t;
      body();
    }
}
// kernels with warp / block synchronisation: one OS thread per CUDA thread, block after block
template <class Body>
static void launch_threads(int64_t grid, unsigned block, Body&& body) {
  for (unsigned b = 0; b < (unsigned)grid; ++b) {
    EmuGroup blk((int)block);
    std::vector<std::unique_ptr<EmuGroup>> warps;
    for (unsigned w = 0; w < (block + 31) / 32; ++w) warps.emplace_back(new EmuGroup((int)std::min(32u, block - 32 * w)));
    std::vector<std::thread> th;
    for (unsigned t = 0; t < block; ++t)
      th.emplace_back([&, t] {
        gridDim.x = (unsigned)grid;
        blockDim.x = block;
        blockIdx.x = b;
        threadIdx.x = t;
        emu_block = &blk;
        emu_warp = warps[t / 32].get();
        body();
      });
    for (auto& x : th) x.join();
  }
}
static int64_t div_up(int64_t a, int64_t b) { return (a + b - 1) / b; }

struct Lines {
  std::vector<int32_t> dof, ptr, edof;
  std::vector<double> w, g;
  pe_constraint_tables::Transposed T;
  Lines(int64_t n_local, int64_t n, const int32_t* d, const int32_t* p, const int32_t* e, const double* ww, const double* gg)
      : dof(d, d + n), ptr(p, p + n + 1), edof(e, e + p[n]), w(ww, ww + p[n]), g(n, 0.0) {
    if (gg) g.assign(gg, gg + n);
    T = pe_constraint_tables::transpose(n_local, dof, ptr, edof, w);
  }
  HangView view() const { return HangView{T.hline.data(), dof.data(), T.tline_of.data(), T.t_ptr.data(), T.t_line.data(), T.t_w.data()}; }
};

extern "C" {

// pe_build_pattern_lists with the kernels executed on the CPU (count / fill adjacency with integer atomics, row_pattern_var
// with its warp-synchronous bitonic sort); the exclusive scans of kernels_pattern.cu are done on the host.  Call with
// col == NULL to get nnz and rowptr, then again to fill col.  With use_kernels == 0 the RESULT is restated instead (sorted
// union of the lists around a row) — the two must agree.
int64_t emu_pattern(int64_t n_cells, int nloc, int ncomp, int64_t n_dofs, const int32_t* cell_dofs, int64_t n_lines, const int32_t* line_dof,
                    const int32_t* line_ptr, const int32_t* edof, int32_t* rowptr, int32_t* col, int64_t* max_candidates, int use_kernels) {
  std::vector<int32_t> d(line_dof, line_dof + n_lines), p(line_ptr, line_ptr + n_lines + 1), e(edof, edof + line_ptr[n_lines]);
  pe_constraint_tables::PatternLists PL = pe_constraint_tables::pattern_lists(n_cells, nloc, ncomp, n_dofs, n_dofs, cell_dofs, d, p, e);
  if (max_candidates) *max_candidates = PL.max_candidates;
  if (!use_kernels) {
    std::vector<std::vector<int32_t>> rows(n_dofs);
    for (int64_t l = 0; l + 1 < (int64_t)PL.lptr.size(); ++l)
      for (int a = PL.lptr[l]; a < PL.lptr[l + 1]; ++a)
        rows[PL.ldofs[a]].insert(rows[PL.ldofs[a]].end(), PL.ldofs.begin() + PL.lptr[l], PL.ldofs.begin() + PL.lptr[l + 1]);
    int64_t nnz = 0;
    rowptr[0] = 0;
    for (int64_t r = 0; r < n_dofs; ++r) {
      std::sort(rows[r].begin(), rows[r].end());
      rows[r].erase(std::unique(rows[r].begin(), rows[r].end()), rows[r].end());
      if (col) std::memcpy(col + nnz, rows[r].data(), rows[r].size() * sizeof(int32_t));
      nnz += (int64_t)rows[r].size();
      rowptr[r + 1] = (int32_t)nnz;
    }
    return nnz;
  }
  const int64_t n_lists = n_cells, n_owned = n_dofs;
  const int T = 256;
  auto exclusive_scan = [](std::vector<int32_t>& v) {  // v has n+1 entries; v[n] receives the total
    int32_t run = 0;
    for (size_t i = 0; i + 1 < v.size(); ++i) { int32_t x = v[i]; v[i] = run; run += x; }
    v.back() = run;
    return run;
  };
  std::vector<int32_t> adj_ptr(n_owned + 1, 0), fill(n_owned, 0);
  launch(div_up(n_lists, T), T, [&] { count_adjacency_var(PL.lptr.data(), PL.ldofs.data(), n_lists, n_owned, adj_ptr.data()); });
  const int32_t n_adj = exclusive_scan(adj_ptr);
  std::vector<int32_t> adj(n_adj);
  launch(div_up(n_lists, T), T, [&] { fill_adjacency_var(PL.lptr.data(), PL.ldofs.data(), n_lists, n_owned, adj_ptr.data(), fill.data(), adj.data()); });
  int cap = 32;
  while (cap < PL.max_candidates) cap <<= 1;
  if ((size_t)ROW_WARPS * cap > ((size_t)1 << 22)) return -2;
  int overflow = 0;
  std::vector<int32_t> rp(n_owned + 1, 0);
  const int64_t blocks = div_up(n_owned, ROW_WARPS);
  launch_threads(blocks, ROW_WARPS * 32, [&] {
    row_pattern_var<false>(PL.lptr.data(), PL.ldofs.data(), n_owned, adj_ptr.data(), adj.data(), cap, rp.data(), nullptr, nullptr, &overflow);
  });
  const int32_t nnz = exclusive_scan(rp);
  std::memcpy(rowptr, rp.data(), rp.size() * sizeof(int32_t));
  if (col)
    launch_threads(blocks, ROW_WARPS * 32, [&] {
      row_pattern_var<true>(PL.lptr.data(), PL.ldofs.data(), n_owned, adj_ptr.data(), adj.data(), cap, nullptr, rp.data(), col, &overflow);
    });
  return overflow ? -1 : nnz;
}

double emu_avg_abs_diag(int64_t n, const int32_t* rowptr, const int32_t* col, const double* val) {
  double out = 0;
  launch_threads(1, 1024, [&] { k_sum_abs_diag(n, rowptr, col, val, &out); });  // pe_avg_abs_diag
  return out / (double)n;
}

int emu_condense_matrix(int64_t n, const int32_t* rowptr, const int32_t* col, const double* src, double* dst, int64_t n_lines, const int32_t* line_dof,
                        const int32_t* line_ptr, const int32_t* edof, const double* w, int keep_diag, double hang_diag) {
  Lines L(n, n_lines, line_dof, line_ptr, edof, w, nullptr);
  const HangView H = L.view();
  const int warps = 8;  // pe_condense_matrix
  launch(div_up(n, warps), warps * 32, [&] { k_condense_matrix(n, rowptr, col, src, dst, H, keep_diag, hang_diag); });
  return 0;
}

int emu_condense_vector(int64_t n, int64_t n_lines, const int32_t* line_dof, const int32_t* line_ptr, const int32_t* edof, const double* w, double* v) {
  Lines L(n, n_lines, line_dof, line_ptr, edof, w, nullptr);
  const int64_t nm = (int64_t)L.T.t_master.size();
  launch(div_up(nm, 256), 256, [&] { k_condense_vector_gather(nm, L.T.t_master.data(), L.T.t_ptr.data(), L.T.t_line.data(), L.T.t_w.data(), L.dof.data(), v); });
  launch(div_up(n_lines, 256), 256, [&] { k_zero_lines(n_lines, L.dof.data(), v); });
  return 0;
}

int emu_distribute(int64_t n, int64_t n_lines, const int32_t* line_dof, const int32_t* line_ptr, const int32_t* edof, const double* w, const double* g,
                   double* v) {
  Lines L(n, n_lines, line_dof, line_ptr, edof, w, g);
  launch(div_up(n_lines, 256), 256, [&] { k_distribute_hanging(n_lines, L.dof.data(), L.ptr.data(), L.edof.data(), L.w.data(), L.g.data(), v); });
  return 0;
}

int emu_scatter_inhomogeneity(int64_t n, int64_t n_lines, const int32_t* line_dof, const double* g, double* v) {
  std::fill(v, v + n, 0.0);
  launch(div_up(n_lines, 256), 256, [&] { k_scatter_lines(n_lines, line_dof, g, v); });
  return 0;
}

}  // extern "C"
