// kernels_solver.cu — K5-K8, K10-K12: CSR SpMV, fused conjugate gradients, pointwise preconditioners,
// vector kernels, halo exchange and reductions.
//
// The CG recurrence is dealii::SolverCG<>::solve of the 8.4 series as called at
// lib/include/PoroElasticPressureSolver.h:176-179, PoroElasticDisplacementSolver.h:300-305 and
// StrainProjector.h:210-214 (g = Ax-b, h = P^-1 g, d = -h, ...; convergence is tested on the
// recursively updated ||g||_2 after the x/g update; failure after max_steps mirrors
// SolverControl::NoConvergence).  PreconditionSSOR (PS:177-178, DS:302-303, SP:211-212) is an
// inherently sequential sweep; north_star allows its replacement by a Chebyshev-Jacobi polynomial
// preconditioner, which is what runs here (PE_PRECOND_CHEBYSHEV), next to plain Jacobi.
//
// Fusion: SpMV + d.h in one kernel; x += a d, g += a h, ||g||^2 and (Jacobi) g.D^-1g in one kernel;
// every Chebyshev step is one SpMV-fused kernel (r -= A d; d' = c1 d + c2 D^-1 r; z += d').
// All dot products are reduced deterministically: per-block partials in a fixed slot, the last block
// to arrive sums them in a fixed order.  Loop control lives on the device (CgState); the host only
// polls a pinned copy every few iterations, one poll behind the launches, so the GPU never idles.
#include <cmath>
#include <cstdlib>
#include <deque>
#include <string>
#include <type_traits>

#include "pe_internal.cuh"

namespace {

constexpr int VEC_T = 256;
constexpr int SPMV_T = 256;

struct RedArgs {
  double* partials;
  unsigned* counter;
  double* out;
  // fused peer-memory allreduce: when `peer` is set the last block also posts the totals to every rank's mailbox
  char* const* peer;
  int nranks, me, epoch;
  // split launches (interior rows, then boundary rows): the second launch adds the first one's local total
  const double* add_from;
};

// Where a consumer kernel finds a reduced scalar: the local slot (1 rank / NCCL already reduced it in place) or
// the per-sender mailboxes of the peer-memory protocol (summed in rank order => identical bits on every rank).
struct RedIn {
  const double* local;
  const P2PControl* ctl;
  int nranks, epoch;
};

// block-cooperative; call from uniform control flow.  *ok is cleared on a peer timeout.
__device__ __forceinline__ double red_fetch(const RedIn& r, int slot, bool* ok) {
  if (r.ctl == nullptr) return r.local[slot];
  __shared__ double s_val[PE_RED_SLOTS];
  __shared__ int s_ok;
  __syncthreads();  // a second fetch in the same kernel must not overwrite s_ok / s_val before every warp has read them
  if (threadIdx.x < 32) {
    bool good = true;
    if ((int)threadIdx.x < r.nranks) good = pe_wait_flag(&r.ctl->red_flag[threadIdx.x], r.epoch);
    good = __all_sync(0xffffffffu, good);
    __threadfence_system();
    if (threadIdx.x == 0) {
      double s = 0.0;
      if (good)
        for (int q = 0; q < r.nranks; ++q) s += pe_ld_mail(&r.ctl->red_val[r.epoch & 1][q][slot]);
      s_val[slot] = s;
      s_ok = good ? 1 : 0;
    }
  }
  __syncthreads();
  if (!s_ok) *ok = false;
  return s_val[slot];
}

// Block-reduce NV values; the last block of the grid adds the per-block partials in a fixed order
// and stores the totals to out[slot0..slot0+NV).  Requires gridDim.x <= PE_MAX_RED_BLOCKS.
template <int NV>
__device__ __forceinline__ void grid_reduce(double (&v)[NV], RedArgs R, int slot0) {
  __shared__ double s_part[NV][32];
  __shared__ bool s_last;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    double x = v[k];
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    if (lane == 0) s_part[k][w] = x;
  }
  __syncthreads();
  if (w == 0) {
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      double x = lane < nw ? s_part[k][lane] : 0.0;
      for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
      if (lane == 0) R.partials[(size_t)(slot0 + k) * PE_MAX_RED_BLOCKS + blockIdx.x] = x;
    }
  }
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned ticket = atomicAdd(R.counter, 1u);
    s_last = (ticket == gridDim.x - 1);
  }
  __syncthreads();
  if (s_last) {
    __threadfence();
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      double x = 0.0;
      for (int i = threadIdx.x; i < (int)gridDim.x; i += blockDim.x) x += ((volatile double*)R.partials)[(size_t)(slot0 + k) * PE_MAX_RED_BLOCKS + i];
      for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
      __syncthreads();
      if (lane == 0) s_part[k][w] = x;
      __syncthreads();
      if (w == 0) {
        double y = lane < nw ? s_part[k][lane] : 0.0;
        for (int o = 16; o > 0; o >>= 1) y += __shfl_xor_sync(0xffffffffu, y, o);
        if (lane == 0) {
          if (R.add_from) y += R.add_from[k];
          R.out[slot0 + k] = y;
          s_part[k][0] = y;
        }
      }
    }
    if (R.peer) {  // fused allreduce: totals -> every rank's mailbox, then fence, then the epoch flag
      __syncthreads();
      if ((int)threadIdx.x < R.nranks) {
        P2PControl* ctl = reinterpret_cast<P2PControl*>(R.peer[threadIdx.x]);
#pragma unroll
        for (int k = 0; k < NV; ++k) ctl->red_val[R.epoch & 1][R.me][slot0 + k] = s_part[k][0];
        __threadfence_system();
        pe_st_flag(&ctl->red_flag[R.me], R.epoch);
      }
    }
    if (threadIdx.x == 0) *R.counter = 0u;
  }
}

// Streaming loads of the matrix (read once per pass: bypass L1) and cached gathers of the vector.  They are
// `volatile` on purpose: ptxas otherwise sinks every dependent gather right behind its column load and
// recycles one register for all of them, which serialises the loads (memory-level parallelism of ~1 per
// warp).  Volatile asm keeps program order, so all column/value loads of a row pair are in flight before
// the first gather waits.
__device__ __forceinline__ double ld_stream(const double* p) {
  double v;
  asm volatile("ld.global.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ float ld_stream(const float* p) {
  float v;
  asm volatile("ld.global.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ int ld_stream(const int* p) {
  int v;
  asm volatile("ld.global.L1::no_allocate.s32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ double ld_gather(const double* p) {
  double v;
  asm volatile("ld.global.f64 %0, [%1];" : "=d"(v) : "l"(p));
  return v;
}

// One row segment [start, end) handled by the LPR lanes of a group: the first four strided chunks live in
// registers (issue(): 4 column + 4 value loads, nothing dependent), finish() gathers x and accumulates;
// rows longer than 4*LPR continue in a plain loop.
template <int LPR>
struct RowChunks {
  int c[4];
  double v[4];
  int start, end, sub;
  __device__ __forceinline__ void issue_cols(const int32_t* __restrict__ col, int start_, int end_, int sub_) {
    start = start_; end = end_; sub = sub_;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int j = start + sub + k * LPR;
      c[k] = 0;
      if (j < end) c[k] = ld_stream(col + j);
    }
  }
  __device__ __forceinline__ void issue_vals(const double* __restrict__ val) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int j = start + sub + k * LPR;
      v[k] = 0.0;
      if (j < end) v[k] = ld_stream(val + j);
    }
  }
  __device__ __forceinline__ double finish(const int32_t* __restrict__ col, const double* __restrict__ val, const double* __restrict__ x) {
    double xv[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int j = start + sub + k * LPR;
      xv[k] = 0.0;
      if (j < end) xv[k] = ld_gather(x + c[k]);
    }
    double s0 = v[0] * xv[0] + v[2] * xv[2], s1 = v[1] * xv[1] + v[3] * xv[3];
    for (int j = start + sub + 4 * LPR; j < end; j += LPR) s0 += ld_stream(val + j) * ld_gather(x + ld_stream(col + j));
    return s0 + s1;
  }
};

template <int LPR>
__device__ __forceinline__ double group_reduce(double s) {
#pragma unroll
  for (int o = LPR / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  return s;
}

// Sums of the 32 consecutive rows of row block `rb` (one warp): lane L returns (A x)[32 rb + L].
template <int LPR>
__device__ __forceinline__ double warp_block_rows(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                                                  const double* __restrict__ val, const double* __restrict__ x, int64_t n, int64_t rb, int lane) {
  constexpr int G = 32 / LPR;  // rows in flight per round
  const int sub = lane % LPR, grp = lane / LPR;
  const int64_t row = (rb << 5) + lane;
  const int64_t rlo = row < n ? row : n, rhi = row + 1 < n ? row + 1 : n;
  const int ptr_lo = rowptr[rlo], ptr_hi = rowptr[rhi];  // rows past the end are empty
  double mine = 0.0;
  for (int t = 0; t < LPR; t += 2) {  // two rounds (2*G rows) in flight
    const int ra = t * G + grp, rb_ = (t + 1) * G + grp;
    const int sa = __shfl_sync(0xffffffffu, ptr_lo, ra), ea = __shfl_sync(0xffffffffu, ptr_hi, ra);
    const int sb = __shfl_sync(0xffffffffu, ptr_lo, rb_), eb = __shfl_sync(0xffffffffu, ptr_hi, rb_);
    RowChunks<LPR> A, B;
    A.issue_cols(col, sa, ea, sub);
    B.issue_cols(col, sb, eb, sub);
    A.issue_vals(val);
    B.issue_vals(val);
    __syncwarp();  // scheduling fence: keeps the 16 streaming loads above ahead of the first dependent gather
    const double s_a = group_reduce<LPR>(A.finish(col, val, x));
    const double s_b = group_reduce<LPR>(B.finish(col, val, x));
    const double va = __shfl_sync(0xffffffffu, s_a, ((lane - t * G) & (G - 1)) * LPR);
    const double vb = __shfl_sync(0xffffffffu, s_b, ((lane - (t + 1) * G) & (G - 1)) * LPR);
    if (lane / G == t) mine = va;
    if (lane / G == t + 1) mine = vb;
  }
  return mine;
}

// epilogue kinds of the fused SpMV
enum { EPI_PLAIN = 0, EPI_DOT = 1, EPI_RESID = 2, EPI_CHEB = 3 };

struct SpmvArgs {
  const int32_t* rowptr;
  const int32_t* col;
  const double* val;
  const double* x;
  double* y;
  int64_t n;              // rows [row0, n) are processed; row0 is a multiple of 32
  int64_t row0;
  // epilogue operands
  const double* b;        // EPI_RESID: y = A x - b
  const double* invdiag;  // EPI_CHEB
  double* r;              // EPI_CHEB: residual (in/out)
  double* d_out;          // EPI_CHEB: new direction
  double* z;              // EPI_CHEB: accumulated result
  double c1, c2;          // EPI_CHEB coefficients
  const CgState* state;   // early exit when state->done != 0 (may be null)
  RedArgs red;
  int slot;
  // fused halo wait (peer-memory protocol): block until every neighbour's values of `halo_epoch` have landed
  const P2PControl* ctl;
  const int32_t* neigh_rank;
  int n_neigh, field, halo_epoch;
  int* comm_err;          // pinned error word (set on a halo timeout, also outside a solve)
};

// CSR SpMV with warp-blocked rows.  A warp owns 32 CONSECUTIVE rows: their row pointers arrive with one
// coalesced load and are handed out by shuffles (no dependent rowptr -> col -> x chain per row); the
// 32/LPR lane groups sweep the rows LPR rounds, two rounds in flight; each lane ends up holding the sum of
// "its" row, so the result store and every fused epilogue (dot, residual, Chebyshev update) are coalesced
// 32-row accesses instead of single-lane traffic.
template <int LPR, int EPI>
__global__ void __launch_bounds__(SPMV_T) k_spmv(SpmvArgs a) {
  if (a.state && a.state->done) return;
  if (a.ctl) {
    bool ok = true;
    if ((int)threadIdx.x < a.n_neigh) ok = pe_wait_flag(&a.ctl->halo_flag[a.field][a.neigh_rank[threadIdx.x]], a.halo_epoch);
    if (!ok) {
      *a.comm_err = 1;
      if (a.state) { CgState* st = const_cast<CgState*>(a.state); st->pad = 1; st->done = -1; }
    }
    __threadfence_system();
    __syncthreads();
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t n_blocks = (a.n + 31) >> 5;
  double acc[1] = {0.0};
  for (int64_t rb = (a.row0 >> 5) + (int64_t)blockIdx.x * (SPMV_T / 32) + warp; rb < n_blocks; rb += (int64_t)gridDim.x * (SPMV_T / 32)) {
    const int64_t row = (rb << 5) + lane;
    const double mine = warp_block_rows<LPR>(a.rowptr, a.col, a.val, a.x, a.n, rb, lane);
    if (row < a.n) {
      if (EPI == EPI_PLAIN) a.y[row] = mine;
      if (EPI == EPI_DOT) { a.y[row] = mine; acc[0] += mine * a.x[row]; }
      if (EPI == EPI_RESID) { const double g = mine - a.b[row]; a.y[row] = g; acc[0] += g * g; }
      if (EPI == EPI_CHEB) {
        const double rn = a.r[row] - mine;
        a.r[row] = rn;
        const double dn = a.c1 * a.x[row] + a.c2 * a.invdiag[row] * rn;
        a.d_out[row] = dn;
        a.z[row] += dn;
      }
    }
  }
  if (EPI == EPI_DOT || EPI == EPI_RESID) grid_reduce<1>(acc, a.red, a.slot);
}


// ------------------------------------------------------------------------------------------------
// Block-CSR SpMV for the vector-valued (displacement) matrix: B x B blocks, one column index per block
// (8 + 4/B^2 bytes per scalar nonzero instead of 12) and one B-wide gather of x per block instead of B^2
// scalar gathers.  Same warp-blocked scheme as k_spmv: a warp owns 32 consecutive block rows, all 32 lanes
// take one block each of the current block row (27 blocks for interior Q1 nodes in 3D), two block rows in
// flight; values are stored component-major inside a block row so every one of the B^2 value loads of a
// warp is a contiguous segment.
struct BsrArgs {
  const int32_t* bptr;
  const int32_t* bcol;
  const double* bval;
  int64_t n_brows;
};

template <int B>
struct BlockRow {
  int c, nb, j;
  bool p;
  double v[B * B];
  size_t vbase;
  __device__ __forceinline__ void issue(const BsrArgs& m, int base, int nb_, int lane) {
    nb = nb_;
    j = lane;
    p = j < nb;
    vbase = (size_t)base * B * B;
    c = 0;
    if (p) c = ld_stream(m.bcol + base + j);
  }
  __device__ __forceinline__ void issue_vals(const BsrArgs& m) {
#pragma unroll
    for (int k = 0; k < B * B; ++k) {
      v[k] = 0.0;
      if (p) v[k] = ld_stream(m.bval + vbase + (size_t)k * nb + j);
    }
  }
  // partial sums of the B rows of this block row over the blocks this lane owns
  __device__ __forceinline__ void finish(const BsrArgs& m, const double* __restrict__ x, int base, double (&s)[B]) {
    double xv[B];
#pragma unroll
    for (int cc = 0; cc < B; ++cc) {
      xv[cc] = 0.0;
      if (p) xv[cc] = ld_gather(x + (size_t)c * B + cc);
    }
#pragma unroll
    for (int r = 0; r < B; ++r) {
      double t = 0.0;
#pragma unroll
      for (int cc = 0; cc < B; ++cc) t += v[r * B + cc] * xv[cc];
      s[r] = t;
    }
    for (int jj = j + 32; jj < nb; jj += 32) {  // block rows with more than 32 blocks (Q2, unstructured)
      const int c2 = ld_stream(m.bcol + base + jj);
      double x2[B];
#pragma unroll
      for (int cc = 0; cc < B; ++cc) x2[cc] = ld_gather(x + (size_t)c2 * B + cc);
#pragma unroll
      for (int r = 0; r < B; ++r)
#pragma unroll
        for (int cc = 0; cc < B; ++cc) s[r] += ld_stream(m.bval + vbase + (size_t)(r * B + cc) * nb + jj) * x2[cc];
    }
  }
};

// Sums of the 32 consecutive block rows of warp block `wb`: lane L returns the B row sums of block row 32 wb + L.
template <int B>
__device__ __forceinline__ void warp_block_brows(const BsrArgs& m, const double* __restrict__ x, int64_t wb, int lane, double (&mine)[B]) {
  const int64_t brow = (wb << 5) + lane;
  const int64_t blo = brow < m.n_brows ? brow : m.n_brows, bhi = brow + 1 < m.n_brows ? brow + 1 : m.n_brows;
  const int ptr_lo = m.bptr[blo], ptr_hi = m.bptr[bhi];
#pragma unroll
  for (int r = 0; r < B; ++r) mine[r] = 0.0;
  for (int t = 0; t < 32; t += 2) {
    const int sa = __shfl_sync(0xffffffffu, ptr_lo, t), ea = __shfl_sync(0xffffffffu, ptr_hi, t);
    const int sb = __shfl_sync(0xffffffffu, ptr_lo, t + 1), eb = __shfl_sync(0xffffffffu, ptr_hi, t + 1);
    BlockRow<B> RA, RB;
    RA.issue(m, sa, ea - sa, lane);
    RB.issue(m, sb, eb - sb, lane);
    RA.issue_vals(m);
    RB.issue_vals(m);
    __syncwarp();  // scheduling fence, see k_spmv
    double s_a[B], s_b[B];
    RA.finish(m, x, sa, s_a);
    RB.finish(m, x, sb, s_b);
#pragma unroll
    for (int r = 0; r < B; ++r) {
      double va = s_a[r], vb = s_b[r];
      for (int o = 16; o > 0; o >>= 1) {
        va += __shfl_xor_sync(0xffffffffu, va, o);
        vb += __shfl_xor_sync(0xffffffffu, vb, o);
      }
      if (lane == t) mine[r] = va;
      if (lane == t + 1) mine[r] = vb;
    }
  }
}

template <int B, int EPI>
__global__ void __launch_bounds__(SPMV_T) k_spmv_bsr(SpmvArgs a, BsrArgs m) {
  if (a.state && a.state->done) return;
  if (a.ctl) {
    bool ok = true;
    if ((int)threadIdx.x < a.n_neigh) ok = pe_wait_flag(&a.ctl->halo_flag[a.field][a.neigh_rank[threadIdx.x]], a.halo_epoch);
    if (!ok) {
      *a.comm_err = 1;
      if (a.state) { CgState* st = const_cast<CgState*>(a.state); st->pad = 1; st->done = -1; }
    }
    __threadfence_system();
    __syncthreads();
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t n_wb = (m.n_brows + 31) >> 5;
  double acc[1] = {0.0};
  for (int64_t wb = (int64_t)blockIdx.x * (SPMV_T / 32) + warp; wb < n_wb; wb += (int64_t)gridDim.x * (SPMV_T / 32)) {
    const int64_t brow = (wb << 5) + lane;
    double mine[B];
    warp_block_brows<B>(m, a.x, wb, lane, mine);
    if (brow < m.n_brows) {
#pragma unroll
      for (int r = 0; r < B; ++r) {
        const int64_t row = brow * B + r;
        if (EPI == EPI_PLAIN) a.y[row] = mine[r];
        if (EPI == EPI_DOT) { a.y[row] = mine[r]; acc[0] += mine[r] * a.x[row]; }
        if (EPI == EPI_RESID) { const double g = mine[r] - a.b[row]; a.y[row] = g; acc[0] += g * g; }
        if (EPI == EPI_CHEB) {
          const double rn = a.r[row] - mine[r];
          a.r[row] = rn;
          const double dn = a.c1 * a.x[row] + a.c2 * a.invdiag[row] * rn;
          a.d_out[row] = dn;
          a.z[row] += dn;
        }
      }
    }
  }
  if (EPI == EPI_DOT || EPI == EPI_RESID) grid_reduce<1>(acc, a.red, a.slot);
}

// ------------------------------------------------------------------------------------------------
// TMA-fed SpMV on the sliced block-ELL copy (kernels_sell.cuh): the default matrix pass of every CG solve.
#include "kernels_sell.cuh"

template <int B, typename T, int EPI>
__global__ void __launch_bounds__(sell::THREADS, 1) k_spmv_sell(SpmvArgs a, sell::Mat m, sell::Work w) {
  extern __shared__ __align__(128) char sell_smem[];
  __shared__ double s_buf[32];
  __shared__ bool s_last;
  if (a.state && a.state->done) return;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  sell::Ring R = sell::ring_setup(sell_smem, warp, lane);
  const uint64_t policy = sell::evict_first_policy();
  bool waited = false;
  double p0[B], p1[B], p2[B], p3[B];
  sell::Pending pend{0u, -1};
  auto ready = [&](int slice) {
    // boundary slices gather ghost values: wait (once per warp) until every neighbour's halo of this epoch has landed
    if (a.ctl && !waited && slice >= m.first_boundary_slice) {
      bool ok = true;
      if (lane < a.n_neigh) ok = pe_wait_flag(&a.ctl->halo_flag[a.field][a.neigh_rank[lane]], a.halo_epoch);
      ok = __all_sync(0xffffffffu, ok);
      if (!ok && lane == 0) {
        *a.comm_err = 1;
        if (a.state) { CgState* st = const_cast<CgState*>(a.state); st->pad = 1; st->done = -1; }
      }
      __threadfence_system();
      waited = true;
    }
  };
  auto pre = [&](int slice) {
    // operands of the fused epilogue: requested now, in flight while the slice streams
    const int64_t brow = (int64_t)slice * 32 + lane;
    if (brow < m.n_brows) {
#pragma unroll
      for (int r = 0; r < B; ++r) {
        const int64_t row = brow * B + r;
        if (EPI == EPI_DOT || EPI == EPI_CHEB) p0[r] = a.x[row];
        if (EPI == EPI_RESID) p0[r] = a.b[row];
        if (EPI == EPI_CHEB) { p1[r] = a.r[row]; p2[r] = a.invdiag[row]; p3[r] = a.z[row]; }
      }
    }
  };
  double v[1] = {0.0};  // this lane's contributions to the current chunk's partial sum
  const int n_chunks = (m.n_slices + m.chunk - 1) / m.chunk;
  auto done = [&](int slice, double (&acc)[B], int chunk) {
    const int64_t brow = (int64_t)slice * 32 + lane;
    if (brow < m.n_brows) {
#pragma unroll
      for (int r = 0; r < B; ++r) {
        const int64_t row = brow * B + r;
        if (EPI == EPI_PLAIN) a.y[row] = acc[r];
        if (EPI == EPI_DOT) { a.y[row] = acc[r]; v[0] += acc[r] * p0[r]; }
        if (EPI == EPI_RESID) { const double g = acc[r] - p0[r]; a.y[row] = g; v[0] += g * g; }
        if (EPI == EPI_CHEB) {
          const double rn = p1[r] - acc[r];
          a.r[row] = rn;
          const double dn = a.c1 * p0[r] + a.c2 * p2[r] * rn;
          a.d_out[row] = dn;
          a.z[row] = p3[r] + dn;
        }
      }
    }
    if ((EPI == EPI_DOT || EPI == EPI_RESID) && chunk >= 0) {
      sell::sums_finish<1>(w, n_chunks, pend, lane);  // the previous chunk's ticket has long arrived
      pend = sell::sums_post<1>(w, chunk, v, lane);
      v[0] = 0.0;
    }
  };
  sell::stream<B, T>(m, a.x, w.claim, (int)blockIdx.x, (int)gridDim.x, warp, R, lane, policy, ready, pre, done);
  if (EPI == EPI_DOT || EPI == EPI_RESID) sell::sums_finish<1>(w, n_chunks, pend, lane);
  // the last CTA adds the group totals in a fixed order, publishes, and re-arms the claim counter
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    s_last = atomicAdd(a.red.counter, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  if (EPI == EPI_DOT || EPI == EPI_RESID) {
    double tot[1];
    sell::sum_groups<1>(w, (n_chunks + 31) >> 5, tot, s_buf);
    if (threadIdx.x == 0) {
      double y = tot[0];
      if (a.red.add_from) y += a.red.add_from[0];
      a.red.out[a.slot] = y;
      s_buf[0] = y;
    }
    if (a.red.peer) {  // fused allreduce: total -> every rank's mailbox, fence, epoch flag (as grid_reduce)
      __syncthreads();
      if ((int)threadIdx.x < a.red.nranks) {
        P2PControl* ctl = reinterpret_cast<P2PControl*>(a.red.peer[threadIdx.x]);
        ctl->red_val[a.red.epoch & 1][a.red.me][a.slot] = s_buf[0];
        __threadfence_system();
        pe_st_flag(&ctl->red_flag[a.red.me], a.red.epoch);
      }
    }
  }
  if (threadIdx.x == 0) {
    *a.red.counter = 0u;
    *w.claim = 0u;
  }
}

// ---- Chebyshev inner pass with the FP32 copy of the block values (opt-in, PE_CHEB_FP32=1) -------------------------------
// Same warp-blocked scheme as k_spmv_bsr, but the nine values of a block arrive as floats (36 + 4 B per block instead of
// 72 + 4) and are widened to double before the multiply: all arithmetic and all vectors stay FP64.  Used only for the
// passes INSIDE the polynomial preconditioner, never for CG's own h = A d or the residual, so the solve still converges to
// the solution of the FP64 system (the preconditioner is a fixed SPD operator; measured in numpy on the assembled 64^3
// matrix: identical iteration counts, |x - x_fp64| / |x| = 1e-15).
template <int B>
struct BlockRowF {
  int c, nb, j;
  bool p;
  float v[B * B];
  size_t vbase;
  __device__ __forceinline__ void issue(const BsrArgs& m, int base, int nb_, int lane) {
    nb = nb_;
    j = lane;
    p = j < nb;
    vbase = (size_t)base * B * B;
    c = 0;
    if (p) c = ld_stream(m.bcol + base + j);
  }
  __device__ __forceinline__ void issue_vals(const float* __restrict__ bval32) {
#pragma unroll
    for (int k = 0; k < B * B; ++k) {
      v[k] = 0.f;
      if (p) v[k] = ld_stream(bval32 + vbase + (size_t)k * nb + j);
    }
  }
  __device__ __forceinline__ void finish(const BsrArgs& m, const float* __restrict__ bval32, const double* __restrict__ x, int base, double (&s)[B]) {
    double xv[B];
#pragma unroll
    for (int cc = 0; cc < B; ++cc) {
      xv[cc] = 0.0;
      if (p) xv[cc] = ld_gather(x + (size_t)c * B + cc);
    }
#pragma unroll
    for (int r = 0; r < B; ++r) {
      double t = 0.0;
#pragma unroll
      for (int cc = 0; cc < B; ++cc) t += (double)v[r * B + cc] * xv[cc];
      s[r] = t;
    }
    for (int jj = j + 32; jj < nb; jj += 32) {
      const int c2 = ld_stream(m.bcol + base + jj);
      double x2[B];
#pragma unroll
      for (int cc = 0; cc < B; ++cc) x2[cc] = ld_gather(x + (size_t)c2 * B + cc);
#pragma unroll
      for (int r = 0; r < B; ++r)
#pragma unroll
        for (int cc = 0; cc < B; ++cc) s[r] += (double)ld_stream(bval32 + vbase + (size_t)(r * B + cc) * nb + jj) * x2[cc];
    }
  }
};

template <int B>
__device__ __forceinline__ void warp_block_brows_f32(const BsrArgs& m, const float* __restrict__ bval32, const double* __restrict__ x, int64_t wb,
                                                     int lane, double (&mine)[B]) {
  const int64_t brow = (wb << 5) + lane;
  const int64_t blo = brow < m.n_brows ? brow : m.n_brows, bhi = brow + 1 < m.n_brows ? brow + 1 : m.n_brows;
  const int ptr_lo = m.bptr[blo], ptr_hi = m.bptr[bhi];
#pragma unroll
  for (int r = 0; r < B; ++r) mine[r] = 0.0;
  for (int t = 0; t < 32; t += 2) {
    const int sa = __shfl_sync(0xffffffffu, ptr_lo, t), ea = __shfl_sync(0xffffffffu, ptr_hi, t);
    const int sb = __shfl_sync(0xffffffffu, ptr_lo, t + 1), eb = __shfl_sync(0xffffffffu, ptr_hi, t + 1);
    BlockRowF<B> RA, RB;
    RA.issue(m, sa, ea - sa, lane);
    RB.issue(m, sb, eb - sb, lane);
    RA.issue_vals(bval32);
    RB.issue_vals(bval32);
    __syncwarp();  // scheduling fence, see k_spmv
    double s_a[B], s_b[B];
    RA.finish(m, bval32, x, sa, s_a);
    RB.finish(m, bval32, x, sb, s_b);
#pragma unroll
    for (int r = 0; r < B; ++r) {
      double va = s_a[r], vb = s_b[r];
      for (int o = 16; o > 0; o >>= 1) {
        va += __shfl_xor_sync(0xffffffffu, va, o);
        vb += __shfl_xor_sync(0xffffffffu, vb, o);
      }
      if (lane == t) mine[r] = va;
      if (lane == t + 1) mine[r] = vb;
    }
  }
}

// r -= A~ d ; d' = c1 d + c2 D^-1 r ; z += d'   (the EPI_CHEB epilogue of k_spmv_bsr)
template <int B>
__global__ void __launch_bounds__(SPMV_T) k_spmv_bsr_cheb_f32(SpmvArgs a, BsrArgs m, const float* __restrict__ bval32) {
  if (a.state && a.state->done) return;
  if (a.ctl) {
    bool ok = true;
    if ((int)threadIdx.x < a.n_neigh) ok = pe_wait_flag(&a.ctl->halo_flag[a.field][a.neigh_rank[threadIdx.x]], a.halo_epoch);
    if (!ok) {
      *a.comm_err = 1;
      if (a.state) { CgState* st = const_cast<CgState*>(a.state); st->pad = 1; st->done = -1; }
    }
    __threadfence_system();
    __syncthreads();
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t n_wb = (m.n_brows + 31) >> 5;
  for (int64_t wb = (int64_t)blockIdx.x * (SPMV_T / 32) + warp; wb < n_wb; wb += (int64_t)gridDim.x * (SPMV_T / 32)) {
    const int64_t brow = (wb << 5) + lane;
    double mine[B];
    warp_block_brows_f32<B>(m, bval32, a.x, wb, lane, mine);
    if (brow < m.n_brows) {
#pragma unroll
      for (int r = 0; r < B; ++r) {
        const int64_t row = brow * B + r;
        const double rn = a.r[row] - mine[r];
        a.r[row] = rn;
        const double dn = a.c1 * a.x[row] + a.c2 * a.invdiag[row] * rn;
        a.d_out[row] = dn;
        a.z[row] += dn;
      }
    }
  }
}

__global__ void k_to_float(int64_t n, const double* __restrict__ in, float* __restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) out[i] = (float)in[i];
}

// pressure residual (PS:113-155): r = -( M t1 + kappa K p + f ), ||r||^2 -> slot
template <int LPR>
__global__ void __launch_bounds__(SPMV_T)
k_pressure_residual(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col, const double* __restrict__ M,
                    const double* __restrict__ K, const double* __restrict__ t1, const double* __restrict__ p, const double* __restrict__ f,
                    double kappa, double* __restrict__ r, int64_t n, RedArgs red, int slot) {
  constexpr int RPB = SPMV_T / LPR;
  const int sub = threadIdx.x % LPR;
  double acc[1] = {0.0};
  const int64_t n_pad = ((n + RPB - 1) / RPB) * RPB;
  for (int64_t row = (int64_t)blockIdx.x * RPB + threadIdx.x / LPR; row < n_pad; row += (int64_t)gridDim.x * RPB) {
    double sm = 0.0, sk = 0.0;
    if (row < n) {
      const int start = rowptr[row], end = rowptr[row + 1];
      for (int j = start + sub; j < end; j += LPR) {
        const int c = ld_stream(col + j);
        sm += ld_stream(M + j) * t1[c];
        sk += ld_stream(K + j) * p[c];
      }
    }
#pragma unroll
    for (int o = LPR / 2; o > 0; o >>= 1) {
      sm += __shfl_xor_sync(0xffffffffu, sm, o);
      sk += __shfl_xor_sync(0xffffffffu, sk, o);
    }
    if (row < n && sub == 0) {
      double v = sm;        // mass_matrix.vmult(residual, tmp1)
      v += sk * kappa;      // residual += (perm/visc) * laplace_matrix * solution
      v += f[row];          // residual += source
      v *= -1.0;            // residual *= -1
      r[row] = v;
      acc[0] += v * v;
    }
  }
  grid_reduce<1>(acc, red, slot);
}

// t1 = (alpha/dt)(ev - ev0) + (p - p_old)/(M_b dt)      PS:120-131
__global__ void k_residual_t1(int64_t n, const double* __restrict__ ev, const double* __restrict__ ev0, const double* __restrict__ p,
                              const double* __restrict__ p_old, double a, double m, double* __restrict__ t1) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double x1 = (ev[i] - ev0[i]) * a;
  const double x2 = (p[i] - p_old[i]) * m;
  t1[i] = x1 + x2;
}

// SolverControl::check for iteration `it` with ||g||^2 = res2; one thread records it.  Returns true when the
// solve is over (every thread of every block computes the same answer from the same reduced value).
__device__ __forceinline__ bool cg_check(CgState* state, double res2, int it, bool comm_ok) {
  const double res = sqrt(res2);
  const bool converged = res <= state->tol;
  const bool failed = !comm_ok || it >= state->max_it || isnan(res);
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    state->it = it;
    state->res = res;
    if (!comm_ok) state->pad = 1;
    if (converged) state->done = 1; else if (failed) state->done = -1;
  }
  return converged || failed;
}

// CG update (after h = A d): alpha = gh/dh; g += alpha h; x += alpha d; res2 = g.g; [Jacobi] z = D^-1 g, gz = g.z
template <bool JACOBI>
__global__ void __launch_bounds__(VEC_T)
k_cg_update(int64_t n, const CgState* __restrict__ state, RedIn dh_in, const double* __restrict__ gh_cur, double* __restrict__ x,
            double* __restrict__ g, const double* __restrict__ d, const double* __restrict__ h, const double* __restrict__ invdiag,
            double* __restrict__ z, RedArgs red) {
  if (state->done) return;
  bool ok = true;
  const double dh = red_fetch(dh_in, 0, &ok);
  const double alpha = *gh_cur / dh;  // a peer timeout yields dh = 0: the next check flags the failure
  double acc[2] = {0.0, 0.0};
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const double gi = g[i] + alpha * h[i];
    g[i] = gi;
    x[i] += alpha * d[i];
    acc[0] += gi * gi;
    if (JACOBI) {
      const double zi = gi * invdiag[i];
      z[i] = zi;
      acc[1] += gi * zi;
    }
  }
  grid_reduce<2>(acc, red, 1);
}

// [check of iteration `it` when do_check] then d = beta d - z with beta = gz_new / gh_old.  gh lives in a two-entry
// ring indexed by iteration parity (the host knows the parity when it enqueues), so no thread reads a value
// another one is writing.
__global__ void __launch_bounds__(VEC_T)
k_cg_direction(int64_t n, CgState* __restrict__ state, RedIn r_in, int it, bool do_check, const double* __restrict__ gh_cur,
               double* __restrict__ gh_next, double* __restrict__ d, const double* __restrict__ z) {
  if (state->done) return;
  bool ok = true;
  if (do_check) {
    const double res2 = red_fetch(r_in, 1, &ok);
    if (cg_check(state, res2, it, ok)) return;
  }
  const double gz = red_fetch(r_in, 2, &ok);
  const double beta = gz / *gh_cur;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) d[i] = beta * d[i] - z[i];
  if (blockIdx.x == 0 && threadIdx.x == 0) *gh_next = gz;
}

// start of a solve: z = P^-1 g was computed; d = -z ; gh = g.z (in red[2])
__global__ void k_cg_start(int64_t n, const CgState* __restrict__ state, RedIn r_in, double* __restrict__ gh0, double* __restrict__ d,
                           const double* __restrict__ z) {
  if (state->done) return;
  bool ok = true;
  const double gz = red_fetch(r_in, 2, &ok);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) d[i] = -z[i];
  if (blockIdx.x == 0 && threadIdx.x == 0) *gh0 = gz;
}
__global__ void k_cg_init_state(CgState* state, const double* __restrict__ red, double tol, int tol_relative, int max_it) {
  // red[0] = ||g0||^2, red[3] = ||b||^2 (when relative)
  const double res = sqrt(red[0]);
  state->res = res;
  state->res0 = res;
  state->tol = tol_relative ? tol * sqrt(red[3]) : tol;
  state->it = 0;
  state->max_it = max_it;
  state->gh = 0.0;
  state->done = 0;
  state->pad = 0;
  if (res <= state->tol) state->done = 1;
  else if (max_it <= 0 || isnan(res)) state->done = -1;
}

// z = D^-1 g, gz = g.z   (Jacobi at solve start)
__global__ void __launch_bounds__(VEC_T)
k_jacobi_dot(int64_t n, const CgState* __restrict__ state, const double* __restrict__ g, const double* __restrict__ invdiag,
             double* __restrict__ z, RedArgs red) {
  if (state && state->done) return;
  double acc[1] = {0.0};
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const double zi = g[i] * invdiag[i];
    z[i] = zi;
    acc[0] += g[i] * zi;
  }
  grid_reduce<1>(acc, red, 2);
}

// generic dot product into a slot
__global__ void __launch_bounds__(VEC_T)
k_dot(int64_t n, const CgState* __restrict__ state, const double* __restrict__ a, const double* __restrict__ b, RedArgs red, int slot) {
  if (state && state->done) return;
  double acc[1] = {0.0};
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) acc[0] += a[i] * b[i];
  grid_reduce<1>(acc, red, slot);
}

// Chebyshev start: r = g ; d = (1/theta) D^-1 r ; z = d
__global__ void __launch_bounds__(VEC_T)
k_cheb_first(int64_t n, CgState* __restrict__ state, RedIn r_in, int it, bool do_check, const double* __restrict__ g,
             const double* __restrict__ invdiag, double inv_theta, double* __restrict__ r, double* __restrict__ d, double* __restrict__ z) {
  if (state && state->done) return;
  if (do_check) {
    bool ok = true;
    const double res2 = red_fetch(r_in, 1, &ok);
    if (cg_check(state, res2, it, ok)) return;
  }
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const double gi = g[i];
    const double di = inv_theta * invdiag[i] * gi;
    r[i] = gi;
    d[i] = di;
    z[i] = di;
  }
}

__global__ void k_axpy(int64_t n, double a, const double* __restrict__ x, double* __restrict__ y) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) y[i] += a * x[i];
}
__global__ void k_axpby(int64_t n, double a, const double* __restrict__ x, double b, const double* __restrict__ y, double* __restrict__ z) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) z[i] = a * x[i] + b * y[i];
}
__global__ void k_set(int64_t n, double a, double* __restrict__ y) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) y[i] = a;
}
__global__ void k_scale_by(int64_t n, const double* __restrict__ s, double* __restrict__ y, double f) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) y[i] *= s[i] * f;
}
__global__ void k_distribute(int64_t n_lines, const int32_t* __restrict__ line_dof, const double* __restrict__ g, double* __restrict__ v) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_lines) v[line_dof[i]] = g[i];
}
__global__ void k_invdiag(int64_t n, const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col, const double* __restrict__ val,
                          double* __restrict__ invdiag) {
  int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  int lo = rowptr[r], hi = rowptr[r + 1];
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    if (col[mid] < r) lo = mid + 1; else hi = mid;
  }
  invdiag[r] = 1.0 / val[lo];
}
// block-wise max |v| -> partials slot 0 ; finished on the host (tiny)
__global__ void __launch_bounds__(VEC_T) k_absmax(int64_t n, const double* __restrict__ v, double* __restrict__ partials) {
  __shared__ double s[VEC_T / 32];
  double m = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) m = fmax(m, fabs(v[i]));
  for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < VEC_T / 32; ++w) m = fmax(m, s[w]);
    partials[blockIdx.x] = m;
  }
}
// sigma = C : eps pointwise (FSS:189-224)
__global__ void k_stress(int64_t n, int dim, double lambda, double mu, const double* e0, const double* e1, const double* e2,
                         const double* e3, const double* e4, const double* e5, double* s0, double* s1, double* s2, double* s3, double* s4,
                         double* s5) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (dim == 2) {  // entries xx, xy, yy
    const double tr = e0[i] + e2[i];
    s0[i] = 2 * mu * e0[i] + lambda * tr;
    s1[i] = 2 * mu * e1[i];
    s2[i] = 2 * mu * e2[i] + lambda * tr;
  } else {  // xx, xy, xz, yy, yz, zz
    const double tr = e0[i] + e3[i] + e5[i];
    s0[i] = 2 * mu * e0[i] + lambda * tr;
    s1[i] = 2 * mu * e1[i];
    s2[i] = 2 * mu * e2[i];
    s3[i] = 2 * mu * e3[i] + lambda * tr;
    s4[i] = 2 * mu * e4[i];
    s5[i] = 2 * mu * e5[i] + lambda * tr;
  }
}

inline RedArgs red_args(pe_ctx* c) { return RedArgs{c->red.partials.p, c->red.counter.p, c->red.out.p, nullptr, 1, 0, 0, nullptr}; }

inline int vec_grid(pe_ctx* c, int64_t n) {
  int64_t want = (n + VEC_T - 1) / VEC_T;
  return (int)std::max<int64_t>(1, std::min<int64_t>(want, (int64_t)c->sm_count * 8));
}

inline int lanes_per_row(const Field& F) {
  // four strided chunks per lane cover a row: LPR ~ avg_row_length / 4
  const double avg = F.n_owned ? (double)F.nnz / (double)F.n_owned : 1.0;
  if (avg > 64) return 32;
  if (avg > 32) return 16;
  if (avg > 12) return 8;
  return 4;
}

inline int spmv_grid(pe_ctx* c, int64_t n, int lpr) {
  (void)lpr;
  const int rpb = SPMV_T;  // 32 rows per warp, SPMV_T/32 warps
  int64_t want = (n + rpb - 1) / rpb;
  static const int mult = std::getenv("PE_SPMV_GRID_MULT") ? std::atoi(std::getenv("PE_SPMV_GRID_MULT")) : 8;
  return (int)std::max<int64_t>(1, std::min<int64_t>(want, std::min<int64_t>(PE_MAX_RED_BLOCKS, (int64_t)c->sm_count * mult)));
}

template <int B, typename T, int EPI>
void launch_sell_t(pe_ctx* c, const SpmvArgs& a, const sell::Mat& m, const sell::Work& w) {
  static bool attr_set = false;
  if (!attr_set) {
    PE_CUDA(cudaFuncSetAttribute(k_spmv_sell<B, T, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, sell::SMEM_BYTES));
    attr_set = true;
  }
  const int grid = std::max(1, std::min(c->sm_count, (m.n_slices + sell::WARPS - 1) / sell::WARPS));
  k_spmv_sell<B, T, EPI><<<grid, sell::THREADS, sell::SMEM_BYTES, c->stream>>>(a, m, w);
}

inline sell::Mat sell_mat(const SellMat& S, int sm_count) {
  const double slice_bytes = (double)S.n_panels * S.panel_bytes / std::max(1, S.n_slices);
  // ~32 KB per claim: the claim's round trip is prefetched (kernels_sell.cuh), so small units cost nothing and keep the
  // end-of-pass tail (at most one unit per warp) short
  static const double chunk_bytes = std::getenv("PE_CHUNK_KB") ? 1024.0 * std::atof(std::getenv("PE_CHUNK_KB")) : 32768.0;
  int chunk = (int)std::min(16.0, std::max(1.0, std::floor(chunk_bytes / slice_bytes + 0.5)));
  // small matrices: never fewer chunks than warps.  (Measured, round 2: asking for ~8 chunks per warp instead — single-slice
  // chunks for the pressure matrices of a 64^3 block — made their CG pass slower, 0.038 -> 0.042 ms: every chunk ends with a
  // release + ticket for the deterministic sums, and a scalar slice is only 10 KB.)
  chunk = std::max(1, std::min(chunk, S.n_slices / (sm_count * sell::WARPS)));
  static const double l2_mb = std::getenv("PE_L2_RESIDENT_MB") ? std::atof(std::getenv("PE_L2_RESIDENT_MB")) : 64.0;
  const int resident = (double)S.n_panels * S.panel_bytes <= l2_mb * 1048576.0 ? 1 : 0;
  return sell::Mat{S.panels.p, S.slice_ptr.p, S.n_slices, S.first_boundary_slice, chunk, resident, 0, S.n_brows};
}
inline sell::Work sell_work(pe_ctx* c, int parity = 0) {
  return sell::Work{c->red.claim.p + parity, c->red.spart.p, c->red.gcnt.p, c->red.gpart.p, c->red.cap, c->red.gcap};
}

template <int EPI>
void launch_sell(pe_ctx* c, const SellMat& S, const SpmvArgs& a) {
  const sell::Mat m = sell_mat(S, c->sm_count);
  const sell::Work w = sell_work(c);
  if (S.f32) {  // only the Chebyshev inner passes read the FP32 copy
    if constexpr (EPI == EPI_CHEB) {
      if (S.B == 3) launch_sell_t<3, float, EPI_CHEB>(c, a, m, w);
      else if (S.B == 2) launch_sell_t<2, float, EPI_CHEB>(c, a, m, w);
      else launch_sell_t<1, float, EPI_CHEB>(c, a, m, w);
      return;
    }
    throw PeError(PE_ERR_STATE, "FP32 matrix copy outside the preconditioner");
  }
  if (S.B == 3) launch_sell_t<3, double, EPI>(c, a, m, w);
  else if (S.B == 2) launch_sell_t<2, double, EPI>(c, a, m, w);
  else launch_sell_t<1, double, EPI>(c, a, m, w);
}

template <int EPI>
void launch_spmv(pe_ctx* c, Field& F, SpmvArgs& a) {
  a.rowptr = F.rowptr.p;
  a.col = F.col.p;
  if (a.n == 0) a.n = F.n_owned;
  if (a.red.partials == nullptr) a.red = red_args(c);
  const int lpr = lanes_per_row(F);
  const int grid = spmv_grid(c, a.n - a.row0, lpr);
  if (c->profiling && !c->prof_hold) pe_prof_begin(c, &F == &c->fu ? 1 : 0);
  if (a.row0 == 0 && a.n == F.n_owned) {  // TMA-fed sliced block-ELL copy (default for every matrix it could be built for)
    static const bool cheb32 = std::getenv("PE_CHEB_FP32") == nullptr || std::string(std::getenv("PE_CHEB_FP32")) != "0";
    const SellMat* S = nullptr;
    if (EPI == EPI_CHEB && cheb32) S = F.find_sell(a.val, true);
    if (!S) S = F.find_sell(a.val, false);
    if (S) {
      launch_sell<EPI>(c, *S, a);
      if (c->profiling && !c->prof_hold) pe_prof_end(c);
      c->st.kernel_launches++;
      return;
    }
  }
  if (F.bsr.B && a.val == c->A.p && a.row0 == 0 && a.n == F.n_owned) {  // the displacement matrix has a block-CSR copy
    BsrArgs m{F.bsr.bptr.p, F.bsr.bcol.p, F.bsr.bval.p, F.bsr.n_brows};
    const int bgrid = spmv_grid(c, F.bsr.n_brows, 32);
    if (EPI == EPI_CHEB && F.bsr.bval32.p) {  // PE_CHEB_FP32=1: the polynomial preconditioner reads the FP32 copy of the values
      if (F.bsr.B == 3) k_spmv_bsr_cheb_f32<3><<<bgrid, SPMV_T, 0, c->stream>>>(a, m, F.bsr.bval32.p);
      else k_spmv_bsr_cheb_f32<2><<<bgrid, SPMV_T, 0, c->stream>>>(a, m, F.bsr.bval32.p);
      if (c->profiling && !c->prof_hold) pe_prof_end(c);
      c->st.kernel_launches++;
      return;
    }
    if (F.bsr.B == 3) k_spmv_bsr<3, EPI><<<bgrid, SPMV_T, 0, c->stream>>>(a, m);
    else k_spmv_bsr<2, EPI><<<bgrid, SPMV_T, 0, c->stream>>>(a, m);
    if (c->profiling && !c->prof_hold) pe_prof_end(c);
    c->st.kernel_launches++;
    return;
  }
  switch (lpr) {
    case 32: k_spmv<32, EPI><<<grid, SPMV_T, 0, c->stream>>>(a); break;
    case 16: k_spmv<16, EPI><<<grid, SPMV_T, 0, c->stream>>>(a); break;
    case 8: k_spmv<8, EPI><<<grid, SPMV_T, 0, c->stream>>>(a); break;
    default: k_spmv<4, EPI><<<grid, SPMV_T, 0, c->stream>>>(a); break;
  }
  if (c->profiling && !c->prof_hold) pe_prof_end(c);
  c->st.kernel_launches++;
}

#include "kernels_pcg.cuh"
#include "kernels_pcg2.cuh"

template <int B, typename TI>
void launch_pcg2_t(pe_ctx* c, Pcg2Args& a) {
  static int per_sm = -1;
  if (per_sm < 0) {
    PE_CUDA(cudaFuncSetAttribute(k_pcg2<B, TI>, cudaFuncAttributeMaxDynamicSharedMemorySize, sell::SMEM_BYTES));
    PE_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_pcg2<B, TI>, sell::THREADS, sell::SMEM_BYTES));
  }
  if (per_sm < 1) throw PeError(PE_ERR_CUDA, "persistent CG kernel does not fit on an SM");
  // one CTA per SM; small systems take fewer CTAs (cheaper barriers), every CTA at least eight slices
  const int grid = std::max(1, std::min(c->sm_count, (a.m64.n_slices + sell::WARPS - 1) / sell::WARPS));
  void* params[] = {&a};
  PE_CUDA(cudaLaunchCooperativeKernel((const void*)k_pcg2<B, TI>, dim3(grid), dim3(sell::THREADS), params, sell::SMEM_BYTES, c->stream));
  c->st.kernel_launches++;
}

template <int LPR, int B>
void launch_pcg_t(pe_ctx* c, PcgArgs& a, int& grid_cache) {
  {  // the cooperative grid belongs to THIS instantiation (the format / lanes per row of a field can change with the mesh)
    static int per_sm = 0;
    if (per_sm == 0) PE_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_pcg<LPR, B>, SPMV_T, 0));
    grid_cache = std::max(1, std::min(per_sm * c->sm_count, PE_MAX_RED_BLOCKS));
  }
  void* params[] = {&a};
  PE_CUDA(cudaLaunchCooperativeKernel((const void*)k_pcg<LPR, B>, dim3(grid_cache), dim3(SPMV_T), params, 0, c->stream));
  c->st.kernel_launches++;
}

}  // namespace

// ------------------------------------------------------------------------------------------------
void pe_prof_begin(pe_ctx* c, int field) {
  if (c->prof_used == PE_PROF_PAIRS) pe_prof_flush(c);
  c->prof_field[c->prof_used] = field;
  cudaEventRecord(c->prof_ev[2 * c->prof_used], c->stream);
}
void pe_prof_end(pe_ctx* c) {
  cudaEventRecord(c->prof_ev[2 * c->prof_used + 1], c->stream);
  c->prof_used++;
}
void pe_prof_flush(pe_ctx* c) {
  if (!c->prof_used) return;
  cudaEventSynchronize(c->prof_ev[2 * c->prof_used - 1]);
  for (int i = 0; i < c->prof_used; ++i) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, c->prof_ev[2 * i], c->prof_ev[2 * i + 1]);
    if (c->prof_field[i]) { c->st.spmv_ms_u += ms; c->st.spmv_timed_u++; }
    else { c->st.spmv_ms_p += ms; c->st.spmv_timed_p++; }
  }
  c->prof_used = 0;
}

void pe_vec_axpy(pe_ctx* c, int64_t n, double a, const double* x, double* y) {
  if (!n) return;
  k_axpy<<<vec_grid(c, n), VEC_T, 0, c->stream>>>(n, a, x, y);
  c->st.kernel_launches++;
}
void pe_vec_copy(pe_ctx* c, int64_t n, const double* x, double* y) {
  if (n) PE_CUDA(cudaMemcpyAsync(y, x, n * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
}
void pe_vec_set(pe_ctx* c, int64_t n, double a, double* y) {
  if (!n) return;
  k_set<<<vec_grid(c, n), VEC_T, 0, c->stream>>>(n, a, y);
  c->st.kernel_launches++;
}
void pe_vec_axpby_vals(pe_ctx* c, int64_t n, double a, const double* x, double b, const double* y, double* z) {
  if (!n) return;
  k_axpby<<<vec_grid(c, n), VEC_T, 0, c->stream>>>(n, a, x, b, y, z);
  c->st.kernel_launches++;
}
void pe_distribute(pe_ctx* c, Field& F, double* v) {
  if (!F.n_lines) return;
  k_distribute<<<pe_div_up(F.n_lines, VEC_T), VEC_T, 0, c->stream>>>(F.n_lines, F.line_dof.p, F.line_g.p, v);
  c->st.kernel_launches++;
}
void pe_extract_invdiag(pe_ctx* c, Field& F, const double* val, double* invdiag) {
  k_invdiag<<<pe_div_up(F.n_owned, VEC_T), VEC_T, 0, c->stream>>>(F.n_owned, F.rowptr.p, F.col.p, val, invdiag);
  c->st.kernel_launches++;
}

double pe_linfty(pe_ctx* c, Field& F, const double* v) {
  const int grid = std::min(vec_grid(c, F.n_owned), PE_MAX_RED_BLOCKS);
  k_absmax<<<grid, VEC_T, 0, c->stream>>>(F.n_owned, v, c->red.partials.p);
  c->st.kernel_launches++;
  std::vector<double> h(grid);
  PE_CUDA(cudaMemcpyAsync(h.data(), c->red.partials.p, grid * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  PE_CUDA(cudaStreamSynchronize(c->stream));
  double m = 0;
  for (double x : h) m = std::max(m, x);
  if (c->nranks > 1) {
    double* dev = c->red.out.p + PE_RED_SLOTS;
    PE_CUDA(cudaMemcpyAsync(dev, &m, sizeof(double), cudaMemcpyHostToDevice, c->stream));
    PE_NCCL(ncclAllReduce(dev, dev, 1, ncclDouble, ncclMax, c->comm, c->stream));
    PE_CUDA(cudaMemcpyAsync(&m, dev, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    PE_CUDA(cudaStreamSynchronize(c->stream));
  }
  return m;
}

// FP32 copy of the block values for the Chebyshev preconditioner (opt-in: PE_CHEB_FP32=1 and a Chebyshev preconditioner)
void pe_build_bsr_fp32(pe_ctx* c, Field& F) {
  F.bsr.bval32.release();
  static const bool want = std::getenv("PE_CHEB_FP32") && std::string(std::getenv("PE_CHEB_FP32")) == "1";
  if (!want || !F.bsr.B || c->prm.preconditioner != PE_PRECOND_CHEBYSHEV) return;
  const int64_t nv = F.bsr.nnzb * F.bsr.B * F.bsr.B;
  F.bsr.bval32.alloc((size_t)nv);
  k_to_float<<<vec_grid(c, nv), VEC_T, 0, c->stream>>>(nv, F.bsr.bval.p, F.bsr.bval32.p);
  c->st.kernel_launches++;
  PE_CUDA(cudaGetLastError());
}

double pe_vec_dot(pe_ctx* c, Field& F, const double* a, const double* b) {
  const int64_t n = F.n_owned;
  RedArgs R = red_args(c);
  k_dot<<<std::min(vec_grid(c, n), PE_MAX_RED_BLOCKS), VEC_T, 0, c->stream>>>(n, nullptr, a, b, R, 0);
  c->st.kernel_launches++;
  pe_allreduce_sum(c, c->red.out.p, 1);
  PE_CUDA(cudaMemcpyAsync(c->h_scalars, c->red.out.p, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  pe_sync_checked(c);
  PE_CUDA(cudaGetLastError());
  return c->h_scalars[0];
}

void pe_stress_kernel(pe_ctx* c) {
  const int64_t n = c->fp.n_owned;
  auto E = [&](int i) { return i < c->n_stress ? c->strains[i].p : nullptr; };
  auto S = [&](int i) { return i < c->n_stress ? c->stresses[i].p : nullptr; };
  k_stress<<<pe_div_up(n, VEC_T), VEC_T, 0, c->stream>>>(n, c->dim, c->prm.lame_lambda, c->prm.shear_modulus, E(0), E(1), E(2), E(3), E(4), E(5),
                                                        S(0), S(1), S(2), S(3), S(4), S(5));
  c->st.kernel_launches++;
}

void pe_spmv_plain(pe_ctx* c, Field& F, const double* val, const double* x, double* y) {
  SpmvArgs a{};
  a.val = val;
  a.x = x;
  a.y = y;
  a.state = nullptr;
  launch_spmv<EPI_PLAIN>(c, F, a);
}

double pe_pressure_residual(pe_ctx* c, double dt) {
  Field& F = c->fp;
  const int64_t n = F.n_owned;
  k_residual_t1<<<pe_div_up(n, VEC_T), VEC_T, 0, c->stream>>>(n, c->ev.p, c->ev0.p, c->p.p, c->p_old.p, c->prm.biot_coef / dt,
                                                             1. / c->prm.m_modulus / dt, c->t1.p);
  c->st.kernel_launches++;
  pe_halo_exchange(c, F, c->t1.p);
  pe_halo_exchange(c, F, c->p.p);
  const int lpr = lanes_per_row(F);
  const int grid = spmv_grid(c, n, lpr);
  RedArgs R = red_args(c);
  switch (lpr) {
    case 32: k_pressure_residual<32><<<grid, SPMV_T, 0, c->stream>>>(F.rowptr.p, F.col.p, c->M.p, c->K.p, c->t1.p, c->p.p, c->frhs.p, c->prm.perm_over_visc, c->resid.p, n, R, 0); break;
    case 16: k_pressure_residual<16><<<grid, SPMV_T, 0, c->stream>>>(F.rowptr.p, F.col.p, c->M.p, c->K.p, c->t1.p, c->p.p, c->frhs.p, c->prm.perm_over_visc, c->resid.p, n, R, 0); break;
    case 8: k_pressure_residual<8><<<grid, SPMV_T, 0, c->stream>>>(F.rowptr.p, F.col.p, c->M.p, c->K.p, c->t1.p, c->p.p, c->frhs.p, c->prm.perm_over_visc, c->resid.p, n, R, 0); break;
    default: k_pressure_residual<4><<<grid, SPMV_T, 0, c->stream>>>(F.rowptr.p, F.col.p, c->M.p, c->K.p, c->t1.p, c->p.p, c->frhs.p, c->prm.perm_over_visc, c->resid.p, n, R, 0); break;
  }
  c->st.kernel_launches++;
  c->st.spmv_launches_p += 2;
  pe_allreduce_sum(c, c->red.out.p, 1);
  PE_CUDA(cudaMemcpyAsync(c->h_scalars, c->red.out.p, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  pe_sync_checked(c);
  PE_CUDA(cudaGetLastError());
  return std::sqrt(c->h_scalars[0]);
}

// lambda_max(D^-1 A) by power iteration from a fixed start vector (upper-bounded by the Gershgorin-free
// safety factor applied by the caller)
double pe_estimate_eig_max(pe_ctx* c, Field& F, const double* val, const double* invdiag) {
  const int64_t n = F.n_owned;
  double* v = c->w_d.p;
  double* y = c->w_h.p;
  std::vector<double> h((size_t)n);
  for (int64_t i = 0; i < n; ++i) h[i] = 0.5 + (double)((uint32_t)(i * 2654435761u) >> 22) / 1024.0;
  PE_CUDA(cudaMemsetAsync(v, 0, F.n_local * sizeof(double), c->stream));
  PE_CUDA(cudaMemcpyAsync(v, h.data(), n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  PE_CUDA(cudaStreamSynchronize(c->stream));
  RedArgs R = red_args(c);
  double lambda = 1.0;
  for (int it = 0; it < 30; ++it) {
    pe_halo_exchange(c, F, v);
    pe_spmv_plain(c, F, val, v, y);
    k_scale_by<<<vec_grid(c, n), VEC_T, 0, c->stream>>>(n, invdiag, y, 1.0);  // y = D^-1 A v
    k_dot<<<std::min(vec_grid(c, n), PE_MAX_RED_BLOCKS), VEC_T, 0, c->stream>>>(n, nullptr, y, y, R, 0);
    k_dot<<<std::min(vec_grid(c, n), PE_MAX_RED_BLOCKS), VEC_T, 0, c->stream>>>(n, nullptr, v, v, R, 1);
    c->st.kernel_launches += 3;
    pe_allreduce_sum(c, c->red.out.p, 2);
    PE_CUDA(cudaMemcpyAsync(c->h_scalars, c->red.out.p, 2 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    pe_sync_checked(c);
    const double ny = std::sqrt(c->h_scalars[0]), nv = std::sqrt(c->h_scalars[1]);
    lambda = ny / nv;
    k_axpby<<<vec_grid(c, n), VEC_T, 0, c->stream>>>(n, 1.0 / ny, y, 0.0, y, v);  // v = y / ||y||
    c->st.kernel_launches++;
  }
  PE_CUDA(cudaGetLastError());
  return lambda;
}

// Preconditioned CG.  x is n_local (ghost slots are scratch), b is read on owned rows.
//
// Communication modes inside the loop:
//   1 rank ............ none;
//   NCCL (PE_COMM=nccl) grouped send/recv before each SpMV, ncclAllReduce in place after each reducing kernel;
//   peer memory ....... k_halo_send stores d into the neighbours' ghost segments, the SpMV itself waits for
//                       the flags; the last block of every reducing kernel posts its totals to all mailboxes and
//                       the consuming kernel sums them in its prologue (no collective launches at all).
CgResult pe_cg_solve(pe_ctx* c, Field& F, const double* val, const double* invdiag, double eig_max, double* x, const double* b, double tol,
                     bool tol_relative_to_b, int64_t* spmv_counter) {
  const int64_t n = F.n_owned;
  double *g = c->w_g.p, *h = c->w_h.p, *d = c->w_d.p, *z = c->w_z.p, *d2 = c->w_d2.p, *r = c->w_r.p;
  CgState* st = c->cg_state.p;
  const RedArgs R0 = red_args(c);
  double* red = c->red.out.p;
  double* ghbuf = red + PE_RED_SLOTS + 2;  // two-entry ring for g.h
  const int vg = std::min(vec_grid(c, n), PE_MAX_RED_BLOCKS);
  const bool cheb = c->prm.preconditioner == PE_PRECOND_CHEBYSHEV && c->prm.chebyshev_degree > 1;
  const int kdeg = c->prm.chebyshev_degree;
  const bool multi = c->nranks > 1;
  const bool fused = multi && c->p2p.on;
  const P2PControl* ctl = fused ? reinterpret_cast<const P2PControl*>(c->p2p.region) : nullptr;
  const int fi = &F == &c->fu ? 1 : 0;
  // Chebyshev interval [lmax/ratio, lmax] on D^-1 A
  const double lmax = eig_max, lmin = eig_max / c->prm.chebyshev_eig_ratio;
  const double theta = 0.5 * (lmax + lmin), delta = 0.5 * (lmax - lmin), sigma = theta / delta;

  // a reducing kernel about to be launched: in fused mode it posts under a fresh epoch
  auto producer = [&](int* epoch) {
    RedArgs R = R0;
    *epoch = 0;
    if (fused) {
      *epoch = (int)(++c->p2p.red_epoch);
      R.peer = c->p2p.d_peer.p;
      R.nranks = c->nranks;
      R.me = c->rank;
      R.epoch = *epoch;
    }
    return R;
  };
  auto consumer = [&](int epoch) { return RedIn{red, fused ? ctl : nullptr, c->nranks, epoch}; };
  // in-place collective for the NCCL mode (the fused mode needs none, one rank neither)
  auto nccl_sum = [&](double* p, int count) {
    if (multi && !fused) pe_allreduce_sum(c, p, count, true);
  };
  // SpMV on a vector that lives in the work area: ship its halo first, let the kernel wait for the flags
  auto spmv_in_solve = [&](SpmvArgs& a, double* xvec, auto epi_tag) {
    constexpr int EPI = decltype(epi_tag)::value;
    int halo_epoch = 0;
    pe_halo_exchange(c, F, xvec, true, fused ? &halo_epoch : nullptr);
    // measured on 8 B200s at 128^3: the extra launch and the short boundary kernel cost more than the hidden halo
    // latency saves (157 vs 147 ms per step), so the split is opt-in
    static const bool want_split = std::getenv("PE_SPLIT_SPMV") != nullptr;
    const bool split = want_split && halo_epoch && F.n_interior >= 4096;
    if (split && c->profiling) {  // one timed "matrix pass" = interior + boundary launch
      pe_prof_begin(c, fi);
      c->prof_hold = true;
    }
    if (split) {
      // rows that reference no ghost column run while the halo is still in flight; their local d.h total is
      // parked in a scratch slot and added by the boundary launch, which alone posts to the peers
      SpmvArgs in = a;
      in.n = F.n_interior;
      in.red = R0;
      in.slot = 3;
      launch_spmv<EPI>(c, F, in);
      a.row0 = F.n_interior;
      if (EPI == EPI_DOT) a.red.add_from = red + 3;
    }
    if (halo_epoch) {
      a.ctl = ctl;
      a.neigh_rank = c->p2p.f[fi].neigh_rank.p;
      a.n_neigh = F.halo.n_neigh;
      a.field = fi;
      a.halo_epoch = halo_epoch;
      a.comm_err = c->h_comm_err;
    }
    launch_spmv<EPI>(c, F, a);
    if (split && c->profiling) {
      c->prof_hold = false;
      pe_prof_end(c);
    }
    (*spmv_counter)++;
  };

  // z = P^-1 g (P = Jacobi or Chebyshev polynomial) and g.z -> slot 2; `check_in`/`it` fold SolverControl::check of
  // the iteration that just updated g into the first kernel (Chebyshev only; Jacobi folds it into k_cg_direction).
  // Returns the epoch under which g.z was posted.
  auto apply_precond_and_dot = [&](bool do_check, RedIn check_in, int it) {
    int e = 0;
    if (!cheb) {
      RedArgs R = producer(&e);
      k_jacobi_dot<<<vg, VEC_T, 0, c->stream>>>(n, st, g, invdiag, z, R);
      c->st.kernel_launches++;
    } else {
      k_cheb_first<<<vg, VEC_T, 0, c->stream>>>(n, st, check_in, it, do_check, g, invdiag, 1.0 / theta, r, d2, z);
      c->st.kernel_launches++;
      double rho = 1.0 / sigma;
      double* din = d2;
      double* dout = h;  // h is free between the update and the next SpMV
      for (int k = 1; k < kdeg; ++k) {
        const double rho_new = 1.0 / (2.0 * sigma - rho);
        SpmvArgs a{};
        a.val = val;
        a.x = din;
        a.invdiag = invdiag;
        a.r = r;
        a.d_out = dout;
        a.z = z;
        a.c1 = rho_new * rho;
        a.c2 = 2.0 * rho_new / delta;
        a.state = st;
        spmv_in_solve(a, din, std::integral_constant<int, EPI_CHEB>{});
        rho = rho_new;
        std::swap(din, dout);
      }
      RedArgs R = producer(&e);
      k_dot<<<vg, VEC_T, 0, c->stream>>>(n, st, g, z, R, 2);
      c->st.kernel_launches++;
    }
    nccl_sum(red + 2, 1);
    return e;
  };

  // g = A x - b, ||g||^2 -> red[0]; ||b||^2 -> red[3]   (cold path: plain collectives, CgState not yet valid)
  pe_halo_exchange(c, F, x);
  {
    SpmvArgs a{};
    a.val = val;
    a.x = x;
    a.y = g;
    a.b = b;
    a.state = nullptr;
    a.slot = 0;
    launch_spmv<EPI_RESID>(c, F, a);
    (*spmv_counter)++;
  }
  if (tol_relative_to_b) {
    k_dot<<<vg, VEC_T, 0, c->stream>>>(n, nullptr, b, b, R0, 3);
    c->st.kernel_launches++;
  }
  pe_allreduce_sum(c, red, PE_RED_SLOTS);
  k_cg_init_state<<<1, 1, 0, c->stream>>>(st, red, tol, tol_relative_to_b ? 1 : 0, c->prm.cg_max_iterations);
  const unsigned red_epoch_start = c->p2p.red_epoch;  // mailbox posts of this solve count from here
  const int max_it = c->prm.cg_max_iterations;
  static const bool pcg_off = std::getenv("PE_PCG") && std::string(std::getenv("PE_PCG")) == "0";
  static const bool pcg2_off = std::getenv("PE_PCG2") && std::string(std::getenv("PE_PCG2")) == "0";
  const SellMat* S64 = F.find_sell(val, false);
  const SellMat* S32 = cheb ? F.find_sell(val, true) : nullptr;
  const bool use_pcg2 = S64 && !pcg_off && !pcg2_off && (!multi || (fused && c->p2p.f[fi].push_ok)) && (!cheb || kdeg <= 8);
  if (!use_pcg2) {  // the multi-kernel loop and the first persistent kernel start from z = P^-1 g, d = -z (k_pcg2 has its own prologue)
    const int e = apply_precond_and_dot(false, consumer(0), 0);
    k_cg_start<<<vg, VEC_T, 0, c->stream>>>(n, st, consumer(e), ghbuf, d, z);
    c->st.kernel_launches += 2;
  }

  // The persistent kernel removes every launch gap and host/NCCL round trip, but its static row assignment cannot
  // balance slow SMs the way CTA scheduling does: measured on B200, it wins when a matrix pass is short
  // (<= ~0.3 ms: 4- and 8-GPU blocks of the 128^3 problem, all pressure solves; 148 vs 161 ms per step at 4 GPUs)
  // and loses ~5-10 % on multi-ms passes.
  static const bool pcg_disabled = std::getenv("PE_PCG") && std::string(std::getenv("PE_PCG")) == "0";
  static const long long pcg_max_nnz = std::getenv("PE_PCG_MAX_NNZ") ? std::atoll(std::getenv("PE_PCG_MAX_NNZ")) : 150000000LL;
  const bool has_bsr = F.bsr.B && val == c->A.p;
  // ---- default: the persistent single-reduction CG kernel on the TMA-fed sliced copy (kernels_pcg2.cuh)
  if (use_pcg2) {
    P2PField& PF = c->p2p.f[fi];
    Pcg2Args pa{};
    pa.m64 = sell_mat(*S64, c->sm_count);
    // (measured, round 2: single-slice chunks for the inner passes alone — they form no sums, so a chunk end is free — gain
    // 11 % on the pressure matrices of a 64^3 mesh on one GPU and LOSE 20 % on the same block as one of 8 ranks; not used)
    if (S32) pa.m32 = sell_mat(*S32, c->sm_count);
    static const bool early_off = std::getenv("PE_HALO_EARLY") && std::string(std::getenv("PE_HALO_EARLY")) == "0";
    pa.m64.boundary_early = pa.m32.boundary_early = (multi && F.halo.n_neigh > 0 && !early_off) ? 1 : 0;
    pa.work = sell_work(c);
    pa.invdiag = invdiag;
    pa.x = x; pa.g = g; pa.d = d; pa.s = c->w_s.p; pa.w = h; pa.z = z; pa.r = r; pa.c0 = d2; pa.c1 = c->w_c1.p;
    pa.n = n;
    pa.n_interior = multi ? F.n_interior : n;
    pa.state = st;
    pa.degree = cheb ? kdeg : 1;
    pa.inv_theta = 1.0 / theta;
    {
      double rho = 1.0 / sigma;
      for (int k = 1; k < pa.degree; ++k) {
        const double rho_new = 1.0 / (2.0 * sigma - rho);
        pa.k1[k] = rho_new * rho;
        pa.k2[k] = 2.0 * rho_new / delta;
        rho = rho_new;
      }
    }
    pa.tickets = c->pcg_tickets.p;
    pa.bar_flag = c->pcg_flags.p;
    pa.abort = c->pcg_flags.p + 1;
    pa.timing = c->pcg_timing.p;
    static const char* trace_prefix = std::getenv("PE_PCG_TRACE");
    const size_t trace_words = (size_t)c->sm_count * sell::WARPS * 8;
    if (trace_prefix && fi == 1) {
      if (!c->pcg_trace.p) c->pcg_trace.alloc_zero(trace_words, c->stream);
      pa.trace = c->pcg_trace.p;
    }
    pa.peer = c->p2p.d_peer.p;
    pa.nranks = c->nranks;
    pa.me = c->rank;
    pa.red_epoch0 = (int)c->p2p.red_epoch;
    pa.n_neigh = multi ? F.halo.n_neigh : 0;
    pa.field = fi;
    pa.halo_epoch0 = (int)PF.epoch;
    pa.neigh_rank = PF.neigh_rank.p;
    pa.push_ptr = PF.push_ptr.p;
    pa.push_dest = PF.push_dest.p;
    pa.push_nb = PF.push_nb.p;
    pa.push_addr = PF.push_addr.p;
    pa.push_src = PF.push_src.p;
    pa.ctrl_bytes = c->p2p.ctrl_bytes;
    const double* w0 = reinterpret_cast<const double*>(c->p2p.region + c->p2p.ctrl_bytes);
    pa.off_z = (size_t)(z - w0);
    pa.off_c0 = (size_t)(d2 - w0);
    pa.off_c1 = (size_t)(c->w_c1.p - w0);
    pa.max_iterations = max_it;
    PE_CUDA(cudaMemsetAsync(c->pcg_timing.p, 0, PE_PCG_TIMING_WORDS * sizeof(unsigned long long), c->stream));
    PE_CUDA(cudaMemsetAsync(c->pcg_flags.p + 1, 0, sizeof(int), c->stream));  // abort flag
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (c->profiling) {
      PE_CUDA(cudaEventCreate(&e0));
      PE_CUDA(cudaEventCreate(&e1));
      PE_CUDA(cudaEventRecord(e0, c->stream));
    }
    const bool f32 = S32 != nullptr;
    if (S64->B == 3) { if (f32) launch_pcg2_t<3, float>(c, pa); else launch_pcg2_t<3, double>(c, pa); }
    else if (S64->B == 2) { if (f32) launch_pcg2_t<2, float>(c, pa); else launch_pcg2_t<2, double>(c, pa); }
    else { if (f32) launch_pcg2_t<1, float>(c, pa); else launch_pcg2_t<1, double>(c, pa); }
    if (c->profiling) PE_CUDA(cudaEventRecord(e1, c->stream));
    unsigned long long h_timing[PE_PCG_TIMING_WORDS] = {0};
    PE_CUDA(cudaMemcpyAsync(&c->h_state[0], st, sizeof(CgState), cudaMemcpyDeviceToHost, c->stream));
    PE_CUDA(cudaMemcpyAsync(h_timing, c->pcg_timing.p, sizeof h_timing, cudaMemcpyDeviceToHost, c->stream));
    pe_sync_checked(c);
    PE_CUDA(cudaGetLastError());
    const CgState last = c->h_state[0];
    if (last.pad) {  // a wait timed out: barriers, claim counters and group counters may be half-used; start clean next time
      PE_CUDA(cudaMemsetAsync(c->pcg_tickets.p, 0, 4 * sizeof(unsigned), c->stream));
      PE_CUDA(cudaMemsetAsync(c->red.claim.p, 0, 6 * sizeof(unsigned), c->stream));
      PE_CUDA(cudaMemsetAsync(c->red.gcnt.p, 0, (size_t)c->red.gcap * sizeof(unsigned), c->stream));
      PE_CUDA(cudaStreamSynchronize(c->stream));
    }
    if (pa.trace) {  // diagnostic dump: <prefix>_rank<r>.bin, 8 u64 per warp (kernels_pcg2.cuh), last iteration of this solve
      std::vector<unsigned long long> h(trace_words);
      PE_CUDA(cudaMemcpy(h.data(), c->pcg_trace.p, trace_words * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
      const std::string path = std::string(trace_prefix) + "_rank" + std::to_string(c->rank) + ".bin";
      if (FILE* f = std::fopen(path.c_str(), "wb")) {
        std::fwrite(h.data(), sizeof(unsigned long long), h.size(), f);
        std::fclose(f);
      }
    }
    c->p2p.red_epoch += (unsigned)h_timing[11];  // identical on every rank
    if (multi) PF.epoch += (unsigned)h_timing[10];
    const long long passes_cg = (long long)last.it + (last.done ? 1 : 0), passes_in = passes_cg * (pa.degree - 1);
    *spmv_counter += passes_cg + passes_in;
    if (c->profiling) {
      float ms = 0.f;
      PE_CUDA(cudaEventElapsedTime(&ms, e0, e1));
      cudaEventDestroy(e0);
      cudaEventDestroy(e1);
      (fi ? c->st.pcg_ms_u : c->st.pcg_ms_p) += ms;
      (fi ? c->st.pcg_iterations_u : c->st.pcg_iterations_p) += last.it;
      // in-kernel %globaltimer of CTA 0: the FP64 CG passes including the reduction that ends them
      (fi ? c->st.spmv_ms_u : c->st.spmv_ms_p) += (double)(h_timing[0] + h_timing[5]) * 1e-6;
      (fi ? c->st.spmv_timed_u : c->st.spmv_timed_p) += (int64_t)h_timing[1];
      if (!fi) {
        const int src[8] = {0, 2, 4, 5, 6, 7, 8, 9};
        for (int k = 0; k < 8; ++k) c->st.phase_ms_p[k] += (double)h_timing[src[k]] * 1e-6;
        c->st.phase_ms_p[8] += (double)h_timing[1];
        c->st.phase_ms_p[9] += (double)h_timing[3];
      }
      if (fi) {
        c->st.inner_ms_u += (double)h_timing[2] * 1e-6;
        c->st.inner_passes_u += (int64_t)h_timing[3];
        c->st.update_ms_u += (double)h_timing[4] * 1e-6;
        c->st.reduce_ms_u += (double)h_timing[5] * 1e-6;
        c->st.wait_inner_ms_u += (double)h_timing[6] * 1e-6;
        c->st.wait_cg_ms_u += (double)h_timing[7] * 1e-6;
        c->st.wait_peer_ms_u += (double)h_timing[8] * 1e-6;
        c->st.wait_update_ms_u += (double)h_timing[9] * 1e-6;
        const SellMat* Sin = S32 ? S32 : S64;
        c->st.inner_bytes_u = (double)Sin->nnzb * (Sin->B * Sin->B * (Sin->f32 ? 4.0 : 8.0) + 4.0) + (double)Sin->n_slices * 4.0 + (double)n * 56.0;
      }
    }
    CgResult out;
    out.its = last.it;
    out.res = last.res;
    out.status = last.done == 1 ? PE_OK : (last.pad ? PE_ERR_NCCL : (std::isnan(last.res) ? PE_ERR_NAN : PE_ERR_NO_CONVERGENCE));
    return out;
  }
  if (!cheb && !pcg_disabled && (!multi || fused) && F.nnz <= pcg_max_nnz) {
    // ---- the whole CG loop in one persistent cooperative launch (kernels_pcg.cuh)
    P2PField& PF = c->p2p.f[fi];
    PcgArgs pa{};
    pa.rowptr = F.rowptr.p;
    pa.col = F.col.p;
    pa.val = val;
    pa.invdiag = invdiag;
    pa.x = x;
    pa.g = g;
    pa.h = h;
    pa.d = d;
    pa.z = z;
    pa.n = n;
    pa.n_interior = F.n_interior;
    pa.state = st;
    pa.gh = ghbuf;
    pa.partials = c->red.partials.p;
    pa.tickets = c->pcg_tickets.p;
    pa.bar_flag = c->pcg_flags.p;
    pa.abort = c->pcg_flags.p + 1;
    pa.timing = c->pcg_timing.p;
    pa.peer = c->p2p.d_peer.p;
    pa.nranks = c->nranks;
    pa.me = c->rank;
    pa.red_epoch0 = (int)c->p2p.red_epoch;
    pa.n_neigh = multi ? F.halo.n_neigh : 0;
    pa.field = fi;
    pa.halo_epoch0 = (int)PF.epoch;
    pa.neigh_rank = PF.neigh_rank.p;
    pa.n_send = multi ? F.halo.n_send() : 0;
    pa.send_idx = F.halo.send_idx.p;
    pa.send_dest = PF.send_dest.p;
    pa.send_nb = PF.send_nb.p;
    pa.ctrl_bytes = c->p2p.ctrl_bytes;
    pa.d_off = (size_t)(d - reinterpret_cast<double*>(c->p2p.region + c->p2p.ctrl_bytes));
    pa.max_iterations = max_it;
    PE_CUDA(cudaMemsetAsync(c->pcg_timing.p, 0, 10 * sizeof(unsigned long long), c->stream));
    PE_CUDA(cudaMemsetAsync(c->pcg_flags.p + 1, 0, sizeof(int), c->stream));  // abort flag
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (c->profiling) {
      PE_CUDA(cudaEventCreate(&e0));
      PE_CUDA(cudaEventCreate(&e1));
      PE_CUDA(cudaEventRecord(e0, c->stream));
    }
    if (has_bsr) {
      pa.bsr = BsrArgs{F.bsr.bptr.p, F.bsr.bcol.p, F.bsr.bval.p, F.bsr.n_brows};
      if (F.bsr.B == 3) launch_pcg_t<32, 3>(c, pa, c->pcg_grid[fi]);
      else launch_pcg_t<32, 2>(c, pa, c->pcg_grid[fi]);
    } else {
      switch (lanes_per_row(F)) {
        case 32: launch_pcg_t<32, 0>(c, pa, c->pcg_grid[fi]); break;
        case 16: launch_pcg_t<16, 0>(c, pa, c->pcg_grid[fi]); break;
        case 8: launch_pcg_t<8, 0>(c, pa, c->pcg_grid[fi]); break;
        default: launch_pcg_t<4, 0>(c, pa, c->pcg_grid[fi]); break;
      }
    }
    if (c->profiling) PE_CUDA(cudaEventRecord(e1, c->stream));
    unsigned long long h_timing[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    PE_CUDA(cudaMemcpyAsync(&c->h_state[0], st, sizeof(CgState), cudaMemcpyDeviceToHost, c->stream));
    PE_CUDA(cudaMemcpyAsync(h_timing, c->pcg_timing.p, sizeof h_timing, cudaMemcpyDeviceToHost, c->stream));
    pe_sync_checked(c);
    PE_CUDA(cudaGetLastError());
    const CgState last = c->h_state[0];
    const int its = last.it;
    if (last.pad) {  // a wait timed out: barriers may have been left half-arrived; make the next launch start clean
      PE_CUDA(cudaMemsetAsync(c->pcg_tickets.p, 0, 4 * sizeof(unsigned), c->stream));
      PE_CUDA(cudaStreamSynchronize(c->stream));
    }
    c->p2p.red_epoch += 2u * (unsigned)its;  // every rank executed the same number of posts
    if (multi) PF.epoch += (unsigned)its;
    *spmv_counter += its;
    for (int k = 0; k < 8; ++k) c->pcg_phase_ns[fi][k] += (double)h_timing[2 + k];
    c->pcg_phase_its[fi] += its;
    if (c->profiling) {
      float ms = 0.f;
      PE_CUDA(cudaEventElapsedTime(&ms, e0, e1));
      cudaEventDestroy(e0);
      cudaEventDestroy(e1);
      (fi ? c->st.pcg_ms_u : c->st.pcg_ms_p) += ms;
      (fi ? c->st.pcg_iterations_u : c->st.pcg_iterations_p) += its;
      (fi ? c->st.spmv_ms_u : c->st.spmv_ms_p) += (double)h_timing[0] * 1e-6;  // in-kernel %globaltimer, SpMV+dot phase of CTA 0
      (fi ? c->st.spmv_timed_u : c->st.spmv_timed_p) += (int64_t)h_timing[1];
    }
    CgResult out;
    out.its = its;
    out.res = last.res;
    out.status = last.done == 1 ? PE_OK : (last.pad ? PE_ERR_NCCL : (std::isnan(last.res) ? PE_ERR_NAN : PE_ERR_NO_CONVERGENCE));
    return out;
  }
  const int interval = c->prm.cg_check_interval > 0 ? c->prm.cg_check_interval : 8;
  int launched = 0;
  std::deque<int> inflight;  // pinned poll slots whose copy is enqueued but not yet consumed
  int slot = 0;
  auto launch_chunk = [&]() {
    const int chunk = std::min(interval, max_it - launched);
    for (int k = 0; k < chunk; ++k) {
      const int it = launched + k + 1;
      double* gh_cur = ghbuf + ((it - 1) & 1);
      double* gh_nxt = ghbuf + (it & 1);
      int e_dh = 0, e_upd = 0;
      SpmvArgs a{};
      a.val = val;
      a.x = d;
      a.y = h;
      a.state = st;
      a.slot = 0;
      a.red = producer(&e_dh);
      spmv_in_solve(a, d, std::integral_constant<int, EPI_DOT>{});  // h = A d, slot 0 = d.h
      nccl_sum(red, 1);
      RedArgs Ru = producer(&e_upd);
      if (!cheb) k_cg_update<true><<<vg, VEC_T, 0, c->stream>>>(n, st, consumer(e_dh), gh_cur, x, g, d, h, invdiag, z, Ru);
      else k_cg_update<false><<<vg, VEC_T, 0, c->stream>>>(n, st, consumer(e_dh), gh_cur, x, g, d, h, invdiag, z, Ru);
      c->st.kernel_launches++;
      nccl_sum(red + 1, cheb ? 1 : 2);
      if (!cheb) {
        k_cg_direction<<<vg, VEC_T, 0, c->stream>>>(n, st, consumer(e_upd), it, true, gh_cur, gh_nxt, d, z);
      } else {
        const int e_gz = apply_precond_and_dot(true, consumer(e_upd), it);
        k_cg_direction<<<vg, VEC_T, 0, c->stream>>>(n, st, consumer(e_gz), it, false, gh_cur, gh_nxt, d, z);
      }
      c->st.kernel_launches++;
    }
    launched += chunk;
    PE_CUDA(cudaMemcpyAsync(&c->h_state[slot], st, sizeof(CgState), cudaMemcpyDeviceToHost, c->stream));
    PE_CUDA(cudaEventRecord(c->ev_poll[slot], c->stream));
    inflight.push_back(slot);
    slot ^= 1;
  };
  CgState last{};
  launch_chunk();
  while (true) {
    if (launched < max_it && inflight.size() < 2) launch_chunk();  // stay one chunk ahead of the poll
    const int s = inflight.front();
    inflight.pop_front();
    PE_CUDA(cudaEventSynchronize(c->ev_poll[s]));
    last = c->h_state[s];
    if (last.done != 0) break;
    if (launched >= max_it && inflight.empty()) break;  // unreachable: the folded check flags failure at max_it
  }
  while (!inflight.empty()) {
    PE_CUDA(cudaEventSynchronize(c->ev_poll[inflight.front()]));
    inflight.pop_front();
  }
  pe_sync_checked(c);
  PE_CUDA(cudaGetLastError());
  if (fused) {
    // Kernels enqueued behind the converging iteration returned at `state->done` without posting, but the host counted
    // an epoch for each of them.  Rewind to the posts that really happened (identical on every rank: all ranks take the
    // same stopping decision), so the next post has the opposite mailbox parity of the last one and can never
    // overwrite a value a slower rank has not read yet.
    const unsigned posts = last.it <= 0 ? 0u : (cheb ? 3u * (unsigned)last.it : 1u + 2u * (unsigned)last.it);
    c->p2p.red_epoch = red_epoch_start + posts;
  }
  CgResult out;
  out.its = last.it;
  out.res = last.res;
  out.status = last.done == 1 ? PE_OK : (last.pad ? PE_ERR_NCCL : (std::isnan(last.res) ? PE_ERR_NAN : PE_ERR_NO_CONVERGENCE));
  return out;
}
