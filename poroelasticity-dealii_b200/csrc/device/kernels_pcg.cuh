// kernels_pcg.cuh — persistent Jacobi-CG kernel; included inside the anonymous namespace of kernels_solver.cu
// (it reuses warp_block_rows<LPR>, the peer-memory primitives and the launch constants defined there).
//
// Every iteration of dealii::SolverCG's loop (h = A d, alpha, g += alpha h, x += alpha d, check, z = D^-1 g, beta,
// d = beta d - z; call sites PS:176-179, DS:300-305, SP:210-214) runs inside ONE cooperative launch.  One CTA set
// stays resident (grid = SMs x occupancy); the three phases of an iteration are separated by on-device
// synchronisation only:
//   * each of the two dot-product reductions is "CTA partials -> grid barrier -> EVERY CTA adds all partials in a
//     fixed order"; with several ranks CTA 0 then posts the rank total to the mailbox of every rank (peer memory
//     over NVLink) and all CTAs of all ranks add the mailboxes in rank order — alpha, beta and the stopping
//     decision are bitwise identical everywhere and no host or NCCL round trip exists inside the loop;
//   * one more grid barrier after the direction update.
// Halo: the CTAs first store the boundary entries of d straight into the neighbours' ghost segments, then
// multiply the interior rows (which reference no ghost column), and only then wait for the neighbours' flags
// before the boundary rows — communication and rank skew hide behind the interior SpMV.
// Cross-phase visibility follows the PTX memory model: writers fence before they arrive at a barrier / post a
// flag, readers fence after they observe it (the gpu-scope fence also drops stale L1 lines of this SM).  Only ONE
// thread (or warp) per CTA executes each fence: the CTA barrier in front of it makes the fence cumulative over the
// whole CTA's writes, and a barrier behind the observer's fence extends the acquire to the whole CTA.  (With a
// fence in every thread each sync point cost ~15 us; fences are per-SM L1 invalidations on this architecture.)

struct PcgArgs {
  const int32_t* rowptr;
  const int32_t* col;
  const double* val;
  const double* invdiag;
  double* x;
  double* g;
  double* h;
  double* d;
  double* z;
  int64_t n, n_interior;
  CgState* state;
  double* gh;            // in: g.h of the start state; out: current value (for a follow-up launch)
  double* partials;      // PE_RED_SLOTS * PE_MAX_RED_BLOCKS
  unsigned* tickets;     // [0] reductions, [1] barrier, [2] halo
  int* bar_flag;         // grid barrier epoch
  int* abort;            // raised by any wait that times out
  unsigned long long* timing;  // [0] ns in the SpMV(+dot) phase as seen by CTA 0, [1] number of phases, [2..9] sub-phase ns (PE_PCG_TIMING)
  // peers
  char* const* peer;
  int nranks, me, red_epoch0;
  // halo of d
  int n_neigh, field, halo_epoch0;
  const int32_t* neigh_rank;
  int64_t n_send;
  const int32_t* send_idx;
  const int32_t* send_dest;
  const int32_t* send_nb;
  size_t ctrl_bytes, d_off;
  int max_iterations;    // iterations to run in this launch at most
  BsrArgs bsr;           // block-CSR copy of the matrix (template parameter B > 0)
};

// Bounded wait used by every spin of the persistent kernel.  A timeout (a peer or CTA that never arrives) raises a
// device-wide abort flag that all other waits poll, so the whole grid leaves its loops within one iteration.
__device__ __forceinline__ bool pcg_wait(const int* flag, int epoch, int* abort) {
  const long long t0 = clock64();
  unsigned spins = 0;
  while ((int)(pe_ld_flag(flag) - epoch) < 0) {
    if ((++spins & 1023u) == 0) {
      if (pe_ld_flag(abort)) return false;
      if (clock64() - t0 > 20000000000LL) {
        atomicExch(abort, 1);
        return false;
      }
    }
  }
  return true;
}

// CTA-wide sums of NV values at once (two barriers in total); the results are valid in every lane of warp 0.
// s_buf holds NV*32 doubles.
template <int NV>
__device__ __forceinline__ void block_sum_n(double (&v)[NV], double* s_buf) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    double x = v[k];
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    v[k] = x;
  }
  __syncthreads();  // s_buf may still be read by the previous user
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < NV; ++k) s_buf[k * 32 + w] = v[k];
  }
  __syncthreads();
  if (w == 0) {
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      double t = lane < nw ? s_buf[k * 32 + lane] : 0.0;
      for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
      v[k] = t;
    }
  }
}

// Release fence of a post: system scope only when another GPU reads the mailbox.
__device__ __forceinline__ void pcg_release_fence(int nranks) {
  if (nranks > 1) __threadfence_system(); else __threadfence();
}

__device__ __forceinline__ void pcg_grid_barrier(const PcgArgs& a, int epoch) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();  // release (cumulative over the CTA)
    if (atomicAdd(&a.tickets[1], 1u) == gridDim.x - 1) {
      a.tickets[1] = 0u;
      __threadfence();
      pe_st_flag(a.bar_flag, epoch);
    } else {
      pcg_wait(a.bar_flag, epoch, a.abort);
    }
    __threadfence();  // acquire
  }
  __syncthreads();
}

// Global sum of NV values over all CTAs of all ranks; every thread of every CTA on every rank returns the same
// bits.  Scheme: (1) CTA partials to a fixed slot, (2) grid barrier, (3) EVERY CTA adds all partials itself in a
// fixed order (a few independent L2 loads per lane — cheaper than a last-CTA serial section followed by a
// publish/fetch round trip), (4) multi-GPU only: CTA 0 posts the rank total to every rank's mailbox, all CTAs
// wait for the flags and add the mailboxes in rank order.
template <int NV>
__device__ __forceinline__ bool pcg_allreduce(const PcgArgs& a, double (&v)[NV], int slot0, int red_epoch, int& bar_epoch, double* s_buf, int* s_ok) {
  block_sum_n<NV>(v, s_buf);
  if (threadIdx.x == 0) {
#pragma unroll
    for (int k = 0; k < NV; ++k) a.partials[(size_t)(slot0 + k) * PE_MAX_RED_BLOCKS + blockIdx.x] = v[k];
  }
  pcg_grid_barrier(a, ++bar_epoch);
  if (threadIdx.x < 32) {
    const int lane = threadIdx.x;
    double tot[NV];
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      double p = 0.0;
      for (int i = lane; i < (int)gridDim.x; i += 32) p += __ldcg(a.partials + (size_t)(slot0 + k) * PE_MAX_RED_BLOCKS + i);
      for (int o = 16; o > 0; o >>= 1) p += __shfl_xor_sync(0xffffffffu, p, o);
      tot[k] = p;
    }
    bool good = true;
    if (a.nranks > 1) {
      if (blockIdx.x == 0 && lane < a.nranks) {
        P2PControl* ctl = reinterpret_cast<P2PControl*>(a.peer[lane]);
#pragma unroll
        for (int k = 0; k < NV; ++k) ctl->red_val[red_epoch & 1][a.me][slot0 + k] = tot[k];
        __threadfence_system();
        pe_st_flag(&ctl->red_flag[a.me], red_epoch);
      }
      const P2PControl* mine = reinterpret_cast<const P2PControl*>(a.peer[a.me]);
      if (lane < a.nranks) good = pcg_wait(&mine->red_flag[lane], red_epoch, a.abort);
      good = __all_sync(0xffffffffu, good);
      __threadfence();  // orders the mailbox loads behind the flag loads (both bypass L1)
      double mail[NV];
#pragma unroll
      for (int k = 0; k < NV; ++k) mail[k] = (good && lane < a.nranks) ? pe_ld_mail(&mine->red_val[red_epoch & 1][lane][slot0 + k]) : 0.0;
#pragma unroll
      for (int k = 0; k < NV; ++k) {
        double sum = 0.0;
        for (int q = 0; q < a.nranks; ++q) sum += __shfl_sync(0xffffffffu, mail[k], q);
        tot[k] = sum;
      }
    }
    if (lane == 0) {
#pragma unroll
      for (int k = 0; k < NV; ++k) s_buf[k] = tot[k];
      *s_ok = good ? 1 : 0;
    }
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < NV; ++k) v[k] = s_buf[k];
  return *s_ok != 0;
}

// B == 0: CSR with LPR lanes per row; B > 0: block CSR with B x B blocks
template <int LPR, int B>
__global__ void __launch_bounds__(SPMV_T) k_pcg(PcgArgs a) {
  __shared__ double s_buf[3 * 32];
  __shared__ bool s_last;
  __shared__ int s_ok;
  CgState* st = a.state;
  if (st->done) return;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t gtid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, gsize = (int64_t)gridDim.x * blockDim.x;
  const int64_t gwarp = (int64_t)blockIdx.x * (SPMV_T / 32) + warp, nwarps = (int64_t)gridDim.x * (SPMV_T / 32);
  constexpr int RB = B > 0 ? B : 1;  // scalar rows per (block) row
  const int64_t n_rows = a.n / RB;   // rows (CSR) or block rows (BSR)
  const int64_t nb_all = (n_rows + 31) >> 5, nb_int = a.n_neigh ? ((a.n_interior / RB) >> 5) : nb_all;
  const int it0 = st->it;
  const double tol = st->tol;
  const int max_it = st->max_it;
  double gh = *a.gh;
  const P2PControl* my_ctl = reinterpret_cast<const P2PControl*>(a.peer[a.me]);
  int bar_epoch = pe_ld_flag(a.bar_flag);  // the same value in every CTA: the flag only moves inside barriers
  unsigned long long t_spmv = 0, n_spmv = 0;
  unsigned long long ph[8] = {0, 0, 0, 0, 0, 0, 0, 0};  // send, interior, halo wait, boundary, reduce+fetch d.h, update, reduce+fetch 2, direction+barrier
  const bool timer = blockIdx.x == 0 && threadIdx.x == 0;
  auto stamp = [&](unsigned long long& last, int slot) {
    if (timer) {
      unsigned long long now;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
      ph[slot] += now - last;
      last = now;
    }
  };
  __syncthreads();

  for (int k = 1; k <= a.max_iterations; ++k) {
    const int it = it0 + k;
    const int e_dh = a.red_epoch0 + 2 * k - 1, e_upd = a.red_epoch0 + 2 * k;
    unsigned long long t0 = 0;
    if (blockIdx.x == 0 && threadIdx.x == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    unsigned long long tl = t0;
    // ---- halo of d: store my boundary entries into the neighbours' ghost segments, then publish
    if (a.n_neigh) {
      for (int64_t i = gtid; i < a.n_send; i += gsize) {
        double* dst = reinterpret_cast<double*>(a.peer[a.neigh_rank[a.send_nb[i]]] + a.ctrl_bytes) + a.d_off;
        dst[a.send_dest[i]] = a.d[a.send_idx[i]];
      }
      __syncthreads();
      if (threadIdx.x == 0) {
        __threadfence_system();  // release of the CTA's remote stores
        s_last = (atomicAdd(&a.tickets[2], 1u) == gridDim.x - 1);
        if (s_last) __threadfence_system();
      }
      __syncthreads();
      if (s_last) {
        if ((int)threadIdx.x < a.n_neigh)
          pe_st_flag(&reinterpret_cast<P2PControl*>(a.peer[a.neigh_rank[threadIdx.x]])->halo_flag[a.field][a.me], a.halo_epoch0 + k);
        if (threadIdx.x == 0) a.tickets[2] = 0u;
      }
    }
    // ---- h = A d on the interior rows, d.h
    double acc1[1] = {0.0};
    auto multiply_blocks = [&](int64_t first, int64_t last) {
      for (int64_t rb = first + gwarp; rb < last; rb += nwarps) {
        if constexpr (B == 0) {
          const double mine = warp_block_rows<LPR>(a.rowptr, a.col, a.val, a.d, a.n, rb, lane);
          const int64_t row = (rb << 5) + lane;
          if (row < a.n) { a.h[row] = mine; acc1[0] += mine * a.d[row]; }
        } else {
          double mine[RB];
          warp_block_brows<RB>(a.bsr, a.d, rb, lane, mine);
          const int64_t brow = (rb << 5) + lane;
          if (brow < n_rows) {
#pragma unroll
            for (int r = 0; r < RB; ++r) { a.h[brow * RB + r] = mine[r]; acc1[0] += mine[r] * a.d[brow * RB + r]; }
          }
        }
      }
    };
    stamp(tl, 0);
    multiply_blocks(0, nb_int);
    stamp(tl, 1);
    bool ok = true;
    if (a.n_neigh) {  // the boundary rows need the neighbours' values
      if ((int)threadIdx.x < a.n_neigh) ok = pcg_wait(&my_ctl->halo_flag[a.field][a.neigh_rank[threadIdx.x]], a.halo_epoch0 + k, a.abort);
      if (threadIdx.x < 32) __threadfence_system();  // acquire by the polling warp, extended to the CTA by the barrier below
      ok = __syncthreads_and(ok ? 1 : 0) != 0;
      stamp(tl, 2);
      multiply_blocks(nb_int, nb_all);
      stamp(tl, 3);
    }
    ok = pcg_allreduce<1>(a, acc1, 0, e_dh, bar_epoch, s_buf, &s_ok) && ok;
    // ---- alpha; g += alpha h; x += alpha d; z = D^-1 g; ||g||^2, g.z
    const double dh[1] = {acc1[0]};
    if (blockIdx.x == 0 && threadIdx.x == 0) {
      unsigned long long t1;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
      t_spmv += t1 - t0;
      n_spmv += 1;
    }
    stamp(tl, 4);
    const double alpha = gh / dh[0];
    double acc2[2] = {0.0, 0.0};
    for (int64_t i = gtid; i < a.n; i += gsize) {
      const double gi = a.g[i] + alpha * a.h[i];
      a.g[i] = gi;
      a.x[i] += alpha * a.d[i];
      const double zi = gi * a.invdiag[i];
      a.z[i] = zi;
      acc2[0] += gi * gi;
      acc2[1] += gi * zi;
    }
    stamp(tl, 5);
    ok = pcg_allreduce<2>(a, acc2, 1, e_upd, bar_epoch, s_buf, &s_ok) && ok;
    // ---- SolverControl::check, beta, d = beta d - z
    const double rz[2] = {acc2[0], acc2[1]};
    stamp(tl, 6);
    const double res = sqrt(rz[0]);
    const bool converged = ok && res <= tol;
    const bool failed = !ok || it >= max_it || isnan(res);
    if (blockIdx.x == 0 && threadIdx.x == 0) {
      st->it = it;
      st->res = res;
      if (!ok) st->pad = 1;
      if (converged) st->done = 1; else if (failed) st->done = -1;
    }
    if (converged || failed) break;
    const double beta = rz[1] / gh;
    gh = rz[1];
    for (int64_t i = gtid; i < a.n; i += gsize) a.d[i] = beta * a.d[i] - a.z[i];
    pcg_grid_barrier(a, ++bar_epoch);
    stamp(tl, 7);
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    *a.gh = gh;
    atomicAdd(&a.timing[0], t_spmv);
    atomicAdd(&a.timing[1], n_spmv);
    for (int k = 0; k < 8; ++k) atomicAdd(&a.timing[2 + k], ph[k]);
  }
}
