// kernels_pcg2.cuh — persistent, single-reduction preconditioned CG on the TMA-fed matrix stream (included inside the
// anonymous namespace of kernels_solver.cu, after kernels_sell.cuh and kernels_pcg.cuh).
//
// The whole loop of dealii::SolverCG (call sites PS:176-179, DS:300-305, SP:210-214) runs in ONE cooperative launch, with
// the recurrence rearranged after Chronopoulos & Gear so that an iteration needs a single global reduction:
//     z = P^-1 g;  w = A z;  gamma = g.z, delta = w.z, rho = g.g          (one pass over A, three fused dot products)
//     beta = gamma/gamma_old;  alpha = gamma / (delta - beta gamma / alpha_old)
//     d = beta d - z;  s = beta s - w;  x += alpha d;  g += alpha s        (s tracks A d; g = A x - b as in deal.II)
// In exact arithmetic the iterates are those of the classic recurrence (profiles/single_reduction_cg_prototype.txt: same
// iteration counts to +-3 % and the same solution to 1e-15 on the assembled matrices).  SolverControl::check sees
// rho = ||g||^2 of the iterate BEFORE the current pass, so convergence is noticed one pass late: x is already final and the
// reported iteration count is the number of updates of x, as in deal.II.
// P^-1 is Jacobi (degree 1) or the Chebyshev polynomial in D^-1 A of kernels_solver.cu, whose inner passes may stream the
// FP32 copy of the matrix (TI = float) — a fixed SPD operator, so CG converges to the FP64 solution.
//
// Phases of an iteration and what separates them (m = polynomial degree):
//   m-1 inner passes  r -= A~ c;  c' = c1 c + c2 D^-1 r;  z += c'     each ends with a grid barrier
//   CG pass           w = A z + the three dot products                ends with the ONE allreduce (grid barrier + mailboxes)
//   update            the four vector recurrences + the first polynomial term of the next iteration, ends with a grid barrier
// Multi-GPU: whoever produces an entry of the next pass's input vector also stores it into the neighbours' ghost segments
// (peer memory over NVLink) — the update phase for z / c0, the epilogue of an inner pass for c' — and the CTA that releases
// the grid barrier publishes the halo epoch.  Receivers wait per WARP, and only the warps that claimed a boundary slice;
// interior slices never wait.  Slices are claimed dynamically, so ranks and SMs that are late (halo, clocks) do not set the pace,
// and all sums are formed in a fixed order (kernels_sell.cuh), so every rank takes bitwise the same decisions.

struct Pcg2Args {
  sell::Mat m64, m32;  // m32.panels == nullptr: the inner passes stream m64
  sell::Work work;     // work.claim points at SIX counters: [0..2] claims (pass number mod 3), [3..5] boundary slices completed in that pass
  const double* invdiag;
  double *x, *g, *d, *s, *w, *z, *r, *c0, *c1;
  int64_t n, n_interior;
  CgState* state;
  int degree;                 // 1 = Jacobi
  double inv_theta;           // first polynomial term c0 = (1/theta) D^-1 g
  double k1[8], k2[8];        // coefficients of inner pass j (1-based): c' = k1[j] c + k2[j] D^-1 r
  unsigned* tickets;          // [1] grid barrier
  int* bar_flag;
  int* abort;
  unsigned long long* trace;  // diagnostic (PE_PCG_TRACE), else nullptr: 8 words per warp — inner pass 1: start, end of own stream, barrier passed,
                              // (smid << 32 | slices); the CG pass: the same four.  Overwritten every iteration: the last one survives.
  unsigned long long* timing; // [0] ns in CG passes (CTA 0), [1] CG passes, [2] ns in inner passes, [3] inner passes, [4] ns in updates,
                              // [5] ns in allreduces, [6..9] ns CTA 0 waited at the barrier behind inner passes / behind the CG pass / for the
                              // peer mailboxes / at the barrier behind updates, [10] halo exchanges posted, [11] reductions posted
  char* const* peer;
  int nranks, me, red_epoch0;
  int n_neigh, field, halo_epoch0;
  const int32_t* neigh_rank;
  const int32_t* push_ptr;    // per boundary row (row - n_interior): its entries in push_dest / push_nb
  const int32_t* push_dest;   // index in the receiver's vector
  const int32_t* push_nb;     // neighbour slot
  const unsigned long long* push_addr;  // the same entry resolved: address of the ghost copy in the receiver's first work vector
  const int32_t* push_src;    // lane * B + component of the entry's row inside its slice
  size_t ctrl_bytes, off_z, off_c0, off_c1;  // offsets (in doubles) of the exchanged vectors inside every rank's work area
  int max_iterations;
};

template <int B, typename TI>
__global__ void __launch_bounds__(sell::THREADS, 1) k_pcg2(Pcg2Args a) {
  extern __shared__ __align__(128) char sell_smem[];
  __shared__ double s_buf[4 * 32];
  __shared__ double s_tot[4];
  __shared__ int s_ok;
  __shared__ int s_halo_seen;  // newest halo epoch a warp of this CTA has waited for AND acquired (system-scope fence)
  CgState* st = a.state;
  if (st->done) return;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t gtid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, gsize = (int64_t)gridDim.x * blockDim.x;
  if (threadIdx.x == 0) s_halo_seen = a.halo_epoch0;
  __syncthreads();
  sell::Ring R = sell::ring_setup(sell_smem, warp, lane);
  const uint64_t policy = sell::evict_first_policy();
  const double tol = st->tol;
  const int max_it = st->max_it;
  const P2PControl* my_ctl = reinterpret_cast<const P2PControl*>(a.peer[a.me]);
  int bar_epoch = pe_ld_flag(a.bar_flag);  // the same value in every CTA: the flag only moves inside barriers
  int halo_seq = 0;                        // halo exchanges published so far (identical on every rank)
  int pass = 0;                            // matrix passes so far (parity selects the claim counter)
  const bool timer = blockIdx.x == 0 && threadIdx.x == 0;
  unsigned long long t_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0}, n_cg = 0, n_in = 0, t_last = 0, t_mark = 0;
  auto now = [&]() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
  };
  auto lap = [&](int slot) {
    if (timer) {
      const unsigned long long t = now();
      t_acc[slot] += t - t_last;
      t_last = t;
    }
  };
  // second clock for the waits INSIDE a phase (slots 4..7: barrier behind an inner pass, barrier behind the CG pass, peer
  // mailboxes, barrier behind the update) — what CTA 0 spends waiting for the slowest warp of the grid / the slowest rank
  unsigned long long* tr = a.trace ? a.trace + ((size_t)blockIdx.x * sell::WARPS + warp) * 8 : nullptr;
  int tr_slices = 0;
  auto trace = [&](int slot) {
    if (tr && lane == 0) {
      if ((slot & 3) == 3) {
        unsigned smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        tr[slot] = ((unsigned long long)smid << 32) | (unsigned)tr_slices;
        tr_slices = 0;
      } else {
        tr[slot] = now();
      }
    }
  };
  auto mark = [&]() { if (timer) t_mark = now(); };
  auto since_mark = [&](int slot) { if (timer) t_acc[slot] += now() - t_mark; };

  // store v into the ghost copies of (boundary) row `row` on the neighbour ranks (the vector phases: one row per thread)
  auto push = [&](int64_t row, double v, size_t off) {
    const int64_t k = row - a.n_interior;
    for (int e = a.push_ptr[k]; e < a.push_ptr[k + 1]; ++e) reinterpret_cast<double*>(a.push_addr[e])[off] = v;
  };
  // the same for a whole boundary slice of a matrix pass, by the warp that owns it: the slice's entries are contiguous in the
  // table, the lanes take them 32 at a time (one coalesced table load, the value fetched from the lane that holds the row) —
  // a handful of independent stores per lane instead of every lane walking its rows' entries one dependent load after another
  auto push_slice = [&](int slice, const double (&v)[B], size_t off) {
    const int per = 32 * B;
    const int64_t n_b = a.n - a.n_interior;
    const int64_t k0 = (int64_t)(slice - a.m64.first_boundary_slice) * per;
    const int e0 = a.push_ptr[k0], e1 = a.push_ptr[min(k0 + per, n_b)];
    for (int eb = e0; eb < e1; eb += 32) {
      const int e = eb + lane;
      const bool on = e < e1;
      const int src = on ? a.push_src[e] : 0;
      const unsigned long long addr = on ? a.push_addr[e] : 0ull;
      const int sl = src / B, sr = src - sl * B;
      double val = 0.0;
#pragma unroll
      for (int r = 0; r < B; ++r) {
        const double t = __shfl_sync(0xffffffffu, v[r], sl);
        if (r == sr) val = t;
      }
      if (on) reinterpret_cast<double*>(addr)[off] = val;
    }
  };
  const int n_boundary_slices = a.m64.n_slices - min(a.m64.first_boundary_slice, a.m64.n_slices);
  // grid barrier; with `publish` the phase stored a halo: it counts as one more exchange, and unless the phase has already
  // published it itself (`early`: the matrix passes, see boundary_done) the releasing CTA posts its epoch to the neighbours
  auto barrier = [&](bool publish, bool early = false) {
    const bool remote = publish && a.n_neigh > 0 && (!early || n_boundary_slices == 0);
    if (publish) ++halo_seq;
    ++bar_epoch;
    __syncthreads();
    if (threadIdx.x == 0) {
      if (remote) __threadfence_system(); else __threadfence();  // release, cumulative over the CTA
      if (atomicAdd(&a.tickets[1], 1u) == gridDim.x - 1) {
        a.tickets[1] = 0u;
        if (remote) {
          __threadfence_system();
          for (int q = 0; q < a.n_neigh; ++q)
            pe_st_flag(&reinterpret_cast<P2PControl*>(a.peer[a.neigh_rank[q]])->halo_flag[a.field][a.me], a.halo_epoch0 + halo_seq);
        }
        __threadfence();
        pe_st_flag(a.bar_flag, bar_epoch);
      } else {
        pcg_wait(a.bar_flag, bar_epoch, a.abort);
      }
      __threadfence();  // acquire (also drops this SM's stale L1 lines)
    }
    __syncthreads();
  };
  // Early publication of a matrix pass's halo.  Boundary slices are claimed in the middle of a pass (Mat::boundary_early); a
  // warp counts the ones it finishes and, when its claims have moved past them (claims only move forward), fences its lanes'
  // remote stores once and adds its count; the warp that completes the count posts the epoch — inside the pass, not at the
  // barrier behind it.  The neighbours' next pass then finds the halo long delivered, the barriers need no system-scope fence,
  // and the only thing that still couples the ranks' clocks is the one reduction per iteration.  Safe to overwrite the
  // neighbour's ghost segment that early: only boundary slices read ghost columns, and a rank's boundary slices of pass k wait
  // for the neighbours' epoch of pass k-1, i.e. for the neighbours having finished THEIR boundary slices — their last
  // readers of the buffer pass k overwrites.
  int my_boundary = 0;  // boundary slices this warp has finished and not yet reported (warp-uniform)
  auto boundary_flush = [&]() {  // all lanes; `pass` = the running pass
    __syncwarp();
    if (lane == 0) {
      // release at system scope, cumulative over the warp's pushes (ordered by the __syncwarp) — as an atomic, not a fence: a
      // fence is also an acquire and would drop this SM's L1 once per warp and pass
      unsigned old;
      asm volatile("atom.release.sys.global.add.u32 %0, [%1], %2;" : "=r"(old) : "l"(a.work.claim + 3 + pass % 3), "r"((unsigned)my_boundary) : "memory");
      if ((int)old + my_boundary == n_boundary_slices) {
        __threadfence_system();  // acquire of the other warps' releases (once per pass), then publish
        for (int q = 0; q < a.n_neigh; ++q)
          pe_st_flag(&reinterpret_cast<P2PControl*>(a.peer[a.neigh_rank[q]])->halo_flag[a.field][a.me], a.halo_epoch0 + halo_seq + 1);
      }
    }
    my_boundary = 0;
  };
  // first polynomial term from a fresh g_i (Jacobi: the whole preconditioner)
  auto precond_first = [&](int64_t i, double gi) {
    if (a.degree <= 1) {
      const double zi = a.invdiag[i] * gi;
      a.z[i] = zi;
      if (a.n_neigh && i >= a.n_interior) push(i, zi, a.off_z);
    } else {
      // r = g and z = c0 are not stored: the first inner pass reads g and c0 in their place
      const double ci = a.inv_theta * a.invdiag[i] * gi;
      a.c0[i] = ci;
      if (a.n_neigh && i >= a.n_interior) push(i, ci, a.off_c0);
    }
  };
  // The acquire behind a halo flag is a system-scope fence, which drops the whole SM's L1 — all eight warps' cached lines of
  // the gathered vector, not only the ghost segment.  So the first warp of the CTA that needs an epoch waits and fences, then
  // leaves the epoch in shared memory; the others find it there and order themselves behind it with a CTA-scope fence (the
  // L1 is the SM's: lines fetched after the first warp's fence are fresh for every warp).
  bool waited = false;
  auto halo_wait = [&](int slice, const sell::Mat& m) {
    if (a.n_neigh && !waited && slice >= m.first_boundary_slice) {
      const int need = a.halo_epoch0 + halo_seq;
      int seen = 0;
      if (lane == 0) seen = *(volatile int*)&s_halo_seen;
      seen = __shfl_sync(0xffffffffu, seen, 0);
      if ((int)(seen - need) < 0) {
        if (lane < a.n_neigh) pcg_wait(&my_ctl->halo_flag[a.field][a.neigh_rank[lane]], need, a.abort);
        __syncwarp();
        __threadfence_system();  // acquire; drops stale L1 lines of the ghost segment
        if (lane == 0) *(volatile int*)&s_halo_seen = need;  // every writer of an epoch writes the same value
      } else {
        __threadfence_block();
      }
      waited = true;
    }
  };
  // Three claim counters in rotation: pass k claims from counter k % 3.  The stream of pass k+1 is begun (first claim included)
  // at the END of pass k, before the barrier, so its counter must be clean by then: it is zeroed at the start of pass k-1,
  // when it has been idle since pass k-2 ended, and that store is published by the barrier between passes k-1 and k.
  auto begin_pass = [&]() {
    waited = false;
    if (blockIdx.x == 0 && threadIdx.x == 0) {
      a.work.claim[(pass + 2) % 3] = 0u;
      a.work.claim[3 + (pass + 2) % 3] = 0u;  // the previous pass's count of boundary slices
    }
  };
  const sell::Mat& m_in = a.m32.panels ? a.m32 : a.m64;
  sell::Stream S;  // the upcoming pass, begun ahead of the barrier in front of it
  auto begin_stream = [&](bool inner) {
    unsigned* ctr = a.work.claim + (pass % 3);
    if (inner && a.m32.panels) sell::stream_begin<B, TI>(S, a.m32, ctr, (int)blockIdx.x, (int)gridDim.x, warp, R, lane, policy);
    else sell::stream_begin<B, double>(S, a.m64, ctr, (int)blockIdx.x, (int)gridDim.x, warp, R, lane, policy);
  };

  // ---- prologue: z (or c0, r) from the start residual; publish its halo
  for (int64_t i = gtid; i < a.n; i += gsize) precond_first(i, a.g[i]);
  begin_stream(a.degree > 1);
  barrier(true);
  if (timer) t_last = now();

  const int n_chunks = (a.m64.n_slices + a.m64.chunk - 1) / a.m64.chunk;  // units of the deterministic sums of the CG pass
  double gamma_old = 1.0, alpha_old = 1.0;
  bool first = true;
  const int it0 = st->it;  // read before anyone can write it (writes happen only on exit, behind a grid barrier)
  for (int k = 1; k <= a.max_iterations + 1; ++k) {
    // ---- m-1 inner passes of the polynomial preconditioner
    for (int j = 1; j < a.degree; ++j) {
      const double* cin = (j & 1) ? a.c0 : a.c1;
      double* cout = (j & 1) ? a.c1 : a.c0;
      const size_t off_out = (j & 1) ? a.off_c1 : a.off_c0;
      const bool last_inner = j == a.degree - 1;
      const double k1 = a.k1[j], k2 = a.k2[j];
      double pr[B], pc[B], pi[B], pz[B];
      begin_pass();
      auto ready = [&](int slice) { halo_wait(slice, m_in); };
      auto pre = [&](int slice) {
        const int64_t brow = (int64_t)slice * 32 + lane;
        if (brow < m_in.n_brows) {
#pragma unroll
          for (int r = 0; r < B; ++r) {  // epilogue operands: in flight while the slice streams
            const int64_t row = brow * B + r;
            pc[r] = cin[row]; pi[r] = a.invdiag[row];
            if (j == 1) { pr[r] = a.g[row]; pz[r] = pc[r]; }  // first pass: r = g, z = c0 (never stored as such)
            else { pr[r] = a.r[row]; pz[r] = a.z[row]; }
          }
        }
      };
      auto done = [&](int slice, double (&acc)[B], int) {
        ++tr_slices;
        const int64_t brow = (int64_t)slice * 32 + lane;
        double out[B];  // what the neighbours need of this row: the next pass's input
#pragma unroll
        for (int r = 0; r < B; ++r) out[r] = 0.0;
        if (brow < m_in.n_brows) {
#pragma unroll
          for (int r = 0; r < B; ++r) {
            const int64_t row = brow * B + r;
            const double rn = pr[r] - acc[r];
            const double cn = k1 * pc[r] + k2 * pi[r] * rn;
            const double zn = pz[r] + cn;
            a.r[row] = rn;
            cout[row] = cn;
            a.z[row] = zn;
            out[r] = last_inner ? zn : cn;
          }
        }
        if (a.n_neigh) {
          if (slice >= m_in.first_boundary_slice) {
            push_slice(slice, out, last_inner ? a.off_z : off_out);
            ++my_boundary;
          } else if (my_boundary) {
            boundary_flush();  // first interior slice behind the boundary ones
          }
        }
      };
      if (j == 1) trace(0);
      if (a.m32.panels) sell::stream_run<B, TI>(S, cin, R, lane, policy, ready, pre, done);
      else sell::stream_run<B, double>(S, cin, R, lane, policy, ready, pre, done);
      if (my_boundary) boundary_flush();
      if (j == 1) trace(1);
      ++pass;
      begin_stream(!last_inner);  // the next pass's first copies fly while this warp waits for the slowest one
      mark();
      barrier(true, true);
      since_mark(4);
      if (j == 1) { trace(2); trace(3); } else tr_slices = 0;
      if (timer) ++n_in;
      lap(1);
    }
    // ---- CG pass: w = A z, gamma = g.z, delta = w.z, rho = g.g
    {
      double pg[B], pz[B];
      sell::Pending pend{0u, -1};
      begin_pass();
      auto ready = [&](int slice) { halo_wait(slice, a.m64); };
      auto pre = [&](int slice) {
        const int64_t brow = (int64_t)slice * 32 + lane;
        if (brow < a.m64.n_brows) {
#pragma unroll
          for (int r = 0; r < B; ++r) { pg[r] = a.g[brow * B + r]; pz[r] = a.z[brow * B + r]; }
        }
      };
      double v[3] = {0.0, 0.0, 0.0};  // this lane's contributions to the current chunk's partial sums
      auto done = [&](int slice, double (&acc)[B], int chunk) {
        ++tr_slices;
        const int64_t brow = (int64_t)slice * 32 + lane;
        if (brow < a.m64.n_brows) {
#pragma unroll
          for (int r = 0; r < B; ++r) {
            a.w[brow * B + r] = acc[r];
            v[0] += pg[r] * pz[r];
            v[1] += acc[r] * pz[r];
            v[2] += pg[r] * pg[r];
          }
        }
        if (chunk >= 0) {
          sell::sums_finish<3>(a.work, n_chunks, pend, lane);  // the previous chunk's ticket has long arrived
          pend = sell::sums_post<3>(a.work, chunk, v, lane);
          v[0] = v[1] = v[2] = 0.0;
        }
      };
      trace(4);
      sell::stream_run<B, double>(S, a.z, R, lane, policy, ready, pre, done);
      sell::sums_finish<3>(a.work, n_chunks, pend, lane);
      trace(5);
      ++pass;
      begin_stream(a.degree > 1);  // first pass of the next iteration (drained below if the solve ends here)
    }
    lap(0);
    if (timer) ++n_cg;
    // ---- the one reduction of the iteration: group totals -> (grid barrier) -> every CTA adds them in the same order
    mark();
    barrier(false);
    since_mark(5);
    trace(6);
    trace(7);
    double tot[3];
    sell::sum_groups<3>(a.work, (n_chunks + 31) >> 5, tot, s_buf);
    if (threadIdx.x < 32) {
      bool good = true;
      if (a.nranks > 1) {
        const int e = a.red_epoch0 + k;
        if (blockIdx.x == 0 && lane < a.nranks) {
          P2PControl* ctl = reinterpret_cast<P2PControl*>(a.peer[lane]);
#pragma unroll
          for (int q = 0; q < 3; ++q) ctl->red_val[e & 1][a.me][q] = tot[q];
          __threadfence_system();
          pe_st_flag(&ctl->red_flag[a.me], e);
        }
        mark();
        if (lane < a.nranks) good = pcg_wait(&my_ctl->red_flag[lane], e, a.abort);
        good = __all_sync(0xffffffffu, good);
        since_mark(6);
        __threadfence();  // orders the mailbox loads behind the flag loads (both bypass L1)
        double mail[3];
#pragma unroll
        for (int q = 0; q < 3; ++q) mail[q] = (good && lane < a.nranks) ? pe_ld_mail(&my_ctl->red_val[e & 1][lane][q]) : 0.0;
#pragma unroll
        for (int q = 0; q < 3; ++q) {
          double sum = 0.0;
          for (int rk = 0; rk < a.nranks; ++rk) sum += __shfl_sync(0xffffffffu, mail[q], rk);
          tot[q] = sum;
        }
      }
      if (lane == 0) {
        s_tot[0] = tot[0]; s_tot[1] = tot[1]; s_tot[2] = tot[2];
        s_ok = (good && !pe_ld_flag(a.abort)) ? 1 : 0;
      }
    }
    __syncthreads();
    const double gamma = s_tot[0], delta = s_tot[1], rho = s_tot[2];
    const bool ok = s_ok != 0;
    lap(3);
    // ---- SolverControl::check on the iterate the pass started from (it has seen k-1 updates)
    const int done_its = it0 + k - 1;
    const double res = sqrt(rho);
    const bool converged = ok && res <= tol;
    const bool failed = !ok || isnan(res) || (!converged && done_its >= max_it);
    if (converged || failed) {
      if (blockIdx.x == 0 && threadIdx.x == 0) {
        st->it = done_its;
        st->res = res;
        if (!ok) st->pad = 1;
        st->done = converged ? 1 : -1;
        a.timing[11] = (unsigned long long)k;
      }
      sell::stream_drain(S, R);
      break;
    }
    // ---- update: d, s, x, g and the first polynomial term of the next iteration (stores the next halo)
    const double beta = first ? 0.0 : gamma / gamma_old;
    const double alpha = first ? gamma / delta : gamma / (delta - beta * gamma / alpha_old);
    // four rows per thread and trip: all 28 loads are issued before the first use (one CTA per SM has only 8 warps to
    // hide DRAM latency with)
    // boundary rows first (multi-GPU): their remote stores are long acknowledged when the system-scope fence of the barrier
    // behind this phase asks for them
    const int64_t rot = a.n_neigh ? a.n_interior : 0;
    auto rotated = [&](int64_t i) { const int64_t r = i + rot; return r >= a.n ? r - a.n : r; };
    for (int64_t i0 = gtid; i0 < a.n; i0 += 4 * gsize) {
      double zi[4], wi[4], di[4], si[4], xi[4], gi[4], vi[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const bool in = i0 + u * gsize < a.n;
        const int64_t i = in ? rotated(i0 + u * gsize) : 0;
        zi[u] = in ? a.z[i] : 0.0;
        wi[u] = in ? a.w[i] : 0.0;
        di[u] = (in && !first) ? a.d[i] : 0.0;
        si[u] = (in && !first) ? a.s[i] : 0.0;
        xi[u] = in ? a.x[i] : 0.0;
        gi[u] = in ? a.g[i] : 0.0;
        vi[u] = in ? a.invdiag[i] : 0.0;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (i0 + u * gsize < a.n) {
          const int64_t i = rotated(i0 + u * gsize);
          const double dn = beta * di[u] - zi[u], sn = beta * si[u] - wi[u];
          const double gn = gi[u] + alpha * sn;
          a.d[i] = dn;
          a.s[i] = sn;
          a.x[i] = xi[u] + alpha * dn;
          a.g[i] = gn;
          if (a.degree <= 1) {
            const double zn = vi[u] * gn;
            a.z[i] = zn;
            if (a.n_neigh && i >= a.n_interior) push(i, zn, a.off_z);
          } else {
            const double cn = a.inv_theta * vi[u] * gn;
            a.c0[i] = cn;
            if (a.n_neigh && i >= a.n_interior) push(i, cn, a.off_c0);
          }
        }
      }
    }
    gamma_old = gamma;
    alpha_old = alpha;
    first = false;
    mark();
    barrier(true);
    since_mark(7);
    lap(2);
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    for (int q = 0; q < 6; ++q) a.work.claim[q] = 0u;
    a.timing[0] += t_acc[0];
    a.timing[1] += n_cg;
    a.timing[2] += t_acc[1];
    a.timing[3] += n_in;
    a.timing[4] += t_acc[2];
    a.timing[5] += t_acc[3];
    a.timing[6] += t_acc[4];
    a.timing[7] += t_acc[5];
    a.timing[8] += t_acc[6];
    a.timing[9] += t_acc[7];
    a.timing[10] = (unsigned long long)halo_seq;
  }
}
