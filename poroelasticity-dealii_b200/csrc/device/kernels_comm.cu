// kernels_comm.cu — K12: ghost-dof halo exchange and the global reductions of CG.
//
// The reference is serial (Vector<double>, SparseMatrix<double>: lib/include/PoroElasticPressureSolver.h:36-44);
// this is the communication of the new cell-partitioned build.  Two transports:
//   * NCCL (grouped ncclSend/ncclRecv, ncclAllReduce) — setup and the handful of exchanges per time step
//     outside the CG loop (p, t1, u before cell kernels, the initial guess of a solve);
//   * peer memory over NVLink/NVSwitch — everything inside the CG loop.  A CG iteration at 8 GPUs is ~0.2 ms
//     of SpMV and needs one halo exchange and two reductions; NCCL's launch+protocol latency per call
//     (tens of microseconds) would eat the strong-scaling target.  Here the SENDER's kernel stores the halo
//     values straight into the ghost segment of the receiver's vector (cudaIpc-mapped), fences at system
//     scope and bumps an epoch flag in the receiver's control block; reductions go through per-sender
//     mailboxes, every rank sums the mailboxes in rank order (bitwise identical result on all ranks).
// All waits are bounded: a peer that never arrives flags CgState and the solve returns PE_ERR_NCCL.
#include <cstdlib>
#include <cstring>

#include "pe_internal.cuh"

namespace {

constexpr int T = 256;

__global__ void k_pack(int64_t n, const int32_t* __restrict__ idx, const double* __restrict__ v, double* __restrict__ buf) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) buf[i] = v[idx[i]];
}

// Sender side of a halo exchange: v_off = offset (in doubles) of the vector inside the work area of every region.
__global__ void __launch_bounds__(T)
k_halo_send(int64_t n_send, const int32_t* __restrict__ send_idx, const int32_t* __restrict__ send_dest, const int32_t* __restrict__ send_nb,
            const int32_t* __restrict__ neigh_rank, int n_neigh, char* const* __restrict__ peer, size_t ctrl_bytes, size_t v_off,
            const double* __restrict__ v, int field, int me, int epoch, unsigned* __restrict__ ticket, const CgState* __restrict__ state) {
  if (state && state->done) return;
  __shared__ bool s_last;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_send; i += (int64_t)gridDim.x * blockDim.x) {
    double* dst = reinterpret_cast<double*>(peer[neigh_rank[send_nb[i]]] + ctrl_bytes) + v_off;
    dst[send_dest[i]] = v[send_idx[i]];
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
  __syncthreads();
  if (s_last) {  // every block's stores are fenced: publish
    __threadfence_system();
    for (int k = threadIdx.x; k < n_neigh; k += blockDim.x) {
      P2PControl* ctl = reinterpret_cast<P2PControl*>(peer[neigh_rank[k]]);
      pe_st_flag(&ctl->halo_flag[field][me], epoch);
    }
    if (threadIdx.x == 0) *ticket = 0u;
  }
}

// Receiver side: returns once every neighbour's halo of this epoch has landed in my ghost segment.
__global__ void k_halo_wait(const P2PControl* __restrict__ ctl, const int32_t* __restrict__ neigh_rank, int n_neigh, int field, int epoch,
                            CgState* __restrict__ state, int* __restrict__ comm_err) {
  if (state && state->done) return;
  bool ok = true;
  for (int k = threadIdx.x; k < n_neigh; k += blockDim.x) ok = pe_wait_flag(&ctl->halo_flag[field][neigh_rank[k]], epoch) && ok;
  __threadfence_system();
  if (!ok) {  // a neighbour never delivered: the ghost values are stale.  Always reported (cold path: through the error word)
    *comm_err = 1;
    if (state) { state->pad = 1; state->done = -1; }
    __threadfence_system();
  }
}

// Allreduce(sum) of `count` doubles in place: post to every rank's mailbox, wait for everyone, sum in rank order.
__global__ void k_allreduce_p2p(double* __restrict__ vals, int count, char* const* __restrict__ peer, int nranks, int me, int epoch,
                                CgState* __restrict__ state, int* __restrict__ comm_err) {
  if (state && state->done) return;
  const int r = threadIdx.x;
  const int par = epoch & 1;
  if (r < nranks) {
    P2PControl* ctl = reinterpret_cast<P2PControl*>(peer[r]);
    for (int k = 0; k < count; ++k) ctl->red_val[par][me][k] = vals[k];
    __threadfence_system();
    pe_st_flag(&ctl->red_flag[me], epoch);
  }
  const P2PControl* mine = reinterpret_cast<const P2PControl*>(peer[me]);
  bool ok = true;
  if (r < nranks) ok = pe_wait_flag(&mine->red_flag[r], epoch);
  ok = __all_sync(0xffffffffu, ok);
  __threadfence_system();
  if (!ok) {  // a peer never posted: never leave the rank-local partial behind as if it were the global sum
    if (r < count) vals[r] = __longlong_as_double(0x7ff8000000000000LL);
    if (r == 0) {
      *comm_err = 2;
      if (state) { state->pad = 1; state->done = -1; }
      __threadfence_system();
    }
    return;
  }
  if (r < count) {
    double s = 0.0;
    for (int q = 0; q < nranks; ++q) s += pe_ld_mail(&mine->red_val[par][q][r]);
    vals[r] = s;
  }
}

bool in_work_area(const pe_ctx* c, const double* v, size_t* off) {
  const char* w0 = c->p2p.region + c->p2p.ctrl_bytes;
  const char* p = reinterpret_cast<const char*>(v);
  if (!c->p2p.region || p < w0 || p >= c->p2p.region + c->p2p.region_bytes) return false;
  *off = (size_t)(p - w0) / sizeof(double);
  return true;
}

}  // namespace

void pe_pack_launch(pe_ctx* c, int64_t n, const int32_t* idx, const double* v, double* buf) {
  if (!n) return;
  k_pack<<<pe_div_up(n, T), T, 0, c->stream>>>(n, idx, v, buf);
  c->st.kernel_launches++;
}

void pe_sync_checked(pe_ctx* c) {
  PE_CUDA(cudaStreamSynchronize(c->stream));
  if (c->h_comm_err && *c->h_comm_err) {
    const int code = *c->h_comm_err;
    *c->h_comm_err = 0;
    throw PeError(PE_ERR_NCCL, code == 1 ? "peer-memory halo exchange timed out (a neighbour rank never delivered its ghost values)"
                                         : "peer-memory reduction timed out (a rank never posted its partial sum)");
  }
}

void pe_allreduce_sum(pe_ctx* c, double* dev, int count, bool in_solve) {
  if (c->nranks <= 1) return;
  if (c->p2p.on && count <= 4) {
    const int epoch = (int)(++c->p2p.red_epoch);
    k_allreduce_p2p<<<1, 32, 0, c->stream>>>(dev, count, c->p2p.d_peer.p, c->nranks, c->rank, epoch, in_solve ? c->cg_state.p : nullptr, c->h_comm_err);
    c->st.kernel_launches++;
    return;
  }
  PE_NCCL(ncclAllReduce(dev, dev, count, ncclDouble, ncclSum, c->comm_nccl(), c->stream));
}

// A rank-local yes/no (did my copy of the matrix build in format X?) turned into a decision every rank shares.  The solver
// paths that follow from such a decision speak different peer-memory protocols, so ranks must never choose on their own:
// a small rank whose sliced copy was rejected for padding next to ranks that kept theirs would wait for messages that
// never come.  Collective; cold path (setup / first assembly / dt change).
bool pe_all_ranks_agree(pe_ctx* c, bool mine) {
  if (c->nranks <= 1) return mine;
  c->h_scalars[0] = mine ? 0.0 : 1.0;
  PE_CUDA(cudaMemcpyAsync(c->red.out.p, c->h_scalars, sizeof(double), cudaMemcpyHostToDevice, c->stream));
  pe_allreduce_sum(c, c->red.out.p, 1);
  PE_CUDA(cudaMemcpyAsync(c->h_scalars, c->red.out.p, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  pe_sync_checked(c);
  return c->h_scalars[0] == 0.0;
}

// ghost values of v (entries [n_owned, n_local)) <- owners
void pe_halo_exchange(pe_ctx* c, Field& F, double* v, bool in_solve, int* fused_epoch) {
  if (fused_epoch) *fused_epoch = 0;
  if (c->nranks <= 1 || F.halo.n_neigh == 0) return;
  Halo& H = F.halo;
  const int64_t ns = H.n_send();
  size_t v_off = 0;
  const int fi = &F == &c->fu ? 1 : 0;
  if (c->p2p.on && in_work_area(c, v, &v_off)) {
    P2PField& P = c->p2p.f[fi];
    const int epoch = (int)(++P.epoch);
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((ns + T - 1) / T, 4 * c->sm_count));
    k_halo_send<<<grid, T, 0, c->stream>>>(ns, H.send_idx.p, P.send_dest.p, P.send_nb.p, P.neigh_rank.p, H.n_neigh, c->p2p.d_peer.p,
                                           c->p2p.ctrl_bytes, v_off, v, fi, c->rank, epoch, c->p2p.ticket.p, in_solve ? c->cg_state.p : nullptr);
    c->st.kernel_launches++;
    if (fused_epoch) {
      *fused_epoch = epoch;
      return;
    }
    k_halo_wait<<<1, 32, 0, c->stream>>>(reinterpret_cast<const P2PControl*>(c->p2p.region), P.neigh_rank.p, H.n_neigh, fi, epoch, in_solve ? c->cg_state.p : nullptr, c->h_comm_err);
    c->st.kernel_launches++;
    return;
  }
  if (c->p2p.on && !in_solve) {
    // Vectors outside the exported work area (p, t1, u, an initial guess): staged through one of TWO exported buffers.  A fast
    // neighbour's next exchange lands in the other buffer, and the one after that needs this rank's next send, which sits
    // behind this copy-back on the stream — so nothing is overwritten before it has been read.  No NCCL in the time loop.
    double* w = (c->p2p.stage_epoch++ & 1u) ? c->w_x1.p : c->w_x0.p;
    PE_CUDA(cudaMemcpyAsync(w, v, (size_t)F.n_owned * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
    pe_halo_exchange(c, F, w, false, nullptr);
    PE_CUDA(cudaMemcpyAsync(v + F.n_owned, w + F.n_owned, (size_t)(F.n_local - F.n_owned) * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
    return;
  }
  pe_pack_launch(c, ns, H.send_idx.p, v, H.send_buf.p);
  PE_NCCL(ncclGroupStart());
  for (int k = 0; k < H.n_neigh; ++k) {
    const int64_t s0 = H.send_ptr[k], s1 = H.send_ptr[k + 1], r0 = H.recv_ptr[k], r1 = H.recv_ptr[k + 1];
    if (s1 > s0) PE_NCCL(ncclSend(H.send_buf.p + s0, (size_t)(s1 - s0), ncclDouble, H.rank[k], c->comm_nccl(), c->stream));
    if (r1 > r0) PE_NCCL(ncclRecv(v + F.n_owned + r0, (size_t)(r1 - r0), ncclDouble, H.rank[k], c->comm_nccl(), c->stream));
  }
  PE_NCCL(ncclGroupEnd());
}

void pe_comm_release(pe_ctx* c) {
  P2P& M = c->p2p;
  for (int r = 0; r < (int)M.peer.size(); ++r)
    if (M.peer[r] && r != c->rank) cudaIpcCloseMemHandle(M.peer[r]);
  M.peer.clear();
  if (M.region) cudaFree(M.region);
  M.region = nullptr;
  M.on = false;
}

// What every rank tells every other rank at setup, exchanged with ONE kind of NCCL collective (ncclAllGather of a
// fixed-size record).  NCCL connects lazily per collective algorithm and per peer-to-peer channel, seconds each at 8 ranks;
// the earlier sequence (allreduce, allgather, grouped send/recv per field, allreduce) paid that three times inside pe_setup.
struct SetupRecord {
  long long stride;                        // work-vector length this rank needs
  cudaIpcMemHandle_t handle;               // its region
  int ok;                                  // could export / open so far
  long long land[2][PE_P2P_MAX_RANKS];     // [field][sender rank]: where that rank's halo values start in MY vectors (-1: not a neighbour)
};

static std::vector<SetupRecord> gather_records(pe_ctx* c, const SetupRecord& mine) {
  cudaStream_t s = c->stream;
  DBuf<char> all;
  all.alloc((size_t)c->nranks * sizeof(SetupRecord));
  PE_CUDA(cudaMemcpyAsync(all.p + (size_t)c->rank * sizeof(SetupRecord), &mine, sizeof(SetupRecord), cudaMemcpyHostToDevice, s));
  PE_NCCL(ncclAllGather(all.p + (size_t)c->rank * sizeof(SetupRecord), all.p, sizeof(SetupRecord), ncclChar, c->comm_nccl(), s));
  std::vector<SetupRecord> out((size_t)c->nranks);
  PE_CUDA(cudaMemcpyAsync(out.data(), all.p, all.n, cudaMemcpyDeviceToHost, s));
  PE_CUDA(cudaStreamSynchronize(s));
  return out;
}

void pe_comm_setup(pe_ctx* c, size_t n_work) {
  pe_comm_release(c);
  P2P& M = c->p2p;
  cudaStream_t s = c->stream;
  M.ctrl_bytes = (sizeof(P2PControl) + 255) / 256 * 256;
  const char* env = std::getenv("PE_COMM");
  const bool want = c->nranks > 1 && !(env && std::strcmp(env, "nccl") == 0) && c->nranks <= PE_P2P_MAX_RANKS;
  // every rank must use the same vector stride so that offsets mean the same thing in every region
  long long stride = (long long)((n_work + 31) / 32 * 32);
  SetupRecord mine;
  std::memset(&mine, 0, sizeof mine);
  std::vector<SetupRecord> all;
  if (c->nranks > 1) {
    mine.stride = stride;
    mine.ok = 1;
    for (int fi = 0; fi < 2; ++fi) {
      Field& F = fi ? c->fu : c->fp;
      for (int r = 0; r < PE_P2P_MAX_RANKS; ++r) mine.land[fi][r] = -1;
      if (c->nranks <= PE_P2P_MAX_RANKS)
        for (int k = 0; k < F.halo.n_neigh; ++k) mine.land[fi][F.halo.rank[k]] = (long long)F.n_owned + F.halo.recv_ptr[k];
    }
    all = gather_records(c, mine);  // round 1: strides (the region cannot be allocated before the common stride is known)
    for (const SetupRecord& r : all) stride = std::max(stride, r.stride);
  }
  M.region_bytes = M.ctrl_bytes + (size_t)PE_WORK_VECTORS * stride * sizeof(double);
  PE_CUDA(cudaMalloc((void**)&M.region, M.region_bytes));
  PE_CUDA(cudaMemsetAsync(M.region, 0, M.region_bytes, s));
  double* w0 = reinterpret_cast<double*>(M.region + M.ctrl_bytes);
  pe_ctx::WPtr* ws[PE_WORK_VECTORS] = {&c->w_g, &c->w_h, &c->w_d, &c->w_z, &c->w_d2, &c->w_r, &c->w_s, &c->w_c1, &c->w_x0, &c->w_x1};
  for (int k = 0; k < PE_WORK_VECTORS; ++k) ws[k]->p = w0 + (size_t)k * stride;
  M.peer.assign(c->nranks, nullptr);
  M.peer[c->rank] = M.region;
  M.on = false;
  M.stage_epoch = 0;
  M.d_peer.upload(M.peer, s);  // one rank: the persistent CG kernels post to their own mailbox
  if (want) {
    // Every step that can fail for environmental reasons (IPC disabled in the container, no peer access between two devices)
    // is followed by a collective look at everybody's flag, so either every rank uses the peer-memory transport or every
    // rank stays on NCCL.
    PE_CUDA(cudaStreamSynchronize(s));
    mine.ok = cudaIpcGetMemHandle(&mine.handle, M.region) == cudaSuccess ? 1 : 0;
    (void)cudaGetLastError();
    if (env && std::strcmp(env, "ipcfail") == 0 && c->rank == c->nranks - 1) mine.ok = 0;  // test hook: one rank cannot export
    all = gather_records(c, mine);  // round 2: handles + landing offsets + export votes
    bool ok = true;
    for (const SetupRecord& r : all) ok = ok && r.ok;
    if (ok) {
      for (int r = 0; r < c->nranks && ok; ++r)
        if (r != c->rank) ok = cudaIpcOpenMemHandle((void**)&M.peer[r], all[r].handle, cudaIpcMemLazyEnablePeerAccess) == cudaSuccess;
      (void)cudaGetLastError();
    }
    mine.ok = ok ? 1 : 0;
    std::vector<SetupRecord> votes = gather_records(c, mine);  // round 3: open votes; also the barrier behind the zeroed regions
    for (const SetupRecord& r : votes) ok = ok && r.ok;
    if (!ok) {  // fall back to NCCL inside the CG loop as well
      for (int r = 0; r < c->nranks; ++r)
        if (r != c->rank && M.peer[r]) { cudaIpcCloseMemHandle(M.peer[r]); M.peer[r] = nullptr; }
      (void)cudaGetLastError();
      std::fprintf(stderr, "[poroel rank %d] peer-memory transport unavailable (cudaIpc), using NCCL inside the CG loop\n", c->rank);
      PE_CUDA(cudaStreamSynchronize(s));
      return;
    }
    M.d_peer.upload(M.peer, s);
    M.ticket.alloc_zero(1, s);
    for (int fi = 0; fi < 2; ++fi) {
      Field& F = fi ? c->fu : c->fp;
      Halo& H = F.halo;
      P2PField& P = M.f[fi];
      P.epoch = 0;
      const int nn = H.n_neigh;
      // neighbour k told everybody where MY values land inside ITS vectors: all[rank k].land[fi][me]
      std::vector<int32_t> dest((size_t)H.n_send()), nb((size_t)H.n_send()), nr(H.rank.begin(), H.rank.end());
      for (int k = 0; k < nn; ++k) {
        const long long theirs = all[H.rank[k]].land[fi][c->rank];
        if (theirs < 0 && H.send_ptr[k + 1] > H.send_ptr[k]) throw PeError(PE_ERR_BAD_INPUT, "halo plans of two ranks do not match");
        for (int64_t i = H.send_ptr[k]; i < H.send_ptr[k + 1]; ++i) {
          dest[i] = (int32_t)(theirs + (i - H.send_ptr[k]));
          nb[i] = k;
        }
      }
      P.send_dest.upload(dest, s);
      P.send_nb.upload(nb, s);
      P.neigh_rank.upload(nr, s);
      // the same entries grouped by source row (counting sort; order inside a row = neighbour order)
      {
        const int64_t n_b = F.n_owned - F.n_interior, ns = H.n_send();
        std::vector<int32_t> h_idx((size_t)ns);
        PE_CUDA(cudaMemcpyAsync(h_idx.data(), H.send_idx.p, (size_t)ns * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
        PE_CUDA(cudaStreamSynchronize(s));
        std::vector<int32_t> ptr((size_t)n_b + 1, 0), pdest((size_t)ns), pnb((size_t)ns);
        bool all_boundary = true;
        for (int64_t i = 0; i < ns; ++i) {
          if (h_idx[i] < F.n_interior) { all_boundary = false; break; }
          ptr[h_idx[i] - F.n_interior + 1]++;
        }
        P.push_ok = pe_all_ranks_agree(c, all_boundary);  // the persistent kernel is chosen by all ranks or by none
        if (P.push_ok) {
          for (int64_t r = 0; r < n_b; ++r) ptr[r + 1] += ptr[r];
          std::vector<int32_t> fill(ptr.begin(), ptr.end() - 1);
          std::vector<unsigned long long> paddr((size_t)ns);
          std::vector<int32_t> psrc((size_t)ns);
          const int64_t per_slice = 32 * (int64_t)F.ncomp;
          for (int64_t i = 0; i < ns; ++i) {
            const int64_t rel = h_idx[i] - F.n_interior;
            const int32_t at = fill[rel]++;
            pdest[at] = dest[i];
            pnb[at] = nb[i];
            // resolved once: the kernels store through one table load instead of chasing neighbour slot -> rank -> region
            paddr[at] = (unsigned long long)(uintptr_t)(M.peer[H.rank[nb[i]]] + M.ctrl_bytes) + 8ull * (unsigned long long)dest[i];
            psrc[at] = (int32_t)(rel % per_slice);
          }
          P.push_ptr.upload(ptr, s);
          P.push_dest.upload(pdest, s);
          P.push_nb.upload(pnb, s);
          P.push_addr.upload(paddr, s);
          P.push_src.upload(psrc, s);
        }
      }
    }
    M.red_epoch = 0;
    M.on = true;
  }
  PE_CUDA(cudaStreamSynchronize(s));
}
