// pe_internal.cuh — internal layout of the device library (not part of the ABI).
#pragma once
#include <cuda_runtime.h>
#include <nccl.h>

#include <cstdint>
#include <cstdio>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../../include/poroel.h"

#define PE_CUDA(call)                                                                                   \
  do {                                                                                                  \
    cudaError_t e_ = (call);                                                                            \
    if (e_ != cudaSuccess)                                                                              \
      throw PeError(PE_ERR_CUDA, std::string(#call) + " failed: " + cudaGetErrorString(e_) + " at " + __FILE__ + ":" + std::to_string(__LINE__)); \
  } while (0)
#define PE_NCCL(call)                                                                                   \
  do {                                                                                                  \
    ncclResult_t r_ = (call);                                                                           \
    if (r_ != ncclSuccess)                                                                              \
      throw PeError(PE_ERR_NCCL, std::string(#call) + " failed: " + ncclGetErrorString(r_) + " at " + __FILE__ + ":" + std::to_string(__LINE__)); \
  } while (0)

struct PeError : std::runtime_error {
  int code;
  PeError(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

template <class T>
struct DBuf {
  T* p = nullptr;
  size_t n = 0;
  DBuf() = default;
  DBuf(const DBuf&) = delete;
  DBuf& operator=(const DBuf&) = delete;
  ~DBuf() { release(); }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    n = 0;
  }
  void alloc(size_t count) {
    release();
    n = count;
    if (count) PE_CUDA(cudaMalloc((void**)&p, count * sizeof(T)));
  }
  void alloc_zero(size_t count, cudaStream_t s) {
    alloc(count);
    if (count) PE_CUDA(cudaMemsetAsync(p, 0, count * sizeof(T), s));
  }
  void upload(const T* h, size_t count, cudaStream_t s) {
    alloc(count);
    if (count) PE_CUDA(cudaMemcpyAsync(p, h, count * sizeof(T), cudaMemcpyHostToDevice, s));
  }
  void upload(const std::vector<T>& h, cudaStream_t s) { upload(h.data(), h.size(), s); }
};

// reference-cell tables of one (finite element, quadrature) pair, all in device memory
struct ShapeTab {
  int ns = 0, nq = 0;
  DBuf<double> N;   // nq*ns
  DBuf<double> dN;  // nq*ns*dim
};
struct QuadTab {
  int nq = 0;
  DBuf<double> w;     // nq
  DBuf<double> geoN;  // nq*vpc   (Q1 mapping shape values)
  DBuf<double> geodN; // nq*vpc*dim
};

struct Halo {
  int n_neigh = 0;
  std::vector<int> rank;
  std::vector<int64_t> send_ptr, recv_ptr;
  DBuf<int32_t> send_idx;
  DBuf<double> send_buf;
  int64_t n_send() const { return send_ptr.empty() ? 0 : send_ptr.back(); }
  void reset() {
    n_neigh = 0;
    rank.clear();
    send_ptr.clear();
    recv_ptr.clear();
    send_idx.release();
    send_buf.release();
  }
};

// Sliced block-ELL copy of a matrix for the TMA-fed CG kernels (layout and rationale: kernels_sell.cuh).
struct SellMat {
  const double* src = nullptr;  // value array this copy was built from (look-up key together with f32)
  bool f32 = false;             // values stored as float (read only inside the Chebyshev preconditioner)
  int B = 0, panel_bytes = 0;
  int n_slices = 0, first_boundary_slice = 0;
  int64_t n_brows = 0, n_panels = 0, nnzb = 0;
  DBuf<int32_t> slice_ptr;      // n_slices + 1 (in panels)
  DBuf<char> panels;
};

struct Field {
  int degree = 1, ncomp = 1, ns = 0, nloc = 0;
  int64_t n_owned = 0, n_local = 0;
  bool have_dofs = false, have_partition = false;
  std::vector<int32_t> h_cell_dofs;
  DBuf<int32_t> cell_dofs;
  // pure-Dirichlet constraint lines
  int64_t n_lines = 0;
  std::vector<int32_t> h_line_dof;
  std::vector<double> h_line_g;
  DBuf<int32_t> cline;     // n_local: line index or -1
  DBuf<int32_t> line_dof;  // n_lines
  DBuf<double> line_g;     // n_lines
  // Hanging-node lines x_h = sum_m w_hm x_m + g_h (adaptive meshes, PS:74-77 / DS:112-114); masters are free dofs.
  // The cell kernels treat hanging dofs like free ones; kernels_constraints.cu condenses matrices / vectors afterwards.
  struct Hanging {
    int64_t n = 0, n_entries = 0, n_masters = 0;
    bool any_g = false;
    std::vector<int32_t> h_dof, h_ptr, h_edof;  // host copies (lines sorted by dof; ptr has n+1 entries)
    std::vector<double> h_w, h_g;
    DBuf<int32_t> hline;                         // n_local: hanging line of a dof or -1
    DBuf<int32_t> dof, ptr, edof;
    DBuf<double> w, g;
    DBuf<int32_t> tline_of;                      // n_local: slot in the transposed table or -1
    DBuf<int32_t> t_master, t_ptr, t_line;       // per master: the hanging lines that refer to it ...
    DBuf<double> t_w;                            // ... and their weights
    void clear() {
      n = n_entries = n_masters = 0;
      any_g = false;
      h_dof.clear(); h_ptr.clear(); h_edof.clear(); h_w.clear(); h_g.clear();
    }
  } hang;
  // CSR pattern of the owned rows, columns ascending (local ids)
  DBuf<int32_t> rowptr, col;
  int64_t nnz = 0;
  int64_t n_interior = 0;  // rows [0, n_interior) reference no ghost column (32-row aligned); == n_owned on one rank
  Halo halo;
  // Block-CSR copy of the vector-valued matrix (dim x dim blocks: all components of two nodes couple, DS:143-145).
  // bval is component-major inside a block row: value k of block j of block row I sits at 9*bptr[I] + k*nb_I + j,
  // so the lanes of a warp (one block each) read consecutive addresses for every k.
  struct Bsr {
    int B = 0;             // block size (0 = not built)
    int64_t n_brows = 0, nnzb = 0;
    DBuf<int32_t> bptr, bcol;
    DBuf<double> bval;
    DBuf<float> bval32;    // FP32 copy of bval, read ONLY inside the Chebyshev preconditioner (PE_CHEB_FP32=1): a fixed SPD
                           // polynomial in D^-1 A~, so CG still converges to the FP64 solution at 40 instead of 76 B per block
  } bsr;
  SellMat sell[5];  // slots: 0 = displacement matrix, 1 = its FP32 copy, 2 = pressure Jacobian, 3 = projection (mass) matrix, 4 = FP32 Jacobian
  const SellMat* find_sell(const double* val, bool f32) const {
    for (const SellMat& m : sell)
      if (m.B && m.src == val && m.f32 == f32) return &m;
    return nullptr;
  }
};

// device-resident scalar state of one CG solve
struct CgState {
  double gh, tol, res, res0;
  int it, done, max_it, pad;
};

static constexpr int PE_MAX_RED_BLOCKS = 4096;
static constexpr int PE_RED_SLOTS = 4;

struct Reducer {
  DBuf<double> partials;   // PE_RED_SLOTS * PE_MAX_RED_BLOCKS
  DBuf<unsigned> counter;  // 1
  DBuf<double> out;        // PE_RED_SLOTS (+ scratch)
  // deterministic sums of the dynamically distributed sliced kernels (kernels_sell.cuh): per-slice and per-group slots
  DBuf<unsigned> claim;    // 6: three claim counters the persistent kernel rotates by pass number + three counts of finished boundary slices
  DBuf<double> spart, gpart;
  DBuf<unsigned> gcnt;
  int cap = 0, gcap = 0;   // slices / groups the slot arrays hold
};
static constexpr int PE_SELL_NV = 4;  // values a pass can reduce at once

// Peer-memory communication over NVLink (kernels_comm.cu): every rank exports one region
// [control block | six CG work vectors] with cudaIpc; halo values are stored by the SENDER straight into the
// ghost segment of the receiver's vector and reductions are exchanged through per-sender mailboxes, both
// published with a system-scope fence + an epoch flag the receiver polls.  NCCL stays for setup and for
// the few exchanges outside the CG loop.
static constexpr int PE_P2P_MAX_RANKS = 16;
static constexpr int PE_WORK_VECTORS = 10;
static constexpr int PE_PCG_TIMING_WORDS = 16;
struct P2PControl {                                   // lives at the start of every rank's region
  int halo_flag[2][PE_P2P_MAX_RANKS];                 // [field][sender rank] = epoch of the last halo it stored here
  int red_flag[PE_P2P_MAX_RANKS];                     // [sender rank] = epoch of its last mailbox post
  double red_val[2][PE_P2P_MAX_RANKS][4];             // [epoch parity][sender rank][slot]
};
struct P2PField {
  DBuf<int32_t> send_dest;   // per send entry: index in the receiver's vector (ghost segment)
  DBuf<int32_t> send_nb;     // per send entry: neighbour slot
  DBuf<int32_t> neigh_rank;  // neighbour ranks (device copy)
  // the same plan indexed by the SOURCE row (boundary rows only, row - n_interior): the persistent CG kernel lets the
  // thread that produces an entry store it to the neighbours itself
  DBuf<int32_t> push_ptr, push_dest, push_nb;
  DBuf<unsigned long long> push_addr;  // per entry: address of the ghost copy inside the owner-of-the-copy's FIRST work vector (add the vector's offset)
  DBuf<int32_t> push_src;              // per entry: (row - n_interior) mod (32 * ncomp) = lane * ncomp + component inside its slice
  bool push_ok = false;      // every sent row is a boundary row (>= n_interior); true for symmetric patterns
  unsigned epoch = 0;        // halo exchanges posted so far
};
struct P2P {
  bool on = false;
  char* region = nullptr;            // my region (cudaMalloc)
  size_t region_bytes = 0, ctrl_bytes = 0;
  std::vector<char*> peer;           // region of every rank as mapped here (peer[rank] == region)
  DBuf<char*> d_peer;                // device copy
  DBuf<unsigned> ticket;             // last-block counter of k_halo_send
  P2PField f[2];
  unsigned red_epoch = 0;
  unsigned stage_epoch = 0;          // staged halo exchanges so far (parity picks the staging buffer)
};

struct pe_ctx {
  int device = 0, rank = 0, nranks = 1;
  cudaStream_t stream = nullptr;
  ncclComm_t comm = nullptr;
  ncclComm_t comm_nccl() const { return comm; }
  std::string err;
  pe_params prm{};
  bool have_params = false, have_mesh = false, setup_done = false;
  int sm_count = 148;

  // mesh
  int dim = 0, vpc = 0;
  int64_t n_vertices = 0, n_cells = 0, n_bfaces = 0;
  std::vector<int32_t> h_cell_vertices;
  DBuf<double> xyz;
  DBuf<int32_t> cell_vertices, bface_cell, bface_id;
  DBuf<int8_t> bface_local;
  std::vector<int32_t> h_bface_cell, h_bface_id;
  std::vector<int8_t> h_bface_local;
  // colouring
  int n_colors = 0;
  std::vector<int64_t> color_ptr;
  DBuf<int32_t> color_cells;
  // neumann
  std::vector<int32_t> nm_label, nm_comp;
  std::vector<double> nm_value;

  Field fp, fu;  // pressure, displacement

  // reference tables
  QuadTab q2, qu;                       // QGauss(2); QGauss(degree_u+1)
  ShapeTab p_q2, p_qu, us_q2, us_qu;    // pressure FE / displacement scalar FE at those points
  DBuf<double> usup;                    // unit support points of the displacement scalar FE (ns*dim)

  // matrices (values on the field patterns) and inverse diagonals
  DBuf<double> M, K, J, A;
  DBuf<double> Mc, Kc;  // ConstraintMatrix::condense of M and K (hanging-node meshes only): J = Mc/(M_b dt) + (k/mu) Kc, projection matrix = Mc
  DBuf<double> invdiag_M, invdiag_J, invdiag_A;
  double jac_dt = -1;
  bool matrix_u_built = false, proj_matrix_ready = false;
  double eig_M = 0, eig_J = 0, eig_A = 0;

  // vectors (n_local of their field)
  DBuf<double> p, p_old, dp, resid, ev, ev0, frhs, t1;
  DBuf<double> u, b, b_const;
  std::vector<DBuf<double>> strains, proj_rhs, stresses;
  int n_stress = 0;
  // CG work vectors sized for the larger field; they live in `comm.region` (IPC-exported when nranks > 1)
  struct WPtr { double* p = nullptr; };
  WPtr w_g, w_h, w_d, w_z, w_d2, w_r, w_s, w_c1, w_x0, w_x1;
  P2P p2p;

  Reducer red;
  // persistent CG kernel scratch: tickets[3], flags {barrier epoch, abort}, timing {ns, count}
  DBuf<unsigned> pcg_tickets;
  DBuf<int> pcg_flags;
  DBuf<unsigned long long> pcg_timing;
  DBuf<unsigned long long> pcg_trace;  // PE_PCG_TRACE=<file prefix>: per-warp timestamps of the last passes of a displacement solve (diagnostic)
  int pcg_grid[2] = {0, 0};  // cooperative grid size per field (0 = not queried yet)
  double pcg_phase_ns[2][8] = {{0}};  // CTA-0 sub-phase times of the persistent kernel (printed with PE_PCG_TIMING=1)
  long long pcg_phase_its[2] = {0, 0};
  DBuf<CgState> cg_state;
  CgState* h_state = nullptr;  // pinned
  double* h_scalars = nullptr; // pinned, PE_RED_SLOTS
  // Communication error word in pinned, device-mapped host memory: every bounded wait of the peer-memory protocol
  // that times out stores a code here (kernels_comm.cu, the fused waits of the SpMV kernels, k_pcg).  The host reads
  // it after its next stream synchronisation (pe_sync_checked) and fails the call with PE_ERR_NCCL, so a late peer can
  // never leave rank-local partial sums or stale ghost values behind silently.
  int* h_comm_err = nullptr;
  cudaEvent_t ev_poll[2] = {nullptr, nullptr};

  pe_stats st{};

  // optional per-launch timing of the matrix passes (roofline evidence): event pairs on `stream`
  bool profiling = false, prof_hold = false;
  std::vector<cudaEvent_t> prof_ev;        // 2 * PROF_PAIRS events
  std::vector<int> prof_field;             // field of each recorded pair
  int prof_used = 0;
};
static constexpr int PE_PROF_PAIRS = 2048;
void pe_prof_begin(pe_ctx* c, int field);
void pe_prof_end(pe_ctx* c);
void pe_prof_flush(pe_ctx* c);

// ---- kernels_pattern.cu
void pe_build_pattern(pe_ctx* c, Field& F);
bool pe_build_bsr(pe_ctx* c, Field& F, const double* csr_val);  // false when the matrix has no block structure
int64_t pe_exclusive_scan_i32(pe_ctx* c, int32_t* data, int64_t n);  // in place, n+1 entries written (last = total)

// ---- kernels_sell.cu
// (Re)builds sell slot `slot` of F from the matrix `val` (CSR for scalar fields, the block-CSR copy otherwise); returns false
// and clears the slot when padding would cost more than PE_SELL_MAX_PAD (default 1.5) times the real blocks.
bool pe_build_sell(pe_ctx* c, Field& F, int slot, const double* val, bool f32);
bool pe_all_ranks_agree(pe_ctx* c, bool mine);  // collective AND of a rank-local flag (kernels_comm.cu)

// ---- kernels_constraints.cu (hanging-node lines; active only when a field has lines with entries)
void pe_hanging_upload(pe_ctx* c, Field& F);
void pe_build_pattern_lists(pe_ctx* c, Field& F);  // pattern of cell lists extended by the masters of hanging dofs
void pe_condense_matrix(pe_ctx* c, Field& F, const double* src, double* dst, bool keep_diag, double hang_diag);
void pe_condense_vector(pe_ctx* c, Field& F, double* v);
void pe_distribute_hanging(pe_ctx* c, Field& F, double* v);
void pe_scatter_hanging_inhomogeneity(pe_ctx* c, Field& F, double* v);  // v = 0 except g_h on hanging dofs
double pe_avg_abs_diag(pe_ctx* c, Field& F, const double* val);

// ---- kernels_assembly.cu
void pe_build_tables(pe_ctx* c);
void pe_color_cells(pe_ctx* c);
void pe_assemble_pressure_matrices(pe_ctx* c);  // M, K, well rhs
void pe_assemble_elasticity(pe_ctx* c);         // A, b_const (Dirichlet + Neumann)
void pe_assemble_u_rhs(pe_ctx* c);              // b = b_const + alpha * int p div(phi)
void pe_assemble_projection_rhs(pe_ctx* c, int n_comp, const int32_t* comps, const int32_t* entries);

// ---- kernels_solver.cu
struct CgResult { int its; double res; int status; };
// in_solve: honour CgState::done.  fused_epoch != null: peer-memory sends skip the separate wait kernel and return the
// epoch the consumer has to wait for itself (0 when nothing has to be waited for).
void pe_halo_exchange(pe_ctx* c, Field& F, double* v, bool in_solve = false, int* fused_epoch = nullptr);
void pe_extract_invdiag(pe_ctx* c, Field& F, const double* val, double* invdiag);
double pe_estimate_eig_max(pe_ctx* c, Field& F, const double* val, const double* invdiag);
CgResult pe_cg_solve(pe_ctx* c, Field& F, const double* val, const double* invdiag, double eig_max, double* x, const double* b,
                     double tol, bool tol_relative_to_b, int64_t* spmv_counter);
void pe_spmv_plain(pe_ctx* c, Field& F, const double* val, const double* x, double* y);
double pe_pressure_residual(pe_ctx* c, double dt);  // returns l2 norm (global)
void pe_vec_axpy(pe_ctx* c, int64_t n, double a, const double* x, double* y);  // y += a x
void pe_vec_copy(pe_ctx* c, int64_t n, const double* x, double* y);
void pe_vec_set(pe_ctx* c, int64_t n, double a, double* y);
void pe_vec_axpby_vals(pe_ctx* c, int64_t n, double a, const double* x, double b, const double* y, double* z);  // z = a x + b y
void pe_distribute(pe_ctx* c, Field& F, double* v);  // constrained dofs <- inhomogeneity
double pe_linfty(pe_ctx* c, Field& F, const double* v);
void pe_stress_kernel(pe_ctx* c);
void pe_allreduce_sum(pe_ctx* c, double* dev, int count, bool in_solve = false);
void pe_sync_checked(pe_ctx* c);  // cudaStreamSynchronize + the communication error word (throws PE_ERR_NCCL)
double pe_vec_dot(pe_ctx* c, Field& F, const double* a, const double* b);
void pe_build_bsr_fp32(pe_ctx* c, Field& F);  // no-op unless PE_CHEB_FP32=1 and the preconditioner is Chebyshev  // global dot product over owned entries
// ---- kernels_comm.cu
void pe_comm_setup(pe_ctx* c, size_t n_work);  // allocates the region (+ IPC exchange when nranks > 1)
void pe_comm_release(pe_ctx* c);
void pe_pack_launch(pe_ctx* c, int64_t n, const int32_t* idx, const double* v, double* buf);

#ifdef __CUDACC__
// ---- device-side primitives of the peer-memory protocol (used by kernels_comm.cu and kernels_solver.cu)
__device__ __forceinline__ int pe_ld_flag(const int* p) {
  int v;
  asm volatile("ld.volatile.global.s32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ void pe_st_flag(int* p, int v) { asm volatile("st.volatile.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ double pe_ld_mail(const double* p) {
  double v;
  asm volatile("ld.volatile.global.f64 %0, [%1];" : "=d"(v) : "l"(p));
  return v;
}
// spin until *flag has reached `epoch` (wrap-safe compare); false after ~10 s
__device__ __forceinline__ bool pe_wait_flag(const int* flag, int epoch) {
  const long long t0 = clock64();
  while ((int)(pe_ld_flag(flag) - epoch) < 0) {
    if (clock64() - t0 > 20000000000LL) return false;
  }
  return true;
}
#endif

static inline int pe_div_up(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }
