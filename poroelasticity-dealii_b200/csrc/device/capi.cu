// capi.cu — extern "C" entry points of libporoel.so (include/poroel.h).
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <string>

#include "pe_internal.cuh"

static thread_local std::string g_create_err;

#define PE_ENTER(ctx)                                \
  if (!(ctx)) return PE_ERR_BAD_INPUT;               \
  try {                                              \
    PE_CUDA(cudaSetDevice((ctx)->device));
#define PE_LEAVE(ctx)                                \
    return PE_OK;                                    \
  } catch (const PeError& e) {                       \
    (ctx)->err = e.what();                           \
    return e.code;                                   \
  } catch (const std::exception& e) {                \
    (ctx)->err = e.what();                           \
    return PE_ERR_BAD_INPUT;                         \
  }

static void require(bool cond, int code, const char* msg) {
  if (!cond) throw PeError(code, msg);
}

static Field& field_of(pe_ctx* c, int f) {
  require(f == PE_FIELD_PRESSURE || f == PE_FIELD_DISPLACEMENT, PE_ERR_BAD_INPUT, "bad field id");
  return f == PE_FIELD_PRESSURE ? c->fp : c->fu;
}

static int sym_entry(int dim, int t) {  // TensorIndexer.h:25-30
  static const int m2[4] = {0, 1, 1, 2}, m3[9] = {0, 1, 2, 1, 3, 4, 2, 4, 5};
  return dim == 2 ? m2[t] : m3[t];
}

struct VecRef { double* p; Field* F; };
static VecRef vec_by_id(pe_ctx* c, int which) {
  switch (which) {
    case PE_VEC_P: return {c->p.p, &c->fp};
    case PE_VEC_P_OLD: return {c->p_old.p, &c->fp};
    case PE_VEC_P_UPDATE: return {c->dp.p, &c->fp};
    case PE_VEC_P_RESIDUAL: return {c->resid.p, &c->fp};
    case PE_VEC_VOL_STRAIN: return {c->ev.p, &c->fp};
    case PE_VEC_VOL_STRAIN0: return {c->ev0.p, &c->fp};
    case PE_VEC_WELL_RHS: return {c->frhs.p, &c->fp};
    case PE_VEC_U: return {c->u.p, &c->fu};
    case PE_VEC_U_RHS: return {c->b.p, &c->fu};
    default: break;
  }
  if (which >= PE_VEC_STRAIN0 && which < PE_VEC_STRAIN0 + c->n_stress) return {c->strains[which - PE_VEC_STRAIN0].p, &c->fp};
  if (which >= PE_VEC_PROJ_RHS0 && which < PE_VEC_PROJ_RHS0 + c->n_stress) return {c->proj_rhs[which - PE_VEC_PROJ_RHS0].p, &c->fp};
  if (which >= PE_VEC_STRESS0 && which < PE_VEC_STRESS0 + c->n_stress) return {c->stresses[which - PE_VEC_STRESS0].p, &c->fp};
  throw PeError(PE_ERR_BAD_INPUT, "bad vector id");
}

struct MatRef { const double* val; Field* F; };
static MatRef mat_by_id(pe_ctx* c, int m) {
  switch (m) {
    case PE_MAT_MASS: return {c->M.p, &c->fp};
    case PE_MAT_LAPLACE: return {c->K.p, &c->fp};
    case PE_MAT_JACOBIAN: return {c->J.p, &c->fp};
    case PE_MAT_ELASTICITY: return {c->A.p, &c->fu};
    case PE_MAT_PROJECTION: return {c->fp.hang.n ? c->Mc.p : c->M.p, &c->fp};
  }
  throw PeError(PE_ERR_BAD_INPUT, "bad matrix id");
}

extern "C" {

int pe_version(void) { return 100; }

int pe_nccl_unique_id(void* out128, size_t* n_bytes) {
  if (!out128 || !n_bytes || *n_bytes < sizeof(ncclUniqueId)) return PE_ERR_BAD_INPUT;
  ncclUniqueId id;
  if (ncclGetUniqueId(&id) != ncclSuccess) return PE_ERR_NCCL;
  std::memcpy(out128, &id, sizeof id);
  *n_bytes = sizeof id;
  return PE_OK;
}

int pe_create(pe_ctx** out, int device, int rank, int nranks, const void* nccl_id, size_t id_bytes) {
  if (!out) return PE_ERR_BAD_INPUT;
  *out = nullptr;
  pe_ctx* c = nullptr;
  try {
    int n_dev = 0;
    cudaError_t e = cudaGetDeviceCount(&n_dev);
    if (e != cudaSuccess || n_dev == 0)
      throw PeError(PE_ERR_CUDA, std::string("no usable CUDA device (this library has no CPU fallback): ") + cudaGetErrorString(e));
    require(device >= 0 && device < n_dev, PE_ERR_BAD_INPUT, "device index out of range");
    require(nranks >= 1 && rank >= 0 && rank < nranks, PE_ERR_BAD_INPUT, "bad rank / nranks");
    PE_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    PE_CUDA(cudaGetDeviceProperties(&prop, device));
    require(prop.major >= 10, PE_ERR_CUDA, "this build targets sm_100a (B200) only");
    c = new pe_ctx();
    c->device = device;
    c->rank = rank;
    c->nranks = nranks;
    c->sm_count = prop.multiProcessorCount;
    PE_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    if (nranks > 1) {
      require(nccl_id && id_bytes == sizeof(ncclUniqueId), PE_ERR_BAD_INPUT, "nranks > 1 needs the 128-byte NCCL unique id of rank 0");
      ncclUniqueId id;
      std::memcpy(&id, nccl_id, sizeof id);
      const auto tn = std::chrono::steady_clock::now();
      PE_NCCL(ncclCommInitRank(&c->comm, nranks, id, rank));
      if (std::getenv("PE_SETUP_TIMING"))
        std::fprintf(stderr, "[pe rank %d] ncclCommInitRank %.1f ms\n", rank, std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - tn).count());
    }
    c->red.partials.alloc_zero((size_t)PE_RED_SLOTS * PE_MAX_RED_BLOCKS, c->stream);
    c->red.counter.alloc_zero(1, c->stream);
    c->red.out.alloc_zero(PE_RED_SLOTS + 8, c->stream);
    c->cg_state.alloc_zero(1, c->stream);
    c->pcg_tickets.alloc_zero(4, c->stream);
    c->pcg_flags.alloc_zero(2, c->stream);
    c->pcg_timing.alloc_zero(PE_PCG_TIMING_WORDS, c->stream);
    PE_CUDA(cudaMallocHost((void**)&c->h_state, 2 * sizeof(CgState)));
    PE_CUDA(cudaMallocHost((void**)&c->h_scalars, (PE_RED_SLOTS + 8) * sizeof(double)));
    PE_CUDA(cudaHostAlloc((void**)&c->h_comm_err, sizeof(int), cudaHostAllocMapped));
    *c->h_comm_err = 0;
    PE_CUDA(cudaEventCreateWithFlags(&c->ev_poll[0], cudaEventDisableTiming));
    PE_CUDA(cudaEventCreateWithFlags(&c->ev_poll[1], cudaEventDisableTiming));
    PE_CUDA(cudaStreamSynchronize(c->stream));
    *out = c;
    return PE_OK;
  } catch (const PeError& e) {
    g_create_err = e.what();
    delete c;
    return e.code;
  } catch (const std::exception& e) {
    g_create_err = e.what();
    delete c;
    return PE_ERR_BAD_INPUT;
  }
}

void pe_destroy(pe_ctx* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  if (std::getenv("PE_PCG_TIMING")) {
    static const char* names[8] = {"halo send", "interior rows", "halo wait", "boundary rows", "reduce+fetch d.h", "update", "reduce+fetch res,gz", "direction+barrier"};
    for (int f = 0; f < 2; ++f) {
      if (!c->pcg_phase_its[f]) continue;
      std::fprintf(stderr, "[pe rank %d] k_pcg %s: %lld iterations, us per iteration (CTA 0):", c->rank, f ? "displacement" : "pressure/projection", c->pcg_phase_its[f]);
      for (int k = 0; k < 8; ++k) std::fprintf(stderr, " %s %.1f;", names[k], c->pcg_phase_ns[f][k] * 1e-3 / (double)c->pcg_phase_its[f]);
      std::fprintf(stderr, "\n");
    }
  }
  if (c->stream) cudaStreamSynchronize(c->stream);
  pe_comm_release(c);
  if (c->comm) ncclCommDestroy(c->comm);
  if (c->h_state) cudaFreeHost(c->h_state);
  if (c->h_scalars) cudaFreeHost(c->h_scalars);
  if (c->h_comm_err) cudaFreeHost(c->h_comm_err);
  for (auto& e : c->ev_poll) if (e) cudaEventDestroy(e);
  for (auto& e : c->prof_ev) if (e) cudaEventDestroy(e);
  cudaStream_t s = c->stream;
  delete c;
  if (s) cudaStreamDestroy(s);
}

const char* pe_last_error(const pe_ctx* c) { return c ? c->err.c_str() : g_create_err.c_str(); }

int pe_set_params(pe_ctx* c, const pe_params* p) {
  PE_ENTER(c)
  require(p != nullptr, PE_ERR_BAD_INPUT, "null params");
  require(p->dim == 2 || p->dim == 3, PE_ERR_BAD_INPUT, "dim must be 2 or 3");
  require(p->degree_p == 1, PE_ERR_UNSUPPORTED, "pressure degree must be 1 (PS:20)");
  require(p->degree_u == 1 || p->degree_u == 2, PE_ERR_UNSUPPORTED, "displacement degree must be 1 or 2");
  require(p->preconditioner == PE_PRECOND_JACOBI || p->preconditioner == PE_PRECOND_CHEBYSHEV, PE_ERR_BAD_INPUT, "unknown preconditioner");
  require(p->cg_max_iterations >= 1, PE_ERR_BAD_INPUT, "cg_max_iterations must be >= 1");
  require(p->preconditioner != PE_PRECOND_CHEBYSHEV || (p->chebyshev_degree >= 1 && p->chebyshev_eig_ratio > 1.0), PE_ERR_BAD_INPUT,
          "chebyshev degree must be >= 1 and the eigenvalue ratio > 1");
  const bool solver_only_change = c->have_params && c->prm.dim == p->dim && c->prm.degree_u == p->degree_u;
  require(!c->setup_done || solver_only_change, PE_ERR_STATE, "dim / degree cannot change after pe_setup");
  c->prm = *p;
  c->have_params = true;
  PE_LEAVE(c)
}

int pe_upload_mesh(pe_ctx* c, int dim, int64_t nv, const double* xyz, int64_t nc, const int32_t* cv, int64_t nbf, const int32_t* bc,
                   const int8_t* bl, const int32_t* bid) {
  PE_ENTER(c)
  require(c->have_params, PE_ERR_STATE, "pe_set_params first");
  require(dim == c->prm.dim, PE_ERR_BAD_INPUT, "mesh dim differs from params");
  require(nv > 0 && nc > 0 && xyz && cv, PE_ERR_BAD_INPUT, "empty mesh");
  require(nc < ((int64_t)1 << 31) && nv < ((int64_t)1 << 31), PE_ERR_UNSUPPORTED, "mesh too large for 32-bit indices");
  c->dim = dim;
  c->vpc = 1 << dim;
  c->n_vertices = nv;
  c->n_cells = nc;
  c->n_bfaces = nbf;
  for (int64_t i = 0; i < nc * c->vpc; ++i) require(cv[i] >= 0 && cv[i] < nv, PE_ERR_BAD_INPUT, "cell vertex index out of range");
  c->h_cell_vertices.assign(cv, cv + nc * c->vpc);
  c->xyz.upload(xyz, (size_t)nv * dim, c->stream);
  c->cell_vertices.upload(cv, (size_t)nc * c->vpc, c->stream);
  c->h_bface_cell.assign(bc, bc + nbf);
  c->h_bface_local.assign(bl, bl + nbf);
  c->h_bface_id.assign(bid, bid + nbf);
  c->bface_cell.upload(c->h_bface_cell, c->stream);
  c->bface_local.upload(c->h_bface_local, c->stream);
  c->bface_id.upload(c->h_bface_id, c->stream);
  PE_CUDA(cudaStreamSynchronize(c->stream));
  c->have_mesh = true;
  c->setup_done = false;
  PE_LEAVE(c)
}

int pe_upload_dofs(pe_ctx* c, int field, int64_t n_local, const int32_t* cell_dofs) {
  PE_ENTER(c)
  require(c->have_mesh, PE_ERR_STATE, "pe_upload_mesh first");
  Field& F = field_of(c, field);
  F.degree = field == PE_FIELD_PRESSURE ? c->prm.degree_p : c->prm.degree_u;
  F.ncomp = field == PE_FIELD_PRESSURE ? 1 : c->dim;
  int ns = 1;
  for (int a = 0; a < c->dim; ++a) ns *= F.degree + 1;
  F.ns = ns;
  F.nloc = ns * F.ncomp;
  require(n_local > 0 && n_local < ((int64_t)1 << 31) && cell_dofs, PE_ERR_BAD_INPUT, "bad dof count");
  const int64_t ne = c->n_cells * F.nloc;
  for (int64_t i = 0; i < ne; ++i) require(cell_dofs[i] >= 0 && cell_dofs[i] < n_local, PE_ERR_BAD_INPUT, "cell dof index out of range");
  F.n_local = n_local;
  F.n_owned = n_local;
  F.h_cell_dofs.assign(cell_dofs, cell_dofs + ne);
  F.cell_dofs.upload(cell_dofs, (size_t)ne, c->stream);
  PE_CUDA(cudaStreamSynchronize(c->stream));
  F.have_dofs = true;
  F.have_partition = false;
  F.n_lines = 0;
  F.h_line_dof.clear();
  F.h_line_g.clear();
  F.hang.clear();
  F.halo.reset();
  c->setup_done = false;
  PE_LEAVE(c)
}

int pe_upload_constraints(pe_ctx* c, int field, int64_t n_lines, const int32_t* line_dof, const int64_t* entry_ptr, const int32_t* entry_dof,
                          const double* entry_w, const double* inhomogeneity) {
  PE_ENTER(c)
  Field& F = field_of(c, field);
  require(F.have_dofs, PE_ERR_STATE, "pe_upload_dofs first");
  require(n_lines >= 0, PE_ERR_BAD_INPUT, "negative line count");
  require(n_lines == 0 || line_dof != nullptr, PE_ERR_BAD_INPUT, "null line array");
  const bool have_entries = entry_ptr && n_lines > 0 && entry_ptr[n_lines] > 0;
  if (have_entries) require(entry_dof && entry_w, PE_ERR_BAD_INPUT, "constraint entries without dof / weight arrays");
  std::vector<uint8_t> is_line((size_t)F.n_local, 0);
  for (int64_t i = 0; i < n_lines; ++i) {
    require(line_dof[i] >= 0 && line_dof[i] < F.n_local, PE_ERR_BAD_INPUT, "constrained dof out of range");
    require(!is_line[line_dof[i]], PE_ERR_BAD_INPUT, "dof constrained twice");
    is_line[line_dof[i]] = 1;
  }
  // lines without entries are Dirichlet values (handled inside the cell kernels); lines with entries are hanging nodes
  // (handled on the assembled objects, kernels_constraints.cu)
  std::vector<int32_t> d_dof;
  std::vector<double> d_g;
  Field::Hanging H;
  H.h_ptr.push_back(0);
  for (int64_t i = 0; i < n_lines; ++i) {
    const int64_t e0 = have_entries ? entry_ptr[i] : 0, e1 = have_entries ? entry_ptr[i + 1] : 0;
    require(e0 <= e1, PE_ERR_BAD_INPUT, "entry_ptr must be non-decreasing");
    const double g = inhomogeneity ? inhomogeneity[i] : 0.0;
    if (e0 == e1) {
      d_dof.push_back(line_dof[i]);
      d_g.push_back(g);
      continue;
    }
    H.h_dof.push_back(line_dof[i]);
    H.h_g.push_back(g);
    if (g != 0.0) H.any_g = true;
    for (int64_t e = e0; e < e1; ++e) {
      require(entry_dof[e] >= 0 && entry_dof[e] < F.n_local, PE_ERR_BAD_INPUT, "constraint entry out of range");
      require(!is_line[entry_dof[e]], PE_ERR_BAD_INPUT, "constraint table is not closed: an entry refers to a constrained dof");
      H.h_edof.push_back(entry_dof[e]);
      H.h_w.push_back(entry_w[e]);
    }
    H.h_ptr.push_back((int32_t)H.h_edof.size());
  }
  H.n = (int64_t)H.h_dof.size();
  H.n_entries = (int64_t)H.h_edof.size();
  if (field == PE_FIELD_PRESSURE)
    require(d_dof.empty() && !H.any_g, PE_ERR_UNSUPPORTED, "pressure constraints are homogeneous hanging-node lines only (PS:71-78)");
  F.n_lines = (int64_t)d_dof.size();
  F.h_line_dof.swap(d_dof);
  F.h_line_g.swap(d_g);
  F.hang.clear();
  F.hang.n = H.n;
  F.hang.n_entries = H.n_entries;
  F.hang.any_g = H.any_g;
  F.hang.h_dof.swap(H.h_dof);
  F.hang.h_ptr.swap(H.h_ptr);
  F.hang.h_edof.swap(H.h_edof);
  F.hang.h_w.swap(H.h_w);
  F.hang.h_g.swap(H.h_g);
  c->setup_done = false;
  PE_LEAVE(c)
}

int pe_upload_neumann(pe_ctx* c, int n, const int32_t* label, const int32_t* comp, const double* value) {
  PE_ENTER(c)
  require(n >= 0, PE_ERR_BAD_INPUT, "negative count");
  for (int i = 0; i < n; ++i) require(comp[i] >= 0 && comp[i] < c->prm.dim, PE_ERR_BAD_INPUT, "Neumann component >= dim (BC:55-57)");
  c->nm_label.assign(label, label + n);
  c->nm_comp.assign(comp, comp + n);
  c->nm_value.assign(value, value + n);
  c->setup_done = false;
  PE_LEAVE(c)
}

int pe_upload_partition(pe_ctx* c, int field, int64_t n_owned, int n_neigh, const int32_t* neigh_rank, const int64_t* send_ptr,
                        const int32_t* send_idx, const int64_t* recv_ptr) {
  PE_ENTER(c)
  Field& F = field_of(c, field);
  require(F.have_dofs, PE_ERR_STATE, "pe_upload_dofs first");
  require(n_owned >= 0 && n_owned <= F.n_local, PE_ERR_BAD_INPUT, "n_owned out of range");
  require(n_neigh >= 0, PE_ERR_BAD_INPUT, "negative neighbour count");
  F.n_owned = n_owned;
  Halo& H = F.halo;
  H.n_neigh = n_neigh;
  H.rank.assign(neigh_rank, neigh_rank + n_neigh);
  H.send_ptr.assign(send_ptr, send_ptr + n_neigh + 1);
  H.recv_ptr.assign(recv_ptr, recv_ptr + n_neigh + 1);
  for (int r : H.rank) require(r >= 0 && r < c->nranks && r != c->rank, PE_ERR_BAD_INPUT, "bad neighbour rank");
  require(n_neigh == 0 || H.recv_ptr.back() == F.n_local - n_owned, PE_ERR_BAD_INPUT, "receive plan does not cover the ghost dofs");
  require(n_neigh > 0 || F.n_local == n_owned, PE_ERR_BAD_INPUT, "ghost dofs without neighbours");
  const int64_t ns = H.n_send();
  for (int64_t i = 0; i < ns; ++i) require(send_idx[i] >= 0 && send_idx[i] < n_owned, PE_ERR_BAD_INPUT, "send index is not an owned dof");
  H.send_idx.upload(send_idx, (size_t)ns, c->stream);
  H.send_buf.alloc((size_t)ns);
  PE_CUDA(cudaStreamSynchronize(c->stream));
  F.have_partition = true;
  c->setup_done = false;
  PE_LEAVE(c)
}

int pe_setup(pe_ctx* c) {
  PE_ENTER(c)
  require(c->have_mesh && c->fp.have_dofs && c->fu.have_dofs, PE_ERR_STATE, "mesh and both dof maps must be uploaded");
  require(c->nranks == 1 || (c->fp.have_partition && c->fu.have_partition), PE_ERR_STATE, "nranks > 1 needs pe_upload_partition for both fields");
  require(c->fp.n_lines == 0, PE_ERR_UNSUPPORTED, "pressure constraints are hanging-node lines only (PS:71-78)");
  require((c->fp.hang.n == 0 && c->fu.hang.n == 0) || c->nranks == 1, PE_ERR_UNSUPPORTED,
          "hanging-node (adaptive) meshes run on one rank; partitioned runs need uniform meshes");
  auto t0 = std::chrono::steady_clock::now();
  cudaStream_t s = c->stream;
  // PE_SETUP_TIMING=1: wall time of every phase of pe_setup on stderr (init_s of partitioned runs is mostly this call)
  static const bool timing = std::getenv("PE_SETUP_TIMING") != nullptr;
  auto t_last = t0;
  std::string t_report;
  auto tick = [&](const char* what) {
    if (!timing) return;
    cudaStreamSynchronize(s);
    const auto now = std::chrono::steady_clock::now();
    t_report += std::string(" ") + what + " " + std::to_string(std::chrono::duration<double, std::milli>(now - t_last).count()).substr(0, 7) + " ms;";
    t_last = now;
  };
  for (Field* F : {&c->fp, &c->fu}) {
    for (SellMat& S : F->sell) {  // sliced copies belong to the previous mesh
      S.B = 0;
      S.src = nullptr;
      S.panels.release();
      S.slice_ptr.release();
    }
    std::vector<int32_t> cl((size_t)F->n_local, -1);
    for (int64_t i = 0; i < F->n_lines; ++i) {
      require(cl[F->h_line_dof[i]] < 0, PE_ERR_BAD_INPUT, "dof constrained twice");
      cl[F->h_line_dof[i]] = (int32_t)i;
    }
    F->cline.upload(cl, s);
    F->line_dof.upload(F->h_line_dof, s);
    F->line_g.upload(F->h_line_g, s);
    PE_CUDA(cudaStreamSynchronize(s));
    F->bsr.B = 0;
    if (F->hang.n) {
      pe_hanging_upload(c, *F);
      pe_build_pattern_lists(c, *F);  // make_sparsity_pattern(dh, dsp, constraints, true), PS:80-88 / DS:140-146
    } else
      pe_build_pattern(c, *F);
  }
  tick("patterns");
  pe_build_tables(c);
  pe_color_cells(c);
  tick("tables+colouring");
  const int64_t npl = c->fp.n_local, nul = c->fu.n_local;
  for (DBuf<double>* v : {&c->p, &c->p_old, &c->dp, &c->resid, &c->ev, &c->ev0, &c->frhs, &c->t1}) v->alloc_zero((size_t)npl, s);
  for (DBuf<double>* v : {&c->u, &c->b, &c->b_const}) v->alloc_zero((size_t)nul, s);
  c->n_stress = (c->dim * c->dim + c->dim) / 2;
  c->strains = std::vector<DBuf<double>>(c->n_stress);
  c->proj_rhs = std::vector<DBuf<double>>(c->n_stress);
  c->stresses = std::vector<DBuf<double>>(c->n_stress);
  for (int e = 0; e < c->n_stress; ++e) {
    c->strains[e].alloc_zero((size_t)npl, s);
    c->proj_rhs[e].alloc_zero((size_t)npl, s);
    c->stresses[e].alloc_zero((size_t)npl, s);
  }
  tick("vectors");
  pe_comm_setup(c, (size_t)std::max(npl, nul));
  tick("comm setup (IPC region, NCCL handshakes)");
  c->M.alloc((size_t)c->fp.nnz);
  c->K.alloc((size_t)c->fp.nnz);
  c->J.alloc_zero((size_t)c->fp.nnz, s);
  c->A.alloc_zero((size_t)c->fu.nnz, s);
  c->invdiag_M.alloc((size_t)c->fp.n_owned);
  c->invdiag_J.alloc((size_t)c->fp.n_owned);
  c->invdiag_A.alloc((size_t)c->fu.n_owned);
  pe_assemble_pressure_matrices(c);  // PS:96-101 (+ cached well source, PS:142-147)
  tick("M, K, well rhs");
  if (c->fp.hang.n) {
    // constraints.condense() of PS:168 and SP:105 is linear in the matrix, so it is applied once to M and K;
    // constrained diagonals get the average |diagonal| of the uncondensed matrix (ConstraintMatrix::condense)
    c->Mc.alloc((size_t)c->fp.nnz);
    c->Kc.alloc((size_t)c->fp.nnz);
    const double avg_m = pe_avg_abs_diag(c, c->fp, c->M.p), avg_k = pe_avg_abs_diag(c, c->fp, c->K.p);
    pe_condense_matrix(c, c->fp, c->M.p, c->Mc.p, false, avg_m);
    pe_condense_matrix(c, c->fp, c->K.p, c->Kc.p, false, avg_k);
  } else {
    c->Mc.release();
    c->Kc.release();
  }
  const double* m_proj = c->fp.hang.n ? c->Mc.p : c->M.p;
  pe_extract_invdiag(c, c->fp, m_proj, c->invdiag_M.p);
  // the projection solves (SP:201-232) stream this copy; format decisions are collective (pe_all_ranks_agree)
  if (!pe_all_ranks_agree(c, pe_build_sell(c, c->fp, 3, m_proj, false))) c->fp.sell[3].B = 0;
  c->eig_M = 1.1 * pe_estimate_eig_max(c, c->fp, m_proj, c->invdiag_M.p);
  tick("sliced copy + eigenvalue estimate of M");
  if (timing) std::fprintf(stderr, "[pe rank %d] pe_setup:%s\n", c->rank, t_report.c_str());
  c->jac_dt = -1;
  c->matrix_u_built = false;
  c->proj_matrix_ready = false;
  PE_CUDA(cudaStreamSynchronize(s));
  c->st = pe_stats{};
  c->st.n_cells = c->n_cells;
  c->st.n_dofs_p = c->fp.n_owned;
  c->st.n_dofs_u = c->fu.n_owned;
  c->st.nnz_p = c->fp.nnz;
  c->st.nnz_u = c->fu.nnz;
  c->st.spmv_bytes_p = (double)c->fp.nnz * 12.0 + (double)c->fp.n_owned * 20.0;
  c->st.spmv_bytes_u = (double)c->fu.nnz * 12.0 + (double)c->fu.n_owned * 20.0;
  c->st.eig_max_m = c->eig_M;
  c->st.setup_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
  c->setup_done = true;
  PE_LEAVE(c)
}

#define PE_NEED_SETUP(c) require((c)->setup_done, PE_ERR_STATE, "pe_setup has not been called")

int pe_pressure_set_uniform(pe_ctx* c, double v) {
  PE_ENTER(c)
  PE_NEED_SETUP(c);
  pe_vec_set(c, c->fp.n_local, v, c->p.p);
  PE_LEAVE(c)
}
int pe_pressure_begin_step(pe_ctx* c) {
  PE_ENTER(c)
  PE_NEED_SETUP(c);
  pe_vec_copy(c, c->fp.n_local, c->p.p, c->p_old.p);
  PE_LEAVE(c)
}
int pe_pressure_zero_update(pe_ctx* c) {
  PE_ENTER(c)
  PE_NEED_SETUP(c);
  PE_CUDA(cudaMemsetAsync(c->dp.p, 0, c->fp.n_local * sizeof(double), c->stream));
  PE_LEAVE(c)
}
int pe_pressure_update_volumetric_strain(pe_ctx* c) {
  PE_ENTER(c)
  PE_NEED_SETUP(c);
  pe_vec_axpy(c, c->fp.n_owned, c->prm.biot_coef / c->prm.bulk_modulus, c->dp.p, c->ev.p);
  PE_LEAVE(c)
}
int pe_pressure_assemble_residual(pe_ctx* c, double dt, double* l2) {
  PE_ENTER(c)
  PE_NEED_SETUP(c);
  require(dt > 0, PE_ERR_BAD_INPUT, "time step must be positive");
  double n = pe_pressure_residual(c, dt);
  if (c->fp.hang.n) {  // constraints.condense(residual), PS:153; FSS:364/405 take the norm afterwards
    pe_condense_vector(c, c->fp, c->resid.p);
    n = std::sqrt(pe_vec_dot(c, c->fp, c->resid.p, c->resid.p));
  }
  if (l2) *l2 = n;
  if (std::isnan(n)) throw PeError(PE_ERR_NAN, "pressure residual is NaN");
  PE_LEAVE(c)
}
int pe_pressure_assemble_jacobian(pe_ctx* c, double dt) {
  PE_ENTER(c)
  PE_NEED_SETUP(c);
  require(dt > 0, PE_ERR_BAD_INPUT, "time step must be positive");
  if (c->jac_dt != dt) {  // J = M/(M_b dt) + (k/mu) K is constant while dt is (PS:162-167)
    const double* Ms = c->fp.hang.n ? c->Mc.p : c->M.p;  // condensed copies on hanging-node meshes (PS:168)
    const double* Ks = c->fp.hang.n ? c->Kc.p : c->K.p;
    pe_vec_axpby_vals(c, c->fp.nnz, 1. / c->prm.m_modulus / dt, Ms, c->prm.perm_over_visc, Ks, c->J.p);
    pe_extract_invdiag(c, c->fp, c->J.p, c->invdiag_J.p);
    const bool have_j = pe_all_ranks_agree(c, pe_build_sell(c, c->fp, 2, c->J.p, false));
    if (!have_j) c->fp.sell[2].B = 0;
    {  // FP32 twin for the passes inside the Chebyshev polynomial (as for the displacement matrix)
      const char* f32 = std::getenv("PE_CHEB_FP32");
      if (have_j && c->prm.preconditioner == PE_PRECOND_CHEBYSHEV && c->prm.chebyshev_degree > 1 && !(f32 && std::string(f32) == "0"))
        { if (!pe_all_ranks_agree(c, pe_build_sell(c, c->fp, 4, c->J.p, true))) c->fp.sell[4].B = 0; }
      else
        c->fp.sell[4].B = 0;
    }
    c->eig_J = 1.1 * pe_estimate_eig_max(c, c->fp, c->J.p, c->invdiag_J.p);
    c->st.eig_max_p = c->eig_J;
    c->jac_dt = dt;
  }
  PE_LEAVE(c)
}
int pe_pressure_solve(pe_ctx* c, int* its, double* res) {
  PE_ENTER(c)
  PE_NEED_SETUP(c);
  require(c->jac_dt > 0, PE_ERR_STATE, "assemble_jacobian first");
  CgResult r = pe_cg_solve(c, c->fp, c->J.p, c->invdiag_J.p, c->eig_J, c->dp.p, c->resid.p, c->prm.cg_rel_tol_pressure, true, &c->st.spmv_launches_p);
  pe_distribute_hanging(c, c->fp, c->dp.p);  // constraints.distribute(solution_update), PS:180
  c->st.cg_iterations_pressure += r.its;
  c->st.cg_solves_pressure++;
  if (its) *its = r.its;
  if (res) *res = r.res;
  if (r.status != PE_OK) throw PeError(r.status, r.status == PE_ERR_NCCL ? "pressure CG: peer communication timed out" : "pressure CG did not converge (SolverControl::NoConvergence, PS:175)");
  PE_LEAVE(c)
}
int pe_pressure_add_update(pe_ctx* c) {
  PE_ENTER(c)
  PE_NEED_SETUP(c);
  pe_vec_axpy(c, c->fp.n_owned, 1.0, c->dp.p, c->p.p);
  PE_LEAVE(c)
}
int pe_pressure_linfty(pe_ctx* c, double* v) {
  PE_ENTER(c)
  PE_NEED_SETUP(c);
  require(v != nullptr, PE_ERR_BAD_INPUT, "null output");
  *v = pe_linfty(c, c->fp, c->p.p);
  PE_LEAVE(c)
}

int pe_displacement_assemble(pe_ctx* c) {
  PE_ENTER(c)
  PE_NEED_SETUP(c);
  if (!c->matrix_u_built) {  // rebuild_system_matrix (DS:137, DS:280-290)
    c->fu.bsr.B = 0;
    pe_assemble_elasticity(c);
    if (c->fu.hang.n) {
      // The cell kernel treated hanging dofs as free.  Lines that carry an inhomogeneity (a hanging node with a parent on
      // a Dirichlet face) are eliminated like Dirichlet values: b_const -= A g~ with g~ = g_h on hanging dofs; then the
      // matrix is condensed (constrained diagonal = the assembled one = sum of the cells' |a_hh|, DS:281-283).
      if (c->fu.hang.any_g) {
        pe_scatter_hanging_inhomogeneity(c, c->fu, c->w_d.p);
        pe_spmv_plain(c, c->fu, c->A.p, c->w_d.p, c->w_h.p);
        pe_vec_axpy(c, c->fu.n_owned, -1.0, c->w_h.p, c->b_const.p);
      }
      DBuf<double> uncondensed;
      uncondensed.alloc((size_t)c->fu.nnz);
      PE_CUDA(cudaMemcpyAsync(uncondensed.p, c->A.p, c->fu.nnz * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
      pe_condense_matrix(c, c->fu, uncondensed.p, c->A.p, true, 0.0);
      PE_CUDA(cudaStreamSynchronize(c->stream));
    }
    pe_extract_invdiag(c, c->fu, c->A.p, c->invdiag_A.p);
    {
      const char* fmt = std::getenv("PE_FORMAT");
      const bool want_bsr = !(fmt && std::string(fmt) == "csr");
      if (want_bsr && pe_all_ranks_agree(c, pe_build_bsr(c, c->fu, c->A.p))) {
        const Field::Bsr& S = c->fu.bsr;
        c->st.spmv_bytes_u = (double)S.nnzb * (S.B * S.B * 8.0 + 4.0) + (double)S.n_brows * 4.0 + (double)c->fu.n_owned * 16.0;
        c->st.bsr_block_size = S.B;
        pe_build_bsr_fp32(c, c->fu);
        // TMA-fed sliced copy (+ its FP32 twin for the passes inside the Chebyshev polynomial unless PE_CHEB_FP32=0)
        const bool have_sell = pe_all_ranks_agree(c, pe_build_sell(c, c->fu, 0, c->A.p, false));
        if (!have_sell) c->fu.sell[0].B = 0;
        c->st.sell_format_u = have_sell ? 1 : 0;
        const char* f32 = std::getenv("PE_CHEB_FP32");
        if (have_sell && c->prm.preconditioner == PE_PRECOND_CHEBYSHEV && c->prm.chebyshev_degree > 1 && !(f32 && std::string(f32) == "0"))
          { if (!pe_all_ranks_agree(c, pe_build_sell(c, c->fu, 1, c->A.p, true))) c->fu.sell[1].B = 0; }
        else
          c->fu.sell[1].B = 0;
      } else {
        c->fu.bsr.B = 0;
        c->fu.bsr.bval32.release();
        c->st.bsr_block_size = 0;
        c->fu.sell[0].B = c->fu.sell[1].B = 0;
      }
    }
    c->eig_A = 1.1 * pe_estimate_eig_max(c, c->fu, c->A.p, c->invdiag_A.p);
    c->st.eig_max_u = c->eig_A;
    c->matrix_u_built = true;
  }
  pe_assemble_u_rhs(c);
  pe_condense_vector(c, c->fu, c->b.p);  // rows of hanging dofs go to their masters (DS:281-286); no-op on uniform meshes
  PE_LEAVE(c)
}
int pe_displacement_solve(pe_ctx* c, int* its, double* res) {
  PE_ENTER(c)
  PE_NEED_SETUP(c);
  require(c->matrix_u_built, PE_ERR_STATE, "displacement_assemble first");
  CgResult r = pe_cg_solve(c, c->fu, c->A.p, c->invdiag_A.p, c->eig_A, c->u.p, c->b.p, c->prm.cg_abs_tol_displacement, false, &c->st.spmv_launches_u);
  pe_distribute(c, c->fu, c->u.p);  // constraints.distribute(solution), DS:306
  pe_distribute_hanging(c, c->fu, c->u.p);
  c->st.cg_iterations_displacement += r.its;
  c->st.cg_solves_displacement++;
  if (its) *its = r.its;
  if (res) *res = r.res;
  if (r.status != PE_OK) throw PeError(r.status, r.status == PE_ERR_NCCL ? "displacement CG: peer communication timed out" : "displacement CG did not converge (SolverControl::NoConvergence, DS:299)");
  PE_LEAVE(c)
}

int pe_project_assemble_matrix(pe_ctx* c) {  // SP:101-106: a copy of the mass matrix; the storage is shared
  PE_ENTER(c)
  PE_NEED_SETUP(c);
  c->proj_matrix_ready = true;
  PE_LEAVE(c)
}
int pe_project_assemble_rhs(pe_ctx* c, int n_comp, const int32_t* comps) {
  PE_ENTER(c)
  PE_NEED_SETUP(c);
  require(n_comp >= 1 && n_comp <= 6 && comps, PE_ERR_BAD_INPUT, "bad component list");
  int32_t entries[6];
  for (int k = 0; k < n_comp; ++k) {
    require(comps[k] >= 0 && comps[k] < c->dim * c->dim, PE_ERR_BAD_INPUT, "tensor component out of range");
    entries[k] = sym_entry(c->dim, comps[k]);
  }
  pe_assemble_projection_rhs(c, n_comp, comps, entries);
  for (int k = 0; k < n_comp; ++k) pe_condense_vector(c, c->fp, c->proj_rhs[entries[k]].p);  // SP:193-194
  PE_LEAVE(c)
}
int pe_project_solve(pe_ctx* c, int entry, int* its) {
  PE_ENTER(c)
  PE_NEED_SETUP(c);
  require(c->proj_matrix_ready, PE_ERR_STATE, "project_assemble_matrix first");
  require(entry >= 0 && entry < c->n_stress, PE_ERR_BAD_INPUT, "rhs entry out of range");
  int64_t dummy = 0;
  const double* m_proj = c->fp.hang.n ? c->Mc.p : c->M.p;  // SP:104-105: the condensed mass matrix
  CgResult r = pe_cg_solve(c, c->fp, m_proj, c->invdiag_M.p, c->eig_M, c->strains[entry].p, c->proj_rhs[entry].p, c->prm.cg_rel_tol_projection, true, &dummy);
  pe_distribute_hanging(c, c->fp, c->strains[entry].p);  // SP:215
  c->st.spmv_launches_p += dummy;
  c->st.cg_iterations_projection += r.its;
  c->st.cg_solves_projection++;
  if (its) *its = r.its;
  if (r.status != PE_OK) throw PeError(r.status, "projection CG did not converge (SolverControl::NoConvergence, SP:209)");
  PE_LEAVE(c)
}
int pe_volumetric_strain_from_projection(pe_ctx* c, int n, const int32_t* entries, int as_initial) {
  PE_ENTER(c)
  PE_NEED_SETUP(c);
  PE_CUDA(cudaMemsetAsync(c->ev.p, 0, c->fp.n_local * sizeof(double), c->stream));
  for (int k = 0; k < n; ++k) {
    require(entries[k] >= 0 && entries[k] < c->n_stress, PE_ERR_BAD_INPUT, "rhs entry out of range");
    pe_vec_axpy(c, c->fp.n_owned, 1.0, c->strains[entries[k]].p, c->ev.p);
  }
  if (as_initial) pe_vec_copy(c, c->fp.n_local, c->ev.p, c->ev0.p);
  PE_LEAVE(c)
}
int pe_effective_stresses(pe_ctx* c) {
  PE_ENTER(c)
  PE_NEED_SETUP(c);
  pe_stress_kernel(c);
  PE_LEAVE(c)
}

int pe_spmv(pe_ctx* c, int matrix, const double* x_host, double* y_host, int reps, float* ms_per_rep) {
  PE_ENTER(c)
  PE_NEED_SETUP(c);
  MatRef m = mat_by_id(c, matrix);
  Field& F = *m.F;
  require(reps >= 1, PE_ERR_BAD_INPUT, "reps must be >= 1");
  double* x = c->w_d.p;
  double* y = c->w_h.p;
  if (x_host) PE_CUDA(cudaMemcpyAsync(x, x_host, F.n_owned * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  else pe_vec_set(c, F.n_local, 1.0, x);
  pe_halo_exchange(c, F, x);
  cudaEvent_t e0, e1;
  PE_CUDA(cudaEventCreate(&e0));
  PE_CUDA(cudaEventCreate(&e1));
  pe_spmv_plain(c, F, m.val, x, y);  // warm-up
  PE_CUDA(cudaEventRecord(e0, c->stream));
  for (int r = 0; r < reps; ++r) pe_spmv_plain(c, F, m.val, x, y);
  PE_CUDA(cudaEventRecord(e1, c->stream));
  PE_CUDA(cudaEventSynchronize(e1));
  float ms = 0;
  PE_CUDA(cudaEventElapsedTime(&ms, e0, e1));
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  if (ms_per_rep) *ms_per_rep = ms / reps;
  if (y_host) {
    PE_CUDA(cudaMemcpyAsync(y_host, y, F.n_owned * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    PE_CUDA(cudaStreamSynchronize(c->stream));
  }
  PE_LEAVE(c)
}

int pe_get_vector(pe_ctx* c, int which, double* host, int64_t n) {
  PE_ENTER(c)
  PE_NEED_SETUP(c);
  VecRef v = vec_by_id(c, which);
  require(host && n == v.F->n_owned, PE_ERR_BAD_INPUT, "size must equal the number of owned dofs");
  PE_CUDA(cudaMemcpyAsync(host, v.p, n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  pe_sync_checked(c);
  PE_LEAVE(c)
}
int pe_set_vector(pe_ctx* c, int which, const double* host, int64_t n) {
  PE_ENTER(c)
  PE_NEED_SETUP(c);
  VecRef v = vec_by_id(c, which);
  require(host && n == v.F->n_owned, PE_ERR_BAD_INPUT, "size must equal the number of owned dofs");
  PE_CUDA(cudaMemcpyAsync(v.p, host, n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  PE_CUDA(cudaStreamSynchronize(c->stream));
  PE_LEAVE(c)
}
int pe_get_matrix_size(pe_ctx* c, int matrix, int64_t* n_rows, int64_t* nnz) {
  PE_ENTER(c)
  PE_NEED_SETUP(c);
  MatRef m = mat_by_id(c, matrix);
  if (n_rows) *n_rows = m.F->n_owned;
  if (nnz) *nnz = m.F->nnz;
  PE_LEAVE(c)
}
int pe_get_matrix(pe_ctx* c, int matrix, int64_t* rowptr, int32_t* col, double* val) {
  PE_ENTER(c)
  PE_NEED_SETUP(c);
  MatRef m = mat_by_id(c, matrix);
  Field& F = *m.F;
  if (matrix == PE_MAT_ELASTICITY) require(c->matrix_u_built, PE_ERR_STATE, "elasticity matrix not assembled yet");
  if (matrix == PE_MAT_JACOBIAN) require(c->jac_dt > 0, PE_ERR_STATE, "jacobian not assembled yet");
  std::vector<int32_t> rp((size_t)F.n_owned + 1);
  PE_CUDA(cudaMemcpyAsync(rp.data(), F.rowptr.p, rp.size() * sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
  if (col) PE_CUDA(cudaMemcpyAsync(col, F.col.p, F.nnz * sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
  if (val) PE_CUDA(cudaMemcpyAsync(val, m.val, F.nnz * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  PE_CUDA(cudaStreamSynchronize(c->stream));
  if (rowptr) for (size_t i = 0; i < rp.size(); ++i) rowptr[i] = rp[i];
  PE_LEAVE(c)
}
int pe_set_profiling(pe_ctx* c, int on) {
  PE_ENTER(c)
  if (on && c->prof_ev.empty()) {
    c->prof_ev.resize(2 * PE_PROF_PAIRS);
    c->prof_field.assign(PE_PROF_PAIRS, 0);
    for (auto& e : c->prof_ev) PE_CUDA(cudaEventCreate(&e));
  }
  if (!on) pe_prof_flush(c);
  c->profiling = on != 0;
  PE_LEAVE(c)
}
int pe_get_stats(pe_ctx* c, pe_stats* s) {
  PE_ENTER(c)
  require(s != nullptr, PE_ERR_BAD_INPUT, "null output");
  pe_prof_flush(c);
  *s = c->st;
  PE_LEAVE(c)
}
int pe_reset_stats(pe_ctx* c) {
  PE_ENTER(c)
  c->st.cg_iterations_pressure = c->st.cg_iterations_displacement = c->st.cg_iterations_projection = 0;
  c->st.cg_solves_pressure = c->st.cg_solves_displacement = c->st.cg_solves_projection = 0;
  c->st.spmv_launches_p = c->st.spmv_launches_u = 0;
  c->st.kernel_launches = 0;
  pe_prof_flush(c);
  c->st.spmv_ms_p = c->st.spmv_ms_u = 0;
  c->st.spmv_timed_p = c->st.spmv_timed_u = 0;
  c->st.pcg_ms_p = c->st.pcg_ms_u = 0;
  c->st.pcg_iterations_p = c->st.pcg_iterations_u = 0;
  c->st.inner_ms_u = c->st.update_ms_u = c->st.reduce_ms_u = 0;
  c->st.wait_inner_ms_u = c->st.wait_cg_ms_u = c->st.wait_peer_ms_u = c->st.wait_update_ms_u = 0;
  for (double& v : c->st.phase_ms_p) v = 0;
  c->st.inner_passes_u = 0;
  PE_LEAVE(c)
}
int pe_synchronize(pe_ctx* c) {
  PE_ENTER(c)
  pe_sync_checked(c);
  PE_LEAVE(c)
}
void* pe_stream(pe_ctx* c) { return c ? (void*)c->stream : nullptr; }

}  // extern "C"
