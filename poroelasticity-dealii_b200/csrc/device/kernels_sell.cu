// kernels_sell.cu — builds the sliced block-ELL copies the TMA-fed CG kernels stream (layout: kernels_sell.cuh).
//
// One-time per matrix (the displacement matrix after DS:155-291's first assembly, the pressure Jacobian whenever dt
// changes at PS:158-169, the projection/mass matrix after SP:101-106).  Source is the block-CSR copy (B = dim) or, for
// scalar fields, the CSR matrix itself read as 1x1 blocks.
#include <cstdlib>

#include "pe_internal.cuh"

namespace {

// panels of slice s = longest block row among its 32
__global__ void k_slice_panels(int n_slices, int64_t n_brows, const int32_t* __restrict__ bptr, int32_t* __restrict__ count) {
  const int s = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (s >= n_slices) return;
  const int64_t r = (int64_t)s * 32 + lane;
  int nb = r < n_brows ? bptr[r + 1] - bptr[r] : 0;
  for (int o = 16; o > 0; o >>= 1) nb = max(nb, __shfl_xor_sync(0xffffffffu, nb, o));
  if (lane == 0) count[s] = nb;
}

// one warp per slice, lane = block row.  Block-CSR values are component-major inside a block row:
// value k of block j of block row I sits at B*B*bptr[I] + k*nb_I + j (for B = 1 that is plain CSR).
template <int B, typename T>
__global__ void k_fill_panels(int n_slices, int64_t n_brows, const int32_t* __restrict__ bptr, const int32_t* __restrict__ bcol,
                              const double* __restrict__ bval, const int32_t* __restrict__ slice_ptr, char* __restrict__ panels) {
  constexpr int PANEL = 128 + B * B * 32 * (int)sizeof(T);
  const int s = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (s >= n_slices) return;
  const int64_t r = (int64_t)s * 32 + lane;
  const bool real = r < n_brows;
  const int base = real ? bptr[r] : 0, nb = real ? bptr[r + 1] - base : 0;
  const int p0 = slice_ptr[s], np = slice_ptr[s + 1] - p0;
  for (int j = 0; j < np; ++j) {
    char* P = panels + (size_t)(p0 + j) * PANEL;
    const bool have = j < nb;
    reinterpret_cast<int32_t*>(P)[lane] = have ? bcol[base + j] : (real ? (int32_t)r : 0);  // padding points at the row itself
    T* V = reinterpret_cast<T*>(P + 128);
#pragma unroll
    for (int k = 0; k < B * B; ++k) V[k * 32 + lane] = have ? (T)bval[(size_t)base * B * B + (size_t)k * nb + j] : (T)0;
  }
}

}  // namespace

bool pe_build_sell(pe_ctx* c, Field& F, int slot, const double* val, bool f32) {
  SellMat& S = F.sell[slot];
  S.B = 0;
  S.src = nullptr;
  S.panels.release();
  S.slice_ptr.release();
  static const bool off = std::getenv("PE_FORMAT") && std::string(std::getenv("PE_FORMAT")) != "sell";
  static const double max_pad = std::getenv("PE_SELL_MAX_PAD") ? std::atof(std::getenv("PE_SELL_MAX_PAD")) : 1.5;
  if (off) return false;
  const int B = F.ncomp;
  const int32_t *bptr, *bcol;
  const double* bval;
  int64_t n_brows, nnzb;
  if (B == 1) {
    bptr = F.rowptr.p; bcol = F.col.p; bval = val; n_brows = F.n_owned; nnzb = F.nnz;
  } else {
    if (F.bsr.B != B || val != c->A.p) return false;  // needs the block-CSR copy of this very matrix
    bptr = F.bsr.bptr.p; bcol = F.bsr.bcol.p; bval = F.bsr.bval.p; n_brows = F.bsr.n_brows; nnzb = F.bsr.nnzb;
  }
  if (n_brows == 0 || B > 3) return false;
  const int n_slices = (int)((n_brows + 31) / 32);
  S.slice_ptr.alloc((size_t)n_slices + 1);
  const int warps = 8;
  k_slice_panels<<<pe_div_up(n_slices, warps), warps * 32, 0, c->stream>>>(n_slices, n_brows, bptr, S.slice_ptr.p);
  const int64_t n_panels = pe_exclusive_scan_i32(c, S.slice_ptr.p, n_slices);
  c->st.kernel_launches++;
  if (n_panels <= 0 || (double)n_panels * 32.0 > max_pad * (double)nnzb) {  // too much padding: the row-wise kernels serve this matrix
    S.slice_ptr.release();
    return false;
  }
  const int panel_bytes = 128 + B * B * 32 * (f32 ? 4 : 8);
  S.panels.alloc((size_t)n_panels * panel_bytes);
  const int grid = pe_div_up(n_slices, warps);
#define PE_FILL(BB, TT) k_fill_panels<BB, TT><<<grid, warps * 32, 0, c->stream>>>(n_slices, n_brows, bptr, bcol, bval, S.slice_ptr.p, S.panels.p)
  if (B == 1) { if (f32) PE_FILL(1, float); else PE_FILL(1, double); }
  else if (B == 2) { if (f32) PE_FILL(2, float); else PE_FILL(2, double); }
  else { if (f32) PE_FILL(3, float); else PE_FILL(3, double); }
#undef PE_FILL
  c->st.kernel_launches++;
  PE_CUDA(cudaGetLastError());
  S.src = val;
  S.f32 = f32;
  S.B = B;
  S.panel_bytes = panel_bytes;
  S.n_slices = n_slices;
  S.n_brows = n_brows;
  S.n_panels = n_panels;
  S.nnzb = nnzb;
  S.first_boundary_slice = (int)(F.n_interior / (32 * B));  // n_interior is a multiple of 32*B (kernels_pattern.cu)
  // reduction scratch sized for the largest matrix seen so far
  if (n_slices > c->red.cap) {
    c->red.cap = n_slices;
    c->red.gcap = (n_slices + 31) / 32;
    c->red.spart.alloc((size_t)PE_SELL_NV * c->red.cap);
    c->red.gpart.alloc((size_t)PE_SELL_NV * c->red.gcap);
    c->red.gcnt.alloc_zero((size_t)c->red.gcap, c->stream);
  }
  if (!c->red.claim.p) c->red.claim.alloc_zero(6, c->stream);
  PE_CUDA(cudaStreamSynchronize(c->stream));
  return true;
}
