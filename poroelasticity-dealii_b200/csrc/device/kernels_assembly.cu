// kernels_assembly.cu — K1-K4, K9: cell-loop assembly in FP64 with a deterministic coloured scatter.
//
// One warp owns one cell.  The reference-cell tables (shape values / gradients at the Gauss points,
// Q1 mapping gradients) are staged in shared memory once per CTA; each warp then computes the
// physical gradients J^-T grad(N) and JxW of its cell into its own shared-memory slice and the lanes
// split the local rows / node pairs.  Cells are processed colour by colour: two cells of one colour
// never share a vertex (hence never a dof), so the global "+=" needs no atomics and the summation
// order of every matrix / vector entry is fixed -> bitwise reproducible run to run.
//
// Reference code restated here (paths relative to /root/reference/lib/include):
//   pressure mass / Laplace ... MatrixCreator::create_mass_matrix / create_laplace_matrix, PS:96-101
//   well source ............... VectorTools::create_right_hand_side + SinglePhaseWell::value, PS:142-147, RHS:99-116
//   elasticity matrix ......... DS:216-246 with isotropic_gassman_tensor (CM:45-57), get_strain_tensor (CM:9-24)
//   Dirichlet elimination ..... ConstraintMatrix::distribute_local_to_global, DS:279-286
//   Neumann faces ............. DS:249-277
//   pore-pressure coupling .... DS:232-234  (alpha p_h tr(eps_i) JxW)
//   strain projection rhs ..... SP:159-196 with get_strain_tensor(grad) (CM:27-42)
#include <algorithm>
#include <cmath>

#include "pe_internal.cuh"

namespace {

// ------------------------------------------------------------------------------------------------
// host: reference-cell tables
// ------------------------------------------------------------------------------------------------
void gauss1d(int n, std::vector<double>& x, std::vector<double>& w) {
  if (n == 2) {
    const double a = 0.5 / std::sqrt(3.0);
    x = {0.5 - a, 0.5 + a};
    w = {0.5, 0.5};
  } else if (n == 3) {
    const double a = 0.5 * std::sqrt(0.6);
    x = {0.5 - a, 0.5, 0.5 + a};
    w = {5.0 / 18.0, 4.0 / 9.0, 5.0 / 18.0};
  } else
    throw PeError(PE_ERR_UNSUPPORTED, "only QGauss(2) and QGauss(3) are tabulated");
}

struct HostQuad { int nq; std::vector<double> pts, w; };
HostQuad tensor_gauss(int dim, int n1d) {
  std::vector<double> x, w;
  gauss1d(n1d, x, w);
  HostQuad Q;
  Q.nq = 1;
  for (int a = 0; a < dim; ++a) Q.nq *= n1d;
  Q.pts.resize((size_t)Q.nq * dim);
  Q.w.resize(Q.nq);
  for (int q = 0; q < Q.nq; ++q) {
    int r = q;
    double ww = 1.0;
    for (int a = 0; a < dim; ++a) {  // x runs fastest
      Q.pts[(size_t)q * dim + a] = x[r % n1d];
      ww *= w[r % n1d];
      r /= n1d;
    }
    Q.w[q] = ww;
  }
  return Q;
}

// unit support points in deal.II's FE_Q local order: vertices, lines, quads, hex
std::vector<double> support_points_unit(int dim, int degree) {
  std::vector<double> s;
  const int nv = 1 << dim;
  for (int v = 0; v < nv; ++v)
    for (int a = 0; a < dim; ++a) s.push_back((v >> a) & 1);
  if (degree == 1) return s;
  auto mid = [&](std::initializer_list<int> vs) {
    for (int a = 0; a < dim; ++a) {
      double m = 0;
      for (int v : vs) m += (v >> a) & 1;
      s.push_back(m / (double)vs.size());
    }
  };
  if (dim == 2) {
    mid({0, 2}); mid({1, 3}); mid({0, 1}); mid({2, 3});
    mid({0, 1, 2, 3});
  } else {
    mid({0, 2}); mid({1, 3}); mid({0, 1}); mid({2, 3});
    mid({4, 6}); mid({5, 7}); mid({4, 5}); mid({6, 7});
    mid({0, 4}); mid({1, 5}); mid({2, 6}); mid({3, 7});
    mid({0, 2, 4, 6}); mid({1, 3, 5, 7}); mid({0, 1, 4, 5}); mid({2, 3, 6, 7}); mid({0, 1, 2, 3}); mid({4, 5, 6, 7});
    mid({0, 1, 2, 3, 4, 5, 6, 7});
  }
  return s;
}

// 1D Lagrange polynomial through equidistant nodes on [0,1] that is 1 at `sigma`
void lagrange(int degree, double sigma, double x, double& val, double& der) {
  val = 1.0;
  der = 0.0;
  for (int t = 0; t <= degree; ++t) {
    const double nt = (double)t / degree;
    if (std::fabs(nt - sigma) < 1e-12) continue;
    // (val * f)' = der * f + val * f'
    const double f = (x - nt) / (sigma - nt), fp = 1.0 / (sigma - nt);
    der = der * f + val * fp;
    val = val * f;
  }
}

void tabulate(int dim, int degree, const std::vector<double>& pts, int npts, std::vector<double>& N, std::vector<double>& dN, int& ns) {
  std::vector<double> sup = support_points_unit(dim, degree);
  ns = (int)sup.size() / dim;
  N.assign((size_t)npts * ns, 0.0);
  dN.assign((size_t)npts * ns * dim, 0.0);
  for (int q = 0; q < npts; ++q)
    for (int s = 0; s < ns; ++s) {
      double v[3] = {1, 1, 1}, d[3] = {0, 0, 0};
      for (int a = 0; a < dim; ++a) lagrange(degree, sup[(size_t)s * dim + a], pts[(size_t)q * dim + a], v[a], d[a]);
      N[(size_t)q * ns + s] = v[0] * v[1] * v[2];
      for (int a = 0; a < dim; ++a) {
        double g = d[a];
        for (int b = 0; b < dim; ++b)
          if (b != a) g *= v[b];
        dN[((size_t)q * ns + s) * dim + a] = g;
      }
    }
}

void make_quad_tab(pe_ctx* c, const HostQuad& Q, QuadTab& T) {
  std::vector<double> N, dN;
  int ns;
  tabulate(c->dim, 1, Q.pts, Q.nq, N, dN, ns);
  T.nq = Q.nq;
  T.w.upload(Q.w, c->stream);
  T.geoN.upload(N, c->stream);
  T.geodN.upload(dN, c->stream);
}
void make_shape_tab(pe_ctx* c, int degree, const HostQuad& Q, ShapeTab& T) {
  std::vector<double> N, dN;
  int ns;
  tabulate(c->dim, degree, Q.pts, Q.nq, N, dN, ns);
  T.ns = ns;
  T.nq = Q.nq;
  T.N.upload(N, c->stream);
  T.dN.upload(dN, c->stream);
}

// ------------------------------------------------------------------------------------------------
// device helpers
// ------------------------------------------------------------------------------------------------
constexpr int ASM_WARPS = 4;

template <int DIM>
__device__ __forceinline__ double jac_inverse_T(const double* J, double* JiT) {  // J row-major J[a*DIM+b] = dx_a/dxi_b
  if (DIM == 2) {
    const double det = J[0] * J[3] - J[1] * J[2];
    const double id = 1.0 / det;
    // J^-1 = 1/det [ J3 -J1; -J2 J0 ];  J^-T = transpose
    JiT[0] = J[3] * id; JiT[1] = -J[2] * id;
    JiT[2] = -J[1] * id; JiT[3] = J[0] * id;
    return det;
  } else {
    const double c00 = J[4] * J[8] - J[5] * J[7], c01 = J[5] * J[6] - J[3] * J[8], c02 = J[3] * J[7] - J[4] * J[6];
    const double det = J[0] * c00 + J[1] * c01 + J[2] * c02;
    const double id = 1.0 / det;
    // cofactor matrix C (C_ab = cofactor of J_ab); J^-T = C / det
    JiT[0] = c00 * id; JiT[1] = c01 * id; JiT[2] = c02 * id;
    JiT[3] = (J[2] * J[7] - J[1] * J[8]) * id; JiT[4] = (J[0] * J[8] - J[2] * J[6]) * id; JiT[5] = (J[1] * J[6] - J[0] * J[7]) * id;
    JiT[6] = (J[1] * J[5] - J[2] * J[4]) * id; JiT[7] = (J[2] * J[3] - J[0] * J[5]) * id; JiT[8] = (J[0] * J[4] - J[1] * J[3]) * id;
    return det;
  }
}

// Per-warp geometry: fills JxW[nq], grad[nq*ns*DIM] (physical gradients of the `ns` scalar shapes) and
// optionally xq[nq*DIM].  Lanes run over (q, s) pairs; every lane of a pair recomputes J(q).
template <int DIM>
__device__ __forceinline__ void warp_geometry(int lane, const double* __restrict__ Xv, int nq, int ns, const double* __restrict__ s_w,
                                              const double* __restrict__ s_geodN, const double* __restrict__ s_geoN,
                                              const double* __restrict__ s_dN, double* __restrict__ JxW, double* __restrict__ grad,
                                              double* __restrict__ xq) {
  constexpr int VPC = 1 << DIM;
  for (int item = lane; item < nq * ns; item += 32) {
    const int q = item / ns, s = item - q * ns;
    double J[DIM * DIM], JiT[DIM * DIM];
#pragma unroll
    for (int k = 0; k < DIM * DIM; ++k) J[k] = 0.0;
#pragma unroll
    for (int v = 0; v < VPC; ++v) {
      const double* g = s_geodN + ((size_t)q * VPC + v) * DIM;
#pragma unroll
      for (int a = 0; a < DIM; ++a)
#pragma unroll
        for (int b = 0; b < DIM; ++b) J[a * DIM + b] += Xv[v * DIM + a] * g[b];
    }
    const double det = jac_inverse_T<DIM>(J, JiT);
    const double* dn = s_dN + ((size_t)q * ns + s) * DIM;
#pragma unroll
    for (int a = 0; a < DIM; ++a) {
      double t = 0.0;
#pragma unroll
      for (int b = 0; b < DIM; ++b) t += JiT[a * DIM + b] * dn[b];
      grad[((size_t)q * ns + s) * DIM + a] = t;
    }
    if (s == 0) {
      JxW[q] = det * s_w[q];
      if (xq) {
#pragma unroll
        for (int a = 0; a < DIM; ++a) {
          double x = 0.0;
#pragma unroll
          for (int v = 0; v < VPC; ++v) x += s_geoN[q * VPC + v] * Xv[v * DIM + a];
          xq[q * DIM + a] = x;
        }
      }
    }
  }
}

__device__ __forceinline__ int csr_find(const int32_t* __restrict__ col, int lo, int hi, int32_t target) {
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    const int32_t v = col[mid];
    if (v < target) lo = mid + 1; else hi = mid;
  }
  return lo;  // caller guarantees presence
}

__device__ __forceinline__ void stage(double* dst, const double* __restrict__ src, int n) {
  for (int i = threadIdx.x; i < n; i += blockDim.x) dst[i] = src[i];
}

struct CellArgs {
  const double* xyz;
  const int32_t* cell_vertices;
  const int32_t* cells;  // cells of this colour
  int n_cells;
};

template <int DIM>
__device__ __forceinline__ void load_vertices(int lane, const CellArgs& A, int cell, double* Xv) {
  constexpr int VPC = 1 << DIM;
  for (int i = lane; i < VPC * DIM; i += 32) {
    const int v = i / DIM, a = i - v * DIM;
    Xv[i] = A.xyz[(size_t)A.cell_vertices[(size_t)cell * VPC + v] * DIM + a];
  }
}

// ------------------------------------------------------------------------------------------------
// K1 + K4: pressure mass / Laplace matrices and the well source vector
// ------------------------------------------------------------------------------------------------
template <int DIM>
__global__ void __launch_bounds__(ASM_WARPS * 32)
k_pressure_matrices(CellArgs A, const int32_t* __restrict__ cell_dofs, int64_t n_owned, const int32_t* __restrict__ rowptr,
                    const int32_t* __restrict__ col, int nq, const double* __restrict__ g_w, const double* __restrict__ g_geoN,
                    const double* __restrict__ g_geodN, const double* __restrict__ g_N, double well_r2, double well_val,
                    double* __restrict__ M, double* __restrict__ K, double* __restrict__ frhs) {
  constexpr int VPC = 1 << DIM, NS = VPC;
  extern __shared__ double sm[];
  double* s_w = sm;
  double* s_geoN = s_w + nq;
  double* s_geodN = s_geoN + nq * VPC;
  double* s_N = s_geodN + nq * VPC * DIM;
  double* warp_base = s_N + nq * NS;
  const int per_warp = VPC * DIM + nq + nq * NS * DIM + nq * DIM;
  stage(s_w, g_w, nq);
  stage(s_geoN, g_geoN, nq * VPC);
  stage(s_geodN, g_geodN, nq * VPC * DIM);
  stage(s_N, g_N, nq * NS);
  __syncthreads();
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  double* Xv = warp_base + (size_t)w * per_warp;
  double* JxW = Xv + VPC * DIM;
  double* grad = JxW + nq;
  double* xq = grad + nq * NS * DIM;
  for (int ci = blockIdx.x * ASM_WARPS + w; ci < A.n_cells; ci += gridDim.x * ASM_WARPS) {
    const int cell = A.cells[ci];
    __syncwarp();
    load_vertices<DIM>(lane, A, cell, Xv);
    __syncwarp();
    warp_geometry<DIM>(lane, Xv, nq, NS, s_w, s_geodN, s_geoN, s_geodN /* Q1 field == mapping */, JxW, grad, xq);
    __syncwarp();
    const int32_t* cd = cell_dofs + (size_t)cell * NS;
    for (int pair = lane; pair < NS * NS; pair += 32) {
      const int a = pair / NS, b = pair - a * NS;
      const int32_t row = cd[a];
      if (row >= n_owned) continue;
      double m = 0.0, k = 0.0;
      for (int q = 0; q < nq; ++q) {
        const double* ga = grad + ((size_t)q * NS + a) * DIM;
        const double* gb = grad + ((size_t)q * NS + b) * DIM;
        double gg = 0.0;
#pragma unroll
        for (int x = 0; x < DIM; ++x) gg += ga[x] * gb[x];
        m += s_N[q * NS + a] * s_N[q * NS + b] * JxW[q];
        k += gg * JxW[q];
      }
      const int pos = csr_find(col, rowptr[row], rowptr[row + 1], cd[b]);
      M[pos] += m;
      K[pos] += k;
    }
    for (int a = lane; a < NS; a += 32) {
      const int32_t row = cd[a];
      if (row >= n_owned) continue;
      double f = 0.0;
      for (int q = 0; q < nq; ++q) {
        const double r2 = xq[q * DIM] * xq[q * DIM] + xq[q * DIM + 1] * xq[q * DIM + 1];  // RHS:106 (x,y only)
        const double fq = (r2 <= well_r2) ? well_val : 0.0;
        f += fq * s_N[q * NS + a] * JxW[q];
      }
      frhs[row] += f;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// K2: elasticity matrix (free-free entries, constrained diagonals) and the constant Dirichlet rhs part
// ------------------------------------------------------------------------------------------------
template <int DIM>
__global__ void __launch_bounds__(ASM_WARPS * 32)
k_elasticity(CellArgs A, const int32_t* __restrict__ cell_dofs, int ns, int64_t n_owned, const int32_t* __restrict__ rowptr,
             const int32_t* __restrict__ col, const int32_t* __restrict__ cline, const double* __restrict__ line_g, int nq,
             const double* __restrict__ g_w, const double* __restrict__ g_geodN, const double* __restrict__ g_dN, double lambda, double mu,
             double* __restrict__ Aval, double* __restrict__ b_const) {
  constexpr int VPC = 1 << DIM;
  extern __shared__ double sm[];
  double* s_w = sm;
  double* s_geodN = s_w + nq;
  double* s_dN = s_geodN + nq * VPC * DIM;
  double* warp_base = s_dN + (size_t)nq * ns * DIM;
  const int per_warp = VPC * DIM + nq + nq * ns * DIM;
  stage(s_w, g_w, nq);
  stage(s_geodN, g_geodN, nq * VPC * DIM);
  stage(s_dN, g_dN, nq * ns * DIM);
  __syncthreads();
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  double* Xv = warp_base + (size_t)w * per_warp;
  double* JxW = Xv + VPC * DIM;
  double* grad = JxW + nq;
  const int nloc = ns * DIM;
  for (int ci = blockIdx.x * ASM_WARPS + w; ci < A.n_cells; ci += gridDim.x * ASM_WARPS) {
    const int cell = A.cells[ci];
    __syncwarp();
    load_vertices<DIM>(lane, A, cell, Xv);
    __syncwarp();
    warp_geometry<DIM>(lane, Xv, nq, ns, s_w, s_geodN, nullptr, s_dN, JxW, grad, nullptr);
    __syncwarp();
    const int32_t* cd = cell_dofs + (size_t)cell * nloc;
    // free-free entries and |diagonal| of constrained rows: lanes over node pairs (a,b)
    for (int pair = lane; pair < ns * ns; pair += 32) {
      const int a = pair / ns, b = pair - a * ns;
      double G[DIM][DIM];
#pragma unroll
      for (int k = 0; k < DIM; ++k)
#pragma unroll
        for (int l = 0; l < DIM; ++l) G[k][l] = 0.0;
      for (int q = 0; q < nq; ++q) {
        const double* ga = grad + ((size_t)q * ns + a) * DIM;
        const double* gb = grad + ((size_t)q * ns + b) * DIM;
        const double jw = JxW[q];
#pragma unroll
        for (int k = 0; k < DIM; ++k)
#pragma unroll
          for (int l = 0; l < DIM; ++l) G[k][l] += ga[k] * gb[l] * jw;
      }
      double tr = 0.0;
#pragma unroll
      for (int k = 0; k < DIM; ++k) tr += G[k][k];
#pragma unroll
      for (int cc = 0; cc < DIM; ++cc) {
        const int32_t row = cd[a * DIM + cc];
        if (row >= n_owned) continue;
        const bool row_c = cline[row] >= 0;
        const int r0 = rowptr[row], r1 = rowptr[row + 1];
#pragma unroll
        for (int dd = 0; dd < DIM; ++dd) {
          const int32_t cj = cd[b * DIM + dd];
          const double val = lambda * G[cc][dd] + mu * G[dd][cc] + (cc == dd ? mu * tr : 0.0);
          if (row_c) {
            if (cj == row) Aval[csr_find(col, r0, r1, cj)] += fabs(val);  // keeps the matrix invertible (ConstraintMatrix)
          } else if (cline[cj] < 0) {
            Aval[csr_find(col, r0, r1, cj)] += val;
          }
        }
      }
    }
    // b_const[i] -= sum_{j constrained} a_ij g_j : lanes over local rows, sequential over j
    bool any = false;
    for (int j = lane; j < nloc; j += 32) {
      const int lj = cline[cd[j]];
      if (lj >= 0 && line_g[lj] != 0.0) any = true;
    }
    if (__any_sync(0xffffffffu, any)) {
      for (int i = lane; i < nloc; i += 32) {
        const int32_t row = cd[i];
        if (row >= n_owned || cline[row] >= 0) continue;
        const int a = i / DIM, cc = i - a * DIM;
        double acc = 0.0;
        for (int j = 0; j < nloc; ++j) {
          const int lj = cline[cd[j]];
          if (lj < 0) continue;
          const double gj = line_g[lj];
          if (gj == 0.0) continue;
          const int b = j / DIM, dd = j - b * DIM;
          double val = 0.0;
          for (int q = 0; q < nq; ++q) {
            const double* ga = grad + ((size_t)q * ns + a) * DIM;
            const double* gb = grad + ((size_t)q * ns + b) * DIM;
            double t = lambda * ga[cc] * gb[dd] + mu * ga[dd] * gb[cc];
            if (cc == dd) {
              double gg = 0.0;
#pragma unroll
              for (int k = 0; k < DIM; ++k) gg += ga[k] * gb[k];
              t += mu * gg;
            }
            val += t * JxW[q];
          }
          acc -= val * gj;
        }
        b_const[row] += acc;
      }
    }
  }
}

// Neumann faces (DS:249-277): one CTA walks the boundary faces in order, lanes over the scalar nodes.
template <int DIM>
__global__ void k_neumann(int n_faces, const int32_t* __restrict__ bface_cell, const int8_t* __restrict__ bface_local,
                          const int32_t* __restrict__ bface_id, const double* __restrict__ xyz, const int32_t* __restrict__ cell_vertices,
                          const int32_t* __restrict__ cell_dofs, int ns, int64_t n_owned, const int32_t* __restrict__ cline, int n_cond,
                          const int32_t* __restrict__ nm_label, const int32_t* __restrict__ nm_comp, const double* __restrict__ nm_value,
                          int nqf, const double* __restrict__ f_w /* nqf */, const double* __restrict__ f_geodN /* 2DIM*nqf*VPC*DIM */,
                          const double* __restrict__ f_N /* 2DIM*nqf*ns */, double* __restrict__ b_const) {
  constexpr int VPC = 1 << DIM;
  for (int f = 0; f < n_faces; ++f) {
    const int cell = bface_cell[f], face = bface_local[f], id = bface_id[f];
    const int axis = face / 2, side = face % 2;
    for (int l = 0; l < n_cond; ++l) {
      if (nm_label[l] != id) continue;
      const int comp = nm_comp[l];
      for (int s = threadIdx.x; s < ns; s += blockDim.x) {
        const int32_t row = cell_dofs[(size_t)cell * ns * DIM + s * DIM + comp];
        if (row >= n_owned || cline[row] >= 0) continue;
        double acc = 0.0;
        for (int q = 0; q < nqf; ++q) {
          double J[DIM * DIM], JiT[DIM * DIM];
          for (int k = 0; k < DIM * DIM; ++k) J[k] = 0.0;
          for (int v = 0; v < VPC; ++v) {
            const double* X = xyz + (size_t)cell_vertices[(size_t)cell * VPC + v] * DIM;
            const double* g = f_geodN + (((size_t)face * nqf + q) * VPC + v) * DIM;
            for (int a = 0; a < DIM; ++a)
              for (int b = 0; b < DIM; ++b) J[a * DIM + b] += X[a] * g[b];
          }
          const double det = jac_inverse_T<DIM>(J, JiT);
          // n ~ J^-T n_ref ; dS = |det| |J^-T n_ref| w
          double nv[DIM], nn = 0.0;
          for (int a = 0; a < DIM; ++a) { nv[a] = JiT[a * DIM + axis] * (side ? 1.0 : -1.0); nn += nv[a] * nv[a]; }
          nn = sqrt(nn);
          const double jxw = fabs(det) * nn * f_w[q];
          acc += f_N[((size_t)face * nqf + q) * ns + s] * (nm_value[l] * nv[comp] / nn) * jxw;
        }
        b_const[row] += acc;
      }
      __syncthreads();
    }
  }
}

// ------------------------------------------------------------------------------------------------
// K3: displacement rhs  b += alpha * sum_q p_h(x_q) d_c N_a JxW   (free owned rows)
// ------------------------------------------------------------------------------------------------
template <int DIM>
__global__ void __launch_bounds__(ASM_WARPS * 32)
k_u_rhs(CellArgs A, const int32_t* __restrict__ cd_u, int ns, const int32_t* __restrict__ cd_p, int64_t n_owned,
        const int32_t* __restrict__ cline, int nq, const double* __restrict__ g_w, const double* __restrict__ g_geodN,
        const double* __restrict__ g_dN, const double* __restrict__ g_Np /* nq*VPC */, const double* __restrict__ p, double alpha,
        double* __restrict__ b) {
  constexpr int VPC = 1 << DIM;
  extern __shared__ double sm[];
  double* s_w = sm;
  double* s_geodN = s_w + nq;
  double* s_dN = s_geodN + nq * VPC * DIM;
  double* s_Np = s_dN + (size_t)nq * ns * DIM;
  double* warp_base = s_Np + nq * VPC;
  const int per_warp = VPC * DIM + nq + nq * ns * DIM + nq;
  stage(s_w, g_w, nq);
  stage(s_geodN, g_geodN, nq * VPC * DIM);
  stage(s_dN, g_dN, nq * ns * DIM);
  stage(s_Np, g_Np, nq * VPC);
  __syncthreads();
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  double* Xv = warp_base + (size_t)w * per_warp;
  double* JxW = Xv + VPC * DIM;
  double* grad = JxW + nq;
  double* pq = grad + (size_t)nq * ns * DIM;
  const int nloc = ns * DIM;
  for (int ci = blockIdx.x * ASM_WARPS + w; ci < A.n_cells; ci += gridDim.x * ASM_WARPS) {
    const int cell = A.cells[ci];
    __syncwarp();
    load_vertices<DIM>(lane, A, cell, Xv);
    for (int q = lane; q < nq; q += 32) {  // pressure_fe_values.get_function_values (DS:211-212)
      double s = 0.0;
      for (int k = 0; k < VPC; ++k) s += p[cd_p[(size_t)cell * VPC + k]] * s_Np[q * VPC + k];
      pq[q] = s;
    }
    __syncwarp();
    warp_geometry<DIM>(lane, Xv, nq, ns, s_w, s_geodN, nullptr, s_dN, JxW, grad, nullptr);
    __syncwarp();
    for (int i = lane; i < nloc; i += 32) {
      const int32_t row = cd_u[(size_t)cell * nloc + i];
      if (row >= n_owned || cline[row] >= 0) continue;
      const int a = i / DIM, cc = i - a * DIM;
      double acc = 0.0;
      for (int q = 0; q < nq; ++q) acc += (alpha * pq[q] * grad[((size_t)q * ns + a) * DIM + cc]) * JxW[q];
      b[row] += acc;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// K9: projection rhs  rhs^c_i = sum_q psi_i eps_c(u_h) JxW  with QGauss(2)
// ------------------------------------------------------------------------------------------------
struct ProjOut { double* rhs[6]; int comp[6]; int n; };

template <int DIM>
__global__ void __launch_bounds__(ASM_WARPS * 32)
k_projection_rhs(CellArgs A, const int32_t* __restrict__ cd_u, int ns, const int32_t* __restrict__ cd_p, int64_t n_owned_p, int nq,
                 const double* __restrict__ g_w, const double* __restrict__ g_geodN, const double* __restrict__ g_dN,
                 const double* __restrict__ g_Np, const double* __restrict__ u, ProjOut out) {
  constexpr int VPC = 1 << DIM;
  extern __shared__ double sm[];
  double* s_w = sm;
  double* s_geodN = s_w + nq;
  double* s_dN = s_geodN + nq * VPC * DIM;
  double* s_Np = s_dN + (size_t)nq * ns * DIM;
  double* warp_base = s_Np + nq * VPC;
  const int nloc = ns * DIM;
  const int per_warp = VPC * DIM + nq + nq * ns * DIM + nloc + nq * DIM * DIM;
  stage(s_w, g_w, nq);
  stage(s_geodN, g_geodN, nq * VPC * DIM);
  stage(s_dN, g_dN, nq * ns * DIM);
  stage(s_Np, g_Np, nq * VPC);
  __syncthreads();
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  double* Xv = warp_base + (size_t)w * per_warp;
  double* JxW = Xv + VPC * DIM;
  double* grad = JxW + nq;
  double* ul = grad + (size_t)nq * ns * DIM;
  double* eps = ul + nloc;  // nq * DIM*DIM strain tensor at the q points
  for (int ci = blockIdx.x * ASM_WARPS + w; ci < A.n_cells; ci += gridDim.x * ASM_WARPS) {
    const int cell = A.cells[ci];
    __syncwarp();
    load_vertices<DIM>(lane, A, cell, Xv);
    for (int i = lane; i < nloc; i += 32) ul[i] = u[cd_u[(size_t)cell * nloc + i]];
    __syncwarp();
    warp_geometry<DIM>(lane, Xv, nq, ns, s_w, s_geodN, nullptr, s_dN, JxW, grad, nullptr);
    __syncwarp();
    for (int item = lane; item < nq * DIM * DIM; item += 32) {
      const int q = item / (DIM * DIM), ij = item - q * DIM * DIM, i = ij / DIM, j = ij - i * DIM;
      double gij = 0.0, gji = 0.0;  // du_i/dx_j, du_j/dx_i  (get_function_gradients, SP:165)
      for (int s = 0; s < ns; ++s) {
        const double* g = grad + ((size_t)q * ns + s) * DIM;
        gij += ul[s * DIM + i] * g[j];
        gji += ul[s * DIM + j] * g[i];
      }
      eps[item] = (i == j) ? gij : (gij + gji) / 2;  // CM:27-42
    }
    __syncwarp();
    for (int item = lane; item < out.n * VPC; item += 32) {
      const int cidx = item / VPC, i = item - cidx * VPC;
      const int32_t row = cd_p[(size_t)cell * VPC + i];
      if (row >= n_owned_p) continue;
      double acc = 0.0;
      for (int q = 0; q < nq; ++q) acc += s_Np[q * VPC + i] * eps[q * DIM * DIM + out.comp[cidx]] * JxW[q];
      out.rhs[cidx][row] += acc;
    }
  }
}

template <class KernelT>
void set_smem(KernelT k, size_t bytes) {
  if (bytes > 48 * 1024) PE_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
}

int asm_grid(pe_ctx* c, int64_t n_cells) {
  int64_t want = (n_cells + ASM_WARPS - 1) / ASM_WARPS;
  int64_t cap = (int64_t)c->sm_count * 8;
  return (int)std::max<int64_t>(1, std::min(want, cap));
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// host entry points
// ------------------------------------------------------------------------------------------------
void pe_build_tables(pe_ctx* c) {
  const int dim = c->dim;
  HostQuad Q2 = tensor_gauss(dim, 2);
  HostQuad Qu = tensor_gauss(dim, c->prm.degree_u + 1);
  make_quad_tab(c, Q2, c->q2);
  make_quad_tab(c, Qu, c->qu);
  make_shape_tab(c, 1, Q2, c->p_q2);
  make_shape_tab(c, 1, Qu, c->p_qu);
  make_shape_tab(c, c->prm.degree_u, Q2, c->us_q2);
  make_shape_tab(c, c->prm.degree_u, Qu, c->us_qu);
  c->usup.upload(support_points_unit(dim, c->prm.degree_u), c->stream);
}

// greedy vertex-conflict colouring in cell order; 2^dim colours on structured grids
void pe_color_cells(pe_ctx* c) {
  const int vpc = c->vpc;
  std::vector<uint64_t> vmask((size_t)c->n_vertices, 0);
  std::vector<uint8_t> color((size_t)c->n_cells);
  int n_colors = 0;
  for (int64_t cell = 0; cell < c->n_cells; ++cell) {
    uint64_t used = 0;
    for (int v = 0; v < vpc; ++v) used |= vmask[c->h_cell_vertices[cell * vpc + v]];
    int col = 0;
    while (col < 64 && (used >> col) & 1) ++col;
    if (col >= 64) throw PeError(PE_ERR_UNSUPPORTED, "cell colouring needs more than 64 colours");
    color[cell] = (uint8_t)col;
    n_colors = std::max(n_colors, col + 1);
    for (int v = 0; v < vpc; ++v) vmask[c->h_cell_vertices[cell * vpc + v]] |= (uint64_t)1 << col;
  }
  c->n_colors = n_colors;
  c->color_ptr.assign(n_colors + 1, 0);
  for (int64_t cell = 0; cell < c->n_cells; ++cell) c->color_ptr[color[cell] + 1]++;
  for (int k = 0; k < n_colors; ++k) c->color_ptr[k + 1] += c->color_ptr[k];
  std::vector<int32_t> list((size_t)c->n_cells);
  std::vector<int64_t> pos(c->color_ptr.begin(), c->color_ptr.end() - 1);
  for (int64_t cell = 0; cell < c->n_cells; ++cell) list[pos[color[cell]]++] = (int32_t)cell;
  c->color_cells.upload(list, c->stream);
}

template <int DIM>
static void assemble_pressure_t(pe_ctx* c) {
  Field& F = c->fp;
  const int nq = c->q2.nq, vpc = 1 << DIM;
  const size_t smem = sizeof(double) * ((size_t)nq + nq * vpc + nq * vpc * DIM + nq * vpc + ASM_WARPS * (vpc * DIM + nq + nq * vpc * DIM + nq * DIM));
  set_smem(k_pressure_matrices<DIM>, smem);
  const double rw = c->prm.well_radius;
  const double well_val = -c->prm.flow_rate / (3.1415926 * rw * rw);  // RHS:109 (truncated pi kept)
  for (int k = 0; k < c->n_colors; ++k) {
    CellArgs A{c->xyz.p, c->cell_vertices.p, c->color_cells.p + c->color_ptr[k], (int)(c->color_ptr[k + 1] - c->color_ptr[k])};
    if (!A.n_cells) continue;
    k_pressure_matrices<DIM><<<asm_grid(c, A.n_cells), ASM_WARPS * 32, smem, c->stream>>>(
        A, F.cell_dofs.p, F.n_owned, F.rowptr.p, F.col.p, nq, c->q2.w.p, c->q2.geoN.p, c->q2.geodN.p, c->p_q2.N.p, rw * rw, well_val, c->M.p,
        c->K.p, c->frhs.p);
    c->st.kernel_launches++;
  }
  PE_CUDA(cudaGetLastError());
}

void pe_assemble_pressure_matrices(pe_ctx* c) {
  PE_CUDA(cudaMemsetAsync(c->M.p, 0, c->fp.nnz * sizeof(double), c->stream));
  PE_CUDA(cudaMemsetAsync(c->K.p, 0, c->fp.nnz * sizeof(double), c->stream));
  PE_CUDA(cudaMemsetAsync(c->frhs.p, 0, c->fp.n_local * sizeof(double), c->stream));
  if (c->dim == 2) assemble_pressure_t<2>(c); else assemble_pressure_t<3>(c);
}

template <int DIM>
static void assemble_elasticity_t(pe_ctx* c) {
  Field& F = c->fu;
  const int nq = c->qu.nq, vpc = 1 << DIM, ns = F.ns;
  const size_t smem = sizeof(double) * ((size_t)nq + nq * vpc * DIM + (size_t)nq * ns * DIM + ASM_WARPS * ((size_t)vpc * DIM + nq + (size_t)nq * ns * DIM));
  set_smem(k_elasticity<DIM>, smem);
  for (int k = 0; k < c->n_colors; ++k) {
    CellArgs A{c->xyz.p, c->cell_vertices.p, c->color_cells.p + c->color_ptr[k], (int)(c->color_ptr[k + 1] - c->color_ptr[k])};
    if (!A.n_cells) continue;
    k_elasticity<DIM><<<asm_grid(c, A.n_cells), ASM_WARPS * 32, smem, c->stream>>>(
        A, F.cell_dofs.p, ns, F.n_owned, F.rowptr.p, F.col.p, F.cline.p, F.line_g.p, nq, c->qu.w.p, c->qu.geodN.p, c->us_qu.dN.p,
        c->prm.lame_lambda, c->prm.shear_modulus, c->A.p, c->b_const.p);
    c->st.kernel_launches++;
  }
  PE_CUDA(cudaGetLastError());
  // Neumann part of the constant right-hand side
  if (!c->nm_label.empty() && c->n_bfaces > 0) {
    const int dim = DIM, n1d = c->prm.degree_u + 1;
    HostQuad Qf;
    if (dim == 2) {
      std::vector<double> x, w;
      gauss1d(n1d, x, w);
      Qf.nq = n1d; Qf.pts = x; Qf.w = w;
    } else
      Qf = tensor_gauss(2, n1d);
    const int nqf = Qf.nq;
    std::vector<double> f_geodN, f_N;
    for (int face = 0; face < 2 * dim; ++face) {
      const int axis = face / 2, side = face % 2;
      std::vector<double> pts((size_t)nqf * dim);
      for (int q = 0; q < nqf; ++q) {
        int t = 0;
        for (int a = 0; a < dim; ++a) pts[(size_t)q * dim + a] = (a == axis) ? (double)side : Qf.pts[(size_t)q * (dim - 1) + t++];
      }
      std::vector<double> N, dN;
      int nsx;
      tabulate(dim, 1, pts, nqf, N, dN, nsx);
      f_geodN.insert(f_geodN.end(), dN.begin(), dN.end());
      tabulate(dim, c->prm.degree_u, pts, nqf, N, dN, nsx);
      f_N.insert(f_N.end(), N.begin(), N.end());
    }
    DBuf<double> d_w, d_geodN, d_N, d_val;
    DBuf<int32_t> d_label, d_comp;
    d_w.upload(Qf.w, c->stream);
    d_geodN.upload(f_geodN, c->stream);
    d_N.upload(f_N, c->stream);
    d_label.upload(c->nm_label, c->stream);
    d_comp.upload(c->nm_comp, c->stream);
    d_val.upload(c->nm_value, c->stream);
    k_neumann<DIM><<<1, 32, 0, c->stream>>>((int)c->n_bfaces, c->bface_cell.p, c->bface_local.p, c->bface_id.p, c->xyz.p, c->cell_vertices.p,
                                            F.cell_dofs.p, ns, F.n_owned, F.cline.p, (int)c->nm_label.size(), d_label.p, d_comp.p, d_val.p, nqf,
                                            d_w.p, d_geodN.p, d_N.p, c->b_const.p);
    c->st.kernel_launches++;
    PE_CUDA(cudaStreamSynchronize(c->stream));
    PE_CUDA(cudaGetLastError());
  }
}

void pe_assemble_elasticity(pe_ctx* c) {
  PE_CUDA(cudaMemsetAsync(c->A.p, 0, c->fu.nnz * sizeof(double), c->stream));
  PE_CUDA(cudaMemsetAsync(c->b_const.p, 0, c->fu.n_local * sizeof(double), c->stream));
  if (c->dim == 2) assemble_elasticity_t<2>(c); else assemble_elasticity_t<3>(c);
}

template <int DIM>
static void assemble_u_rhs_t(pe_ctx* c) {
  Field& F = c->fu;
  const int nq = c->qu.nq, vpc = 1 << DIM, ns = F.ns;
  const size_t smem = sizeof(double) * ((size_t)nq + nq * vpc * DIM + (size_t)nq * ns * DIM + nq * vpc +
                                        ASM_WARPS * ((size_t)vpc * DIM + nq + (size_t)nq * ns * DIM + nq));
  set_smem(k_u_rhs<DIM>, smem);
  for (int k = 0; k < c->n_colors; ++k) {
    CellArgs A{c->xyz.p, c->cell_vertices.p, c->color_cells.p + c->color_ptr[k], (int)(c->color_ptr[k + 1] - c->color_ptr[k])};
    if (!A.n_cells) continue;
    k_u_rhs<DIM><<<asm_grid(c, A.n_cells), ASM_WARPS * 32, smem, c->stream>>>(A, F.cell_dofs.p, ns, c->fp.cell_dofs.p, F.n_owned, F.cline.p, nq,
                                                                              c->qu.w.p, c->qu.geodN.p, c->us_qu.dN.p, c->p_qu.N.p, c->p.p,
                                                                              c->prm.biot_coef, c->b.p);
    c->st.kernel_launches++;
  }
  PE_CUDA(cudaGetLastError());
}

void pe_assemble_u_rhs(pe_ctx* c) {
  pe_halo_exchange(c, c->fp, c->p.p);  // pressure at ghost dofs of the local cells
  pe_vec_copy(c, c->fu.n_local, c->b_const.p, c->b.p);
  if (c->dim == 2) assemble_u_rhs_t<2>(c); else assemble_u_rhs_t<3>(c);
}

template <int DIM>
static void projection_rhs_t(pe_ctx* c, int n_comp, const int32_t* comps, const int32_t* entries) {
  Field& F = c->fu;
  const int nq = c->q2.nq, vpc = 1 << DIM, ns = F.ns, nloc = ns * DIM;
  const size_t smem = sizeof(double) * ((size_t)nq + nq * vpc * DIM + (size_t)nq * ns * DIM + nq * vpc +
                                        ASM_WARPS * ((size_t)vpc * DIM + nq + (size_t)nq * ns * DIM + nloc + nq * DIM * DIM));
  set_smem(k_projection_rhs<DIM>, smem);
  ProjOut out{};
  out.n = n_comp;
  for (int k = 0; k < n_comp; ++k) {
    out.rhs[k] = c->proj_rhs[entries[k]].p;
    out.comp[k] = comps[k];
    PE_CUDA(cudaMemsetAsync(out.rhs[k], 0, c->fp.n_local * sizeof(double), c->stream));
  }
  for (int k = 0; k < c->n_colors; ++k) {
    CellArgs A{c->xyz.p, c->cell_vertices.p, c->color_cells.p + c->color_ptr[k], (int)(c->color_ptr[k + 1] - c->color_ptr[k])};
    if (!A.n_cells) continue;
    k_projection_rhs<DIM><<<asm_grid(c, A.n_cells), ASM_WARPS * 32, smem, c->stream>>>(A, F.cell_dofs.p, ns, c->fp.cell_dofs.p, c->fp.n_owned, nq,
                                                                                       c->q2.w.p, c->q2.geodN.p, c->us_q2.dN.p, c->p_q2.N.p,
                                                                                       c->u.p, out);
    c->st.kernel_launches++;
  }
  PE_CUDA(cudaGetLastError());
}

void pe_assemble_projection_rhs(pe_ctx* c, int n_comp, const int32_t* comps, const int32_t* entries) {
  pe_halo_exchange(c, c->fu, c->u.p);
  if (c->dim == 2) projection_rhs_t<2>(c, n_comp, comps, entries); else projection_rhs_t<3>(c, n_comp, comps, entries);
}
