// kernels_sell.cuh — the TMA-fed matrix stream of the CG kernels (included by kernels_solver.cu).
//
// Replaces SparseMatrix<double>::vmult as called by SolverCG at lib/include/PoroElasticPressureSolver.h:176-179,
// PoroElasticDisplacementSolver.h:300-305 and StrainProjector.h:210-214.
//
// Format ("sliced block ELL", built once per matrix by kernels_sell.cu): 32 consecutive block rows form a SLICE; a
// slice is stored as nb_max PANELS, panel j holding block j of each of the 32 block rows:
//     int32 col[32] | T val[B*B][32]          (128 + B*B*32*sizeof(T) bytes, always a multiple of 128)
// so a slice is ONE contiguous, 128-byte aligned byte range.  Rows shorter than the slice maximum are padded with
// zero blocks that point at the row itself (+1.6 % bytes on the 128^3 Q1 mesh).  8 + 4/B^2 bytes per scalar
// nonzero, as block CSR.
//
// Data movement: every warp owns a ring of NST shared-memory stages.  Lane 0 arms the stage's mbarrier with the
// byte count and issues ONE cp.async.bulk (the 1-D TMA copy, SASS UBLKCP) for up to SJ panels; the bytes travel
// DRAM -> L2 -> shared memory without passing through registers or the LSU issue slots, with an evict-first L2
// policy so the matrix stream does not push the gathered vector out of L2.  The lanes then read "their" column
// and values with conflict-free LDS (lane = block row), gather x, and accumulate in a fixed order — there is no
// cross-lane reduction at all, every lane ends up with the B row sums of its block row, so the store and every
// fused epilogue are coalesced.  Measured reason for this design: with per-lane global loads the FP32 copy of the
// matrix was exactly as slow as the FP64 one (profiles/README.md, round 2) — the pass was bound by load
// instructions / bytes in flight per warp, not by DRAM.
//
// Work distribution: slices are claimed dynamically (one atomic per 32 block rows, issued a whole stage ring ahead
// of use) so slow SMs or ranks with uneven halos do not set the pace.  Dot products stay bitwise reproducible:
// every slice writes its partial sums to a fixed slot; the warp that completes a group of 32 slices adds the
// group's partials in a fixed order; the kernel's last CTA (or every CTA of the persistent kernel) adds the group
// totals in a fixed order.  Who computed what never enters the result.
#pragma once

namespace sell {

constexpr int WARPS = 8;          // per CTA; one CTA per SM (the rings take ~180 KB of shared memory)
constexpr int THREADS = WARPS * 32;
constexpr int NST = 3;            // stages per warp
constexpr int STAGE_BYTES = 7680; // largest stage of any configuration
constexpr int RING_BYTES = NST * STAGE_BYTES;
constexpr int SMEM_BYTES = WARPS * RING_BYTES + WARPS * NST * 8 + 64;

template <int B, typename T>
struct Cfg {
  static constexpr int PANEL = 128 + B * B * 32 * (int)sizeof(T);
  static constexpr int SJ = (B == 1) ? 14 : (STAGE_BYTES / PANEL);  // panels per stage: (3,f64) 3, (3,f32) 6, (2,f64) 6, (2,f32) 12
  static_assert(SJ >= 1 && SJ * PANEL <= STAGE_BYTES, "stage does not fit");
  // panels per GROUP, the unit of the gather pipeline (stream_run): two groups' gathers are in flight per lane
  static constexpr int G = (B == 1) ? 7 : (B == 2 && sizeof(T) == 4) ? 6 : 3;
  static constexpr int NG = SJ / G;
  static_assert(SJ % G == 0, "a stage is a whole number of groups");
  // Gather pipeline on/off.  Measured on B200 (C4, C3; profiles/README.md round 2): it lifts the scalar passes (pressure
  // Jacobian 3.92 -> 4.24 TB/s), which wait on gather latency, and LOWERS the 3x3 block passes (FP64 0.716 -> 0.749 ms,
  // FP32 0.476 -> 0.545 ms): those are bound by L1/TEX throughput, and looking one stage ahead costs them a stage of TMA depth
  // (the next stage must have landed before the current one is consumed).
  static constexpr bool PIPELINED = (B == 1);
};

struct Mat {
  const char* panels;
  const int32_t* slice_ptr;  // n_slices + 1, in panels
  int n_slices;
  int first_boundary_slice;  // slices before this one reference no ghost column
  int chunk;                 // slices per claim (~64 KB of panels), 1..16
  int l2_resident;           // the panels are small enough to live in L2 between passes: no evict-first hint
  int boundary_early;        // multi-GPU: claim order [interior | boundary | interior] with the boundary slices ending at ~70 % of the
                             // pass instead of [interior | boundary]: the halo they produce leaves well before the pass ends, and the
                             // halo they consume (published at the end of the phase before) has ~60 % of a pass to arrive
  int64_t n_brows;
};

// scratch of the deterministic reductions and of the dynamic distribution (per context)
struct Work {
  unsigned* claim;    // next unclaimed slice
  double* spart;      // [NV][cap] per-slice partial sums
  unsigned* gcnt;     // [groups] slices of the group that have reported
  double* gpart;      // [NV][gcap] per-group sums
  int cap, gcap;
};

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ uint64_t evict_first_policy() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
// 1-D bulk copy global -> shared, completion counted in bytes on `bar` (TMA engine; SASS: UBLKCP).  policy != 0: L2 cache
// hint (evict-first for matrices that cannot stay in L2 anyway, so they do not push the gathered vector out); policy == 0:
// default policy — small matrices (a pressure Jacobian of an 8-GPU block is ~60 MB) then stay resident in the 126 MB L2
// from one pass to the next.
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar, uint64_t policy) {
  if (policy)
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(dst), "l"(src),
                 "r"(bytes), "r"(bar), "l"(policy)
                 : "memory");
  else
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ int lds_i32(uint32_t addr) {
  int v;
  asm volatile("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ double lds_val(uint32_t addr, double) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ double lds_val(uint32_t addr, float) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
  return (double)v;
}

// Per-warp state that survives between calls of stream() inside a persistent kernel: the mbarriers keep flipping.
struct Ring {
  uint32_t base;     // shared address of the warp's first stage
  uint32_t bar;      // shared address of the warp's first mbarrier
  uint32_t parity;   // bit s = parity the NEXT completion of stage s will have
};

__device__ __forceinline__ Ring ring_setup(char* smem, int warp, int lane) {
  Ring r;
  r.base = smem_addr(smem) + (uint32_t)warp * RING_BYTES;
  r.bar = smem_addr(smem) + (uint32_t)WARPS * RING_BYTES + (uint32_t)warp * NST * 8;
  r.parity = 0;
  if (lane == 0) {
    for (int s = 0; s < NST; ++s) mbar_init(r.bar + 8 * s, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  return r;
}

// ---- per-warp stream of claimed slices ------------------------------------------------------------------------------
// Claims: work is handed out in CHUNKS of m.chunk consecutive slices (~64 KB of matrix, so the atomic and the slice-pointer
// load are paid once per 64 KB whatever the block size).  A warp's first two chunks are static (its global warp index, then
// + n_warps: no atomic before the first TMA copy), every later one is `2 n_warps + atomicAdd(claim, 1)`.  Both latencies are off the
// critical path: the atomic for the chunk after next and the slice pointers of the next chunk are always in flight
// while the current chunk streams.
//
// The stream is split in two so that a persistent kernel can hide the start-up latency of a pass as well: stream_begin()
// only needs the MATRIX (which never changes), so it is called for the next pass BEFORE the grid barrier that ends the
// current one — slice pointers, first claim and the first NST bulk copies are in flight while the warp waits for the
// slowest warp of the grid; stream_run() consumes (and keeps issuing) behind the barrier.  stream_drain() waits for copies
// that were issued but will never be consumed (a kernel must not exit with bulk copies in flight into its shared memory).
struct Stream {
  const char* panels;
  const int32_t* slice_ptr;
  unsigned* claim;
  int n_slices, CH, n_chunks, n_warps, panel, sj;
  int v_a, v_nb, v_fb;       // claim order -> slice: [0, v_a) | [v_fb, v_fb + v_nb) | [v_a, v_fb)   (identity when v_a == v_fb)
  int c_cur, c_next, sp_cur, sp_cur_e, sp_next, sp_next_e, q, nq;
  unsigned a_nn;
  bool exhausted;
  int f_slice, f_j, f_np;  // fetch cursor: slice, next panel, panels of the slice
  int64_t f_base;
  int d_slice[NST], d_cnt[NST], d_chunk[NST];
  bool d_first[NST], d_last[NST], d_cend[NST];
  bool resident;
};

__device__ __forceinline__ int stream_map(const Stream& S, int v) {  // position in the claim order -> slice
  return v < S.v_a ? v : (v < S.v_a + S.v_nb ? S.v_fb + (v - S.v_a) : v - S.v_nb);
}
// lane q < CH: first panel (and one past the last) of the q-th slice of the chunk
__device__ __forceinline__ void stream_load_ptrs(const Stream& S, int chunk, int lane, int& b, int& e) {
  const int v = chunk * S.CH + lane;
  b = e = 0;
  if (chunk < S.n_chunks && lane < S.CH && v < S.n_slices) {
    const int s = stream_map(S, v);
    b = S.slice_ptr[s];
    e = S.slice_ptr[s + 1];
  }
}
__device__ __forceinline__ unsigned stream_claim(const Stream& S, int lane) {
  unsigned v = 0;
  if (lane == 0) v = atomicAdd(S.claim, 1u);
  return v;  // consumed (shuffled) one chunk later
}
__device__ __forceinline__ void stream_issue(Stream& S, int st, Ring& R, int lane, uint64_t policy) {
  while (S.f_j >= S.f_np && !S.exhausted) {
    if (++S.q >= S.nq) {  // next chunk
      S.c_cur = S.c_next;
      S.sp_cur = S.sp_next;
      S.sp_cur_e = S.sp_next_e;
      if (S.c_cur >= S.n_chunks) {  // checked BEFORE the next claim: a ticket taken by a warp that is done would be a lost chunk
        S.exhausted = true;
        break;
      }
      S.c_next = 2 * S.n_warps + (int)__shfl_sync(0xffffffffu, S.a_nn, 0);  // a_nn is valid: it was claimed when c_cur proved valid
      stream_load_ptrs(S, S.c_next, lane, S.sp_next, S.sp_next_e);
      if (S.c_next < S.n_chunks) S.a_nn = stream_claim(S, lane);
      S.q = 0;
      S.nq = min(S.CH, S.n_slices - S.c_cur * S.CH);
    }
    S.f_slice = stream_map(S, S.c_cur * S.CH + S.q);
    const int p0 = __shfl_sync(0xffffffffu, S.sp_cur, S.q), p1 = __shfl_sync(0xffffffffu, S.sp_cur_e, S.q);
    S.f_base = p0;
    S.f_np = p1 - p0;
    S.f_j = 0;
  }
  if (S.f_j >= S.f_np) {
    S.d_cnt[st] = 0;
    return;
  }
  const int cnt = min(S.sj, S.f_np - S.f_j);
  S.d_slice[st] = S.f_slice;
  S.d_cnt[st] = cnt;
  S.d_first[st] = S.f_j == 0;
  S.d_last[st] = S.f_j + cnt == S.f_np;
  S.d_chunk[st] = S.c_cur;
  S.d_cend[st] = S.q == S.nq - 1;  // (with d_last) the chunk ends here: its partial sums are complete
  if (lane == 0) {
    const uint32_t bytes = (uint32_t)cnt * S.panel;
    mbar_expect_tx(R.bar + 8 * st, bytes);
    bulk_g2s(R.base + (uint32_t)st * STAGE_BYTES, S.panels + (size_t)(S.f_base + S.f_j) * S.panel, bytes, R.bar + 8 * st, S.resident ? 0ull : policy);
  }
  S.f_j += cnt;
}

// Which chunk a warp starts on: ranks are CTA-major, so the warps of a CTA start on neighbouring slices and share gathered
// lines in L1.  (Tried and measured on B200, round 2: "even rounds" — parking warps so that n_chunks is a whole number of
// rounds, 1073 of 1184 warps for the 8582 slices of a 64^3 block — shortens the wait at the barrier behind the inner passes
// from 8 to 5 us but lengthens every pass by 2.7 us, because the streaming rate follows the number of warps in flight; the
// spread of the warps' finishing times is about one slice whatever the round structure.  All warps take part.)
__device__ __forceinline__ void stream_roles(int n_chunks, int cta, int n_ctas, int warp, int& rank, int& n_active) {
  n_active = min(n_chunks, n_ctas * WARPS);
  const int q = n_active / n_ctas, rem = n_active % n_ctas;
  rank = (warp < q + (cta < rem ? 1 : 0)) ? cta * q + min(cta, rem) + warp : -1;
}

template <int B, typename T>
__device__ __forceinline__ void stream_begin(Stream& S, const Mat& m, unsigned* claim, int cta, int n_ctas, int warp, Ring& R, int lane, uint64_t policy) {
  typedef Cfg<B, T> C;
  S.panels = m.panels;
  S.slice_ptr = m.slice_ptr;
  S.claim = claim;
  S.n_slices = m.n_slices;
  S.CH = m.chunk;
  S.n_chunks = (m.n_slices + m.chunk - 1) / m.chunk;
  int rank, n_active;
  stream_roles(S.n_chunks, cta, n_ctas, warp, rank, n_active);
  S.n_warps = n_active;
  S.panel = C::PANEL;
  S.sj = C::SJ;
  S.resident = m.l2_resident != 0;
  S.v_fb = min(m.first_boundary_slice, m.n_slices);
  S.v_nb = m.n_slices - S.v_fb;
  S.v_a = m.boundary_early ? max(0, min(S.v_fb, (int)(0.7f * (float)m.n_slices) - S.v_nb)) : S.v_fb;
  S.c_cur = rank >= 0 ? rank : S.n_chunks;  // the first two chunks of a warp are static: nothing to wait for at the start of a pass
  stream_load_ptrs(S, S.c_cur, lane, S.sp_cur, S.sp_cur_e);
  S.c_next = rank >= 0 ? rank + n_active : S.n_chunks;
  stream_load_ptrs(S, S.c_next, lane, S.sp_next, S.sp_next_e);
  S.a_nn = S.c_next < S.n_chunks ? stream_claim(S, lane) : 0u;  // chunk ids from here on: 2 n_active + ticket
  S.q = -1;
  S.nq = S.c_cur < S.n_chunks ? min(S.CH, S.n_slices - S.c_cur * S.CH) : 0;
  S.exhausted = S.c_cur >= S.n_chunks;
  S.f_slice = -1;
  S.f_j = 0;
  S.f_np = 0;
  S.f_base = 0;
#pragma unroll
  for (int st = 0; st < NST; ++st) stream_issue(S, st, R, lane, policy);
}

// Consumes the stream.  Hooks, each called once per slice by all lanes (warp-uniform control flow):
//   ready(slice)  ahead of the slice's first gather (halo wait of boundary slices);
//   pre(slice)    operand prefetch of the fused epilogue — issued as soon as the PREVIOUS slice's done() has run, i.e. while
//                 the slice's own first stage is already being gathered, so the loads fly for the whole slice;
//   done(slice, acc, chunk)  receives the B row sums of block row 32*slice + lane and, when the slice is the last one of its
//                 chunk, the chunk's number (else -1): partial sums are formed per CHUNK.
// Software pipeline (scalar matrices, Cfg::PIPELINED): a stage is consumed in GROUPS of Cfg::G panels, and the gathers of the
// next group (column loads from shared memory, then the x loads) are issued BEFORE the multiply-adds of the current one,
// across stage and slice boundaries, so a lane always has two groups' gathers in flight.  One CTA per SM has only 8 warps;
// a scalar pass has one gather per 12 bytes of matrix and was waiting on gather latency (3.9 of 6.5 TB/s at 128^3).
// x is gathered with plain cached loads.
template <int B, typename T, class Ready, class Pre, class Done>
__device__ __forceinline__ void stream_run_pipelined(Stream& S, const double* __restrict__ x, Ring& R, int lane, uint64_t policy, Ready&& ready, Pre&& pre,
                                                     Done&& done) {
  typedef Cfg<B, T> C;
  constexpr int G = C::G, NG = C::NG;
  constexpr int UNR = ((NST * NG) % 2 == 0) ? NST * NG : 2 * NST * NG;  // whole ring cycles, even (two gather buffers)
  if (S.d_cnt[0] == 0) return;
  double acc[B];
  double xv[2][G][B];
  // gathers of group g of the stage in ring slot st into buffer buf (all three compile-time after unrolling); panels beyond
  // the stage's count read nothing and contribute zeros
  auto fetch = [&](int st, int g, int buf) {
    if (g == 0) {
      if (S.d_first[st]) ready(S.d_slice[st]);
      while (!mbar_try_wait(R.bar + 8 * st, (R.parity >> st) & 1u)) {}
      R.parity ^= 1u << st;
    }
    const int cnt = S.d_cnt[st];
    const uint32_t sb = R.base + (uint32_t)st * STAGE_BYTES;
    int col[G];
#pragma unroll
    for (int jj = 0; jj < G; ++jj) col[jj] = g * G + jj < cnt ? lds_i32(sb + (g * G + jj) * C::PANEL + lane * 4) : 0;
#pragma unroll
    for (int jj = 0; jj < G; ++jj)
#pragma unroll
      for (int cc = 0; cc < B; ++cc) xv[buf][jj][cc] = g * G + jj < cnt ? ld_gather(x + (size_t)col[jj] * B + cc) : 0.0;
  };
  fetch(0, 0, 0);
  pre(S.d_slice[0]);
  bool more = true;
  while (more) {
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      const int st = (u / NG) % NST, g = u % NG, nx = (st + 1) % NST, cur = u & 1;
      const int cnt = S.d_cnt[st];
      // the next group's gathers fly during this group's arithmetic
      if (g + 1 < NG) fetch(st, g + 1, cur ^ 1);
      else if (S.d_cnt[nx]) fetch(nx, 0, cur ^ 1);
      if (g == 0 && S.d_first[st]) {
#pragma unroll
        for (int r = 0; r < B; ++r) acc[r] = 0.0;
      }
      const uint32_t sb = R.base + (uint32_t)st * STAGE_BYTES;
#pragma unroll
      for (int jj = 0; jj < G; ++jj) {
        if (g * G + jj < cnt) {
#pragma unroll
          for (int r = 0; r < B; ++r)
#pragma unroll
            for (int cc = 0; cc < B; ++cc)
              acc[r] += lds_val(sb + (g * G + jj) * C::PANEL + 128 + ((r * B + cc) * 32 + lane) * (int)sizeof(T), T()) * xv[cur][jj][cc];
        }
      }
      if (g == NG - 1) {  // the stage is consumed
        const bool last = S.d_last[st];
        const int slice = S.d_slice[st], chunk = S.d_cend[st] ? S.d_chunk[st] : -1;
        const int ncnt = S.d_cnt[nx];
        __syncwarp();  // every lane has read the stage: lane 0 may re-arm it
        stream_issue(S, st, R, lane, policy);
        if (last) {
          done(slice, acc, chunk);
          if (ncnt) pre(S.d_slice[nx]);  // the stage after a slice's last one is the first of the next slice
        }
        if (!ncnt) {
          more = false;
          break;
        }
      }
    }
  }
}

// One stage at a time: all gathers of the stage are issued (independent loads in flight), then values + FMAs.
template <int B, typename T, class Ready, class Pre, class Done>
__device__ __forceinline__ void stream_run_staged(Stream& S, const double* __restrict__ x, Ring& R, int lane, uint64_t policy, Ready&& ready, Pre&& pre,
                                                  Done&& done) {
  typedef Cfg<B, T> C;
  double acc[B];
  bool more = true;
  while (more) {
#pragma unroll
    for (int st = 0; st < NST; ++st) {
      if (!more) break;
      if (S.d_cnt[st] == 0) {
        more = false;
        break;
      }
      const int cnt = S.d_cnt[st];
      if (S.d_first[st]) {
        ready(S.d_slice[st]);
        pre(S.d_slice[st]);
#pragma unroll
        for (int r = 0; r < B; ++r) acc[r] = 0.0;
      }
      while (!mbar_try_wait(R.bar + 8 * st, (R.parity >> st) & 1u)) {}
      R.parity ^= 1u << st;
      const uint32_t sb = R.base + (uint32_t)st * STAGE_BYTES;
      int col[C::SJ];
#pragma unroll
      for (int jj = 0; jj < C::SJ; ++jj) col[jj] = jj < cnt ? lds_i32(sb + jj * C::PANEL + lane * 4) : 0;
      double xv[C::SJ][B];
#pragma unroll
      for (int jj = 0; jj < C::SJ; ++jj)
#pragma unroll
        for (int cc = 0; cc < B; ++cc) xv[jj][cc] = jj < cnt ? ld_gather(x + (size_t)col[jj] * B + cc) : 0.0;
#pragma unroll
      for (int jj = 0; jj < C::SJ; ++jj) {
        if (jj < cnt) {
#pragma unroll
          for (int r = 0; r < B; ++r)
#pragma unroll
            for (int cc = 0; cc < B; ++cc)
              acc[r] += lds_val(sb + jj * C::PANEL + 128 + ((r * B + cc) * 32 + lane) * (int)sizeof(T), T()) * xv[jj][cc];
        }
      }
      __syncwarp();  // every lane has read the stage: lane 0 may re-arm it
      const bool last = S.d_last[st];
      const int slice = S.d_slice[st], chunk = S.d_cend[st] ? S.d_chunk[st] : -1;
      stream_issue(S, st, R, lane, policy);
      if (last) done(slice, acc, chunk);
    }
  }
}

template <int B, typename T, class Ready, class Pre, class Done>
__device__ __forceinline__ void stream_run(Stream& S, const double* __restrict__ x, Ring& R, int lane, uint64_t policy, Ready&& ready, Pre&& pre,
                                           Done&& done) {
  if constexpr (Cfg<B, T>::PIPELINED) stream_run_pipelined<B, T>(S, x, R, lane, policy, ready, pre, done);
  else stream_run_staged<B, T>(S, x, R, lane, policy, ready, pre, done);
}

// Waits for the copies of a stream that was begun but is not going to be consumed.
__device__ __forceinline__ void stream_drain(Stream& S, Ring& R) {
#pragma unroll
  for (int st = 0; st < NST; ++st) {
    if (S.d_cnt[st] == 0) break;
    while (!mbar_try_wait(R.bar + 8 * st, (R.parity >> st) & 1u)) {}
    R.parity ^= 1u << st;
  }
}

// one pass, start to end (the stand-alone SpMV kernels)
template <int B, typename T, class Ready, class Pre, class Done>
__device__ __forceinline__ void stream(const Mat& m, const double* __restrict__ x, unsigned* claim, int cta, int n_ctas, int warp, Ring& R, int lane,
                                       uint64_t policy, Ready&& ready, Pre&& pre, Done&& done) {
  Stream S;
  stream_begin<B, T>(S, m, claim, cta, n_ctas, warp, R, lane, policy);
  stream_run<B, T>(S, x, R, lane, policy, ready, pre, done);
}

// Deterministic sums with dynamic work distribution, in two halves so that the atomic's round trip overlaps the next slice:
// sums_post() is called by all lanes when a CHUNK is complete (v = the lane's contributions over the chunk's slices, a fixed
// set processed in a fixed order by one warp) and returns the ticket of the chunk's group (meaningful in lane 0, not yet waited
// for); sums_finish() is called with it one chunk later (and once after the stream) and, if this warp's chunk completed its
// group of 32, adds the group's partials in chunk order.  Slot layout: spart[k*cap + chunk], gpart[k*gcap + group].
struct Pending {
  unsigned ticket;
  int slice;  // -1: nothing pending
};
template <int NV>
__device__ __forceinline__ Pending sums_post(const Work& w, int slice, double (&v)[NV], int lane) {
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    double t = v[k];
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    v[k] = t;
  }
  Pending p{0u, slice};
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < NV; ++k) __stcg(w.spart + (size_t)k * w.cap + slice, v[k]);
    // release only: a gpu-scope __threadfence() also invalidates the SM's L1 (CCTL.IVALL), which would throw away the
    // cached x lines once per slice; the acquire side runs once per 32 slices
    asm volatile("atom.release.gpu.global.add.u32 %0, [%1], 1;" : "=r"(p.ticket) : "l"(w.gcnt + (slice >> 5)) : "memory");
  }
  return p;
}
template <int NV>
__device__ __forceinline__ void sums_finish(const Work& w, int n_slices, Pending& p, int lane) {
  if (p.slice < 0) return;
  const int g = p.slice >> 5, g0 = g << 5;
  const unsigned old = __shfl_sync(0xffffffffu, p.ticket, 0);
  const int gsize = min(32, n_slices - g0);
  p.slice = -1;
  if ((int)old == gsize - 1) {  // this warp completed the group: add its partials in slice order
    __threadfence();
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      double t = lane < gsize ? __ldcg(w.spart + (size_t)k * w.cap + g0 + lane) : 0.0;
      for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
      if (lane == 0) __stcg(w.gpart + (size_t)k * w.gcap + g, t);
    }
    if (lane == 0) w.gcnt[g] = 0u;  // ready for the next pass
  }
}

// Fixed-order sum of the group totals by one CTA (all threads call; result valid in thread 0).
template <int NV>
__device__ __forceinline__ void sum_groups(const Work& w, int n_groups, double (&tot)[NV], double* s_buf /* NV*32 */) {
  const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    double t = 0.0;
    for (int i = threadIdx.x; i < n_groups; i += blockDim.x) t += __ldcg(w.gpart + (size_t)k * w.gcap + i);
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    if (lane == 0) s_buf[k * 32 + wp] = t;
  }
  __syncthreads();
  if (wp == 0) {
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      double t = lane < nw ? s_buf[k * 32 + lane] : 0.0;
      for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
      tot[k] = t;
    }
  }
}

}  // namespace sell
