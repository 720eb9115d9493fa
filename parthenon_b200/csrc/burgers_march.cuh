// burgers_march.cuh — y / z sweeps of the FAST burgers stage (included by burgers_sweep.cu).
//
// A thread owns one (i, other) column of a meshblock and marches along the sweep direction
// (lanes along i: every access of a warp is one coalesced 256-byte row).  The march is cut
// into chunks of kChunk cells.  Per chunk the thread
//   1. reconstructs the sweep-direction velocity and turns every face's HLL wave speeds
//      (burgers_package.hpp:31-43) into two coefficients P, Q with  F = P qL + Q qR  for ALL
//      components of that face — they stay in registers for the chunk;
//   2. runs the other components one after the other (a rolled loop: small code, few live
//      registers) over the same chunk.  The stencil of a component lives in registers in
//      difference form: first differences and the curvature parts of the smoothness indicators
//      are computed ONCE per row and shared by the three cells that use them
//      (weno_fast.cuh: WENO5Z_diff), so a reconstruction costs ~85 FP64 instructions instead
//      of the ~165 of the reference's expression tree (recon.hpp:43-99);
//   3. in the LAST sweep of a stage finishes the chunk's cells: derived field, min dt
//      (burgers_package.cpp:143-200).
// The rows of a component are staged asynchronously (cp.async) into a per-thread shared-memory
// column one component ahead of the arithmetic.  What a component carries from chunk to chunk
// (left state and flux of the last face) sits in a per-thread shared-memory column as well; the
// four overlapping stencil rows are re-read at a chunk start (L2 hits).
//
// Ghost rows: with Args::nbr (the mesh's same-device neighbour table) the stencil rows beyond a
// block face are read from the INTERIOR rows of the block on the other side of that face instead
// of from the block's own ghost rows — the values SendBoundBufs + SetBounds
// (boundary_communication.cpp:95-140, :273-334) would have copied there.  The caller can then
// leave the same-device ghost exchange out of the cycle and refresh ghost cells only when
// something else wants to read them.  Faces without a same-device neighbour (physical
// boundary, another GPU) keep reading the block's ghost rows.
//
// GEO = 32 fixes the geometry at compile time (32^3 cells, 4 ghosts, 3-D: the benchmark's
// block): every component / row offset becomes an immediate of the load or store.
#pragma once

namespace pb2 {
namespace PB2_SWEEP_NS {

#ifndef PB2_CHUNK
#define PB2_CHUNK 4
#endif
#ifndef PB2_L2_HINTS
#define PB2_L2_HINTS 0
#endif
#ifndef PB2_CHUNK_MINB
#define PB2_CHUNK_MINB 4
#endif
constexpr int kChunk = PB2_CHUNK;

template <int GEO>
struct GeoT;
template <>
struct GeoT<0> {
  const Geom &g;
  __device__ __forceinline__ explicit GeoT(const Geom &g_) : g(g_) {}
  __device__ __forceinline__ int nx(int d) const { return g.nx[d]; }
  __device__ __forceinline__ int is(int d) const { return g.is[d]; }
  __device__ __forceinline__ int n(int d) const { return g.n[d]; }
  __device__ __forceinline__ int ndim() const { return g.ndim; }
  __device__ __forceinline__ int64_t sj() const { return g.sj; }
  __device__ __forceinline__ int64_t sk() const { return g.sk; }
  __device__ __forceinline__ int64_t sc() const { return g.sc; }
};
template <>
struct GeoT<32> {
  __device__ __forceinline__ explicit GeoT(const Geom &) {}
  __device__ __forceinline__ constexpr int nx(int) const { return 32; }
  __device__ __forceinline__ constexpr int is(int) const { return 4; }
  __device__ __forceinline__ constexpr int n(int) const { return 40; }
  __device__ __forceinline__ constexpr int ndim() const { return 3; }
  __device__ __forceinline__ constexpr int64_t sj() const { return 40; }
  __device__ __forceinline__ constexpr int64_t sk() const { return 1600; }
  __device__ __forceinline__ constexpr int64_t sc() const { return 64000; }
};
inline bool geo32(const Geom &g) {
  return g.ndim == 3 && g.nx[0] == 32 && g.nx[1] == 32 && g.nx[2] == 32 && g.is[0] == 4;
}

using fastmath::Linear_diff;
using fastmath::weno_curv;
using fastmath::WENO5Z_diff;

// HLL coefficients of a face: F = P qL + Q qR (velocities carry an extra 1/2, folded into the
// update coefficient).  From  F = (sr upl qL - sl upr qR + sl sr (qR - qL)) / (sr - sl + [sl sr
// == 0])  (burgers_package.cpp:326-333):  P = sr (upl - sl) inv,  Q = sl (sr - upr) inv.
__device__ __forceinline__ void face_pq(const double upl, const double upr, double &P,
                                        double &Q) {
  const double sl = min_std(min_std(upl, upr), 0.0);
  const double sr = max_std(max_std(upl, upr), 0.0);
  const double inv = rcp_fast(sr - sl + (sl * sr == 0.0 ? 1.0 : 0.0));
  P = (sr * inv) * (upl - sl);
  Q = (sl * inv) * (sr - upr);
}

// ---- asynchronous staging: global -> per-thread shared-memory column -------------------------
// Every thread copies the rows of ITS OWN column (cp.async, 8 bytes each) into its own column
// of a stage buffer, so completion is a per-thread cp.async.wait_group: no barrier, no bank
// conflicts.  While component m of a chunk is being computed the rows of component m+1 (or of
// the next chunk's first component) are in flight.
__device__ __forceinline__ void cp_async8(const uint32_t saddr, const void *g) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(saddr), "l"(g) : "memory");
}
// ... with an L2 eviction priority: the rows two consecutive chunks share are fetched
// "evict_last" the first time and "evict_first" the second (their last use), everything that is
// read once streams through "evict_first" — the re-read of a row then finds it in L2 even though
// the rows all resident CTAs touch between the two reads (~100 MB) are as large as the L2
__device__ __forceinline__ void cp_async8_hint(const uint32_t saddr, const void *g,
                                               const uint64_t policy) {
  asm volatile("cp.async.ca.shared.global.L2::cache_hint [%0], [%1], 8, %2;" ::"r"(saddr), "l"(g),
               "l"(policy)
               : "memory");
}
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ void cp_async_commit() {
  asm volatile("cp.async.commit_group;" ::: "memory");
}
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

template <int RECON>
struct StencilRows {
  static constexpr int kLo = RECON == PB2_RECON_WENO5 ? -2 : -1;   // first row, relative to cell
  static constexpr int kExtra = RECON == PB2_RECON_WENO5 ? 4 : 2;  // rows beyond the chunk's cells
};
constexpr int kURows = kChunk + 4; // rows of u per stage buffer (weno5)
constexpr int kStageDoubles = kURows + kChunk; // + the old `out` values of the chunk's cells

// rows of u (stencil) and old values of out for CH cells starting at march cell s0; pu / po
// point at the component's row of cell s0.  EDGE: some rows lie beyond the block's ends; they
// come from the neighbour block (element offsets dlo / dhi from the own-row address, 0 = own
// ghost row).
template <int RECON, int CH, bool EDGE>
__device__ __forceinline__ void stage_issue(const uint32_t st, const double *__restrict__ pu,
                                            const double *po, const int64_t sd, const int s0,
                                            const int nd, const long long dlo,
                                            const long long dhi, const uint64_t keep,
                                            const uint64_t stream) {
  constexpr int lo = StencilRows<RECON>::kLo, nrow = CH + StencilRows<RECON>::kExtra;
#pragma unroll
  for (int r = 0; r < nrow; ++r) {
    const double *src = pu + (lo + r) * sd;
    if (EDGE) {
      const int rr = s0 + lo + r;
      src += rr < 0 ? dlo : (rr >= nd ? dhi : 0);
    }
    // the last kExtra rows are the first rows of the next chunk
#if PB2_L2_HINTS
    cp_async8_hint(st + r * kThreads * 8, src, r >= CH ? keep : stream);
#else
    cp_async8(st + r * kThreads * 8, src);
#endif
  }
#pragma unroll
  for (int t = 0; t < CH; ++t)
    if (s0 + t >= 1) {
#if PB2_L2_HINTS
      cp_async8_hint(st + (kURows + t) * kThreads * 8, po + (t - 1) * sd, stream);
#else
      cp_async8(st + (kURows + t) * kThreads * 8, po + (t - 1) * sd);
#endif
    }
}

// ---- the marching stencil of one component -------------------------------------------------
// rows come from the thread's stage column: su[r * kThreads] = row (cell s0 + kLo + r)
template <int RECON>
struct MarchStencil;

template <>
struct MarchStencil<PB2_RECON_WENO5> {
  double qc, qp1, qp2;     // q_c, q_{c+1}, q_{c+2}
  double d1, d2, d3, d4;   // q_{c-1}-q_{c-2}, q_c-q_{c-1}, q_{c+1}-q_c, q_{c+2}-q_{c+1}
  double A0, A1, A2;       // curvature terms centred at c-1, c, c+1
  __device__ __forceinline__ void init(const double *su) {
    const double qm2 = su[0], qm1 = su[kThreads];
    qc = su[2 * kThreads];
    qp1 = su[3 * kThreads];
    qp2 = su[4 * kThreads];
    d1 = qm1 - qm2;
    d2 = qc - qm1;
    d3 = qp1 - qc;
    d4 = qp2 - qp1;
    A0 = weno_curv(d1, d2);
    A1 = weno_curv(d2, d3);
    A2 = weno_curv(d3, d4);
  }
  __device__ __forceinline__ void recon(double &ql, double &qr) const {
    WENO5Z_diff(d1, d2, d3, d4, A0, A1, A2, qc, ql, qr);
  }
  // step t -> t+1: takes row t + 5 of the stage column
  __device__ __forceinline__ void advance(const double *su, const int t) {
    const double qnew = su[(t + 5) * kThreads];
    const double dn = qnew - qp2;
    d1 = d2;
    d2 = d3;
    d3 = d4;
    d4 = dn;
    A0 = A1;
    A1 = A2;
    A2 = weno_curv(d3, d4);
    qc = qp1;
    qp1 = qp2;
    qp2 = qnew;
  }
};

template <>
struct MarchStencil<PB2_RECON_LINEAR> {
  double qc, qp1, dm, dp;
  __device__ __forceinline__ void init(const double *su) {
    const double qm1 = su[0];
    qc = su[kThreads];
    qp1 = su[2 * kThreads];
    dm = qc - qm1;
    dp = qp1 - qc;
  }
  __device__ __forceinline__ void recon(double &ql, double &qr) const {
    Linear_diff(dm, dp, qc, ql, qr);
  }
  __device__ __forceinline__ void advance(const double *su, const int t) {
    const double qnew = su[(t + 3) * kThreads];
    dm = dp;
    dp = qnew - qp1;
    qc = qp1;
    qp1 = qnew;
  }
};

// CH cells of one component out of the stage column `st` (rows of u, then old values of out).
// Cell s gives face s its right state and cell s-1 its last flux; the previous cell's left
// state L and the previous face's flux F come in and go out.  po points at the component's
// row of cell s0 in `out`.
__device__ __forceinline__ void st_stream(double *p, const double v, const uint64_t policy) {
#if PB2_L2_HINTS
  asm volatile("st.global.L2::cache_hint.f64 [%0], %1, %2;" ::"l"(p), "d"(v), "l"(policy)
               : "memory");
#else
  *p = v;
#endif
}

template <int RECON, bool COEF, int CH>
__device__ __forceinline__ void comp_chunk(const double *st, double *po, const int64_t sd,
                                           const int s0, const double cd, double &L, double &F,
                                           double (&P)[CH], double (&Q)[CH],
                                           const uint64_t stream) {
  MarchStencil<RECON> ms;
  ms.init(st);
#pragma unroll
  for (int t = 0; t < CH; ++t) {
    double ql, qr;
    ms.recon(ql, qr);
    if (COEF) face_pq(L, qr, P[t], Q[t]);
    const double f = fma(P[t], L, Q[t] * qr);
    if (s0 + t >= 1) // face s and face s-1 are known: cell s-1 is complete
      st_stream(po + (t - 1) * sd, fma(cd, f - F, st[(kURows + t) * kThreads]), stream);
    F = f;
    L = ql;
    if (t + 1 < CH) ms.advance(st, t);
  }
}

// LAST sweep: derived field and time-step rate of the chunk's finished cells kk0 .. kk1
// (interior index along the march) of this thread's column, read back from `out`
template <int DIR, int GEO>
__device__ __forceinline__ void finish_chunk(const Args &a, const GeoT<GEO> &G, const int b,
                                             double *ob, const int64_t col0, const int kk0,
                                             const int kk1, const double idx0, const double idx1,
                                             const double idx2, double &rate) {
  const int64_t sd = (DIR == 1) ? G.sj() : G.sk();
  const int ds = G.is(DIR);
  const int ndim = G.ndim();
  for (int kb = kk0; kb <= kk1; kb += 4) {
    double v[4][4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int64_t cell = (int64_t)(ds + min(kb + u, kk1)) * sd;
#pragma unroll
      for (int n = 0; n < 4; ++n) v[u][n] = ob[n * G.sc() + cell];
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (kb + u > kk1) continue;
      const int64_t cell = (int64_t)(ds + kb + u) * sd;
      if (a.derived)
        a.derived[(int64_t)b * G.sc() + col0 + cell] =
            0.5 * v[u][3] * (v[u][0] * v[u][0] + v[u][1] * v[u][1] + v[u][2] * v[u][2]);
      const double r = fabs(v[u][0]) * idx0 + (ndim > 1 ? fabs(v[u][1]) * idx1 : 0.0) +
                       (ndim > 2 ? fabs(v[u][2]) * idx2 : 0.0);
      rate = max_std(rate, r);
    }
  }
}

// everything a chunk needs from the thread's set-up
struct ColumnCtx {
  const double *ub; // u,   block b, component 0, this column, row 0
  double *ob;       // out, same place
  double *sL, *sF;  // per-thread carry columns in shared memory (+ n * kThreads)
  double *stage;    // per-thread stage columns: [2][kStageDoubles] (+ i * kThreads)
  long long dlo, dhi; // ghost rows below / above the block: offset to the neighbour's rows
  uint64_t keep, stream; // L2 eviction policies (cp_async8_hint)
  int64_t col0;
  int b, nc, nd;
  double cdir, idx0, idx1, idx2;
};

// component processed m-th in a chunk: the sweep direction's velocity first (its states give
// the face coefficients), then the others in order
template <int DIR>
__device__ __forceinline__ int comp_of(const int m) {
  return m == 0 ? DIR : (m <= DIR ? m - 1 : m);
}

// issue the staging of item (s0n, m) — a whole chunk if enough cells are left, else one cell
template <int RECON, int DIR, int GEO>
__device__ __forceinline__ void issue_item(const GeoT<GEO> &G, const ColumnCtx &c, const int buf,
                                           const int s0n, const int m) {
  const int64_t sd = (DIR == 1) ? G.sj() : G.sk();
  const int n = comp_of<DIR>(m);
  const int64_t off = n * G.sc() + (int64_t)(G.is(DIR) + s0n) * sd;
  const uint32_t st = static_cast<uint32_t>(__cvta_generic_to_shared(
      c.stage + (size_t)buf * kStageDoubles * kThreads));
  constexpr int lo = StencilRows<RECON>::kLo, extra = StencilRows<RECON>::kExtra;
  if (c.nd + 1 - s0n >= kChunk) {
    if (s0n + lo < 0 || s0n + lo + kChunk + extra > c.nd)
      stage_issue<RECON, kChunk, true>(st, c.ub + off, c.ob + off, sd, s0n, c.nd, c.dlo, c.dhi,
                                       c.keep, c.stream);
    else
      stage_issue<RECON, kChunk, false>(st, c.ub + off, c.ob + off, sd, s0n, c.nd, 0, 0, c.keep,
                                        c.stream);
  } else {
    stage_issue<RECON, 1, true>(st, c.ub + off, c.ob + off, sd, s0n, c.nd, c.dlo, c.dhi, c.keep,
                                c.stream);
  }
  cp_async_commit();
}

template <int RECON, int DIR, bool LAST, int GEO, int CH>
__device__ __forceinline__ void run_chunk(const Args &a, const GeoT<GEO> &G, const ColumnCtx &c,
                                          const int s0, int &buf, double &rate) {
  const int64_t sd = (DIR == 1) ? G.sj() : G.sk();
  double P[CH], Q[CH];
  const int64_t row0 = (int64_t)(G.is(DIR) + s0) * sd;
#pragma unroll 1
  for (int m = 0; m < c.nc; ++m) {
    // next item in flight: component m+1 of this chunk, or the first of the next one
    const bool more = m + 1 < c.nc || s0 + CH <= c.nd;
    if (more) {
      if (m + 1 < c.nc)
        issue_item<RECON, DIR, GEO>(G, c, buf ^ 1, s0, m + 1);
      else
        issue_item<RECON, DIR, GEO>(G, c, buf ^ 1, s0 + CH, 0);
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    const int n = comp_of<DIR>(m);
    const double *st = c.stage + (size_t)buf * kStageDoubles * kThreads;
    double L = c.sL[n * kThreads], F = c.sF[n * kThreads];
    const double cd = n < 3 ? 0.5 * c.cdir : c.cdir;
    double *po = c.ob + n * G.sc() + row0;
    if (m == 0)
      comp_chunk<RECON, true, CH>(st, po, sd, s0, cd, L, F, P, Q, c.stream);
    else
      comp_chunk<RECON, false, CH>(st, po, sd, s0, cd, L, F, P, Q, c.stream);
    c.sL[n * kThreads] = L;
    c.sF[n * kThreads] = F;
    buf ^= 1;
  }
  if (LAST) {
    const int kk0 = max(s0, 1) - 1, kk1 = s0 + CH - 2;
    if (kk0 <= kk1)
      finish_chunk<DIR, GEO>(a, G, c.b, c.ob, c.col0, kk0, kk1, c.idx0, c.idx1, c.idx2, rate);
  }
}

// index of a face neighbour in the 27-entry table: (ox+1) + 3 (oy+1) + 9 (oz+1)
__device__ __forceinline__ constexpr int face_slot(const int dir, const int side) {
  return 13 + (dir == 0 ? 1 : (dir == 1 ? 3 : 9)) * (side ? 1 : -1);
}

template <int RECON, int DIR, bool LAST, int GEO>
__global__ void __launch_bounds__(kThreads, PB2_CHUNK_MINB) sweep_chunk_kernel(const Args a) {
  const GeoT<GEO> G(a.g);
  const int OD = (DIR == 1) ? 2 : 1;
  const int ncol = G.nx(OD) * G.nx(0);
  const int ctas_per_block = (ncol + kThreads - 1) / kThreads;
  const int bi = blockIdx.x / ctas_per_block;
  const int b = a.block_ids ? a.block_ids[bi] : bi;
  const int col = (blockIdx.x % ctas_per_block) * kThreads + threadIdx.x;
  extern __shared__ double smem[];
  double rate = 0.0;
  if (col < ncol) {
    ColumnCtx c;
    c.b = b;
    c.nc = a.g.ncomp;
    const int ci = col % G.nx(0), co = col / G.nx(0);
    const int64_t so = (DIR == 1) ? G.sk() : G.sj();
    const int64_t sd = (DIR == 1) ? G.sj() : G.sk();
    c.nd = G.nx(DIR);
    c.col0 = (int64_t)(G.is(OD) + co) * so + (G.is(0) + ci); // inside a component
    const int64_t sb = a.g.sb;
    c.ub = a.u + (int64_t)b * sb + c.col0;
    c.ob = a.out + (int64_t)b * sb + c.col0;
    c.idx0 = 1.0 / a.dx[3 * b];
    c.idx1 = 1.0 / a.dx[3 * b + 1];
    c.idx2 = 1.0 / a.dx[3 * b + 2];
    c.cdir = -a.bdt * (DIR == 1 ? c.idx1 : c.idx2);
    c.sL = smem + threadIdx.x;
    c.sF = c.sL + (size_t)c.nc * kThreads;
    c.stage = c.sF + (size_t)c.nc * kThreads;
    c.dlo = 0;
    c.dhi = 0;
    c.keep = l2_policy_evict_last();
    c.stream = l2_policy_evict_first();
    if (a.nbr) {
      // rows beyond the block's ends: the same columns of the face neighbour's interior
      const int nlo = a.nbr[b * 27 + face_slot(DIR, 0)], nhi = a.nbr[b * 27 + face_slot(DIR, 1)];
      if (nlo >= 0) c.dlo = (long long)(nlo - b) * sb + (long long)c.nd * sd;
      if (nhi >= 0) c.dhi = (long long)(nhi - b) * sb - (long long)c.nd * sd;
    }
    for (int n = 0; n < c.nc; ++n) {
      c.sL[n * kThreads] = 0.0;
      c.sF[n * kThreads] = 0.0;
    }
    // cells s = -1 .. nd: whole chunks, then single cells; the first item's rows start now
    int s0 = -1, buf = 0;
    issue_item<RECON, DIR, GEO>(G, c, 0, s0, 0);
#pragma unroll 1
    for (; c.nd + 1 - s0 >= kChunk; s0 += kChunk)
      run_chunk<RECON, DIR, LAST, GEO, kChunk>(a, G, c, s0, buf, rate);
#pragma unroll 1
    for (; s0 <= c.nd; ++s0) run_chunk<RECON, DIR, LAST, GEO, 1>(a, G, c, s0, buf, rate);
  }
  if (LAST) reduce_dt(a, rate);
  if (LAST && a.progress != nullptr && bi < a.progress_blocks) {
    // this block's results are in memory: tell whoever waits on another stream
    __syncthreads();
    if (threadIdx.x == 0) {
      __threadfence();
      atomicAdd(a.progress, 1);
    }
  }
}

inline size_t chunk_smem_bytes(int ncomp) {
  return sizeof(double) * kThreads * (2 * ncomp + 2 * kStageDoubles);
}

} // namespace PB2_SWEEP_NS
} // namespace pb2
