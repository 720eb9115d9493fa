// burgers_xsweep.cuh — x sweep of the FAST burgers stage (included by burgers_sweep.cu).
//
// The sweep direction is the contiguous one, so lanes run along x.  A lane owns a PAIR of
// neighbouring cells (2p-1, 2p), p = 0 .. nx1/2, of one row: (row, pair) items of a group of
// rows are flattened over the lanes of a warp (32 rows x 17 pairs = 17 full passes for the
// benchmark's block, no idle lanes).  Owning two cells lets a lane share first differences and
// the curvature parts of the smoothness indicators between its two reconstructions
// (weno_fast.cuh: WENO5Z_diff), and halves the traffic between lanes: per component one
// rotate-shuffle hands the pair's last left state and last face flux to the next lane (lane 31
// offers what it held in the previous pass, parked in shared memory).
// Per pass the lane first reconstructs the x velocity and turns the HLL wave speeds of its two
// faces into coefficients P, Q (F = P qL + Q qR for every component, burgers_march.cuh:
// face_pq); the components then run one after the other with the next component's stencil
// loads in flight.  Cells 2p-2 and 2p-1 are complete once faces 2p-1 and 2p are known:
//   out = beta u + (1 - beta) base - (beta dt / dx1) (F(i+1) - F(i))      (update.hpp:43-137)
// With Args::nbr the stencil values beyond the row's ends come from the interior cells of the
// x neighbours instead of the block's own ghost cells (see burgers_march.cuh, "Ghost rows").
#pragma once

namespace pb2 {
namespace PB2_SWEEP_NS {

#ifndef PB2_XPAIR_MINB
#define PB2_XPAIR_MINB 4
#endif
constexpr int kXRows = 32; // rows per warp task

template <int RECON>
struct PairStencil;

template <>
struct PairStencil<PB2_RECON_WENO5> {
  static constexpr int kLo = -2, kN = 6; // loads cells c0-2 .. c0+3 (c0 = first cell of the pair)
  double q[6];
  __device__ __forceinline__ void recon(double &l0, double &r0, double &l1, double &r1) const {
    const double da = q[1] - q[0], db = q[2] - q[1], dc = q[3] - q[2], dd = q[4] - q[3],
                 de = q[5] - q[4];
    const double A1 = weno_curv(da, db), A2 = weno_curv(db, dc), A3 = weno_curv(dc, dd),
                 A4 = weno_curv(dd, de);
    WENO5Z_diff(da, db, dc, dd, A1, A2, A3, q[2], l0, r0);
    WENO5Z_diff(db, dc, dd, de, A2, A3, A4, q[3], l1, r1);
  }
  __device__ __forceinline__ double u0() const { return q[1]; } // cell c0-1
  __device__ __forceinline__ double u1() const { return q[2]; } // cell c0
};

template <>
struct PairStencil<PB2_RECON_LINEAR> {
  static constexpr int kLo = -1, kN = 4;
  double q[4];
  __device__ __forceinline__ void recon(double &l0, double &r0, double &l1, double &r1) const {
    const double da = q[1] - q[0], db = q[2] - q[1], dc = q[3] - q[2];
    Linear_diff(da, db, q[1], l0, r0);
    Linear_diff(db, dc, q[2], l1, r1);
  }
  __device__ __forceinline__ double u0() const { return q[0]; }
  __device__ __forceinline__ double u1() const { return q[1]; }
};

// Element offsets that move a load from the block's own cells to the x neighbour's when the
// cells it fetches lie beyond the row's ends: four load groups of a pair whose first cell is c0
// (cells c0-2 | c0-1, c0 | c0+1, c0+2 | c0+3; c0 is odd, so a group never straddles an end).
struct PairShift {
  long long o[4];
};
template <int RECON>
__device__ __forceinline__ PairShift pair_shift(const int c0, const int nx, const long long dlo,
                                                const long long dhi) {
  PairShift s;
  if (RECON == PB2_RECON_WENO5) {
    s.o[0] = c0 - 2 < 0 ? dlo : 0;
    s.o[1] = c0 < 0 ? dlo : 0;      // cells c0-1, c0 (c0 = -1: both below the row)
    s.o[2] = c0 + 1 >= nx ? dhi : 0; // cells c0+1, c0+2
    s.o[3] = c0 + 3 >= nx ? dhi : 0;
  } else { // cells c0-1 | c0 | c0+1 | c0+2
    s.o[0] = c0 - 1 < 0 ? dlo : 0;
    s.o[1] = c0 < 0 ? dlo : 0;
    s.o[2] = c0 + 1 >= nx ? dhi : 0;
    s.o[3] = c0 + 2 >= nx ? dhi : 0;
  }
  return s;
}

// ---- staging: the stencil of a pair (and the pair's two `base` values) travel global ->
// per-thread shared-memory column with cp.async, kXAhead components ahead of the arithmetic
constexpr int kXStages = 3, kXAhead = kXStages - 1, kXSlots = 8;

// what a lane needs to address one (pass, component) item
struct PairItem {
  int64_t off;   // element offset of the pair's first cell inside a component
  PairShift sh;  // ... and where its four load groups really come from
  bool ldb;      // the pair completes two cells whose `base` values are needed
};

template <int RECON>
__device__ __forceinline__ void issue_pair(const uint32_t st, const double *__restrict__ ub,
                                           const double *bb, const int64_t sc, const int n,
                                           const PairItem &it) {
  const double *p = ub + n * sc + it.off;
  constexpr uint32_t K = kThreads * 8;
  if (RECON == PB2_RECON_WENO5) {
    cp_async8(st, p + it.sh.o[0] - 2);
    cp_async8(st + K, p + it.sh.o[1] - 1);
    cp_async8(st + 2 * K, p + it.sh.o[1]);
    cp_async8(st + 3 * K, p + it.sh.o[2] + 1);
    cp_async8(st + 4 * K, p + it.sh.o[2] + 2);
    cp_async8(st + 5 * K, p + it.sh.o[3] + 3);
  } else {
#pragma unroll
    for (int t = 0; t < 4; ++t) cp_async8(st + t * K, p + it.sh.o[t] - 1 + t);
  }
  if (it.ldb) {
    const double *q = bb + n * sc + it.off;
    cp_async8(st + 6 * K, q - 1);
    cp_async8(st + 7 * K, q);
  }
}

template <int RECON, bool LAST, int GEO>
__global__ void __launch_bounds__(kThreads, PB2_XPAIR_MINB) sweep_xpair_kernel(const Args a) {
  const GeoT<GEO> G(a.g);
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int warp_global = blockIdx.x * (kThreads / 32) + wid;
  const int nrows = G.nx(1) * G.nx(2);
  const int warps_per_block = (nrows + kXRows - 1) / kXRows;
  const int bi = warp_global / warps_per_block;
  double rate = 0.0;
  __shared__ double sL[kThreads / 32][kMaxComp], sF[kThreads / 32][kMaxComp];
  __shared__ double xs[kXStages][kXSlots][kThreads];
  if (bi < a.g.nblocks) { // whole warp together
    const int b = a.block_ids ? a.block_ids[bi] : bi;
    const int row0 = (warp_global % warps_per_block) * kXRows;
    const int rows = min(kXRows, nrows - row0);
    const int npair = G.nx(0) / 2 + 1;
    const int items = rows * npair;
    const int nc = a.g.ncomp;
    const int64_t sc = G.sc();
    const double *__restrict__ ub = a.u + (int64_t)b * a.g.sb;
    // `out` may alias `base` (second RK stage): each cell of base is read by the one lane that
    // then writes it, so neither pointer is restrict
    const double *bb = a.base + (int64_t)b * a.g.sb;
    double *ob = a.out + (int64_t)b * a.g.sb;
    const double idx0 = 1.0 / a.dx[3 * b];
    const double cdir = -a.bdt * idx0;
    const bool use_base = a.w2 != 0.0;
    const int src = (lane + 31) & 31;
    long long dlo = 0, dhi = 0; // to the x neighbours' interiors, or 0: own ghost cells
    if (a.nbr) {
      const int nlo = a.nbr[b * 27 + face_slot(0, 0)], nhi = a.nbr[b * 27 + face_slot(0, 1)];
      if (nlo >= 0) dlo = (long long)(nlo - b) * a.g.sb + G.nx(0);
      if (nhi >= 0) dhi = (long long)(nhi - b) * a.g.sb - G.nx(0);
    }
    if (lane < kMaxComp) {
      sL[wid][lane] = 0.0;
      sF[wid][lane] = 0.0;
    }
    __syncwarp();
    const uint32_t xs0 = static_cast<uint32_t>(__cvta_generic_to_shared(&xs[0][0][threadIdx.x]));
    constexpr uint32_t kStageBytes = kXSlots * kThreads * 8;

    // the lane's item of pass f0: which pair of which row, where its cells live
    auto item_of = [&](const int f0, PairItem &it, int &p_out, bool &upd_out) {
      const int f = f0 + lane;
      const bool live = f < items;
      const int fr = live ? f : items - 1; // inactive lanes recompute a valid item
      int r, p;
      if (GEO == 32) {
        r = fr / 17;
        p = fr - r * 17;
      } else {
        r = (int)a.dnpair.div((uint32_t)fr);
        p = fr - r * npair;
      }
      const int row = row0 + r;
      int rk, rj;
      if (GEO == 32) {
        rk = row >> 5;
        rj = row & 31;
      } else {
        rk = (int)a.dnx1.div((uint32_t)row);
        rj = row - rk * G.nx(1);
      }
      // element offset of cell c0 = 2p - 1 inside a component
      it.off = (int64_t)(G.is(2) + rk) * G.sk() + (int64_t)(G.is(1) + rj) * G.sj() +
               (G.is(0) + 2 * p - 1);
      it.sh = pair_shift<RECON>(2 * p - 1, G.nx(0), dlo, dhi);
      upd_out = live && p >= 1; // this lane completes cells 2p-2 and 2p-1
      it.ldb = use_base && upd_out;
      p_out = p;
    };

    PairItem cur, nxt;
    int p, pn;
    bool upd, updn;
    item_of(0, cur, p, upd);
    nxt = cur;
    // fill the pipeline: the first kXAhead components of the first pass
    int qs = 0; // stage of the item about to be computed
#pragma unroll
    for (int k = 0; k < kXAhead; ++k) {
      if (k < nc) issue_pair<RECON>(xs0 + k * kStageBytes, ub, bb, sc, k, cur);
      cp_async_commit();
    }
    for (int f0 = 0; f0 < items; f0 += 32) {
      const bool more = f0 + 32 < items;
      if (more) item_of(f0 + 32, nxt, pn, updn);
      double P0, Q0, P1, Q1;
      double sq0 = 0.0, sq1 = 0.0; // LAST (1-D meshes): sum of squared velocities, cells 2p-2 / 2p-1
      const int64_t off = cur.off;
#pragma unroll 1
      for (int n = 0; n < nc; ++n) {
        // item kXAhead ahead: a later component of this pass or an early one of the next
        {
          const int na = n + kXAhead;
          int sa = qs + kXAhead;
          if (sa >= kXStages) sa -= kXStages;
          if (na < nc)
            issue_pair<RECON>(xs0 + sa * kStageBytes, ub, bb, sc, na, cur);
          else if (more && na - nc < nc)
            issue_pair<RECON>(xs0 + sa * kStageBytes, ub, bb, sc, na - nc, nxt);
          cp_async_commit();
          cp_async_wait<kXAhead>();
        }
        const double *st = &xs[qs][0][threadIdx.x];
        PairStencil<RECON> q;
#pragma unroll
        for (int t = 0; t < PairStencil<RECON>::kN; ++t) q.q[t] = st[t * kThreads];
        double b0 = 0.0, b1 = 0.0;
        if (cur.ldb) {
          b0 = st[6 * kThreads];
          b1 = st[7 * kThreads];
        }
        if (++qs == kXStages) qs = 0;
        double l0, r0, l1, r1;
        q.recon(l0, r0, l1, r1);
        // left state of cell 2p-2 from the previous lane (lane 31's value of the previous pass
        // for lane 0)
        const double cL = sL[wid][n], cF = sF[wid][n];
        const double Lp = __shfl_sync(full, lane == 31 ? cL : l1, src);
        if (n == 0) {
          face_pq(Lp, r0, P0, Q0);
          face_pq(l0, r1, P1, Q1);
        }
        const double fa = fma(P0, Lp, Q0 * r0); // face 2p-1
        const double fb = fma(P1, l0, Q1 * r1); // face 2p
        const double Fp = __shfl_sync(full, lane == 31 ? cF : fb, src); // face 2p-2
        __syncwarp();
        if (lane == 31) {
          sL[wid][n] = l1;
          sF[wid][n] = fb;
        }
        __syncwarp();
        if (upd) {
          const double cd = n < 3 ? 0.5 * cdir : cdir;
          const double o0 = fma(cd, fa - Fp, fma(a.w2, b0, a.beta * q.u0()));
          const double o1 = fma(cd, fb - fa, fma(a.w2, b1, a.beta * q.u1()));
          if (GEO == 32) {
            *reinterpret_cast<double2 *>(ob + n * sc + off - 1) = make_double2(o0, o1);
          } else {
            ob[n * sc + off - 1] = o0;
            ob[n * sc + off] = o1;
          }
          if (LAST) { // a 1-D mesh: CalculateDerived and the time-step rate right here
            if (n < 3) {
              sq0 = fma(o0, o0, sq0);
              sq1 = fma(o1, o1, sq1);
              if (n == 0) rate = max_std(rate, max_std(fabs(o0), fabs(o1)) * idx0);
            } else if (n == 3 && a.derived) {
              a.derived[(int64_t)b * sc + off - 1] = 0.5 * o0 * sq0;
              a.derived[(int64_t)b * sc + off] = 0.5 * o1 * sq1;
            }
          }
        }
      }
      cur = nxt;
      p = pn;
      upd = updn;
    }
    cp_async_wait<0>();
    (void)p;
  }
  if (LAST) reduce_dt(a, rate);
}

} // namespace PB2_SWEEP_NS
} // namespace pb2
