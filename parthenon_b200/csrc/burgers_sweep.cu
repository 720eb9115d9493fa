// burgers_sweep.cu — the FAST burgers stage: three direction sweeps that never store a flux.
//
// Replaces, for PB2_MATH_FAST, the whole dense part of a stage of the reference's task list
// (benchmarks/burgers/burgers_driver.cpp:92-127): CalculateFluxes (burgers_package.cpp:202-404)
// + FluxDivergence (update.cpp:63-86) + AverageIndependentData + UpdateIndependentData
// (update.hpp:71-137) + CalculateDerived (:143-167) + EstimateTimestepMesh (:170-200).
//
// The reference writes 6 x 11 reconstructed fields, reads them back, writes 3 x 11 fluxes,
// reads them back and then runs three more full-array passes.  Here each direction is ONE
// kernel in which the flux of a face lives only in registers:
//   x sweep:  out  = beta*u + (1-beta)*base - (beta*dt/dx1) * (F1(i+1) - F1(i))
//   y sweep:  out -= (beta*dt/dx2) * (F2(j+1) - F2(j))
//   z sweep:  out -= (beta*dt/dx3) * (F3(k+1) - F3(k));  derived;  min dt
// (algebraically update.hpp:43-58 with A_d / V = 1 / dx_d on a uniform Cartesian block).
// The kernels are bound by the FP64 pipe (33 WENO5-Z reconstructions per zone and stage),
// so the arithmetic is reorganised to issue fewer FP64 instructions than the reference's
// expression tree (weno_fast.cuh: difference form, 4 shared reciprocals instead of 8
// divisions), the HLL denominators of a face are inverted once for all 11 components, and
// stencil values are software-prefetched one component ahead.  Results stay
// within 1e-12 relative of the reference (tests/test_burgers_sim_gpu.py); the bit-exact
// arithmetic lives in burgers_strict.cu.  Valid for |q| < ~1e40 (products of three
// smoothness indicators must not overflow).
#include <cfloat>
#include <cstdlib>

#include "common.cuh"
#include "weno_fast.cuh"

#ifndef PB2_SWEEP_MINB
#define PB2_SWEEP_MINB 4
#endif
#ifndef PB2_SWEEP_UNR
#define PB2_SWEEP_UNR 2
#endif
#ifndef PB2_SWEEP_NS
#define PB2_SWEEP_NS sweep
#endif
#define PB2_CAT_(a, b) a##b
#define PB2_CAT(a, b) PB2_CAT_(a, b)

namespace pb2 {
namespace PB2_SWEEP_NS {

constexpr int kThreads = 128;
constexpr int kMaxComp = 16;
// component count the kernels are specialised for (the benchmark's 3 velocities + 8 scalars):
// with a compile-time count the scalar loops unroll and every ring / carry address becomes an
// immediate offset (~10 % fewer issued instructions; measured 5 % on the march kernels)
constexpr int kSpecialNC = 11;

struct Geom {
  int nblocks, ncomp, ndim;
  int nx[3], is[3], n[3];
  int64_t sj, sk, sc, sb;
};

struct Args {
  Geom g;
  const double *u, *base;
  double *out;
  double *derived;           // LAST sweep only, or null
  unsigned long long *dtmin; // LAST sweep only, or null
  const double *dx;          // [nblocks][3]
  const int *block_ids;      // launch block -> block of the batch, or null (identity)
  const int *nbr;            // [nblocks][27] same-device neighbour blocks (-1: none) whose
                             // interiors stand in for this block's ghost rows, or null
  int *progress;             // LAST sweep: += 1 per finished CTA of the first progress_blocks
  int progress_blocks;       // launched blocks, or null
  double beta, w2, bdt;      // w2 = 1 - beta, bdt = beta * dt
  FastDiv dncell, dnx1;      // x sweep: division by (nx1 + 2) and by nx2 as multiply-shift
  FastDiv dnpair;            // paired x sweep: division by nx1 / 2 + 1
};

using fastmath::Linear;
using fastmath::max_std;
using fastmath::min_std;
using fastmath::rcp_fast;
using fastmath::WENO5Z;

// pull the line holding p into L1 ahead of the pass / march step that will read it
__device__ __forceinline__ void prefetch_l1(const void *p) {
  asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
}

__device__ __forceinline__ void prefetch_l2(const void *p) {
  asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}

// per-face HLL coefficients shared by all components (burgers_package.hpp:31-43 and
// burgers_package.cpp:326-333):  F = A*qL - B*qR + C*(qR - qL), velocities carry an extra 1/2
struct FaceCoef {
  double A, B, C;
};
__device__ __forceinline__ FaceCoef face_coef(const double upl, const double upr) {
  const double sl = min_std(min_std(upl, upr), 0.0);
  const double sr = max_std(max_std(upl, upr), 0.0);
  const double slsr = sl * sr;
  const double inv = rcp_fast(sr - sl + (slsr == 0.0 ? 1.0 : 0.0));
  FaceCoef f;
  f.A = sr * upl * inv;
  f.B = sl * upr * inv;
  f.C = slsr * inv;
  return f;
}
__device__ __forceinline__ double face_flux(const FaceCoef &f, const double ql, const double qr) {
  return fma(f.A, ql, fma(-f.B, qr, f.C * (qr - ql)));
}

template <int RECON>
__device__ __forceinline__ void load_stencil(const double *__restrict__ p, const int64_t sd,
                                             double q[5]) {
  if (RECON == PB2_RECON_WENO5) {
    q[0] = __ldg(p - 2 * sd);
    q[4] = __ldg(p + 2 * sd);
  }
  q[1] = __ldg(p - sd);
  q[2] = __ldg(p);
  q[3] = __ldg(p + sd);
}
template <int RECON>
__device__ __forceinline__ void recon(const double q[5], double &ql, double &qr) {
  if (RECON == PB2_RECON_WENO5)
    WENO5Z(q[0], q[1], q[2], q[3], q[4], ql, qr);
  else
    Linear(q[1], q[2], q[3], ql, qr);
}

__device__ __forceinline__ void finish_cell(const Args &a, const int b, const int64_t cell,
                                            const double v0, const double v1, const double v2,
                                            const double v3, const double idx0,
                                            const double idx1, const double idx2,
                                            double &inv) {
  if (a.derived)
    a.derived[(int64_t)b * a.g.sc + cell] = 0.5 * v3 * (v0 * v0 + v1 * v1 + v2 * v2);
  const double rate = fabs(v0) * idx0 + (a.g.ndim > 1 ? fabs(v1) * idx1 : 0.0) +
                      (a.g.ndim > 2 ? fabs(v2) * idx2 : 0.0);
  inv = max_std(inv, rate); // min of 1/rate == 1 / max(rate)
}

__device__ __forceinline__ void reduce_dt(const Args &a, double rate) {
  if (!a.dtmin) return;
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) rate = max_std(rate, __shfl_xor_sync(0xffffffffu, rate, s));
  if ((threadIdx.x & 31) == 0) {
    const double inv = 1.0 / rate;
    atomicMin(a.dtmin, static_cast<unsigned long long>(__double_as_longlong(inv)));
  }
}

// ---- y / z sweeps: a thread owns one (i, other) column and marches along DIR ---------------
// The five stencil rows of every component live in a per-thread ring in shared memory
// (ring[comp][slot][thread]: thread-private columns, so no barriers and no bank conflicts).
// Each march step reads ONE new row per component from global memory — issued before the
// reconstruction whose oldest row it will overwrite, so HBM/L2 latency hides behind ~300 FP64
// instructions — and takes its 5-point stencil from the ring.  (With plain global loads the
// 11 x 5 rows of 4 resident CTAs overflow L1: 10 % hit rate, every stencil load an L2 trip.)
constexpr int kRing = 5;

__host__ __device__ inline size_t march_smem_bytes(int ncomp) {
  return sizeof(double) * kThreads * (static_cast<size_t>(ncomp) * kRing + 2 * (ncomp - 3));
}

template <int RECON, int DIR, bool LAST, int NC>
__global__ void __launch_bounds__(kThreads, 3) sweep_march_kernel(const Args a) {
  constexpr int kUnr = NC ? (NC - 3 + 1) / 2 : 1; // trips of the two-scalar loop
  const Geom &g = a.g;
  const int ncol_other = (DIR == 1) ? g.nx[2] : g.nx[1];
  const int ncol = ncol_other * g.nx[0];
  const int ctas_per_block = (ncol + kThreads - 1) / kThreads;
  const int bi = blockIdx.x / ctas_per_block;
  const int b = a.block_ids ? a.block_ids[bi] : bi;
  const int col = (blockIdx.x % ctas_per_block) * kThreads + threadIdx.x;
  const int nc = NC ? NC : g.ncomp;
  extern __shared__ double smem[];
  double *ring = smem + threadIdx.x;                      // + (n * kRing + slot) * kThreads
  double *sL = smem + (size_t)nc * kRing * kThreads + threadIdx.x; // + (n - 3) * kThreads
  double *sF = sL + (size_t)(nc - 3) * kThreads;
  double rate = 0.0;
  if (col < ncol) {
    const int i = g.is[0] + col % g.nx[0];
    const int o = col / g.nx[0];
    const int64_t sd = (DIR == 1) ? g.sj : g.sk;
    const int64_t so = (DIR == 1) ? g.sk : g.sj;
    const int os = (DIR == 1) ? g.is[2] : g.is[1];
    const int ds = g.is[DIR], nd = g.nx[DIR];
    const int64_t col0 = (int64_t)(os + o) * so + i; // offset inside one component
    const double *__restrict__ ub = a.u + (int64_t)b * g.sb + col0;
    double *__restrict__ ob = a.out + (int64_t)b * g.sb + col0;
    const double idx0 = 1.0 / a.dx[3 * b], idx1 = 1.0 / a.dx[3 * b + 1],
                 idx2 = 1.0 / a.dx[3 * b + 2];
    const double cdir = -a.bdt * (DIR == 1 ? idx1 : idx2);

    // fill the ring with rows ds-3 .. ds+1 (the stencil of cell ds-1); slot t <-> row ds-3+t
    for (int n = 0; n < nc; ++n) {
#pragma unroll
      for (int t = 0; t < kRing; ++t)
        ring[(n * kRing + t) * kThreads] =
            __ldg(ub + n * g.sc + (int64_t)max(ds - 3 + t, 0) * sd); // (linear: row -1 unused)
    }
    double Lc[3] = {0, 0, 0}, Fc[3] = {0, 0, 0};
    for (int n = 3; n < nc; ++n) {
      sL[(n - 3) * kThreads] = 0.0;
      sF[(n - 3) * kThreads] = 0.0;
    }
    int s0 = 0; // slot holding row s-2
    // cells ds-1 .. ds+nd, faces ds .. ds+nd; cell s-1 is complete once face s is known
    for (int s = -1; s <= nd; ++s) {
      const int64_t off = (int64_t)(ds + s) * sd;
      const bool upd = s >= 1, more = s < nd;
      int sl[kRing]; // ring slots of rows s-2 .. s+2, in units of kThreads doubles
#pragma unroll
      for (int t = 0; t < kRing; ++t) sl[t] = ((s0 + t) % kRing) * kThreads;
      double *po = ob + off - sd;              // cell s-1, component 0
      // row s+3, component 0 (clamped: with nghost 2 the linear stencil never reads it)
      const double *pn = ub + (int64_t)min(ds + s + 3, g.n[DIR] - 1) * sd;
      if (more) {
        // next step's lines start their trip from HBM now: its new stencil row into L2 (L1 is
        // mostly carved out as shared memory here), the `out` cell it completes into L1
        const double *pf = ub + (int64_t)min(ds + s + 4, g.n[DIR] - 1) * sd;
        for (int n = 0; n < nc; ++n) prefetch_l2(pf + n * g.sc);
        if (upd)
          for (int n = 0; n < nc; ++n) prefetch_l1(po + sd + n * g.sc);
      }

      double ql[3], qr[3], old[3];
#pragma unroll
      for (int n = 0; n < 3; ++n) {
        const double nw = more ? __ldg(pn + n * g.sc) : 0.0;
        old[n] = upd ? po[n * g.sc] : 0.0;
        double q[5];
        const double *rg = ring + n * kRing * kThreads;
#pragma unroll
        for (int t = 0; t < kRing; ++t) q[t] = rg[sl[t]];
        recon<RECON>(q, ql[n], qr[n]);
        ring[n * kRing * kThreads + sl[0]] = nw; // row s-2 is dead: it becomes row s+3
      }
      const FaceCoef fc = face_coef(Lc[DIR], qr[DIR]);
      double v[4] = {0, 0, 0, 0};
#pragma unroll
      for (int n = 0; n < 3; ++n) {
        const double f = 0.5 * face_flux(fc, Lc[n], qr[n]);
        if (upd) {
          v[n] = fma(cdir, f - Fc[n], old[n]);
          po[n * g.sc] = v[n];
        }
        Fc[n] = f;
        Lc[n] = ql[n];
      }
      // two scalars per trip: their reconstructions are independent instruction streams
#pragma unroll kUnr
      for (int n = 3; n < nc; n += 2) {
        const bool two = n + 1 < nc;
        const int n2 = two ? n + 1 : n;
        const double nw = more ? __ldg(pn + n * g.sc) : 0.0;
        const double nw2 = more ? __ldg(pn + n2 * g.sc) : 0.0;
        const double old1 = upd ? po[n * g.sc] : 0.0;
        const double old2 = upd ? po[n2 * g.sc] : 0.0;
        double q[5], q2[5];
        double *rg = ring + n * kRing * kThreads, *rg2 = ring + n2 * kRing * kThreads;
#pragma unroll
        for (int t = 0; t < kRing; ++t) {
          q[t] = rg[sl[t]];
          q2[t] = rg2[sl[t]];
        }
        double l, r, l2, r2;
        recon<RECON>(q, l, r);
        recon<RECON>(q2, l2, r2);
        double *pL = sL + (n - 3) * kThreads, *pF = sF + (n - 3) * kThreads;
        double *pL2 = sL + (n2 - 3) * kThreads, *pF2 = sF + (n2 - 3) * kThreads;
        const double f = face_flux(fc, *pL, r);
        const double f2 = face_flux(fc, *pL2, r2);
        if (upd) {
          const double val = fma(cdir, f - *pF, old1);
          po[n * g.sc] = val;
          if (n == 3) v[3] = val;
          if (two) po[n2 * g.sc] = fma(cdir, f2 - *pF2, old2);
        }
        *pF = f;
        *pL = l;
        rg[sl[0]] = nw;
        if (two) {
          *pF2 = f2;
          *pL2 = l2;
          rg2[sl[0]] = nw2;
        }
      }
      if (LAST && upd) finish_cell(a, b, col0 + off - sd, v[0], v[1], v[2], v[3], idx0, idx1, idx2, rate);
      s0 = (s0 + 1) % kRing;
    }
  }
  if (LAST) reduce_dt(a, rate);
}

} // namespace PB2_SWEEP_NS
} // namespace pb2
#include "burgers_march.cuh"
#include "burgers_xsweep.cuh"
namespace pb2 {
namespace PB2_SWEEP_NS {

// ---- x sweep: (row, cell) items flattened over the lanes of a warp --------------------------
constexpr int kRowsPerWarp = 16;

template <int RECON, bool LAST, int NC>
__global__ void __launch_bounds__(kThreads, PB2_SWEEP_MINB) sweep_x_kernel(const Args a) {
  constexpr int kUnr = NC ? (NC - 3 + 1) / 2 : 1;
  const Geom &g = a.g;
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int warp_global = blockIdx.x * (kThreads / 32) + wid;
  const int nrows = g.nx[1] * g.nx[2];
  const int warps_per_block = (nrows + kRowsPerWarp - 1) / kRowsPerWarp;
  const int bi = warp_global / warps_per_block;
  double rate = 0.0;
  __shared__ double sL[kThreads / 32][kMaxComp], sF[kThreads / 32][kMaxComp];
  if (bi < g.nblocks) { // whole warp together
    const int b = a.block_ids ? a.block_ids[bi] : bi;
    const int row0 = (warp_global % warps_per_block) * kRowsPerWarp;
    const int rows = min(kRowsPerWarp, nrows - row0);
    const int ncell = g.nx[0] + 2;
    const int items = rows * ncell;
    const int nc = NC ? NC : g.ncomp;
    const double *__restrict__ ub = a.u + (int64_t)b * g.sb;
    // `out` may alias `base` (second RK stage: base <- 0.5*"1" + 0.5*base + ...): each cell of
    // base is read by the one thread that then writes it, so neither pointer is restrict
    const double *bb = a.base + (int64_t)b * g.sb;
    double *ob = a.out + (int64_t)b * g.sb;
    const double idx0 = 1.0 / a.dx[3 * b], idx1 = 1.0 / a.dx[3 * b + 1],
                 idx2 = 1.0 / a.dx[3 * b + 2];
    const double cdir = -a.bdt * idx0;
    const bool use_base = a.w2 != 0.0;
    const int src = (lane + 31) & 31;

    // values of lane 31 in the previous pass, handed to lane 0 of this pass
    double pL[3] = {0, 0, 0}, pF[3] = {0, 0, 0};
    if (lane < kMaxComp) {
      sL[wid][lane] = 0.0;
      sF[wid][lane] = 0.0;
    }
    __syncwarp();
    for (int f0 = 0; f0 < items; f0 += 32) {
      const int f = f0 + lane;
      const bool cell = f < items;
      const int fr = cell ? f : items - 1; // inactive lanes recompute a valid cell
      const int r = (int)a.dncell.div((uint32_t)fr), c = fr - r * ncell;
      const int row = row0 + r;
      const int rk = (int)a.dnx1.div((uint32_t)row);
      const int k = g.is[2] + rk, j = g.is[1] + (row - rk * g.nx[1]);
      const int i = g.is[0] - 1 + c;
      const int64_t off = (int64_t)k * g.sk + (int64_t)j * g.sj + i;
      const bool upd = cell && c >= 2; // this lane completes cell i-1
      double q[5], qn[5], bv[3] = {0, 0, 0}, bvc = 0.0, bvn = 0.0;
      double ql[3], qr[3], um[3];
      const bool ldb = use_base && upd; // base(i-1) is fetched together with the stencil
      if (f0 + 32 < items) {
        // lines of the NEXT pass (32 items further along the flattened rows)
        const int f2 = min(f + 32, items - 1);
        const int r2 = (int)a.dncell.div((uint32_t)f2), c2 = f2 - r2 * ncell, row2 = row0 + r2;
        const int rk2 = (int)a.dnx1.div((uint32_t)row2);
        const int64_t off2 = (int64_t)(g.is[2] + rk2) * g.sk +
                             (int64_t)(g.is[1] + (row2 - rk2 * g.nx[1])) * g.sj + (g.is[0] - 1 + c2);
        for (int n = 0; n < nc; ++n) prefetch_l1(ub + n * g.sc + off2);
        if (use_base)
          for (int n = 0; n < nc; ++n) prefetch_l1(bb + n * g.sc + off2);
      }
      load_stencil<RECON>(ub + off, 1, q);
#pragma unroll
      for (int n = 0; n < 3; ++n) {
        load_stencil<RECON>(ub + (n + 1) * g.sc + off, 1, qn);
        if (ldb) bv[n] = bb[n * g.sc + off - 1];
        um[n] = q[1];
        recon<RECON>(q, ql[n], qr[n]);
#pragma unroll
        for (int t = 0; t < 5; ++t) q[t] = qn[t];
      }
      double Lp[3];
#pragma unroll
      for (int n = 0; n < 3; ++n) {
        Lp[n] = __shfl_sync(full, lane == 31 ? pL[n] : ql[n], src);
        pL[n] = ql[n];
      }
      const FaceCoef fc = face_coef(Lp[0], qr[0]);
      double v[4] = {0, 0, 0, 0};
#pragma unroll
      for (int n = 0; n < 3; ++n) {
        const double fl = 0.5 * face_flux(fc, Lp[n], qr[n]);
        const double fprev = __shfl_sync(full, lane == 31 ? pF[n] : fl, src);
        pF[n] = fl;
        if (upd) {
          const double avg = fma(a.w2, bv[n], a.beta * um[n]);
          v[n] = fma(cdir, fl - fprev, avg);
          ob[n * g.sc + off - 1] = v[n];
        }
      }
      if (ldb) bvc = bb[3 * g.sc + off - 1];
#if PB2_SWEEP_UNR == 2
      (void)bvn;
#pragma unroll kUnr
      for (int n = 3; n < nc; n += 2) {
        const bool two = n + 1 < nc;
        const int n2 = two ? n + 1 : n;
        load_stencil<RECON>(ub + n2 * g.sc + off, 1, qn);
        const double bv2 = ldb ? bb[n2 * g.sc + off - 1] : 0.0;
        double l, rr, l2, rr2;
        recon<RECON>(q, l, rr);
        recon<RECON>(qn, l2, rr2);
        const double cL = sL[wid][n], cF = sF[wid][n], cL2 = sL[wid][n2], cF2 = sF[wid][n2];
        const double Lq = __shfl_sync(full, lane == 31 ? cL : l, src);
        const double Lq2 = __shfl_sync(full, lane == 31 ? cL2 : l2, src);
        const double fl = face_flux(fc, Lq, rr), fl2 = face_flux(fc, Lq2, rr2);
        const double fprev = __shfl_sync(full, lane == 31 ? cF : fl, src);
        const double fprev2 = __shfl_sync(full, lane == 31 ? cF2 : fl2, src);
        __syncwarp();
        if (lane == 31) {
          sL[wid][n] = l;
          sF[wid][n] = fl;
          if (two) {
            sL[wid][n2] = l2;
            sF[wid][n2] = fl2;
          }
        }
        __syncwarp();
        if (upd) {
          const double val = fma(cdir, fl - fprev, fma(a.w2, bvc, a.beta * q[1]));
          ob[n * g.sc + off - 1] = val;
          if (n == 3) v[3] = val;
          if (two)
            ob[n2 * g.sc + off - 1] = fma(cdir, fl2 - fprev2, fma(a.w2, bv2, a.beta * qn[1]));
        }
        if (n + 2 < nc) {
          load_stencil<RECON>(ub + (n + 2) * g.sc + off, 1, q);
          if (ldb) bvc = bb[(n + 2) * g.sc + off - 1];
        }
      }
#else
#pragma unroll 1
      for (int n = 3; n < nc; ++n) {
        if (n + 1 < nc) {
          load_stencil<RECON>(ub + (n + 1) * g.sc + off, 1, qn);
          if (ldb) bvn = bb[(n + 1) * g.sc + off - 1];
        }
        double l, rr;
        recon<RECON>(q, l, rr);
        // lane 31 offers what it held in the previous pass (parked in shared memory)
        const double carryL = sL[wid][n], carryF = sF[wid][n];
        const double Lq = __shfl_sync(full, lane == 31 ? carryL : l, src);
        const double fl = face_flux(fc, Lq, rr);
        const double fprev = __shfl_sync(full, lane == 31 ? carryF : fl, src);
        __syncwarp();
        if (lane == 31) {
          sL[wid][n] = l;
          sF[wid][n] = fl;
        }
        __syncwarp();
        if (upd) {
          const double avg = fma(a.w2, bvc, a.beta * q[1]);
          const double val = fma(cdir, fl - fprev, avg);
          ob[n * g.sc + off - 1] = val;
          if (n == 3) v[3] = val;
        }
        bvc = bvn;
#pragma unroll
        for (int t = 0; t < 5; ++t) q[t] = qn[t];
      }
#endif
      if (LAST && upd) finish_cell(a, b, off - 1, v[0], v[1], v[2], v[3], idx0, idx1, idx2, rate);
    }
  }
  if (LAST) reduce_dt(a, rate);
}

template <int RECON, int NC>
int launch_nc(const pb2_burgers_args *args, cudaStream_t st) {
  Args a;
  const pb2_pack_geom &pg = args->geom;
  Geom &g = a.g;
  g.nblocks = pg.nblocks;
  g.ncomp = pg.ncomp;
  g.ndim = pg.ndim;
  for (int d = 0; d < 3; ++d) {
    const bool sym = d >= pg.ndim;
    g.nx[d] = sym ? 1 : pg.nx[d];
    g.is[d] = sym ? 0 : pg.ng;
    g.n[d] = sym ? 1 : pg.nx[d] + 2 * pg.ng;
  }
  g.sj = g.n[0];
  g.sk = (int64_t)g.n[0] * g.n[1];
  g.sc = g.sk * g.n[2];
  g.sb = pg.block_stride;
  a.u = args->u;
  a.base = args->base;
  a.out = args->out;
  a.dx = pg.dx;
  a.block_ids = args->block_ids;
  if (args->block_ids) g.nblocks = args->num_block_ids;
  if (g.nblocks == 0) return PB2_OK;
  a.dncell.init(static_cast<uint32_t>(g.nx[0] + 2));
  a.dnx1.init(static_cast<uint32_t>(g.nx[1]));
  a.dnpair.init(static_cast<uint32_t>(g.nx[0] / 2 + 1));
  a.beta = args->beta;
  a.w2 = 1.0 - args->beta;
  a.bdt = args->beta * args->dt;
  a.derived = nullptr;
  a.dtmin = nullptr;
  a.progress = nullptr;
  a.progress_blocks = 0;
  // PB2_SWEEP_V1=1 selects the kernels of round 1 (one cell per lane in x, shared-memory-ring
  // march in y / z) for A/B runs
  static const bool v1 = std::getenv("PB2_SWEEP_V1") != nullptr;
  // the march kernels keep their stencil rows in dynamic shared memory (opt-in above 48 KB)
  const size_t smem = march_smem_bytes(g.ncomp);
  static bool attr_set = false;
  if (!attr_set) {
    const int maxb = 200 * 1024;
    PB2_CUDA_CHECK(cudaFuncSetAttribute(sweep_march_kernel<RECON, 1, true, NC>, cudaFuncAttributeMaxDynamicSharedMemorySize, maxb));
    PB2_CUDA_CHECK(cudaFuncSetAttribute(sweep_march_kernel<RECON, 1, false, NC>, cudaFuncAttributeMaxDynamicSharedMemorySize, maxb));
    PB2_CUDA_CHECK(cudaFuncSetAttribute(sweep_march_kernel<RECON, 2, true, NC>, cudaFuncAttributeMaxDynamicSharedMemorySize, maxb));
    attr_set = true;
  }
  auto last = [&](Args &x) {
    x.progress = v1 ? nullptr : args->progress;
    x.progress_blocks = args->progress_blocks;
    x.derived = args->derived;
    x.dtmin = reinterpret_cast<unsigned long long *>(args->dt_min);
  };
  const double zones = (double)g.nblocks * g.nx[0] * g.nx[1] * g.nx[2]; // of the launched blocks
  const auto al16 = [](const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; };
  const bool g32 = geo32(g) && al16(a.u) && al16(a.base) && al16(a.out) && g.sb % 2 == 0;
  a.nbr = v1 ? nullptr : args->nbr_direct;
  const int sweeps = args->sweeps ? args->sweeps : 7;
  PB2_REQUIRE(!v1 || sweeps == 7, "PB2_SWEEP_V1 kernels run whole stages only");
  if (!(sweeps & 1)) {
    // x1 sweep done by an earlier call
  } else if (!v1 && g.nx[0] % 2 == 0) {
    const int nrows = g.nx[1] * g.nx[2];
    const int warps = g.nblocks * ((nrows + kXRows - 1) / kXRows);
    const int wpc = kThreads / 32;
    const int ctas = (warps + wpc - 1) / wpc;
    ProfScope prof(K_SWEEP_XPAIR, st, zones);
    if (g.ndim == 1) {
      last(a);
      sweep_xpair_kernel<RECON, true, 0><<<ctas, kThreads, 0, st>>>(a);
    } else if (g32) {
      sweep_xpair_kernel<RECON, false, 32><<<ctas, kThreads, 0, st>>>(a);
    } else {
      sweep_xpair_kernel<RECON, false, 0><<<ctas, kThreads, 0, st>>>(a);
    }
    PB2_LAUNCH_CHECK();
  } else
  {
    const int nrows = g.nx[1] * g.nx[2];
    const int warps = g.nblocks * ((nrows + kRowsPerWarp - 1) / kRowsPerWarp);
    const int wpc = kThreads / 32;
    const int ctas = (warps + wpc - 1) / wpc;
    ProfScope prof(K_SWEEP_X, st, zones);
    if (g.ndim == 1) {
      last(a);
      sweep_x_kernel<RECON, true, NC><<<ctas, kThreads, 0, st>>>(a);
    } else {
      sweep_x_kernel<RECON, false, NC><<<ctas, kThreads, 0, st>>>(a);
    }
    PB2_LAUNCH_CHECK();
  }
  // y / z sweeps: chunked register march (burgers_march.cuh)
  if (!v1) {
    auto march = [&](auto kern, int dir, bool is_last) -> int {
      if (!(sweeps & (1 << dir))) return PB2_OK;
      const int ncol = g.nx[dir == 1 ? 2 : 1] * g.nx[0];
      const int ctas = g.nblocks * ((ncol + kThreads - 1) / kThreads);
      if (is_last) last(a);
      const size_t csm = chunk_smem_bytes(g.ncomp);
      PB2_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          static_cast<int>(csm)));
      ProfScope prof(dir == 1 ? K_SWEEP_CHUNK_Y : K_SWEEP_CHUNK_Z, st, zones);
      kern<<<ctas, kThreads, csm, st>>>(a);
      PB2_LAUNCH_CHECK();
      return PB2_OK;
    };
    if (g.ndim == 2) {
      return march(sweep_chunk_kernel<RECON, 1, true, 0>, 1, true);
    } else if (g.ndim == 3) {
      if (g32) {
        if (int rc = march(sweep_chunk_kernel<RECON, 1, false, 32>, 1, false)) return rc;
        return march(sweep_chunk_kernel<RECON, 2, true, 32>, 2, true);
      }
      if (int rc = march(sweep_chunk_kernel<RECON, 1, false, 0>, 1, false)) return rc;
      return march(sweep_chunk_kernel<RECON, 2, true, 0>, 2, true);
    }
    return PB2_OK;
  }
  if (g.ndim > 1) {
    const int ncol = g.nx[2] * g.nx[0];
    const int ctas = g.nblocks * ((ncol + kThreads - 1) / kThreads);
    ProfScope prof(K_SWEEP_Y, st, zones);
    if (g.ndim == 2) {
      last(a);
      sweep_march_kernel<RECON, 1, true, NC><<<ctas, kThreads, smem, st>>>(a);
    } else {
      sweep_march_kernel<RECON, 1, false, NC><<<ctas, kThreads, smem, st>>>(a);
    }
    PB2_LAUNCH_CHECK();
  }
  if (g.ndim > 2) {
    const int ncol = g.nx[1] * g.nx[0];
    const int ctas = g.nblocks * ((ncol + kThreads - 1) / kThreads);
    last(a);
    ProfScope prof(K_SWEEP_Z, st, zones);
    sweep_march_kernel<RECON, 2, true, NC><<<ctas, kThreads, smem, st>>>(a);
    PB2_LAUNCH_CHECK();
  }
  return PB2_OK;
}

template <int RECON>
int launch(const pb2_burgers_args *args, cudaStream_t st) {
  if (args->geom.ncomp == kSpecialNC) return launch_nc<RECON, kSpecialNC>(args, st);
  return launch_nc<RECON, 0>(args, st);
}

} // namespace PB2_SWEEP_NS

int PB2_CAT(burgers_stage_, PB2_SWEEP_NS)(const pb2_burgers_args *args, cudaStream_t st) {
  if (args->recon == PB2_RECON_WENO5) return PB2_SWEEP_NS::launch<PB2_RECON_WENO5>(args, st);
  return PB2_SWEEP_NS::launch<PB2_RECON_LINEAR>(args, st);
}

} // namespace pb2
