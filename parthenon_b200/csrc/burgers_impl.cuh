// burgers_impl.cuh — Parthenon-VIBE (benchmarks/burgers) stencil kernels for sm_100a.
//
// Included twice: burgers_strict.cu (nvcc -fmad=false, PB2_NS = strict) keeps the
// reference's expression order with no FMA contraction and is bit-identical to the
// reference's CPU build; burgers_fast.cu (FMA contraction on, PB2_NS = fast) is the
// throughput build (<= 1e-12 relative).
//
// Reference dataflow being replaced (benchmarks/burgers/burgers_package.cpp:202-404):
// a reconstruction kernel writes six 11-component scratch fields Ulx..Urz to memory and a
// Riemann kernel reads them back.  Here the left/right states never leave registers:
//   * y / z sweeps: a thread owns one (i, other) column and marches along the sweep
//     direction; lanes run along i, so every load is a coalesced 256 B row.  Each step
//     reconstructs the cell (5-point stencil re-read through L1), forms the flux at the
//     cell's lower face from the carried left state of the previous cell, and stores it.
//   * x sweep: rows are flattened to (row, cell) items, 32 consecutive items per warp
//     pass, so loads stay coalesced along i; the left state of the previous cell comes
//     from the neighbouring lane by warp shuffle (lane 31 of the previous pass for lane 0).
// Both do exactly (n+2)/n reconstructions per face row like the reference — no halo
// recomputation — and no block barriers (shared memory only parks the carried scalar
// left states, one slot per thread).  The FP64 pipe is the bound
// (SURVEY.md §8d); memory traffic per zone-stage is 11 reads + 33 flux writes.
#pragma once
#include <cfloat>

#include "common.cuh"

#ifndef PB2_NS
#error "define PB2_NS before including burgers_impl.cuh"
#endif

namespace pb2 {
namespace PB2_NS {

// std::min / std::max semantics (returns first argument on ties, incl. signed zeros)
__device__ __forceinline__ double min_std(double a, double b) { return (b < a) ? b : a; }
__device__ __forceinline__ double max_std(double a, double b) { return (a < b) ? b : a; }

// recon.hpp:27-32
__device__ __forceinline__ double mc(const double dm, const double dp) {
  const double dc = (dm * dp > 0.0) * 0.5 * (dm + dp);
  return copysign(min_std(fabs(dc), 2.0 * min_std(fabs(dm), fabs(dp))), dc);
}

// recon.hpp:34-40
__device__ __forceinline__ void Linear(const double qm, const double q0, const double qp,
                                       double &ql, double &qr) {
  double dq = qp - q0;
  dq = 0.5 * mc(q0 - qm, dq);
  ql = q0 + dq;
  qr = q0 - dq;
}

// recon.hpp:42-99
__device__ __forceinline__ void WENO5Z(const double q0, const double q1, const double q2,
                                       const double q3, const double q4, double &ql,
                                       double &qr) {
  constexpr double a00 = 1.0 / 3.0, a01 = -7.0 / 6.0, a02 = 11.0 / 6.0;
  constexpr double a10 = -1.0 / 6.0, a11 = 5.0 / 6.0, a12 = 1.0 / 3.0;
  constexpr double a20 = 1.0 / 3.0, a21 = 5.0 / 6.0, a22 = -1.0 / 6.0;
  constexpr double g0 = 0.1, g1 = 0.6, g2 = 0.3;
  constexpr double eps = 10.0 * DBL_EPSILON; // robust::EPS(), utils/robust.hpp:39-42
  constexpr double thirteen_thirds = 13.0 / 3.0;

  double a = q0 - 2 * q1 + q2;
  double b = q0 - 4.0 * q1 + 3.0 * q2;
  double beta0 = thirteen_thirds * a * a + b * b + eps;
  a = q1 - 2.0 * q2 + q3;
  b = q3 - q1;
  double beta1 = thirteen_thirds * a * a + b * b + eps;
  a = q2 - 2.0 * q3 + q4;
  b = q4 - 4.0 * q3 + 3.0 * q2;
  double beta2 = thirteen_thirds * a * a + b * b + eps;
  const double tau5 = fabs(beta2 - beta0);

  beta0 = (beta0 + tau5) / beta0;
  beta1 = (beta1 + tau5) / beta1;
  beta2 = (beta2 + tau5) / beta2;

  double w0 = g0 * beta0 + eps;
  double w1 = g1 * beta1 + eps;
  double w2 = g2 * beta2 + eps;
  double wsum = 1.0 / (w0 + w1 + w2);
  ql = w0 * (a00 * q0 + a01 * q1 + a02 * q2);
  ql += w1 * (a10 * q1 + a11 * q2 + a12 * q3);
  ql += w2 * (a20 * q2 + a21 * q3 + a22 * q4);
  ql *= wsum;
  const double alpha_l =
      3.0 * wsum * w0 * w1 * w2 / (g2 * w0 * w1 + g1 * w0 * w2 + g0 * w1 * w2) + eps;

  w0 = g0 * beta2 + eps;
  w1 = g1 * beta1 + eps;
  w2 = g2 * beta0 + eps;
  wsum = 1.0 / (w0 + w1 + w2);
  qr = w0 * (a00 * q4 + a01 * q3 + a02 * q2);
  qr += w1 * (a10 * q3 + a11 * q2 + a12 * q1);
  qr += w2 * (a20 * q2 + a21 * q1 + a22 * q0);
  qr *= wsum;
  const double alpha_r =
      3.0 * wsum * w0 * w1 * w2 / (g2 * w0 * w1 + g1 * w0 * w2 + g0 * w1 * w2) + eps;

  double dq = q3 - q2;
  dq = 0.5 * mc(q2 - q1, dq);

  const double alpha_lin = 2.0 * alpha_l * alpha_r / (alpha_l + alpha_r);
  ql = alpha_lin * ql + (1.0 - alpha_lin) * (q2 + dq);
  qr = alpha_lin * qr + (1.0 - alpha_lin) * (q2 - dq);
}

// reconstruct the cell at p along a direction with stride sd
template <int RECON>
__device__ __forceinline__ void recon_cell(const double *__restrict__ p, const int64_t sd,
                                           double &ql, double &qr) {
  if (RECON == PB2_RECON_WENO5) {
    WENO5Z(__ldg(p - 2 * sd), __ldg(p - sd), __ldg(p), __ldg(p + sd), __ldg(p + 2 * sd), ql,
           qr);
  } else {
    Linear(__ldg(p - sd), __ldg(p), __ldg(p + sd), ql, qr);
  }
}

// burgers_package.hpp:31-43
__device__ __forceinline__ void lr_to_flux(const double uxl, const double uxr,
                                           const double uyl, const double uyr,
                                           const double uzl, const double uzr,
                                           const double upl, const double upr, double &sl,
                                           double &sr, double &fux, double &fuy,
                                           double &fuz) {
  sl = min_std(min_std(upl, upr), 0.0);
  sr = max_std(max_std(upl, upr), 0.0);
  const double islsr = 1.0 / (sr - sl + (sl * sr == 0.0));
  fux = 0.5 * (sr * uxl * upl - sl * uxr * upr + sl * sr * (uxr - uxl)) * islsr;
  fuy = 0.5 * (sr * uyl * upl - sl * uyr * upr + sl * sr * (uyr - uyl)) * islsr;
  fuz = 0.5 * (sr * uzl * upl - sl * uzr * upr + sl * sr * (uzr - uzl)) * islsr;
}

// burgers_package.cpp:326-333 (qflux_loop)
__device__ __forceinline__ double scalar_flux(const double upl, const double upr,
                                              const double ql, const double qr,
                                              const double sl, const double sr) {
  return (sr * upl * ql - sl * upr * qr + sl * sr * (qr - ql)) /
         (sr - sl + (sl * sr == 0.0));
}

struct FluxGeom {
  int nblocks, ncomp, ndim;
  int nx[3], is[3], n[3]; // interior cells, interior start, full extents (i,j,k)
  int64_t sj, sk, sc, sb; // strides in Reals: j, k, component, block
};

constexpr int kFluxThreads = 128;
constexpr int kMaxComp = 16;

// ---- y / z sweeps -------------------------------------------------------------------------
// grid.x = blocks * ceil(columns / kFluxThreads), DIR = 1 (y) or 2 (z)
template <int RECON, int DIR>
__global__ void __launch_bounds__(kFluxThreads)
    flux_march_kernel(const FluxGeom g, const double *__restrict__ u,
                      double *__restrict__ flux, const int *__restrict__ block_ids) {
  const int ncol_other = (DIR == 1) ? g.nx[2] : g.nx[1]; // k for y sweep, j for z sweep
  const int ncol = ncol_other * g.nx[0];
  const int ctas_per_block = (ncol + kFluxThreads - 1) / kFluxThreads;
  const int bi = blockIdx.x / ctas_per_block;
  const int b = block_ids ? block_ids[bi] : bi;
  const int col = (blockIdx.x % ctas_per_block) * kFluxThreads + threadIdx.x;
  if (col >= ncol) return;
  const int i = g.is[0] + col % g.nx[0];
  const int o = col / g.nx[0];
  const int64_t sd = (DIR == 1) ? g.sj : g.sk;
  const int64_t so = (DIR == 1) ? g.sk : g.sj;
  const int os = (DIR == 1) ? g.is[2] : g.is[1];
  const int ds = g.is[DIR], nd = g.nx[DIR];
  const int64_t base = (int64_t)b * g.sb + (int64_t)(os + o) * so + i;
  const double *__restrict__ ub = u + base;
  double *__restrict__ fb = flux + base;
  const int nc = g.ncomp;

  // left state at the lower face of the current cell: velocities in registers, scalars
  // in a per-thread shared-memory column (the scalar loop is deliberately not unrolled so
  // the kernel body stays within the instruction cache)
  double carry[3] = {0.0, 0.0, 0.0};
  __shared__ double scarry[kMaxComp][kFluxThreads];
  for (int n = 3; n < nc; ++n) scarry[n][threadIdx.x] = 0.0;

  // cells ds-1 .. ds+nd ; faces ds .. ds+nd
  for (int s = -1; s <= nd; ++s) {
    const int64_t off = (int64_t)(ds + s) * sd;
    const bool face = s >= 0;
    double L[3], R[3], nl[3];
#pragma unroll
    for (int n = 0; n < 3; ++n) {
      double ql, qr;
      recon_cell<RECON>(ub + n * g.sc + off, sd, ql, qr);
      L[n] = carry[n];
      R[n] = qr;
      nl[n] = ql;
    }
    double sl = 0, sr = 0;
    const double upl = L[DIR], upr = R[DIR];
    if (face) {
      double f0, f1, f2;
      lr_to_flux(L[0], R[0], L[1], R[1], L[2], R[2], upl, upr, sl, sr, f0, f1, f2);
      fb[off] = f0;
      fb[g.sc + off] = f1;
      fb[2 * g.sc + off] = f2;
    }
#pragma unroll
    for (int n = 0; n < 3; ++n) carry[n] = nl[n];
#pragma unroll 1
    for (int n = 3; n < nc; ++n) {
      double ql, qr;
      recon_cell<RECON>(ub + n * g.sc + off, sd, ql, qr);
      if (face) fb[n * g.sc + off] = scalar_flux(upl, upr, scarry[n][threadIdx.x], qr, sl, sr);
      scarry[n][threadIdx.x] = ql;
    }
  }
}

// ---- x sweep ------------------------------------------------------------------------------
// A warp owns kRowsPerWarp consecutive (k,j) rows of one block; items are (row, c) with
// c = 0..nx1+1 <-> i = is-1+c, flattened row-major and taken 32 per pass.
constexpr int kRowsPerWarp = 16;

template <int RECON>
__global__ void __launch_bounds__(kFluxThreads)
    flux_x_kernel(const FluxGeom g, const double *__restrict__ u, double *__restrict__ flux,
                  const int *__restrict__ block_ids) {
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const int warp_global = blockIdx.x * (kFluxThreads / 32) + (threadIdx.x >> 5);
  const int nrows = g.nx[1] * g.nx[2];
  const int warps_per_block = (nrows + kRowsPerWarp - 1) / kRowsPerWarp;
  const int bi = warp_global / warps_per_block;
  if (bi >= g.nblocks) return; // whole warp exits together
  const int b = block_ids ? block_ids[bi] : bi;
  const int row0 = (warp_global % warps_per_block) * kRowsPerWarp;
  const int rows = min(kRowsPerWarp, nrows - row0);
  const int ncell = g.nx[0] + 2;
  const int items = rows * ncell;
  const int nc = g.ncomp;
  const double *__restrict__ ub = u + (int64_t)b * g.sb;
  double *__restrict__ fb = flux + (int64_t)b * g.sb;

  // ql of lane 31 in the previous pass (same value in every lane); scalars live in a
  // per-warp shared-memory row so the scalar loop need not be unrolled
  double last[3] = {0.0, 0.0, 0.0};
  __shared__ double slast[kFluxThreads / 32][kMaxComp];
  double *wl = slast[threadIdx.x >> 5];
  if (lane < kMaxComp) wl[lane] = 0.0;
  __syncwarp();

  for (int f0 = 0; f0 < items; f0 += 32) {
    const int f = f0 + lane;
    const bool cell = f < items;
    const int fr = cell ? f : items - 1; // clamp: inactive lanes recompute a valid cell
    const int r = fr / ncell, c = fr - r * ncell;
    const int row = row0 + r;
    const int k = g.is[2] + row / g.nx[1], j = g.is[1] + row % g.nx[1];
    const int i = g.is[0] - 1 + c;
    const int64_t off = (int64_t)k * g.sk + (int64_t)j * g.sj + i;
    const bool face = cell && c >= 1;
    double L[3], R[3];
#pragma unroll
    for (int n = 0; n < 3; ++n) {
      double ql, qr;
      recon_cell<RECON>(ub + n * g.sc + off, 1, ql, qr);
      double prev = __shfl_up_sync(full, ql, 1);
      if (lane == 0) prev = last[n];
      last[n] = __shfl_sync(full, ql, 31);
      L[n] = prev;
      R[n] = qr;
    }
    double sl = 0, sr = 0;
    const double upl = L[0], upr = R[0];
    if (face) {
      double fx, fy, fz;
      lr_to_flux(L[0], R[0], L[1], R[1], L[2], R[2], upl, upr, sl, sr, fx, fy, fz);
      fb[off] = fx;
      fb[g.sc + off] = fy;
      fb[2 * g.sc + off] = fz;
    }
#pragma unroll 1
    for (int n = 3; n < nc; ++n) {
      double ql, qr;
      recon_cell<RECON>(ub + n * g.sc + off, 1, ql, qr);
      double prev = __shfl_up_sync(full, ql, 1);
      if (lane == 0) prev = wl[n];
      __syncwarp();
      if (lane == 31) wl[n] = ql;
      __syncwarp();
      if (face) fb[n * g.sc + off] = scalar_flux(upl, upr, prev, qr, sl, sr);
    }
  }
}

// ---- update -------------------------------------------------------------------------------
// out = (beta*u + (1-beta)*base) + (beta*dt) * dudt,  dudt = -div F   (interior cells)
//   FluxDivHelper update.hpp:43-58, WeightedSumData update.hpp:71-91 (Average then Update,
//   burgers_driver.cpp:98-104), CalculateDerived burgers_package.cpp:143-167,
//   EstimateTimestepMesh :170-200.
struct UpdateArgs {
  FluxGeom g;
  const double *u, *base;
  double *out;
  const double *fx, *fy, *fz;
  const int *block_ids;     // launch block -> block of the batch, or null (identity)
  double *derived;          // [nblocks][nk][nj][ni] or null
  unsigned long long *dtmin; // bit pattern of a positive double, or null
  const double *dx;          // [nblocks][3]
  double beta, dt;
};

constexpr int kUpdThreads = 256;

__global__ void __launch_bounds__(kUpdThreads) update_kernel(const UpdateArgs a) {
  const FluxGeom &g = a.g;
  const int ncell = g.nx[0] * g.nx[1] * g.nx[2];
  const int ctas_per_block = (ncell + kUpdThreads - 1) / kUpdThreads;
  const int bi = blockIdx.x / ctas_per_block;
  const int b = a.block_ids ? a.block_ids[bi] : bi;
  const int t = (blockIdx.x % ctas_per_block) * kUpdThreads + threadIdx.x;
  double inv = DBL_MAX;
  if (t < ncell) {
    const int i = g.is[0] + t % g.nx[0];
    const int tj = t / g.nx[0];
    const int j = g.is[1] + tj % g.nx[1];
    const int k = g.is[2] + tj / g.nx[1];
    const int64_t p = (int64_t)b * g.sb + (int64_t)k * g.sk + (int64_t)j * g.sj + i;
    const double dx0 = a.dx[3 * b], dx1 = a.dx[3 * b + 1], dx2 = a.dx[3 * b + 2];
    // uniform_cartesian.hpp:36-39
    const double a1 = dx1 * dx2, a2 = dx0 * dx2, a3 = dx0 * dx1;
    const double vol = dx0 * dx1 * dx2;
    const double w2 = 1.0 - a.beta, bdt = a.beta * a.dt;
    double v4[4] = {0, 0, 0, 0};
#pragma unroll 1
    for (int n = 0; n < g.ncomp; ++n) {
      const int64_t q = p + n * g.sc;
      double du = (a1 * __ldg(a.fx + q + 1) - a1 * __ldg(a.fx + q));
      if (g.ndim >= 2) du += (a2 * __ldg(a.fy + q + g.sj) - a2 * __ldg(a.fy + q));
      if (g.ndim == 3) du += (a3 * __ldg(a.fz + q + g.sk) - a3 * __ldg(a.fz + q));
      const double dudt = -du / vol;
      const double avg = a.beta * __ldg(a.u + q) + w2 * __ldg(a.base + q);
      const double val = 1.0 * avg + bdt * dudt;
      a.out[q] = val;
      if (n < 4) v4[n] = val;
    }
    if (a.derived) {
      const int64_t pd = ((int64_t)b * g.n[2] + k) * g.sk + (int64_t)j * g.sj + i;
      a.derived[pd] = 0.5 * v4[3] * (v4[0] * v4[0] + v4[1] * v4[1] + v4[2] * v4[2]);
    }
    inv = 1.0 / ((fabs(v4[0])) / dx0 + (g.ndim > 1) * (fabs(v4[1])) / dx1 +
                 (g.ndim > 2) * (fabs(v4[2])) / dx2);
  }
  if (a.dtmin) {
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) inv = min_std(inv, __shfl_xor_sync(0xffffffffu, inv, s));
    if ((threadIdx.x & 31) == 0)
      atomicMin(a.dtmin, static_cast<unsigned long long>(__double_as_longlong(inv)));
  }
}

inline int make_geom(const pb2_pack_geom &pg, FluxGeom &g) {
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
  return 0;
}

template <int RECON>
int launch_fluxes_t(const pb2_burgers_args *args, cudaStream_t st) {
  FluxGeom g;
  make_geom(args->geom, g);
  const int *ids = args->block_ids;
  if (ids) g.nblocks = args->num_block_ids; // the launch covers the listed blocks only
  if (g.nblocks == 0) return PB2_OK;
  {
    const int nrows = g.nx[1] * g.nx[2];
    const int warps = g.nblocks * ((nrows + kRowsPerWarp - 1) / kRowsPerWarp);
    const int wpc = kFluxThreads / 32;
    ProfScope prof(K_FLUX_X, st);
    flux_x_kernel<RECON><<<(warps + wpc - 1) / wpc, kFluxThreads, 0, st>>>(g, args->u,
                                                                          args->flux[0], ids);
    PB2_LAUNCH_CHECK();
  }
  if (g.ndim > 1) {
    const int ncol = g.nx[2] * g.nx[0];
    const int ctas = g.nblocks * ((ncol + kFluxThreads - 1) / kFluxThreads);
    ProfScope prof(K_FLUX_Y, st);
    flux_march_kernel<RECON, 1><<<ctas, kFluxThreads, 0, st>>>(g, args->u, args->flux[1], ids);
    PB2_LAUNCH_CHECK();
  }
  if (g.ndim > 2) {
    const int ncol = g.nx[1] * g.nx[0];
    const int ctas = g.nblocks * ((ncol + kFluxThreads - 1) / kFluxThreads);
    ProfScope prof(K_FLUX_Z, st);
    flux_march_kernel<RECON, 2><<<ctas, kFluxThreads, 0, st>>>(g, args->u, args->flux[2], ids);
    PB2_LAUNCH_CHECK();
  }
  return PB2_OK;
}

inline int launch_fluxes(const pb2_burgers_args *args, cudaStream_t st) {
  if (args->recon == PB2_RECON_WENO5) return launch_fluxes_t<PB2_RECON_WENO5>(args, st);
  return launch_fluxes_t<PB2_RECON_LINEAR>(args, st);
}

inline int launch_update(const pb2_burgers_args *args, cudaStream_t st) {
  UpdateArgs a;
  make_geom(args->geom, a.g);
  a.u = args->u;
  a.base = args->base;
  a.out = args->out;
  a.fx = args->flux[0];
  a.fy = args->flux[1];
  a.fz = args->flux[2];
  a.block_ids = args->block_ids;
  if (a.block_ids) a.g.nblocks = args->num_block_ids;
  if (a.g.nblocks == 0) return PB2_OK;
  a.derived = args->derived;
  a.dtmin = reinterpret_cast<unsigned long long *>(args->dt_min);
  a.dx = args->geom.dx;
  a.beta = args->beta;
  a.dt = args->dt;
  const int ncell = a.g.nx[0] * a.g.nx[1] * a.g.nx[2];
  const int ctas = a.g.nblocks * ((ncell + kUpdThreads - 1) / kUpdThreads);
  ProfScope prof(K_UPDATE, st);
  update_kernel<<<ctas, kUpdThreads, 0, st>>>(a);
  PB2_LAUNCH_CHECK();
  return PB2_OK;
}

} // namespace PB2_NS
} // namespace pb2
