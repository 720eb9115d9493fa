// update.cu — dense per-stage updates (reference src/interface/update.hpp:43-91,
// update.cpp:63-86).  Compiled with -fmad=false: these are memory-bound, so keeping the
// reference's rounding costs nothing.
#include <algorithm>
#include <cfloat>

#include "common.cuh"

namespace pb2 {

// z = w1*x + w2*y (WeightedSumData update.hpp:71-91); 16-byte vectors, grid-stride
__global__ void __launch_bounds__(256)
    weighted_sum_kernel(const double *__restrict__ x, const double *__restrict__ y, double w1,
                        double w2, double *__restrict__ z, int64_t n) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool vec = ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y) |
                     reinterpret_cast<uintptr_t>(z)) & 15u) == 0;
  if (vec) {
    const int64_t nv = n / 2;
    const double2 *x2 = reinterpret_cast<const double2 *>(x);
    const double2 *y2 = reinterpret_cast<const double2 *>(y);
    double2 *z2 = reinterpret_cast<double2 *>(z);
    for (int64_t v = i; v < nv; v += stride) {
      const double2 a = x2[v], b = y2[v];
      double2 c;
      c.x = w1 * a.x + w2 * b.x;
      c.y = w1 * a.y + w2 * b.y;
      z2[v] = c;
    }
    if (i == 0 && (n & 1)) z[n - 1] = w1 * x[n - 1] + w2 * y[n - 1];
  } else {
    for (; i < n; i += stride) z[i] = w1 * x[i] + w2 * y[i];
  }
}

struct DivGeom {
  int nblocks, ncomp, ndim;
  int nx[3], is[3], n[3];
  int64_t sj, sk, sc, sb;
};

// WeightedSumData (update.hpp:71-91) restricted to the GHOST cells of every block: the part of
// the reference's full-extent passes that a fused interior update does not perform.  Needed on
// multilevel meshes, where fine ghosts facing a coarser block are not refreshed by the stage's
// exchange and so carry these values into the next stencil.
__global__ void __launch_bounds__(256)
    weighted_sum_ghost_kernel(const DivGeom g, const double *x, const double *y, double w1,
                              double w2, double *z, int64_t total,
                              const int *__restrict__ block_ids) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t ncell = g.sc;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
    const int64_t bc = e / ncell; // (block, component)
    int64_t t = e - bc * ncell;
    const int i = (int)(t % g.n[0]);
    t /= g.n[0];
    const int j = (int)(t % g.n[1]);
    const int k = (int)(t / g.n[1]);
    const bool interior = i >= g.is[0] && i < g.is[0] + g.nx[0] && j >= g.is[1] &&
                          j < g.is[1] + g.nx[1] && k >= g.is[2] && k < g.is[2] + g.nx[2];
    if (interior) continue;
    const int64_t bi = bc / g.ncomp, c = bc - bi * g.ncomp;
    const int64_t b = block_ids ? block_ids[bi] : bi;
    const int64_t p = b * g.sb + c * g.sc + (e - bc * ncell);
    z[p] = w1 * x[p] + w2 * y[p];
  }
}

// FluxDivergence<MeshData> update.cpp:63-86 + FluxDivHelper update.hpp:43-58
__global__ void __launch_bounds__(256)
    flux_div_kernel(const DivGeom g, const double *__restrict__ fx,
                    const double *__restrict__ fy, const double *__restrict__ fz,
                    const double *__restrict__ dx, double *__restrict__ dudt,
                    const int32_t *__restrict__ mask, const int32_t *__restrict__ list) {
  const int ncell = g.nx[0] * g.nx[1] * g.nx[2];
  const int ctas_per_block = (ncell + 255) / 256;
  const int slot = blockIdx.x / ctas_per_block;
  const int b = list != nullptr ? list[slot] : slot;
  if (mask != nullptr && mask[b] == 0) return;
  const int t = (blockIdx.x % ctas_per_block) * 256 + threadIdx.x;
  if (t >= ncell) return;
  const int i = g.is[0] + t % g.nx[0];
  const int tj = t / g.nx[0];
  const int j = g.is[1] + tj % g.nx[1];
  const int k = g.is[2] + tj / g.nx[1];
  const int64_t p = (int64_t)b * g.sb + (int64_t)k * g.sk + (int64_t)j * g.sj + i;
  const double dx0 = dx[3 * b], dx1 = dx[3 * b + 1], dx2 = dx[3 * b + 2];
  const double a1 = dx1 * dx2, a2 = dx0 * dx2, a3 = dx0 * dx1, vol = dx0 * dx1 * dx2;
  for (int n = 0; n < g.ncomp; ++n) {
    const int64_t q = p + n * g.sc;
    double du = (a1 * fx[q + 1] - a1 * fx[q]);
    if (g.ndim >= 2) du += (a2 * fy[q + g.sj] - a2 * fy[q]);
    if (g.ndim == 3) du += (a3 * fz[q + g.sk] - a3 * fz[q]);
    dudt[q] = -du / vol;
  }
}

// WeightedSumData for a sparse field: whole blocks, skipped where the field is unallocated
__global__ void __launch_bounds__(256)
    weighted_sum_blocks_kernel(const double *x, const double *y, double w1, double w2, double *z,
                               int64_t block_stride, int64_t per_block,
                               const int32_t *__restrict__ mask, int ctas_per_block,
                               const int32_t *__restrict__ list) {
  const int slot = blockIdx.x / ctas_per_block;
  const int b = list != nullptr ? list[slot] : slot;
  if (mask != nullptr && mask[b] == 0) return;
  const int64_t o = (int64_t)b * block_stride;
  for (int64_t i = (int64_t)(blockIdx.x % ctas_per_block) * 256 + threadIdx.x; i < per_block;
       i += (int64_t)ctas_per_block * 256)
    z[o + i] = w1 * x[o + i] + w2 * y[o + i];
}

// SparseDealloc update.cpp:161-186: is every value of the block within the threshold?
__global__ void __launch_bounds__(256)
    block_quiet_kernel(const double *__restrict__ u, int64_t block_stride, int64_t per_block,
                       double threshold, const int32_t *__restrict__ mask,
                       int32_t *__restrict__ quiet) {
  const int b = blockIdx.x;
  if (mask != nullptr && mask[b] == 0) return;
  const double *p = u + (int64_t)b * block_stride;
  bool loud = false;
  for (int64_t i = threadIdx.x; i < per_block; i += 256) loud = loud || (fabs(p[i]) > threshold);
  const int any = __syncthreads_or(loud);
  if (threadIdx.x == 0) quiet[b] = any ? 0 : 1;
}

// per-block minimum / maximum over all components and the ENTIRE extents: the reduction of
// example/advection's CheckRefinement (advection_package.cpp:252-263, Kokkos::MinMax)
__global__ void __launch_bounds__(256)
    block_minmax_kernel(const double *__restrict__ u, int64_t block_stride, int64_t per_block,
                        const int32_t *__restrict__ mask, double *__restrict__ out) {
  const int b = blockIdx.x;
  if (mask != nullptr && mask[b] == 0) return;
  const double *p = u + (int64_t)b * block_stride;
  double mn = DBL_MAX, mx = -DBL_MAX;
  for (int64_t i = threadIdx.x; i < per_block; i += 256) {
    const double v = p[i];
    mn = v < mn ? v : mn;
    mx = v > mx ? v : mx;
  }
  __shared__ double smn[256], smx[256];
  smn[threadIdx.x] = mn;
  smx[threadIdx.x] = mx;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) {
      smn[threadIdx.x] = smn[threadIdx.x + s] < smn[threadIdx.x] ? smn[threadIdx.x + s] : smn[threadIdx.x];
      smx[threadIdx.x] = smx[threadIdx.x + s] > smx[threadIdx.x] ? smx[threadIdx.x + s] : smx[threadIdx.x];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    out[2 * b] = smn[0];
    out[2 * b + 1] = smx[0];
  }
}

// Refinement::FirstDerivative / SecondDerivative (amr_criteria/refinement_package.cpp:92-150):
// the largest normalised first (ORDER 1) or second (ORDER 2) difference of one component over
// the interior cells of every block.  max is exact in any order, so a plain tree reduction
// reproduces the reference's Kokkos::Max bit for bit.
template <int ORDER>
__global__ void __launch_bounds__(256)
    block_derivative_kernel(const DivGeom g, const double *__restrict__ u, int comp,
                            double *__restrict__ out) {
  const int b = blockIdx.x;
  const int ncell = g.nx[0] * g.nx[1] * g.nx[2];
  const double *q0 = u + (int64_t)b * g.sb + (int64_t)comp * g.sc;
  const int64_t str[3] = {1, g.sj, g.sk};
  double maxd = 0.0;
  for (int t = threadIdx.x; t < ncell; t += 256) {
    const int i = g.is[0] + t % g.nx[0];
    const int tj = t / g.nx[0];
    const int j = g.is[1] + tj % g.nx[1];
    const int k = g.is[2] + tj / g.nx[1];
    const double *q = q0 + (int64_t)k * g.sk + (int64_t)j * g.sj + i;
    const double c = q[0];
    for (int d = 0; d < g.ndim; ++d) {
      double v;
      if (ORDER == 1) {
        v = 0.5 * fabs((q[str[d]] - q[-str[d]])) / (fabs(c) + 1.0e-20);
      } else {
        const double aqt = fabs(c) + 1.0e-20;
        const double qavg = 0.5 * (q[str[d]] + q[-str[d]]);
        v = fabs(qavg - c) / (fabs(qavg) + aqt);
      }
      maxd = v > maxd ? v : maxd;
    }
  }
  __shared__ double sm[256];
  sm[threadIdx.x] = maxd;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) sm[threadIdx.x] = sm[threadIdx.x + s] > sm[threadIdx.x] ? sm[threadIdx.x + s] : sm[threadIdx.x];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[b] = sm[0];
}

// example/advection CalculateFluxes with a constant velocity (advection_package.cpp:540-646;
// DonorCellX1/2/3 reconstruct/dc_inline.hpp:31-71): the flux through the lower d-face of a
// cell is the upwind cell value times v_d.  One thread per (block, component, cell of the
// interior extended by one layer on the upper side); it writes the d-faces it owns.
__global__ void __launch_bounds__(256)
    advection_flux_kernel(const DivGeom g, const double *__restrict__ u,
                          double *__restrict__ fx, double *__restrict__ fy,
                          double *__restrict__ fz, double vx, double vy, double vz,
                          int64_t total, const int32_t *__restrict__ mask,
                          const int32_t *__restrict__ list) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int e0 = g.nx[0] + 1, e1 = g.nx[1] + (g.ndim > 1), e2 = g.nx[2] + (g.ndim > 2);
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
    int64_t t = e;
    const int di = (int)(t % e0);
    t /= e0;
    const int dj = (int)(t % e1);
    t /= e1;
    const int dk = (int)(t % e2);
    t /= e2;
    const int c = (int)(t % g.ncomp);
    const int64_t slot = t / g.ncomp;
    const int64_t b = list != nullptr ? list[slot] : slot;
    if (mask != nullptr && mask[b] == 0) continue;
    const int64_t p = b * g.sb + c * g.sc + (int64_t)(g.is[2] + dk) * g.sk +
                      (int64_t)(g.is[1] + dj) * g.sj + (g.is[0] + di);
    const double q = u[p];
    const bool in0 = di < g.nx[0], in1 = dj < g.nx[1], in2 = dk < g.nx[2];
    if (in1 && in2) fx[p] = vx > 0.0 ? u[p - 1] * vx : q * vx;
    if (g.ndim > 1 && in0 && in2) fy[p] = vy > 0.0 ? u[p - g.sj] * vy : q * vy;
    if (g.ndim > 2 && in0 && in1) fz[p] = vz > 0.0 ? u[p - g.sk] * vz : q * vz;
  }
}

// interior cells <-> packed host-layout staging buffer [block][comp][nx3][nx2][nx1]
// (what an application's host arrays hold: no ghosts).  16-byte vectors when nx1 is even.
template <bool SCATTER>
__global__ void __launch_bounds__(256)
    interior_kernel(const DivGeom g, double *__restrict__ field, double *__restrict__ packed,
                    int64_t total) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int nx0 = g.nx[0], nx1 = g.nx[1], nx2 = g.nx[2];
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
    int64_t t = e;
    const int i = (int)(t % nx0);
    t /= nx0;
    const int j = (int)(t % nx1);
    t /= nx1;
    const int k = (int)(t % nx2);
    t /= nx2;
    const int c = (int)(t % g.ncomp);
    const int64_t b = t / g.ncomp;
    const int64_t f = b * g.sb + c * g.sc + (int64_t)(k + g.is[2]) * g.sk +
                      (int64_t)(j + g.is[1]) * g.sj + (i + g.is[0]);
    if (SCATTER)
      field[f] = packed[e];
    else
      packed[e] = field[f];
  }
}

} // namespace pb2

using namespace pb2;

extern "C" {

int pb2_weighted_sum(const double *x, const double *y, double w1, double w2, double *z,
                     int64_t n, pb2_stream_t stream) {
  PB2_REQUIRE(x && y && z && n >= 0, "bad arguments");
  if (int rc = require_device()) return rc;
  if (n == 0) return PB2_OK;
  int64_t ctas = (n / 2 + 255) / 256;
  if (ctas > 148 * 16) ctas = 148 * 16;
  if (ctas < 1) ctas = 1;
  ProfScope prof(K_WEIGHTED_SUM, as_stream(stream), static_cast<double>(n));
  weighted_sum_kernel<<<static_cast<unsigned>(ctas), 256, 0, as_stream(stream)>>>(x, y, w1, w2,
                                                                                 z, n);
  PB2_LAUNCH_CHECK();
  return PB2_OK;
}

int pb2_weighted_sum_ghosts(const pb2_pack_geom *pg, const double *x, const double *y, double w1,
                            double w2, double *z, pb2_stream_t stream) {
  return pb2_weighted_sum_ghosts_blocks(pg, x, y, w1, w2, z, nullptr, 0, stream);
}

int pb2_weighted_sum_ghosts_blocks(const pb2_pack_geom *pg, const double *x, const double *y,
                                   double w1, double w2, double *z, const int32_t *block_ids,
                                   int32_t num_block_ids, pb2_stream_t stream) {
  PB2_REQUIRE(pg && x && y && z, "bad arguments");
  if (int rc = require_device()) return rc;
  DivGeom g;
  g.nblocks = block_ids ? num_block_ids : pg->nblocks;
  g.ncomp = pg->ncomp;
  g.ndim = pg->ndim;
  for (int d = 0; d < 3; ++d) {
    const bool sym = d >= pg->ndim;
    g.nx[d] = sym ? 1 : pg->nx[d];
    g.is[d] = sym ? 0 : pg->ng;
    g.n[d] = sym ? 1 : pg->nx[d] + 2 * pg->ng;
  }
  g.sj = g.n[0];
  g.sk = (int64_t)g.n[0] * g.n[1];
  g.sc = g.sk * g.n[2];
  g.sb = pg->block_stride;
  const int64_t total = (int64_t)g.nblocks * g.ncomp * g.sc;
  if (total == 0) return PB2_OK;
  const unsigned ctas = static_cast<unsigned>(std::min<int64_t>((total + 255) / 256, 148 * 32));
  ProfScope prof(K_WEIGHTED_SUM, as_stream(stream), static_cast<double>(total));
  weighted_sum_ghost_kernel<<<ctas, 256, 0, as_stream(stream)>>>(g, x, y, w1, w2, z, total,
                                                                 block_ids);
  PB2_LAUNCH_CHECK();
  return PB2_OK;
}

static int interior_copy(const pb2_pack_geom *pg, double *field, double *packed, bool scatter,
                         pb2_stream_t stream) {
  PB2_REQUIRE(pg && field && packed, "bad arguments");
  if (int rc = require_device()) return rc;
  DivGeom g;
  g.nblocks = pg->nblocks;
  g.ncomp = pg->ncomp;
  g.ndim = pg->ndim;
  for (int d = 0; d < 3; ++d) {
    const bool sym = d >= pg->ndim;
    g.nx[d] = sym ? 1 : pg->nx[d];
    g.is[d] = sym ? 0 : pg->ng;
    g.n[d] = sym ? 1 : pg->nx[d] + 2 * pg->ng;
  }
  g.sj = g.n[0];
  g.sk = (int64_t)g.n[0] * g.n[1];
  g.sc = g.sk * g.n[2];
  g.sb = pg->block_stride;
  const int64_t total = (int64_t)g.nblocks * g.ncomp * g.nx[0] * g.nx[1] * g.nx[2];
  if (total == 0) return PB2_OK;
  const unsigned ctas = static_cast<unsigned>(std::min<int64_t>((total + 255) / 256, 148 * 32));
  ProfScope prof(K_INTERIOR, as_stream(stream), static_cast<double>(total));
  if (scatter)
    interior_kernel<true><<<ctas, 256, 0, as_stream(stream)>>>(g, field, packed, total);
  else
    interior_kernel<false><<<ctas, 256, 0, as_stream(stream)>>>(g, field, packed, total);
  PB2_LAUNCH_CHECK();
  return PB2_OK;
}

int pb2_interior_scatter(const pb2_pack_geom *g, const double *packed, double *field,
                         pb2_stream_t stream) {
  return interior_copy(g, field, const_cast<double *>(packed), true, stream);
}
int pb2_interior_gather(const pb2_pack_geom *g, const double *field, double *packed,
                        pb2_stream_t stream) {
  return interior_copy(g, const_cast<double *>(field), packed, false, stream);
}

int pb2_advection_fluxes(const pb2_pack_geom *pg, const double *u, double *const flux[3],
                         const double v[3], pb2_stream_t stream) {
  return pb2_advection_fluxes_blocks(pg, u, flux, v, nullptr, stream);
}

int pb2_advection_fluxes_blocks(const pb2_pack_geom *pg, const double *u, double *const flux[3],
                                const double v[3], const int32_t *block_mask,
                                pb2_stream_t stream) {
  PB2_REQUIRE(pg && u && flux && v && flux[0], "bad arguments");
  PB2_REQUIRE(pg->ndim < 2 || flux[1], "null x2 flux");
  PB2_REQUIRE(pg->ndim < 3 || flux[2], "null x3 flux");
  PB2_REQUIRE(pg->ng >= 1, "donor cell needs one ghost layer");
  if (int rc = require_device()) return rc;
  DivGeom g;
  g.nblocks = pg->nblocks;
  g.ncomp = pg->ncomp;
  g.ndim = pg->ndim;
  for (int d = 0; d < 3; ++d) {
    const bool sym = d >= pg->ndim;
    g.nx[d] = sym ? 1 : pg->nx[d];
    g.is[d] = sym ? 0 : pg->ng;
    g.n[d] = sym ? 1 : pg->nx[d] + 2 * pg->ng;
  }
  g.sj = g.n[0];
  g.sk = (int64_t)g.n[0] * g.n[1];
  g.sc = g.sk * g.n[2];
  g.sb = pg->block_stride;
  const int64_t total = (int64_t)(pg->block_list ? pg->nlist : g.nblocks) * g.ncomp *
                        (g.nx[0] + 1) * (g.nx[1] + (g.ndim > 1)) * (g.nx[2] + (g.ndim > 2));
  if (total == 0) return PB2_OK;
  const unsigned ctas = static_cast<unsigned>(std::min<int64_t>((total + 255) / 256, 148 * 32));
  ProfScope prof(K_ADVECTION_FLUX, as_stream(stream), static_cast<double>(total));
  advection_flux_kernel<<<ctas, 256, 0, as_stream(stream)>>>(g, u, flux[0], flux[1], flux[2],
                                                            v[0], v[1], v[2], total, block_mask,
                                                            pg->block_list);
  PB2_LAUNCH_CHECK();
  return PB2_OK;
}

int pb2_flux_divergence(const pb2_pack_geom *pg, const double *const flux[3], double *dudt,
                        pb2_stream_t stream) {
  return pb2_flux_divergence_blocks(pg, flux, dudt, nullptr, stream);
}

int pb2_weighted_sum_blocks(const pb2_pack_geom *pg, const double *x, const double *y, double w1,
                            double w2, double *z, const int32_t *block_mask,
                            pb2_stream_t stream) {
  PB2_REQUIRE(pg && x && y && z, "bad arguments");
  if (int rc = require_device()) return rc;
  if (pg->nblocks == 0) return PB2_OK;
  int64_t per_block = pg->ncomp;
  for (int d = 0; d < 3; ++d) per_block *= d >= pg->ndim ? 1 : pg->nx[d] + 2 * pg->ng;
  const int cpb = static_cast<int>(std::min<int64_t>((per_block + 255) / 256, 64));
  const int nslots = pg->block_list ? pg->nlist : pg->nblocks;
  ProfScope prof(K_WEIGHTED_SUM, as_stream(stream), static_cast<double>(per_block) * nslots);
  if (nslots == 0) return PB2_OK;
  weighted_sum_blocks_kernel<<<nslots * cpb, 256, 0, as_stream(stream)>>>(
      x, y, w1, w2, z, pg->block_stride, per_block, block_mask, cpb, pg->block_list);
  PB2_LAUNCH_CHECK();
  return PB2_OK;
}

int pb2_block_derivative(const pb2_pack_geom *pg, const double *u, int comp, int order,
                         double *maxd, pb2_stream_t stream) {
  PB2_REQUIRE(pg && u && maxd, "bad arguments");
  PB2_REQUIRE(comp >= 0 && comp < pg->ncomp, "component out of range");
  PB2_REQUIRE(order == 1 || order == 2, "order must be 1 or 2");
  PB2_REQUIRE(pg->ng >= 1, "the criterion reads one ghost layer");
  if (int rc = require_device()) return rc;
  if (pg->nblocks == 0) return PB2_OK;
  DivGeom g;
  g.nblocks = pg->nblocks;
  g.ncomp = pg->ncomp;
  g.ndim = pg->ndim;
  for (int d = 0; d < 3; ++d) {
    const bool sym = d >= pg->ndim;
    g.nx[d] = sym ? 1 : pg->nx[d];
    g.is[d] = sym ? 0 : pg->ng;
    g.n[d] = sym ? 1 : pg->nx[d] + 2 * pg->ng;
  }
  g.sj = g.n[0];
  g.sk = (int64_t)g.n[0] * g.n[1];
  g.sc = g.sk * g.n[2];
  g.sb = pg->block_stride;
  ProfScope prof(K_WEIGHTED_SUM, as_stream(stream));
  if (order == 1)
    block_derivative_kernel<1><<<g.nblocks, 256, 0, as_stream(stream)>>>(g, u, comp, maxd);
  else
    block_derivative_kernel<2><<<g.nblocks, 256, 0, as_stream(stream)>>>(g, u, comp, maxd);
  PB2_LAUNCH_CHECK();
  return PB2_OK;
}

int pb2_block_minmax(const pb2_pack_geom *pg, const double *u, const int32_t *block_mask,
                     double *minmax, pb2_stream_t stream) {
  PB2_REQUIRE(pg && u && minmax, "bad arguments");
  if (int rc = require_device()) return rc;
  if (pg->nblocks == 0) return PB2_OK;
  int64_t per_block = pg->ncomp;
  for (int d = 0; d < 3; ++d) per_block *= d >= pg->ndim ? 1 : pg->nx[d] + 2 * pg->ng;
  ProfScope prof(K_WEIGHTED_SUM, as_stream(stream));
  block_minmax_kernel<<<pg->nblocks, 256, 0, as_stream(stream)>>>(u, pg->block_stride, per_block,
                                                                  block_mask, minmax);
  PB2_LAUNCH_CHECK();
  return PB2_OK;
}

int pb2_block_quiet_flags(const pb2_pack_geom *pg, const double *u, double threshold,
                          const int32_t *block_mask, int32_t *quiet, pb2_stream_t stream) {
  PB2_REQUIRE(pg && u && quiet, "bad arguments");
  if (int rc = require_device()) return rc;
  if (pg->nblocks == 0) return PB2_OK;
  int64_t per_block = pg->ncomp;
  for (int d = 0; d < 3; ++d) per_block *= d >= pg->ndim ? 1 : pg->nx[d] + 2 * pg->ng;
  ProfScope prof(K_WEIGHTED_SUM, as_stream(stream));
  block_quiet_kernel<<<pg->nblocks, 256, 0, as_stream(stream)>>>(u, pg->block_stride, per_block,
                                                                 threshold, block_mask, quiet);
  PB2_LAUNCH_CHECK();
  return PB2_OK;
}

int pb2_flux_divergence_blocks(const pb2_pack_geom *pg, const double *const flux[3],
                               double *dudt, const int32_t *block_mask, pb2_stream_t stream) {
  PB2_REQUIRE(pg && flux && dudt && pg->dx, "bad arguments");
  if (int rc = require_device()) return rc;
  DivGeom g;
  g.nblocks = pg->nblocks;
  g.ncomp = pg->ncomp;
  g.ndim = pg->ndim;
  for (int d = 0; d < 3; ++d) {
    const bool sym = d >= pg->ndim;
    g.nx[d] = sym ? 1 : pg->nx[d];
    g.is[d] = sym ? 0 : pg->ng;
    g.n[d] = sym ? 1 : pg->nx[d] + 2 * pg->ng;
  }
  g.sj = g.n[0];
  g.sk = (int64_t)g.n[0] * g.n[1];
  g.sc = g.sk * g.n[2];
  g.sb = pg->block_stride;
  const int ncell = g.nx[0] * g.nx[1] * g.nx[2];
  const int ctas = (pg->block_list ? pg->nlist : g.nblocks) * ((ncell + 255) / 256);
  if (ctas == 0) return PB2_OK;
  ProfScope prof(K_FLUX_DIV, as_stream(stream), static_cast<double>(ncell) * g.ncomp *
                                                    (pg->block_list ? pg->nlist : g.nblocks));
  flux_div_kernel<<<ctas, 256, 0, as_stream(stream)>>>(g, flux[0], flux[1], flux[2], pg->dx,
                                                      dudt, block_mask, pg->block_list);
  PB2_LAUNCH_CHECK();
  return PB2_OK;
}

} // extern "C"
