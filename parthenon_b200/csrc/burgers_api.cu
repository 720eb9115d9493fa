// burgers_api.cu — C-ABI entry points of the burgers stencil; dispatch strict / fast builds.
#include <cfloat>
#include <cstdlib>

#include "common.cuh"

namespace pb2 {
int burgers_fluxes_strict(const pb2_burgers_args *, cudaStream_t);
int burgers_update_strict(const pb2_burgers_args *, cudaStream_t);
int burgers_fluxes_fast(const pb2_burgers_args *, cudaStream_t);
int burgers_update_fast(const pb2_burgers_args *, cudaStream_t);
int burgers_stage_sweep(const pb2_burgers_args *, cudaStream_t);

static int check_args(const pb2_burgers_args *a, bool need_update, bool need_flux = true) {
  PB2_REQUIRE(a, "null args");
  const pb2_pack_geom &g = a->geom;
  PB2_REQUIRE(g.ndim >= 1 && g.ndim <= 3, "ndim must be 1..3");
  PB2_REQUIRE(g.ncomp >= 4 && g.ncomp <= 16, "ncomp must be 4..16 (3 velocities + scalars)");
  PB2_REQUIRE(g.nblocks >= 0, "negative block count");
  PB2_REQUIRE(a->recon == PB2_RECON_WENO5 || a->recon == PB2_RECON_LINEAR, "unknown recon");
  PB2_REQUIRE(a->math == PB2_MATH_STRICT || a->math == PB2_MATH_FAST, "unknown math mode");
  PB2_REQUIRE(g.ng >= (a->recon == PB2_RECON_WENO5 ? 3 : 2),
              "not enough ghost cells for the reconstruction stencil");
  PB2_REQUIRE(a->u, "null field pointer");
  if (need_flux)
    for (int d = 0; d < g.ndim; ++d) PB2_REQUIRE(a->flux[d], "null flux pointer");
  if (need_update) PB2_REQUIRE(a->base && a->out && g.dx, "null update pointer");
  return PB2_OK;
}

struct CellGeom {
  int nblocks, ncomp, ndim;
  int nx[3], is[3], n[3];
  int64_t sj, sk, sc, sb;
};

static void make_cell_geom(const pb2_pack_geom &pg, CellGeom &g) {
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
}

// CalculateDerived burgers_package.cpp:143-167 and EstimateTimestepMesh :170-200 as
// stand-alone passes (the fused stage kernel produces both on the fly)
__global__ void __launch_bounds__(256)
    derived_dt_kernel(const CellGeom g, const double *__restrict__ u,
                      const double *__restrict__ dx, double *__restrict__ derived,
                      unsigned long long *__restrict__ dtmin) {
  const int ncell = g.nx[0] * g.nx[1] * g.nx[2];
  const int ctas_per_block = (ncell + 255) / 256;
  const int b = blockIdx.x / ctas_per_block;
  const int t = (blockIdx.x % ctas_per_block) * 256 + threadIdx.x;
  double inv = DBL_MAX;
  if (t < ncell) {
    const int i = g.is[0] + t % g.nx[0];
    const int tj = t / g.nx[0];
    const int j = g.is[1] + tj % g.nx[1];
    const int k = g.is[2] + tj / g.nx[1];
    const int64_t cell = (int64_t)k * g.sk + (int64_t)j * g.sj + i;
    const int64_t p = (int64_t)b * g.sb + cell;
    const double u0 = u[p], u1 = u[p + g.sc], u2 = u[p + 2 * g.sc];
    if (derived) derived[(int64_t)b * g.sc + cell] = 0.5 * u[p + 3 * g.sc] * (u0 * u0 + u1 * u1 + u2 * u2);
    const double dx0 = dx[3 * b], dx1 = dx[3 * b + 1], dx2 = dx[3 * b + 2];
    inv = 1.0 / ((fabs(u0)) / dx0 + (g.ndim > 1) * (fabs(u1)) / dx1 + (g.ndim > 2) * (fabs(u2)) / dx2);
  }
  if (dtmin) {
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) {
      const double o = __shfl_xor_sync(0xffffffffu, inv, s);
      inv = (o < inv) ? o : inv;
    }
    if ((threadIdx.x & 31) == 0)
      atomicMin(dtmin, static_cast<unsigned long long>(__double_as_longlong(inv)));
  }
}

// MassHistory burgers_package.cpp:406-439: out[o] = sum mask_o * q^2 * vol / (mesh_vol+1e-20)
struct HistGeom {
  int nblocks, ncomp, ndim;
  int nx[3], is[3], n[3];
  int64_t sj, sk, sc, sb;
  double lo[3], mid[3], hi[3], mesh_vol;
};

__global__ void __launch_bounds__(256)
    history_kernel(const HistGeom g, const double *__restrict__ u,
                   const double *__restrict__ dx, const double *__restrict__ bxmin,
                   double *__restrict__ partial) {
  // one CTA per (block, component); threads stride over interior cells
  const int b = blockIdx.x / g.ncomp, v = blockIdx.x % g.ncomp;
  const int ncell = g.nx[0] * g.nx[1] * g.nx[2];
  const double dx0 = dx[3 * b], dx1 = dx[3 * b + 1], dx2 = dx[3 * b + 2];
  const double weight = (dx0 * dx1 * dx2) / (g.mesh_vol + 1e-20);
  // UniformCartesian xmin_ = block xmin - istart*dx (uniform_cartesian.hpp:33-35)
  const double x0 = bxmin[3 * b] - g.is[0] * dx0, y0 = bxmin[3 * b + 1] - g.is[1] * dx1,
               z0 = bxmin[3 * b + 2] - g.is[2] * dx2;
  double acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int t = threadIdx.x; t < ncell; t += blockDim.x) {
    const int i = g.is[0] + t % g.nx[0];
    const int tj = t / g.nx[0];
    const int j = g.is[1] + tj % g.nx[1];
    const int k = g.is[2] + tj / g.nx[1];
    const double x1 = x0 + (i + 0.5) * dx0, x2 = y0 + (j + 0.5) * dx1,
                 x3 = z0 + (k + 0.5) * dx2;
    const double q =
        u[(int64_t)b * g.sb + v * g.sc + (int64_t)k * g.sk + (int64_t)j * g.sj + i];
    const double w = q * q * weight;
#pragma unroll
    for (int o = 0; o < 8; ++o) {
      const int s1 = (o >> 2) & 1, s2 = (o >> 1) & 1, s3 = o & 1; // octant order :94-107
      const double l1 = s1 ? g.mid[0] : g.lo[0], h1 = s1 ? g.hi[0] : g.mid[0];
      const double l2 = s2 ? g.mid[1] : g.lo[1], h2 = s2 ? g.hi[1] : g.mid[1];
      const double l3 = s3 ? g.mid[2] : g.lo[2], h3 = s3 ? g.hi[2] : g.mid[2];
      const bool in = (l1 <= x1) && (x1 <= h1) && (l2 <= x2) && (x2 <= h2) && (l3 <= x3) &&
                      (x3 <= h3);
      if (in) acc[o] += w;
    }
  }
  __shared__ double red[8][256];
#pragma unroll
  for (int o = 0; o < 8; ++o) red[o][threadIdx.x] = acc[o];
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) {
#pragma unroll
      for (int o = 0; o < 8; ++o) red[o][threadIdx.x] += red[o][threadIdx.x + s];
    }
    __syncthreads();
  }
  if (threadIdx.x < 8) partial[(int64_t)blockIdx.x * 8 + threadIdx.x] = red[threadIdx.x][0];
}

__global__ void history_final_kernel(const double *__restrict__ partial, int n,
                                     double *__restrict__ out) {
  // fixed-order (deterministic) tree over CTA partials, one warp-sized CTA per octant
  const int o = blockIdx.x;
  __shared__ double red[256];
  double acc = 0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) acc += partial[(int64_t)i * 8 + o];
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[o] = red[0];
}

} // namespace pb2

using namespace pb2;

extern "C" {

int pb2_burgers_calculate_fluxes(const pb2_burgers_args *args, pb2_stream_t stream) {
  if (int rc = check_args(args, false)) return rc;
  if (int rc = require_device()) return rc;
  if (args->geom.nblocks == 0) return PB2_OK;
  return args->math == PB2_MATH_STRICT ? burgers_fluxes_strict(args, as_stream(stream))
                                       : burgers_fluxes_fast(args, as_stream(stream));
}

int pb2_burgers_update(const pb2_burgers_args *args, pb2_stream_t stream) {
  if (int rc = check_args(args, true)) return rc;
  if (int rc = require_device()) return rc;
  if (args->geom.nblocks == 0) return PB2_OK;
  return args->math == PB2_MATH_STRICT ? burgers_update_strict(args, as_stream(stream))
                                       : burgers_update_fast(args, as_stream(stream));
}

int pb2_burgers_stage(const pb2_burgers_args *args, pb2_stream_t stream) {
  // FAST: three flux-free direction sweeps (burgers_sweep.cu); STRICT: the reference's
  // dataflow with stored fluxes, bit-exact
  if (args && args->math == PB2_MATH_FAST) {
    if (int rc = check_args(args, true, false)) return rc;
    if (int rc = require_device()) return rc;
    if (args->geom.nblocks == 0) return PB2_OK;
    PB2_REQUIRE(args->out != args->u, "the fused stage cannot write its stencil input");
    PB2_REQUIRE(args->progress == nullptr || args->geom.ndim >= 2,
                "the progress counter needs a 2-D or 3-D mesh");
    return burgers_stage_sweep(args, as_stream(stream));
  }
  if (int rc = pb2_burgers_calculate_fluxes(args, stream)) return rc;
  return pb2_burgers_update(args, stream);
}

int32_t pb2_burgers_progress_target(const pb2_pack_geom *g, int32_t nblocks) {
  if (!g || g->ndim < 2 || nblocks <= 0) return 0;
  // thread blocks per meshblock of the last sweep: 128 columns each (burgers_march.cuh)
  const int ncol = g->nx[0] * (g->ndim == 2 ? 1 : g->nx[1]);
  return nblocks * ((ncol + 127) / 128);
}

int pb2_burgers_derived_dt(const pb2_pack_geom *pg, const double *u, double *derived,
                           double *dt_min, pb2_stream_t stream) {
  PB2_REQUIRE(pg && u && pg->dx && (derived || dt_min), "bad arguments");
  PB2_REQUIRE(pg->ncomp >= 4, "ncomp must be >= 4");
  if (int rc = require_device()) return rc;
  if (pg->nblocks == 0) return PB2_OK;
  CellGeom g;
  make_cell_geom(*pg, g);
  const int ncell = g.nx[0] * g.nx[1] * g.nx[2];
  const int ctas = g.nblocks * ((ncell + 255) / 256);
  ProfScope prof(K_DERIVED_DT, as_stream(stream));
  derived_dt_kernel<<<ctas, 256, 0, as_stream(stream)>>>(
      g, u, pg->dx, derived, reinterpret_cast<unsigned long long *>(dt_min));
  PB2_LAUNCH_CHECK();
  return PB2_OK;
}

int pb2_burgers_history(const pb2_pack_geom *pg, const double *u, const double *block_xmin,
                        const double mesh_xmin[3], const double mesh_xmax[3], double out[8],
                        pb2_stream_t stream) {
  PB2_REQUIRE(pg && u && block_xmin && mesh_xmin && mesh_xmax && out && pg->dx,
              "bad arguments");
  if (int rc = require_device()) return rc;
  HistGeom g;
  g.nblocks = pg->nblocks;
  g.ncomp = pg->ncomp;
  g.ndim = pg->ndim;
  g.mesh_vol = 1;
  for (int d = 0; d < 3; ++d) {
    const bool sym = d >= pg->ndim;
    g.nx[d] = sym ? 1 : pg->nx[d];
    g.is[d] = sym ? 0 : pg->ng;
    g.n[d] = sym ? 1 : pg->nx[d] + 2 * pg->ng;
    g.lo[d] = mesh_xmin[d];
    g.hi[d] = mesh_xmax[d];
    g.mid[d] = 0.5 * (mesh_xmin[d] + mesh_xmax[d]);
    g.mesh_vol *= (mesh_xmax[d] - mesh_xmin[d]);
  }
  g.sj = g.n[0];
  g.sk = (int64_t)g.n[0] * g.n[1];
  g.sc = g.sk * g.n[2];
  g.sb = pg->block_stride;
  const int nct = g.nblocks * g.ncomp;
  for (int o = 0; o < 8; ++o) out[o] = 0;
  if (nct == 0) return PB2_OK;
  double *partial = nullptr, *dout = nullptr;
  PB2_CUDA_CHECK(cudaMallocAsync(&partial, sizeof(double) * 8 * (size_t)nct, as_stream(stream)));
  PB2_CUDA_CHECK(cudaMallocAsync(&dout, sizeof(double) * 8, as_stream(stream)));
  ProfScope prof(K_HISTORY, as_stream(stream));
  history_kernel<<<nct, 256, 0, as_stream(stream)>>>(g, u, pg->dx, block_xmin, partial);
  PB2_LAUNCH_CHECK();
  history_final_kernel<<<8, 256, 0, as_stream(stream)>>>(partial, nct, dout);
  PB2_LAUNCH_CHECK();
  PB2_CUDA_CHECK(cudaMemcpyAsync(out, dout, sizeof(double) * 8, cudaMemcpyDeviceToHost,
                                 as_stream(stream)));
  PB2_CUDA_CHECK(cudaFreeAsync(partial, as_stream(stream)));
  PB2_CUDA_CHECK(cudaFreeAsync(dout, as_stream(stream)));
  PB2_CUDA_CHECK(cudaStreamSynchronize(as_stream(stream)));
  return PB2_OK;
}

} // extern "C"
