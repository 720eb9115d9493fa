// prores.cu — restriction and prolongation at fine-coarse boundaries, cell-centred fields.
//
// Replaces refinement::Restrict / ProlongateShared (reference
// src/prolong_restrict/prolong_restrict.cpp:37-79) and the team loops of
// src/prolong_restrict/pr_loops.hpp:113-155: one launch covers every ProResInfo region of a
// MeshData.  Operators follow src/prolong_restrict/pr_ops.hpp:105-165 (RestrictAverage) and
// :167-280 (ProlongateSharedGeneral) with the same expression order; this file is compiled
// with -fmad=false so results are bit-identical to the reference's CPU build.
#include <vector>

#include "common.cuh"
#include "tables.cuh"

namespace pb2 {

constexpr int kPrThreads = 128;
constexpr int kPrPerThread = 2;

__device__ __forceinline__ double sign_of(double x) { return (x > 0) - (x < 0); }

// util::GradMinMod pr_ops.hpp:95-101
__device__ __forceinline__ double grad_minmod(double fc, double fm, double fp, double dxm,
                                              double dxp, double &gxm, double &gxp) {
  gxm = (fc - fm) / dxm;
  gxp = (fp - fc) / dxp;
  return 0.5 * (sign_of(gxm) + sign_of(gxp)) * fmin(fabs(gxm), fabs(gxp));
}

__device__ __forceinline__ bool cell_of(const pb2_prores_region &r, uint32_t e, int &c, int &k,
                                        int &j, int &i) {
  const uint32_t total = (uint32_t)r.ncomp * r.n[0] * r.n[1] * r.n[2];
  if (e >= total) return false;
  i = e % r.n[0];
  uint32_t t = e / r.n[0];
  j = t % r.n[1];
  t /= r.n[1];
  k = t % r.n[2];
  c = t / r.n[2];
  i += r.s[0];
  j += r.s[1];
  k += r.s[2];
  return true;
}

// RestrictAverage::Do pr_ops.hpp:105-165 (uniform Cartesian: every fine cell has the same
// volume dx1*dx2*dx3, uniform_cartesian.hpp:39; the weighted form is kept for bit parity)
__global__ void __launch_bounds__(kPrThreads)
    restrict_kernel(const pb2_prores_region *__restrict__ regions,
                    const Chunk *__restrict__ chunks) {
  const Chunk ch = chunks[blockIdx.x];
  const pb2_prores_region &r = regions[ch.region];
  if (!(r.status & PB2_REGION_ALLOCATED)) return;
  const int DIM = r.ndim;
  const double cellvol = r.fine_dx[0] * r.fine_dx[1] * r.fine_dx[2];
#pragma unroll
  for (int u = 0; u < kPrPerThread; ++u) {
    int c, ck, cj, ci;
    if (!cell_of(r, ch.first_vec + u * kPrThreads + threadIdx.x, c, ck, cj, ci)) continue;
    const int i = (ci - r.coarse_is[0]) * 2 + r.fine_is[0];
    const int j = DIM > 1 ? (cj - r.coarse_is[1]) * 2 + r.fine_is[1] : r.fine_is[1];
    const int k = DIM > 2 ? (ck - r.coarse_is[2]) * 2 + r.fine_is[2] : r.fine_is[2];
    const double *f = r.fine + (int64_t)c * r.fine_stride_c;
    double vol[2][2][2], terms[2][2][2];
#pragma unroll
    for (int ok = 0; ok < 2; ++ok)
#pragma unroll
      for (int oj = 0; oj < 2; ++oj)
#pragma unroll
        for (int oi = 0; oi < 2; ++oi) {
          const bool on = (ok == 0 || DIM > 2) && (oj == 0 || DIM > 1);
          vol[ok][oj][oi] = on ? cellvol : 0.0;
          terms[ok][oj][oi] =
              on ? vol[ok][oj][oi] * f[(int64_t)(k + ok) * r.fine_stride_k +
                                       (int64_t)(j + oj) * r.fine_stride_j + (i + oi)]
                 : 0.0;
        }
    const double tvol = ((vol[0][0][0] + vol[0][1][0]) + (vol[0][0][1] + vol[0][1][1])) +
                        ((vol[1][0][0] + vol[1][1][0]) + (vol[1][0][1] + vol[1][1][1]));
    r.coarse[(int64_t)c * r.coarse_stride_c + (int64_t)ck * r.coarse_stride_k +
             (int64_t)cj * r.coarse_stride_j + ci] =
        (((terms[0][0][0] + terms[0][1][0]) + (terms[0][0][1] + terms[0][1][1])) +
         ((terms[1][0][0] + terms[1][1][0]) + (terms[1][0][1] + terms[1][1][1]))) /
        tvol;
  }
}

// ProlongateSharedGeneral::Do pr_ops.hpp:167-280, cell centred; GetGridSpacings :76-93 with
// UniformCartesian::X (uniform_cartesian.hpp:166-178)
__global__ void __launch_bounds__(kPrThreads)
    prolongate_kernel(const pb2_prores_region *__restrict__ regions,
                      const Chunk *__restrict__ chunks, int op) {
  const Chunk ch = chunks[blockIdx.x];
  const pb2_prores_region &r = regions[ch.region];
  if (!(r.status & PB2_REGION_ALLOCATED)) return;
  const int DIM = r.ndim;
#pragma unroll
  for (int u = 0; u < kPrPerThread; ++u) {
    int c, k, j, i;
    if (!cell_of(r, ch.first_vec + u * kPrThreads + threadIdx.x, c, k, j, i)) continue;
    const int fi = (i - r.coarse_is[0]) * 2 + r.fine_is[0];
    const int fj = DIM > 1 ? (j - r.coarse_is[1]) * 2 + r.fine_is[1] : r.fine_is[1];
    const int fk = DIM > 2 ? (k - r.coarse_is[2]) * 2 + r.fine_is[2] : r.fine_is[2];
    const double *cs = r.coarse + (int64_t)c * r.coarse_stride_c +
                       (int64_t)k * r.coarse_stride_k + (int64_t)j * r.coarse_stride_j + i;
    const double fc = cs[0];
    double gm[3] = {0, 0, 0}, gp[3] = {0, 0, 0}, dxfm[3] = {0, 0, 0}, dxfp[3] = {0, 0, 0};
    const int cc[3] = {i, j, k}, ff[3] = {fi, fj, fk};
    const int64_t cstr[3] = {1, r.coarse_stride_j, r.coarse_stride_k};
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      if (d < DIM) {
        const double xm = r.coarse_xmin[d] + ((cc[d] - 1) + 0.5) * r.coarse_dx[d];
        const double xc = r.coarse_xmin[d] + (cc[d] + 0.5) * r.coarse_dx[d];
        const double xp = r.coarse_xmin[d] + ((cc[d] + 1) + 0.5) * r.coarse_dx[d];
        const double dxm = xc - xm, dxp = xp - xc;
        const double fxm = r.fine_xmin[d] + (ff[d] + 0.5) * r.fine_dx[d];
        const double fxp = r.fine_xmin[d] + ((ff[d] + 1) + 0.5) * r.fine_dx[d];
        dxfm[d] = xc - fxm;
        dxfp[d] = fxp - xc;
        double gxm, gxp;
        const double gc = grad_minmod(fc, cs[-cstr[d]], cs[cstr[d]], dxm, dxp, gxm, gxp);
        if (op == PB2_PROLONG_MINMOD) {
          gm[d] = gc;
          gp[d] = gc;
        } else if (op == PB2_PROLONG_LINEAR) {
          gm[d] = gxm;
          gp[d] = gxp;
        } // piecewise constant: zero slopes
      }
    }
    const double gx1m = gm[0], gx1p = gp[0], gx2m = gm[1], gx2p = gp[1], gx3m = gm[2],
                 gx3p = gp[2];
    const double dx1fm = dxfm[0], dx1fp = dxfp[0], dx2fm = dxfm[1], dx2fp = dxfp[1],
                 dx3fm = dxfm[2], dx3fp = dxfp[2];
    double *f = r.fine + (int64_t)c * r.fine_stride_c + (int64_t)fk * r.fine_stride_k +
                (int64_t)fj * r.fine_stride_j + fi;
    const int64_t sj = r.fine_stride_j, sk = r.fine_stride_k;
    f[0] = fc - (gx1m * dx1fm + gx2m * dx2fm + gx3m * dx3fm);
    f[1] = fc + (gx1p * dx1fp - gx2m * dx2fm - gx3m * dx3fm);
    if (DIM > 1) {
      f[sj] = fc - (gx1m * dx1fm - gx2p * dx2fp + gx3m * dx3fm);
      f[sj + 1] = fc + (gx1p * dx1fp + gx2p * dx2fp - gx3m * dx3fm);
    }
    if (DIM > 2) {
      f[sk] = fc - (gx1m * dx1fm + gx2m * dx2fm - gx3p * dx3fp);
      f[sk + 1] = fc + (gx1p * dx1fp - gx2m * dx2fm + gx3p * dx3fp);
      f[sk + sj] = fc - (gx1m * dx1fm - gx2p * dx2fp - gx3p * dx3fp);
      f[sk + sj + 1] = fc + (gx1p * dx1fp + gx2p * dx2fp + gx3p * dx3fp);
    }
  }
}


// ---------------------------------------------------------------------------------------
// face / edge / node elements (pb2_prores_region::ftop / ctop): the element forms of the same
// operators.  A region addresses ONE topological element of a field (its `fine` / `coarse`
// pointers start at the element's first slab component); ftop says in which directions the
// element is displaced by half a cell (TopologicalOffsetI/J/K, basic_types.hpp:195-203).
// ---------------------------------------------------------------------------------------

// RestrictAverage::Do<DIM, el> pr_ops.hpp:105-165: average over the directions the element is
// centred in (INCLUDE_Xd), weights Volume<el> (uniform_cartesian.hpp:247-270): cell volume,
// face area, edge length or 1 — the product of dx over the centred directions
__global__ void __launch_bounds__(kPrThreads)
    restrict_te_kernel(const pb2_prores_region *__restrict__ regions,
                       const Chunk *__restrict__ chunks) {
  const Chunk ch = chunks[blockIdx.x];
  const pb2_prores_region &r = regions[ch.region];
  if (!(r.status & PB2_REGION_ALLOCATED)) return;
  const int DIM = r.ndim;
  const bool inc0 = DIM > 0 && !r.ftop[0], inc1 = DIM > 1 && !r.ftop[1],
             inc2 = DIM > 2 && !r.ftop[2];
  double elvol = 1.0;
#pragma unroll
  for (int d = 0; d < 3; ++d)
    if (!r.ftop[d]) elvol = elvol * r.fine_dx[d];
#pragma unroll
  for (int u = 0; u < kPrPerThread; ++u) {
    int c, ck, cj, ci;
    if (!cell_of(r, ch.first_vec + u * kPrThreads + threadIdx.x, c, ck, cj, ci)) continue;
    const int i = (ci - r.coarse_is[0]) * 2 + r.fine_is[0];
    const int j = DIM > 1 ? (cj - r.coarse_is[1]) * 2 + r.fine_is[1] : r.fine_is[1];
    const int k = DIM > 2 ? (ck - r.coarse_is[2]) * 2 + r.fine_is[2] : r.fine_is[2];
    const double *f = r.fine + (int64_t)c * r.fine_stride_c;
    double vol[2][2][2], terms[2][2][2];
#pragma unroll
    for (int ok = 0; ok < 2; ++ok)
#pragma unroll
      for (int oj = 0; oj < 2; ++oj)
#pragma unroll
        for (int oi = 0; oi < 2; ++oi) {
          const bool on = (ok == 0 || inc2) && (oj == 0 || inc1) && (oi == 0 || inc0);
          vol[ok][oj][oi] = on ? elvol : 0.0;
          terms[ok][oj][oi] =
              on ? vol[ok][oj][oi] * f[(int64_t)(k + ok) * r.fine_stride_k +
                                       (int64_t)(j + oj) * r.fine_stride_j + (i + oi)]
                 : 0.0;
        }
    const double tvol = ((vol[0][0][0] + vol[0][1][0]) + (vol[0][0][1] + vol[0][1][1])) +
                        ((vol[1][0][0] + vol[1][1][0]) + (vol[1][0][1] + vol[1][1][1]));
    r.coarse[(int64_t)c * r.coarse_stride_c + (int64_t)ck * r.coarse_stride_k +
             (int64_t)cj * r.coarse_stride_j + ci] =
        (((terms[0][0][0] + terms[0][1][0]) + (terms[0][0][1] + terms[0][1][1])) +
         ((terms[1][0][0] + terms[1][1][0]) + (terms[1][0][1] + terms[1][1][1]))) /
        tvol;
  }
}

// ProlongateSharedGeneral::Do<DIM, el> pr_ops.hpp:167-280: the fine elements that coincide
// with coarse element (k, j, i); slopes exist only in the directions the element is centred in
// (there GetGridSpacings :76-93 uses cell-centre positions)
__global__ void __launch_bounds__(kPrThreads)
    prolongate_te_kernel(const pb2_prores_region *__restrict__ regions,
                         const Chunk *__restrict__ chunks, int op) {
  const Chunk ch = chunks[blockIdx.x];
  const pb2_prores_region &r = regions[ch.region];
  if (!(r.status & PB2_REGION_ALLOCATED)) return;
  const int DIM = r.ndim;
  const bool inc[3] = {DIM > 0 && !r.ftop[0], DIM > 1 && !r.ftop[1], DIM > 2 && !r.ftop[2]};
#pragma unroll
  for (int u = 0; u < kPrPerThread; ++u) {
    int c, k, j, i;
    if (!cell_of(r, ch.first_vec + u * kPrThreads + threadIdx.x, c, k, j, i)) continue;
    const int fi = (i - r.coarse_is[0]) * 2 + r.fine_is[0];
    const int fj = DIM > 1 ? (j - r.coarse_is[1]) * 2 + r.fine_is[1] : r.fine_is[1];
    const int fk = DIM > 2 ? (k - r.coarse_is[2]) * 2 + r.fine_is[2] : r.fine_is[2];
    const double *cs = r.coarse + (int64_t)c * r.coarse_stride_c +
                       (int64_t)k * r.coarse_stride_k + (int64_t)j * r.coarse_stride_j + i;
    const double fc = cs[0];
    double gm[3] = {0, 0, 0}, gp[3] = {0, 0, 0}, dxfm[3] = {0, 0, 0}, dxfp[3] = {0, 0, 0};
    const int cc[3] = {i, j, k}, ff[3] = {fi, fj, fk};
    const int64_t cstr[3] = {1, r.coarse_stride_j, r.coarse_stride_k};
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      if (inc[d]) {
        const double xm = r.coarse_xmin[d] + ((cc[d] - 1) + 0.5) * r.coarse_dx[d];
        const double xc = r.coarse_xmin[d] + (cc[d] + 0.5) * r.coarse_dx[d];
        const double xp = r.coarse_xmin[d] + ((cc[d] + 1) + 0.5) * r.coarse_dx[d];
        const double dxm = xc - xm, dxp = xp - xc;
        const double fxm = r.fine_xmin[d] + (ff[d] + 0.5) * r.fine_dx[d];
        const double fxp = r.fine_xmin[d] + ((ff[d] + 1) + 0.5) * r.fine_dx[d];
        dxfm[d] = xc - fxm;
        dxfp[d] = fxp - xc;
        double gxm, gxp;
        const double gc = grad_minmod(fc, cs[-cstr[d]], cs[cstr[d]], dxm, dxp, gxm, gxp);
        if (op == PB2_PROLONG_MINMOD) {
          gm[d] = gc;
          gp[d] = gc;
        } else if (op == PB2_PROLONG_LINEAR) {
          gm[d] = gxm;
          gp[d] = gxp;
        }
      }
    }
    const double gx1m = gm[0], gx1p = gp[0], gx2m = gm[1], gx2p = gp[1], gx3m = gm[2],
                 gx3p = gp[2];
    const double dx1fm = dxfm[0], dx1fp = dxfp[0], dx2fm = dxfm[1], dx2fp = dxfp[1],
                 dx3fm = dxfm[2], dx3fp = dxfp[2];
    double *f = r.fine + (int64_t)c * r.fine_stride_c + (int64_t)fk * r.fine_stride_k +
                (int64_t)fj * r.fine_stride_j + fi;
    const int64_t sj = r.fine_stride_j, sk = r.fine_stride_k;
    f[0] = fc - (gx1m * dx1fm + gx2m * dx2fm + gx3m * dx3fm);
    if (inc[0]) f[1] = fc + (gx1p * dx1fp - gx2m * dx2fm - gx3m * dx3fm);
    if (inc[1]) f[sj] = fc - (gx1m * dx1fm - gx2p * dx2fp + gx3m * dx3fm);
    if (inc[1] && inc[0]) f[sj + 1] = fc + (gx1p * dx1fp + gx2p * dx2fp - gx3m * dx3fm);
    if (inc[2]) f[sk] = fc - (gx1m * dx1fm + gx2m * dx2fm - gx3p * dx3fp);
    if (inc[2] && inc[0]) f[sk + 1] = fc + (gx1p * dx1fp - gx2m * dx2fm + gx3p * dx3fp);
    if (inc[2] && inc[1]) f[sk + sj] = fc - (gx1m * dx1fm - gx2p * dx2fp - gx3p * dx3fp);
    if (inc[2] && inc[1] && inc[0])
      f[sk + sj + 1] = fc + (gx1p * dx1fp + gx2p * dx2fp + gx3p * dx3fp);
  }
}

// ProlongateInternalAverage::Do<DIM, fel, cel> pr_ops.hpp:291-382: the fine elements of the
// field (ftop) strictly inside coarse container element cel (ctop) at (k, j, i) become the
// average of the fine shared elements around them.  Runs after prolongate_te_kernel of the
// same exchange has completed; reads only shared elements, writes only internal ones.
__global__ void __launch_bounds__(kPrThreads)
    prolongate_internal_kernel(const pb2_prores_region *__restrict__ regions,
                               const Chunk *__restrict__ chunks) {
  const Chunk ch = chunks[blockIdx.x];
  const pb2_prores_region &r = regions[ch.region];
  if (!(r.status & PB2_REGION_ALLOCATED)) return;
  const int DIM = r.ndim;
  int center[3], stencil[3];
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    center[d] = d < DIM && !r.ftop[d];
    stencil[d] = d < DIM && !center[d] && !r.ctop[d];
  }
  const double w = 1.0 / ((1.0 + stencil[2]) * (1.0 + stencil[1]) * (1.0 + stencil[0]));
  const int64_t sj = r.fine_stride_j, sk = r.fine_stride_k;
#pragma unroll
  for (int u = 0; u < kPrPerThread; ++u) {
    int c, k, j, i;
    if (!cell_of(r, ch.first_vec + u * kPrThreads + threadIdx.x, c, k, j, i)) continue;
    const int fi = (i - r.coarse_is[0]) * 2 + r.fine_is[0];
    const int fj = DIM > 1 ? (j - r.coarse_is[1]) * 2 + r.fine_is[1] : r.fine_is[1];
    const int fk = DIM > 2 ? (k - r.coarse_is[2]) * 2 + r.fine_is[2] : r.fine_is[2];
    double *f = r.fine + (int64_t)c * r.fine_stride_c;
    for (int ok = 0; ok < 1 + center[2]; ++ok)
      for (int oj = 0; oj < 1 + center[1]; ++oj)
        for (int oi = 0; oi < 1 + center[0]; ++oi) {
          const int tk = fk + ok + stencil[2], tj = fj + oj + stencil[1],
                    ti = fi + oi + stencil[0];
          double v = 0.0;
          for (int stk = -stencil[2]; stk <= stencil[2]; stk += 2)
            for (int stj = -stencil[1]; stj <= stencil[1]; stj += 2)
              for (int sti = -stencil[0]; sti <= stencil[0]; sti += 2)
                v += w * f[(int64_t)(tk + stk) * sk + (int64_t)(tj + stj) * sj + (ti + sti)];
          f[(int64_t)tk * sk + (int64_t)tj * sj + ti] = v;
        }
  }
}

// ProlongateInternalTothAndRoe::Do<DIM, fel, CC> pr_ops.hpp:384-470: the fine faces of element
// fel inside coarse cell (k, j, i) from the fine faces on the cell's surface, so that the fine
// divergence equals the coarse one (Toth & Roe 2002).  The region's `fine` points at element F1
// of the face field (the operator reads all three), ncomp = tensor components per element;
// ftop names fel.  Written for the x-component, the others by cyclic permutation.
__global__ void __launch_bounds__(kPrThreads)
    prolongate_toth_roe_kernel(const pb2_prores_region *__restrict__ regions,
                               const Chunk *__restrict__ chunks) {
  const Chunk ch = chunks[blockIdx.x];
  const pb2_prores_region &r = regions[ch.region];
  if (!(r.status & PB2_REGION_ALLOCATED)) return;
  const int DIM = r.ndim;
  const int el = r.ftop[0] ? 0 : (r.ftop[1] ? 1 : 2);
  const int g3 = DIM > 2, g2 = DIM > 1;
  const int64_t sj = r.fine_stride_j, sk = r.fine_stride_k;
  const double d1 = r.coarse_dx[el], d2 = r.coarse_dx[(el + 1) % 3],
               d3 = r.coarse_dx[(el + 2) % 3];
  const double dx2 = d1 * d1, dy2 = d2 * d2, dz2 = d3 * d3;
  const double vfac = 0.125 * dz2 / (dx2 + dz2), wfac = 0.125 * dy2 / (dx2 + dy2);
#pragma unroll
  for (int u = 0; u < kPrPerThread; ++u) {
    int c, k, j, i;
    if (!cell_of(r, ch.first_vec + u * kPrThreads + threadIdx.x, c, k, j, i)) continue;
    const int fi = (i - r.coarse_is[0]) * 2 + r.fine_is[0];
    const int fj = DIM > 1 ? (j - r.coarse_is[1]) * 2 + r.fine_is[1] : r.fine_is[1];
    const int fk = DIM > 2 ? (k - r.coarse_is[2]) * 2 + r.fine_is[2] : r.fine_is[2];
    auto fp = [&](int eidx, int ok, int oj, int oi) -> double * {
      double *f = r.fine + ((int64_t)((el + eidx) % 3) * r.ncomp + c) * r.fine_stride_c;
      if (el == 0) return f + (int64_t)(fk + ok * g3) * sk + (int64_t)(fj + oj * g2) * sj + (fi + oi);
      if (el == 1) return f + (int64_t)(fk + oj * g3) * sk + (int64_t)(fj + oi * g2) * sj + (fi + ok);
      return f + (int64_t)(fk + oi * g3) * sk + (int64_t)(fj + ok * g2) * sj + (fi + oj);
    };
    auto sg = [](int o) { return o == 0 ? -1.0 : 1.0; };
    double Uxx = 0.0, Vxyz = 0.0, Wxyz = 0.0;
    for (int v = 0; v <= 1; ++v)
      for (int w = 0; w <= 2; w += 2)
        for (int t = 0; t <= 1; ++t) {
          const double fine2 = *fp(1, v, w, t);
          const double fine3 = *fp(2, w, v, t);
          Uxx += sg(t) * sg(w) * (fine2 + fine3);
          Vxyz += sg(t) * sg(w) * sg(v) * fine2;
          Wxyz += sg(t) * sg(w) * sg(v) * fine3;
        }
    Uxx *= 0.125;
    Vxyz *= vfac;
    Wxyz *= wfac;
    for (int ok = 0; ok <= 1; ++ok)
      for (int oj = 0; oj <= 1; ++oj)
        *fp(0, ok, oj, 1) =
            0.5 * (*fp(0, ok, oj, 0) + *fp(0, ok, oj, 2)) + Uxx + sg(ok) * Vxyz + sg(oj) * Wxyz;
  }
}

// Flux correction: RestrictAverage::Do pr_ops.hpp:105-165 with el = F_dir over the fine faces
// tiling one coarse face, delivered straight to the coarser block's flux array (or a slab).
// Weights are coords.Volume<F_dir> (the fine block's face area, uniform_cartesian.hpp:36-38);
// the summation tree :155-162 is kept term for term (absent children contribute +0).
__global__ void __launch_bounds__(kPrThreads)
    flux_correct_kernel(const pb2_flxcor_region *__restrict__ regions,
                        const Chunk *__restrict__ chunks, double *__restrict__ slab) {
  const Chunk ch = chunks[blockIdx.x];
  const pb2_flxcor_region &r = regions[ch.region];
  if (!(r.status & PB2_REGION_ALLOCATED)) return;
  const int DIM = r.ndim;
  const bool inc0 = r.dir != 0, inc1 = DIM > 1 && r.dir != 1, inc2 = DIM > 2 && r.dir != 2;
  const uint32_t total = (uint32_t)r.ncomp * r.n[0] * r.n[1] * r.n[2];
#pragma unroll
  for (int u = 0; u < kPrPerThread; ++u) {
    const uint32_t e = ch.first_vec + u * kPrThreads + threadIdx.x;
    if (e >= total) continue;
    const int ci = e % r.n[0];
    uint32_t t = e / r.n[0];
    const int cj = t % r.n[1];
    t /= r.n[1];
    const int ck = t % r.n[2];
    const int c = t / r.n[2];
    // the box is one face thick along dir: tangential coarse steps are two fine faces
    const int i = r.fs[0] + (inc0 ? 2 * ci : 0);
    const int j = r.fs[1] + (inc1 ? 2 * cj : 0);
    const int k = r.fs[2] + (inc2 ? 2 * ck : 0);
    const double *f = r.fine + (int64_t)c * r.fine_stride_c;
    double vol[2][2][2], terms[2][2][2];
#pragma unroll
    for (int ok = 0; ok < 2; ++ok)
#pragma unroll
      for (int oj = 0; oj < 2; ++oj)
#pragma unroll
        for (int oi = 0; oi < 2; ++oi) {
          const bool on = (ok == 0 || inc2) && (oj == 0 || inc1) && (oi == 0 || inc0);
          vol[ok][oj][oi] = on ? r.area : 0.0;
          terms[ok][oj][oi] =
              on ? vol[ok][oj][oi] * f[(int64_t)(k + ok) * r.fine_stride_k +
                                       (int64_t)(j + oj) * r.fine_stride_j + (i + oi)]
                 : 0.0;
        }
    const double tvol = ((vol[0][0][0] + vol[0][1][0]) + (vol[0][0][1] + vol[0][1][1])) +
                        ((vol[1][0][0] + vol[1][1][0]) + (vol[1][0][1] + vol[1][1][1]));
    const double val =
        (((terms[0][0][0] + terms[0][1][0]) + (terms[0][0][1] + terms[0][1][1])) +
         ((terms[1][0][0] + terms[1][1][0]) + (terms[1][0][1] + terms[1][1][1]))) /
        tvol;
    if (r.coarse)
      r.coarse[(int64_t)c * r.coarse_stride_c + (int64_t)(r.ds[2] + ck) * r.coarse_stride_k +
               (int64_t)(r.ds[1] + cj) * r.coarse_stride_j + (r.ds[0] + ci)] = val;
    else
      slab[r.buf_off + e] = val;
  }
}

// GenericBC<DIR, SIDE, Outflow|Reflect> boundary_conditions_generic.hpp:174-268 for cell-centred
// fields, every listed (block, face) in one launch.  The ghost slab of a face covers the whole
// extents of the other two directions.
__global__ void __launch_bounds__(kPrThreads)
    apply_bc_kernel(const pb2_bc_region *__restrict__ regions, const Chunk *__restrict__ chunks) {
  const Chunk ch = chunks[blockIdx.x];
  const pb2_bc_region &r = regions[ch.region];
  const int d = r.face >> 1;
  const bool inner = (r.face & 1) == 0;
  const int ref = inner ? r.is : r.ie;
  const int offset = 2 * ref + (inner ? -1 : 1); // reflections
  int ext[3] = {r.n[0], r.n[1], r.n[2]};
  const int lo = inner ? 0 : r.ie + 1;
  ext[d] = inner ? r.is : r.n[d] - (r.ie + 1);
  const uint32_t total = (uint32_t)r.ncomp * ext[0] * ext[1] * ext[2];
  const int64_t sj = r.stride_j ? r.stride_j : r.n[0],
                sk = r.stride_k ? r.stride_k : (int64_t)r.n[0] * r.n[1];
#pragma unroll
  for (int u = 0; u < kPrPerThread; ++u) {
    const uint32_t e = ch.first_vec + u * kPrThreads + threadIdx.x;
    if (e >= total) continue;
    int idx[3];
    idx[0] = e % ext[0];
    uint32_t t = e / ext[0];
    idx[1] = t % ext[1];
    t /= ext[1];
    idx[2] = t % ext[2];
    const int c = t / ext[2];
    idx[d] += lo;
    int src[3] = {idx[0], idx[1], idx[2]};
    src[d] = r.type == PB2_BC_REFLECT ? offset - idx[d] : ref;
    double *f = r.var + (int64_t)c * r.stride_c;
    const double v = f[src[2] * sk + src[1] * sj + src[0]];
    const bool flip = r.type == PB2_BC_REFLECT && ((r.flip_mask >> c) & 1u);
    f[idx[2] * sk + idx[1] * sj + idx[0]] = flip ? -1.0 * v : 1.0 * v;
  }
}

} // namespace pb2

using namespace pb2;

extern "C" {

int pb2_prores_table_create(pb2_bnd_table **table, const pb2_prores_region *regions,
                            int64_t n) {
  PB2_REQUIRE(table && (regions || n == 0) && n >= 0, "bad arguments");
  if (int rc = require_device()) return rc;
  std::vector<Chunk> chunks;
  int64_t elements = 0;
  for (int64_t r = 0; r < n; ++r) {
    const int64_t total =
        (int64_t)regions[r].ncomp * regions[r].n[0] * regions[r].n[1] * regions[r].n[2];
    PB2_REQUIRE(total >= 0 && total < (1ll << 31), "bad region extent");
    elements += total;
    for (int64_t v = 0; v < total; v += kPrThreads * kPrPerThread)
      chunks.push_back(Chunk{static_cast<int32_t>(r), static_cast<uint32_t>(v)});
  }
  auto *t = new pb2_bnd_table();
  t->kind = kProRes;
  t->nregions = n;
  t->nchunks = static_cast<int64_t>(chunks.size());
  t->elements = elements;
  t->d_regions = nullptr;
  t->d_chunks = nullptr;
  t->d_prores = nullptr;
  if (n > 0) {
    cudaError_t e = table_alloc(reinterpret_cast<void **>(&t->d_prores), n * sizeof(pb2_prores_region));
    if (e == cudaSuccess) e = table_alloc(reinterpret_cast<void **>(&t->d_chunks), (chunks.size() + 1) * sizeof(Chunk));
    if (e == cudaSuccess)
      e = cudaMemcpy(t->d_prores, regions, n * sizeof(pb2_prores_region),
                     cudaMemcpyHostToDevice);
    if (e == cudaSuccess && !chunks.empty())
      e = cudaMemcpy(t->d_chunks, chunks.data(), chunks.size() * sizeof(Chunk),
                     cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
      set_error("prores table upload failed: %s", cudaGetErrorString(e));
      table_free(t->d_prores);
      table_free(t->d_chunks);
      delete t;
      return PB2_ERR_CUDA;
    }
  }
  *table = t;
  return PB2_OK;
}

int pb2_restrict(const pb2_bnd_table *table, pb2_stream_t stream) {
  PB2_REQUIRE(table && table->kind == kProRes, "restrict needs a prores table");
  if (table->nchunks == 0) return PB2_OK;
  ProfScope prof(K_RESTRICT, as_stream(stream), static_cast<double>(table->elements));
  restrict_kernel<<<static_cast<unsigned>(table->nchunks), kPrThreads, 0,
                    as_stream(stream)>>>(table->d_prores, table->d_chunks);
  PB2_LAUNCH_CHECK();
  return PB2_OK;
}

int pb2_prolongate(const pb2_bnd_table *table, int op, pb2_stream_t stream) {
  PB2_REQUIRE(table && table->kind == kProRes, "prolongate needs a prores table");
  PB2_REQUIRE(op >= 0 && op <= 2, "unknown prolongation operator");
  if (table->nchunks == 0) return PB2_OK;
  ProfScope prof(K_PROLONGATE, as_stream(stream), static_cast<double>(table->elements));
  prolongate_kernel<<<static_cast<unsigned>(table->nchunks), kPrThreads, 0,
                      as_stream(stream)>>>(table->d_prores, table->d_chunks, op);
  PB2_LAUNCH_CHECK();
  return PB2_OK;
}

int pb2_restrict_te(const pb2_bnd_table *table, pb2_stream_t stream) {
  PB2_REQUIRE(table && table->kind == kProRes, "restrict needs a prores table");
  if (table->nchunks == 0) return PB2_OK;
  ProfScope prof(K_RESTRICT, as_stream(stream), static_cast<double>(table->elements));
  restrict_te_kernel<<<static_cast<unsigned>(table->nchunks), kPrThreads, 0,
                       as_stream(stream)>>>(table->d_prores, table->d_chunks);
  PB2_LAUNCH_CHECK();
  return PB2_OK;
}

int pb2_prolongate_te(const pb2_bnd_table *table, int op, pb2_stream_t stream) {
  PB2_REQUIRE(table && table->kind == kProRes, "prolongate needs a prores table");
  PB2_REQUIRE(op >= 0 && op <= 2, "unknown prolongation operator");
  if (table->nchunks == 0) return PB2_OK;
  ProfScope prof(K_PROLONGATE, as_stream(stream), static_cast<double>(table->elements));
  prolongate_te_kernel<<<static_cast<unsigned>(table->nchunks), kPrThreads, 0,
                         as_stream(stream)>>>(table->d_prores, table->d_chunks, op);
  PB2_LAUNCH_CHECK();
  return PB2_OK;
}

int pb2_prolongate_internal(const pb2_bnd_table *table, pb2_stream_t stream) {
  PB2_REQUIRE(table && table->kind == kProRes, "prolongate needs a prores table");
  if (table->nchunks == 0) return PB2_OK;
  ProfScope prof(K_PROLONGATE, as_stream(stream), static_cast<double>(table->elements));
  prolongate_internal_kernel<<<static_cast<unsigned>(table->nchunks), kPrThreads, 0,
                               as_stream(stream)>>>(table->d_prores, table->d_chunks);
  PB2_LAUNCH_CHECK();
  return PB2_OK;
}

int pb2_prolongate_toth_roe(const pb2_bnd_table *table, pb2_stream_t stream) {
  PB2_REQUIRE(table && table->kind == kProRes, "prolongate needs a prores table");
  if (table->nchunks == 0) return PB2_OK;
  ProfScope prof(K_PROLONGATE, as_stream(stream), static_cast<double>(table->elements));
  prolongate_toth_roe_kernel<<<static_cast<unsigned>(table->nchunks), kPrThreads, 0,
                               as_stream(stream)>>>(table->d_prores, table->d_chunks);
  PB2_LAUNCH_CHECK();
  return PB2_OK;
}

int pb2_flxcor_table_create(pb2_bnd_table **table, const pb2_flxcor_region *regions,
                            int64_t n) {
  PB2_REQUIRE(table && (regions || n == 0) && n >= 0, "bad arguments");
  if (int rc = require_device()) return rc;
  std::vector<Chunk> chunks;
  int64_t elements = 0;
  for (int64_t r = 0; r < n; ++r) {
    const pb2_flxcor_region &q = regions[r];
    PB2_REQUIRE(q.dir >= 0 && q.dir < 3 && q.dir < q.ndim && q.n[q.dir] == 1,
                "a flux-correction box is one face thick along its normal");
    const int64_t total = (int64_t)q.ncomp * q.n[0] * q.n[1] * q.n[2];
    PB2_REQUIRE(total >= 0 && total < (1ll << 31), "bad region extent");
    elements += total;
    for (int64_t v = 0; v < total; v += kPrThreads * kPrPerThread)
      chunks.push_back(Chunk{static_cast<int32_t>(r), static_cast<uint32_t>(v)});
  }
  auto *t = new pb2_bnd_table();
  t->kind = kFlxCor;
  t->nregions = n;
  t->nchunks = static_cast<int64_t>(chunks.size());
  t->elements = elements;
  t->d_regions = nullptr;
  t->d_chunks = nullptr;
  t->d_prores = nullptr;
  t->d_flxcor = nullptr;
  if (n > 0) {
    cudaError_t e = table_alloc(reinterpret_cast<void **>(&t->d_flxcor), n * sizeof(pb2_flxcor_region));
    if (e == cudaSuccess) e = table_alloc(reinterpret_cast<void **>(&t->d_chunks), (chunks.size() + 1) * sizeof(Chunk));
    if (e == cudaSuccess)
      e = cudaMemcpy(t->d_flxcor, regions, n * sizeof(pb2_flxcor_region),
                     cudaMemcpyHostToDevice);
    if (e == cudaSuccess && !chunks.empty())
      e = cudaMemcpy(t->d_chunks, chunks.data(), chunks.size() * sizeof(Chunk),
                     cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
      set_error("flux-correction table upload failed: %s", cudaGetErrorString(e));
      table_free(t->d_flxcor);
      table_free(t->d_chunks);
      delete t;
      return PB2_ERR_CUDA;
    }
  }
  *table = t;
  return PB2_OK;
}

int pb2_flux_correct(const pb2_bnd_table *table, double *slab, pb2_stream_t stream) {
  PB2_REQUIRE(table && table->kind == kFlxCor, "flux correction needs a flxcor table");
  if (table->nchunks == 0) return PB2_OK;
  ProfScope prof(K_FLUX_CORRECT, as_stream(stream), static_cast<double>(table->elements));
  flux_correct_kernel<<<static_cast<unsigned>(table->nchunks), kPrThreads, 0,
                        as_stream(stream)>>>(table->d_flxcor, table->d_chunks, slab);
  PB2_LAUNCH_CHECK();
  return PB2_OK;
}

int pb2_bc_table_create(pb2_bnd_table **table, const pb2_bc_region *regions, int64_t n) {
  PB2_REQUIRE(table && (regions || n == 0) && n >= 0, "bad arguments");
  if (int rc = require_device()) return rc;
  std::vector<Chunk> chunks;
  int64_t elements = 0;
  for (int64_t r = 0; r < n; ++r) {
    const pb2_bc_region &q = regions[r];
    PB2_REQUIRE(q.face >= 0 && q.face < 6, "face must be 0..5");
    PB2_REQUIRE(q.type == PB2_BC_OUTFLOW || q.type == PB2_BC_REFLECT, "unknown boundary type");
    PB2_REQUIRE(q.ncomp >= 0 && q.ncomp <= 32, "at most 32 components per boundary region");
    const int d = q.face >> 1;
    const int depth = (q.face & 1) == 0 ? q.is : q.n[d] - (q.ie + 1);
    PB2_REQUIRE(depth >= 0 && (q.type != PB2_BC_REFLECT || depth <= q.ie - q.is + 1),
                "ghost slab deeper than the interior it mirrors");
    int64_t total = (int64_t)q.ncomp * depth;
    for (int o = 0; o < 3; ++o)
      if (o != d) total *= q.n[o];
    PB2_REQUIRE(total >= 0 && total < (1ll << 31), "bad region extent");
    elements += total;
    for (int64_t v = 0; v < total; v += kPrThreads * kPrPerThread)
      chunks.push_back(Chunk{static_cast<int32_t>(r), static_cast<uint32_t>(v)});
  }
  auto *t = new pb2_bnd_table();
  t->kind = kBc;
  t->nregions = n;
  t->nchunks = static_cast<int64_t>(chunks.size());
  t->elements = elements;
  t->d_regions = nullptr;
  t->d_chunks = nullptr;
  t->d_prores = nullptr;
  if (n > 0) {
    cudaError_t e = table_alloc(reinterpret_cast<void **>(&t->d_bc), n * sizeof(pb2_bc_region));
    if (e == cudaSuccess) e = table_alloc(reinterpret_cast<void **>(&t->d_chunks), (chunks.size() + 1) * sizeof(Chunk));
    if (e == cudaSuccess)
      e = cudaMemcpy(t->d_bc, regions, n * sizeof(pb2_bc_region), cudaMemcpyHostToDevice);
    if (e == cudaSuccess && !chunks.empty())
      e = cudaMemcpy(t->d_chunks, chunks.data(), chunks.size() * sizeof(Chunk),
                     cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
      set_error("boundary-condition table upload failed: %s", cudaGetErrorString(e));
      table_free(t->d_bc);
      table_free(t->d_chunks);
      delete t;
      return PB2_ERR_CUDA;
    }
  }
  *table = t;
  return PB2_OK;
}

int pb2_apply_bcs(const pb2_bnd_table *table, pb2_stream_t stream) {
  PB2_REQUIRE(table && table->kind == kBc, "apply_bcs needs a boundary-condition table");
  if (table->nchunks == 0) return PB2_OK;
  ProfScope prof(K_APPLY_BC, as_stream(stream), static_cast<double>(table->elements));
  apply_bc_kernel<<<static_cast<unsigned>(table->nchunks), kPrThreads, 0, as_stream(stream)>>>(
      table->d_bc, table->d_chunks);
  PB2_LAUNCH_CHECK();
  return PB2_OK;
}

} // extern "C"
