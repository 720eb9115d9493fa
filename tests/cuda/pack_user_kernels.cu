// pack_user_kernels.cu — "downstream application" kernels written against the device pack header
// (include/parthenon_b200_pack.h) only: what a Parthenon user kernel that indexes
// pack(b, n, k, j, i) looks like on this framework.  Compiled by tests/test_sparse_pack_gpu.py with
// nvcc; mirrors the checks of the reference's tst/unit/test_sparse_pack.cpp:47-330.
#include <cuda_runtime.h>

#include "parthenon_b200_pack.h"

using pb2::PackIdx;
using pb2::SparsePackView;

// value the tests put into component c of variable v on block b (test_sparse_pack.cpp:170)
__host__ __device__ inline double pattern(int b, int v, int c, int k, int j, int i) {
  return i + 1e1 * j + 1e2 * k + 1e4 * c + 1e5 * v + 1e3 * b;
}

// counts mismatches of variable `var` (descriptor index) against the pattern with id `vid`,
// through the two accessors of test_sparse_pack.cpp:223-245: pack(b, lo + c, ...) and
// pack(b, PackIdx + c, ...)
__global__ void check_var_kernel(const SparsePackView pack, const int var, const int vid,
                                 int *nwrong, int *nseen) {
  const int b = blockIdx.x;
  const PackIdx iv(var);
  const int lo = pack.GetLowerBound(b, iv), hi = pack.GetUpperBound(b, iv);
  const int ncell = pack.ni * pack.nj * pack.nk;
  int wrong = 0, seen = 0;
  for (int t = threadIdx.x; t < ncell; t += blockDim.x) {
    const int i = t % pack.ni, j = (t / pack.ni) % pack.nj, k = t / (pack.ni * pack.nj);
    for (int c = 0; c <= hi - lo; ++c) {
      const double n = pattern(b, vid, c, k, j, i);
      if (n != pack(b, lo + c, k, j, i)) ++wrong;
      if (n != pack(b, iv + c, k, j, i)) ++wrong;
      ++seen;
    }
  }
  atomicAdd(nwrong, wrong);
  atomicAdd(nseen, seen);
}

// a flattened pack: one unified index over every (block, component) (test_sparse_pack.cpp:290-306)
__global__ void check_flat_kernel(const SparsePackView pack, int *nwrong) {
  const int v = blockIdx.x;
  const int ncell = pack.ni * pack.nj * pack.nk;
  int wrong = 0;
  for (int t = threadIdx.x; t < ncell; t += blockDim.x) {
    const int i = t % pack.ni, j = (t / pack.ni) % pack.nj, k = t / (pack.ni * pack.nj);
    const int n = i + 10 * j + 100 * k;
    if (n != static_cast<int>(pack(v, k, j, i)) % 1000) ++wrong;
  }
  atomicAdd(nwrong, wrong);
}

// a user update written like a Parthenon package task: for every allocated component of the pack
//   u(b, n) <- u(b, n) - dt / dx1 * (flux1(b, n, i+1) - flux1(b, n, i))   over the interior,
// skipping what is not allocated (Contains / bounds), using GetCoordinates for the cell width
__global__ void flux_update_kernel(const SparsePackView pack, const double dt) {
  const int b = blockIdx.x;
  if (!pack.Contains(b)) return;
  const double dx = pack.GetCoordinates(b).dx[0];
  const int nxi = pack.ie - pack.is + 1, nxj = pack.je - pack.js + 1, nxk = pack.ke - pack.ks + 1;
  for (int n = pack.GetLowerBound(b); n <= pack.GetUpperBound(b); ++n) {
    for (int t = threadIdx.x; t < nxi * nxj * nxk; t += blockDim.x) {
      const int i = pack.is + t % nxi, j = pack.js + (t / nxi) % nxj, k = pack.ks + t / (nxi * nxj);
      pack(b, n, k, j, i) -= dt / dx * (pack.flux(b, 1, n, k, j, i + 1) - pack.flux(b, 1, n, k, j, i));
    }
  }
}

extern "C" {

int pack_check_var(pb2_sparse_pack pod, int var, int vid, int *nwrong_out, int *nseen_out) {
  int *d = nullptr;
  if (cudaMalloc(&d, 2 * sizeof(int)) != cudaSuccess) return 1;
  cudaMemset(d, 0, 2 * sizeof(int));
  check_var_kernel<<<pod.nblocks, 128>>>(SparsePackView(pod), var, vid, d, d + 1);
  int h[2] = {-1, -1};
  const cudaError_t e = cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  cudaFree(d);
  *nwrong_out = h[0];
  *nseen_out = h[1];
  return e == cudaSuccess ? 0 : 2;
}

int pack_check_flat(pb2_sparse_pack pod, int *nwrong_out) {
  int *d = nullptr;
  if (cudaMalloc(&d, sizeof(int)) != cudaSuccess) return 1;
  cudaMemset(d, 0, sizeof(int));
  check_flat_kernel<<<pod.maxvars, 128>>>(SparsePackView(pod), d);
  int h = -1;
  const cudaError_t e = cudaMemcpy(&h, d, sizeof(h), cudaMemcpyDeviceToHost);
  cudaFree(d);
  *nwrong_out = h;
  return e == cudaSuccess ? 0 : 2;
}

int pack_flux_update(pb2_sparse_pack pod, double dt) {
  flux_update_kernel<<<pod.nblocks, 128>>>(SparsePackView(pod), dt);
  return cudaDeviceSynchronize() == cudaSuccess ? 0 : 2;
}

} // extern "C"
