// parthenon_app_inputs.cpp — initial condition and package list of the burgers benchmark
// (reference benchmarks/burgers/parthenon_app_inputs.cpp:35-87).
//
// The initial condition is evaluated on the HOST with libm (tanh, cos, sin, exp) and
// uploaded, so it is bit-identical to the reference's CPU build; it runs once, outside the
// hot path.  Ghost cells are filled by the first exchange (Mesh::Initialize).
#include <cmath>
#include <vector>

#include "burgers_driver.hpp"
#include "burgers_package.hpp"

namespace burgers_benchmark {
using namespace parthenon;

void MeshProblemGenerator(MeshData<Real> *md, ParameterInput *) {
  Variable &u = md->Get("U");
  const int ncomp = u.NumComponents();
  const IndexRange ib = md->GetBoundsI(IndexDomain::interior);
  const IndexRange jb = md->GetBoundsJ(IndexDomain::interior);
  const IndexRange kb = md->GetBoundsK(IndexDomain::interior);
  const int nb = md->NumBlocks();
  // stage a few blocks at a time so a 512^3 mesh does not need a second copy in host RAM
  const int chunk = std::max(1, std::min(nb, 64));
  std::vector<Real> h(static_cast<size_t>(chunk) * u.block_stride);
  for (int b0 = 0; b0 < nb; b0 += chunk) {
    const int n = std::min(chunk, nb - b0);
    std::fill(h.begin(), h.end(), 0.0);
#pragma omp parallel for schedule(static)
    for (int bb = 0; bb < n; ++bb) {
      const auto &coords = md->GetBlock(b0 + bb)->coords;
      Real *ub = h.data() + static_cast<size_t>(bb) * u.block_stride;
      for (int k = kb.s; k <= kb.e; ++k)
        for (int j = jb.s; j <= jb.e; ++j)
          for (int i = ib.s; i <= ib.e; ++i) {
            const Real x = coords.Xc<1>(i), y = coords.Xc<2>(j), z = coords.Xc<3>(k);
            const size_t cell = (static_cast<size_t>(k) * u.nj + j) * u.ni + i;
            ub[cell] = (std::tanh(-20.0 * x) * std::cos(M_PI * x) + 1.0) *
                       std::exp(-30.0 * y * y) * std::exp(-30.0 * z * z);
            ub[u.comp_stride + cell] = (std::sin(M_PI * y) + 0.2) * std::exp(-30.0 * x * x) *
                                       std::exp(-30.0 * z * z);
            ub[2 * u.comp_stride + cell] = (std::tanh(-20. * z) * std::cos(M_PI * z) + 0.5) *
                                           std::exp(-30.0 * x * x) * std::exp(-30.0 * y * y);
            Real q = 1;
            if (std::abs(x) < 0.025 && std::abs(y) < 0.15 && std::abs(z) < 0.025) q += 10.0;
            for (int c = 3; c < ncomp; ++c) ub[c * u.comp_stride + cell] = q;
          }
    }
    PB2_CHECK(pb2_memcpy_h2d(u.data() + static_cast<int64_t>(b0) * u.block_stride, h.data(),
                             sizeof(Real) * static_cast<size_t>(n) * u.block_stride, md->stream()));
    PB2_CHECK(pb2_stream_sync(md->stream()));
  }
}

Packages_t ProcessPackages(std::unique_ptr<ParameterInput> &pin) {
  Packages_t packages;
  packages.Add(burgers_package::Initialize(pin.get()));
  return packages;
}

} // namespace burgers_benchmark
