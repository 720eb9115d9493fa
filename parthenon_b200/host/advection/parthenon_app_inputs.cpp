// parthenon_app_inputs.cpp — initial condition and package list of example/advection
// (reference example/advection/parthenon_app_inputs.cpp:40-100, 275-286).  The profile is
// evaluated on the HOST with libm and uploaded (bit-identical to the reference's CPU build;
// once, outside the hot path); ghosts are filled by the first exchange.
#include <cmath>
#include <vector>

#include "advection_driver.hpp"
#include "advection_package.hpp"

namespace advection_example {
using namespace parthenon;

void MeshProblemGenerator(MeshData<Real> *md, ParameterInput *) {
  auto pkg = md->GetMeshPointer()->packages.Get("advection_package");
  const Real amp = pkg->Param<Real>("amp");
  const std::string profile = pkg->Param<std::string>("profile");
  PARTHENON_REQUIRE_THROWS(profile == "smooth_gaussian" || profile == "hard_sphere",
                           "Advection/profile must be smooth_gaussian or hard_sphere here");
  const bool gaussian = profile == "smooth_gaussian";
  const IndexRange ib = md->GetBoundsI(IndexDomain::interior);
  const IndexRange jb = md->GetBoundsJ(IndexDomain::interior);
  const IndexRange kb = md->GetBoundsK(IndexDomain::interior);
  const int nb = md->NumBlocks();
  for (Variable *pu : md->GetVariablesByFlag({Metadata::Independent})) {
    Variable &u = *pu;
    std::vector<Real> h(static_cast<size_t>(nb) * u.block_stride, 0.0);
#pragma omp parallel for schedule(static)
    for (int b = 0; b < nb; ++b) {
      const auto &coords = md->GetBlock(b)->coords;
      Real *ub = h.data() + static_cast<size_t>(b) * u.block_stride;
      for (int n = 0; n < u.NumComponents(); ++n)
        for (int k = kb.s; k <= kb.e; ++k)
          for (int j = jb.s; j <= jb.e; ++j)
            for (int i = ib.s; i <= ib.e; ++i) {
              const Real rsq = coords.Xc<1>(i) * coords.Xc<1>(i) +
                               coords.Xc<2>(j) * coords.Xc<2>(j) +
                               coords.Xc<3>(k) * coords.Xc<3>(k);
              const size_t cell = (static_cast<size_t>(k) * u.nj + j) * u.ni + i;
              ub[n * u.comp_stride + cell] =
                  gaussian ? 1. + amp * std::exp(-100.0 * rsq) : (rsq < 0.15 * 0.15 ? 1.0 : 0.0);
            }
    }
    PB2_CHECK(pb2_memcpy_h2d(u.data(), h.data(), sizeof(Real) * h.size(), md->stream()));
    PB2_CHECK(pb2_stream_sync(md->stream()));
  }
}

Packages_t ProcessPackages(std::unique_ptr<ParameterInput> &pin) {
  Packages_t packages;
  packages.Add(advection_package::Initialize(pin.get()));
  return packages;
}

} // namespace advection_example
