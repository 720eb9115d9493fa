// parthenon_app_inputs.cpp — initial condition and package list of example/sparse_advection
// (reference example/sparse_advection/parthenon_app_inputs.cpp:42-105, 187-197): field f is
// allocated on the blocks its initial disc touches and set to 1 inside / 0 outside there.
#include <cmath>
#include <string>
#include <vector>

#include "sparse_advection_driver.hpp"
#include "sparse_advection_package.hpp"

namespace sparse_advection_example {
using namespace parthenon;
using sparse_advection_package::NUM_FIELDS;
using sparse_advection_package::RealArr_t;

void MeshProblemGenerator(MeshData<Real> *md, ParameterInput *) {
  Mesh *pm = md->GetMeshPointer();
  auto pkg = pm->packages.Get("sparse_advection_package");
  const Real r = pkg->Param<Real>("init_size");
  const Real size = r * r;
  const auto &x0s = pkg->Param<RealArr_t>("x0");
  const auto &y0s = pkg->Param<RealArr_t>("y0");
  const IndexRange ib = md->GetBoundsI(IndexDomain::interior);
  const IndexRange jb = md->GetBoundsJ(IndexDomain::interior);
  const IndexRange kb = md->GetBoundsK(IndexDomain::interior);
  for (int f = 0; f < NUM_FIELDS; ++f) {
    const std::string label = "sparse_" + std::to_string(f);
    Variable &u = md->Get(label);
    std::vector<Real> h(static_cast<size_t>(u.block_stride));
    for (int b = 0; b < md->NumBlocks(); ++b) {
      const auto &coords = md->GetBlock(b)->coords;
      bool any_nonzero = false;
      std::fill(h.begin(), h.end(), 0.0);
      for (int k = kb.s; k <= kb.e; ++k)
        for (int j = jb.s; j <= jb.e; ++j)
          for (int i = ib.s; i <= ib.e; ++i) {
            const Real x = coords.Xc<1>(i) - x0s[f], y = coords.Xc<2>(j) - y0s[f],
                       z = coords.Xc<3>(k);
            const Real r2 = x * x + y * y + z * z;
            if (r2 < size) any_nonzero = true;
            h[(static_cast<size_t>(k) * u.nj + j) * u.ni + i] = (r2 < size ? 1.0 : 0.0);
          }
      if (!any_nonzero) continue;
      pm->AllocateSparse(label, md->GetBlock(b)->lid);
      PB2_CHECK(pb2_memcpy_h2d(u.block(b), h.data(), sizeof(Real) * h.size(), md->stream()));
      PB2_CHECK(pb2_stream_sync(md->stream()));
    }
  }
}

Packages_t ProcessPackages(std::unique_ptr<ParameterInput> &pin) {
  Packages_t packages;
  packages.Add(sparse_advection_package::Initialize(pin.get()));
  return packages;
}

} // namespace sparse_advection_example
