// sparse_advection_package.cpp — see sparse_advection_package.hpp.
#include "sparse_advection_package.hpp"

#include <algorithm>
#include <cmath>
#include <limits>

namespace sparse_advection_package {

std::shared_ptr<StateDescriptor> Initialize(ParameterInput *pin) {
  auto pkg = std::make_shared<StateDescriptor>("sparse_advection_package");

  PARTHENON_REQUIRE_THROWS(!pin->GetOrAddBoolean("sparse_advection", "restart_test", false),
                           "sparse_advection/restart_test is not supported");
  pkg->AddParam("cfl", pin->GetOrAddReal("sparse_advection", "cfl", 0.45));
  pkg->AddParam("refine_tol", pin->GetOrAddReal("sparse_advection", "refine_tol", 0.3));
  pkg->AddParam("derefine_tol", pin->GetOrAddReal("sparse_advection", "derefine_tol", 0.03));
  pkg->AddParam("init_size", pin->GetOrAddReal("sparse_advection", "init_size", 0.1));

  // starting positions (sparse_advection_package.cpp:60-63)
  const Real pos = 0.8;
  pkg->AddParam("x0", RealArr_t{pos, -pos, -pos, pos});
  pkg->AddParam("y0", RealArr_t{pos, pos, -pos, -pos});
  // field 0 moves in (-1,-1) direction, 1 in (1,-1), 2 in (1,1) and 3 in (-1,1) (:65-70)
  const Real speed = pin->GetOrAddReal("sparse_advection", "speed", 1.0) / std::sqrt(2.0);
  pkg->AddParam("vx", RealArr_t{-speed, speed, speed, -speed});
  pkg->AddParam("vy", RealArr_t{-speed, -speed, speed, speed});
  pkg->AddParam("vz", RealArr_t{0.0, 0.0, 0.0, 0.0});

  Metadata m({Metadata::Cell, Metadata::Independent, Metadata::WithFluxes, Metadata::FillGhost,
              Metadata::Sparse});
  std::vector<int> ids;
  for (int sid = 0; sid < NUM_FIELDS; ++sid) ids.push_back(sid);
  pkg->AddSparsePool("sparse", m, ids);

  pkg->EstimateTimestepMesh = EstimateTimestepMesh;
  pkg->CheckRefinementMesh = CheckRefinement;
  return pkg;
}

// CheckRefinement sparse_advection_package.cpp:110-143: minimum and maximum over the ALLOCATED
// fields of a block (entire extents); a block on which nothing is allocated keeps the reduction's
// identity (max = lowest) and is tagged for derefinement.  One device reduction per field for the
// whole batch instead of one par_reduce per block.
void CheckRefinement(MeshData<Real> *md, std::vector<AmrTag> &tags) {
  auto pkg = md->GetMeshPointer()->packages.Get("sparse_advection_package");
  const Real refine_tol = pkg->Param<Real>("refine_tol");
  const Real derefine_tol = pkg->Param<Real>("derefine_tol");
  const int nb = md->NumBlocks();
  std::vector<Real> mn(nb, std::numeric_limits<Real>::max()),
      mx(nb, std::numeric_limits<Real>::lowest());
  DeviceBuffer dev;
  dev.Allocate(sizeof(Real) * 2 * nb, md->stream());
  std::vector<Real> h(2 * nb);
  for (Variable *u : md->GetVariablesByFlag({Metadata::Sparse})) {
    if (u->label().compare(0, 6, "sparse") != 0) continue;
    const pb2_pack_geom g = md->Geometry(*u);
    PB2_CHECK(pb2_block_minmax(&g, u->data(), u->DeviceMask(), dev.get<Real>(), md->stream()));
    PB2_CHECK(pb2_memcpy_d2h(h.data(), dev.get(), sizeof(Real) * h.size(), md->stream()));
    PB2_CHECK(pb2_stream_sync(md->stream()));
    for (int b = 0; b < nb; ++b) {
      if (!u->IsAllocated(b)) continue; // (the kernel leaves masked blocks untouched)
      mn[b] = std::min(mn[b], h[2 * b]);
      mx[b] = std::max(mx[b], h[2 * b + 1]);
    }
  }
  for (int b = 0; b < nb; ++b) {
    if (mx[b] > refine_tol && mn[b] < derefine_tol)
      tags[b] = AmrTag::refine;
    else if (mx[b] < derefine_tol)
      tags[b] = AmrTag::derefine;
    else
      tags[b] = AmrTag::same;
  }
}

TaskStatus CalculateFluxes(MeshData<Real> *md) {
  auto pkg = md->GetMeshPointer()->packages.Get("sparse_advection_package");
  const auto &vx = pkg->Param<RealArr_t>("vx");
  const auto &vy = pkg->Param<RealArr_t>("vy");
  const auto &vz = pkg->Param<RealArr_t>("vz");
  // The reference stops here in 3-D ("Sparse Advection example must be 2D", :256-257) because it
  // never wrote the x3 flux loop.  BASELINE.json's config 4 names a 3-D shape, so the x3 flux
  // (same donor-cell formula, vz = 0 as registered above) is computed as well; 3-D runs have
  // no reference to be compared with, only the CPU oracle.
  PARTHENON_REQUIRE_THROWS(md->GetMeshPointer()->ndim >= 2, "Sparse Advection needs 2 or 3 dimensions");
  for (Variable *u : md->GetVariablesByFlag({Metadata::WithFluxes})) {
    const int f = u->sparse_id() % NUM_FIELDS;
    const double v[3] = {vx[f], vy[f], vz[f]};
    const pb2_pack_geom g = md->Geometry(*u);
    double *flux[3] = {u->flux(1), u->flux(2), g.ndim > 2 ? u->flux(3) : nullptr};
    PB2_CHECK(pb2_advection_fluxes_blocks(&g, u->data(), flux, v, u->DeviceMask(), md->stream()));
  }
  return TaskStatus::complete;
}

Real EstimateTimestepMesh(MeshData<Real> *md) {
  auto pkg = md->GetMeshPointer()->packages.Get("sparse_advection_package");
  const Real cfl = pkg->Param<Real>("cfl");
  const auto &vx = pkg->Param<RealArr_t>("vx");
  const auto &vy = pkg->Param<RealArr_t>("vy");
  const auto &vz = pkg->Param<RealArr_t>("vz");
  // constant velocities: the reference's per-cell reduction collapses to the cell widths;
  // every field votes, allocated or not (:153-163)
  Real dt_min = std::numeric_limits<Real>::max();
  for (auto &pmb : md->GetBlockList()) {
    const auto dx = pmb->coords.Dx();
    Real min_dt = std::numeric_limits<Real>::max();
    for (int v = 0; v < NUM_FIELDS; ++v) {
      if (vx[v] != 0.0) min_dt = std::min(min_dt, dx[0] / std::abs(vx[v]));
      if (vy[v] != 0.0) min_dt = std::min(min_dt, dx[1] / std::abs(vy[v]));
      if (vz[v] != 0.0) min_dt = std::min(min_dt, dx[2] / std::abs(vz[v]));
    }
    dt_min = std::min(dt_min, cfl * min_dt);
  }
  return dt_min;
}

} // namespace sparse_advection_package
