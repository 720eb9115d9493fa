// burgers_package.hpp — Parthenon-VIBE (benchmarks/burgers) package on the B200 hot path.
// Same entry points as the reference's benchmarks/burgers/burgers_package.hpp:25-52:
// Initialize, CalculateFluxes, EstimateTimestepMesh, CalculateDerived, MassHistory — plus
// FusedStage, the single-launch-group stage the B200 driver uses by default.
#pragma once
#include <memory>
#include <vector>

#include "pb2/parthenon.hpp"

namespace burgers_package {
using namespace parthenon;

std::shared_ptr<StateDescriptor> Initialize(ParameterInput *pin);
// reference-shaped pieces (one C-ABI launch group each)
TaskStatus CalculateFluxes(MeshData<Real> *md);
void CalculateDerived(MeshData<Real> *md);
Real EstimateTimestepMesh(MeshData<Real> *md);
// the eight octant sums "MS Mass 0..7" in one device pass (burgers_package.cpp:406-439)
std::vector<Real> MassHistory(MeshData<Real> *md);

// CalculateFluxes + FluxDivergence + AverageIndependentData + UpdateIndependentData +
// CalculateDerived (+ EstimateTimestepMesh on the last stage) of burgers_driver.cpp:92-127:
//   mc1.U = (beta*mc0.U + (1-beta)*base.U) - beta*dt*div F(mc0.U)    on interior cells
TaskStatus FusedStage(MeshData<Real> *mc0, MeshData<Real> *mbase, MeshData<Real> *mc1,
                      Real beta, Real dt, bool last_stage);

// last stage of the fused path: read back the min dt the stage kernel reduced and let the
// blocks vote (what Update::EstimateTimestep does after the package hook)
TaskStatus CollectFusedTimestep(MeshData<Real> *mc1);

} // namespace burgers_package
