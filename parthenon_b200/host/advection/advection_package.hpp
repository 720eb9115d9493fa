// advection_package.hpp — example/advection package (constant-velocity branch) on the B200
// hot path.  Same entry points as the reference's example/advection/advection_package.hpp:
// Initialize, CalculateFluxes, EstimateTimestep — batched over a whole MeshData.
#pragma once
#include <memory>
#include <vector>

#include "pb2/parthenon.hpp"

namespace advection_package {
using namespace parthenon;

std::shared_ptr<StateDescriptor> Initialize(ParameterInput *pin);
// donor-cell fluxes of every WithFluxes field of the batch (advection_package.cpp:540-646)
TaskStatus CalculateFluxes(MeshData<Real> *md);
// cfl * min over blocks and directions of dx_d / |v_d| (advection_package.cpp:505-536)
Real EstimateTimestepMesh(MeshData<Real> *md);
// refinement tags of every block of the batch (advection_package.cpp:239-273)
void CheckRefinement(MeshData<Real> *md, std::vector<AmrTag> &tags);

} // namespace advection_package
