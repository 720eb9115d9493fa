// sparse_advection_package.hpp — example/sparse_advection package on the B200 hot path:
// NUM_FIELDS sparse fields "sparse_<id>", each advected with its own constant velocity
// (reference example/sparse_advection/sparse_advection_package.hpp).
#pragma once
#include <array>
#include <memory>

#include "pb2/parthenon.hpp"

namespace sparse_advection_package {
using namespace parthenon;

static constexpr int NUM_FIELDS = 4;
using RealArr_t = std::array<Real, NUM_FIELDS>;

std::shared_ptr<StateDescriptor> Initialize(ParameterInput *pin);
// donor-cell fluxes of every allocated (block, field) of the batch
// (sparse_advection_package.cpp:173-258)
TaskStatus CalculateFluxes(MeshData<Real> *md);
Real EstimateTimestepMesh(MeshData<Real> *md); // :136-168
// refinement tags from the allocated fields (:110-143)
void CheckRefinement(MeshData<Real> *md, std::vector<AmrTag> &tags);

} // namespace sparse_advection_package
