// sparse_advection_driver.hpp — application driver of example/sparse_advection
// (reference example/sparse_advection/sparse_advection_driver.hpp).
#pragma once
#include "pb2/parthenon.hpp"

namespace sparse_advection_example {
using namespace parthenon::driver::prelude;

class SparseAdvectionDriver : public MultiStageDriver {
 public:
  SparseAdvectionDriver(ParameterInput *pin, ApplicationInput *app_in, Mesh *pm);
  // sparse_advection_driver.cpp:56-149
  TaskCollection MakeTaskCollection(BlockList_t &blocks, int stage) override;
};

void MeshProblemGenerator(parthenon::MeshData<parthenon::Real> *md, ParameterInput *pin);
parthenon::Packages_t ProcessPackages(std::unique_ptr<ParameterInput> &pin);

} // namespace sparse_advection_example
