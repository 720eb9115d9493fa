// advection_driver.hpp — application driver of example/advection
// (reference example/advection/advection_driver.hpp).
#pragma once
#include "pb2/parthenon.hpp"

namespace advection_example {
using namespace parthenon::driver::prelude;

class AdvectionDriver : public MultiStageDriver {
 public:
  AdvectionDriver(ParameterInput *pin, ApplicationInput *app_in, Mesh *pm);
  // one task collection per integrator stage (advection_driver.cpp:56-163)
  TaskCollection MakeTaskCollection(BlockList_t &blocks, int stage) override;
};

// example/advection/parthenon_app_inputs.cpp
void MeshProblemGenerator(parthenon::MeshData<parthenon::Real> *md, ParameterInput *pin);
parthenon::Packages_t ProcessPackages(std::unique_ptr<ParameterInput> &pin);

} // namespace advection_example
