// burgers_driver.hpp — application driver of the Parthenon-VIBE benchmark
// (reference benchmarks/burgers/burgers_driver.hpp).
#pragma once
#include "pb2/parthenon.hpp"

namespace burgers_benchmark {
using namespace parthenon::driver::prelude;

class BurgersDriver : public MultiStageDriver {
 public:
  BurgersDriver(ParameterInput *pin, ApplicationInput *app_in, Mesh *pm);
  // one task collection per integrator stage (burgers_driver.cpp:53-148)
  TaskCollection MakeTaskCollection(BlockList_t &blocks, int stage) override;
};

// benchmarks/burgers/parthenon_app_inputs.cpp
void MeshProblemGenerator(parthenon::MeshData<parthenon::Real> *md, ParameterInput *pin);
parthenon::Packages_t ProcessPackages(std::unique_ptr<ParameterInput> &pin);

} // namespace burgers_benchmark
