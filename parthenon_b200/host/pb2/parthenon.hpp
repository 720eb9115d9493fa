// parthenon.hpp — umbrella header of the host library (what the reference spreads over
// <parthenon/driver.hpp> and <parthenon/package.hpp>).
#pragma once
#include "bvals.hpp"
#include "driver.hpp"
#include "mesh.hpp"
#include "mesh_data.hpp"
#include "parameter_input.hpp"
#include "state.hpp"
#include "tasks.hpp"
#include "types.hpp"
#include "update.hpp"

namespace parthenon {
namespace package {
namespace prelude {
using ::parthenon::Metadata;
using ::parthenon::MeshData;
using ::parthenon::Packages_t;
using ::parthenon::ParameterInput;
using ::parthenon::Real;
using ::parthenon::StateDescriptor;
using ::parthenon::TaskStatus;
} // namespace prelude
} // namespace package
namespace driver {
namespace prelude {
using namespace ::parthenon::package::prelude;
using ::parthenon::ApplicationInput;
using ::parthenon::BlockList_t;
using ::parthenon::DriverStatus;
using ::parthenon::Mesh;
using ::parthenon::MeshBlock;
using ::parthenon::MultiStageDriver;
using ::parthenon::TaskCollection;
using ::parthenon::TaskID;
using ::parthenon::TaskList;
using ::parthenon::TaskListStatus;
using ::parthenon::TaskRegion;
} // namespace prelude
} // namespace driver

// parthenon_manager.hpp: parse the deck, build packages and mesh, run the problem generator
class ParthenonManager {
 public:
  ParthenonManager() { app_input = std::make_unique<ApplicationInput>(); }
  ~ParthenonManager();
  enum class ParthenonStatus { ok, complete, error };
  // argv: -i <deck> [block/key=value ...]; rank/nranks/nccl_id describe the GPU job
  ParthenonStatus ParthenonInitEnv(int argc, char *argv[]);
  ParthenonStatus ParthenonInitEnvFromString(const std::string &deck,
                                             const std::vector<std::string> &overrides);
  void SetRank(int rank, int nranks, const unsigned char *nccl_id);
  void ParthenonInitPackagesAndMesh(const std::vector<LogicalLocation> &leaves = {});
  // ... on a forest given face by face (parthenon_manager.cpp:170-190)
  void ParthenonInitPackagesAndMesh(const forest::ForestDefinition &forest_def);
  ParthenonStatus ParthenonFinalize();
  std::unique_ptr<ParameterInput> pinput;
  std::unique_ptr<ApplicationInput> app_input;
  std::unique_ptr<Mesh> pmesh;

 private:
  void FinishMesh();
  int rank_ = 0, nranks_ = 1;
  std::vector<unsigned char> nccl_id_;
  pb2_comm *comm_ = nullptr;
};

} // namespace parthenon
