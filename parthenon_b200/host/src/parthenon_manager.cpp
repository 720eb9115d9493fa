// parthenon_manager.cpp — environment + mesh bring-up (reference src/parthenon_manager.cpp:
// ParthenonInitEnv :48-168, ParthenonInitPackagesAndMesh :170-250) and Mesh::Initialize
// (mesh.cpp:745-860): problem generator, first ghost exchange, FillDerived.
#include <cstring>

#include "pb2/parthenon.hpp"

namespace parthenon {

ParthenonManager::~ParthenonManager() { ParthenonFinalize(); }

ParthenonManager::ParthenonStatus ParthenonManager::ParthenonInitEnv(int argc, char *argv[]) {
  std::string deck;
  for (int i = 1; i < argc; ++i)
    if (std::strcmp(argv[i], "-i") == 0 && i + 1 < argc) deck = argv[i + 1];
  if (deck.empty()) {
    std::fprintf(stderr, "Usage: %s -i <input deck> [block/key=value ...]\n", argv[0]);
    return ParthenonStatus::error;
  }
  pinput = std::make_unique<ParameterInput>();
  pinput->LoadFromFile(deck);
  pinput->ModifyFromCmdline(argc, argv);
  return ParthenonStatus::ok;
}

ParthenonManager::ParthenonStatus
ParthenonManager::ParthenonInitEnvFromString(const std::string &deck,
                                             const std::vector<std::string> &overrides) {
  pinput = std::make_unique<ParameterInput>();
  pinput->LoadFromString(deck);
  for (auto &o : overrides) pinput->ModifyFromString(o);
  return ParthenonStatus::ok;
}

void ParthenonManager::SetRank(int rank, int nranks, const unsigned char *nccl_id) {
  rank_ = rank;
  nranks_ = nranks;
  if (nccl_id) nccl_id_.assign(nccl_id, nccl_id + PB2_NCCL_UNIQUE_ID_BYTES);
}

void ParthenonManager::ParthenonInitPackagesAndMesh(const std::vector<LogicalLocation> &leaves) {
  PARTHENON_REQUIRE(app_input->ProcessPackages != nullptr, "ProcessPackages must be set");
  // parthenon_manager.cpp:141: nghost must be known before packages read it
  pinput->GetOrAddInteger("parthenon/mesh", "nghost", 2);
  Packages_t packages = app_input->ProcessPackages(pinput);
  pmesh = std::make_unique<Mesh>(pinput.get(), app_input.get(), packages, rank_, nranks_, leaves);
  FinishMesh();
}

void ParthenonManager::ParthenonInitPackagesAndMesh(const forest::ForestDefinition &forest_def) {
  PARTHENON_REQUIRE(app_input->ProcessPackages != nullptr, "ProcessPackages must be set");
  pinput->GetOrAddInteger("parthenon/mesh", "nghost", 2);
  Packages_t packages = app_input->ProcessPackages(pinput);
  pmesh = std::make_unique<Mesh>(pinput.get(), app_input.get(), packages, forest_def, rank_, nranks_);
  FinishMesh();
}

void ParthenonManager::FinishMesh() {
  pb2_stream_t st = nullptr, cs = nullptr;
  PB2_CHECK(pb2_stream_create(&st));
  PB2_CHECK(pb2_stream_create_priority(&cs, 1)); // halo pack / NCCL: first in line for SMs
  pmesh->stream = st;
  pmesh->comm_stream = cs;
  if (nranks_ > 1) {
    PARTHENON_REQUIRE(nccl_id_.size() == PB2_NCCL_UNIQUE_ID_BYTES,
                      "multi-rank run needs the NCCL unique id (SetRank)");
    PB2_CHECK(pb2_comm_create(&comm_, rank_, nranks_, nccl_id_.data()));
    pmesh->comm = comm_;
  }
  pmesh->Initialize(true, pinput.get(), app_input.get());
}

ParthenonManager::ParthenonStatus ParthenonManager::ParthenonFinalize() {
  if (pmesh) {
    pb2_stream_t st = pmesh->stream, cs = pmesh->comm_stream;
    if (st) pb2_stream_sync(st);
    pmesh.reset();
    if (st) pb2_stream_destroy(st);
    if (cs) pb2_stream_destroy(cs);
  }
  if (comm_) pb2_comm_destroy(comm_);
  comm_ = nullptr;
  return ParthenonStatus::complete;
}

void Mesh::Initialize(bool init_problem, ParameterInput *pin, ApplicationInput *app_in) {
  bool init_done = true;
  do { // mesh.cpp:745-860: on adaptive meshes, regenerate the problem until the mesh settles
    const int np = DefaultNumPartitions();
    if (init_problem && app_in && app_in->MeshProblemGenerator)
      for (int p = 0; p < np; ++p)
        app_in->MeshProblemGenerator(mesh_data.GetOrAdd("base", p).get(), pin);
    // mesh.cpp:640-706 CommunicateBoundaries (+ prolongation on multilevel meshes), then
    // FillDerived on every batch
    auto &base0 = mesh_data.GetOrAdd("base", 0);
    CommunicateBoundaries(base0, true);
    for (int p = 0; p < np; ++p) Update::FillDerived(mesh_data.GetOrAdd("base", p).get());
    PB2_CHECK(pb2_stream_sync(stream));
    init_done = true;
    if (init_problem && adaptive) {
      for (int p = 0; p < np; ++p) Refinement::Tag(mesh_data.GetOrAdd("base", p).get());
      const int nb_before = nbtotal;
      LoadBalancingAndAdaptiveMeshRefinement(pin, app_in);
      init_done = nbtotal == nb_before;
    }
  } while (!init_done);
}

} // namespace parthenon
