// main.cpp — sparse_advection-example executable, same command line as the reference's
// example/sparse_advection/main.cpp.
#include <cstdio>
#include <exception>

#include "sparse_advection_driver.hpp"

int main(int argc, char *argv[]) {
  using parthenon::ParthenonManager;
  try {
    ParthenonManager pman;
    pman.app_input->ProcessPackages = sparse_advection_example::ProcessPackages;
    pman.app_input->MeshProblemGenerator = sparse_advection_example::MeshProblemGenerator;
    if (pman.ParthenonInitEnv(argc, argv) != ParthenonManager::ParthenonStatus::ok) return 1;
    pman.ParthenonInitPackagesAndMesh();
    {
      sparse_advection_example::SparseAdvectionDriver driver(
          pman.pinput.get(), pman.app_input.get(), pman.pmesh.get());
      const auto status = driver.Execute();
      if (status == parthenon::DriverStatus::failed) return 2;
    }
    pman.ParthenonFinalize();
  } catch (const std::exception &e) {
    std::fprintf(stderr, "%s\n", e.what());
    return 1;
  }
  return 0;
}
