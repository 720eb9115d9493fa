// main.cpp — burgers-benchmark executable, same command line as the reference's
// benchmarks/burgers/main.cpp: burgers-benchmark -i burgers.pin [block/key=value ...]
#include <cstdio>
#include <exception>

#include "burgers_driver.hpp"

int main(int argc, char *argv[]) {
  using parthenon::ParthenonManager;
  try {
    ParthenonManager pman;
    pman.app_input->ProcessPackages = burgers_benchmark::ProcessPackages;
    pman.app_input->MeshProblemGenerator = burgers_benchmark::MeshProblemGenerator;
    if (pman.ParthenonInitEnv(argc, argv) != ParthenonManager::ParthenonStatus::ok) return 1;
    pman.ParthenonInitPackagesAndMesh();
    {
      burgers_benchmark::BurgersDriver driver(pman.pinput.get(), pman.app_input.get(),
                                              pman.pmesh.get());
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
