// main.cpp — advection-example executable, same command line as the reference's
// example/advection/main.cpp: advection-example -i parthinput.advection [block/key=value ...]
#include <cstdio>
#include <exception>

#include "advection_driver.hpp"

int main(int argc, char *argv[]) {
  using parthenon::ParthenonManager;
  try {
    ParthenonManager pman;
    pman.app_input->ProcessPackages = advection_example::ProcessPackages;
    pman.app_input->MeshProblemGenerator = advection_example::MeshProblemGenerator;
    if (pman.ParthenonInitEnv(argc, argv) != ParthenonManager::ParthenonStatus::ok) return 1;
    pman.ParthenonInitPackagesAndMesh();
    {
      advection_example::AdvectionDriver driver(pman.pinput.get(), pman.app_input.get(),
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
