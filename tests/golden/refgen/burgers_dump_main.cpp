// Fixture generator (test infrastructure, NOT product code).
//
// A replacement main() for the reference's benchmarks/burgers app that dumps the raw
// "U" field of every meshblock (full extents, ghosts included) before the time loop and
// after every cycle.  It is compiled against the UNMODIFIED reference sources where they
// lie under /root/reference (burgers_driver.cpp, burgers_package.cpp,
// parthenon_app_inputs.cpp) and linked to a libparthenon.a built out-of-tree (see
// make_fixtures.sh).  Only this file is ours; it calls the reference's public API:
//   ApplicationInput hooks     src/application_input.hpp:43-70
//   MeshBlockData::Get(label)  src/interface/meshblock_data.hpp:262
//
// Dump layout (little endian):  int32 magic=0x50423230, nblocks, ncomp, nk, nj, ni, cycle
//   float64 time, dt
//   then per block: int32 gid, level, lx1, lx2, lx3 (tree-relative, forest.cpp:104-141);
//   float64 xmin[3], xmax[3] ; float64 data[ncomp][nk][nj][ni]
#include <cstdio>
#include <cstdlib>
#include <string>

#include "parthenon_manager.hpp"

#include "burgers_driver.hpp"

namespace {
std::string g_prefix = "dump";
void DumpU(parthenon::Mesh *pmesh, int cycle, double time, double dt) {
  const std::string fname = g_prefix + "." + std::to_string(cycle) + ".bin";
  FILE *fp = std::fopen(fname.c_str(), "wb");
  if (!fp) std::abort();
  auto &first = pmesh->block_list[0]->meshblock_data.Get()->Get("U");
  int hdr[7] = {0x50423230,
                static_cast<int>(pmesh->block_list.size()),
                first.GetDim(4),
                first.GetDim(3),
                first.GetDim(2),
                first.GetDim(1),
                cycle};
  std::fwrite(hdr, sizeof(int), 7, fp);
  double td[2] = {time, dt};
  std::fwrite(td, sizeof(double), 2, fp);
  for (auto &pmb : pmesh->block_list) {
    auto &v = pmb->meshblock_data.Get()->Get("U");
    auto h = v.data.GetHostMirrorAndCopy();
    int bh[5] = {pmb->gid, pmb->loc.level(), static_cast<int>(pmb->loc.lx1()),
                 static_cast<int>(pmb->loc.lx2()), static_cast<int>(pmb->loc.lx3())};
    std::fwrite(bh, sizeof(int), 5, fp);
    double bb[6] = {pmb->block_size.xmin(parthenon::X1DIR), pmb->block_size.xmin(parthenon::X2DIR),
                    pmb->block_size.xmin(parthenon::X3DIR), pmb->block_size.xmax(parthenon::X1DIR),
                    pmb->block_size.xmax(parthenon::X2DIR), pmb->block_size.xmax(parthenon::X3DIR)};
    std::fwrite(bb, sizeof(double), 6, fp);
    for (int n = 0; n < hdr[2]; ++n)
      for (int k = 0; k < hdr[3]; ++k)
        for (int j = 0; j < hdr[4]; ++j)
          for (int i = 0; i < hdr[5]; ++i) {
            double x = h(n, k, j, i);
            std::fwrite(&x, sizeof(double), 1, fp);
          }
  }
  std::fclose(fp);
}
} // namespace

int main(int argc, char *argv[]) {
  using parthenon::ParthenonManager;
  using parthenon::ParthenonStatus;
  ParthenonManager pman;
  if (const char *p = std::getenv("PB2_DUMP_PREFIX")) g_prefix = p;

  pman.app_input->ProcessPackages = burgers_benchmark::ProcessPackages;
  pman.app_input->ProblemGenerator = burgers_benchmark::ProblemGenerator;
  // "reflecting" mesh boundaries need their functions enrolled (application_input.hpp);
  // periodic / outflow decks are unaffected
  pman.app_input->RegisterDefaultReflectingBoundaryConditions();
  pman.app_input->UserWorkBeforeLoop = [](parthenon::Mesh *pm, parthenon::ParameterInput *,
                                          parthenon::SimTime &tm) {
    DumpU(pm, 0, tm.time, tm.dt);
  };
  pman.app_input->PostStepMeshUserWorkInLoop =
      [](parthenon::Mesh *pm, parthenon::ParameterInput *, parthenon::SimTime const &tm) {
        // called before ncycle/time are advanced (driver.cpp:113-123)
        DumpU(pm, tm.ncycle + 1, tm.time + tm.dt, tm.dt);
      };

  auto manager_status = pman.ParthenonInitEnv(argc, argv);
  if (manager_status == ParthenonStatus::complete) {
    pman.ParthenonFinalize();
    return 0;
  }
  if (manager_status == ParthenonStatus::error) {
    pman.ParthenonFinalize();
    return 1;
  }
  pman.ParthenonInitPackagesAndMesh();
  {
    burgers_benchmark::BurgersDriver driver(pman.pinput.get(), pman.app_input.get(),
                                            pman.pmesh.get());
    driver.Execute();
  }
  pman.ParthenonFinalize();
  return 0;
}
