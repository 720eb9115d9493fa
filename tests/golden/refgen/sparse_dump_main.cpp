// Fixture generator (test infrastructure, NOT product code).
//
// A replacement main() for the reference's example/sparse_advection app that dumps the four
// sparse fields "sparse_0" .. "sparse_3" of every meshblock (full extents, ghosts included)
// before the time loop and after every cycle.  Compiled against the UNMODIFIED reference
// sources where they lie under /root/reference (sparse_advection_driver.cpp,
// sparse_advection_package.cpp, parthenon_app_inputs.cpp) and linked to a libparthenon.a
// built out-of-tree (see make_fixtures.sh).  Only this file is ours.
//
// Dump layout: the one of burgers_dump_main.cpp with ncomp = 4; a field that is NOT
// ALLOCATED on a block is written as NaN (pack_dumps.py needs no change; tests read the
// allocation status as ~isnan).
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <limits>
#include <string>

#include "parthenon_manager.hpp"

#include "sparse_advection_driver.hpp"
#include "sparse_advection_package.hpp"

namespace {
std::string g_prefix = "dump";
int g_every = 1; // $PB2_DUMP_EVERY: dump cycles 0, N, 2N, ...
constexpr int NF = sparse_advection_package::NUM_FIELDS;
void DumpU(parthenon::Mesh *pmesh, int cycle, double time, double dt) {
  if (cycle % g_every != 0) return;
  const std::string fname = g_prefix + "." + std::to_string(cycle) + ".bin";
  FILE *fp = std::fopen(fname.c_str(), "wb");
  if (!fp) std::abort();
  auto &cb = pmesh->block_list[0]->cellbounds;
  const int nk = cb.ncellsk(parthenon::IndexDomain::entire),
            nj = cb.ncellsj(parthenon::IndexDomain::entire),
            ni = cb.ncellsi(parthenon::IndexDomain::entire);
  int hdr[7] = {0x50423230, static_cast<int>(pmesh->block_list.size()), NF, nk, nj, ni, cycle};
  std::fwrite(hdr, sizeof(int), 7, fp);
  double td[2] = {time, dt};
  std::fwrite(td, sizeof(double), 2, fp);
  for (auto &pmb : pmesh->block_list) {
    int bh[5] = {pmb->gid, pmb->loc.level(), static_cast<int>(pmb->loc.lx1()),
                 static_cast<int>(pmb->loc.lx2()), static_cast<int>(pmb->loc.lx3())};
    std::fwrite(bh, sizeof(int), 5, fp);
    double bb[6] = {pmb->block_size.xmin(parthenon::X1DIR), pmb->block_size.xmin(parthenon::X2DIR),
                    pmb->block_size.xmin(parthenon::X3DIR), pmb->block_size.xmax(parthenon::X1DIR),
                    pmb->block_size.xmax(parthenon::X2DIR), pmb->block_size.xmax(parthenon::X3DIR)};
    std::fwrite(bb, sizeof(double), 6, fp);
    auto rc = pmb->meshblock_data.Get();
    for (int f = 0; f < NF; ++f) {
      const bool alloc = rc->IsAllocated("sparse", f);
      if (alloc) {
        auto &v = rc->Get("sparse", f);
        auto h = v.data.GetHostMirrorAndCopy();
        for (int k = 0; k < nk; ++k)
          for (int j = 0; j < nj; ++j)
            for (int i = 0; i < ni; ++i) {
              double x = h(0, k, j, i);
              std::fwrite(&x, sizeof(double), 1, fp);
            }
      } else {
        const double x = std::numeric_limits<double>::quiet_NaN();
        for (int n = 0; n < nk * nj * ni; ++n) std::fwrite(&x, sizeof(double), 1, fp);
      }
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
  if (const char *p = std::getenv("PB2_DUMP_EVERY")) g_every = std::atoi(p);

  pman.app_input->ProcessPackages = sparse_advection_example::ProcessPackages;
  pman.app_input->ProblemGenerator = sparse_advection_example::ProblemGenerator;
  pman.app_input->RegisterDefaultReflectingBoundaryConditions(); // as the example's main.cpp:30
  pman.app_input->UserWorkBeforeLoop = [](parthenon::Mesh *pm, parthenon::ParameterInput *,
                                          parthenon::SimTime &tm) {
    DumpU(pm, 0, tm.time, tm.dt);
  };
  pman.app_input->PostStepMeshUserWorkInLoop =
      [](parthenon::Mesh *pm, parthenon::ParameterInput *, parthenon::SimTime const &tm) {
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
    sparse_advection_example::SparseAdvectionDriver driver(
        pman.pinput.get(), pman.app_input.get(), pman.pmesh.get());
    driver.Execute();
  }
  pman.ParthenonFinalize();
  return 0;
}
