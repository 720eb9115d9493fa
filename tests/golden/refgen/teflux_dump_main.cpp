// Fixture generator (test infrastructure, NOT product code).
//
// A minimal application on the UNMODIFIED reference library (libparthenon.a built out-of-tree,
// see make_fixtures.sh) that pins FLUX CORRECTION OF A FACE FIELD: the flux of a face-centred
// field is an edge-centred field ("bnd_flux::B", Metadata::Flux | Edge), and at fine-coarse
// boundaries the fine blocks restrict their edge fluxes on shared faces AND shared block edges
// and send them to the coarser neighbour (GetFluxCorrectionElements, bnd_info.cpp:71-103;
// ForEachBoundary<flxcor_*>, loop_utils.hpp:134-158), where only the values of the owning fine
// block land (block_ownership.cpp).  One package with a face field B (Metadata::WithFluxes) on a
// statically refined mesh; every entry of the flux field gets the block-dependent code
//     value = (gid + 1) * 1e6 + element * 1e5 + flat (k, j, i) index
// then SendBoundBufs<flxcor_send> / ReceiveBoundBufs<flxcor_recv> / SetBounds<flxcor_recv> run
// once (AddFluxCorrectionTasks, boundary_communication.cpp:454-461) and the flux field is dumped.
// Only this file is ours.
//
// Dump layout: the one of burgers_dump_main.cpp; one file, ncomp = 3 edge elements, extents are
// the padded array extents of the flux field.
#include <cstdio>
#include <cstdlib>
#include <memory>
#include <string>
#include <vector>

#include "bvals/comms/bvals_in_one.hpp"
#include "parthenon_manager.hpp"
#include <parthenon/package.hpp>

namespace {
using namespace parthenon;
using namespace parthenon::package::prelude;
std::string g_prefix = "dump";
const char *kFlux = "bnd_flux::B";

Packages_t ProcessPackages(std::unique_ptr<ParameterInput> &pin) {
  Packages_t packages;
  auto pkg = std::make_shared<StateDescriptor>("teflux");
  Metadata m({Metadata::Face, Metadata::Independent, Metadata::WithFluxes, Metadata::FillGhost});
  pkg->AddField("B", m);
  packages.Add(pkg);
  return packages;
}

void FillFlux(Mesh *pmesh) {
  for (auto &pmb : pmesh->block_list) {
    auto &v = pmb->meshblock_data.Get()->Get(kFlux);
    auto h = v.data.GetHostMirror();
    const int ne = v.data.GetDim(7), nc = v.data.GetDim(4), nk = v.data.GetDim(3),
              nj = v.data.GetDim(2), ni = v.data.GetDim(1);
    for (int e = 0; e < ne; ++e)
      for (int c = 0; c < nc; ++c)
        for (int k = 0; k < nk; ++k)
          for (int j = 0; j < nj; ++j)
            for (int i = 0; i < ni; ++i)
              h(e, 0, 0, c, k, j, i) = (pmb->gid + 1) * 1.0e6 + e * 1.0e5 + ((k * nj + j) * ni + i);
    v.data.DeepCopy(h);
  }
}

void Dump(Mesh *pmesh) {
  const std::string fname = g_prefix + ".0.bin";
  FILE *fp = std::fopen(fname.c_str(), "wb");
  if (!fp) std::abort();
  auto &v0 = pmesh->block_list[0]->meshblock_data.Get()->Get(kFlux);
  const int ne = v0.data.GetDim(7), nc = v0.data.GetDim(4), nk = v0.data.GetDim(3),
            nj = v0.data.GetDim(2), ni = v0.data.GetDim(1);
  int hdr[7] = {0x50423230, static_cast<int>(pmesh->block_list.size()), ne * nc, nk, nj, ni, 0};
  std::fwrite(hdr, sizeof(int), 7, fp);
  double td[2] = {0.0, 0.0};
  std::fwrite(td, sizeof(double), 2, fp);
  for (auto &pmb : pmesh->block_list) {
    int bh[5] = {pmb->gid, pmb->loc.level(), static_cast<int>(pmb->loc.lx1()),
                 static_cast<int>(pmb->loc.lx2()), static_cast<int>(pmb->loc.lx3())};
    std::fwrite(bh, sizeof(int), 5, fp);
    double bb[6] = {pmb->block_size.xmin(X1DIR), pmb->block_size.xmin(X2DIR),
                    pmb->block_size.xmin(X3DIR), pmb->block_size.xmax(X1DIR),
                    pmb->block_size.xmax(X2DIR), pmb->block_size.xmax(X3DIR)};
    std::fwrite(bb, sizeof(double), 6, fp);
    auto &v = pmb->meshblock_data.Get()->Get(kFlux);
    auto h = v.data.GetHostMirrorAndCopy();
    for (int e = 0; e < ne; ++e)
      for (int c = 0; c < nc; ++c)
        for (int k = 0; k < nk; ++k)
          for (int j = 0; j < nj; ++j)
            for (int i = 0; i < ni; ++i) {
              double x = h(e, 0, 0, c, k, j, i);
              std::fwrite(&x, sizeof(double), 1, fp);
            }
  }
  std::fclose(fp);
}
} // namespace

int main(int argc, char *argv[]) {
  ParthenonManager pman;
  if (const char *p = std::getenv("PB2_DUMP_PREFIX")) g_prefix = p;
  pman.app_input->ProcessPackages = ProcessPackages;
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
  Mesh *pmesh = pman.pmesh.get();
  FillFlux(pmesh);
  {
    // AddFluxCorrectionTasks on every partition of "base" (boundary_communication.cpp:454-461)
    TaskCollection tc;
    const int np = pmesh->DefaultNumPartitions();
    TaskRegion &region = tc.AddRegion(np);
    for (int i = 0; i < np; ++i) {
      auto &md = pmesh->mesh_data.GetOrAdd("base", i);
      TaskID none(0);
      auto start = region[i].AddTask(none, StartReceiveFluxCorrections, md);
      AddFluxCorrectionTasks(start, region[i], md, pmesh->multilevel);
    }
    auto status = tc.Execute();
    if (status != TaskListStatus::complete) std::abort();
  }
  Dump(pmesh);
  pman.ParthenonFinalize();
  return 0;
}
