// burgers_driver.cpp — per-stage task graph of the burgers benchmark.
//
// Two shapes, selected by <pb2>/fused_stage:
//  * false: task for task the list of the reference's burgers_driver.cpp:78-127
//    (CalculateFluxes, FluxDivergence, AverageIndependentData, UpdateIndependentData,
//    Send/Receive/SetBounds in the local/nonlocal split, FillDerived, EstimateTimestep),
//    each task one C-ABI launch group;
//  * true (default): the dense tasks collapse into burgers_package::FusedStage, the
//    exchange tasks stay as they are.
#include "burgers_driver.hpp"

#include "burgers_package.hpp"

namespace burgers_benchmark {
using namespace parthenon;

BurgersDriver::BurgersDriver(ParameterInput *pin, ApplicationInput *app_in, Mesh *pm)
    : MultiStageDriver(pin, app_in, pm) {
  pin->CheckRequired("parthenon/mesh", "ix1_bc");
  pin->CheckRequired("parthenon/mesh", "ox1_bc");
  pin->CheckRequired("parthenon/mesh", "ix2_bc");
  pin->CheckRequired("parthenon/mesh", "ox2_bc");
  pin->CheckDesired("parthenon/mesh", "refinement");
  pin->CheckDesired("parthenon/mesh", "numlevel");
}

TaskCollection BurgersDriver::MakeTaskCollection(BlockList_t &blocks, const int stage) {
  using namespace parthenon::Update;
  TaskCollection tc;
  TaskID none(0);

  const Real beta = integrator->beta[stage - 1];
  const Real dt = integrator->dt;
  const auto &stage_name = integrator->stage_name;
  const bool fused = pmesh->packages.Get("burgers_package")->Param<bool>("fused_stage");
  const bool last = stage == integrator->nstages;

  const int num_partitions = pmesh->DefaultNumPartitions();
  TaskRegion &region = tc.AddRegion(num_partitions);
  for (int i = 0; i < num_partitions; i++) {
    auto &tl = region[i];
    auto &mbase = pmesh->mesh_data.GetOrAdd("base", i);
    auto &mc0 = pmesh->mesh_data.GetOrAdd(stage_name[stage - 1], i);
    auto &mc1 = pmesh->mesh_data.GetOrAdd(stage_name[stage], i);

    const auto any = BoundaryType::any;
    const auto local = BoundaryType::local;
    const auto nonlocal = BoundaryType::nonlocal;

    auto start_bnd = tl.AddTask(none, StartReceiveBoundBufs<any>, mc1);
    TaskID update = none;
    if (fused) {
      update = tl.AddTask(none, burgers_package::FusedStage, mc0.get(), mbase.get(), mc1.get(),
                          beta, dt, last);
    } else {
      auto &mdudt = pmesh->mesh_data.GetOrAdd("dUdt", i);
      auto start_flx_recv = tl.AddTask(none, StartReceiveFluxCorrections, mc0);
      auto flx = tl.AddTask(none, burgers_package::CalculateFluxes, mc0.get());
      auto send_flx = tl.AddTask(flx, LoadAndSendFluxCorrections, mc0);
      auto recv_flx = tl.AddTask(start_flx_recv | send_flx, ReceiveFluxCorrections, mc0);
      auto set_flx = tl.AddTask(recv_flx, SetFluxCorrections, mc0);
      auto flux_div =
          tl.AddTask(set_flx, FluxDivergence<MeshData<Real>>, mc0.get(), mdudt.get());
      auto avg_data = tl.AddTask(flux_div, AverageIndependentData<MeshData<Real>>, mc0.get(),
                                 mbase.get(), beta);
      update = tl.AddTask(avg_data, UpdateIndependentData<MeshData<Real>>, mc0.get(),
                          mdudt.get(), beta * dt, mc1.get());
    }

    // boundary exchange, local / nonlocal split (burgers_driver.cpp:106-119)
    auto send = tl.AddTask(update, SendBoundBufs<nonlocal>, mc1);
    auto send_local = tl.AddTask(update, SendBoundBufs<local>, mc1);
    auto recv_local = tl.AddTask(update | send_local, ReceiveBoundBufs<local>, mc1);
    auto set_local = tl.AddTask(recv_local, SetBounds<local>, mc1);
    auto recv = tl.AddTask(start_bnd | update | send, ReceiveBoundBufs<nonlocal>, mc1);
    auto set = tl.AddTask(recv | set_local, SetBounds<nonlocal>, mc1);

    // set physical boundaries (the per-block second region of burgers_driver.cpp:129-147, here
    // one launch per direction for the whole batch; nothing to do on periodic meshes)
    auto set_bc = tl.AddTask(set, ApplyBoundaryConditionsMD, mc1);
    // Update refinement (burgers_driver.cpp:139-145)
    if (last && pmesh->adaptive) tl.AddTask(set_bc, Refinement::Tag, mc1.get());

    if (fused) {
      if (last) tl.AddTask(set, burgers_package::CollectFusedTimestep, mc1.get());
    } else {
      tl.AddTask(update, FillDerived<MeshData<Real>>, mc1.get());
      if (last) tl.AddTask(update, EstimateTimestep<MeshData<Real>>, mc1.get());
    }
  }
  // second region of the reference (per-block ApplyBoundaryConditions / Refinement::Tag):
  // periodic static meshes have nothing to do there
  (void)blocks;
  return tc;
}

} // namespace burgers_benchmark
