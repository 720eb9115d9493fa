// advection_driver.cpp — per-stage task graph of example/advection: the GENERIC shape of a
// Parthenon finite-volume stage (reference example/advection/advection_driver.cpp:56-163):
//   CalculateFluxes -> AddFluxCorrectionTasks -> FluxDivergence -> AverageIndependentData ->
//   UpdateIndependentData -> AddBoundaryExchangeTasks (Send/Receive/SetBounds, and on
//   multilevel meshes restriction + ProlongateBounds every stage) -> FillDerived ->
//   EstimateTimestep.
// The reference runs CalculateFluxes / FillDerived / EstimateTimestep block by block; here
// they take the whole MeshData batch, one launch each.
#include "advection_driver.hpp"

#include "advection_package.hpp"

namespace advection_example {
using namespace parthenon;

AdvectionDriver::AdvectionDriver(ParameterInput *pin, ApplicationInput *app_in, Mesh *pm)
    : MultiStageDriver(pin, app_in, pm) {
  pin->CheckRequired("parthenon/mesh", "ix1_bc");
  pin->CheckRequired("parthenon/mesh", "ox1_bc");
  pin->CheckRequired("parthenon/mesh", "ix2_bc");
  pin->CheckRequired("parthenon/mesh", "ox2_bc");
  pin->CheckDesired("parthenon/mesh", "refinement");
  pin->CheckDesired("parthenon/mesh", "numlevel");
  pin->CheckDesired("Advection", "cfl");
  pin->CheckDesired("Advection", "vx");
  pin->CheckDesired("Advection", "refine_tol");
  pin->CheckDesired("Advection", "derefine_tol");
}

TaskCollection AdvectionDriver::MakeTaskCollection(BlockList_t &blocks, const int stage) {
  using namespace parthenon::Update;
  TaskCollection tc;
  TaskID none(0);

  const Real beta = integrator->beta[stage - 1];
  const Real dt = integrator->dt;
  const auto &stage_name = integrator->stage_name;

  const int num_partitions = pmesh->DefaultNumPartitions();
  TaskRegion &region = tc.AddRegion(num_partitions);
  for (int i = 0; i < num_partitions; i++) {
    auto &tl = region[i];
    auto &mbase = pmesh->mesh_data.GetOrAdd("base", i);
    auto &mc0 = pmesh->mesh_data.GetOrAdd(stage_name[stage - 1], i);
    auto &mc1 = pmesh->mesh_data.GetOrAdd(stage_name[stage], i);
    auto &mdudt = pmesh->mesh_data.GetOrAdd("dUdt", i);

    const auto any = BoundaryType::any;
    auto start_bnd = tl.AddTask(none, StartReceiveBoundBufs<any>, mc1);
    auto start_flx = tl.AddTask(none, StartReceiveFluxCorrections, mc0);

    auto advect_flux = tl.AddTask(none, advection_package::CalculateFluxes, mc0.get());

    // AddFluxCorrectionTasks (boundary_communication.cpp:454-461)
    auto set_flx = advect_flux | start_flx;
    if (pmesh->multilevel) {
      auto send_flx = tl.AddTask(advect_flux, LoadAndSendFluxCorrections, mc0);
      auto recv_flx = tl.AddTask(start_flx | send_flx, ReceiveFluxCorrections, mc0);
      set_flx = tl.AddTask(recv_flx, SetFluxCorrections, mc0);
    }

    auto flux_div = tl.AddTask(set_flx, FluxDivergence<MeshData<Real>>, mc0.get(), mdudt.get());
    auto avg_data = tl.AddTask(flux_div, AverageIndependentData<MeshData<Real>>, mc0.get(),
                               mbase.get(), beta);
    auto update = tl.AddTask(avg_data, UpdateIndependentData<MeshData<Real>>, mc0.get(),
                             mdudt.get(), beta * dt, mc1.get());

    auto bnd = AddBoundaryExchangeTasks(update | start_bnd, tl, mc1, pmesh->multilevel);

    auto fill_derived = tl.AddTask(bnd, FillDerived<MeshData<Real>>, mc1.get());
    if (stage == integrator->nstages) {
      tl.AddTask(fill_derived, EstimateTimestep<MeshData<Real>>, mc1.get());
      // Update refinement (advection_driver.cpp:156-159)
      if (pmesh->adaptive) tl.AddTask(fill_derived, Refinement::Tag, mc1.get());
    }
  }
  (void)blocks; // per-block region of the reference: periodic static meshes have no work there
  return tc;
}

} // namespace advection_example
