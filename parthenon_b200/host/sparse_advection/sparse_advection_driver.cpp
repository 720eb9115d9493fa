// sparse_advection_driver.cpp — per-stage task graph of example/sparse_advection (reference
// sparse_advection_driver.cpp:56-149): the generic stage list of example/advection over
// sparse fields, plus Update::SparseDealloc after the last stage's exchange.
#include "sparse_advection_driver.hpp"

#include "sparse_advection_package.hpp"

namespace sparse_advection_example {
using namespace parthenon;

SparseAdvectionDriver::SparseAdvectionDriver(ParameterInput *pin, ApplicationInput *app_in,
                                             Mesh *pm)
    : MultiStageDriver(pin, app_in, pm) {
  pin->CheckRequired("parthenon/mesh", "ix1_bc");
  pin->CheckRequired("parthenon/mesh", "ox1_bc");
  pin->CheckRequired("parthenon/mesh", "ix2_bc");
  pin->CheckRequired("parthenon/mesh", "ox2_bc");
  pin->CheckDesired("parthenon/mesh", "refinement");
  pin->CheckDesired("parthenon/mesh", "numlevel");
  pin->CheckDesired("sparse_advection", "cfl");
  pin->CheckDesired("sparse_advection", "refine_tol");
  pin->CheckDesired("sparse_advection", "derefine_tol");
}

TaskCollection SparseAdvectionDriver::MakeTaskCollection(BlockList_t &blocks, const int stage) {
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
    auto advect_flux = tl.AddTask(none, sparse_advection_package::CalculateFluxes, mc0.get());
    auto start_flxcor = tl.AddTask(none, StartReceiveFluxCorrections, mc0);
    auto start_bound = tl.AddTask(none, StartReceiveBoundBufs<any>, mc1);

    auto set_flxcor = advect_flux | start_flxcor;
    if (pmesh->multilevel) {
      auto send_flx = tl.AddTask(set_flxcor, LoadAndSendFluxCorrections, mc0);
      auto recv_flx = tl.AddTask(send_flx, ReceiveFluxCorrections, mc0);
      set_flxcor = tl.AddTask(recv_flx, SetFluxCorrections, mc0);
    }

    auto flux_div =
        tl.AddTask(set_flxcor, FluxDivergence<MeshData<Real>>, mc0.get(), mdudt.get());
    auto avg_data = tl.AddTask(flux_div, AverageIndependentData<MeshData<Real>>, mc0.get(),
                               mbase.get(), beta);
    auto update = tl.AddTask(avg_data, UpdateIndependentData<MeshData<Real>>, mc0.get(),
                             mdudt.get(), beta * dt, mc1.get());

    auto boundary = AddBoundaryExchangeTasks(update | start_bound, tl, mc1, pmesh->multilevel);

    // if this is the last stage, check if we can deallocate any sparse variables
    if (stage == integrator->nstages) {
      auto dealloc = tl.AddTask(boundary, SparseDealloc, mc1.get());
      auto new_dt = tl.AddTask(dealloc, EstimateTimestep<MeshData<Real>>, mc1.get());
      // update refinement (sparse_advection_driver.cpp:147-151)
      if (pmesh->adaptive) tl.AddTask(new_dt, Refinement::Tag, mc1.get());
    }
  }
  (void)blocks;
  return tc;
}

} // namespace sparse_advection_example
