// driver.hpp — evolution loop, multi-stage stepping, integrator tables, history output.
//
// Host-side mirror of src/driver/driver.{hpp,cpp} (EvolutionDriver::Execute :67-193,
// SetGlobalTimeStep :210-270, zone-cycle accounting :57-63,:124,:146-149),
// src/driver/multistage.hpp:39 (MultiStageDriver::Step) and
// src/time_integration/low_storage_integrator.cpp:30-180 (rk1/rk2/vl2/rk3 tables, stage names
// staged_integrator.cpp:23-30).  The only output type kept is the `.hst` history text file
// (src/outputs/history.cpp), which is the reference's own parity artifact for this path.
#pragma once
#include <chrono>
#include <memory>
#include <string>
#include <vector>

#include "bvals.hpp"
#include "mesh_data.hpp"
#include "tasks.hpp"
#include "update.hpp"

namespace parthenon {

class StagedIntegrator {
 public:
  StagedIntegrator() = default;
  explicit StagedIntegrator(ParameterInput *pin);
  int nstages = 1, nbuffers = 1;
  Real dt = 0.0;
  std::vector<Real> delta, beta, gam0, gam1, c;
  std::vector<std::string> stage_name; // "base", "1", ..., "base"
  const std::string &GetName() const { return name_; }

 private:
  std::string name_;
};
using LowStorageIntegrator = StagedIntegrator;

class Driver {
 public:
  Driver(ParameterInput *pin, ApplicationInput *app_in, Mesh *pm)
      : pinput(pin), app_input(app_in), pmesh(pm) {}
  virtual ~Driver() = default;
  virtual DriverStatus Execute() = 0;
  ParameterInput *pinput;
  ApplicationInput *app_input;
  Mesh *pmesh;
};

class EvolutionDriver : public Driver {
 public:
  EvolutionDriver(ParameterInput *pin, ApplicationInput *app_in, Mesh *pm);
  DriverStatus Execute() override;
  virtual TaskListStatus Step() = 0;
  void SetGlobalTimeStep();
  void InitializeBlockTimeSteps();
  // one cycle of the main loop of Execute (Step + bookkeeping + dt), for callers that
  // drive the loop themselves (bench.py, tests)
  TaskListStatus DoCycle();
  // the two halves of a cycle (driver.cpp:112-123 and :125-129): Step + time advance, then
  // LoadBalancingAndAdaptiveMeshRefinement + SetGlobalTimeStep
  TaskListStatus StepAndAdvanceTime();
  void AdaptMeshAndSetTimeStep();
  // call once before the first DoCycle (what Execute does before its loop)
  void PreExecute();
  void OutputCycleDiagnostics();
  // zone-cycles per wall second since the timer reset (driver.cpp:57-63)
  double ZoneCyclesPerSecond() const;
  SimTime tm;
  bool quiet = false;

 protected:
  void MakeHistoryOutput(bool force);
  Real dt_init = std::numeric_limits<Real>::max(), dt_user = std::numeric_limits<Real>::max(),
       dt_force = -1.0, dt_factor = 2.0, dt_floor = std::numeric_limits<Real>::min(),
       dt_ceil = std::numeric_limits<Real>::max();
  // "off the rails" guards of driver.cpp:243-263
  Real dt_min = std::numeric_limits<Real>::min(), dt_max = std::numeric_limits<Real>::max();
  int dt_min_count = 0, dt_max_count = 0, dt_min_count_max = 10, dt_max_count_max = 1;
  bool dt_init_force = false;
  int perf_cycle_offset = 0;
  std::chrono::steady_clock::time_point timer_main_;
  // history outputs (<parthenon/outputN> with file_type = hst)
  struct HistoryOutput {
    std::string filename, data_format;
    Real dt, next_time;
    bool header_written = false;
  };
  std::vector<HistoryOutput> hst_outputs_;
};

class MultiStageDriver : public EvolutionDriver {
 public:
  MultiStageDriver(ParameterInput *pin, ApplicationInput *app_in, Mesh *pm)
      : EvolutionDriver(pin, app_in, pm), integrator(std::make_unique<StagedIntegrator>(pin)) {}
  // multistage.hpp:39-55
  TaskListStatus Step() override;
  virtual TaskCollection MakeTaskCollection(BlockList_t &blocks, int stage) = 0;

 protected:
  std::unique_ptr<StagedIntegrator> integrator;
};

} // namespace parthenon
