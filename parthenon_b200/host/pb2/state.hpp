// state.hpp — package / field registry: Metadata, Params, StateDescriptor, Packages_t.
//
// Host-only mirror of the reference's interface layer: Metadata flags
// (src/interface/metadata.hpp:42-127), Params (src/interface/params.hpp), StateDescriptor
// with its std::function hooks (src/interface/state_descriptor.hpp:79-410) and Packages_t.
// Applications register fields exactly as in Parthenon:
//   Metadata m({Metadata::Cell, Metadata::Independent, Metadata::FillGhost,
//               Metadata::WithFluxes}, std::vector<int>{ncomp});
//   pkg->AddField("U", m);
// The registry decides which arrays MeshData allocates in HBM and which of them the
// ghost-exchange tables cover (FillGhost).
#pragma once
#include <any>
#include <functional>
#include <map>
#include <memory>
#include <string>
#include <typeinfo>
#include <utility>
#include <vector>

#include "types.hpp"

namespace parthenon {

class Mesh;
class MeshBlock;
class ParameterInput;
template <typename T>
class MeshData;
template <typename T>
class MeshBlockData;

// prolongation operator ids understood by the CUDA library (PB2_PROLONG_* of the C ABI)
namespace refinement_ops {
struct ProlongateSharedMinMod { static constexpr int id = 0; };
struct ProlongateSharedLinear { static constexpr int id = 1; };
struct ProlongatePiecewiseConstant { static constexpr int id = 2; };
struct RestrictAverage { static constexpr int id = 0; };
struct ProlongateInternalAverage { static constexpr int id = 0; };
struct ProlongateInternalTothAndRoe { static constexpr int id = 1; }; // face fields only
} // namespace refinement_ops

struct MetadataFlag {
  int bit;
  bool operator==(const MetadataFlag &o) const { return bit == o.bit; }
};

class Metadata {
 public:
  // metadata.hpp:42-127 (same names; numbering is internal)
  static constexpr MetadataFlag None{1}, Cell{2}, Face{3}, Edge{4}, Node{5}, Particle{6},
      Swarm{7}, Private{8}, Provides{9}, Requires{10}, Overridable{11}, Vector{12}, Tensor{13},
      Boolean{14}, Integer{15}, Real{16}, Independent{17}, Derived{18}, Advected{19},
      Conserved{20}, Intensive{21}, Restart{22}, Sparse{23}, SparseCommunication{24},
      OneCopy{25}, FillGhost{26}, WithFluxes{27}, ForceRemeshComm{28}, GMGProlongate{29},
      GMGRestrict{30}, ForceAllocOnNewBlocks{31}, Fine{32}, Flux{33}, CellMemAligned{34},
      CoordinatesVec{35};

  Metadata() = default;
  explicit Metadata(const std::vector<MetadataFlag> &flags, const std::vector<int> &shape = {},
                    const std::vector<std::string> &component_labels = {})
      : shape_(shape), component_labels_(component_labels) {
    for (auto f : flags) Set(f);
    // defaults: metadata.cpp:60-110 (topology None, role Provides, Real, Derived)
    if (!(IsSet(Independent) || IsSet(Derived))) Set(Derived);
    if (!(IsSet(Private) || IsSet(Requires) || IsSet(Overridable))) Set(Provides);
    if (IsSet(FillGhost) || IsSet(WithFluxes)) refinement_registered_ = true;
  }
  void Set(MetadataFlag f) { bits_ |= (1ull << f.bit); }
  void Unset(MetadataFlag f) { bits_ &= ~(1ull << f.bit); }
  bool IsSet(MetadataFlag f) const { return (bits_ >> f.bit) & 1ull; }
  bool AllFlagsSet(const std::vector<MetadataFlag> &fs) const {
    for (auto f : fs)
      if (!IsSet(f)) return false;
    return true;
  }
  bool AnyFlagsSet(const std::vector<MetadataFlag> &fs) const {
    for (auto f : fs)
      if (IsSet(f)) return true;
    return false;
  }
  const std::vector<int> &Shape() const { return shape_; }
  // number of flattened tensor components (t,u,v)
  int NumComponents() const {
    int n = 1;
    for (int s : shape_) n *= s;
    return n;
  }
  const std::vector<std::string> &getComponentLabels() const { return component_labels_; }

  // sparse fields (metadata.hpp:470-510)
  bool IsSparse() const { return IsSet(Sparse); }
  void SetSparseThresholds(parthenon::Real alloc, parthenon::Real dealloc,
                           parthenon::Real default_val = 0.0) {
    allocation_threshold_ = alloc;
    deallocation_threshold_ = dealloc;
    default_value_ = default_val;
  }
  parthenon::Real GetAllocationThreshold() const { return allocation_threshold_; }
  parthenon::Real GetDeallocationThreshold() const { return deallocation_threshold_; }
  parthenon::Real GetDefaultValue() const { return default_value_; }

  // metadata.hpp:553-568: ops are identified by type; the CUDA library implements the
  // reference's stock cell-centred operators
  template <class ProlongationOp, class RestrictionOp,
            class InternalOp = refinement_ops::ProlongateInternalAverage>
  void RegisterRefinementOps() {
    prolongation_op_ = ProlongationOp::id;
    restriction_op_ = RestrictionOp::id;
    internal_op_ = InternalOp::id;
    refinement_registered_ = true;
  }
  bool IsRefined() const { return refinement_registered_; }
  int ProlongationOp() const { return prolongation_op_; }
  int RestrictionOp() const { return restriction_op_; }
  int InternalProlongationOp() const { return internal_op_; }

 private:
  uint64_t bits_ = 0;
  std::vector<int> shape_;
  std::vector<std::string> component_labels_;
  parthenon::Real allocation_threshold_ = 0.0, deallocation_threshold_ = 0.0,
                  default_value_ = 0.0;
  int prolongation_op_ = 0, restriction_op_ = 0; // MinMod / Average (metadata.hpp:337)
  int internal_op_ = 0;                          // ProlongateInternalAverage
  bool refinement_registered_ = false;
};

// params.hpp: typed key/value store
class Params {
 public:
  template <typename T>
  void Add(const std::string &key, T value, bool is_mutable = false) {
    PARTHENON_REQUIRE(!map_.count(key), "Key " + key + " already exists in Params");
    map_[key] = std::make_pair(std::any(std::move(value)), is_mutable);
  }
  template <typename T>
  void Update(const std::string &key, T value) {
    auto it = map_.find(key);
    PARTHENON_REQUIRE(it != map_.end(), "Key " + key + " missing from Params");
    PARTHENON_REQUIRE(it->second.second, "Parameter " + key + " must be marked as mutable");
    it->second.first = std::any(std::move(value));
  }
  template <typename T>
  const T &Get(const std::string &key) const {
    auto it = map_.find(key);
    PARTHENON_REQUIRE(it != map_.end(), "Key " + key + " doesn't exist in Params");
    const T *p = std::any_cast<T>(&it->second.first);
    PARTHENON_REQUIRE(p != nullptr, "WRONG TYPE FOR KEY '" + key + "'");
    return *p;
  }
  bool hasKey(const std::string &key) const { return map_.count(key) > 0; }

 private:
  std::map<std::string, std::pair<std::any, bool>> map_;
};

// history output registration (src/outputs/outputs.hpp UserHistoryOperation / HistoryOutputVar)
enum class UserHistoryOperation { sum, max, min };
struct HistoryOutputVar {
  UserHistoryOperation hst_op;
  std::function<Real(MeshData<Real> *)> hst_fun;
  std::string label;
  HistoryOutputVar(UserHistoryOperation op, std::function<Real(MeshData<Real> *)> f,
                   std::string l)
      : hst_op(op), hst_fun(std::move(f)), label(std::move(l)) {}
};
using HstVar_list = std::vector<HistoryOutputVar>;
// a package may instead register ONE function that returns all its history columns from a
// single device pass (B200: 8 octant sums of the burgers benchmark in one launch)
struct HistoryOutputVec {
  UserHistoryOperation hst_op;
  std::function<std::vector<Real>(MeshData<Real> *)> hst_fun;
  std::vector<std::string> labels;
};
inline const std::string hist_param_key = "HistoryFunctions";
inline const std::string hist_vec_param_key = "HistoryVectorFunctions";

struct FieldEntry {
  std::string name;
  Metadata m;
  int sparse_id = -1; // >= 0 for members of a sparse pool ("base_name_<id>")
};

class StateDescriptor {
 public:
  explicit StateDescriptor(std::string label) : label_(std::move(label)) {}
  const std::string &label() const { return label_; }

  template <typename T>
  void AddParam(const std::string &key, T value, bool is_mutable = false) {
    params_.Add<T>(key, std::move(value), is_mutable);
  }
  template <typename T>
  void UpdateParam(const std::string &key, T value) {
    params_.Update<T>(key, std::move(value));
  }
  template <typename T>
  const T &Param(const std::string &key) const {
    return params_.Get<T>(key);
  }
  Params &AllParams() { return params_; }
  const Params &AllParams() const { return params_; }

  // state_descriptor.hpp:172
  bool AddField(const std::string &field_name, const Metadata &m) {
    for (auto &f : fields_)
      if (f.name == field_name) return false;
    // the flux of a face field is a field of its own, one topological type up: an edge field
    // "bnd_flux::<name>" with Metadata::Flux, shared between the containers
    // (state_descriptor.cpp:313-318, metadata.cpp:175-205).  Cell-centred fields keep their
    // face fluxes inside the Variable (Variable::flux).
    if (m.IsSet(Metadata::Face) && m.IsSet(Metadata::WithFluxes)) {
      Metadata fm({Metadata::Edge, Metadata::Flux, Metadata::OneCopy, Metadata::Derived}, m.Shape());
      fields_.push_back(FieldEntry{"bnd_flux::" + field_name, fm, -1});
    }
    fields_.push_back(FieldEntry{field_name, m, -1});
    return true;
  }
  // state_descriptor.hpp:183-195: one field "base_<id>" per sparse id
  bool AddSparsePool(const std::string &base_name, const Metadata &m_in,
                     const std::vector<int> &sparse_ids) {
    Metadata m = m_in;
    m.Set(Metadata::Sparse);
    for (int id : sparse_ids) {
      const std::string name = base_name + "_" + std::to_string(id);
      for (auto &f : fields_)
        if (f.name == name) return false;
      fields_.push_back(FieldEntry{name, m, id});
    }
    return true;
  }
  const std::vector<FieldEntry> &AllFields() const { return fields_; }
  bool FieldPresent(const std::string &name) const {
    for (auto &f : fields_)
      if (f.name == name) return true;
    return false;
  }

  // hooks, state_descriptor.hpp:368-394
  std::function<void(MeshData<Real> *)> PreCommFillDerivedMesh = nullptr;
  std::function<void(MeshData<Real> *)> PreFillDerivedMesh = nullptr;
  std::function<void(MeshData<Real> *)> FillDerivedMesh = nullptr;
  std::function<void(MeshData<Real> *)> PostFillDerivedMesh = nullptr;
  std::function<void(MeshBlockData<Real> *)> PreFillDerivedBlock = nullptr;
  std::function<void(MeshBlockData<Real> *)> FillDerivedBlock = nullptr;
  std::function<void(MeshBlockData<Real> *)> PostFillDerivedBlock = nullptr;
  std::function<Real(MeshData<Real> *)> EstimateTimestepMesh = nullptr;
  std::function<Real(MeshBlockData<Real> *)> EstimateTimestepBlock = nullptr;
  std::function<AmrTag(MeshBlockData<Real> *)> CheckRefinementBlock = nullptr;
  // batch form: one tag per block of the MeshData (one device reduction for all blocks)
  std::function<void(MeshData<Real> *, std::vector<AmrTag> &)> CheckRefinementMesh = nullptr;
  std::function<void(MeshData<Real> *)> InitNewlyAllocatedVarsMesh = nullptr;
  std::function<void(Mesh *, ParameterInput *, SimTime &)> UserWorkBeforeLoopMesh = nullptr;
  std::function<void(SimTime const &, MeshData<Real> *)> PreStepDiagnosticsMesh = nullptr;
  std::function<void(SimTime const &, MeshData<Real> *)> PostStepDiagnosticsMesh = nullptr;

 private:
  std::string label_;
  Params params_;
  std::vector<FieldEntry> fields_;
};

class Packages_t {
 public:
  void Add(const std::shared_ptr<StateDescriptor> &pkg) {
    PARTHENON_REQUIRE(!packages_.count(pkg->label()),
                      "Package name " + pkg->label() + " must be unique.");
    packages_[pkg->label()] = pkg;
    order_.push_back(pkg->label());
  }
  const std::shared_ptr<StateDescriptor> &Get(const std::string &name) const {
    auto it = packages_.find(name);
    PARTHENON_REQUIRE(it != packages_.end(), "Package " + name + " doesn't exist");
    return it->second;
  }
  const std::map<std::string, std::shared_ptr<StateDescriptor>> &AllPackages() const {
    return packages_;
  }
  // registration order (field layout in HBM follows it)
  const std::vector<std::string> &Order() const { return order_; }

 private:
  std::map<std::string, std::shared_ptr<StateDescriptor>> packages_;
  std::vector<std::string> order_;
};

} // namespace parthenon
