// sparse_pack.hpp — SparsePack / MakePackDescriptor over the slab layout (host side).
//
// Mirrors src/interface/sparse_pack.hpp:53-420, sparse_pack_base.hpp:41-250 and
// make_pack_descriptor.hpp:42-128 of the reference: a descriptor selects variables of the
// registry by name (or regular expression), Metadata flags and options; GetPack(md) resolves it
// against one MeshData batch into the device tables of include/parthenon_b200_pack.h and hands
// back an object whose device half (`pb2::SparsePackView`) is passed to kernels by value:
//
//     auto desc = parthenon::MakePackDescriptor(pkg.get(), {"v5", "v3"}, {Metadata::WithFluxes});
//     auto pack = desc.GetPack(md);                    // cached in md, rebuilt when an
//     auto map = desc.GetMap();                        // allocation status changes
//     pb2::PackIdx iv3(map["v3"]);
//     my_kernel<<<...>>>(pack.view(), iv3);            // pack(b, iv3 + c, k, j, i)
//
// and with variable-name types (pack_utils.hpp:60-150):
//
//     struct v3 : parthenon::variable_names::base_t<false, 3> { ... static name() "v3" };
//     auto desc = parthenon::MakePackDescriptor<v5, v3>(pkg.get());
//     auto pack = desc.GetPack(md);                    // pack.view()(b, v3(c), k, j, i)
//
// Differences that follow from the layout: a pack holds fields of ONE topological type (all
// cell-centred, or all face / edge / node: their arrays have different extents), and entries
// are component pointers into the field slabs instead of view handles.
#pragma once
#include <map>
#include <memory>
#include <regex>
#include <set>
#include <string>
#include <vector>

#include "mesh_data.hpp"
#include "parthenon_b200_pack.h"

namespace parthenon {

using pb2::PackIdx;
enum class PDOpt { WithFluxes, Coarse, Flatten };

// name -> index of the variable group in the descriptor (sparse_pack_base.hpp:41-57)
class SparsePackIdxMap {
 public:
  const std::size_t &operator[](const std::string &key) const {
    auto it = map_.find(key);
    PARTHENON_REQUIRE(it != map_.end(), "Key " + key + " does not exist in SparsePackIdxMap");
    return it->second;
  }
  void insert(const std::string &key, std::size_t idx) { map_[key] = idx; }
  std::size_t size() const { return map_.size(); }

 private:
  std::map<std::string, std::size_t> map_;
};

namespace impl {
struct PackDescriptor {
  PackDescriptor() = default;
  using Selector = std::function<bool(int, const FieldEntry &)>;
  // fields of `packages` (registration order), grouped by the selector they match and sorted
  // by (base name, sparse id) inside a group (sparse_pack_base.hpp:194-218)
  PackDescriptor(const std::vector<const StateDescriptor *> &packages,
                 const std::vector<std::string> &group_names, const Selector &selector,
                 const std::set<PDOpt> &options);
  int nvar_groups = 0;
  std::vector<std::string> var_group_names;
  std::vector<std::vector<std::string>> var_groups; // field labels per group
  bool with_fluxes = false, coarse = false, flat = false;
  std::string identifier;
  std::size_t nvar_tot = 0;
};
} // namespace impl

// device tables + host mirrors of one resolved pack; owned by the MeshData's cache
struct SparsePackStorage {
  pb2_sparse_pack view{};
  DeviceBuffer ptr, bounds, coords, block_props;
  std::vector<int32_t> bounds_h;       // [2][nblocks_md][nvar + 1]
  std::vector<int32_t> block_props_h;  // [nblocks_md][28]: neighbour levels, gid
  std::vector<std::string> labels_h;   // [nblocks][maxvars]
  std::vector<uint8_t> alloc_status;   // allocation status it was built from
  std::vector<bool> include_block;
};

class SparsePack {
 public:
  SparsePack() = default;
  explicit SparsePack(std::shared_ptr<SparsePackStorage> s) : s_(std::move(s)) {}

  // what a kernel takes by value
  pb2::SparsePackView view() const { return pb2::SparsePackView(s_->view); }
  const pb2_sparse_pack &pod() const { return s_->view; }

  int GetNBlocks() const { return s_->view.nblocks; }
  int GetMaxNumberOfVars() const { return s_->view.maxvars; }
  int GetSize() const { return s_->view.size; }

  // host bound overloads (sparse_pack.hpp:169-196)
  int GetLowerBoundHost(int b) const {
    return (s_->view.flat && b > 0) ? bound(1, b - 1, s_->view.nvar) + 1 : 0;
  }
  int GetUpperBoundHost(int b) const { return bound(1, b, s_->view.nvar); }
  int GetLowerBoundHost(int b, PackIdx idx) const { return bound(0, b, idx.VariableIdx()); }
  int GetUpperBoundHost(int b, PackIdx idx) const { return bound(1, b, idx.VariableIdx()); }
  int GetSizeHost(int b, PackIdx idx) const {
    return GetUpperBoundHost(b, idx) - GetLowerBoundHost(b, idx) + 1;
  }
  bool ContainsHost(int b) const { return GetUpperBoundHost(b) >= 0; }
  bool ContainsHost(int b, PackIdx idx) const { return GetUpperBoundHost(b, idx) >= 0; }
  int GetLevelHost(int b, int off3, int off2, int off1) const {
    return s_->block_props_h[b * 28 + (off3 + 1) + 3 * ((off2 + 1) + 3 * (off1 + 1))];
  }
  int GetGIDHost(int b) const { return s_->block_props_h[b * 28 + 27]; }
  const std::string &LabelHost(int b, int idx) const {
    return s_->labels_h[static_cast<size_t>(b) * s_->view.maxvars + idx];
  }

  class Descriptor : public impl::PackDescriptor {
   public:
    Descriptor() = default;
    explicit Descriptor(const impl::PackDescriptor &d) : impl::PackDescriptor(d) {}
    // resolved against md; taken from md's pack cache unless an allocation status (or the
    // block selection) changed since it was built (sparse_pack_base.cpp:320-370)
    SparsePack GetPack(MeshData<Real> *md, const std::vector<bool> &include_block = {}) const;
    SparsePackIdxMap GetMap() const {
      SparsePackIdxMap m;
      for (int i = 0; i < nvar_groups; ++i) m.insert(var_group_names[i], i);
      return m;
    }
  };

 protected:
  int bound(int which, int b, int v) const {
    const auto &p = s_->view;
    return s_->bounds_h[(static_cast<size_t>(which) * p.nblocks_md + b) * (p.nvar + 1) + v];
  }
  std::shared_ptr<SparsePackStorage> s_;
};

// ---- variable-name types (pack_utils.hpp:60-150) ------------------------------------------------
namespace variable_names {
constexpr int ANYDIM = -1234; // must be the slowest-moving index
template <bool REGEX, int... NCOMP>
struct base_t {
  PB2_PACK_HD base_t() : idx(0) {}
  PB2_PACK_HD explicit base_t(int idx1) : idx(idx1) {}
  template <typename... Args,
            typename = std::enable_if_t<(sizeof...(Args) == sizeof...(NCOMP)) &&
                                        (sizeof...(Args) > 1)>>
  PB2_PACK_HD explicit base_t(Args... args) : idx(GetIndex_(args...)) {}
  static bool regex() { return REGEX; }
  static std::vector<int> GetShape() { return std::vector<int>{NCOMP...}; }
  const int idx;

 private:
  template <class... Args>
  PB2_PACK_HD static int GetIndex_(Args... args) {
    int i = 0;
    const int dims[] = {NCOMP...};
    const int a[] = {static_cast<int>(args)...};
    for (unsigned d = 0; d < sizeof...(NCOMP); ++d) i = i * dims[d] + a[d];
    return i;
  }
};
struct any : public base_t<true> {
  template <class... Ts>
  PB2_PACK_HD any(Ts &&...args) : base_t<true>(static_cast<Ts &&>(args)...) {}
  static std::string name() { return ".*"; }
};
} // namespace variable_names

namespace impl {
template <class T, class... Ts>
struct TypeIdx;
template <class T, class... Ts>
struct TypeIdx<T, T, Ts...> { static constexpr int value = 0; };
template <class T, class U, class... Ts>
struct TypeIdx<T, U, Ts...> { static constexpr int value = 1 + TypeIdx<T, Ts...>::value; };
} // namespace impl

// device view with the type-based accessors: pack(b, v3(c), k, j, i), GetLowerBound(b, v3())
template <class... Ts>
struct TypedPackView : pb2::SparsePackView {
  TypedPackView() = default;
  PB2_PACK_HD explicit TypedPackView(const pb2_sparse_pack &p) : pb2::SparsePackView(p) {}
  using pb2::SparsePackView::operator();
  using pb2::SparsePackView::Contains;
  using pb2::SparsePackView::flux;
  using pb2::SparsePackView::GetLowerBound;
  using pb2::SparsePackView::GetSize;
  using pb2::SparsePackView::GetUpperBound;
  template <class T, int I = impl::TypeIdx<T, Ts...>::value>
  PB2_PACK_HD int GetLowerBound(int b, const T &) const { return bound(0, b, I); }
  template <class T, int I = impl::TypeIdx<T, Ts...>::value>
  PB2_PACK_HD int GetUpperBound(int b, const T &) const { return bound(1, b, I); }
  template <class T, int I = impl::TypeIdx<T, Ts...>::value>
  PB2_PACK_HD int GetSize(int b, const T &t) const {
    return GetUpperBound(b, t) - GetLowerBound(b, t) + 1;
  }
  template <class T, int I = impl::TypeIdx<T, Ts...>::value>
  PB2_PACK_HD bool Contains(int b, const T &t) const { return GetUpperBound(b, t) >= 0; }
  template <class T, int I = impl::TypeIdx<T, Ts...>::value>
  PB2_PACK_HD int GetIndex(int b, const T &t) const { return bound(0, b, I) + t.idx; }
  template <class T, int I = impl::TypeIdx<T, Ts...>::value>
  PB2_PACK_HD double &operator()(int b, const T &t, int k, int j, int i) const {
    return Component(0, b, bound(0, b, I) + t.idx)[Cell(k, j, i)];
  }
  template <class T, int I = impl::TypeIdx<T, Ts...>::value>
  PB2_PACK_HD double &flux(int b, int dir, const T &t, int k, int j, int i) const {
    return Component(dir, b, bound(0, b, I) + t.idx)[Cell(k, j, i)];
  }
};

template <class... Ts>
class TypedSparsePack : public SparsePack {
 public:
  TypedSparsePack() = default;
  explicit TypedSparsePack(const SparsePack &p) : SparsePack(p) {}
  TypedPackView<Ts...> view() const { return TypedPackView<Ts...>(s_->view); }
  using SparsePack::ContainsHost;
  using SparsePack::GetLowerBoundHost;
  using SparsePack::GetSizeHost;
  using SparsePack::GetUpperBoundHost;
  template <class T, int I = impl::TypeIdx<T, Ts...>::value>
  int GetLowerBoundHost(int b, const T &) const { return bound(0, b, I); }
  template <class T, int I = impl::TypeIdx<T, Ts...>::value>
  int GetUpperBoundHost(int b, const T &) const { return bound(1, b, I); }
  template <class T, int I = impl::TypeIdx<T, Ts...>::value>
  int GetSizeHost(int b, const T &t) const {
    return GetUpperBoundHost(b, t) - GetLowerBoundHost(b, t) + 1;
  }
  template <class T, int I = impl::TypeIdx<T, Ts...>::value>
  bool ContainsHost(int b, const T &t) const { return GetUpperBoundHost(b, t) >= 0; }
  template <class T1, class T2, class... Rest>
  bool ContainsHost(int b, const T1 &t1, const T2 &t2, const Rest &...rest) const {
    return ContainsHost(b, t1) && ContainsHost(b, t2, rest...);
  }
  template <class... Args, typename = std::enable_if_t<(sizeof...(Args) > 0)>>
  bool ContainsHost(int b) const {
    return (... && ContainsHost(b, Args()));
  }

  class Descriptor : public SparsePack::Descriptor {
   public:
    Descriptor() = default;
    explicit Descriptor(const impl::PackDescriptor &d) : SparsePack::Descriptor(d) {}
    TypedSparsePack GetPack(MeshData<Real> *md, const std::vector<bool> &include_block = {}) const {
      return TypedSparsePack(SparsePack::Descriptor::GetPack(md, include_block));
    }
  };
};

// ---- make_pack_descriptor.hpp:42-128 ------------------------------------------------------------
SparsePack::Descriptor MakePackDescriptor(const std::vector<const StateDescriptor *> &packages,
                                          const std::vector<std::string> &vars,
                                          const std::vector<bool> &use_regex,
                                          const std::vector<MetadataFlag> &flags = {},
                                          const std::set<PDOpt> &options = {});
inline SparsePack::Descriptor MakePackDescriptor(StateDescriptor *psd,
                                                 const std::vector<std::string> &vars,
                                                 const std::vector<bool> &use_regex,
                                                 const std::vector<MetadataFlag> &flags = {},
                                                 const std::set<PDOpt> &options = {}) {
  return MakePackDescriptor(std::vector<const StateDescriptor *>{psd}, vars, use_regex, flags,
                            options);
}
inline SparsePack::Descriptor MakePackDescriptor(StateDescriptor *psd,
                                                 const std::vector<std::string> &vars,
                                                 const std::vector<MetadataFlag> &flags = {},
                                                 const std::set<PDOpt> &options = {}) {
  return MakePackDescriptor(psd, vars, std::vector<bool>(vars.size(), false), flags, options);
}
// every package of the mesh (the reference's resolved_packages)
SparsePack::Descriptor MakePackDescriptor(MeshData<Real> *md, const std::vector<std::string> &vars,
                                          const std::vector<MetadataFlag> &flags = {},
                                          const std::set<PDOpt> &options = {});
template <class... Ts>
inline typename TypedSparsePack<Ts...>::Descriptor
MakePackDescriptor(StateDescriptor *psd, const std::vector<MetadataFlag> &flags = {},
                   const std::set<PDOpt> &options = {}) {
  static_assert(sizeof...(Ts) > 0, "Must have at least one variable type for type pack");
  return typename TypedSparsePack<Ts...>::Descriptor(static_cast<impl::PackDescriptor>(
      MakePackDescriptor(psd, std::vector<std::string>{Ts::name()...},
                         std::vector<bool>{Ts::regex()...}, flags, options)));
}

} // namespace parthenon
