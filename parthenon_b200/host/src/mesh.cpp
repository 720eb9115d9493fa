// mesh.cpp — mesh topology: tree of leaves, Morton order, rank assignment, neighbour lists.
// See pb2/mesh.hpp for the reference files this follows.
#include "pb2/mesh.hpp"

#include <algorithm>
#include <cmath>
#include <numeric>

#include "pb2/mesh_data.hpp"

namespace parthenon {

namespace Globals {
int my_rank = 0, nranks = 1, nghost = 2;
}

uint64_t LogicalLocation::MortonKey(int maxlevel) const {
  uint64_t key = 0;
  const int sh = maxlevel - level;
  for (int bit = 0; bit < maxlevel; ++bit)
    for (int d = 0; d < 3; ++d) {
      const uint64_t c = static_cast<uint64_t>(lx[d]) << sh;
      key |= ((c >> bit) & 1ull) << (3 * bit + d);
    }
  return key;
}

std::array<int, 3> LogicalLocation::GetSameLevelOffsets(const LogicalLocation &nb) const {
  std::array<int, 3> off;
  const int sn = std::max(nb.level - level, 0), sm = std::max(level - nb.level, 0);
  for (int d = 0; d < 3; ++d) off[d] = static_cast<int>((nb.lx[d] >> sn) - (lx[d] >> sm));
  return off;
}

bool LogicalLocation::IsNeighbor(const LogicalLocation &in) const {
  const int max_level = std::max(in.level, level);
  const int64_t bs_in = int64_t{1} << (max_level - in.level);
  const int64_t bs_this = int64_t{1} << (max_level - level);
  for (int d = 0; d < 3; ++d) {
    const int64_t low = lx[d] * bs_this - 1, hi = low + bs_this + 1;
    const int64_t low_in = in.lx[d] * bs_in, hi_in = low_in + bs_in - 1;
    if (hi < low_in || low > hi_in) return false;
  }
  return true;
}

bool LogicalLocation::IsNeighborOfTE(const LogicalLocation &in,
                                     const std::array<int, 3> &te_offset) const {
  const int max_level = std::max(in.level, level);
  const int64_t bs_in = int64_t{1} << (max_level - in.level);
  const int64_t bs_this = int64_t{1} << (max_level - level);
  for (int d = 0; d < 3; ++d) {
    int64_t low = lx[d] * bs_this, hi = low + bs_this - 1;
    if (te_offset[d] == -1) {
      low -= 1;
      hi = low + 1;
    } else if (te_offset[d] == 1) {
      hi += 1;
      low = hi - 1;
    }
    const int64_t low_in = in.lx[d] * bs_in, hi_in = low_in + bs_in - 1;
    if (hi < low_in || low > hi_in) return false;
  }
  return true;
}

namespace {
// logical_location.cpp:61-74: position in [-0.5, 0.5] built from integers so that the
// mesh is bitwise symmetric about its centre
Real SymmetrizedCoordinate(int64_t index, int bloc, int64_t nrange) {
  const int64_t noffset = index - nrange / 2;
  const int64_t noffset_ceil = index - (nrange + 1) / 2;
  return static_cast<Real>(noffset + noffset_ceil + static_cast<int64_t>(bloc)) /
         (2.0 * static_cast<Real>(nrange));
}
// defs.hpp:98-101
Real LogicalToActual(Real u, Real xmin, Real xmax) {
  return static_cast<Real>(0.5) * (xmin + xmax) + (u * xmax - u * xmin);
}
// anything that is not a stock condition names a user-registered one
// (ApplicationInput::RegisterBoundaryCondition)
BoundaryFlag ParseBoundary(const std::string &s) {
  if (s == "periodic") return BoundaryFlag::periodic;
  if (s == "outflow") return BoundaryFlag::outflow;
  if (s == "reflecting" || s == "reflect") return BoundaryFlag::reflect;
  return BoundaryFlag::user;
}
} // namespace

void Mesh::AssignBlocks(const std::vector<double> &costlist, int nranks,
                        std::vector<int> &ranklist) {
  // equal-cost contiguous gid ranges filled from the last rank backwards, so rank 0 gets
  // the lighter share (mesh-amr_loadbalance.cpp:362-386)
  ranklist.assign(costlist.size(), 0);
  const double total = std::accumulate(costlist.begin(), costlist.end(), 0.0);
  int rank = nranks - 1;
  double target = total / nranks, mine = 0.0, remaining = total;
  for (int b = static_cast<int>(costlist.size()) - 1; b >= 0; --b) {
    PARTHENON_REQUIRE(target != 0.0, "There is at least one process which has no MeshBlock");
    mine += costlist[b];
    ranklist[b] = rank;
    if (mine >= target && rank > 0) {
      --rank;
      remaining -= mine;
      mine = 0.0;
      target = remaining / (rank + 1);
    }
  }
}

int64_t Mesh::BlocksAtLevel(int level, int d) const {
  if (d >= ndim) return 1;
  return static_cast<int64_t>(nrbx[d]) << (level - root_level);
}

bool Mesh::WrapLocation(const LogicalLocation &in, LogicalLocation &out) const {
  out = in;
  for (int d = 0; d < 3; ++d) {
    const int64_t n = BlocksAtLevel(in.level, d);
    if (in.lx[d] < 0 || in.lx[d] >= n) {
      if (d >= ndim) return false;
      if (mesh_bcs[2 * d] != BoundaryFlag::periodic) return false;
      out.lx[d] = ((in.lx[d] % n) + n) % n;
    }
  }
  return true;
}

RegionSize Mesh::GetBlockSize(const LogicalLocation &loc) const {
  RegionSize rs = base_block_size;
  for (int d = 0; d < 3; ++d) {
    if (d < ndim) {
      // blocks spanning the mesh at this level.  Single tree (cubic 2^n root grid): 2^level,
      // exactly the reference's tree.cpp:297-312; other root grids use the same formula on
      // the mesh as a whole
      const int64_t ntot = BlocksAtLevel(std::max(loc.level, root_level), d);
      rs.xmin_[d] = LogicalToActual(SymmetrizedCoordinate(loc.lx[d], 0, ntot), mesh_size.xmin_[d],
                                    mesh_size.xmax_[d]);
      rs.xmax_[d] = LogicalToActual(SymmetrizedCoordinate(loc.lx[d], 2, ntot), mesh_size.xmin_[d],
                                    mesh_size.xmax_[d]);
    } else {
      rs.xmin_[d] = mesh_size.xmin_[d];
      rs.xmax_[d] = mesh_size.xmax_[d];
    }
  }
  return rs;
}

void Mesh::BuildTree(ParameterInput *pin, const std::vector<LogicalLocation> &leaves_in) {
  std::vector<LogicalLocation> leaves = leaves_in;
  if (leaves.empty()) {
    for (int64_t k = 0; k < nrbx[2]; ++k)
      for (int64_t j = 0; j < nrbx[1]; ++j)
        for (int64_t i = 0; i < nrbx[0]; ++i) {
          LogicalLocation l;
          l.level = root_level;
          l.lx[0] = i;
          l.lx[1] = j;
          l.lx[2] = k;
          leaves.push_back(l);
        }
    // <parthenon/static_refinementN>: refine every leaf that overlaps the region down to
    // the requested level (mesh.cpp:1084-1170), then restore 2:1 nesting
    std::unordered_map<LogicalLocation, int, LogicalLocationHash> set;
    for (auto &l : leaves) set[l] = 1;
    auto refine = [&](const LogicalLocation &l) {
      set.erase(l);
      for (int dk = 0; dk < (ndim > 2 ? 2 : 1); ++dk)
        for (int dj = 0; dj < (ndim > 1 ? 2 : 1); ++dj)
          for (int di = 0; di < 2; ++di) {
            LogicalLocation c;
            c.level = l.level + 1;
            c.lx[0] = 2 * l.lx[0] + di;
            c.lx[1] = ndim > 1 ? 2 * l.lx[1] + dj : 0;
            c.lx[2] = ndim > 2 ? 2 * l.lx[2] + dk : 0;
            set[c] = 1;
          }
    };
    bool any_static = false;
    for (auto &bname : pin->BlockNames()) {
      if (bname.compare(0, 27, "parthenon/static_refinement") != 0) continue;
      any_static = true;
      Real rmin[3], rmax[3];
      for (int d = 0; d < 3; ++d) {
        const std::string x = "x" + std::to_string(d + 1);
        rmin[d] = d < ndim ? pin->GetReal(bname, x + "min") : mesh_size.xmin_[d];
        rmax[d] = d < ndim ? pin->GetReal(bname, x + "max") : mesh_size.xmax_[d];
      }
      const int ref_lev = pin->GetInteger(bname, "level");
      PARTHENON_REQUIRE(ref_lev >= 1, "Refinement level must be larger than 0 (root level = 0)");
      for (int lev = root_level; lev < root_level + ref_lev; ++lev) {
        std::vector<LogicalLocation> todo;
        for (auto &kv : set) {
          if (kv.first.level != lev) continue;
          const RegionSize rs = GetBlockSize(kv.first);
          bool overlap = true;
          for (int d = 0; d < ndim; ++d)
            overlap = overlap && rs.xmax_[d] > rmin[d] && rs.xmin_[d] < rmax[d];
          if (overlap) todo.push_back(kv.first);
        }
        for (auto &l : todo) refine(l);
      }
    }
    if (any_static) {
      // 2:1 balance: a leaf must not touch a leaf more than one level finer
      bool changed = true;
      while (changed) {
        changed = false;
        std::vector<LogicalLocation> cur;
        for (auto &kv : set) cur.push_back(kv.first);
        for (auto &l : cur) {
          if (!set.count(l) || l.level <= root_level + 0) continue;
          // every neighbour position of l's parent must exist at level >= l.level - 1
          const LogicalLocation par = l.GetParent();
          for (int o3 = (ndim > 2 ? -1 : 0); o3 <= (ndim > 2 ? 1 : 0); ++o3)
            for (int o2 = (ndim > 1 ? -1 : 0); o2 <= (ndim > 1 ? 1 : 0); ++o2)
              for (int o1 = -1; o1 <= 1; ++o1) {
                LogicalLocation n = par, w;
                n.lx[0] += o1;
                n.lx[1] += o2;
                n.lx[2] += o3;
                if (!WrapLocation(n, w)) continue;
                // walk up: if an ancestor of w (coarser than w) is a leaf, refine it
                LogicalLocation a = w;
                while (a.level > root_level) {
                  a = a.GetParent();
                  if (set.count(a)) {
                    refine(a);
                    changed = true;
                    break;
                  }
                }
              }
        }
      }
      leaves.clear();
      for (auto &kv : set) leaves.push_back(kv.first);
    }
  }
  int maxlevel = 0;
  for (auto &l : leaves) maxlevel = std::max(maxlevel, l.level);
  std::vector<std::pair<uint64_t, LogicalLocation>> ent;
  ent.reserve(leaves.size());
  for (auto &l : leaves) ent.emplace_back(l.MortonKey(maxlevel), l);
  std::sort(ent.begin(), ent.end(), [](const auto &a, const auto &b) {
    if (a.first != b.first) return a.first < b.first;
    return a.second.level < b.second.level;
  });
  loclist.clear();
  leaf_gid_.clear();
  ownership_.clear();
  internal_.clear();
  current_level = maxlevel;
  multilevel = false;
  for (size_t g = 0; g < ent.size(); ++g) {
    loclist.push_back(ent[g].second);
    leaf_gid_[ent[g].second] = static_cast<int>(g);
    if (ent[g].second.level != root_level) multilevel = true;
    LogicalLocation a = ent[g].second;
    while (a.level > 0) {
      a = a.GetParent();
      internal_[a] = 1;
    }
  }
  nbtotal = static_cast<int>(loclist.size());
}

// neighbour search on the leaf grid, tree.cpp:139-226; offsets iterate ox1 slowest like the
// reference's 3-D indexer (indexer.hpp:117-144)
void Mesh::FindNeighbors(MeshBlock &mb) const { mb.neighbors = FindNeighbors(mb.loc); }

// ownership of shared faces / edges / nodes: the block of the highest level, then of the
// highest (tree, Morton) number — on leaves that is the highest gid of the level — owns an
// element (block_ownership.cpp:42-83, without newly refined blocks)
const std::array<bool, 27> &Mesh::Ownership(int gid) const {
  auto it = ownership_.find(gid);
  if (it != ownership_.end()) return it->second;
  const LogicalLocation &loc = loclist[gid];
  const std::vector<NeighborBlock> nbs = FindNeighbors(loc);
  std::array<bool, 27> owns;
  for (int o1 = -1; o1 <= 1; ++o1)
    for (int o2 = -1; o2 <= 1; ++o2)
      for (int o3 = -1; o3 <= 1; ++o3) {
        bool own = true;
        for (const NeighborBlock &n : nbs) {
          const int la = 2 * loc.level - (newly_refined_.count(loc) ? 1 : 0);
          const int lb = 2 * n.loc.level - (newly_refined_.count(n.loc) ? 1 : 0);
          const bool less = la != lb ? la < lb : gid < n.gid;
          if (less && loc.IsNeighborOfTE(n.origin_loc, {o1, o2, o3})) {
            own = false;
            break;
          }
        }
        owns[(o1 + 1) + 3 * (o2 + 1) + 9 * (o3 + 1)] = own;
      }
  return ownership_.emplace(gid, owns).first->second;
}

std::vector<NeighborBlock> Mesh::FindNeighbors(const LogicalLocation &loc) const {
  struct {
    std::vector<NeighborBlock> neighbors;
  } mb;
  if (forest) {
    // mesh-gmg.cpp:48-101 over Forest::FindNeighbors: offsets come from the neighbour's location
    // in THIS block's index space, the transformation says how to read what it sends
    for (const forest::NeighborLocation &nl : forest->FindNeighbors(loc)) {
      NeighborBlock nb;
      nb.gid = leaf_gid_.at(nl.global_loc);
      nb.rank = ranklist[nb.gid];
      nb.lid = nb.gid - nslist[nb.rank];
      nb.loc = nl.global_loc;
      nb.origin_loc = nl.origin_loc;
      const auto off = loc.GetSameLevelOffsets(nl.origin_loc);
      for (int d = 0; d < 3; ++d) nb.offsets[d] = off[d];
      nb.lcoord_trans = nl.lcoord_trans;
      nb.transformed = !nl.lcoord_trans.IsIdentity();
      mb.neighbors.push_back(nb);
    }
    return mb.neighbors;
  }
  auto add = [&](int gid, const LogicalLocation &wrapped, const LogicalLocation &origin) {
    NeighborBlock nb;
    nb.gid = gid;
    nb.rank = ranklist[gid];
    nb.lid = gid - nslist[nb.rank];
    nb.loc = wrapped;
    nb.origin_loc = origin;
    const auto off = loc.GetSameLevelOffsets(origin); // mesh-gmg.cpp:64
    for (int d = 0; d < 3; ++d) nb.offsets[d] = off[d];
    mb.neighbors.push_back(nb);
  };
  int lo[3], hi[3];
  for (int d = 0; d < 3; ++d) {
    lo[d] = d < ndim ? -1 : 0;
    hi[d] = d < ndim ? 1 : 0;
  }
  for (int o1 = lo[0]; o1 <= hi[0]; ++o1)
    for (int o2 = lo[1]; o2 <= hi[1]; ++o2)
      for (int o3 = lo[2]; o3 <= hi[2]; ++o3) {
        if (o1 == 0 && o2 == 0 && o3 == 0) continue;
        LogicalLocation neigh = loc, w;
        neigh.lx[0] += o1;
        neigh.lx[1] += o2;
        neigh.lx[2] += o3;
        if (!WrapLocation(neigh, w)) continue;
        auto leaf = leaf_gid_.find(w);
        if (leaf != leaf_gid_.end()) {
          add(leaf->second, w, neigh);
        } else if (internal_.count(w)) {
          // finer neighbours: the daughters of the position that touch this block
          for (int d3 = 0; d3 < (ndim > 2 ? 2 : 1); ++d3)
            for (int d2 = 0; d2 < (ndim > 1 ? 2 : 1); ++d2)
              for (int d1 = 0; d1 < 2; ++d1) {
                LogicalLocation dn, dw;
                dn.level = neigh.level + 1;
                dn.lx[0] = (neigh.lx[0] << 1) + d1;
                dn.lx[1] = ndim > 1 ? (neigh.lx[1] << 1) + d2 : 0;
                dn.lx[2] = ndim > 2 ? (neigh.lx[2] << 1) + d3 : 0;
                if (!loc.IsNeighbor(dn)) continue;
                WrapLocation(dn, dw);
                auto dl = leaf_gid_.find(dw);
                PARTHENON_REQUIRE(dl != leaf_gid_.end(), "mesh violates 2:1 nesting");
                add(dl->second, dw, dn);
              }
        } else {
          // coarser neighbour: the parent of the position, if it sits at this offset
          LogicalLocation par = neigh.GetParent(), pw;
          if (ndim < 2) par.lx[1] = 0;
          if (ndim < 3) par.lx[2] = 0;
          if (!WrapLocation(par, pw)) continue;
          auto pl = leaf_gid_.find(pw);
          if (pl == leaf_gid_.end()) continue;
          const auto so = loc.GetSameLevelOffsets(par);
          if (so[0] == o1 && so[1] == o2 && so[2] == o3) add(pl->second, pw, par);
        }
      }
  return mb.neighbors;
}

Mesh::Mesh(ParameterInput *pin, ApplicationInput *app_in, Packages_t &pkgs, int rank, int nranks_in,
           const std::vector<LogicalLocation> &leaves)
    : my_rank(rank), nranks(nranks_in), packages(pkgs) {
  Globals::my_rank = rank;
  Globals::nranks = nranks_in;
  Globals::nghost = pin->GetOrAddInteger("parthenon/mesh", "nghost", 2);
  const char *bc_names[6] = {"ix1_bc", "ox1_bc", "ix2_bc", "ox2_bc", "ix3_bc", "ox3_bc"};
  for (int d = 0; d < 3; ++d) {
    const std::string n = std::to_string(d + 1);
    mesh_size.nx_[d] = pin->GetOrAddInteger("parthenon/mesh", "nx" + n, 1);
    mesh_size.xmin_[d] = pin->GetOrAddReal("parthenon/mesh", "x" + n + "min", -0.5);
    mesh_size.xmax_[d] = pin->GetOrAddReal("parthenon/mesh", "x" + n + "max", 0.5);
  }
  ndim = mesh_size.nx_[2] > 1 ? 3 : (mesh_size.nx_[1] > 1 ? 2 : 1);
  for (int d = 0; d < 3; ++d) {
    mesh_size.symmetry_[d] = d >= ndim;
    const std::string n = std::to_string(d + 1);
    base_block_size.nx_[d] =
        d < ndim ? pin->GetOrAddInteger("parthenon/meshblock", "nx" + n, mesh_size.nx_[d]) : 1;
    base_block_size.symmetry_[d] = d >= ndim;
    PARTHENON_REQUIRE(mesh_size.nx_[d] % base_block_size.nx_[d] == 0,
                      "the Mesh must be evenly divisible by the MeshBlock");
    nrbx[d] = mesh_size.nx_[d] / base_block_size.nx_[d];
    mesh_bcs[2 * d] = ParseBoundary(pin->GetOrAddString("parthenon/mesh", bc_names[2 * d], "periodic"));
    mesh_bcs[2 * d + 1] =
        ParseBoundary(pin->GetOrAddString("parthenon/mesh", bc_names[2 * d + 1], "periodic"));
    if (d < ndim) {
      for (int f = 2 * d; f < 2 * d + 2; ++f) {
        if (mesh_bcs[f] != BoundaryFlag::user) continue;
        const std::string name = pin->GetString("parthenon/mesh", bc_names[f]);
        const bool have = app_in != nullptr && app_in->boundary_conditions_[f].count(name) > 0;
        PARTHENON_REQUIRE(have, "boundary condition '" + name + "' of " + bc_names[f] +
                                    " is neither periodic / outflow / reflecting nor registered "
                                    "with ApplicationInput::RegisterBoundaryCondition");
        user_bcs[f] = app_in->boundary_conditions_[f].at(name);
      }
      PARTHENON_REQUIRE((mesh_bcs[2 * d] == BoundaryFlag::periodic) ==
                            (mesh_bcs[2 * d + 1] == BoundaryFlag::periodic),
                        "a direction is periodic on both faces or on neither");
    }
  }
  // Leaves are ordered by their Morton key in the smallest 2^n cube that holds the root
  // grid.  For a cubic 2^n root grid this is the reference's single-tree order
  // (forest.cpp:73-145); elongated root grids (512x256x256 ...) come out as consecutive
  // cubic sub-trees, which is what keeps each rank's gid range a compact brick.
  int maxrb = 1;
  for (int d = 0; d < ndim; ++d) maxrb = std::max(maxrb, nrbx[d]);
  root_level = 0;
  while ((1 << root_level) < maxrb) ++root_level;
  ReadKnobs(pin);
  BuildTree(pin, leaves);
  FinishConstruction(pin);
}

// Mesh::Mesh(pin, app_in, packages, forest_def), mesh.cpp:189-218
Mesh::Mesh(ParameterInput *pin, ApplicationInput *app_in, Packages_t &pkgs,
           const forest::ForestDefinition &forest_def, int rank, int nranks_in)
    : my_rank(rank), nranks(nranks_in), packages(pkgs) {
  Globals::my_rank = rank;
  Globals::nranks = nranks_in;
  Globals::nghost = pin->GetOrAddInteger("parthenon/mesh", "nghost", 2);
  ndim = 2;
  mesh_size = RegionSize(); // {0,0,0}-{1,1,0}: the trees carry their own domains
  mesh_size.symmetry_ = {false, false, true};
  const char *bc_names[6] = {"ix1_bc", "ox1_bc", "ix2_bc", "ox2_bc", "ix3_bc", "ox3_bc"};
  for (int d = 0; d < 3; ++d) {
    const std::string n = std::to_string(d + 1);
    base_block_size.nx_[d] = d < ndim ? pin->GetOrAddInteger("parthenon/meshblock", "nx" + n, 1) : 1;
    base_block_size.symmetry_[d] = d >= ndim;
    nrbx[d] = 1;
  }
  PARTHENON_REQUIRE(base_block_size.nx_[0] == base_block_size.nx_[1],
                    "blocks of a forest must be square: a neighbouring tree may swap the axes");
  // edges flagged `user` take the deck's condition of that face, outflow unless named otherwise
  // (Mesh::SetBCNames_ mesh.cpp:1179-1186, Tree::EnrollBndryFncts tree.cpp:388-406)
  for (int f = 0; f < 6; ++f) {
    const std::string name = pin->GetOrAddString("parthenon/mesh", bc_names[f], "outflow");
    mesh_bcs[f] = f < 4 ? ParseBoundary(name) : BoundaryFlag::periodic;
    PARTHENON_REQUIRE(f >= 4 || mesh_bcs[f] != BoundaryFlag::periodic,
                      "periodic is not a condition for the outer edges of a forest");
    if (f < 4 && mesh_bcs[f] == BoundaryFlag::user) {
      const bool have = app_in != nullptr && app_in->boundary_conditions_[f].count(name) > 0;
      PARTHENON_REQUIRE(have, "boundary condition '" + name + "' of " + bc_names[f] +
                                  " is neither outflow / reflecting nor registered with "
                                  "ApplicationInput::RegisterBoundaryCondition");
      user_bcs[f] = app_in->boundary_conditions_[f].at(name);
    }
  }
  root_level = 0;
  ReadKnobs(pin);
  PARTHENON_REQUIRE(!adaptive, "adaptive refinement of forests is not built");
  forest = std::make_shared<forest::Forest>(forest_def);
  loclist = forest->GetMeshBlockList();
  leaf_gid_.clear();
  current_level = 0;
  for (size_t g = 0; g < loclist.size(); ++g) {
    leaf_gid_[loclist[g]] = static_cast<int>(g);
    current_level = std::max(current_level, loclist[g].level);
  }
  multilevel = current_level > 0;
  nbtotal = static_cast<int>(loclist.size());
  FinishConstruction(pin);
}

void Mesh::ReadKnobs(ParameterInput *pin) {
  const std::string refinement = pin->GetOrAddString("parthenon/mesh", "refinement", "none");
  adaptive = refinement == "adaptive";
  pack_size_ = pin->GetOrAddInteger("parthenon/mesh", "pack_size", -1);
  virtual_ranks = pin->GetOrAddInteger("pb2", "virtual_ranks", 1);
  table_halo = pin->GetOrAddBoolean("pb2", "table_halo", false);
  peer_push = pin->GetOrAddBoolean("pb2", "peer_push", nranks > 1);
  {
    const std::string mode = pin->GetOrAddString("pb2", "peer_push_mode", "ce");
    PARTHENON_REQUIRE(mode == "ce" || mode == "sm" || mode == "direct",
                      "pb2/peer_push_mode must be ce, sm or direct");
    peer_push_mode = mode == "ce" ? PeerPush::ce : (mode == "sm" ? PeerPush::sm : PeerPush::direct);
  }
  unverified_sparse_multilevel = pin->GetOrAddBoolean("pb2", "unverified_sparse_multilevel", false);
  sparse_config.enabled = pin->GetOrAddBoolean("parthenon/sparse", "enable_sparse", true);
  sparse_config.allocation_threshold = pin->GetOrAddReal("parthenon/sparse", "alloc_threshold", 1e-12);
  sparse_config.deallocation_threshold =
      pin->GetOrAddReal("parthenon/sparse", "dealloc_threshold", 1e-14);
  sparse_config.deallocation_count = pin->GetOrAddInteger("parthenon/sparse", "dealloc_count", 5);
}

void Mesh::FinishConstruction(ParameterInput *pin) {
  const std::string refinement = pin->GetOrAddString("parthenon/mesh", "refinement", "none");
  if (refinement != "none") multilevel = true; // coarse buffers exist (mesh.cpp:118-140)

  std::vector<double> cost(nbtotal, 1.0);
  AssignBlocks(cost, nranks, ranklist);
  nblist.assign(nranks, 0);
  for (int r : ranklist) nblist[r]++;
  nslist.assign(nranks, 0);
  for (int r = 1; r < nranks; ++r) nslist[r] = nslist[r - 1] + nblist[r - 1];

  max_level = pin->GetOrAddInteger("parthenon/mesh", "numlevel", 1) + root_level - 1;
  derefine_count = pin->GetOrAddInteger("parthenon/mesh", "derefine_count", 10);
  // Refinement::Initialize (refinement_package.cpp:36-50)
  for (int n = 0;; ++n) {
    const std::string block = "parthenon/refinement" + std::to_string(n);
    if (!pin->DoesBlockExist(block)) break;
    AMRCriterion c;
    const std::string method = pin->GetOrAddString(block, "method", "PLEASE SPECIFY method");
    PARTHENON_REQUIRE_THROWS(method == "derivative_order_1" || method == "derivative_order_2",
                             "\n  Invalid selection for refinment method in " + block + ": " + method);
    c.order = method == "derivative_order_1" ? 1 : 2;
    c.field = pin->GetOrAddString(block, "field", "NO FIELD WAS SET");
    PARTHENON_REQUIRE_THROWS(c.field != "NO FIELD WAS SET", "Error in " + block + ": no field set");
    PARTHENON_REQUIRE_THROWS(!pin->DoesParameterExist(block, "tensor_ij") &&
                                 !pin->DoesParameterExist(block, "tensor_ijk"),
                             "tensor-valued refinement fields are not supported");
    c.comp = pin->GetOrAddInteger(block, "vector_i", 0);
    c.refine_criteria = pin->GetOrAddReal(block, "refine_tol", 0.5);
    c.derefine_criteria = pin->GetOrAddReal(block, "derefine_tol", 0.05);
    const int global_max_level = pin->GetOrAddInteger("parthenon/mesh", "numlevel", 1);
    c.max_level = std::min(pin->GetOrAddInteger(block, "max_level", global_max_level), global_max_level);
    c.max_level += root_level;
    amr_criteria.push_back(c);
  }
  BuildBlockList(nullptr);
  for (auto &name : packages.Order())
    for (auto &f : packages.Get(name)->AllFields()) {
      resolved_fields.push_back(f);
      // metadata.cpp:160-171: sparse fields take the global thresholds, dense ones 0
      FieldEntry &e = resolved_fields.back();
      if (e.m.IsSparse())
        e.m.SetSparseThresholds(sparse_config.allocation_threshold,
                                sparse_config.deallocation_threshold, e.m.GetDefaultValue());
    }
}

void Mesh::AllocateSparse(const std::string &label, int lid) {
  MeshBlock *pmb = block_list[lid].get();
  for (auto &kv : mesh_data.All()) {
    MeshData<Real> *md = kv.second.get();
    if (md->partition_id() != pmb->partition || !md->HasVariable(label)) continue;
    Variable &v = md->Get(label);
    if (v.IsAllocated(pmb->pack_index)) continue; // OneCopy fields are shared between stages
    v.AllocateBlock(pmb->pack_index);
    md->alloc_generation++;
  }
}

void Mesh::DeallocateSparse(const std::string &label, int lid) {
  MeshBlock *pmb = block_list[lid].get();
  for (auto &kv : mesh_data.All()) {
    MeshData<Real> *md = kv.second.get();
    if (md->partition_id() != pmb->partition || !md->HasVariable(label)) continue;
    Variable &v = md->Get(label);
    if (!v.IsAllocated(pmb->pack_index)) continue;
    v.SetAllocated(pmb->pack_index, false);
    md->alloc_generation++;
  }
}

void Mesh::BuildBlockList(const BlockList_t *keep) {
  const int ng = Globals::nghost;
  const int rank = my_rank;
  std::unordered_map<LogicalLocation, std::shared_ptr<MeshBlock>, LogicalLocationHash> old;
  if (keep)
    for (auto &pmb : *keep) old[pmb->loc] = pmb;
  // every block's geometry and neighbour search is independent of the others (the tree is only
  // read): the rebuild after a remesh of thousands of blocks runs on all host threads
  BlockList_t blocks(static_cast<size_t>(nblist[rank]));
  const int gid0 = nslist[rank], gid1 = nslist[rank] + nblist[rank];
#pragma omp parallel for schedule(static) if (gid1 - gid0 > 256)
  for (int gid = gid0; gid < gid1; ++gid) {
    auto it = old.find(loclist[gid]);
    auto mb = it != old.end() ? it->second : std::make_shared<MeshBlock>();
    mb->gid = gid;
    mb->lid = gid - nslist[rank];
    mb->loc = loclist[gid];
    mb->block_size = GetBlockSize(mb->loc);
    if (forest) forest->GetBlockDomain(mb->loc, mb->block_size.xmin_.data(), mb->block_size.xmax_.data());
    const int nx1 = base_block_size.nx_[0], nx2 = ndim > 1 ? base_block_size.nx_[1] : 0,
              nx3 = ndim > 2 ? base_block_size.nx_[2] : 0;
    mb->cellbounds = IndexShape(nx3, nx2, nx1, ng);
    // meshblock.cpp:204-216
    mb->c_cellbounds = IndexShape(ndim > 2 ? std::max(1, nx3 / 2) : 0,
                                  ndim > 1 ? std::max(1, nx2 / 2) : 0, std::max(1, nx1 / 2), ng);
    mb->coords = UniformCartesian(mb->block_size, ng);
    for (int f = 0; f < 6; ++f) mb->boundary_flag[f] = BoundaryFlag::block;
    for (int d = 0; d < ndim && !forest; ++d) {
      if (mb->loc.lx[d] == 0) mb->boundary_flag[2 * d] = mesh_bcs[2 * d];
      if (mb->loc.lx[d] == BlocksAtLevel(mb->loc.level, d) - 1)
        mb->boundary_flag[2 * d + 1] = mesh_bcs[2 * d + 1];
    }
    if (forest) {
      // the tree's flags where the block touches the tree boundary; an edge flagged `user` takes
      // the deck's condition of that face (outflow unless the deck names another)
      const auto bcs = forest->GetBlockBCs(mb->loc);
      for (int f = 0; f < 4; ++f)
        mb->boundary_flag[f] = bcs[f] == BoundaryFlag::user ? mesh_bcs[f] : bcs[f];
    }
    mb->pmy_mesh = this;
    const int ps = DefaultPackSizeFor(static_cast<int>(nblist[rank]));
    mb->partition = mb->lid / ps;
    mb->pack_index = mb->lid % ps;
    FindNeighbors(*mb);
    blocks[gid - gid0] = mb;
  }
  block_list = std::move(blocks);
  fine_coarse_faces_ = -1;
  plan_cache.clear();
}

Mesh::~Mesh() = default;

int Mesh::DefaultPackSizeFor(int nblocks) const {
  return pack_size_ < 1 ? std::max(nblocks, 1) : pack_size_;
}
int Mesh::DefaultPackSize() const { return DefaultPackSizeFor(GetNumMeshBlocksThisRank()); }
int Mesh::DefaultNumPartitions() const {
  const int ps = DefaultPackSize();
  return (GetNumMeshBlocksThisRank() + ps - 1) / ps;
}

Real *Mesh::ScratchReal() {
  if (!scratch_) scratch_.Allocate(64 * sizeof(Real), stream);
  return scratch_.get<Real>();
}

void Mesh::ReduceHistory(std::vector<Real> &vals) {
  if (nranks == 1 || vals.empty()) return;
  PARTHENON_REQUIRE(vals.size() <= 64, "too many history columns");
  PARTHENON_REQUIRE(comm != nullptr, "multi-rank mesh without a communicator");
  Real *d = ScratchReal();
  PB2_CHECK(pb2_memcpy_h2d(d, vals.data(), sizeof(Real) * vals.size(), stream));
  PB2_CHECK(pb2_comm_allreduce_sum(comm, d, static_cast<int64_t>(vals.size()), stream));
  PB2_CHECK(pb2_memcpy_d2h(vals.data(), d, sizeof(Real) * vals.size(), stream));
  PB2_CHECK(pb2_stream_sync(stream));
}

void Mesh::AllReduceSum(std::vector<Real> &vals) {
  if (nranks == 1 || vals.empty()) return;
  PARTHENON_REQUIRE(comm != nullptr, "multi-rank mesh without a communicator");
  DeviceBuffer d;
  d.Allocate(sizeof(Real) * vals.size(), stream);
  PB2_CHECK(pb2_memcpy_h2d(d.get(), vals.data(), sizeof(Real) * vals.size(), stream));
  PB2_CHECK(pb2_comm_allreduce_sum(comm, d.get<Real>(), static_cast<int64_t>(vals.size()), stream));
  PB2_CHECK(pb2_memcpy_d2h(vals.data(), d.get(), sizeof(Real) * vals.size(), stream));
  PB2_CHECK(pb2_stream_sync(stream));
}

int Mesh::SlabCapacity(int nblocks) {
  if (!adaptive) return nblocks;
  // grown by half when exceeded (shrunk when less than a third is used): every change of the
  // capacity re-allocates all field slabs, and cudaMalloc of GB-sized slabs was measured at
  // 30-270 ms per event on a B200 (profiles/README.md, session r03e) — rarer events matter more
  // than the spare blocks
  if (nblocks > slab_capacity_ || 3 * nblocks < slab_capacity_)
    slab_capacity_ = (nblocks + nblocks / 2 + 63) / 64 * 64;
  return slab_capacity_;
}

bool Mesh::HasFineCoarseFaces() const {
  if (!multilevel) return false;
  if (fine_coarse_faces_ < 0) {
    fine_coarse_faces_ = 0;
    for (auto &pmb : block_list)
      for (auto &nb : pmb->neighbors)
        if (nb.loc.level != pmb->loc.level) fine_coarse_faces_ = 1;
  }
  return fine_coarse_faces_ == 1;
}

int Mesh::VirtualRankOf(int gid) const {
  if (virtual_ranks <= 1) return 0;
  // contiguous Morton ranges inside this rank, like the real partition
  const int r = ranklist[gid];
  const int64_t l = gid - nslist[r];
  return static_cast<int>(l * virtual_ranks / std::max(nblist[r], 1));
}

} // namespace parthenon
