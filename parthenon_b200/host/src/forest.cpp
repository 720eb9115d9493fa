// forest.cpp — see pb2/forest.hpp for what this replaces in the reference.
#include "pb2/forest.hpp"

#include <algorithm>
#include <cstdlib>

namespace parthenon {
namespace forest {

namespace {
// Face::node_to_offset, forest_topology.hpp:122-124 (the x3 entry is -1 for every node and only
// serves the reference's IsEdge() test; two active directions are kept here)
constexpr int kNodeOffset[4][2] = {{-1, -1}, {1, -1}, {-1, 1}, {1, 1}};
inline int Idx9(int ox, int oy) { return (ox + 1) + 3 * (oy + 1); }
inline int Idx27(int o1, int o2, int o3) { return (o1 + 1) + 3 * (o2 + 1) + 9 * (o3 + 1); }

// logical_location.cpp:48-56
int NeighborTreeIndex(const LogicalLocation &loc) {
  const int64_t up = int64_t{1} << std::max(loc.level, 0);
  int i[3];
  for (int d = 0; d < 3; ++d) i[d] = (loc.lx[d] >= 0) - (loc.lx[d] < up) + 1;
  return i[0] + 3 * i[1] + 9 * i[2];
}
// logical_location.cpp:160-181 (locations in the halo of a tree and negative levels included)
LogicalLocation ParentOf(const LogicalLocation &loc) {
  const int64_t norig = int64_t{1} << std::max(loc.level, 0);
  const int64_t nparent = int64_t{1} << std::max(loc.level - 1, 0);
  constexpr int64_t nmax = 5;
  LogicalLocation p = loc;
  p.level = loc.level - 1;
  for (int d = 0; d < 3; ++d) {
    const int64_t off_l = loc.lx[d] + nmax * norig;
    p.lx[d] = ((off_l % norig) >> 1) + (off_l / norig - nmax) * nparent;
  }
  return p;
}
// logical_location.cpp:183-200 with ndim = 2 (first direction outermost)
std::array<LogicalLocation, 4> DaughtersOf(const LogicalLocation &loc) {
  std::array<LogicalLocation, 4> out;
  int n = 0;
  for (int i = 0; i < 2; ++i)
    for (int j = 0; j < 2; ++j)
      out[n++] = LogicalLocation(loc.tree, loc.level + 1, (loc.lx[0] << 1) + i, (loc.lx[1] << 1) + j, 0);
  return out;
}
std::array<int64_t, 3> Key(const LogicalLocation &l) { return {l.level, l.lx[0], l.lx[1]}; }

// logical_location.cpp:61-74, defs.hpp:98-101
Real SymmetrizedCoordinate(int64_t index, int bloc, int64_t nrange) {
  const int64_t noffset = index - nrange / 2, noffset_ceil = index - (nrange + 1) / 2;
  return static_cast<Real>(noffset + noffset_ceil + static_cast<int64_t>(bloc)) /
         (2.0 * static_cast<Real>(nrange));
}
Real LogicalToActual(Real u, Real xmin, Real xmax) {
  return static_cast<Real>(0.5) * (xmin + xmax) + (u * xmax - u * xmin);
}

struct FaceTopo {
  std::array<uint64_t, 4> nodes;
  std::array<std::vector<std::pair<int, LogicalCoordinateTransformation>>, 9> nbr;
  std::map<int, std::array<int, 2>> offset_of; // neighbour face -> offset it sits at
  int IndexOf(uint64_t node) const {
    for (int i = 0; i < 4; ++i)
      if (nodes[i] == node) return i;
    return -1;
  }
  std::vector<uint64_t> Overlap(const FaceTopo &o) const { // in this face's node order
    std::vector<uint64_t> out;
    for (uint64_t n : nodes)
      if (o.IndexOf(n) >= 0) out.push_back(n);
    return out;
  }
  // Face::GetEdgeDirections (forest_topology.cpp:82-103): the signed direction along the edge
  // nodes[0] -> nodes[1], the direction normal to it and the side of the face it lies on
  void EdgeDirections(uint64_t n0, uint64_t n1, int &dir_tang, int &dir_norm, int &side) const {
    const int i0 = IndexOf(n0), i1 = IndexOf(n1);
    const int diff = std::abs(i1 - i0);
    PARTHENON_REQUIRE(diff == 1 || diff == 2, "the two nodes are not an edge of the face");
    dir_tang = (i1 - i0 > 0 ? 1 : -1) * diff; // IntegerLog2Floor(diff) + 1 for diff = 1, 2
    const int o0 = (kNodeOffset[i0][0] + kNodeOffset[i1][0]) / 2;
    const int o1 = (kNodeOffset[i0][1] + kNodeOffset[i1][1]) / 2;
    if (std::abs(o0) == 1) {
      dir_norm = 1;
      side = o0;
    } else {
      PARTHENON_REQUIRE(std::abs(o1) == 1, "the two nodes are not an edge of the face");
      dir_norm = 2;
      side = o1;
    }
  }
};
} // namespace

// ---------------------------------------------------------------------------------------
// logical_coordinate_transformation.cpp:38-118
// ---------------------------------------------------------------------------------------
LogicalLocation LogicalCoordinateTransformation::Transform(const LogicalLocation &loc,
                                                           int64_t destination) const {
  const int64_t nblock = int64_t{1} << std::max(loc.level, 0);
  LogicalLocation out = loc;
  out.tree = destination;
  for (int d = 0; d < 3; ++d) {
    int64_t l_in = loc.lx[d];
    // back into the interior of the bordering tree as if it had our orientation ...
    if (use_offset)
      l_in -= offset[d] * nblock;
    else
      l_in = (l_in + nblock) % nblock;
    // ... then permute / flip into its own coordinates
    out.lx[std::abs(dir_connection[d])] = dir_flip[d] ? nblock - 1 - l_in : l_in;
  }
  return out;
}

LogicalLocation LogicalCoordinateTransformation::InverseTransform(const LogicalLocation &loc,
                                                                  int64_t origin) const {
  const int64_t nblock = int64_t{1} << std::max(loc.level, 0);
  LogicalLocation out = loc;
  out.tree = origin;
  for (int d = 0; d < 3; ++d) {
    const int64_t l_in = loc.lx[std::abs(dir_connection[d])];
    out.lx[d] = dir_flip[d] ? nblock - 1 - l_in : l_in;
    if (use_offset)
      out.lx[d] += offset[d] * nblock;
    else
      out.lx[d] = (out.lx[d] + nblock) % nblock;
  }
  return out;
}

std::array<int, 3> LogicalCoordinateTransformation::Transform(const std::array<int, 3> &in) const {
  std::array<int, 3> out{0, 0, 0};
  for (int d = 0; d < 3; ++d) out[std::abs(dir_connection[d])] = dir_flip[d] ? -in[d] : in[d];
  return out;
}

LogicalCoordinateTransformation ComposeTransformations(const LogicalCoordinateTransformation &first,
                                                       const LogicalCoordinateTransformation &second) {
  LogicalCoordinateTransformation out;
  for (int d = 0; d < 3; ++d) {
    out.dir_connection[d] = second.dir_connection[first.dir_connection[d]];
    out.dir_flip[d] = second.dir_flip[first.dir_connection[d]] != first.dir_flip[d];
    out.offset[d] = second.offset[first.dir_connection[d]] * (first.dir_flip[d] ? -1 : 1) +
                    first.offset[d];
  }
  for (int d = 0; d < 3; ++d) out.dir_connection_inverse[out.dir_connection[d]] = d;
  out.use_offset = first.use_offset && second.use_offset;
  return out;
}

// ---------------------------------------------------------------------------------------
// Forest::Make2D (forest.cpp:204-297) with the face topology of forest_topology.cpp:57-174
// ---------------------------------------------------------------------------------------
Forest::Forest(const ForestDefinition &def) {
  const int nf = static_cast<int>(def.faces.size());
  PARTHENON_REQUIRE(nf > 0, "a forest needs at least one face");
  std::vector<FaceTopo> faces(nf);
  std::map<uint64_t, std::vector<int>> faces_of_node;
  for (int f = 0; f < nf; ++f) {
    faces[f].nodes = def.faces[f].nodes;
    for (uint64_t n : faces[f].nodes) faces_of_node[n].push_back(f);
  }
  // neighbours: faces that share one node (corner) or two (edge); the offset is the mean of the
  // shared nodes' positions in this face (SetNeighbors :57-80)
  for (int f = 0; f < nf; ++f)
    for (int g = 0; g < nf; ++g) {
      if (g == f) continue;
      const auto ov = faces[f].Overlap(faces[g]);
      if (ov.empty()) continue;
      PARTHENON_REQUIRE(ov.size() <= 2, "two faces of a forest share more than an edge");
      int off[2] = {0, 0};
      for (uint64_t n : ov)
        for (int o = 0; o < 2; ++o) off[o] += kNodeOffset[faces[f].IndexOf(n)][o];
      for (int o = 0; o < 2; ++o) off[o] /= static_cast<int>(ov.size());
      faces[f].nbr[Idx9(off[0], off[1])].emplace_back(g, LogicalCoordinateTransformation());
      faces[f].offset_of[g] = {off[0], off[1]};
    }
  // across an edge (SetEdgeCoordinateTransforms :105-138): the direction along the edge maps to
  // the neighbour's direction along it (reversed if the neighbour runs it the other way), the
  // normal maps to the neighbour's normal (flipped if both faces see the edge on the same side)
  for (int f = 0; f < nf; ++f)
    for (int ox = -1; ox <= 1; ++ox)
      for (int oy = -1; oy <= 1; ++oy) {
        if (std::abs(ox) + std::abs(oy) != 1) continue;
        for (auto &[g, ct] : faces[f].nbr[Idx9(ox, oy)]) {
          const auto ov = faces[f].Overlap(faces[g]); // sorted by this face's node index
          PARTHENON_REQUIRE(ov.size() == 2, "an edge neighbour must share two nodes");
          int t, n, side, tn, nn, siden;
          faces[f].EdgeDirections(ov[0], ov[1], t, n, side);
          faces[g].EdgeDirections(ov[0], ov[1], tn, nn, siden);
          LogicalCoordinateTransformation c;
          c.SetDirection(std::abs(t), std::abs(tn), tn < 0);
          c.SetDirection(n, nn, side == siden);
          c.offset = {ox, oy, 0};
          c.use_offset = true;
          ct = c;
        }
      }
  // across a corner (SetNodeCoordinateTransforms :140-174): through a face that is an edge
  // neighbour of both, composing the two edge transformations
  for (int f = 0; f < nf; ++f)
    for (int ox = -1; ox <= 1; ox += 2)
      for (int oy = -1; oy <= 1; oy += 2)
        for (auto &[g, ct] : faces[f].nbr[Idx9(ox, oy)]) {
          const auto ov = faces[f].Overlap(faces[g]);
          PARTHENON_REQUIRE(ov.size() == 1, "a corner neighbour must share one node");
          bool found = false;
          for (int h : faces_of_node[ov[0]]) {
            if (!faces[f].offset_of.count(h) || !faces[g].offset_of.count(h)) continue;
            const auto o1 = faces[f].offset_of[h], og = faces[g].offset_of[h];
            if (std::abs(o1[0]) + std::abs(o1[1]) != 1 || std::abs(og[0]) + std::abs(og[1]) != 1)
              continue;
            const auto &ct1 = faces[f].nbr[Idx9(o1[0], o1[1])][0].second;
            const auto o2 = faces[h].offset_of[g];
            const auto &ct2 = faces[h].nbr[Idx9(o2[0], o2[1])][0].second;
            const auto composed = ComposeTransformations(ct1, ct2);
            // the reference takes whichever such face its hash set yields first; the result is
            // only defined if every path around the node gives the same transformation
            PARTHENON_REQUIRE(!found || composed == ct,
                              "corner transformation depends on the path around the node");
            ct = composed;
            found = true;
          }
          PARTHENON_REQUIRE(found, "no common edge neighbour for a corner neighbour");
        }

  // trees: x3 periodic, the listed edges carry their flag, every face towards another tree is
  // `block` (forest.cpp:213-251, Tree::AddNeighborTree tree.cpp:334-344)
  trees_.resize(nf);
  for (int f = 0; f < nf; ++f) {
    Tree &t = trees_[f];
    t.id = def.faces[f].id;
    PARTHENON_REQUIRE(!index_of_.count(t.id), "two faces of a forest have the same id");
    index_of_[t.id] = f;
    t.nodes = def.faces[f].nodes;
    t.xmin = def.faces[f].xmin;
    t.xmax = def.faces[f].xmax;
    t.bcs = {BoundaryFlag::block, BoundaryFlag::block, BoundaryFlag::block,
             BoundaryFlag::block, BoundaryFlag::periodic, BoundaryFlag::periodic};
    t.leaves.insert({0, 0, 0});
    LogicalCoordinateTransformation self; // the tree is its own central neighbour
    self.use_offset = true;
    t.neighbors[13].push_back(TreeNeighbor{f, self});
  }
  for (const auto &bc : def.bc_edges)
    for (int f = 0; f < nf; ++f) {
      const int i0 = faces[f].IndexOf(bc.nodes[0]), i1 = faces[f].IndexOf(bc.nodes[1]);
      if (i0 < 0 || i1 < 0) continue;
      const int o0 = (kNodeOffset[i0][0] + kNodeOffset[i1][0]) / 2;
      const int o1 = (kNodeOffset[i0][1] + kNodeOffset[i1][1]) / 2;
      if (std::abs(o0) + std::abs(o1) != 1) continue; // a diagonal is not an edge (Face::IsEdge)
      if (o0 == -1)
        trees_[f].bcs[0] = bc.flag;
      else if (o0 == 1)
        trees_[f].bcs[1] = bc.flag;
      else if (o1 == -1)
        trees_[f].bcs[2] = bc.flag;
      else
        trees_[f].bcs[3] = bc.flag;
    }
  for (int f = 0; f < nf; ++f)
    for (int ox = -1; ox <= 1; ++ox)
      for (int oy = -1; oy <= 1; ++oy)
        for (auto &[g, ct] : faces[f].nbr[Idx9(ox, oy)]) {
          auto &slot = trees_[f].neighbors[Idx27(ox, oy, 0)];
          bool have = false;
          for (auto &tn : slot) have = have || tn.tree == g;
          if (!have) slot.push_back(TreeNeighbor{g, ct});
          if (std::abs(ox) + std::abs(oy) == 1)
            trees_[f].bcs[ox != 0 ? (ox > 0 ? 1 : 0) : (oy > 0 ? 3 : 2)] = BoundaryFlag::block;
        }
  for (const LogicalLocation &loc : def.refinement_locations) {
    PARTHENON_REQUIRE(index_of_.count(loc.tree), "initial refinement names an unknown tree");
    AddMeshBlock(index_of_.at(loc.tree), loc);
  }
}

const Forest::Tree &Forest::TreeOf(const LogicalLocation &loc) const {
  auto it = index_of_.find(loc.tree);
  PARTHENON_REQUIRE(it != index_of_.end(), "location on an unknown tree");
  return trees_[it->second];
}

bool Forest::IsLeaf(const LogicalLocation &loc) const {
  auto it = index_of_.find(loc.tree);
  return it != index_of_.end() && trees_[it->second].leaves.count(Key(loc)) > 0;
}

int Forest::AddMeshBlock(int t, const LogicalLocation &loc) {
  Tree &tr = trees_[t];
  if (tr.internal.count(Key(loc))) return -1;
  if (tr.leaves.count(Key(loc))) return 0;
  std::vector<LogicalLocation> todo; // ancestors up to the first one that is a leaf
  LogicalLocation parent = ParentOf(loc);
  for (int l = loc.level - 1; l >= 0; --l) {
    todo.push_back(parent);
    if (tr.leaves.count(Key(parent))) break;
    parent = ParentOf(parent);
  }
  int added = 0;
  for (auto it = todo.rbegin(); it != todo.rend(); ++it) added += Refine(t, *it);
  return added;
}

int Forest::Refine(int t, const LogicalLocation &ref) {
  Tree &tr = trees_[t];
  if (!tr.leaves.count(Key(ref))) return 0; // (negative levels are never leaves)
  tr.leaves.erase(Key(ref));
  tr.internal.insert(Key(ref));
  for (const LogicalLocation &d : DaughtersOf(ref)) tr.leaves.insert(Key(d));
  int nadded = 3;
  // proper nesting: the neighbours of the parent on this block's side must exist, also in the
  // bordering trees (tree.cpp:113-134)
  const LogicalLocation parent = ParentOf(ref);
  const int64_t ox1 = ref.lx[0] - (parent.lx[0] << 1), ox2 = ref.lx[1] - (parent.lx[1] << 1);
  for (int j = 0; j < 2; ++j)
    for (int i = 0; i < 2; ++i) {
      LogicalLocation neigh = parent;
      neigh.lx[0] += i + ox1 - 1;
      neigh.lx[1] += j + ox2 - 1;
      // (copy: refining a neighbour tree never edits this tree's neighbour table)
      const auto slot = trees_[t].neighbors[NeighborTreeIndex(neigh)];
      for (const TreeNeighbor &tn : slot)
        nadded += Refine(tn.tree, tn.ct.Transform(neigh, trees_[tn.tree].id));
    }
  return nadded;
}

std::vector<LogicalLocation> Forest::GetMeshBlockList() const {
  int maxlevel = 0;
  for (const Tree &t : trees_)
    for (const auto &k : t.leaves) maxlevel = std::max<int>(maxlevel, static_cast<int>(k[0]));
  std::vector<LogicalLocation> out;
  for (const auto &[id, f] : index_of_) { // trees by id
    std::vector<LogicalLocation> mine;
    for (const auto &k : trees_[f].leaves)
      mine.emplace_back(id, static_cast<int>(k[0]), k[1], k[2], 0);
    std::sort(mine.begin(), mine.end(), [&](const LogicalLocation &a, const LogicalLocation &b) {
      const uint64_t ka = a.MortonKey(maxlevel), kb = b.MortonKey(maxlevel);
      return ka != kb ? ka < kb : a.level < b.level;
    });
    out.insert(out.end(), mine.begin(), mine.end());
  }
  return out;
}

std::vector<NeighborLocation> Forest::FindNeighbors(const LogicalLocation &loc) const {
  const Tree &me = TreeOf(loc);
  std::vector<NeighborLocation> out;
  for (int o1 = -1; o1 <= 1; ++o1)
    for (int o2 = -1; o2 <= 1; ++o2) {
      if (o1 == 0 && o2 == 0) continue;
      LogicalLocation neigh = loc;
      neigh.lx[0] += o1;
      neigh.lx[1] += o2;
      for (const TreeNeighbor &tn : me.neighbors[NeighborTreeIndex(neigh)]) {
        const Tree &nt = trees_[tn.tree];
        const LogicalLocation tneigh = tn.ct.Transform(neigh, nt.id);
        const LogicalLocation tloc = tn.ct.Transform(loc, nt.id);
        PARTHENON_REQUIRE(tn.ct.InverseTransform(tloc, me.id) == loc, "inverse transform not working");
        if (nt.leaves.count(Key(tneigh))) {
          out.push_back({tneigh, tn.ct.InverseTransform(tneigh, me.id), tn.ct});
        } else if (nt.internal.count(Key(tneigh))) {
          for (const LogicalLocation &d : DaughtersOf(tneigh))
            if (tloc.IsNeighbor(d)) out.push_back({d, tn.ct.InverseTransform(d, me.id), tn.ct});
        } else if (tneigh.level > 0 && nt.leaves.count(Key(ParentOf(tneigh)))) {
          // a coarser neighbour covers several offsets: listed at the one it sits at only
          const LogicalLocation tpar = ParentOf(tneigh);
          const LogicalLocation neighp = tn.ct.InverseTransform(tpar, me.id);
          const auto so = loc.GetSameLevelOffsets(neighp);
          if (so[0] == o1 && so[1] == o2 && so[2] == 0) out.push_back({tpar, neighp, tn.ct});
        }
      }
    }
  return out;
}

std::array<BoundaryFlag, 6> Forest::GetBlockBCs(const LogicalLocation &loc) const {
  std::array<BoundaryFlag, 6> out = TreeOf(loc).bcs;
  const int64_t nblock = int64_t{1} << std::max(loc.level, 0);
  for (int d = 0; d < 3; ++d) {
    if (loc.lx[d] != 0) out[2 * d] = BoundaryFlag::block;
    if (loc.lx[d] != nblock - 1) out[2 * d + 1] = BoundaryFlag::block;
  }
  return out;
}

void Forest::GetBlockDomain(const LogicalLocation &loc, Real xmin[3], Real xmax[3]) const {
  const Tree &t = TreeOf(loc);
  const int64_t nblock = int64_t{1} << std::max(loc.level, 0);
  for (int d = 0; d < 3; ++d) {
    if (d < 2) {
      xmin[d] = LogicalToActual(SymmetrizedCoordinate(loc.lx[d], 0, nblock), t.xmin[d], t.xmax[d]);
      xmax[d] = LogicalToActual(SymmetrizedCoordinate(loc.lx[d], 2, nblock), t.xmin[d], t.xmax[d]);
    } else {
      xmin[d] = t.xmin[d];
      xmax[d] = t.xmax[d];
    }
  }
}

} // namespace forest
} // namespace parthenon
