// forest.hpp — forests of differently oriented trees (host only, 2-D like the reference's).
//
// What the ghost-zone path needs from the reference's src/mesh/forest: a mesh described as
// quadrilateral faces over shared nodes (ForestDefinition, forest.hpp:47-78), each face one tree;
// the trees a tree meets across its edges and corners and the LogicalCoordinateTransformation
// that takes its logical coordinates into theirs (forest_topology.cpp:57-174,
// logical_coordinate_transformation.{hpp,cpp}); refinement that keeps 2:1 nesting across tree
// boundaries (tree.cpp:73-137); the leaf list in gid order (forest.cpp:39-63) and the neighbour
// search (tree.cpp:139-226).  example/boundary_exchange is the application that uses it.
//
// Layout here: faces are indexed 0..n-1 in the order they were added (their ids are kept for
// LogicalLocation::tree), nodes are plain integers, and every face carries a table
// [9 offsets] -> list of (neighbour face, transformation) instead of pointer graphs.
#pragma once
#include <array>
#include <map>
#include <memory>
#include <set>
#include <vector>

#include "types.hpp"

namespace parthenon {

struct LogicalLocation {
  int level = 0;
  int64_t lx[3] = {0, 0, 0};
  int64_t tree = 0; // id of the tree (forest meshes); 0 on hyper-rectangular meshes
  LogicalLocation() = default;
  LogicalLocation(int lev, int64_t l1, int64_t l2, int64_t l3) : level(lev), lx{l1, l2, l3} {}
  LogicalLocation(int64_t t, int lev, int64_t l1, int64_t l2, int64_t l3)
      : level(lev), lx{l1, l2, l3}, tree(t) {}
  int64_t lx1() const { return lx[0]; }
  int64_t lx2() const { return lx[1]; }
  int64_t lx3() const { return lx[2]; }
  bool operator==(const LogicalLocation &o) const {
    return level == o.level && lx[0] == o.lx[0] && lx[1] == o.lx[1] && lx[2] == o.lx[2] &&
           tree == o.tree;
  }
  bool operator!=(const LogicalLocation &o) const { return !(*this == o); }
  LogicalLocation GetParent() const {
    LogicalLocation p = *this;
    p.level = level - 1;
    for (int d = 0; d < 3; ++d) p.lx[d] = lx[d] >> 1;
    return p;
  }
  // z-order key at `maxlevel` resolution, x in the lowest interleaved bit
  // (utils/morton_number.hpp:43)
  uint64_t MortonKey(int maxlevel) const;
  // logical_location.cpp:98-108
  std::array<int, 3> GetSameLevelOffsets(const LogicalLocation &neighbor) const;
  // logical_location.cpp:110-129
  bool IsNeighbor(const LogicalLocation &in) const;
  // logical_location.cpp:131-158: does block `in` touch the topological element of this block
  // at te_offset (a face, edge or node of the block; {0,0,0} is its volume)
  bool IsNeighborOfTE(const LogicalLocation &in, const std::array<int, 3> &te_offset) const;
};

struct LogicalLocationHash {
  size_t operator()(const LogicalLocation &l) const {
    uint64_t h = (static_cast<uint64_t>(l.level) + 31 * static_cast<uint64_t>(l.tree)) *
                 0x9E3779B97F4A7C15ull;
    for (int d = 0; d < 3; ++d)
      h ^= (static_cast<uint64_t>(l.lx[d]) + 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2));
    return static_cast<size_t>(h);
  }
};

namespace forest {

// logical_coordinate_transformation.hpp:36-112: how the logical coordinates of one tree read in
// a neighbouring tree — a permutation of the axes, flips, and the offset of the neighbour
struct LogicalCoordinateTransformation {
  std::array<int, 3> dir_connection{0, 1, 2}, dir_connection_inverse{0, 1, 2};
  std::array<bool, 3> dir_flip{false, false, false};
  std::array<int, 3> offset{0, 0, 0};
  bool use_offset = false;
  void SetDirection(int origin, int neighbor, bool reversed = false) { // 1-based directions
    dir_connection[origin - 1] = neighbor - 1;
    dir_connection_inverse[neighbor - 1] = origin - 1;
    dir_flip[origin - 1] = reversed;
  }
  LogicalLocation Transform(const LogicalLocation &loc, int64_t destination) const;
  LogicalLocation InverseTransform(const LogicalLocation &loc, int64_t origin) const;
  std::array<int, 3> Transform(const std::array<int, 3> &offsets) const; // CellCentOffsets
  bool IsIdentity() const {
    return dir_connection == std::array<int, 3>{0, 1, 2} && !dir_flip[0] && !dir_flip[1] &&
           !dir_flip[2];
  }
  bool operator==(const LogicalCoordinateTransformation &o) const {
    return dir_connection == o.dir_connection && dir_flip == o.dir_flip && offset == o.offset &&
           use_offset == o.use_offset;
  }
};
LogicalCoordinateTransformation ComposeTransformations(const LogicalCoordinateTransformation &first,
                                                       const LogicalCoordinateTransformation &second);

// forest_node.hpp: a corner shared by faces; the position is for plots only
struct Node {
  uint64_t id;
  std::array<Real, 2> x;
  static std::shared_ptr<Node> create(uint64_t id, std::array<Real, 2> pos) {
    return std::make_shared<Node>(Node{id, pos});
  }
};

struct Edge { // forest_topology.hpp:44-66
  std::array<std::shared_ptr<Node>, 2> nodes;
  Edge() = default;
  explicit Edge(std::array<std::shared_ptr<Node>, 2> n) : nodes(std::move(n)) {}
};

// forest.hpp:47-78.  Node order of a face:   2---3
//                                            |   |     X1 from 0 to 1, X2 from 0 to 2
//                                            0---1
class ForestDefinition {
 public:
  using ar3_t = std::array<Real, 3>;
  void AddFace(std::size_t id, std::array<std::shared_ptr<Node>, 4> nodes,
               ar3_t xmin = {0.0, 0.0, 0.0}, ar3_t xmax = {1.0, 1.0, 1.0}) {
    faces.push_back(FaceDef{static_cast<int64_t>(id), {nodes[0]->id, nodes[1]->id, nodes[2]->id, nodes[3]->id}, xmin, xmax});
  }
  void AddBC(Edge edge, BoundaryFlag bf = BoundaryFlag::user) {
    PARTHENON_REQUIRE(bf != BoundaryFlag::periodic,
                      "periodic connections between forest edges are not supported");
    bc_edges.push_back(BcEdge{{edge.nodes[0]->id, edge.nodes[1]->id}, bf});
  }
  void AddInitialRefinement(const LogicalLocation &loc) { refinement_locations.push_back(loc); }

  struct FaceDef {
    int64_t id;
    std::array<uint64_t, 4> nodes;
    ar3_t xmin, xmax;
  };
  struct BcEdge {
    std::array<uint64_t, 2> nodes;
    BoundaryFlag flag;
  };
  std::vector<FaceDef> faces;
  std::vector<BcEdge> bc_edges;
  std::vector<LogicalLocation> refinement_locations;
};

// tree.cpp:139-226 NeighborLocation
struct NeighborLocation {
  LogicalLocation global_loc; // the neighbour as stored in its own tree
  LogicalLocation origin_loc; // ... in the index space of the block that asked
  LogicalCoordinateTransformation lcoord_trans;
};

class Forest {
 public:
  // Forest::Make2D, forest.cpp:204-297
  explicit Forest(const ForestDefinition &def);
  // forest.cpp:39-63: leaves of every tree in (tree id, Morton number, level) order
  std::vector<LogicalLocation> GetMeshBlockList() const;
  std::vector<NeighborLocation> FindNeighbors(const LogicalLocation &loc) const;
  // tree.cpp:320-332: the tree's flags where the block touches the tree boundary, block elsewhere
  std::array<BoundaryFlag, 6> GetBlockBCs(const LogicalLocation &loc) const;
  // tree.cpp:297-318 for the two active directions
  void GetBlockDomain(const LogicalLocation &loc, Real xmin[3], Real xmax[3]) const;
  bool IsLeaf(const LogicalLocation &loc) const;
  int NumTrees() const { return static_cast<int>(trees_.size()); }

 private:
  struct TreeNeighbor {
    int tree; // index into trees_
    LogicalCoordinateTransformation ct;
  };
  struct Tree {
    int64_t id = 0;
    std::array<uint64_t, 4> nodes{};
    std::array<Real, 3> xmin{}, xmax{};
    std::array<BoundaryFlag, 6> bcs{};
    std::array<std::vector<TreeNeighbor>, 27> neighbors; // by location index of the offset
    std::set<std::array<int64_t, 3>> leaves;   // (level, lx1, lx2)
    std::set<std::array<int64_t, 3>> internal; // refined positions
  };
  std::vector<Tree> trees_;
  std::map<int64_t, int> index_of_; // tree id -> index
  const Tree &TreeOf(const LogicalLocation &loc) const;
  int Refine(int t, const LogicalLocation &loc);      // tree.cpp:96-137, proper nesting kept
  int AddMeshBlock(int t, const LogicalLocation &loc); // tree.cpp:73-94
};

} // namespace forest
} // namespace parthenon
