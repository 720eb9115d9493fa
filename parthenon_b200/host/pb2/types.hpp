// types.hpp — basic host-side types of the B200-native Parthenon hot path.
//
// Mirrors the names and meaning of the reference's basic_types.hpp (Real, TaskStatus,
// AmrTag, SimTime :30-242), defs.hpp (CoordinateDirection, BoundaryFlag, IndexDomain) and
// mesh/domain.hpp (IndexRange :32, IndexShape :83, RegionSize) so application code written
// against Parthenon reads the same.  Host only: no device code lives in this library — every
// kernel is reached through the C ABI in include/parthenon_b200.h.
#pragma once
#include <array>
#include <cstdint>
#include <limits>
#include <sstream>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>

namespace parthenon {

using Real = double; // reference basic_types.hpp:30-39 (double is the default build)

enum class TaskStatus { fail, complete, incomplete, iterate }; // basic_types.hpp:59
enum class TaskListStatus { running, stuck, complete, nothing_to_do };
enum class AmrTag : int { derefine = -1, same = 0, refine = 1 };
enum class DriverStatus { complete, timeout, failed };

enum CoordinateDirection { NODIR = -1, X0DIR = 0, X1DIR = 1, X2DIR = 2, X3DIR = 3 };
enum class BoundaryFlag { block = -1, undef, reflect, outflow, periodic, user };
// where on the cell a value lives (basic_types.hpp:156-166); element % 3 is the index of the
// element inside a face / edge field
enum class TopologicalElement : std::size_t { CC = 0, F1 = 3, F2 = 4, F3 = 5, E1 = 6, E2 = 7, E3 = 8, NN = 9 };
enum class TopologicalType { Cell, Face, Edge, Node };
using TE = TopologicalElement;
// is the element displaced by half a cell from the cell centre along I / J / K
// (basic_types.hpp:195-203)
constexpr int TopologicalOffset(TE el, int dir) {
  return dir == 0   ? (el == TE::F1 || el == TE::E2 || el == TE::E3 || el == TE::NN)
         : dir == 1 ? (el == TE::F2 || el == TE::E3 || el == TE::E1 || el == TE::NN)
                    : (el == TE::F3 || el == TE::E2 || el == TE::E1 || el == TE::NN);
}
constexpr int TopologicalOffsetI(TE el) { return TopologicalOffset(el, 0); }
constexpr int TopologicalOffsetJ(TE el) { return TopologicalOffset(el, 1); }
constexpr int TopologicalOffsetK(TE el) { return TopologicalOffset(el, 2); }
// the elements a field of a topological type holds, in storage order
inline std::vector<TE> GetTopologicalElements(TopologicalType tt) {
  switch (tt) {
  case TopologicalType::Face: return {TE::F1, TE::F2, TE::F3};
  case TopologicalType::Edge: return {TE::E1, TE::E2, TE::E3};
  case TopologicalType::Node: return {TE::NN};
  default: return {TE::CC};
  }
}

enum class IndexDomain {
  entire,
  interior,
  inner_x1,
  outer_x1,
  inner_x2,
  outer_x2,
  inner_x3,
  outer_x3
};

// bvals/comms/bnd_info.hpp / basic_types.hpp BoundaryType
enum class BoundaryType : int {
  local = 0,
  nonlocal,
  any,
  flxcor_send,
  flxcor_recv,
  gmg_same,
  gmg_restrict_send,
  gmg_restrict_recv,
  gmg_prolongate_send,
  gmg_prolongate_recv
};
constexpr int kNumBoundaryTypes = 10;

struct IndexRange {
  int s = 0;
  int e = 0;
};

// Cell-centred index shape of one block (interior + nghost each side; symmetry directions
// collapse to a single cell) — mesh/domain.hpp:83-330.
class IndexShape {
 public:
  IndexShape() = default;
  // nx = 0 (or 1 with sym) marks a symmetry direction
  IndexShape(int nx3, int nx2, int nx1, int ng) {
    const int nx[3] = {nx1, nx2, nx3};
    for (int d = 0; d < 3; ++d) {
      if (nx[d] == 0) {
        x_[d] = IndexRange{0, 0};
        n_[d] = 1;
      } else {
        x_[d] = IndexRange{ng, ng + nx[d] - 1};
        n_[d] = nx[d] + 2 * ng;
      }
    }
  }
  IndexRange GetBoundsI(IndexDomain d) const { return Bounds(0, d); }
  IndexRange GetBoundsJ(IndexDomain d) const { return Bounds(1, d); }
  IndexRange GetBoundsK(IndexDomain d) const { return Bounds(2, d); }
  IndexRange Bounds(int dir, IndexDomain d) const {
    if (d == IndexDomain::interior) return x_[dir];
    return IndexRange{0, n_[dir] - 1};
  }
  // bounds of the entries of topological element `el` (mesh/domain.hpp:162-251): one more
  // entry at the upper end of every direction the element is displaced in
  IndexRange Bounds(int dir, IndexDomain d, TE el) const {
    IndexRange r = Bounds(dir, d);
    if (n_[dir] > 1) r.e += TopologicalOffset(el, dir);
    return r;
  }
  int is(IndexDomain d) const { return Bounds(0, d).s; }
  int ie(IndexDomain d) const { return Bounds(0, d).e; }
  int js(IndexDomain d) const { return Bounds(1, d).s; }
  int je(IndexDomain d) const { return Bounds(1, d).e; }
  int ks(IndexDomain d) const { return Bounds(2, d).s; }
  int ke(IndexDomain d) const { return Bounds(2, d).e; }
  int ncellsi(IndexDomain d) const { return ncells(0, d); }
  int ncellsj(IndexDomain d) const { return ncells(1, d); }
  int ncellsk(IndexDomain d) const { return ncells(2, d); }
  int ncells(int dir, IndexDomain d) const {
    const IndexRange r = Bounds(dir, d);
    return r.e - r.s + 1;
  }
  int GetTotal(IndexDomain d) const { return ncellsi(d) * ncellsj(d) * ncellsk(d); }

 private:
  std::array<IndexRange, 3> x_{};
  std::array<int, 3> n_{1, 1, 1};
};

// mesh/domain.hpp RegionSize: physical extent + cell counts (+ symmetry) of mesh or block
struct RegionSize {
  std::array<Real, 3> xmin_{0, 0, 0}, xmax_{1, 1, 1};
  std::array<int, 3> nx_{1, 1, 1};
  std::array<bool, 3> symmetry_{false, false, false};
  Real xmin(CoordinateDirection d) const { return xmin_[d - 1]; }
  Real xmax(CoordinateDirection d) const { return xmax_[d - 1]; }
  int nx(CoordinateDirection d) const { return nx_[d - 1]; }
  bool symmetry(CoordinateDirection d) const { return symmetry_[d - 1]; }
};

// basic_types.hpp SimTime :208-242
struct SimTime {
  Real start_time = 0.0, time = 0.0, tlim = 0.0, dt = 0.0;
  int ncycle = 0, nlim = -1, ncycle_out = 1, ncycle_out_mesh = 0;
  bool KeepGoing() const { return (time < tlim) && (nlim < 0 || ncycle < nlim); }
};

template <typename T>
using Dictionary = std::unordered_map<std::string, T>;

// utils/error_checking.hpp: hard errors throw (never cross the C ABI as exceptions; the C
// entry points of the host library catch and report)
[[noreturn]] inline void Fail(const std::string &msg, const char *file, int line) {
  std::ostringstream s;
  s << "### PARTHENON ERROR\n  Message:     " << msg << "\n  File:        " << file
    << "\n  Line number: " << line;
  throw std::runtime_error(s.str());
}

} // namespace parthenon

#define PARTHENON_REQUIRE(cond, msg)                                                      \
  do {                                                                                    \
    if (!(cond)) ::parthenon::Fail(msg, __FILE__, __LINE__);                              \
  } while (0)
#define PARTHENON_REQUIRE_THROWS(cond, msg) PARTHENON_REQUIRE(cond, msg)
#define PARTHENON_FAIL(msg) ::parthenon::Fail(msg, __FILE__, __LINE__)
#define PARTHENON_THROW(msg) ::parthenon::Fail(msg, __FILE__, __LINE__)
#define PARTHENON_DEBUG_REQUIRE(cond, msg) ((void)0)
#define PARTHENON_WARN(msg)                                                               \
  do {                                                                                    \
    std::fprintf(stderr, "### PARTHENON WARNING: %s\n", std::string(msg).c_str());        \
  } while (0)
#define PARTHENON_INSTRUMENT
// every C-ABI call is checked; failure is a hard error (there is no CPU fallback)
#define PB2_CHECK(expr)                                                                   \
  do {                                                                                    \
    int rc__ = (expr);                                                                    \
    if (rc__ != 0)                                                                        \
      ::parthenon::Fail(std::string(#expr) + " failed (" + std::to_string(rc__) +         \
                            "): " + pb2_last_error(),                                     \
                        __FILE__, __LINE__);                                              \
  } while (0)
