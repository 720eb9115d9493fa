// bvals_plan.cpp — the pure-topology half of the ghost-zone exchange: index boxes
// (CalcIndices and its flux / topological-element forms), ownership masks resolved into boxes,
// and the channel plan with its per-peer slab layout.  Nothing here touches a device, which is
// what the CPU-only tests exercise through pb2h_topology_create / pb2h_sim_plan_boxes.
// See pb2/bvals.hpp for the design and the reference files each piece replaces.
#include "pb2/bvals.hpp"

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include <tuple>

namespace parthenon {

// ---------------------------------------------------------------------------------------
// CalcIndices for cell-centred, non-flux fields with the identity logical-coordinate
// transform and full ownership (what a single-tree mesh produces): bnd_info.cpp:105-252
// ---------------------------------------------------------------------------------------
IndexBox CalcIndices(const NeighborBlock &nb, const MeshBlock *pmb, IndexRangeType ir_type,
                     bool prores) {
  const LogicalLocation &loc = pmb->loc;
  const int ng = Globals::nghost;
  // prolongation/restriction work in the coarse index space; so does any exchange with a
  // coarser neighbour (:121-125)
  const bool use_coarse = prores || nb.loc.level < loc.level;
  const IndexShape &shape = use_coarse ? pmb->c_cellbounds : pmb->cellbounds;
  const int coarse_fac = nb.loc.level > loc.level ? 2 : 1; // :130
  const int interior_offset = ir_type == IndexRangeType::BoundaryInteriorSend ? ng : 0;
  int exterior_offset = ir_type == IndexRangeType::BoundaryExteriorRecv ? ng : 0;
  if (prores) exterior_offset /= 2; // only coarse ghosts that have fine ghosts (:161-166)
  IndexBox box;
  for (int d = 0; d < 3; ++d) {
    const bool not_sym = !pmb->block_size.symmetry_[d];
    const IndexRange b = shape.Bounds(d, IndexDomain::interior);
    // the neighbour's interior extent in the index space we exchange in (:131-135)
    const int nb_n = not_sym ? pmb->block_size.nx_[d] / coarse_fac : 1;
    int &s = box.s[d], &e = box.e[d];
    if (nb.offsets[d] == 0) {
      s = b.s;
      e = b.e;
      if (loc.level < nb.origin_loc.level && not_sym) {
        // finer neighbour abuts half of this face; send ng extra interior zones so the
        // receiver can prolongate (:173-192)
        const int extra = (b.e - b.s + 1) - nb_n;
        const bool upper_half = ((nb.origin_loc.lx[d] % 2) + 2) % 2 == 1;
        s += upper_half ? extra - interior_offset : 0;
        e -= upper_half ? 0 : extra - interior_offset;
        if (ir_type == IndexRangeType::InteriorSend && !prores) {
          s -= ng;
          e += ng;
        }
      }
      if (loc.level > nb.origin_loc.level && not_sym) {
        // coarser neighbour: it sent extra zones on the side away from our corner (:193-204)
        s -= loc.lx[d] % 2 == 1 ? exterior_offset : 0;
        e += loc.lx[d] % 2 == 0 ? exterior_offset : 0;
        if (ir_type == IndexRangeType::InteriorRecv && !prores) {
          s -= ng;
          e += ng;
        }
      }
      if (prores && not_sym && ir_type == IndexRangeType::InteriorRecv) {
        s -= ng / 2;
        e += ng / 2;
      }
    } else if (nb.offsets[d] > 0) {
      s = b.e + (-interior_offset + 1);
      e = b.e + exterior_offset;
    } else {
      s = b.s - exterior_offset;
      e = b.s + (interior_offset - 1);
    }
  }
  return box;
}

IndexBox CalcIndicesTE(const NeighborBlock &nb, const MeshBlock *pmb, TE el,
                       IndexRangeType ir_type, bool prores) {
  const LogicalLocation &loc = pmb->loc;
  const int ng = Globals::nghost;
  const bool use_coarse = prores || nb.loc.level < loc.level;
  const IndexShape &shape = use_coarse ? pmb->c_cellbounds : pmb->cellbounds;
  const int coarse_fac = nb.loc.level > loc.level ? 2 : 1;
  const int interior_offset = ir_type == IndexRangeType::BoundaryInteriorSend ? ng : 0;
  int exterior_offset = ir_type == IndexRangeType::BoundaryExteriorRecv ? ng : 0;
  if (prores) exterior_offset /= 2;
  IndexBox box;
  for (int d = 0; d < 3; ++d) {
    const bool not_sym = !pmb->block_size.symmetry_[d];
    const IndexRange b = shape.Bounds(d, IndexDomain::interior, el);
    const int top = not_sym ? TopologicalOffset(el, d) : 0;
    // entries of the element the neighbour holds along d in the index space of the exchange
    const int nb_n = not_sym ? pmb->block_size.nx_[d] / coarse_fac + top : 1;
    int &s = box.s[d], &e = box.e[d];
    if (nb.offsets[d] == 0) {
      s = b.s;
      e = b.e;
      if (loc.level < nb.origin_loc.level && not_sym) { // :173-192
        const int extra = (b.e - b.s + 1) - nb_n;
        const bool upper_half = ((nb.origin_loc.lx[d] % 2) + 2) % 2 == 1;
        s += upper_half ? extra - interior_offset : 0;
        e -= upper_half ? 0 : extra - interior_offset;
      }
      if (loc.level > nb.origin_loc.level && not_sym) { // :193-204
        s -= loc.lx[d] % 2 == 1 ? exterior_offset : 0;
        e += loc.lx[d] % 2 == 0 ? exterior_offset : 0;
      }
    } else if (nb.offsets[d] > 0) {
      // a neighbour duplicates the shared elements: its boundary lies one deeper (:150-155)
      s = b.e + (-interior_offset + 1 - top);
      e = b.e + exterior_offset;
    } else {
      s = b.s - exterior_offset;
      e = b.s + (interior_offset - 1 + top);
    }
  }
  return box;
}

// bnd_info.cpp:105-252 with flux = true for one element of an edge-centred flux field (the flux
// of a face field): along a direction the neighbour is offset in, the box is the one plane of
// the element's interior range that lies on the boundary (:207-218); tangentially as above
IndexBox CalcIndicesFluxTE(const NeighborBlock &nb, const MeshBlock *pmb, TE el,
                           IndexRangeType ir_type, bool prores) {
  IndexBox box = CalcIndicesTE(nb, pmb, el, ir_type, prores);
  const bool use_coarse = prores || nb.loc.level < pmb->loc.level;
  const IndexShape &shape = use_coarse ? pmb->c_cellbounds : pmb->cellbounds;
  for (int d = 0; d < 3; ++d) {
    if (nb.offsets[d] == 0) continue;
    const IndexRange b = shape.Bounds(d, IndexDomain::interior, el);
    box.s[d] = box.e[d] = nb.offsets[d] > 0 ? b.e : b.s;
  }
  return box;
}

// GetFluxCorrectionElements (bnd_info.cpp:71-103) for an edge-centred flux field: the two edge
// elements tangent to a shared face, the one along a shared block edge; none across a corner
std::vector<TE> FluxCorrectionEdgeElements(const int off[3]) {
  const int nz = (off[0] != 0) + (off[1] != 0) + (off[2] != 0);
  if (nz == 1) {
    if (off[0]) return {TE::E2, TE::E3};
    if (off[1]) return {TE::E3, TE::E1};
    return {TE::E1, TE::E2};
  }
  if (nz == 2) return {off[0] == 0 ? TE::E1 : (off[1] == 0 ? TE::E2 : TE::E3)};
  return {};
}

std::array<bool, 27> RecvMask(const Mesh *pm, const NeighborBlock &nb, const MeshBlock *pmb,
                              TE el) {
  int sox[3] = {-nb.offsets[0], -nb.offsets[1], -nb.offsets[2]};
  if (nb.origin_loc.level < pmb->loc.level)
    // a coarser sender passes zones from an interior corner only, never a whole face or edge
    for (int d = 0; d < 3; ++d)
      if (sox[d] == 0) sox[d] = pmb->loc.lx[d] % 2 == 1 ? 1 : -1;
  return IndexRangeMask(el, pm->Ownership(nb.gid), sox);
}

std::array<bool, 27> IndexRangeMask(TE el, const std::array<bool, 27> &sender,
                                    const int sox[3]) {
  auto at = [](int i, int j, int k) { return (i + 1) + 3 * (j + 1) + 9 * (k + 1); };
  std::array<bool, 27> m;
  // block ownership -> ownership of the element's entries: directions the element is not
  // displaced in have no shared entries and follow the block's interior
  for (int i = -1; i <= 1; ++i)
    for (int j = -1; j <= 1; ++j)
      for (int k = -1; k <= 1; ++k)
        m[at(i, j, k)] = sender[at(TopologicalOffsetI(el) ? i : 0, TopologicalOffsetJ(el) ? j : 0,
                                   TopologicalOffsetK(el) ? k : 0)];
  // the box is a slice of the sender next to its (sox) boundary: the side of it that faces the
  // sender's interior holds interior entries
  if (sox[0] != 0)
    for (int j = -1; j <= 1; ++j)
      for (int k = -1; k <= 1; ++k) m[at(-sox[0], j, k)] = m[at(0, j, k)];
  if (sox[1] != 0)
    for (int i = -1; i <= 1; ++i)
      for (int k = -1; k <= 1; ++k) m[at(i, -sox[1], k)] = m[at(i, 0, k)];
  if (sox[2] != 0)
    for (int i = -1; i <= 1; ++i)
      for (int j = -1; j <= 1; ++j) m[at(i, j, -sox[2])] = m[at(i, j, 0)];
  return m;
}

std::vector<IndexBox> ActivePieces(const int n[3], const std::array<bool, 27> &mask) {
  // per direction: first entry (-1), inner run (0), last entry (+1); a single entry counts as
  // inner (indexer.hpp:171-173: (i == end) - (i == start))
  struct Seg {
    int s, e, idx;
  };
  std::vector<Seg> segs[3];
  for (int d = 0; d < 3; ++d) {
    if (n[d] == 1) {
      segs[d] = {{0, 0, 0}};
    } else {
      segs[d].push_back({0, 0, -1});
      if (n[d] > 2) segs[d].push_back({1, n[d] - 2, 0});
      segs[d].push_back({n[d] - 1, n[d] - 1, 1});
    }
  }
  std::vector<IndexBox> out;
  for (const Seg &k : segs[2])
    for (const Seg &j : segs[1])
      for (const Seg &i : segs[0]) {
        if (!mask[(i.idx + 1) + 3 * (j.idx + 1) + 9 * (k.idx + 1)]) continue;
        IndexBox b;
        b.s[0] = i.s, b.e[0] = i.e, b.s[1] = j.s, b.e[1] = j.e, b.s[2] = k.s, b.e[2] = k.e;
        out.push_back(b);
      }
  // glue boxes that share a whole side (fewer, longer regions); deterministic, so the sending
  // and the receiving device arrive at the same list
  bool merged = true;
  while (merged) {
    merged = false;
    for (size_t a = 0; a < out.size() && !merged; ++a)
      for (size_t b = 0; b < out.size() && !merged; ++b) {
        if (a == b) continue;
        for (int d = 0; d < 3 && !merged; ++d) {
          const int d1 = (d + 1) % 3, d2 = (d + 2) % 3;
          if (out[a].e[d] + 1 == out[b].s[d] && out[a].s[d1] == out[b].s[d1] &&
              out[a].e[d1] == out[b].e[d1] && out[a].s[d2] == out[b].s[d2] &&
              out[a].e[d2] == out[b].e[d2]) {
            out[a].e[d] = out[b].e[d];
            out.erase(out.begin() + static_cast<std::ptrdiff_t>(b));
            merged = true;
          }
        }
      }
  }
  return out;
}

// bnd_info.cpp:105-252 with flux = true and el = F_dir: the box is the shared face itself
// (:207-211), tangentially the whole face of a finer sender (coarse index space) or the half
// (quarter in 3-D) of a coarser receiver's face that the finer neighbour abuts (:173-192)
IndexBox CalcIndicesFlux(const NeighborBlock &nb, const MeshBlock *pmb) {
  const LogicalLocation &loc = pmb->loc;
  const bool use_coarse = nb.loc.level < loc.level;
  const IndexShape &shape = use_coarse ? pmb->c_cellbounds : pmb->cellbounds;
  const int coarse_fac = nb.loc.level > loc.level ? 2 : 1;
  IndexBox box;
  for (int d = 0; d < 3; ++d) {
    const bool not_sym = !pmb->block_size.symmetry_[d];
    const IndexRange b = shape.Bounds(d, IndexDomain::interior);
    int &s = box.s[d], &e = box.e[d];
    if (nb.offsets[d] == 0) {
      s = b.s;
      e = b.e;
      if (loc.level < nb.origin_loc.level && not_sym) {
        const int extra = (b.e - b.s + 1) - pmb->block_size.nx_[d] / coarse_fac;
        const bool upper_half = ((nb.origin_loc.lx[d] % 2) + 2) % 2 == 1;
        s += upper_half ? extra : 0;
        e -= upper_half ? 0 : extra;
      }
    } else if (nb.offsets[d] > 0) {
      s = e = b.e + (not_sym ? 1 : 0); // upper face: index ie + 1 of the cell-aligned array
    } else {
      s = e = b.s;
    }
  }
  return box;
}

// ---------------------------------------------------------------------------------------
// channel plan (pure topology)
// ---------------------------------------------------------------------------------------
int OffsetIndexOf(int o1, int o2, int o3) { return (o1 + 1) + 3 * (o2 + 1) + 9 * (o3 + 1); }

// the entry of `sender`'s neighbour list that describes the channel towards `receiver_gid`
// seen from the receiver at offsets `roff` (bvals_utils.hpp:43-67: channels are keyed by
// sender gid, receiver gid and the location index)
const NeighborBlock *MatchingNeighbor(const MeshBlock *sender, int receiver_gid,
                                      const int roff[3]) {
  for (auto &q : sender->neighbors)
    if (q.gid == receiver_gid && q.offsets[0] == -roff[0] && q.offsets[1] == -roff[1] &&
        q.offsets[2] == -roff[2])
      return &q;
  return nullptr;
}

// The offsets of neighbour `nb` as the NEIGHBOUR's tree sees them (ReceiveKey,
// bvals_utils.hpp:57-67: lcoord_trans.Transform(nb.offsets)): the sender's own offsets towards
// the receiver are the reverse of these.
std::array<int, 3> OffsetsInSenderFrame(const NeighborBlock &nb) {
  std::array<int, 3> off{nb.offsets[0], nb.offsets[1], nb.offsets[2]};
  return nb.transformed ? nb.lcoord_trans.Transform(off) : off;
}

// bnd_info.cpp:216-228: a receive box goes to the sender's logical coordinates — both corners
// through LogicalCoordinateTransformation::Transform(ijk) (.hpp:79-88) with ncell = the array's
// extent along x1, re-sorted — and SetBounds walks THAT box in buffer order, writing each element
// at InverseTransform (boundary_communication.cpp:282-308)
IndexBox TransformBox(const IndexBox &box, const forest::LogicalCoordinateTransformation &ct,
                      int ncell) {
  IndexBox out;
  for (int d = 0; d < 3; ++d) {
    const int o = std::abs(ct.dir_connection[d]);
    out.s[o] = ct.dir_flip[d] ? ncell - 1 - box.s[d] : box.s[d];
    out.e[o] = ct.dir_flip[d] ? ncell - 1 - box.e[d] : box.e[d];
  }
  for (int d = 0; d < 3; ++d)
    if (out.s[d] > out.e[d]) std::swap(out.s[d], out.e[d]);
  return out;
}
namespace {
auto ChannelKey(const Channel &c) {
  return std::make_tuple(c.sender_gid, c.receiver_gid, c.var, c.offset_index, c.piece);
}
IndexBox SubBox(const IndexBox &box, const IndexBox &rel) {
  IndexBox b;
  for (int d = 0; d < 3; ++d) {
    b.s[d] = box.s[d] + rel.s[d];
    b.e[d] = box.s[d] + rel.e[d];
  }
  return b;
}
} // namespace

EdgeFluxPlan BuildEdgeFluxPlan(const Mesh *pm, const BlockList_t &blocks, int ncomp) {
  EdgeFluxPlan plan;
  const int V = pm->virtual_ranks > 1 ? pm->virtual_ranks : 1;
  plan.npeers = V > 1 ? V * V : pm->nranks;
  for (auto &pmb : blocks) {
    const int my_vr = pm->VirtualRankOf(pmb->gid);
    for (auto &nb : pmb->neighbors) {
      // ForEachBoundary<flxcor_*> (loop_utils.hpp:134-158): neighbours one level apart across a
      // face or a block edge
      if (std::abs(nb.loc.level - pmb->loc.level) != 1) continue;
      const std::vector<TE> els = FluxCorrectionEdgeElements(nb.offsets);
      if (els.empty()) continue;
      PARTHENON_REQUIRE(!nb.transformed, "flux correction of face fields on forests is not built");
      const int nb_vr = nb.rank == pm->my_rank ? pm->VirtualRankOf(nb.gid) : 0;
      const bool local = nb.rank == pm->my_rank && nb_vr == my_vr;
      const int pass = els.size() == 2 ? 1 : 0;
      const bool pmb_sends = nb.loc.level == pmb->loc.level - 1;
      // the fine SENDER's half of a same-device pair is listed from the receiver's side below, so
      // that a partition's tables hold everything its own blocks need, in order
      if (pmb_sends && local) continue;
      for (TE el : els) {
        const int e = static_cast<int>(el) % 3; // E1, E2, E3 -> 0, 1, 2 in storage order
        if (pmb_sends) {
          // fine block of this rank, coarser neighbour on another device: restrict the shared
          // elements into the coarse buffer (ProResInfo::GetSend, bnd_info.cpp:387-403) and send
          // the entries this block owns (bnd_info.cpp:232-248; the mask is taken in the sender's
          // view of the receiver, the same numbers the receiver derives below)
          plan.send_restricts.push_back(
              {pmb->gid, e,
               CalcIndicesFluxTE(nb, pmb.get(), el, IndexRangeType::BoundaryInteriorSend, true)});
          const IndexBox sbox =
              CalcIndicesFluxTE(nb, pmb.get(), el, IndexRangeType::BoundaryInteriorSend, false);
          const int n[3] = {sbox.n(0), sbox.n(1), sbox.n(2)};
          const auto mask = IndexRangeMask(el, pm->Ownership(pmb->gid), nb.offsets);
          int sub = 0;
          for (const IndexBox &rel : ActivePieces(n, mask)) {
            EdgeFluxPiece pc{pmb->gid, nb.gid, e, pass, SubBox(sbox, rel), IndexBox{}};
            pc.seg = V > 1 ? my_vr * V + nb_vr : nb.rank;
            pc.offset_index = nb.OffsetIndex();
            pc.sub = sub++;
            plan.send.push_back(pc);
          }
          continue;
        }
        // coarse RECEIVER: the sender's coarse-buffer box -> this block's flux array, the
        // entries the sender owns
        const IndexBox rbox =
            CalcIndicesFluxTE(nb, pmb.get(), el, IndexRangeType::BoundaryExteriorRecv, false);
        const int n[3] = {rbox.n(0), rbox.n(1), rbox.n(2)};
        const int sox[3] = {-nb.offsets[0], -nb.offsets[1], -nb.offsets[2]};
        const auto mask = IndexRangeMask(el, pm->Ownership(nb.gid), sox);
        if (!local) {
          int sub = 0;
          for (const IndexBox &rel : ActivePieces(n, mask)) {
            EdgeFluxPiece pc{nb.gid, pmb->gid, e, pass, IndexBox{}, SubBox(rbox, rel)};
            pc.seg = V > 1 ? nb_vr * V + my_vr : nb.rank;
            pc.offset_index = OffsetIndexOf(sox[0], sox[1], sox[2]);
            pc.sub = sub++;
            plan.recv.push_back(pc);
          }
          continue;
        }
        const MeshBlock *sb = pm->block_list[nb.lid].get();
        const NeighborBlock *q = MatchingNeighbor(sb, pmb->gid, nb.offsets);
        PARTHENON_REQUIRE(q != nullptr, "no matching flux-correction sender");
        // the sender restricts (RestrictAverage) the shared elements into its coarse buffer
        plan.restricts.push_back(
            {nb.gid, e, CalcIndicesFluxTE(*q, sb, el, IndexRangeType::BoundaryInteriorSend, true)});
        const IndexBox sbox =
            CalcIndicesFluxTE(*q, sb, el, IndexRangeType::BoundaryInteriorSend, false);
        for (int d = 0; d < 3; ++d)
          PARTHENON_REQUIRE(sbox.n(d) == n[d], "flux-correction extents differ");
        for (const IndexBox &rel : ActivePieces(n, mask))
          plan.pieces.push_back({nb.gid, pmb->gid, e, pass, SubBox(sbox, rel), SubBox(rbox, rel)});
      }
    }
  }
  // both sides order a peer segment by the same key => identical offsets, no handshake
  auto layout = [&](std::vector<EdgeFluxPiece> &pcs, std::vector<int64_t> &seg_off, bool send) {
    std::stable_sort(pcs.begin(), pcs.end(), [](const EdgeFluxPiece &a, const EdgeFluxPiece &b) {
      return std::make_tuple(a.seg, a.sender_gid, a.receiver_gid, a.offset_index, a.el, a.sub) <
             std::make_tuple(b.seg, b.sender_gid, b.receiver_gid, b.offset_index, b.el, b.sub);
    });
    std::vector<int64_t> seg_size(plan.npeers, 0);
    for (EdgeFluxPiece &pc : pcs) {
      const int64_t n = static_cast<int64_t>(ncomp) * (send ? pc.send_box : pc.recv_box).size();
      pc.slab_off = seg_size[pc.seg];
      seg_size[pc.seg] += n + (n & 1); // 16-byte aligned pieces
    }
    seg_off.assign(plan.npeers + 1, 0);
    for (int p = 0; p < plan.npeers; ++p) seg_off[p + 1] = seg_off[p] + seg_size[p];
  };
  layout(plan.send, plan.send_off, true);
  layout(plan.recv, plan.recv_off, false);
  return plan;
}

ExchangePlan BuildExchangePlan(const Mesh *pm, const BlockList_t &blocks,
                               const std::vector<PlanVar> &vars) {
  ExchangePlan plan;
  const int V = pm->virtual_ranks > 1 ? pm->virtual_ranks : 1;
  PARTHENON_REQUIRE(!(V > 1 && pm->nranks > 1), "pb2/virtual_ranks needs a single real rank");
  plan.npeers = V > 1 ? V * V : pm->nranks;
  const int nvar = static_cast<int>(vars.size());
  std::vector<std::pair<int, Channel>> send, recv; // (segment, channel)
  // a face / edge / node channel: one piece per element and active sub-box of the ownership
  // mask of its sender; `emit` receives each piece with boxes, component range and piece id set
  auto pieces = [&](Channel base, const NeighborBlock &nb, const MeshBlock *pmb, bool pmb_sends,
                    const MeshBlock *local_sender, const NeighborBlock *q, auto &&emit) {
    const PlanVar &pv = vars[base.var];
    const std::vector<TE> els = GetTopologicalElements(pv.tt);
    // the mask is the RECEIVER's (bnd_info.cpp:232-248): offsets of the receiver seen from the
    // sender, replaced by the receiver's position inside its parent where a coarser sender
    // faces it with offset 0
    const LogicalLocation &recv_loc = pmb_sends ? nb.origin_loc : pmb->loc;
    const int sender_gid = pmb_sends ? pmb->gid : nb.gid;
    const int sender_level = pmb_sends ? pmb->loc.level : nb.loc.level;
    int sox[3];
    for (int d = 0; d < 3; ++d) {
      sox[d] = pmb_sends ? nb.offsets[d] : -nb.offsets[d];
      if (sender_level < recv_loc.level && sox[d] == 0)
        sox[d] = ((recv_loc.lx[d] % 2) + 2) % 2 == 1 ? 1 : -1;
    }
    for (size_t e = 0; e < els.size(); ++e) {
      const IndexBox mine = CalcIndicesTE(nb, pmb, els[e],
                                          pmb_sends ? IndexRangeType::BoundaryInteriorSend
                                                    : IndexRangeType::BoundaryExteriorRecv);
      IndexBox other = mine;
      if (local_sender)
        other = CalcIndicesTE(*q, local_sender, els[e], IndexRangeType::BoundaryInteriorSend);
      int n[3];
      for (int d = 0; d < 3; ++d) {
        n[d] = mine.n(d);
        PARTHENON_REQUIRE(other.n(d) == n[d], "send/receive extents of a channel differ");
      }
      const auto mask = IndexRangeMask(els[e], pm->Ownership(sender_gid), sox);
      int sub = 0;
      for (const IndexBox &rel : ActivePieces(n, mask)) {
        Channel c = base;
        c.piece = static_cast<int>(e) * 32 + sub++;
        c.comp0 = static_cast<int>(e) * pv.ncomp;
        c.ncomp = pv.ncomp;
        c.send_box = SubBox(pmb_sends ? mine : other, rel);
        c.recv_box = SubBox(pmb_sends ? other : mine, rel);
        emit(c);
      }
    }
  };
  // Channels of different blocks are independent: built on all host threads into per-block
  // lists that are concatenated in block order (the order of the serial loop).  Face / edge /
  // node fields consult the lazily filled ownership cache of the mesh, so they stay serial.
  bool all_cell = true;
  for (const PlanVar &pv : vars) all_cell = all_cell && pv.tt == TopologicalType::Cell;
  const int nblk = static_cast<int>(blocks.size());
  static const bool timing = std::getenv("PB2_TIME_HOST") != nullptr;
  const auto tp0 = std::chrono::steady_clock::now();
  std::vector<std::vector<Channel>> local_b(nblk);
  std::vector<std::vector<std::pair<int, Channel>>> send_b(nblk), recv_b(nblk);
  std::string failure;
#pragma omp parallel for schedule(dynamic, 16) if (all_cell && nblk > 256)
  for (int ib = 0; ib < nblk; ++ib) {
   try {
    const auto &pmb = blocks[ib];
    std::vector<Channel> &local_out = local_b[ib];
    std::vector<std::pair<int, Channel>> &send = send_b[ib], &recv = recv_b[ib];
    const int my_vr = pm->VirtualRankOf(pmb->gid);
    for (auto &nb : pmb->neighbors) {
      const int nb_vr = nb.rank == pm->my_rank ? pm->VirtualRankOf(nb.gid) : 0;
      // a neighbour in a differently oriented tree is unpacked through its transformation, which
      // the fused same-device copy does not do: such channels take the slab path
      const bool local = nb.rank == pm->my_rank && nb_vr == my_vr && !nb.transformed;
      const std::array<int, 3> soff = OffsetsInSenderFrame(nb);
      for (int v = 0; v < nvar; ++v) {
        PARTHENON_REQUIRE(!nb.transformed || vars[v].tt == TopologicalType::Cell,
                          "forests of rotated trees exchange cell-centred fields only");
        // this block as RECEIVER of the channel nb -> pmb
        Channel rc;
        rc.sender_gid = nb.gid;
        rc.receiver_gid = pmb->gid;
        rc.var = v;
        rc.offset_index = OffsetIndexOf(-soff[0], -soff[1], -soff[2]);
        rc.recv_box = CalcIndices(nb, pmb.get(), IndexRangeType::BoundaryExteriorRecv, false);
        rc.recv_coarse = nb.loc.level < pmb->loc.level; // bnd_info.cpp:285-289
        if (nb.transformed) {
          rc.transformed = true;
          rc.lcoord_trans = nb.lcoord_trans;
          const IndexShape &shape = rc.recv_coarse ? pmb->c_cellbounds : pmb->cellbounds;
          rc.ncell = shape.Bounds(0, IndexDomain::entire).e + 1; // var.GetDim(1), bnd_info.cpp:297
          rc.recv_box = TransformBox(rc.recv_box, nb.lcoord_trans, rc.ncell);
        }
        rc.send_coarse = pmb->loc.level < nb.loc.level;
        rc.sender_rank = nb.rank;
        rc.receiver_rank = pm->my_rank;
        rc.sender_vrank = nb_vr;
        rc.receiver_vrank = my_vr;
        rc.send_box = rc.recv_box;
        rc.ncomp = vars[v].ncomp;
        const bool cell = vars[v].tt == TopologicalType::Cell;
        if (local) {
          const MeshBlock *sender = pm->block_list[nb.lid].get();
          const int roff[3] = {soff[0], soff[1], soff[2]};
          const NeighborBlock *q = MatchingNeighbor(sender, pmb->gid, roff);
          PARTHENON_REQUIRE(q != nullptr, "no matching send region for a local channel");
          auto emit = [&](const Channel &c) { local_out.push_back(c); };
          if (cell) {
            rc.send_box = CalcIndices(*q, sender, IndexRangeType::BoundaryInteriorSend, false);
            for (int d = 0; d < 3; ++d)
              PARTHENON_REQUIRE(rc.send_box.n(d) == rc.recv_box.n(d),
                                "send/receive extents of a channel differ");
            emit(rc);
          } else {
            pieces(rc, nb, pmb.get(), false, sender, q, emit);
          }
        } else {
          const int seg = V > 1 ? nb_vr * V + my_vr : nb.rank;
          auto emit = [&](const Channel &c) { recv.emplace_back(seg, c); };
          if (cell)
            emit(rc);
          else
            pieces(rc, nb, pmb.get(), false, nullptr, nullptr, emit);
        }
        // this block as SENDER of the channel pmb -> nb
        if (!local) {
          Channel sc;
          sc.sender_gid = pmb->gid;
          sc.receiver_gid = nb.gid;
          sc.var = v;
          sc.offset_index = nb.OffsetIndex();
          sc.send_box = CalcIndices(nb, pmb.get(), IndexRangeType::BoundaryInteriorSend, false);
          sc.recv_box = sc.send_box;
          sc.send_coarse = nb.loc.level < pmb->loc.level;
          sc.recv_coarse = pmb->loc.level < nb.loc.level;
          sc.sender_rank = pm->my_rank;
          sc.receiver_rank = nb.rank;
          sc.sender_vrank = my_vr;
          sc.receiver_vrank = nb_vr;
          sc.ncomp = vars[v].ncomp;
          const int seg = V > 1 ? my_vr * V + nb_vr : nb.rank;
          auto emit = [&](const Channel &c) { send.emplace_back(seg, c); };
          if (vars[v].tt == TopologicalType::Cell)
            emit(sc);
          else
            pieces(sc, nb, pmb.get(), true, nullptr, nullptr, emit);
        }
      }
    }
   } catch (const std::exception &e) {
#pragma omp critical
    failure = e.what();
   }
  }
  PARTHENON_REQUIRE(failure.empty(), failure);
  const auto tp1 = std::chrono::steady_clock::now();
  // concatenate in block order (the order of a serial loop): prefix sums of the list lengths,
  // then every block copies its own lists into place — on an adaptive mesh of a few thousand
  // blocks these are tens of MB
  std::vector<size_t> lo(nblk + 1, 0), so(nblk + 1, 0), ro(nblk + 1, 0);
  for (int ib = 0; ib < nblk; ++ib) {
    lo[ib + 1] = lo[ib] + local_b[ib].size();
    so[ib + 1] = so[ib] + send_b[ib].size();
    ro[ib + 1] = ro[ib] + recv_b[ib].size();
  }
  plan.local.resize(lo[nblk]);
  send.resize(so[nblk]);
  recv.resize(ro[nblk]);
  int64_t local_elements = 0;
#pragma omp parallel for schedule(static) reduction(+ : local_elements) if (nblk > 256)
  for (int ib = 0; ib < nblk; ++ib) {
    for (const Channel &c : local_b[ib]) local_elements += c.recv_box.size() * c.ncomp;
    std::copy(local_b[ib].begin(), local_b[ib].end(), plan.local.begin() + lo[ib]);
    std::copy(send_b[ib].begin(), send_b[ib].end(), send.begin() + so[ib]);
    std::copy(recv_b[ib].begin(), recv_b[ib].end(), recv.begin() + ro[ib]);
    std::vector<Channel>().swap(local_b[ib]); // freed by the thread that holds it in cache
  }
  plan.local_elements = local_elements;
  // both sides of a peer segment order its channels by the same key, so slab offsets agree
  // without any handshake
  auto layout = [&](std::vector<std::pair<int, Channel>> &chs, std::vector<Channel> &out,
                    std::vector<int64_t> &seg_off, int64_t &total) {
    std::stable_sort(chs.begin(), chs.end(), [](const auto &a, const auto &b) {
      if (a.first != b.first) return a.first < b.first;
      return ChannelKey(a.second) < ChannelKey(b.second);
    });
    seg_off.assign(plan.npeers + 1, 0);
    std::vector<int64_t> seg_size(plan.npeers, 0);
    for (auto &sc : chs) {
      Channel c = sc.second;
      c.slab_off = seg_size[sc.first];
      int64_t n = c.send_box.size() * c.ncomp;
      n += n & 1; // keep every channel 16-byte aligned for vector access
      seg_size[sc.first] += n;
      out.push_back(c);
    }
    for (int p = 0; p < plan.npeers; ++p) seg_off[p + 1] = seg_off[p] + seg_size[p];
    total = seg_off[plan.npeers];
    // make slab_off absolute
    size_t i = 0;
    for (auto &sc : chs) out[i++].slab_off += seg_off[sc.first];
  };
  layout(send, plan.send, plan.send_off, plan.send_elements);
  layout(recv, plan.recv, plan.recv_off, plan.recv_elements);
  if (timing) {
    const auto tp2 = std::chrono::steady_clock::now();
    auto ms = [](auto a, auto b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
    std::fprintf(stderr, "[pb2 plan] %d blocks: channels %.2f ms, concatenate + layout %.2f ms\n", nblk,
                 ms(tp0, tp1), ms(tp1, tp2));
  }
  return plan;
}

} // namespace parthenon
