// forest — a minimal application on a FOREST of differently oriented trees, the counterpart of
// the reference's example/boundary_exchange (and of the fixture generator
// tests/golden/refgen/forest_dump_main.cpp, which runs the reference on the same forests): nine
// nodes on a 3 x 3 lattice, four faces around the central node, user (= the deck's, outflow by
// default) conditions on the outer edges, one cell-centred field of eight components with
// ProlongatePiecewiseConstant / RestrictAverage.  The problem generator writes a block-dependent
// code into EVERY entry, so after the boundary exchange of Mesh::Initialize every entry tells
// which block, and which cell of it, it came from — through every rotation and reflection.
#pragma once
#include <memory>

#include "pb2/parthenon.hpp"

namespace forest_example {
// forest/variant of the deck (same numbering as forest_dump_main.cpp):
//   0  example/boundary_exchange as shipped: face 0 listed as {n1, n2, n0, n3} (rotated), block
//      (tree 0, level 1, 0, 0) refined;  1  the same faces, no refinement;
//   2  all four faces in different orientations (one a reflection);  3  as 2 with blocks
//      (tree 3, level 1, 1, 0) and (tree 4, level 1, 0, 1) refined
parthenon::forest::ForestDefinition MakeForest(int variant);
parthenon::Packages_t ProcessPackages(std::unique_ptr<parthenon::ParameterInput> &pin);
void MeshProblemGenerator(parthenon::MeshData<parthenon::Real> *md, parthenon::ParameterInput *pin);
} // namespace forest_example
