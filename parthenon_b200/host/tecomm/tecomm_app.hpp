// tecomm — a minimal application that exercises the ghost exchange of NON-CELL-CENTRED fields:
// one package with a face field (2 components), an edge field and a node field, all
// Metadata::FillGhost.  It is the counterpart of the application the parity fixtures were made
// with on the reference (tests/golden/refgen/tecomm_dump_main.cpp): the problem generator
// writes the same block-dependent integer code into EVERY entry, so after the boundary exchange
// of Mesh::Initialize every entry tells which block (and which entry of it) it came from.
#pragma once
#include <memory>

#include "pb2/parthenon.hpp"

namespace tecomm_example {
// adaptive runs (refinement = adaptive): the criterion of the reference-side fixture generator
// tests/golden/refgen/teamr_dump_main.cpp — refine the blocks whose centre lies within 0.2 of a
// point that moves with the cycle number, derefine all others.  The fields never evolve.
void SetCriterionCycle(int cycle);
parthenon::Packages_t ProcessPackages(std::unique_ptr<parthenon::ParameterInput> &pin);
void MeshProblemGenerator(parthenon::MeshData<parthenon::Real> *md, parthenon::ParameterInput *pin);
} // namespace tecomm_example
