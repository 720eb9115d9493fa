// tables.cuh — device-resident region tables shared by exchange.cu and prores.cu.
#pragma once
#include "common.cuh"

namespace pb2 {

constexpr int kThreads = 256;
constexpr int kUnroll = 4;

enum TableKind { kBnd = 0, kCopy = 1, kProRes = 2, kFlxCor = 3, kBc = 4 };

struct DevRegion {
  double *var;       // array side (pack source / unpack destination / copy destination)
  const double *src; // copy source
  int64_t buf_off;
  int32_t s[3];
  int32_t sj, sk, sc;
  int32_t ss[3]; // copy: source start
  int32_t ssj, ssk, ssc;
  FastDiv dni, dnj, dnk; // ni is in units of vectors
  uint32_t total_vec;    // total vectors in the region
  uint32_t vec;          // 1 or 2 doubles per vector
  int32_t flag_slot;
  uint32_t status;
  double value;
  double default_value;
  // unpack through a LogicalCoordinateTransformation (pb2_bnd_region::lcoord_*)
  int32_t tr_on, tr_dir[3], tr_flip[3], tr_ncell;
  double fac;
};

struct Chunk {
  int32_t region;
  uint32_t first_vec;
};

} // namespace pb2

struct pb2_bnd_table {
  int kind;
  int64_t nregions;
  int64_t nchunks;
  int64_t elements;
  pb2::DevRegion *d_regions;
  pb2::Chunk *d_chunks;
  // prores tables keep the API struct on device
  pb2_prores_region *d_prores;
  pb2_flxcor_region *d_flxcor = nullptr;
  pb2_bc_region *d_bc = nullptr;
};

