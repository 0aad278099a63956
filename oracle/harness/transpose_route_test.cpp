// transpose_route_test.cpp -- TEST INFRASTRUCTURE: runs pcfd::RouteTransposedGhostBlocks (include/pcfd_host.hpp, the
// MPI routing DropIn::CRSTranspose uses for PObj::TransposeCommCRS, parallel.tcc:54-338) over the process-based MPI shim,
// without a device: every rank reads <dir>/route_in.<rank>.bin written by tests/test_crs_transpose.py
//   int32: nnode, ngedge, n2, gnode; int32[2*ngedge] ghost half-edges; int32[gnode] gNodeOwner; int32[gnode] gNodeLocalId;
//   float64[ngedge*n2] ghost-column blocks after the local transpose
// and writes the routed blocks to <dir>/route_out.<rank>.bin.  PCFD_MPI_NP ranks (mpi_shim).
#include <mpi.h>

#include <cstdio>
#include <vector>

#define PCFD_HOST_ROUTING_ONLY
#include "pcfd_host.hpp"

int main(int argc, char** argv) {
  MPI_Init(&argc, &argv);
  int rank = 0;
  MPI_Comm_rank(MPI_COMM_WORLD, &rank);
  if (argc < 2) return 2;
  char path[4096];
  std::snprintf(path, sizeof path, "%s/route_in.%d.bin", argv[1], rank);
  FILE* f = std::fopen(path, "rb");
  if (!f) { std::perror(path); return 3; }
  int hdr[4];
  if (std::fread(hdr, sizeof(int), 4, f) != 4) return 4;
  const int nnode = hdr[0], ngedge = hdr[1], n2 = hdr[2], gnode = hdr[3];
  std::vector<int> ge(2 * (size_t)ngedge + 1), owner((size_t)gnode + 1), lid((size_t)gnode + 1);
  std::vector<double> blocks((size_t)ngedge * n2 + 1);
  bool ok = std::fread(ge.data(), sizeof(int), 2 * (size_t)ngedge, f) == 2 * (size_t)ngedge;
  ok = ok && std::fread(owner.data(), sizeof(int), gnode, f) == (size_t)gnode;
  ok = ok && std::fread(lid.data(), sizeof(int), gnode, f) == (size_t)gnode;
  ok = ok && std::fread(blocks.data(), sizeof(double), (size_t)ngedge * n2, f) == (size_t)ngedge * n2;
  std::fclose(f);
  if (!ok) return 5;
  const bool routed = pcfd::RouteTransposedGhostBlocks(nnode, ngedge, ge.data(), owner.data(), lid.data(), n2, blocks.data());
  std::snprintf(path, sizeof path, "%s/route_out.%d.bin", argv[1], rank);
  f = std::fopen(path, "wb");
  std::fwrite(blocks.data(), sizeof(double), (size_t)ngedge * n2, f);
  std::fclose(f);
  MPI_Finalize();
  return routed ? 0 : 6;
}
