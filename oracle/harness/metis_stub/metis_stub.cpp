/* metis_stub.cpp -- see metis.h (TEST INFRASTRUCTURE ONLY). */
#include "metis.h"
#include <cstdio>
#include <cstdlib>

static int ReadPartition(idx_t nvtxs, idx_t nparts, idx_t* xadj, idx_t* adjncy, idx_t* edgecut, idx_t* part)
{
  const char* fn = getenv("PCFD_PARTITION_FILE");
  if(!fn){
    fprintf(stderr, "metis_stub: set PCFD_PARTITION_FILE (one partition id per node)\n");
    exit(3);
  }
  FILE* f = fopen(fn, "r");
  if(!f){ perror(fn); exit(3); }
  for(idx_t i = 0; i < nvtxs; i++){
    if(fscanf(f, "%d", &part[i]) != 1 || part[i] < 0 || part[i] >= nparts){
      fprintf(stderr, "metis_stub: bad partition entry %d\n", i);
      exit(3);
    }
  }
  fclose(f);
  idx_t cut = 0;
  for(idx_t i = 0; i < nvtxs; i++){
    for(idx_t k = xadj[i]; k < xadj[i+1]; k++) if(part[adjncy[k]] != part[i]) cut++;
  }
  *edgecut = cut/2;
  return METIS_OK;
}

extern "C" {
int METIS_SetDefaultOptions(idx_t* options)
{
  for(int i = 0; i < METIS_NOPTIONS; i++) options[i] = -1;
  return METIS_OK;
}
int METIS_PartGraphRecursive(idx_t* nvtxs, idx_t*, idx_t* xadj, idx_t* adjncy, idx_t*, idx_t*, idx_t*,
			     idx_t* nparts, real_t*, real_t*, idx_t*, idx_t* edgecut, idx_t* part)
{
  return ReadPartition(*nvtxs, *nparts, xadj, adjncy, edgecut, part);
}
int METIS_PartGraphKway(idx_t* nvtxs, idx_t*, idx_t* xadj, idx_t* adjncy, idx_t*, idx_t*, idx_t*,
			idx_t* nparts, real_t*, real_t*, idx_t*, idx_t* edgecut, idx_t* part)
{
  return ReadPartition(*nvtxs, *nparts, xadj, adjncy, edgecut, part);
}
}
