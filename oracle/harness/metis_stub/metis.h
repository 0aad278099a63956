/*
 * metis.h -- stand-in for METIS 5 used ONLY to compile the reference's
 * ucs/decomp.cpp as part of the parity oracle (TEST INFRASTRUCTURE ONLY).
 * The graph partitioner itself is not on the hot path: the halo maps the
 * reference derives (decomp.cpp:122-273, parallel.tcc:461-554) are a pure
 * function of the partition vector, which this stub reads from the file named
 * by $PCFD_PARTITION_FILE (one int per node) instead of computing it.
 */
#ifndef PCFD_METIS_STUB_H
#define PCFD_METIS_STUB_H
#ifdef __cplusplus
extern "C" {
#endif
typedef int idx_t;
typedef float real_t;
#define METIS_NOPTIONS 40
#define METIS_OK 1
int METIS_SetDefaultOptions(idx_t* options);
int METIS_PartGraphRecursive(idx_t* nvtxs, idx_t* ncon, idx_t* xadj, idx_t* adjncy, idx_t* vwgt,
			     idx_t* vsize, idx_t* adjwgt, idx_t* nparts, real_t* tpwgts, real_t* ubvec,
			     idx_t* options, idx_t* edgecut, idx_t* part);
int METIS_PartGraphKway(idx_t* nvtxs, idx_t* ncon, idx_t* xadj, idx_t* adjncy, idx_t* vwgt,
			idx_t* vsize, idx_t* adjwgt, idx_t* nparts, real_t* tpwgts, real_t* ubvec,
			idx_t* options, idx_t* edgecut, idx_t* part);
#ifdef __cplusplus
}
#endif
#endif
