/*
 * mpi.h -- minimal process-based MPI shim (TEST INFRASTRUCTURE ONLY).
 *
 * The reference (ngcurrier/ProteusCFD) is MPI-only and this image has no MPI.
 * This header + mpi_shim.cpp provide the ~20 MPI entry points the reference
 * uses (grep over ucs/: Isend, Irecv, Wait, Allreduce, Reduce, Bcast,
 * Allgather, Alltoall, Barrier, Abort, Comm_rank/size, Init, Finalize) so the
 * UNMODIFIED reference sources compile and run as the parity oracle and the
 * CPU baseline.  Ranks are forked from MPI_Init (count = $PCFD_MPI_NP,
 * default 1) and talk over AF_UNIX socket pairs.  Nothing under
 * proteuscfd_b200/ may include or link this.
 */
#ifndef PCFD_MPI_SHIM_H
#define PCFD_MPI_SHIM_H

#ifdef __cplusplus
extern "C" {
#endif

typedef int MPI_Comm;
typedef int MPI_Datatype;
typedef int MPI_Op;
typedef int MPI_Request;
typedef struct { int MPI_SOURCE, MPI_TAG, MPI_ERROR; } MPI_Status;

#define MPI_COMM_WORLD 0
#define MPI_SUCCESS 0
#define MPI_IN_PLACE ((void*)1)
#define MPI_STATUS_IGNORE ((MPI_Status*)0)
#define MPI_REQUEST_NULL (-1)

/* datatypes (ids; sizes are in mpi_shim.cpp) */
#define MPI_BYTE 1
#define MPI_CHAR 2
#define MPI_INT 3
#define MPI_FLOAT 4
#define MPI_DOUBLE 5
#define MPI_COMPLEX 6
#define MPI_DOUBLE_COMPLEX 7
#define MPI_DOUBLE_INT 8
#define MPI_UNSIGNED 9
#define MPI_LONG 10

#define MPI_SUM 1
#define MPI_MAX 2
#define MPI_MIN 3
#define MPI_MINLOC 4
#define MPI_MAXLOC 5

int MPI_Init(int* argc, char*** argv);
int MPI_Finalize(void);
int MPI_Abort(MPI_Comm comm, int code);
int MPI_Comm_rank(MPI_Comm comm, int* rank);
int MPI_Comm_size(MPI_Comm comm, int* size);
int MPI_Barrier(MPI_Comm comm);
int MPI_Isend(const void* buf, int count, MPI_Datatype t, int dest, int tag, MPI_Comm comm, MPI_Request* req);
int MPI_Irecv(void* buf, int count, MPI_Datatype t, int src, int tag, MPI_Comm comm, MPI_Request* req);
int MPI_Wait(MPI_Request* req, MPI_Status* status);
int MPI_Send(const void* buf, int count, MPI_Datatype t, int dest, int tag, MPI_Comm comm);
int MPI_Recv(void* buf, int count, MPI_Datatype t, int src, int tag, MPI_Comm comm, MPI_Status* status);
int MPI_Bcast(void* buf, int count, MPI_Datatype t, int root, MPI_Comm comm);
int MPI_Reduce(const void* sendbuf, void* recvbuf, int count, MPI_Datatype t, MPI_Op op, int root, MPI_Comm comm);
int MPI_Allreduce(const void* sendbuf, void* recvbuf, int count, MPI_Datatype t, MPI_Op op, MPI_Comm comm);
int MPI_Allgather(const void* sendbuf, int scount, MPI_Datatype st, void* recvbuf, int rcount, MPI_Datatype rt, MPI_Comm comm);
int MPI_Alltoall(const void* sendbuf, int scount, MPI_Datatype st, void* recvbuf, int rcount, MPI_Datatype rt, MPI_Comm comm);
double MPI_Wtime(void);

#ifdef __cplusplus
}
#endif
#endif
