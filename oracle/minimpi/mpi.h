/*
 * mpi.h -- minimpi: a tiny shared-memory stand-in for the MPI subset the reference's four MPI
 * programs use (the d2q9-bgk.c of MPI, MPI_Waitall and MPI_Testall_OptimizedVersion), so that they can be compiled UNMODIFIED from /root/reference and run
 * on the host cores of a box without an MPI installation.
 *
 * TEST / BASELINE INFRASTRUCTURE ONLY (lives under oracle/): never part of the product.
 *
 * Model: the program is started once; MPI_Init() forks MINIMPI_NP-1 children (env MINIMPI_NP, default 1)
 * after mapping one shared-memory region; parent = rank 0.  Point-to-point messages are EAGER: a send
 * copies its payload into a per-(source,destination) FIFO in shared memory and completes at once; a
 * receive is matched (same source, same tag, FIFO order) whenever the receiving rank makes progress,
 * i.e. inside any MPI call.  Receives that are never waited for (the reference's un-waited
 * MPI_Testall, MPI_Testall_OptimizedVersion/d2q9-bgk.c:279-280) stay posted and are completed by
 * a later call's progress, exactly like a real MPI library's progress engine would.
 */
#ifndef MINIMPI_MPI_H
#define MINIMPI_MPI_H

#ifdef __cplusplus
extern "C" {
#endif

typedef int MPI_Comm;
typedef int MPI_Datatype; /* low 24 bits: size in bytes; bits 24..: kind (0 derived, 1 float, 2 int) */
typedef int MPI_Op;
typedef struct minimpi_request* MPI_Request;
typedef struct { int MPI_SOURCE, MPI_TAG, MPI_ERROR; } MPI_Status;

#define MPI_COMM_WORLD 0
#define MPI_SUCCESS 0
#define MPI_FLOAT ((MPI_Datatype)((1 << 24) | 4))
#define MPI_INT ((MPI_Datatype)((2 << 24) | 4))
#define MPI_SUM 1
#define MPI_STATUS_IGNORE ((MPI_Status*)0)
#define MPI_STATUSES_IGNORE ((MPI_Status*)0)
#define MPI_REQUEST_NULL ((MPI_Request)0)

int MPI_Init(int* argc, char*** argv);
int MPI_Finalize(void);
int MPI_Comm_size(MPI_Comm comm, int* size);
int MPI_Comm_rank(MPI_Comm comm, int* rank);
int MPI_Type_contiguous(int count, MPI_Datatype oldtype, MPI_Datatype* newtype);
int MPI_Type_commit(MPI_Datatype* type);
int MPI_Send(const void* buf, int count, MPI_Datatype type, int dest, int tag, MPI_Comm comm);
int MPI_Recv(void* buf, int count, MPI_Datatype type, int source, int tag, MPI_Comm comm, MPI_Status* status);
int MPI_Sendrecv(const void* sendbuf, int sendcount, MPI_Datatype sendtype, int dest, int sendtag, void* recvbuf,
                 int recvcount, MPI_Datatype recvtype, int source, int recvtag, MPI_Comm comm, MPI_Status* status);
int MPI_Isend(const void* buf, int count, MPI_Datatype type, int dest, int tag, MPI_Comm comm, MPI_Request* request);
int MPI_Irecv(void* buf, int count, MPI_Datatype type, int source, int tag, MPI_Comm comm, MPI_Request* request);
int MPI_Wait(MPI_Request* request, MPI_Status* status);
int MPI_Waitall(int count, MPI_Request requests[], MPI_Status statuses[]);
int MPI_Testall(int count, MPI_Request requests[], int* flag, MPI_Status statuses[]);
int MPI_Reduce(const void* sendbuf, void* recvbuf, int count, MPI_Datatype type, MPI_Op op, int root, MPI_Comm comm);
int MPI_Barrier(MPI_Comm comm);
double MPI_Wtime(void);
int MPI_Abort(MPI_Comm comm, int errorcode);

#ifdef __cplusplus
}
#endif
#endif
