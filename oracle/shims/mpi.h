/* Single-rank MPI stand-in used ONLY to compile the unmodified CGFD3D reference
 * sources into the parity oracle (oracle/_ref). Test infrastructure, not product.
 * Covers the 35 symbols the reference uses. */
#ifndef CGFD_ORACLE_MPI_SHIM_H
#define CGFD_ORACLE_MPI_SHIM_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif
typedef int MPI_Comm;
typedef int MPI_Request;
typedef int MPI_Datatype;
typedef int MPI_Op;
typedef struct { int MPI_SOURCE, MPI_TAG, MPI_ERROR; } MPI_Status;
#define MPI_COMM_WORLD 0
#define MPI_PROC_NULL (-1)
#define MPI_SUCCESS 0
#define MPI_CHAR   1
#define MPI_INT    4
#define MPI_LONG   8
#define MPI_FLOAT  14
#define MPI_REAL   14
#define MPI_SUM 1
#define MPI_MAX 2
#define MPI_MAX_PROCESSOR_NAME 256
#define MPI_STATUS_IGNORE ((MPI_Status *)0)
#define MPI_STATUSES_IGNORE ((MPI_Status *)0)
int MPI_Init(int *argc, char ***argv);
int MPI_Finalize(void);
int MPI_Abort(MPI_Comm comm, int code);
int MPI_Comm_rank(MPI_Comm comm, int *rank);
int MPI_Comm_size(MPI_Comm comm, int *size);
int MPI_Get_processor_name(char *name, int *len);
int MPI_Bcast(void *buf, int n, MPI_Datatype t, int root, MPI_Comm comm);
int MPI_Barrier(MPI_Comm comm);
int MPI_Allreduce(const void *s, void *r, int n, MPI_Datatype t, MPI_Op op, MPI_Comm comm);
int MPI_Allgather(const void *s, int ns, MPI_Datatype ts, void *r, int nr, MPI_Datatype tr, MPI_Comm comm);
int MPI_Cart_create(MPI_Comm comm, int ndims, const int *dims, const int *periods, int reorder, MPI_Comm *out);
int MPI_Cart_coords(MPI_Comm comm, int rank, int maxdims, int *coords);
int MPI_Cart_shift(MPI_Comm comm, int dir, int disp, int *src, int *dst);
int MPI_Send_init(const void *buf, int n, MPI_Datatype t, int dest, int tag, MPI_Comm comm, MPI_Request *req);
int MPI_Recv_init(void *buf, int n, MPI_Datatype t, int src, int tag, MPI_Comm comm, MPI_Request *req);
int MPI_Startall(int n, MPI_Request *reqs);
int MPI_Waitall(int n, MPI_Request *reqs, MPI_Status *st);
int MPI_Sendrecv(const void *s, int ns, MPI_Datatype ts, int dest, int stag,
                 void *r, int nr, MPI_Datatype tr, int src, int rtag, MPI_Comm comm, MPI_Status *st);
int MPI_Type_vector(int count, int blocklen, int stride, MPI_Datatype old, MPI_Datatype *newt);
int MPI_Type_commit(MPI_Datatype *t);
#ifdef __cplusplus
}
#endif
#endif
