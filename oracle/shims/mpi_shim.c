/* Single-rank MPI stand-in (see mpi.h). Every collective is a local copy; every
 * neighbour is MPI_PROC_NULL; persistent requests are no-ops. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "mpi.h"

static size_t tsize(MPI_Datatype t) { return t == MPI_CHAR ? 1 : (t == MPI_LONG ? 8 : 4); }

int MPI_Init(int *argc, char ***argv) { (void)argc; (void)argv; return 0; }
int MPI_Finalize(void) { return 0; }
int MPI_Abort(MPI_Comm comm, int code) { (void)comm; fflush(NULL); exit(code ? code : 1); }
int MPI_Comm_rank(MPI_Comm comm, int *rank) { (void)comm; *rank = 0; return 0; }
int MPI_Comm_size(MPI_Comm comm, int *size) { (void)comm; *size = 1; return 0; }
int MPI_Get_processor_name(char *name, int *len) { strcpy(name, "localhost"); *len = 9; return 0; }
int MPI_Bcast(void *b, int n, MPI_Datatype t, int root, MPI_Comm c) { (void)b; (void)n; (void)t; (void)root; (void)c; return 0; }
int MPI_Barrier(MPI_Comm c) { (void)c; return 0; }
int MPI_Allreduce(const void *s, void *r, int n, MPI_Datatype t, MPI_Op op, MPI_Comm c)
{ (void)op; (void)c; memmove(r, s, (size_t)n * tsize(t)); return 0; }
int MPI_Allgather(const void *s, int ns, MPI_Datatype ts, void *r, int nr, MPI_Datatype tr, MPI_Comm c)
{ (void)nr; (void)tr; (void)c; memmove(r, s, (size_t)ns * tsize(ts)); return 0; }
int MPI_Cart_create(MPI_Comm comm, int ndims, const int *dims, const int *periods, int reorder, MPI_Comm *out)
{
  (void)comm; (void)periods; (void)reorder;
  for (int i = 0; i < ndims; i++) if (dims[i] != 1) {
    fprintf(stderr, "mpi shim: single-rank only, got dims[%d]=%d\n", i, dims[i]); exit(3);
  }
  *out = 1; return 0;
}
int MPI_Cart_coords(MPI_Comm c, int rank, int maxdims, int *coords)
{ (void)c; (void)rank; for (int i = 0; i < maxdims; i++) coords[i] = 0; return 0; }
int MPI_Cart_shift(MPI_Comm c, int dir, int disp, int *src, int *dst)
{ (void)c; (void)dir; (void)disp; *src = MPI_PROC_NULL; *dst = MPI_PROC_NULL; return 0; }
int MPI_Send_init(const void *b, int n, MPI_Datatype t, int d, int tag, MPI_Comm c, MPI_Request *q)
{ (void)b; (void)n; (void)t; (void)d; (void)tag; (void)c; *q = 0; return 0; }
int MPI_Recv_init(void *b, int n, MPI_Datatype t, int s, int tag, MPI_Comm c, MPI_Request *q)
{ (void)b; (void)n; (void)t; (void)s; (void)tag; (void)c; *q = 0; return 0; }
int MPI_Startall(int n, MPI_Request *q) { (void)n; (void)q; return 0; }
int MPI_Waitall(int n, MPI_Request *q, MPI_Status *s) { (void)n; (void)q; (void)s; return 0; }
int MPI_Sendrecv(const void *s, int ns, MPI_Datatype ts, int dest, int stag,
                 void *r, int nr, MPI_Datatype tr, int src, int rtag, MPI_Comm c, MPI_Status *st)
{ (void)s; (void)ns; (void)ts; (void)dest; (void)stag; (void)r; (void)nr; (void)tr; (void)src; (void)rtag; (void)c; (void)st; return 0; }
int MPI_Type_vector(int count, int bl, int stride, MPI_Datatype old, MPI_Datatype *newt)
{ (void)count; (void)bl; (void)stride; *newt = old; return 0; }
int MPI_Type_commit(MPI_Datatype *t) { (void)t; return 0; }
