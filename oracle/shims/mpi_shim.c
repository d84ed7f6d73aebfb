/* MPI stand-in used ONLY to build and run the unmodified CGFD3D reference sources as the parity oracle / CPU baseline
 * (oracle/_ref). Test infrastructure, not product. Covers the 35 symbols the reference uses.
 *
 * Ranks: CGFD_SHIM_NPROCS in the environment (default 1). With N > 1, MPI_Init maps one shared-memory region and fork()s
 * N - 1 children, so `ref_main case.json` becomes an N-rank run on this host without any MPI installation:
 *   - point-to-point (persistent requests of the halo exchange, MPI_Sendrecv of the metric exchange): one mailbox per ordered
 *     rank pair in the shared region; Startall only marks requests active, Waitall / Sendrecv drive every pending request until
 *     all are complete (a sender copies into the mailbox when it is empty, a receiver copies out when tag and source match), so
 *     no ordering of posts can deadlock;
 *   - collectives (Barrier, Bcast, Allreduce, Allgather): a sense-reversing barrier and a per-rank scratch slot;
 *   - MPI_Type_vector: (count, blocklen, stride) records, packed / unpacked around the mailbox copy;
 *   - Cartesian topology: row-major ranks, no periodicity, no reordering (what the reference asks for, forward/mympi_t.c:32-44).
 * Rank 0 waits for the children in MPI_Finalize and turns a failed child into a non-zero exit status; MPI_Abort raises a
 * shared flag that every waiting rank polls.
 * With N = 1 every collective is a local copy and every neighbour is MPI_PROC_NULL, as before. */
#define _GNU_SOURCE
#include <sched.h>
#include <signal.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/wait.h>
#include <unistd.h>
#include "mpi.h"

#define MAXR 64
#define SCRATCH_BYTES (1 << 20)
#define TYPE_BASE 1000   /* derived datatypes are TYPE_BASE + index */
#define MAX_TYPES 64
#define MAX_REQS 4096

typedef struct {
  volatile int full;      /* 0: empty, 1: holds a message */
  int tag;
  size_t nbytes;
} box_hdr_t;

typedef struct {
  volatile int abort_code;
  volatile int bar_count;
  volatile int bar_sense;
  int nprocs;
  size_t box_cap;         /* payload bytes per mailbox */
} ctrl_t;

typedef struct { int count, blocklen, stride; MPI_Datatype old; } vtype_t;
typedef struct {
  int kind;               /* 0 unused, 1 send, 2 recv */
  void *buf; size_t nbytes; int peer, tag; MPI_Datatype type; int count;
  int active;
} req_t;

static int g_n = 1, g_rank = 0;
static ctrl_t *g_ctrl = NULL;
static unsigned char *g_boxes = NULL;      /* [src][dst]: box_hdr_t + payload */
static unsigned char *g_scratch = NULL;    /* [rank][SCRATCH_BYTES] */
static size_t g_box_stride = 0;
static pid_t g_children[MAXR];
static vtype_t g_types[MAX_TYPES];
static int g_ntypes = 0;
static req_t g_reqs[MAX_REQS];
static int g_nreqs = 1;                    /* request 0 = the null request */
static int g_dims[2] = {1, 1};
static int g_local_sense = 0;

static size_t tsize(MPI_Datatype t) { return t == MPI_CHAR ? 1 : (t == MPI_LONG ? 8 : 4); }
static size_t msg_bytes(int count, MPI_Datatype t)
{
  if (t >= TYPE_BASE) { vtype_t *v = &g_types[t - TYPE_BASE]; return (size_t)count * v->count * v->blocklen * tsize(v->old); }
  return (size_t)count * tsize(t);
}
static void pack(void *dst, const void *src, int count, MPI_Datatype t)
{
  if (t < TYPE_BASE) { memcpy(dst, src, msg_bytes(count, t)); return; }
  vtype_t *v = &g_types[t - TYPE_BASE];
  size_t es = tsize(v->old), bl = (size_t)v->blocklen * es;
  unsigned char *d = dst; const unsigned char *s = src;
  for (int c = 0; c < count * v->count; c++) { memcpy(d, s + (size_t)c * v->stride * es, bl); d += bl; }
}
static void unpack(void *dst, const void *src, int count, MPI_Datatype t)
{
  if (t < TYPE_BASE) { memcpy(dst, src, msg_bytes(count, t)); return; }
  vtype_t *v = &g_types[t - TYPE_BASE];
  size_t es = tsize(v->old), bl = (size_t)v->blocklen * es;
  unsigned char *d = dst; const unsigned char *s = src;
  for (int c = 0; c < count * v->count; c++) { memcpy(d + (size_t)c * v->stride * es, s, bl); s += bl; }
}

static void die_if_aborted(void)
{
  if (g_ctrl && g_ctrl->abort_code) { fflush(NULL); _exit(g_ctrl->abort_code); }
}
static pid_t g_parent = 0;
static void relax(unsigned *spins)
{
  die_if_aborted();
  if (++*spins > 200) sched_yield();
  if (*spins > 20000) usleep(50);
  if ((*spins & 4095) == 0) {
    /* a rank that died without MPI_Abort (crash, exit() inside the reference) must not leave the others spinning */
    if (g_rank == 0) {
      for (int r = 1; r < g_n; r++) {
        int st = 0;
        if (g_children[r] > 0 && waitpid(g_children[r], &st, WNOHANG) == g_children[r]) {
          g_children[r] = -1;
          if (!WIFEXITED(st) || WEXITSTATUS(st) != 0) { fprintf(stderr, "mpi shim: rank %d died\n", r); g_ctrl->abort_code = 5; }
        }
      }
    } else if (getppid() != g_parent) {
      _exit(6);
    }
  }
}
static box_hdr_t *box(int src, int dst) { return (box_hdr_t *)(g_boxes + ((size_t)src * g_n + dst) * g_box_stride); }

/* try to complete one request; 1 = done */
static int progress(req_t *q)
{
  if (!q->active) return 1;
  if (q->peer == MPI_PROC_NULL) { q->active = 0; return 1; }
  if (q->kind == 1) {
    box_hdr_t *b = box(g_rank, q->peer);
    if (b->full) return 0;
    if (q->nbytes > g_ctrl->box_cap) { fprintf(stderr, "mpi shim: message of %zu bytes exceeds CGFD_SHIM_MSG_MB\n", q->nbytes); MPI_Abort(0, 7); }
    pack((unsigned char *)(b + 1), q->buf, q->count, q->type);
    b->tag = q->tag; b->nbytes = q->nbytes;
    __sync_synchronize();
    b->full = 1;
    q->active = 0;
    return 1;
  }
  box_hdr_t *b = box(q->peer, g_rank);
  if (!b->full) return 0;
  __sync_synchronize();
  if (b->tag != q->tag) return 0;   /* another request of this rank owns that message */
  if (b->nbytes != q->nbytes) {
    fprintf(stderr, "mpi shim: rank %d expected %zu bytes from %d (tag %d), got %zu\n", g_rank, q->nbytes, q->peer, q->tag, b->nbytes);
    MPI_Abort(0, 8);
  }
  unpack(q->buf, (unsigned char *)(b + 1), q->count, q->type);
  __sync_synchronize();
  b->full = 0;
  q->active = 0;
  return 1;
}
static void drive(req_t **qs, int n)
{
  unsigned spins = 0;
  for (;;) {
    int left = 0;
    for (int i = 0; i < n; i++) if (qs[i] && !progress(qs[i])) left++;
    if (!left) return;
    relax(&spins);
  }
}

static void barrier(void)
{
  if (g_n == 1) return;
  g_local_sense = !g_local_sense;
  if (__sync_add_and_fetch(&g_ctrl->bar_count, 1) == g_n) {
    g_ctrl->bar_count = 0;
    __sync_synchronize();
    g_ctrl->bar_sense = g_local_sense;
  } else {
    unsigned spins = 0;
    while (g_ctrl->bar_sense != g_local_sense) relax(&spins);
  }
  __sync_synchronize();
}

int MPI_Init(int *argc, char ***argv)
{
  (void)argc; (void)argv;
  const char *e = getenv("CGFD_SHIM_NPROCS");
  g_n = e ? atoi(e) : 1;
  if (g_n < 1) g_n = 1;
  if (g_n > MAXR) { fprintf(stderr, "mpi shim: at most %d ranks\n", MAXR); exit(3); }
  if (g_n == 1) return 0;
  size_t cap = (size_t)(getenv("CGFD_SHIM_MSG_MB") ? atoi(getenv("CGFD_SHIM_MSG_MB")) : 32) << 20;
  g_box_stride = sizeof(box_hdr_t) + cap;
  g_box_stride = (g_box_stride + 63) & ~(size_t)63;
  size_t total = 4096 + (size_t)g_n * g_n * g_box_stride + (size_t)g_n * SCRATCH_BYTES;
  unsigned char *m = mmap(NULL, total, PROT_READ | PROT_WRITE, MAP_SHARED | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
  if (m == MAP_FAILED) { perror("mpi shim: mmap"); exit(3); }
  g_ctrl = (ctrl_t *)m; g_boxes = m + 4096; g_scratch = g_boxes + (size_t)g_n * g_n * g_box_stride;
  g_ctrl->abort_code = 0; g_ctrl->bar_count = 0; g_ctrl->bar_sense = 0; g_ctrl->nprocs = g_n; g_ctrl->box_cap = cap;
  fflush(NULL);
  g_parent = getpid();
  for (int r = 1; r < g_n; r++) {
    pid_t p = fork();
    if (p < 0) { perror("mpi shim: fork"); exit(3); }
    if (p == 0) { g_rank = r; return 0; }
    g_children[r] = p;
  }
  g_rank = 0;
  return 0;
}
int MPI_Finalize(void)
{
  if (g_n == 1) return 0;
  barrier();
  fflush(NULL);
  if (g_rank != 0) _exit(0);
  int bad = 0;
  for (int r = 1; r < g_n; r++) {
    int st = 0;
    if (g_children[r] <= 0) continue;   /* already reaped by relax(): it had exited cleanly */
    if (waitpid(g_children[r], &st, 0) < 0 || !WIFEXITED(st) || WEXITSTATUS(st) != 0) bad = 1;
  }
  if (bad) { fprintf(stderr, "mpi shim: a rank failed\n"); exit(4); }
  return 0;
}
int MPI_Abort(MPI_Comm comm, int code)
{
  (void)comm;
  fflush(NULL);
  if (g_ctrl) g_ctrl->abort_code = code ? code : 1;
  if (g_n > 1 && g_rank == 0) {   /* give the children a moment to see the flag, then make sure they are gone */
    usleep(200000);
    for (int r = 1; r < g_n; r++) if (g_children[r] > 0) kill(g_children[r], SIGKILL);
  }
  _exit(code ? code : 1);
}
int MPI_Comm_rank(MPI_Comm comm, int *rank) { (void)comm; *rank = g_rank; return 0; }
int MPI_Comm_size(MPI_Comm comm, int *size) { (void)comm; *size = g_n; return 0; }
int MPI_Get_processor_name(char *name, int *len) { strcpy(name, "localhost"); *len = 9; return 0; }
int MPI_Barrier(MPI_Comm c) { (void)c; barrier(); return 0; }

int MPI_Bcast(void *b, int n, MPI_Datatype t, int root, MPI_Comm c)
{
  (void)c;
  if (g_n == 1) return 0;
  size_t left = (size_t)n * tsize(t); unsigned char *p = b;
  while (left) {   /* through the root's scratch slot, one chunk at a time */
    size_t k = left < SCRATCH_BYTES ? left : SCRATCH_BYTES;
    if (g_rank == root) memcpy(g_scratch + (size_t)root * SCRATCH_BYTES, p, k);
    barrier();
    if (g_rank != root) memcpy(p, g_scratch + (size_t)root * SCRATCH_BYTES, k);
    barrier();
    p += k; left -= k;
  }
  return 0;
}
int MPI_Allgather(const void *s, int ns, MPI_Datatype ts, void *r, int nr, MPI_Datatype tr, MPI_Comm c)
{
  (void)nr; (void)tr; (void)c;
  size_t k = (size_t)ns * tsize(ts);
  if (g_n == 1) { memmove(r, s, k); return 0; }
  if (k > SCRATCH_BYTES) { fprintf(stderr, "mpi shim: Allgather chunk too large\n"); MPI_Abort(0, 9); }
  memcpy(g_scratch + (size_t)g_rank * SCRATCH_BYTES, s, k);
  barrier();
  for (int q = 0; q < g_n; q++) memcpy((unsigned char *)r + (size_t)q * k, g_scratch + (size_t)q * SCRATCH_BYTES, k);
  barrier();
  return 0;
}
int MPI_Allreduce(const void *s, void *r, int n, MPI_Datatype t, MPI_Op op, MPI_Comm c)
{
  (void)c;
  size_t k = (size_t)n * tsize(t);
  if (g_n == 1) { memmove(r, s, k); return 0; }
  if (k > SCRATCH_BYTES) { fprintf(stderr, "mpi shim: Allreduce chunk too large\n"); MPI_Abort(0, 9); }
  memcpy(g_scratch + (size_t)g_rank * SCRATCH_BYTES, s, k);
  barrier();
  for (int i = 0; i < n; i++) {   /* every rank reduces in rank order: identical results everywhere */
    if (t == MPI_FLOAT) {
      float acc = ((float *)(g_scratch))[i];
      for (int q = 1; q < g_n; q++) { float v = ((float *)(g_scratch + (size_t)q * SCRATCH_BYTES))[i]; acc = (op == MPI_MAX) ? (v > acc ? v : acc) : acc + v; }
      ((float *)r)[i] = acc;
    } else if (t == MPI_INT) {
      int acc = ((int *)(g_scratch))[i];
      for (int q = 1; q < g_n; q++) { int v = ((int *)(g_scratch + (size_t)q * SCRATCH_BYTES))[i]; acc = (op == MPI_MAX) ? (v > acc ? v : acc) : acc + v; }
      ((int *)r)[i] = acc;
    } else { fprintf(stderr, "mpi shim: Allreduce datatype %d not supported\n", t); MPI_Abort(0, 9); }
  }
  barrier();
  return 0;
}

int MPI_Cart_create(MPI_Comm comm, int ndims, const int *dims, const int *periods, int reorder, MPI_Comm *out)
{
  (void)comm; (void)periods; (void)reorder;
  int tot = 1;
  for (int i = 0; i < ndims && i < 2; i++) { g_dims[i] = dims[i]; tot *= dims[i]; }
  if (tot != g_n) {
    fprintf(stderr, "mpi shim: process grid %dx%d needs %d ranks, CGFD_SHIM_NPROCS gives %d\n", g_dims[0], g_dims[1], tot, g_n);
    MPI_Abort(0, 3);
  }
  *out = 1; return 0;
}
int MPI_Cart_coords(MPI_Comm c, int rank, int maxdims, int *coords)
{
  (void)c;
  for (int i = 0; i < maxdims; i++) coords[i] = 0;
  coords[0] = rank / g_dims[1];
  if (maxdims > 1) coords[1] = rank % g_dims[1];
  return 0;
}
int MPI_Cart_shift(MPI_Comm c, int dir, int disp, int *src, int *dst)
{
  (void)c;
  int co[2] = { g_rank / g_dims[1], g_rank % g_dims[1] };
  int lo[2] = { co[0], co[1] }, hi[2] = { co[0], co[1] };
  lo[dir] -= disp; hi[dir] += disp;
  *src = (lo[dir] < 0 || lo[dir] >= g_dims[dir]) ? MPI_PROC_NULL : lo[0] * g_dims[1] + lo[1];
  *dst = (hi[dir] < 0 || hi[dir] >= g_dims[dir]) ? MPI_PROC_NULL : hi[0] * g_dims[1] + hi[1];
  return 0;
}

static int new_req(int kind, void *buf, int n, MPI_Datatype t, int peer, int tag, MPI_Request *q)
{
  if (g_nreqs >= MAX_REQS) { fprintf(stderr, "mpi shim: too many persistent requests\n"); MPI_Abort(0, 9); }
  req_t *r = &g_reqs[g_nreqs];
  r->kind = kind; r->buf = buf; r->count = n; r->type = t; r->nbytes = msg_bytes(n, t); r->peer = peer; r->tag = tag; r->active = 0;
  *q = g_nreqs++;
  return 0;
}
int MPI_Send_init(const void *b, int n, MPI_Datatype t, int d, int tag, MPI_Comm c, MPI_Request *q)
{ (void)c; return new_req(1, (void *)b, n, t, d, tag, q); }
int MPI_Recv_init(void *b, int n, MPI_Datatype t, int s, int tag, MPI_Comm c, MPI_Request *q)
{ (void)c; return new_req(2, b, n, t, s, tag, q); }
int MPI_Startall(int n, MPI_Request *q)
{
  for (int i = 0; i < n; i++) if (q[i] > 0) g_reqs[q[i]].active = 1;
  return 0;
}
int MPI_Waitall(int n, MPI_Request *q, MPI_Status *s)
{
  (void)s;
  if (g_n == 1) { for (int i = 0; i < n; i++) if (q[i] > 0) g_reqs[q[i]].active = 0; return 0; }
  req_t *qs[64];
  if (n > 64) { fprintf(stderr, "mpi shim: Waitall of %d requests\n", n); MPI_Abort(0, 9); }
  for (int i = 0; i < n; i++) qs[i] = q[i] > 0 ? &g_reqs[q[i]] : NULL;
  drive(qs, n);
  return 0;
}
int MPI_Sendrecv(const void *s, int ns, MPI_Datatype ts, int dest, int stag,
                 void *r, int nr, MPI_Datatype tr, int src, int rtag, MPI_Comm c, MPI_Status *st)
{
  (void)c; (void)st;
  if (g_n == 1) return 0;
  req_t a = { 1, (void *)s, msg_bytes(ns, ts), dest, stag, ts, ns, 1 };
  req_t b = { 2, r, msg_bytes(nr, tr), src, rtag, tr, nr, 1 };
  req_t *qs[2] = { &a, &b };
  drive(qs, 2);
  return 0;
}
int MPI_Type_vector(int count, int bl, int stride, MPI_Datatype old, MPI_Datatype *newt)
{
  if (g_ntypes >= MAX_TYPES) g_ntypes = 0;   /* the reference creates two per metric exchange and never frees them */
  g_types[g_ntypes].count = count; g_types[g_ntypes].blocklen = bl; g_types[g_ntypes].stride = stride; g_types[g_ntypes].old = old;
  *newt = TYPE_BASE + g_ntypes++;
  return 0;
}
int MPI_Type_commit(MPI_Datatype *t) { (void)t; return 0; }
