#include "halo.h"

#include <dlfcn.h>
#include <cstdio>
#include <cstring>
#include <string>

#include "aux_kernels.cuh"

// The few NCCL entry points used, resolved at run time so that single-GPU use of the library has no
// NCCL dependency (the torch-bundled libnccl.so.2 is already loaded in a torch process).
typedef struct ncclComm *ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
enum { ncclFloat = 7 };
struct NcclApi {
  void *lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*Send)(const void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
};
static NcclApi g_nccl;
static thread_local std::string g_herr;
const char *halo_error() { return g_herr.c_str(); }

static int load_nccl()
{
  if (g_nccl.lib) return 0;
  const char *names[] = {"libnccl.so.2", "libnccl.so", nullptr};
  void *lib = nullptr;
  for (int n = 0; names[n] && !lib; n++) lib = dlopen(names[n], RTLD_NOW | RTLD_GLOBAL);
  if (!lib) { g_herr = std::string("cannot load NCCL: ") + dlerror(); return 1; }
#define SYM(field, name) \
  *(void **)(&g_nccl.field) = dlsym(lib, name); \
  if (!g_nccl.field) { g_herr = std::string("NCCL symbol missing: ") + name; return 1; }
  SYM(GetUniqueId, "ncclGetUniqueId") SYM(CommInitRank, "ncclCommInitRank") SYM(CommDestroy, "ncclCommDestroy")
  SYM(Send, "ncclSend") SYM(Recv, "ncclRecv") SYM(GroupStart, "ncclGroupStart") SYM(GroupEnd, "ncclGroupEnd")
  SYM(GetErrorString, "ncclGetErrorString")
#undef SYM
  g_nccl.lib = lib;
  return 0;
}
#define NC(call) do { ncclResult_t r_ = (call); if (r_ != 0) { g_herr = std::string(#call) + ": " + g_nccl.GetErrorString(r_); return 1; } } while (0)

struct HaloComm {
  ncclComm_t comm = nullptr;
  int rank = 0, nranks = 1;
  int neigh[4];
  cgfd_grid_t g;
  int ncmp = 9;
  size_t V = 0;
  int pitch = 0;   // padded x pitch of the device arrays
  float *sbuf[4] = {nullptr, nullptr, nullptr, nullptr};
  float *rbuf[4] = {nullptr, nullptr, nullptr, nullptr};
  size_t plane[4];    // floats per exchanged plane per side
};

int halo_unique_id(char id[128])
{
  if (load_nccl()) return 1;
  ncclUniqueId u;
  NC(g_nccl.GetUniqueId(&u));
  memcpy(id, u.internal, 128);
  return 0;
}

HaloComm *halo_create(const char id[128], int rank, int nranks, const int neigh[4], const cgfd_grid_t &g, int ncmp, size_t V,
                      int pitch, cudaStream_t st)
{
  (void)st;
  if (load_nccl()) return nullptr;
  HaloComm *h = new HaloComm();
  h->rank = rank; h->nranks = nranks; h->g = g; h->ncmp = ncmp; h->V = V; h->pitch = pitch;
  for (int n = 0; n < 4; n++) h->neigh[n] = neigh[n];
  ncclUniqueId u;
  memcpy(u.internal, id, 128);
  ncclResult_t r = g_nccl.CommInitRank(&h->comm, nranks, u, rank);
  if (r != 0) { g_herr = std::string("ncclCommInitRank: ") + g_nccl.GetErrorString(r); delete h; return nullptr; }
  const size_t nj = g.nj2 - g.nj1 + 1, nk = g.nk2 - g.nk1 + 1, ni = g.ni2 - g.ni1 + 1;
  for (int n = 0; n < 4; n++) {
    h->plane[n] = (n < 2 ? nj * nk : ni * nk) * (size_t)ncmp;
    if (neigh[n] < 0) continue;
    if (cudaMalloc((void **)&h->sbuf[n], h->plane[n] * 3 * sizeof(float)) != cudaSuccess ||
        cudaMalloc((void **)&h->rbuf[n], h->plane[n] * 3 * sizeof(float)) != cudaSuccess) {
      g_herr = "halo buffers: cudaMalloc failed"; delete h; return nullptr;
    }
  }
  return h;
}

void halo_destroy(HaloComm *h)
{
  if (!h) return;
  for (int n = 0; n < 4; n++) { if (h->sbuf[n]) cudaFree(h->sbuf[n]); if (h->rbuf[n]) cudaFree(h->rbuf[n]); }
  if (h->comm) g_nccl.CommDestroy(h->comm);
  delete h;
}

int halo_launches_per_exchange(HaloComm *h)
{
  int n = 0;
  for (int s = 0; s < 4; s++) if (h->neigh[s] >= 0) n += 2;
  return n;
}

// Which strips travel for an operator with direction indices (dirx, diry): operator widths are
// dir 0 -> left 1 / right 3, dir 1 -> left 3 / right 1 (forward/fd_t.c:133-147); the message to x1 carries my
// first `right` physical planes, the message to x2 my last `left` ones, and the ghosts filled from x1 / x2 are
// `left` / `right` wide (forward/blk_t.c:497-510, 594-679, 700-808); y likewise. No edges or corners.
// box = {i1, ni, j1, nj, k1, nk} in local indices including ghosts.
void halo_plan(const cgfd_grid_t &g, int dirx, int diry, int side, int send_box[6], int recv_box[6])
{
  const int ni = g.ni2 - g.ni1 + 1, nj = g.nj2 - g.nj1 + 1, nk = g.nk2 - g.nk1 + 1;
  const int lx = dirx ? 3 : 1, rx = dirx ? 1 : 3, ly = diry ? 3 : 1, ry = diry ? 1 : 3;
  const int send_w[4] = {rx, lx, ry, ly};
  const int recv_w[4] = {lx, rx, ly, ry};
  const int send_i1[4] = {g.ni1, g.ni2 - lx + 1, g.ni1, g.ni1};
  const int send_j1[4] = {g.nj1, g.nj1, g.nj1, g.nj2 - ly + 1};
  const int recv_i1[4] = {g.ni1 - lx, g.ni2 + 1, g.ni1, g.ni1};
  const int recv_j1[4] = {g.nj1, g.nj1, g.nj1 - ly, g.nj2 + 1};
  const int s = side;
  send_box[0] = send_i1[s]; send_box[1] = (s < 2) ? send_w[s] : ni;
  send_box[2] = send_j1[s]; send_box[3] = (s < 2) ? nj : send_w[s];
  send_box[4] = g.nk1; send_box[5] = nk;
  recv_box[0] = recv_i1[s]; recv_box[1] = (s < 2) ? recv_w[s] : ni;
  recv_box[2] = recv_j1[s]; recv_box[3] = (s < 2) ? nj : recv_w[s];
  recv_box[4] = g.nk1; recv_box[5] = nk;
}

int halo_exchange(HaloComm *h, float *w, int dirx, int diry, cudaStream_t st)
{
  using cgfd::k_halo_copy;
  const cgfd_grid_t &g = h->g;
  int sb[4][6], rb[4][6];
  size_t cnt_s[4], cnt_r[4];
  for (int s = 0; s < 4; s++) {
    if (h->neigh[s] < 0) continue;
    halo_plan(g, dirx, diry, s, sb[s], rb[s]);
    cnt_s[s] = (size_t)sb[s][1] * sb[s][3] * sb[s][5] * h->ncmp;
    cnt_r[s] = (size_t)rb[s][1] * rb[s][3] * rb[s][5] * h->ncmp;
    k_halo_copy<<<(unsigned)((cnt_s[s] + 255) / 256), 256, 0, st>>>(w, h->sbuf[s], h->V, h->ncmp, h->pitch, g.ny, sb[s][0], sb[s][1],
                                                                   sb[s][2], sb[s][3], sb[s][4], sb[s][5], 0);
  }
  NC(g_nccl.GroupStart());
  for (int s = 0; s < 4; s++) {
    if (h->neigh[s] < 0) continue;
    NC(g_nccl.Recv(h->rbuf[s], cnt_r[s], ncclFloat, h->neigh[s], h->comm, st));
    NC(g_nccl.Send(h->sbuf[s], cnt_s[s], ncclFloat, h->neigh[s], h->comm, st));
  }
  NC(g_nccl.GroupEnd());
  for (int s = 0; s < 4; s++) {
    if (h->neigh[s] < 0) continue;
    k_halo_copy<<<(unsigned)((cnt_r[s] + 255) / 256), 256, 0, st>>>(w, h->rbuf[s], h->V, h->ncmp, h->pitch, g.ny, rb[s][0], rb[s][1],
                                                                   rb[s][2], rb[s][3], rb[s][4], rb[s][5], 1);
  }
  if (cudaGetLastError() != cudaSuccess) { g_herr = "halo kernels failed to launch"; return 1; }
  return 0;
}
