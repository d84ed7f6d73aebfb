// x-y halo exchange between subdomains: GPU pack -> NCCL send/recv group over NVLink -> GPU unpack.
// Replaces blk_macdrp_pack_mesg / MPI_Startall / MPI_Waitall / blk_macdrp_unpack_mesg
// (forward/blk_t.c:576-808, forward/drv_rk_curv_col.c:289, 308-312, 448-469).
#pragma once
#include <cuda_runtime.h>
#include "../../include/cgfd3d_b200.h"

struct HaloComm;
const char *halo_error();
int halo_unique_id(char id[128]);
HaloComm *halo_create(const char id[128], int rank, int nranks, const int neigh[4], const cgfd_grid_t &g, int ncmp, size_t V,
                      int pitch, cudaStream_t st);
void halo_destroy(HaloComm *h);
// refresh the ghosts of level w for an operator with direction indices (dirx, diry)
int halo_exchange(HaloComm *h, float *w, int dirx, int diry, cudaStream_t st);
int halo_launches_per_exchange(HaloComm *h);
// strips of one side (0..3 = x1,x2,y1,y2) as {i1, ni, j1, nj, k1, nk}; pure host logic
void halo_plan(const cgfd_grid_t &g, int dirx, int diry, int side, int send_box[6], int recv_box[6]);
