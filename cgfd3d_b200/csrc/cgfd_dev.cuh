// Device-side parameter blocks shared by the kernels and the host API of libcgfd3d_b200.so.
#pragma once
#include <cstddef>
#include <cstdint>
#include <cuda_runtime.h>

namespace cgfd {

// wavefield component order, forward/wav_t.c:93-136
enum { VX = 0, VY, VZ, TXX, TYY, TZZ, TYZ, TXZ, TXY, NCMP_EL = 9 };
// metric order, forward/gd_t.c:101-180
enum { M_JAC = 0, M_XIX, M_XIY, M_XIZ, M_ETX, M_ETY, M_ETZ, M_ZTX, M_ZTY, M_ZTZ, NMETRIC = 10 };

enum { KIND_FIRST = 0, KIND_MID = 1, KIND_LAST = 2 };

constexpr int MAX_MEDIA = 24;
constexpr int MAX_MAXWELL = 8;

// interior 5-point one-sided operators (forward/fd_t.c:89-113) and the near-surface 2-/3-point
// variants used by the stress RHS in the top rows (forward/fd_t.c:75-88); values come from the
// caller's fd tables (cgfd_fd_t), offsets are compile-time: dir 0 = {-1..3}, dir 1 = {-3..1}.
struct FdConst {
  float coef[2][5];
  float lay2[2][2];  // layer 1: dir 0 over {0,1}, dir 1 over {-1,0}
  float lay3[2][3];  // layer 2: dir 0 over {0,1,2}, dir 1 over {-2,-1,0}
};
extern __constant__ FdConst c_fd;

struct PmlFaceDev {
  int on;
  int i1, i2, j1, j2, k1, k2;   // slab range, inclusive (forward/bdry_t.c:161-186)
  int sni, snj;                 // slab extents in i and j
  size_t siz;                   // slab points per component
  const float *A, *B, *D;       // [nlay+1] device
  const float *aux_cur;         // level read by this stage
  const float *aux_pre;         // level n
  float *aux_tmp;               // level written for the next stage
  float *aux_end;               // accumulating level n+1
};

struct StageArgs {
  int nx, ny, nz;
  int ni1, ni2, nj1, nj2, nk1, nk2;
  int kbeg, kend;               // rows handled by this launch, inclusive
  int zchunk;                   // rows per block along z (main kernel)
  size_t siz_line, siz_slice, siz_vol;
  const float *cur;             // w_cur  [ncmp][nz][ny][nx]
  const float *pre;             // w_pre
  float *tmp;                   // w_tmp written for the next stage
  float *end;                   // w_end
  float a, b;                   // rk_a[s]*dt, rk_b[s]*dt (forward/drv_rk_curv_col.c:294-295)
  const float *metric[NMETRIC];
  const float *media[MAX_MEDIA];
  int nmaxwell;
  float wl[MAX_MAXWELL];
  PmlFaceDev pml[3][2];
  int free_top;
  int timg_mode;
  const float *matVx2Vz, *matVy2Vz, *matF2Vz, *matD;
  const float *TxSrc, *TySrc, *TzSrc, *VxSrc, *VySrc, *VzSrc;  // [ny][nx] or nullptr (== 0)
};

// launchers (kernels_*.cu); dir = direction index per axis of this stage's operator
void launch_iso_stage(const StageArgs &P, int dx, int dy, int dz, int kind, int variant, cudaStream_t st,
                      cudaEvent_t ev0, cudaEvent_t ev1, int *nlaunch);

}  // namespace cgfd
