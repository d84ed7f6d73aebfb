// Device-side parameter blocks shared by the kernels and the host API of libcgfd3d_b200.so.
#pragma once
#include <cstddef>
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>

namespace cgfd {

// wavefield component order, forward/wav_t.c:93-136
enum { VX = 0, VY, VZ, TXX, TYY, TZZ, TYZ, TXZ, TXY, NCMP_EL = 9 };
// metric order, forward/gd_t.c:101-180
enum { M_JAC = 0, M_XIX, M_XIY, M_XIZ, M_ETX, M_ETY, M_ETZ, M_ZTX, M_ZTY, M_ZTZ, NMETRIC = 10 };

// RK stage kinds. The wavefield update skips every w_end access that can be reconstructed (see rk_wave in physics.cuh):
//   FIRST (stage 0): tmp = cur + a h                                   loads cur            stores tmp
//   MID   (stage 1): tmp = pre + a h ; end = pre + c (cur-pre) + b h   loads cur, pre       stores tmp, end
//   THIRD (stage 2): tmp = pre + a h                                   loads cur, pre       stores tmp
//   LAST  (stage 3): end = end + c (cur-pre) + b h                     loads cur, pre, end  stores end
// with c = b_prev / a_prev, because cur = pre + a_prev h_prev. The PML auxiliary variables and the visco-elastic memory
// variables advance the same way.
enum { KIND_FIRST = 0, KIND_MID = 1, KIND_LAST = 2, KIND_THIRD = 3 };

// Device order of the metric arrays: the five that differ from zero on a grid whose x-y lines stay Cartesian (only the
// vertical coordinate is deformed: topography on a regular horizontal mesh) come first, so that the kernels specialised for
// such grids (GZ: xi_y = xi_z = eta_x = eta_z == 0 everywhere) fetch them with one 5-component TMA box.
//   block:  jac | xi_x eta_y zeta_x zeta_y zeta_z | xi_y xi_z eta_x eta_z
__host__ __device__ constexpr int metric_dev_slot(int m)
{
  return m == M_JAC ? 0 : m == M_XIX ? 1 : m == M_ETY ? 2 : m == M_ZTX ? 3 : m == M_ZTY ? 4 : m == M_ZTZ ? 5
       : m == M_XIY ? 6 : m == M_XIZ ? 7 : m == M_ETX ? 8 : 9;
}

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

// PML auxiliary variables live as one record of AUX_REC floats per slab point and level:
//   [Txx Tyy Tzz Tyz Txz Txy | Vx Vy Vz | pad]   (the stress part first: it is read first)
// so that a point's variables come with three 8-byte loads + one 8-byte and one 4-byte load instead of nine scalar
// loads from nine different cache lines. aux_slot(c) = position of wavefield component c inside the record.
constexpr int AUX_REC = 10;
__host__ __device__ constexpr int aux_slot(int c) { return c >= 3 ? c - 3 : 6 + c; }

struct PmlFaceDev {
  int on;
  int i1, i2, j1, j2, k1, k2;   // slab range, inclusive (forward/bdry_t.c:161-186)
  int sni, snj;                 // slab extents in i and j
  size_t siz;                   // slab points
  const float *A, *B, *D;       // [nlay+1] device
  const float *aux_cur;         // level read by this stage; every level is [slab point][AUX_REC]
  const float *aux_pre;         // level n
  float *aux_tmp;               // level written for the next stage
  float *aux_end;               // accumulating level n+1
};

// Device arrays keep the reference's [nz][ny][nx] order but with a padded x pitch: every row starts
// `shift` floats into a pitch that is a multiple of 32 floats, so that the first physical point
// (i = ni1) of every row sits on a 128-byte boundary. All 3-D base pointers handed to the kernels
// are already advanced by `shift`, i.e. they are indexed with the reference's own (i,j,k).
// 2-D per-surface-point arrays (matrices, source slices, PG maps) are unpadded [ny][nx].
struct StageArgs {
  int nx, ny, nz;               // logical extents incl. ghosts
  int shift;                    // x offset of index 0 inside a padded row
  int ni1, ni2, nj1, nj2, nk1, nk2;
  int kbeg, kend;               // rows handled by this launch, inclusive
  int zchunk;                   // rows per block along z (main kernel)
  int bx0, by0;                 // first tile of this launch along x / y (main kernel)
  int nbx, nby;                 // tiles of this launch along x / y: block b works on tile (b % nbx, b / nbx % nby), chunk b / (nbx nby)
  const int *order;             // optional permutation of the block indices: the slow (PML) tiles first (longest job first)
  int l2mode;                   // bit 0: wavefield tiles evict_last, bit 1: touch-once operands and results evict_first;
                                // bit 2: no PML-free fast path (every block-plane runs the PML copy of the loop body)
                                // bits 5-6: L2 prefetch distance of the interior kernel, planes beyond the ring (0 = off)
                                // bits 7, 8: traffic DIAGNOSTICS (results are wrong): no z-ahead loads / no priming of the zeta queue
                                //            (profiles/r2_experiments.txt, r2a: both kinds of loads hit in L2)
  size_t siz_line, siz_slice, siz_vol;   // padded pitch, pitch*ny, pitch*ny*nz
  const float *cur;             // w_cur  [ncmp][nz][ny][nx]
  const float *pre;             // w_pre
  float *tmp;                   // w_tmp written for the next stage
  float *end;                   // w_end
  float a, b;                   // rk_a[s]*dt, rk_b[s]*dt (forward/drv_rk_curv_col.c:294-295)
  float c;                      // rk_b[s-1]/rk_a[s-1]: weight of (cur - pre) in the w_end update of stages 1 and 3
  const float *metric[NMETRIC];
  const float *media[MAX_MEDIA];
  int nmaxwell;
  int vis_staged;               // visco-elastic interior kernel: memory variables and Y arrays staged by TMA (nmaxwell <= VIS_MAX_STAGED)
  float wl[MAX_MAXWELL];
  const float *qatt;            // Graves' attenuation factor exp(-pi f0 dt / Qs) per point, applied to w_end by the last stage; or nullptr
  PmlFaceDev pml[3][2];
  int free_top;
  int fuse_top;                 // the free-surface rows k >= nk2 - 3 are planes of the interior kernel's top z chunk (a launch of its
                                // TOPK instantiations; no k_top launch)
  int timg_mode;
  const float *matVx2Vz, *matVy2Vz, *matF2Vz, *matD;
  const float *TxSrc, *TySrc, *TzSrc, *VxSrc, *VySrc, *VzSrc;  // [ny][nx] or nullptr (== 0)
};

// tensor maps of one stage launch (TMA variant of the interior kernel)
struct TmaMaps {
  CUtensorMap cur;   // w_cur, box (TX+2*HALO_X, TY+4, 1, 9)
  CUtensorMap met;   // the 9 metric arrays in device order (metric_dev_slot), box (TX, TY, 1, 9)
  CUtensorMap met5;  // the same tensor, box (TX, TY, 1, 5): the arrays a vertically deformed grid needs (GZ kernels)
  CUtensorMap med;   // media, box (TX, TY, 1, nmedia)
  CUtensorMap pre;   // w_pre, box (TX, TY, 1, 9)
  CUtensorMap za;    // w_cur again, box (TX, TY, 1, 9): the centre box of the plane AHEAD of the march (media with ZA tiles)
  CUtensorMap end;   // w_end, box (TX, TY, 1, 9)
  // store maps: the PHYSICAL x-y range only (origin (ni1,nj1), extents ni x nj), so that the parts of a tile that hang
  // over the physical range are clipped by the TMA unit and ghosts are never written
  CUtensorMap out_tmp, out_end;
  // visco-elastic medium, staged variant: the 6N memory variables of the three levels read (box (TX, TY, 1, 6N) of a tensor
  // that starts at component 9 of the level), the 2N Ylam / Ymu arrays (tensor = media arrays 3 ..), and the store maps
  CUtensorMap jcur, jpre, jend, ymed, jout_tmp, jout_end;
};
constexpr int VIS_MAX_STAGED = 3;   // Maxwell bodies the staged variant has shared memory for (218 KB per block at 3)

// launchers, one explicit instantiation per medium (kernels_{iso,vti,aniso,vis}.cu); dir = direction index per axis of
// this stage's operator; MED = MED_* of physics.cuh
template <int MED> int med_kernels_init();   // one-time function attributes (dynamic shared memory)
template <int MED> int med_blocks_per_sm();  // resident blocks of the interior kernel per SM
// interior rows of the tile rectangle rect = {bx0, bx1, by0, by1} (tiles of TILE_X x TILE_Y points from (ni1, nj1))
// gz = 1: the grid has xi_y = xi_z = eta_x = eta_z == 0 (checked by the caller): kernels that never read those arrays
// rows [P.kbeg, P.kend]; topk = 1: the launch holds the free-surface rows k >= nk2 - 3 (kernels with the free-surface plane function)
template <int MED>
void med_launch_main(const StageArgs &P, const TmaMaps *maps, int dx, int dy, int dz, int kind, int gz, int topk, int zchunk, const int rect[4],
                     cudaStream_t st, cudaEvent_t ev0, cudaEvent_t ev1, int *nlaunch);
// the four free-surface rows, whole x-y range (no-op without a free top)
template <int MED> void med_launch_top(const StageArgs &P, int dx, int dy, int dz, int kind, cudaStream_t st, int *nlaunch);
// tile of the interior kernel (one thread per column); overridable at build time for tile-shape experiments
#ifndef CGFD_TILE_X
#define CGFD_TILE_X 32
#endif
#ifndef CGFD_TILE_Y
#define CGFD_TILE_Y 8
#endif
constexpr int TILE_X = CGFD_TILE_X, TILE_Y = CGFD_TILE_Y, HALO_X = 4;

}  // namespace cgfd
