/*
 * cgfd3d_b200.h -- C ABI of the B200-native CGFD3D time-stepping hot path.
 *
 * Plain C types only (pointers, sizes, PODs). This is the boundary a CGFD3D maintainer binds
 * from the C host program: the reference's own call site is the single call
 *     drv_rk_curv_col_allstep(...)                    forward/main_curv_col_el_3d.c:849-858
 * declared at forward/drv_rk_curv_col.h:17-37. The replacement driver (integration/
 * drv_rk_curv_col_b200.c) flattens the reference structs into
 * cgfd_problem_t and calls the entry points below.
 *
 * Every array is float32, x fastest: iptr = i + j*nx + k*nx*ny, 3 ghost layers per side
 * (forward/wav_t.c:35-38, forward/gd_t.c:2890-2897). Host pointers unless stated otherwise.
 *
 * All functions return 0 on success, non-zero on error; cgfd_b200_last_error() gives the text.
 * There is NO CPU fallback: without a CUDA device every compute entry point fails.
 */
#ifndef CGFD3D_B200_H
#define CGFD3D_B200_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CGFD_ABI_VERSION 2

/* medium types: values of forward/constants.h:21-26 */
#define CGFD_MEDIUM_ELASTIC_ISO      2
#define CGFD_MEDIUM_ELASTIC_VTI      3
#define CGFD_MEDIUM_ELASTIC_ANISO    4
#define CGFD_MEDIUM_VISCOELASTIC_ISO 5

/* spatial source extent: forward/constants.h:32-35 */
#define CGFD_SRC_SPATIAL_POINT    1
#define CGFD_SRC_SPATIAL_GAUSSIAN 2

/* What the traction-image operator does with the image term whose mirror index falls outside
 * its 5-point window (the reference reads out of bounds there, forward/sv_curv_col_el.c:162,
 * 243,290; SURVEY.md §8c hazard 1). ZERO = treat the missing term as 0 (parity default, equals
 * the oracle build ref_main_zero); MIRROR = take it from the grid row it stands for
 * (the in-source commented alternative, forward/sv_curv_col_el.c:154-159). */
#define CGFD_TIMG_ZERO   0
#define CGFD_TIMG_MIRROR 1

#define CGFD_MAX_MEDIA   24   /* aniso: 21 Cij + rho = 22 */
#define CGFD_MAX_MAXWELL 8
#define CGFD_NUM_PAIRS   8
#define CGFD_NUM_STAGES  4

/* wavefield component order (forward/wav_t.c:93-136) */
enum { CGFD_VX = 0, CGFD_VY, CGFD_VZ, CGFD_TXX, CGFD_TYY, CGFD_TZZ, CGFD_TYZ, CGFD_TXZ, CGFD_TXY };

/* metric array order (forward/gd_t.c:101-180) */
enum { CGFD_JAC = 0, CGFD_XI_X, CGFD_XI_Y, CGFD_XI_Z, CGFD_ET_X, CGFD_ET_Y, CGFD_ET_Z,
       CGFD_ZT_X, CGFD_ZT_Y, CGFD_ZT_Z, CGFD_NUM_METRIC };

/* ---- geometry of one subdomain: gd_t index ranges (forward/gd_t.h:27-125) ------------------ */
typedef struct {
  int32_t nx, ny, nz;                     /* including ghosts */
  int32_t ni1, ni2, nj1, nj2, nk1, nk2;   /* physical range, inclusive */
} cgfd_grid_t;

/* ---- fd_t tables the hot path uses (forward/fd_t.c:27-49, 75-114, 182-289) ----------------- */
typedef struct {
  float   rk_a[CGFD_NUM_STAGES];          /* a = {.5,.5,1,-}  */
  float   rk_b[CGFD_NUM_STAGES];          /* b = {1/6,1/3,1/3,1/6} */
  /* direction index per [pair][stage][axis]: 0 -> op over offsets {-1..3}, 1 -> {-3..1} */
  int32_t dir[CGFD_NUM_PAIRS][CGFD_NUM_STAGES][3];
  /* interior 5-point op for direction index 0/1 (pair_fd?_op[..][..][last]) */
  int32_t indx[2][5];
  float   coef[2][5];
  /* near-surface Dz for the stress RHS (vlow): layer n = 1 (2-pt), n = 2 (3-pt), per direction
   * (pair_fdz_op[..][..][n], forward/sv_curv_col_el_iso.c:503-516). Layer 0 is unused. */
  int32_t lay_len[3][2];
  int32_t lay_indx[3][2][5];
  float   lay_coef[3][2][5];
} cgfd_fd_t;

/* ---- CFS-PML of one face (forward/bdry_t.c:115-288): transformed coefficient profiles ------- */
typedef struct {
  int32_t enabled;                        /* bdry->is_sides_pml[idim][iside] */
  int32_t nlay;                           /* number_of_layers; the slab holds nlay+1 points */
  const float *A, *B, *D;                 /* [nlay+1], A = alpha + d/beta, B = 1/beta, D = d/beta */
} cgfd_pml_face_t;

/* ---- src_t fields used per stage (forward/src_t.h:21-128) ---------------------------------- */
typedef struct {
  int32_t total_number;
  int32_t max_nt, max_stage;
  const int32_t *si, *sj, *sk;            /* local indices incl. ghosts */
  const float   *si_inc, *sj_inc, *sk_inc;
  const int32_t *it_begin, *it_end;
  int32_t is_surface_force_strict;
  int32_t total_number_surface_force;
  const int32_t *force_rate_indx;         /* [total_number_surface_force] */
  int32_t itype_spatial_ext;              /* CGFD_SRC_SPATIAL_* */
  int32_t ext_half_npoint;                /* 3 -> 7x7x7 footprint */
  float   ext_func_coef;                  /* 1.5 */
  int32_t force_actived, moment_actived;
  const float *Fx, *Fy, *Fz;              /* [total_number][max_nt][max_stage] */
  const float *Mxx, *Myy, *Mzz, *Mxz, *Myz, *Mxy;
  const float *Fx_rate, *Fy_rate, *Fz_rate; /* [total_number_surface_force][max_nt][max_stage] */
} cgfd_src_t;

/* ---- the whole static problem of one subdomain ---------------------------------------------- */
typedef struct {
  int32_t abi_version;                    /* CGFD_ABI_VERSION */
  cgfd_grid_t grid;
  cgfd_fd_t   fd;
  float   dt;
  int32_t medium_type;                    /* CGFD_MEDIUM_* */
  int32_t nmaxwell;                       /* visco GMB only; ncmp = 9 + 6*nmaxwell */
  int32_t ncmp;
  const float *metric[CGFD_NUM_METRIC];   /* each [nz][ny][nx]; metric and media may be host OR device pointers */
  /* media arrays, each [nz][ny][nx] (forward/md_t.c:57-260), rho already holds 1/rho
   * (forward/main_curv_col_el_3d.c:843):
   *   iso   : lambda, mu, rho
   *   vti   : c11, c13, c33, c55, c66, rho
   *   aniso : c11,c12,c13,c14,c15,c16,c22,c23,c24,c25,c26,c33,c34,c35,c36,c44,c45,c46,c55,c56,c66, rho
   *   visco : lambda, mu, rho, Ylam[0..N-1], Ymu[0..N-1] */
  int32_t nmedia;
  const float *media[CGFD_MAX_MEDIA];
  float   visco_wl[CGFD_MAX_MAXWELL];     /* relaxation frequencies md->wl */
  /* boundaries */
  int32_t free_top;                       /* bdry->is_sides_free[2][1] */
  int32_t timg_mode;                      /* CGFD_TIMG_* */
  cgfd_pml_face_t pml[3][2];              /* [idim][iside] */
  /* free-surface 3x3 matrices per surface point, [(j*nx+i)*9 + row*3 + col]
   * (forward/sv_curv_col_el_iso.c:1258-1375); NULL when free_top == 0 */
  const float *matVx2Vz, *matVy2Vz, *matF2Vz, *matD;
  /* exponential sponge (forward/bdry_t.c:840-890); all NULL when disabled */
  int32_t ablexp_enabled;
  int32_t ablexp_blk[6][7];               /* per block: enable, ni1,ni2,nj1,nj2,nk1,nk2 */
  const float *ablexp_Ex, *ablexp_Ey, *ablexp_Ez;
  cgfd_src_t src;
  /* x-y neighbours (rank ids, -1 = physical boundary): x1, x2, y1, y2 (forward/mympi_t.c:32-49) */
  int32_t neigh[4];
  /* Graves' constant-Q attenuation of the elastic media (md->visco_type == CONST_VISCO_GRAVES_QS): after the last RK stage
   * every component of the new wavefield is multiplied by exp(-pi f0 dt / Qs) at every physical point
   * (sv_curv_col_el_graves_Qs, forward/sv_curv_col_el.c:638-666, called at forward/drv_rk_curv_col.c:413-416).
   * graves_Qs = md->Qs [nz][ny][nx] (host or device pointer), graves_Qs_freq = md->visco_Qs_freq; NULL = off. */
  const float *graves_Qs;
  float   graves_Qs_freq;
} cgfd_problem_t;

typedef struct cgfd_b200_ctx cgfd_b200_ctx;

/* text of the last error on this thread */
const char *cgfd_b200_last_error(void);
int  cgfd_b200_abi_version(void);
/* sizeof(cgfd_problem_t) as compiled into the library, for binding self-checks */
size_t cgfd_b200_sizeof_problem(void);
/* number of visible CUDA devices (0 when none; never an error) */
int  cgfd_b200_device_count(void);

/* Upload the static problem to `device`, allocate the 4 wavefield levels (zeroed) and PML
 * auxiliary variables (zeroed). Replaces the level-pointer set-up of
 * forward/drv_rk_curv_col.c:101-120. */
int  cgfd_b200_create(const cgfd_problem_t *prob, int device, cgfd_b200_ctx **out);
void cgfd_b200_destroy(cgfd_b200_ctx *ctx);

/* wavefield at time level n (w_pre of forward/drv_rk_curv_col.c:101), [ncmp][nz][ny][nx] */
int  cgfd_b200_set_wavefield(cgfd_b200_ctx *ctx, const float *w);
int  cgfd_b200_get_wavefield(cgfd_b200_ctx *ctx, float *w);
/* PML auxiliary variables of one face at time level n, [9][slab k][slab j][slab i]
 * (forward/bdry_t.c:291-327) */
int  cgfd_b200_set_pml_aux(cgfd_b200_ctx *ctx, int idim, int iside, const float *aux);
int  cgfd_b200_get_pml_aux(cgfd_b200_ctx *ctx, int idim, int iside, float *aux);
size_t cgfd_b200_pml_aux_size(cgfd_b200_ctx *ctx, int idim, int iside); /* floats per level */

/* One RHS evaluation: rhs = L(w_cur) for the operator of [ipair][istage] with the sources of
 * (it, istage). The function-level seam for parity: the four *_onestage entry points
 * (forward/sv_curv_col_el_iso.c:22-37 and twins, called at forward/drv_rk_curv_col.c:233-286).
 * w_cur, rhs: host [ncmp][nz][ny][nx]; PML aux "cur" is whatever set_pml_aux stored, the aux rhs
 * of face (idim,iside) can be read back afterwards with cgfd_b200_get_pml_aux_rhs. */
int  cgfd_b200_onestage(cgfd_b200_ctx *ctx, int it, int ipair, int istage,
                        const float *w_cur, float *rhs);
int  cgfd_b200_get_pml_aux_rhs(cgfd_b200_ctx *ctx, int idim, int iside, float *aux_rhs);

/* Advance nsteps RK4 steps starting at step index it0 (ipair = it % 8): the time loop of
 * forward/drv_rk_curv_col.c:167-544 without the file outputs. */
int  cgfd_b200_run(cgfd_b200_ctx *ctx, int it0, int nsteps);

/* ---- output taps (forward/io_funcs.c:1584-1645) --------------------------------------------- */
/* Register n point indices (iptr into one component). After every step the ncmp values at each
 * point of the new time level are appended to a device-side record. */
int  cgfd_b200_set_record_points(cgfd_b200_ctx *ctx, int n, const int64_t *iptr, int max_nt);
/* Copy steps [it_first, it_first+nt) of the record to host as out[(it*ncmp + icmp)*n + ip]. */
int  cgfd_b200_get_record(cgfd_b200_ctx *ctx, int it_first, int nt, float *out);
/* Download a strided sub-box of component icmp of the current level n wavefield:
 * out[kk][jj][ii] = w[icmp][k1+kk*dk][j1+jj*dj][i1+ii*di] (io_snap_nc_put sub-volumes,
 * forward/io_funcs.c:1113-1268). */
int  cgfd_b200_get_box(cgfd_b200_ctx *ctx, int icmp, int i1, int ni, int di, int j1, int nj, int dj,
                       int k1, int nk, int dk, float *out);
/* peak-ground-motion maps accumulated over the steps run so far: 15 x [ny][nx]
 * (PG_calcu, forward/wav_t.c:379-455) */
int  cgfd_b200_get_pg(cgfd_b200_ctx *ctx, float *pg);

/* ---- multi-GPU: one process per GPU, x-y split, NCCL send/recv halo exchange ---------------- */
/* 128-byte NCCL unique id created on rank 0 and broadcast by the host program */
int  cgfd_b200_comm_unique_id(char id[128]);
int  cgfd_b200_comm_init(cgfd_b200_ctx *ctx, const char id[128], int rank, int nranks);

/* The halo plan itself (pure host logic, no GPU needed): the strips of `side` (0..3 = x1,x2,y1,y2) that are
 * sent / received for an operator with direction indices (dirx, diry), as {i1, ni, j1, nj, k1, nk} in local
 * indices including ghosts. Mirrors blk_macdrp_pack_mesg / unpack_mesg (forward/blk_t.c:576-808). */
int  cgfd_b200_halo_plan(const cgfd_grid_t *grid, int dirx, int diry, int side, int send_box[6], int recv_box[6]);

/* ---- one-shot set-up on the device ----------------------------------------------------------- */
/* gd_curv_metric_cal (forward/gd_t.c:190-402, called at forward/main_curv_col_el_3d.c:234): Jacobian and the nine metric
 * derivatives from the coordinate arrays x, y, z [nz][ny][nx] with the centred operator fd->fdc_indx / fdc_coef (fd_len terms),
 * ghosts mirrored in the order x, y, z. Every array may be a host or a device pointer; metric_out = jac, xi_x, xi_y, xi_z, eta_x,
 * eta_y, eta_z, zeta_x, zeta_y, zeta_z. Bit-identical to the reference function (no FMA contraction). Ghosts of inter-rank faces
 * are the caller's gd_curv_metric_exchange (forward/gd_t.c:407-475) as before. */
int  cgfd_b200_metric_from_coords(int device, const cgfd_grid_t *grid, const float *x, const float *y, const float *z, int fd_len,
                                  const int *fd_indx, const float *fd_coef, float *const metric_out[10]);

/* The free-surface conversion matrices of the four constitutive laws: sv_curv_col_el_iso_dvh2dvz (forward/sv_curv_col_el_iso.c:
 * 1258-1375), _vti_ (forward/sv_curv_col_el_vti.c:1085-1205), _aniso_ (forward/sv_curv_col_el_aniso.c:1267-1422) and
 * sv_curv_col_vis_iso_dvh2dvz (forward/sv_curv_col_vis_iso.c:353-507), called by the reference driver at
 * forward/drv_rk_curv_col.c:132-159. Reads medium_type, grid, metric[] and media[] of `prob` (host or device pointers, whole
 * arrays [nz][ny][nx]; only the k = nk2 plane is touched) and writes matVx2Vz, matVy2Vz [ny][nx][9] (host or device), matF2Vz
 * (isotropic medium only; zero otherwise, may be NULL) and matD (visco-elastic medium only; may be NULL otherwise). The
 * visco-elastic medium also needs the coordinate arrays x, y, z [nz][ny][nx] and the centred operator fd->fdc_indx / fdc_coef
 * (surface tangent for matD). Bit-identical to the reference functions (no FMA contraction, same order of operations). */
int  cgfd_b200_dvh2dvz(int device, const cgfd_problem_t *prob, const float *x, const float *y, const float *z, int fd_len,
                       const int *fd_indx, const float *fd_coef, float *matVx2Vz, float *matVy2Vz, float *matF2Vz, float *matD);

/* ---- distributed (finite-fault) sources ------------------------------------------------------- */
/* src_t dd_* (forward/src_t.h:94-126): `n` grid points indx[] = src->dd_indx (flat host index i + j*nx + k*nx*ny) that receive a
 * velocity source vi and / or a moment-rate source mij at every stage of every step, added at the point
 * (sv_curv_col_el_rhs_srcdd, forward/sv_curv_col_el.c:486-632; the smoothed variant does not exist in the reference either).
 * The time functions arrive block by block, as the reference reads them from file (src_dd_accit_loadstf,
 * forward/src_t.c:1929-2006): vi[nt][max_stage][n][3], mij[nt][max_stage][n][6] (xx yy zz yz xz xy) for the steps
 * it_first .. it_first+nt-1. Two blocks are resident: load block b+1 while the steps of block b run; the copy is asynchronous
 * (I/O stream) and the host arrays may be reused as soon as the call returns when they are pageable memory.
 * Steps outside every resident block get no dd source (dd_is_valid = 0 past dd_max_nt in the reference). */
int  cgfd_b200_dd_set_points(cgfd_b200_ctx *ctx, int n, const int64_t *indx, int vi_actived, int mij_actived, int max_stage,
                             int nt_per_block);
int  cgfd_b200_dd_load_block(cgfd_b200_ctx *ctx, int it_first, int nt, const float *vi, const float *mij);

/* ---- streaming outputs ------------------------------------------------------------------------- */
/* Snapshot / slice output without stalling the time loop (replaces the per-step io_snap_nc_put / io_slice_nc_put gathers,
 * forward/io_funcs.c:991-1268, called at forward/drv_rk_curv_col.c:498-512). A snapshot is the strided sub-box
 * box = {i1, ni, di, j1, nj, dj, k1, nk, dk} (local indices including ghosts) of `ncmps` wavefield components; during
 * cgfd_b200_run a frame of the new wavefield is packed on the device after every step it >= it1 with (it - it1) % tinv == 0
 * and copied asynchronously to host_out[frame][cmp][nk][nj][ni] (pinned memory gives a true asynchronous copy). When
 * cgfd_b200_run returns, every frame of the steps it covered is complete. Returns the snapshot id (>= 0) or -1. */
int  cgfd_b200_add_snapshot(cgfd_b200_ctx *ctx, int ncmps, const int *cmps, const int box[9], int it1, int tinv, int max_frames,
                            float *host_out);
/* frames written so far for snapshot `id` (-1: unknown id) */
int  cgfd_b200_snapshot_frames(cgfd_b200_ctx *ctx, int id);

/* Asynchronous time stepping for a host program that post-processes the outputs of one block of steps while the GPU runs the next:
 * cgfd_b200_run_async enqueues the steps and returns; rec_out (may be NULL; pinned memory for a true asynchronous copy) receives
 * the record samples of these steps as [nsteps][ncmp][npoints] once they exist. cgfd_b200_sync waits for everything enqueued: steps,
 * snapshot frames, record copies. cgfd_b200_snapshot_set_output points snapshot `id` at another host buffer for the frames that
 * follow (frame counter restarts at 0; frames already enqueued keep their destination), so that two buffers can alternate between blocks. cgfd_b200_host_alloc / _free give
 * page-locked host memory to a C program that does not link the CUDA runtime itself. */
int  cgfd_b200_run_async(cgfd_b200_ctx *ctx, int it0, int nsteps, float *rec_out);
int  cgfd_b200_sync(cgfd_b200_ctx *ctx);
/* wait for the block of steps a cgfd_b200_run_async call enqueued (identified by its last step) and for its outputs; blocks
 * enqueued after it keep running. The last four blocks can be waited for. */
int  cgfd_b200_wait_block(cgfd_b200_ctx *ctx, int it_last);
int  cgfd_b200_snapshot_set_output(cgfd_b200_ctx *ctx, int id, float *host_out, int max_frames);
int  cgfd_b200_host_alloc(size_t bytes, void **out);
void cgfd_b200_host_free(void *p);

/* The launch plan of the interior kernel for the tile rectangle rect = {bx0, bx1, by0, by1} (tiles of 32 x 8 points counted from
 * (ni1, nj1)); pure host logic, no GPU needed. pml_nlay[idim][iside] = layers of the CFS-PML on that face (0 = none). Returns the
 * number of blocks, *zchunk = rows per z chunk, order[b] = (chunk * ntiles_y + tile_y) * ntiles_x + tile_x of the b-th block:
 * tiles that meet an x / y PML slab first (longest job first); within each class bands of one wave of tiles run through their z chunks
 * in the direction of the kernel's march (dz = the zeta direction index of the stage's operator) before the next band starts.
 * free_top != 0: the top four rows are NOT part of the plan (they are left to the separate free-surface launch, CGFD_FUSE_TOP=0);
 * the default context runs them as planes of the top z chunk, i.e. asks for the plan with free_top = 0. -1 on bad arguments. */
int  cgfd_b200_launch_plan(const cgfd_grid_t *grid, const int pml_nlay[3][2], int free_top, int blocks_per_sm, int dz, const int rect[4],
                           int *zchunk, int *order, int capacity);

/* ---- measurement ----------------------------------------------------------------------------- */
/* When enabled, every launch of the dominant (interior RHS + RK) kernel is bracketed by CUDA
 * events on its own stream; get_profile returns the accumulated milliseconds and launch counts. */
int  cgfd_b200_set_profiling(cgfd_b200_ctx *ctx, int enabled);
int  cgfd_b200_get_profile(cgfd_b200_ctx *ctx, double *main_kernel_ms, int64_t *main_kernel_launches,
                           int64_t *total_launches);
/* time (ms, CUDA events on the compute stream) of the last cgfd_b200_run call */
int  cgfd_b200_last_run_ms(cgfd_b200_ctx *ctx, double *ms);
/* 1 when the context runs the kernels specialised for vertically deformed grids (xi_y = xi_z = eta_x = eta_z == 0 at every
 * physical point, detected at create time; CGFD_GZ=0 in the environment keeps the general kernels), else 0 */
int  cgfd_b200_grid_class(cgfd_b200_ctx *ctx);
/* 1 when the free-surface rows (rhs_timg_z2 / rhs_vlow_z2, the top four rows) run as planes of the interior kernel's top z chunk
 * (the default with a free top), 0 when they are a separate launch (CGFD_FUSE_TOP=0) or the top is not a free surface */
int  cgfd_b200_top_fused(cgfd_b200_ctx *ctx);
/* choose a kernel variant by name; NULL/"" = default */
int  cgfd_b200_set_variant(cgfd_b200_ctx *ctx, const char *name);

#ifdef __cplusplus
}
#endif
#endif
