/*
 * drv_rk_curv_col_b200.c -- drop-in replacement of forward/drv_rk_curv_col.c for the CGFD3D host
 * program: same entry point, same arguments (forward/drv_rk_curv_col.h:17-37), called from
 * forward/main_curv_col_el_3d.c:849-858. Link this file and libcgfd3d_b200.so instead of
 * drv_rk_curv_col.o and sv_curv_col_*.o; the par file, every *_t struct and the SAC / nc output layout
 * stay the reference's own.
 *
 * What runs where:
 *   host (reference code, unchanged): nc file creation,
 *        io_recv_keep / io_line_keep / io_slice_nc_put / io_snap_nc_put / PG_slice_output;
 *   GPU (libcgfd3d_b200.so through the C ABI of include/cgfd3d_b200.h): the whole RK4 stage loop --
 *        RHS, CFS-PML, free surface, sources, RK update, halo exchange, PGV/PGA/PGD maps; the free-surface matrices (*_dvh2dvz).
 * After every step only the samples the output taps need are copied back: the receiver / line points
 * (recorded on the device) and, on the steps a slice or snapshot is due, its strided sub-boxes. They are
 * written into the host wavefield level the reference functions read (w_end), so those run unmodified.
 *
 * This file includes the reference's headers and therefore compiles only where the reference tree is.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <mpi.h>

#include "fdlib_mem.h"
#include "fdlib_math.h"
#include "blk_t.h"
#include "drv_rk_curv_col.h"

#include "cgfd3d_b200.h"

#define DIE(...) do { fprintf(stderr, "cgfd3d_b200 driver: " __VA_ARGS__); fprintf(stderr, "\n"); fflush(stderr); \
                      MPI_Abort(MPI_COMM_WORLD, 1); exit(1); } while (0)
#define GPU(call) do { if ((call) != 0) DIE("%s failed: %s", #call, cgfd_b200_last_error()); } while (0)

static void flatten_fd(const fd_t *fd, cgfd_fd_t *o)
{
  memset(o, 0, sizeof(*o));
  for (int s = 0; s < CGFD_NUM_STAGES; s++) { o->rk_b[s] = fd->rk_b[s]; if (s < CGFD_NUM_STAGES - 1) o->rk_a[s] = fd->rk_a[s]; }
  if (fd->num_of_pairs != CGFD_NUM_PAIRS || fd->num_rk_stages != CGFD_NUM_STAGES) DIE("expected 8 operator pairs and 4 RK stages");
  for (int p = 0; p < CGFD_NUM_PAIRS; p++) for (int s = 0; s < CGFD_NUM_STAGES; s++) {
    fd_op_t *ox = &fd->pair_fdx_op[p][s][fd->num_of_fdx_op - 1];
    fd_op_t *oy = &fd->pair_fdy_op[p][s][fd->num_of_fdy_op - 1];
    fd_op_t *oz = &fd->pair_fdz_op[p][s][fd->num_of_fdz_op - 1];
    fd_op_t *ops[3] = { ox, oy, oz };
    for (int a = 0; a < 3; a++) {
      if (ops[a]->total_len != 5) DIE("interior operator must have 5 points");
      int d = (ops[a]->indx[0] == -1) ? 0 : 1;   /* {-1..3} -> 0, {-3..1} -> 1 */
      o->dir[p][s][a] = d;
      for (int n = 0; n < 5; n++) { o->indx[d][n] = ops[a]->indx[n]; o->coef[d][n] = ops[a]->coef[n]; }
    }
    if (fd->num_of_fdz_op != 4) DIE("expected 4 zeta operators per stage (forward/fd_t.c:197-199)");
    int dz = o->dir[p][s][2];
    for (int lay = 1; lay <= 2; lay++) {
      fd_op_t *l = &fd->pair_fdz_op[p][s][lay];
      o->lay_len[lay][dz] = l->total_len;
      for (int n = 0; n < l->total_len; n++) { o->lay_indx[lay][dz][n] = l->indx[n]; o->lay_coef[lay][dz][n] = l->coef[n]; }
    }
  }
}

/* one streaming output (slice or snapshot): components c1 .. c1+ncmp-1 of the strided box, every tinv steps from it1 on */
typedef struct {
  int id, c1, ncmp, box[9], it1, tinv, cap, frame;
  size_t elems;
  float *buf[2];
} tap_t;
static void tap_set(tap_t *T, int c1, int c2, int i1, int ni, int di, int j1, int nj, int dj, int k1, int nk, int dk, int it1, int tinv)
{
  int b[9] = { i1, ni, di, j1, nj, dj, k1, nk, dk };
  memcpy(T->box, b, sizeof(b));
  T->c1 = c1; T->ncmp = c2 - c1 + 1; T->it1 = it1; T->tinv = tinv > 0 ? tinv : 1;
  T->elems = (size_t)ni * nj * nk;
}
/* if a frame of this tap is due at step `it`: copy it from the block's host buffer into the host level `w` at the box's own indices */
static void tap_scatter(tap_t *T, int set, int it, float *w, size_t siz_icmp, size_t siz_line, size_t siz_slice)
{
  if (it < T->it1 || (it - T->it1) % T->tinv != 0 || T->elems == 0) return;
  const int *b = T->box;
  const float *f = T->buf[set] + (size_t)T->frame * T->ncmp * T->elems;
  for (int c = 0; c < T->ncmp; c++) {
    float *var = w + (size_t)(T->c1 + c) * siz_icmp;
    size_t n = (size_t)c * T->elems;
    for (int kk = 0; kk < b[7]; kk++) for (int jj = 0; jj < b[4]; jj++) {
      float *row = var + (size_t)(b[3] + jj * b[5]) * siz_line + (size_t)(b[6] + kk * b[8]) * siz_slice + b[0];
      for (int ii = 0; ii < b[1]; ii++) row[(size_t)ii * b[2]] = f[n++];
    }
  }
  T->frame++;
}

void
drv_rk_curv_col_allstep(
  fd_t            *fd,
  gd_t        *gd,
  gdcurv_metric_t *metric,
  md_t      *md,
  src_t     *src,
  bdry_t    *bdry,
  wav_t     *wav,
  mympi_t    *mympi,
  iorecv_t   *iorecv,
  ioline_t   *ioline,
  ioslice_t  *ioslice,
  iosnap_t   *iosnap,
  float dt, int nt_total, float t0,
  char *output_fname_part,
  char *output_dir,
  int qc_check_nan_num_of_step,
  const int output_all,
  const int verbose)
{
  int myid = mympi->myid;
  int *topoid = mympi->topoid;
  const int ncmp = wav->ncmp;
  const size_t siz_icmp = wav->siz_icmp;

  /* host levels the reference's output functions work on (forward/drv_rk_curv_col.c:101-104) */
  float *w_pre = wav->v5d + wav->siz_ilevel * 0;
  float *w_rhs = wav->v5d + wav->siz_ilevel * 2;
  float *w_end = wav->v5d + wav->siz_ilevel * 3;

  if (myid == 0 && verbose > 0) fprintf(stdout, "prepare slice nc output ...\n");
  ioslice_nc_t ioslice_nc;
  io_slice_nc_create(ioslice, wav->ncmp, wav->visco_type, wav->cmp_name, gd->ni, gd->nj, gd->nk, topoid, &ioslice_nc);
  if (myid == 0 && verbose > 0) fprintf(stdout, "prepare snap nc output ...\n");
  iosnap_nc_t iosnap_nc;
  io_snap_nc_create(iosnap, &iosnap_nc, topoid);

  /* ---- flatten the reference structs into the C ABI problem description ---------------------------- */
  cgfd_problem_t P;
  memset(&P, 0, sizeof(P));
  P.abi_version = CGFD_ABI_VERSION;
  P.grid.nx = gd->nx; P.grid.ny = gd->ny; P.grid.nz = gd->nz;
  P.grid.ni1 = gd->ni1; P.grid.ni2 = gd->ni2; P.grid.nj1 = gd->nj1; P.grid.nj2 = gd->nj2; P.grid.nk1 = gd->nk1; P.grid.nk2 = gd->nk2;
  flatten_fd(fd, &P.fd);
  P.dt = dt;
  P.medium_type = md->medium_type;
  P.nmaxwell = (md->medium_type == CONST_MEDIUM_VISCOELASTIC_ISO) ? md->nmaxwell : 0;
  P.ncmp = ncmp;
  const float *mt[10] = { metric->jac, metric->xi_x, metric->xi_y, metric->xi_z, metric->eta_x, metric->eta_y, metric->eta_z,
                          metric->zeta_x, metric->zeta_y, metric->zeta_z };
  for (int m = 0; m < 10; m++) P.metric[m] = mt[m];
  if (md->medium_type == CONST_MEDIUM_ELASTIC_ISO) {
    P.nmedia = 3; P.media[0] = md->lambda; P.media[1] = md->mu; P.media[2] = md->rho;
  } else if (md->medium_type == CONST_MEDIUM_ELASTIC_VTI) {
    const float *a[6] = { md->c11, md->c13, md->c33, md->c55, md->c66, md->rho };
    P.nmedia = 6; for (int m = 0; m < 6; m++) P.media[m] = a[m];
  } else if (md->medium_type == CONST_MEDIUM_ELASTIC_ANISO) {
    const float *a[22] = { md->c11, md->c12, md->c13, md->c14, md->c15, md->c16, md->c22, md->c23, md->c24, md->c25, md->c26,
                           md->c33, md->c34, md->c35, md->c36, md->c44, md->c45, md->c46, md->c55, md->c56, md->c66, md->rho };
    P.nmedia = 22; for (int m = 0; m < 22; m++) P.media[m] = a[m];
  } else if (md->medium_type == CONST_MEDIUM_VISCOELASTIC_ISO) {
    P.nmedia = 3 + 2 * md->nmaxwell; P.media[0] = md->lambda; P.media[1] = md->mu; P.media[2] = md->rho;
    if (md->nmaxwell > CGFD_MAX_MAXWELL) DIE("too many Maxwell bodies");
    for (int n = 0; n < md->nmaxwell; n++) { P.media[3 + n] = md->Ylam[n]; P.media[3 + md->nmaxwell + n] = md->Ymu[n]; P.visco_wl[n] = md->wl[n]; }
  } else DIE("medium_type=%d is not supported", md->medium_type);
  if (md->visco_type == CONST_VISCO_GRAVES_QS) { P.graves_Qs = md->Qs; P.graves_Qs_freq = md->visco_Qs_freq; }
  P.free_top = bdry->is_sides_free[CONST_NDIM - 1][1];
  P.timg_mode = getenv("CGFD_TIMG_MIRROR") ? CGFD_TIMG_MIRROR : CGFD_TIMG_ZERO;
  for (int d = 0; d < 3; d++) for (int s = 0; s < 2; s++) {
    if (bdry->is_enable_pml != 1 || bdry->is_sides_pml[d][s] != 1) continue;
    P.pml[d][s].enabled = 1; P.pml[d][s].nlay = bdry->num_of_layers[d][s];
    P.pml[d][s].A = bdry->A[d][s]; P.pml[d][s].B = bdry->B[d][s]; P.pml[d][s].D = bdry->D[d][s];
  }
  if (P.free_top) {
    /* free-surface conversion matrices: the *_dvh2dvz call of forward/drv_rk_curv_col.c:132-159, computed on the device by the
     * library (bit-identical to the reference functions) into the arrays bdry_free_set allocated */
    if (md->medium_type == CONST_MEDIUM_VISCOELASTIC_ISO && md->visco_type != CONST_VISCO_GMB)
      DIE("conversion matrix for visco_type=%d is not implemented (neither is it in the reference, forward/drv_rk_curv_col.c:150-157)", md->visco_type);
    int ndev0 = cgfd_b200_device_count();
    if (ndev0 <= 0) DIE("no CUDA device visible (this driver has no CPU fallback)");
    GPU(cgfd_b200_dvh2dvz(myid % ndev0, &P, gd->x3d, gd->y3d, gd->z3d, fd->fdc_len, fd->fdc_indx, fd->fdc_coef,
                          bdry->matVx2Vz2, bdry->matVy2Vz2, bdry->matF2Vz2, bdry->matD));
    P.matVx2Vz = bdry->matVx2Vz2; P.matVy2Vz = bdry->matVy2Vz2; P.matF2Vz = bdry->matF2Vz2; P.matD = bdry->matD;
  }
  if (bdry->is_enable_ablexp == 1) {
    P.ablexp_enabled = 1;
    for (int n = 0; n < CONST_NDIM_2; n++) {
      bdry_block_t *D = bdry->bdry_blk + n;
      int v[7] = { D->enable, D->ni1, D->ni2, D->nj1, D->nj2, D->nk1, D->nk2 };
      memcpy(P.ablexp_blk[n], v, sizeof(v));
    }
    P.ablexp_Ex = bdry->ablexp_Ex; P.ablexp_Ey = bdry->ablexp_Ey; P.ablexp_Ez = bdry->ablexp_Ez;
  }
  if (src->dd_is_valid == 1 && src->dd_is_add_at_point != 1) DIE("dd sources with spatial smoothing are not implemented (neither are they in the reference, forward/sv_curv_col_el.c:578-582)");
  cgfd_src_t *S = &P.src;
  S->total_number = src->total_number; S->max_nt = src->max_nt; S->max_stage = src->max_stage;
  S->si = src->si; S->sj = src->sj; S->sk = src->sk;
  S->si_inc = src->si_inc; S->sj_inc = src->sj_inc; S->sk_inc = src->sk_inc;
  S->it_begin = src->it_begin; S->it_end = src->it_end;
  S->is_surface_force_strict = src->is_surface_force_strict;
  S->total_number_surface_force = src->total_number_surface_force;
  S->force_rate_indx = src->force_rate_indx;
  S->itype_spatial_ext = src->itype_spatial_ext; S->ext_half_npoint = src->ext_half_npoint; S->ext_func_coef = src->ext_func_coef;
  S->force_actived = src->force_actived; S->moment_actived = src->moment_actived;
  S->Fx = src->Fx; S->Fy = src->Fy; S->Fz = src->Fz;
  S->Mxx = src->Mxx; S->Myy = src->Myy; S->Mzz = src->Mzz; S->Mxz = src->Mxz; S->Myz = src->Myz; S->Mxy = src->Mxy;
  S->Fx_rate = src->Fx_rate; S->Fy_rate = src->Fy_rate; S->Fz_rate = src->Fz_rate;
  for (int n = 0; n < 4; n++) P.neigh[n] = (mympi->neighid[n] == MPI_PROC_NULL) ? -1 : mympi->neighid[n];

  /* one rank drives one GPU: local device = rank modulo the devices of this node */
  int ndev = cgfd_b200_device_count();
  if (ndev <= 0) DIE("no CUDA device visible (this driver has no CPU fallback)");
  cgfd_b200_ctx *ctx = NULL;
  GPU(cgfd_b200_create(&P, myid % ndev, &ctx));
  GPU(cgfd_b200_set_wavefield(ctx, w_pre));
  for (int d = 0; d < 3; d++) for (int s = 0; s < 2; s++)
    if (P.pml[d][s].enabled) {
      /* aux vars hold ncmp components per level in the reference; the first 9 are the ones in use */
      GPU(cgfd_b200_set_pml_aux(ctx, d, s, bdry->auxvar[d][s].var));
    }
  int nproc = mympi->nprocx * mympi->nprocy;
  if (nproc > 1) {
    char id[128];
    if (myid == 0) GPU(cgfd_b200_comm_unique_id(id));
    MPI_Bcast(id, 128, MPI_CHAR, 0, MPI_COMM_WORLD);
    GPU(cgfd_b200_comm_init(ctx, id, myid, nproc));
  }

  /* distributed (finite-fault) sources: points once, the time block src_dd_read2local left in memory first; the following
   * blocks are read from file by the reference's own src_dd_accit_loadstf inside the loop (forward/drv_rk_curv_col.c:179-181) */
  if (src->dd_is_valid == 1) {
    int64_t *dd_indx = (int64_t *)malloc(sizeof(int64_t) * src->dd_total_number);
    for (int q = 0; q < src->dd_total_number; q++) dd_indx[q] = (int64_t)src->dd_indx[q];
    GPU(cgfd_b200_dd_set_points(ctx, src->dd_total_number, dd_indx, src->dd_vi_actived, src->dd_mij_actived, src->max_stage, src->dd_nt_per_read));
    /* the block src_dd_read2local left in memory holds dd_nt_per_read steps (forward/src_t.c:1841); dd_nt_this_read is only
     * assigned by the reloads inside src_dd_accit_loadstf (forward/src_t.c:1962-1964) and is uninitialised here */
    int nt_first = src->dd_nt_per_read < src->dd_max_nt ? src->dd_nt_per_read : src->dd_max_nt;
    GPU(cgfd_b200_dd_load_block(ctx, 0, nt_first, src->dd_vi, src->dd_mij));
    free(dd_indx);
  }

  /* ---- output taps ---------------------------------------------------------------------------------------
   * Every grid point io_recv_keep / io_line_keep will read is a record point (sampled on the device after every step); every slice
   * and snapshot is a streaming tap (cgfd_b200_add_snapshot: frames packed on the device and copied to pinned host memory on an
   * I/O stream while the steps run). The time loop advances in BLOCKS of steps: while the GPU runs block b+1, the host feeds the
   * samples and frames of block b, step by step, into the host wavefield level the reference's own io_recv_keep / io_line_keep /
   * io_slice_nc_put / io_snap_nc_put read (w_end), so those functions run unmodified and the files come out as the reference's. */
  int nrec = iorecv->total_number * CONST_2_NDIM;
  for (int n = 0; n < ioline->num_of_lines; n++) nrec += ioline->line_nr[n];
  int64_t *rec_iptr = (int64_t *)malloc(sizeof(int64_t) * (nrec > 0 ? nrec : 1));
  int ir = 0;
  for (int n = 0; n < iorecv->total_number; n++) for (int q = 0; q < CONST_2_NDIM; q++) rec_iptr[ir++] = iorecv->recvone[n].indx1d[q];
  for (int n = 0; n < ioline->num_of_lines; n++) for (int q = 0; q < ioline->line_nr[n]; q++) rec_iptr[ir++] = ioline->recv_iptr[n][q];
  if (nrec > 0) GPU(cgfd_b200_set_record_points(ctx, nrec, rec_iptr, nt_total));

  int BLK = 32;   /* steps per block */
  if (getenv("CGFD_DRV_BLOCK")) BLK = atoi(getenv("CGFD_DRV_BLOCK"));
  if (BLK < 1 || output_all == 1) BLK = 1;

  /* taps: slices x, y, z first, then the snapshots */
  const int nslice = ioslice_nc.num_of_slice_x + ioslice_nc.num_of_slice_y + ioslice_nc.num_of_slice_z;
  const int ntap = nslice + iosnap->num_of_snap;
  tap_t *tap = (tap_t *)calloc(ntap > 0 ? ntap : 1, sizeof(tap_t));
  {
    int t = 0;
    const int c2s = (wav->visco_type == CONST_VISCO_GMB) ? 8 : ncmp - 1;   /* io_slice_nc_put writes components 0 .. c2s */
    for (int n = 0; n < ioslice_nc.num_of_slice_x; n++, t++)
      tap_set(&tap[t], 0, c2s, ioslice->slice_x_indx[n], 1, 1, gd->nj1, gd->nj, 1, gd->nk1, gd->nk, 1, 0, 1);
    for (int n = 0; n < ioslice_nc.num_of_slice_y; n++, t++)
      tap_set(&tap[t], 0, c2s, gd->ni1, gd->ni, 1, ioslice->slice_y_indx[n], 1, 1, gd->nk1, gd->nk, 1, 0, 1);
    for (int n = 0; n < ioslice_nc.num_of_slice_z; n++, t++)
      tap_set(&tap[t], 0, c2s, gd->ni1, gd->ni, 1, gd->nj1, gd->nj, 1, ioslice->slice_z_indx[n], 1, 1, 0, 1);
    for (int n = 0; n < iosnap->num_of_snap; n++, t++) {
      int c1 = (iosnap->out_vel[n] == 1) ? 0 : 3;
      int c2 = (iosnap->out_stress[n] == 1 || iosnap->out_strain[n] == 1) ? 8 : 2;
      tap_set(&tap[t], c1, c2, iosnap->i1[n], iosnap->ni[n], iosnap->di[n], iosnap->j1[n], iosnap->nj[n], iosnap->dj[n],
              iosnap->k1[n], iosnap->nk[n], iosnap->dk[n], iosnap->it1[n], iosnap->dit[n]);
    }
    for (t = 0; t < ntap; t++) {
      tap_t *T = &tap[t];
      T->cap = BLK / T->tinv + 2;   /* frames one block can produce */
      size_t bytes = (size_t)T->cap * T->ncmp * T->elems * sizeof(float);
      for (int b = 0; b < 2; b++) { void *q = NULL; GPU(cgfd_b200_host_alloc(bytes, &q)); T->buf[b] = (float *)q; }
      int cm[32]; for (int c = 0; c < T->ncmp; c++) cm[c] = T->c1 + c;
      T->id = cgfd_b200_add_snapshot(ctx, T->ncmp, cm, T->box, T->it1, T->tinv, T->cap, T->buf[0]);
      if (T->id < 0) DIE("cgfd_b200_add_snapshot failed: %s", cgfd_b200_last_error());
    }
  }
  float *rec_blk[2] = { NULL, NULL };
  if (nrec > 0) for (int b = 0; b < 2; b++) { void *q = NULL; GPU(cgfd_b200_host_alloc((size_t)BLK * ncmp * nrec * sizeof(float), &q)); rec_blk[b] = (float *)q; }

  if (myid == 0 && verbose > 0) fprintf(stdout, "start time loop (GPU, blocks of %d steps) ...\n", BLK);
  struct timespec ts0, ts1;
  clock_gettime(CLOCK_MONOTONIC, &ts0);
  int it_enq = 0;                 /* next step to enqueue */
  int blk_first[2] = { 0, 0 }, blk_n[2] = { 0, 0 };
  int cur = 0;                    /* buffer set of the block being enqueued */
  /* enqueue the first block, then: wait for block b, enqueue block b+1, post-process block b on the host while b+1 runs */
  while (it_enq < nt_total || blk_n[1 - cur] > 0)
  {
    if (it_enq < nt_total) {
      int n = nt_total - it_enq < BLK ? nt_total - it_enq : BLK;
      if (src->dd_is_valid == 1) {
        /* distributed sources: the reference reloads their time functions from file every dd_nt_per_read steps
         * (src_dd_accit_loadstf, forward/drv_rk_curv_col.c:179-181); a block never crosses a reload, and the reloaded block goes
         * to the device (two blocks resident) before the steps that use it are enqueued */
        src_dd_accit_loadstf(src, it_enq, myid);
        if (src->dd_is_valid == 1 && it_enq > 0 && src->dd_it_here == 0)
          GPU(cgfd_b200_dd_load_block(ctx, it_enq, src->dd_nt_this_read, src->dd_vi, src->dd_mij));
        int m = 1;
        while (m < n && src->dd_is_valid == 1 && src->dd_it_here + 1 < src->dd_nt_per_read && it_enq + m < src->dd_max_nt) {
          src_dd_accit_loadstf(src, it_enq + m, myid);   /* only advances the counter: no reload inside this range */
          m++;
        }
        n = m;
      }
      for (int t = 0; t < ntap; t++) GPU(cgfd_b200_snapshot_set_output(ctx, tap[t].id, tap[t].buf[cur], tap[t].cap));
      GPU(cgfd_b200_run_async(ctx, it_enq, n, rec_blk[cur]));
      blk_first[cur] = it_enq; blk_n[cur] = n;
      it_enq += n;
    } else {
      blk_n[cur] = 0;
    }
    /* post-process the PREVIOUS block while the one just enqueued runs (output_all = 1 dumps the whole device state after every
     * step: no look-ahead then, the block just enqueued is waited for and processed at once) */
    const int prev = (output_all == 1) ? cur : 1 - cur;
    if (blk_n[prev] > 0) {
      GPU(cgfd_b200_wait_block(ctx, blk_first[prev] + blk_n[prev] - 1));
      for (int t = 0; t < ntap; t++) tap[t].frame = 0;
      for (int q = 0; q < blk_n[prev]; q++) {
        const int it = blk_first[prev] + q;
        const float t_end = it * dt + t0 + dt;
        if (myid == 0 && verbose > 10) fprintf(stdout, "-> it=%d, t=%f\n", it, it * dt + t0);
        if (nrec > 0) {
          const float *r = rec_blk[prev] + (size_t)q * ncmp * nrec;
          for (int c = 0; c < ncmp; c++) for (int p = 0; p < nrec; p++) w_end[(size_t)c * siz_icmp + rec_iptr[p]] = r[(size_t)c * nrec + p];
          io_recv_keep(iorecv, w_end, it, ncmp, siz_icmp);
          io_line_keep(ioline, w_end, it, ncmp, siz_icmp);
        }
        for (int t = 0; t < ntap; t++) tap_scatter(&tap[t], prev, it, w_end, siz_icmp, gd->siz_iy, gd->siz_iz);
        if (nslice > 0) io_slice_nc_put(ioslice, &ioslice_nc, gd, w_end, w_rhs, it, t_end, 0, ncmp - 1, wav->visco_type);
        io_snap_nc_put(iosnap, &iosnap_nc, gd, md, wav, w_end, w_rhs, nt_total, it, t_end, 1, 1, 1);
        if (output_all == 1) {
          char ou_file[CONST_MAX_STRLEN];
          GPU(cgfd_b200_get_wavefield(ctx, w_end));
          io_build_fname_time(output_dir, "w3d", ".nc", topoid, it, ou_file);
          io_var3d_export_nc(ou_file, w_end, wav->cmp_pos, wav->cmp_name, wav->ncmp, gd->index_name, gd->nx, gd->ny, gd->nz);
        }
      }
      blk_n[prev] = 0;
    }
    cur = 1 - cur;
  }
  GPU(cgfd_b200_sync(ctx));
  clock_gettime(CLOCK_MONOTONIC, &ts1);
  if (myid == 0 && verbose > 0) {
    double sec = (ts1.tv_sec - ts0.tv_sec) + 1e-9 * (ts1.tv_nsec - ts0.tv_nsec);
    fprintf(stdout, "GPU time loop: %d steps in %.3f s = %.4f Gpoint-updates/s on this rank\n", nt_total, sec,
            (double)gd->ni * gd->nj * gd->nk * nt_total / sec / 1e9);
  }

  if (bdry->is_sides_free[CONST_NDIM - 1][1] == 1) {
    float *PG = (float *)fdlib_mem_calloc_1d_float(CONST_NDIM_5 * gd->ny * gd->nx, 0.0, "PG malloc");
    GPU(cgfd_b200_get_pg(ctx, PG));
    PG_slice_output(PG, gd, output_dir, output_fname_part, topoid);
    free(PG);
  }
  /* leave the final state where the reference leaves it: level n of the host wavefield */
  GPU(cgfd_b200_get_wavefield(ctx, w_pre));
  io_slice_nc_close(&ioslice_nc);
  io_snap_nc_close(&iosnap_nc);
  cgfd_b200_destroy(ctx);
  free(rec_iptr);
  for (int t = 0; t < ntap; t++) for (int b = 0; b < 2; b++) cgfd_b200_host_free(tap[t].buf[b]);
  for (int b = 0; b < 2; b++) cgfd_b200_host_free(rec_blk[b]);
  free(tap);
  (void)qc_check_nan_num_of_step;
  return;
}
