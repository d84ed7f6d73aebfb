/*
 * ref_flat_adapter.c -- TEST INFRASTRUCTURE, NOT PRODUCT.
 *
 * Lets the parity tests and bench.py's cpu_baseline / --impl reference legs call the UNMODIFIED
 * reference functions of the hot path on the flat problem description of
 * include/cgfd3d_b200.h. It rebuilds the reference's own structs (gd_t, gdcurv_metric_t, md_t,
 * wav_t, bdry_t, src_t, fd_t, mympi_t, io*_t) with the reference's own *_init functions where they
 * exist, copies the flat arrays in, and then calls
 *     sv_curv_col_{el_iso,el_vti,el_aniso,vis_iso}_onestage   (forward/sv_curv_col_el_iso.c:22 ...)
 *     drv_rk_curv_col_allstep                                 (forward/drv_rk_curv_col.c:27)
 * from oracle/_ref/libcgfd_ref.so (built by oracle/Makefile from the sources in /root/reference;
 * the traction-image file is the ZERO-guarded copy, see Makefile). Nothing here computes physics.
 *
 * Compiled only where /root/reference exists; the resulting oracle/_ref/libcgfd_ref_flat.so travels
 * to the GPU box.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>

#include "mpi.h"
#include "constants.h"
#include "fdlib_mem.h"
#include "fd_t.h"
#include "gd_t.h"
#include "md_t.h"
#include "wav_t.h"
#include "bdry_t.h"
#include "src_t.h"
#include "mympi_t.h"
#include "io_funcs.h"
#include "blk_t.h"
#include "sv_curv_col_el.h"
#include "sv_curv_col_el_iso.h"
#include "sv_curv_col_el_vti.h"
#include "sv_curv_col_el_aniso.h"
#include "sv_curv_col_vis_iso.h"
#include "drv_rk_curv_col.h"

#include "../include/cgfd3d_b200.h"

typedef struct {
  cgfd_problem_t p;
  fd_t fd;
  gd_t gd;
  gdcurv_metric_t metric;
  md_t md;
  wav_t wav;
  bdry_t bdry;
  src_t src;
  mympi_t mympi;
  iorecv_t iorecv;
  ioline_t ioline;
  ioslice_t ioslice;
  iosnap_t iosnap;
  size_t nvol;
} ref_t;

static float *dupf(const float *s, size_t n)
{
  float *d = (float *)calloc(n ? n : 1, sizeof(float));
  if (s && n) memcpy(d, s, n * sizeof(float));
  return d;
}
static int *dupi(const int32_t *s, size_t n)
{
  int *d = (int *)calloc(n ? n : 1, sizeof(int));
  if (s) for (size_t i = 0; i < n; i++) d[i] = s[i];
  return d;
}

void *cgfd_ref_create(const cgfd_problem_t *p)
{
  if (p->abi_version != CGFD_ABI_VERSION) { fprintf(stderr, "cgfd_ref_create: abi mismatch\n"); return NULL; }
  ref_t *r = (ref_t *)calloc(1, sizeof(ref_t));
  r->p = *p;
  const cgfd_grid_t *g = &p->grid;

  fd_set_macdrp(&r->fd);

  /* gd_t: what gd_indx_set (forward/gd_t.c:2761-2915) leaves behind for one rank */
  gd_t *gd = &r->gd;
  gd->type = GD_TYPE_CURV;
  gd->nx = g->nx; gd->ny = g->ny; gd->nz = g->nz;
  gd->ni1 = g->ni1; gd->ni2 = g->ni2; gd->nj1 = g->nj1; gd->nj2 = g->nj2; gd->nk1 = g->nk1; gd->nk2 = g->nk2;
  gd->ni = g->ni2 - g->ni1 + 1; gd->nj = g->nj2 - g->nj1 + 1; gd->nk = g->nk2 - g->nk1 + 1;
  gd->npoint_ghosts = 3; gd->fdx_nghosts = 3; gd->fdy_nghosts = 3; gd->fdz_nghosts = 3;
  gd->gni1 = 0; gd->gnj1 = 0; gd->gnk1 = 0;
  gd->gni2 = gd->ni - 1; gd->gnj2 = gd->nj - 1; gd->gnk2 = gd->nk - 1;
  gd->siz_iy = gd->siz_line = g->nx;
  gd->siz_iz = gd->siz_slice = (size_t)g->nx * g->ny;
  gd->siz_icmp = gd->siz_volume = (size_t)g->nx * g->ny * g->nz;
  r->nvol = gd->siz_volume;
  {
    static char *index_name[3] = {"i", "j", "k"};
    gd->index_name = index_name;
  }

  gd_curv_metric_init(gd, &r->metric);
  {
    float *dst[10] = { r->metric.jac, r->metric.xi_x, r->metric.xi_y, r->metric.xi_z,
                       r->metric.eta_x, r->metric.eta_y, r->metric.eta_z,
                       r->metric.zeta_x, r->metric.zeta_y, r->metric.zeta_z };
    for (int m = 0; m < 10; m++) memcpy(dst[m], p->metric[m], r->nvol * sizeof(float));
  }

  int visco_type = (p->medium_type == CONST_MEDIUM_VISCOELASTIC_ISO) ? CONST_VISCO_GMB : (p->graves_Qs ? CONST_VISCO_GRAVES_QS : 0);
  md_init(gd, &r->md, p->medium_type, visco_type, p->nmaxwell);
  if (p->graves_Qs) {   /* md_init allocated md->Qs (forward/md_t.c:47-50) */
    memcpy(r->md.Qs, p->graves_Qs, (size_t)g->nx * g->ny * g->nz * sizeof(float));
    r->md.visco_Qs_freq = p->graves_Qs_freq;
  }
  {
    md_t *md = &r->md;
    size_t nb = r->nvol * sizeof(float);
    if (p->medium_type == CONST_MEDIUM_ELASTIC_ISO) {
      memcpy(md->lambda, p->media[0], nb); memcpy(md->mu, p->media[1], nb); memcpy(md->rho, p->media[2], nb);
    } else if (p->medium_type == CONST_MEDIUM_ELASTIC_VTI) {
      float *d[6] = { md->c11, md->c13, md->c33, md->c55, md->c66, md->rho };
      for (int m = 0; m < 6; m++) memcpy(d[m], p->media[m], nb);
    } else if (p->medium_type == CONST_MEDIUM_ELASTIC_ANISO) {
      float *d[22] = { md->c11, md->c12, md->c13, md->c14, md->c15, md->c16, md->c22, md->c23, md->c24, md->c25,
                       md->c26, md->c33, md->c34, md->c35, md->c36, md->c44, md->c45, md->c46, md->c55, md->c56,
                       md->c66, md->rho };
      for (int m = 0; m < 22; m++) memcpy(d[m], p->media[m], nb);
    } else if (p->medium_type == CONST_MEDIUM_VISCOELASTIC_ISO) {
      memcpy(md->lambda, p->media[0], nb); memcpy(md->mu, p->media[1], nb); memcpy(md->rho, p->media[2], nb);
      for (int n = 0; n < p->nmaxwell; n++) {
        memcpy(md->Ylam[n], p->media[3 + n], nb);
        memcpy(md->Ymu[n], p->media[3 + p->nmaxwell + n], nb);
        md->wl[n] = p->visco_wl[n];
      }
    }
  }

  wav_init(gd, &r->wav, 4, visco_type, p->nmaxwell);

  /* bdry_t: bdry_init + what bdry_free_set / bdry_pml_set (forward/bdry_t.c:49-288) store */
  bdry_t *b = &r->bdry;
  bdry_init(b, g->nx, g->ny, g->nz);
  b->is_sides_free[2][1] = p->free_top;
  b->is_enable_free = p->free_top;
  size_t nsl = gd->siz_slice * 9;
  b->matVx2Vz2 = dupf(p->matVx2Vz, nsl);
  b->matVy2Vz2 = dupf(p->matVy2Vz, nsl);
  b->matF2Vz2 = dupf(p->matF2Vz, nsl);
  b->matD = dupf(p->matD, nsl);
  for (int idim = 0; idim < 3; idim++) for (int is = 0; is < 2; is++) {
    const cgfd_pml_face_t *f = &p->pml[idim][is];
    b->is_sides_pml[idim][is] = f->enabled;
    b->num_of_layers[idim][is] = f->enabled ? f->nlay : 0;
    int nl = b->num_of_layers[idim][is];
    b->ni1[idim][is] = g->ni1; b->ni2[idim][is] = g->ni2;
    b->nj1[idim][is] = g->nj1; b->nj2[idim][is] = g->nj2;
    b->nk1[idim][is] = g->nk1; b->nk2[idim][is] = g->nk2;
    if (idim == 0 && is == 0) b->ni2[idim][is] = g->ni1 + nl;
    if (idim == 0 && is == 1) b->ni1[idim][is] = g->ni2 - nl;
    if (idim == 1 && is == 0) b->nj2[idim][is] = g->nj1 + nl;
    if (idim == 1 && is == 1) b->nj1[idim][is] = g->nj2 - nl;
    if (idim == 2 && is == 0) b->nk2[idim][is] = g->nk1 + nl;
    if (idim == 2 && is == 1) b->nk1[idim][is] = g->nk2 - nl;
    if (f->enabled) {
      b->is_enable_pml = 1;
      b->A[idim][is] = dupf(f->A, nl + 1);
      b->B[idim][is] = dupf(f->B, nl + 1);
      b->D[idim][is] = dupf(f->D, nl + 1);
    }
    bdry_pml_auxvar_init(b->ni2[idim][is] - b->ni1[idim][is] + 1, b->nj2[idim][is] - b->nj1[idim][is] + 1,
                         b->nk2[idim][is] - b->nk1[idim][is] + 1, &r->wav, &b->auxvar[idim][is], 0);
    bdrypml_auxvar_t *a = &b->auxvar[idim][is];
    a->pre = a->var; a->cur = a->var;
    if (a->var) { a->tmp = a->var + a->siz_ilevel; a->rhs = a->var + 2 * a->siz_ilevel; a->end = a->var + 3 * a->siz_ilevel; }
  }
  if (p->ablexp_enabled) {
    b->is_enable_ablexp = 1;
    for (int n = 0; n < 6; n++) {
      bdry_block_t *D = &b->bdry_blk[n];
      D->enable = p->ablexp_blk[n][0];
      D->ni1 = p->ablexp_blk[n][1]; D->ni2 = p->ablexp_blk[n][2];
      D->nj1 = p->ablexp_blk[n][3]; D->nj2 = p->ablexp_blk[n][4];
      D->nk1 = p->ablexp_blk[n][5]; D->nk2 = p->ablexp_blk[n][6];
      D->ni = D->ni2 - D->ni1 + 1; D->nj = D->nj2 - D->nj1 + 1; D->nk = D->nk2 - D->nk1 + 1;
    }
    b->ablexp_Ex = dupf(p->ablexp_Ex, g->nx);
    b->ablexp_Ey = dupf(p->ablexp_Ey, g->ny);
    b->ablexp_Ez = dupf(p->ablexp_Ez, g->nz);
  }

  /* src_t */
  src_t *s = &r->src;
  const cgfd_src_t *fs = &p->src;
  s->total_number = fs->total_number;
  s->max_nt = fs->max_nt; s->max_stage = fs->max_stage;
  sprintf(s->evtnm, "evt");
  size_t ns = fs->total_number, ntab = (size_t)fs->total_number * fs->max_nt * fs->max_stage;
  s->si = dupi(fs->si, ns); s->sj = dupi(fs->sj, ns); s->sk = dupi(fs->sk, ns);
  s->si_inc = dupf(fs->si_inc, ns); s->sj_inc = dupf(fs->sj_inc, ns); s->sk_inc = dupf(fs->sk_inc, ns);
  s->it_begin = dupi(fs->it_begin, ns); s->it_end = dupi(fs->it_end, ns);
  s->is_surface_force_strict = fs->is_surface_force_strict;
  s->total_number_surface_force = fs->total_number_surface_force;
  s->force_rate_indx = dupi(fs->force_rate_indx, fs->total_number_surface_force);
  s->itype_spatial_ext = fs->itype_spatial_ext ? fs->itype_spatial_ext : CONST_SRC_SPATIAL_POINT;
  s->ext_half_npoint = fs->ext_half_npoint;
  s->ext_length_npoint = 2 * fs->ext_half_npoint + 1;
  s->ext_size_npoint = s->ext_length_npoint * s->ext_length_npoint * s->ext_length_npoint;
  s->ext_func_coef = fs->ext_func_coef;
  s->force_actived = fs->force_actived; s->moment_actived = fs->moment_actived;
  s->Fx = dupf(fs->Fx, ntab); s->Fy = dupf(fs->Fy, ntab); s->Fz = dupf(fs->Fz, ntab);
  s->Mxx = dupf(fs->Mxx, ntab); s->Myy = dupf(fs->Myy, ntab); s->Mzz = dupf(fs->Mzz, ntab);
  s->Mxz = dupf(fs->Mxz, ntab); s->Myz = dupf(fs->Myz, ntab); s->Mxy = dupf(fs->Mxy, ntab);
  size_t nrate = (size_t)fs->total_number_surface_force * fs->max_nt * fs->max_stage;
  s->Fx_rate = dupf(fs->Fx_rate, nrate); s->Fy_rate = dupf(fs->Fy_rate, nrate); s->Fz_rate = dupf(fs->Fz_rate, nrate);
  s->TxSrc = dupf(NULL, gd->siz_slice); s->TySrc = dupf(NULL, gd->siz_slice); s->TzSrc = dupf(NULL, gd->siz_slice);
  s->VxSrc = dupf(NULL, gd->siz_slice); s->VySrc = dupf(NULL, gd->siz_slice); s->VzSrc = dupf(NULL, gd->siz_slice);
  s->dd_is_valid = 0;

  /* single rank topology (forward/mympi_t.c:15-49 on one rank) + halo buffers */
  mympi_t *m = &r->mympi;
  m->nprocx = 1; m->nprocy = 1; m->myid = 0; m->comm = MPI_COMM_WORLD; m->topocomm = MPI_COMM_WORLD;
  m->topoid[0] = 0; m->topoid[1] = 0;
  for (int n = 0; n < CONST_NDIM_2; n++) m->neighid[n] = MPI_PROC_NULL;
  blk_macdrp_mesg_init(m, &r->fd, gd->ni, gd->nj, gd->nk, r->wav.ncmp);

  return r;
}

int cgfd_ref_ncmp(void *h) { return ((ref_t *)h)->wav.ncmp; }

/* gd_curv_metric_cal (forward/gd_t.c:190-402) on the instance's own gd_t with the coordinates x, y, z: out = jac, xi_x .. zeta_z */
int cgfd_ref_metric_from_coords(void *h, const float *x, const float *y, const float *z, float *out)
{
  ref_t *r = (ref_t *)h;
  gd_t gd = r->gd;
  gdcurv_metric_t M;
  gd.x3d = dupf(x, r->nvol); gd.y3d = dupf(y, r->nvol); gd.z3d = dupf(z, r->nvol);
  gd_curv_metric_init(&gd, &M);
  gd_curv_metric_cal(&gd, &M, r->fd.fdc_len, r->fd.fdc_indx, r->fd.fdc_coef);
  float *src[10] = { M.jac, M.xi_x, M.xi_y, M.xi_z, M.eta_x, M.eta_y, M.eta_z, M.zeta_x, M.zeta_y, M.zeta_z };
  for (int m = 0; m < 10; m++) memcpy(out + (size_t)m * r->nvol, src[m], r->nvol * sizeof(float));
  free(gd.x3d); free(gd.y3d); free(gd.z3d);
  return 0;
}

/* distributed sources the way src_dd_read2local (forward/src_t.c:1181-1927) leaves them behind: points, first time block in
 * memory, the rest of the time functions in two binary files the driver's src_dd_accit_loadstf keeps reading from.
 * vi [nt_total][max_stage][n][3], mij [nt_total][max_stage][n][6] */
int cgfd_ref_set_dd(void *h, int n, const int64_t *indx, int vi_act, int mij_act, int nt_total, int nt_per_read,
                    const float *vi, const float *mij, const char *prefix)
{
  ref_t *r = (ref_t *)h;
  src_t *s = &r->src;
  char fn[1024];
  if (s->max_stage <= 0) s->max_stage = 4;
  s->dd_is_valid = 1; s->dd_is_add_at_point = 1; s->dd_smo_hlen = 0;
  s->dd_vi_actived = vi_act; s->dd_mij_actived = mij_act;
  s->dd_total_number = n; s->dd_max_nt = nt_total; s->dd_nt_per_read = nt_per_read;
  s->dd_nt_this_read = nt_per_read < nt_total ? nt_per_read : nt_total; s->dd_it_here = 0;
  s->dd_indx = (size_t *)malloc(sizeof(size_t) * n);
  for (int q = 0; q < n; q++) s->dd_indx[q] = (size_t)indx[q];
  size_t per_step3 = (size_t)s->max_stage * n * 3, per_step6 = (size_t)s->max_stage * n * 6;
  if (vi_act) {
    snprintf(fn, sizeof(fn), "%s.vi", prefix);
    FILE *f = fopen(fn, "wb"); if (!f) return 1;
    fwrite(vi, sizeof(float), per_step3 * nt_total, f); fclose(f);
    s->fp_vi = fopen(fn, "rb"); if (!s->fp_vi) return 1;
    s->dd_vi = (float *)calloc(per_step3 * nt_per_read, sizeof(float));
    if (fread(s->dd_vi, sizeof(float), per_step3 * s->dd_nt_this_read, s->fp_vi) != per_step3 * s->dd_nt_this_read) return 1;
  }
  if (mij_act) {
    snprintf(fn, sizeof(fn), "%s.mij", prefix);
    FILE *f = fopen(fn, "wb"); if (!f) return 1;
    fwrite(mij, sizeof(float), per_step6 * nt_total, f); fclose(f);
    s->fp_mij = fopen(fn, "rb"); if (!s->fp_mij) return 1;
    s->dd_mij = (float *)calloc(per_step6 * nt_per_read, sizeof(float));
    if (fread(s->dd_mij, sizeof(float), per_step6 * s->dd_nt_this_read, s->fp_mij) != per_step6 * s->dd_nt_this_read) return 1;
  }
  return 0;
}

/* grid coordinates [nz][ny][nx] (gd->x3d/y3d/z3d, forward/gd_t.c:23-95): only sv_curv_col_vis_iso_dvh2dvz reads them
 * (surface tangents for matD, forward/sv_curv_col_vis_iso.c:375-507); the C ABI problem does not carry coordinates */
int cgfd_ref_set_coords(void *h, const float *x, const float *y, const float *z)
{
  ref_t *r = (ref_t *)h;
  r->gd.x3d = dupf(x, r->nvol); r->gd.y3d = dupf(y, r->nvol); r->gd.z3d = dupf(z, r->nvol);
  return 0;
}

size_t cgfd_ref_pml_aux_size(void *h, int idim, int iside)
{
  return ((ref_t *)h)->bdry.auxvar[idim][iside].siz_ilevel;
}
int cgfd_ref_set_pml_aux(void *h, int idim, int iside, const float *aux)
{
  bdrypml_auxvar_t *a = &((ref_t *)h)->bdry.auxvar[idim][iside];
  if (!a->var) return 1;
  memcpy(a->var, aux, a->siz_ilevel * sizeof(float));
  return 0;
}
int cgfd_ref_get_pml_aux(void *h, int idim, int iside, int level, float *aux)
{
  bdrypml_auxvar_t *a = &((ref_t *)h)->bdry.auxvar[idim][iside];
  if (!a->var) return 1;
  memcpy(aux, a->var + (size_t)level * a->siz_ilevel, a->siz_ilevel * sizeof(float));
  return 0;
}

/* free-surface matrices as the reference computes them (drv_rk_curv_col.c:132-159) */
int cgfd_ref_dvh2dvz(void *h, float *matVx2Vz, float *matVy2Vz, float *matF2Vz, float *matD)
{
  ref_t *r = (ref_t *)h;
  size_t nsl = r->gd.siz_slice * 9 * sizeof(float);
  switch (r->md.medium_type) {
    case CONST_MEDIUM_ELASTIC_ISO: sv_curv_col_el_iso_dvh2dvz(&r->gd, &r->metric, &r->md, &r->bdry, 0); break;
    case CONST_MEDIUM_ELASTIC_VTI: sv_curv_col_el_vti_dvh2dvz(&r->gd, &r->metric, &r->md, &r->bdry, 0); break;
    case CONST_MEDIUM_ELASTIC_ANISO: sv_curv_col_el_aniso_dvh2dvz(&r->gd, &r->metric, &r->md, &r->bdry, 0); break;
    case CONST_MEDIUM_VISCOELASTIC_ISO:
      sv_curv_col_vis_iso_dvh2dvz(&r->gd, &r->metric, &r->md, &r->bdry, r->fd.fdc_len, r->fd.fdc_indx, r->fd.fdc_coef, 0);
      break;
    default: return 1;
  }
  if (matVx2Vz) memcpy(matVx2Vz, r->bdry.matVx2Vz2, nsl);
  if (matVy2Vz) memcpy(matVy2Vz, r->bdry.matVy2Vz2, nsl);
  if (matF2Vz) memcpy(matF2Vz, r->bdry.matF2Vz2, nsl);
  if (matD) memcpy(matD, r->bdry.matD, nsl);
  return 0;
}

/* one RHS evaluation with the reference's own *_onestage; aux "cur" = level 0, aux rhs = level 2 */
int cgfd_ref_onestage(void *h, int it, int ipair, int istage, const float *w_cur, float *rhs)
{
  ref_t *r = (ref_t *)h;
  size_t n = r->wav.siz_ilevel;
  float *cur = r->wav.v5d;
  float *out = r->wav.v5d + 2 * n;
  memcpy(cur, w_cur, n * sizeof(float));
  memset(out, 0, n * sizeof(float));
  for (int idim = 0; idim < 3; idim++) for (int is = 0; is < 2; is++) {
    bdrypml_auxvar_t *a = &r->bdry.auxvar[idim][is];
    a->cur = a->var;
    if (a->var) a->rhs = a->var + 2 * a->siz_ilevel;
  }
  fd_t *fd = &r->fd;
  src_set_time(&r->src, it, istage);
  src_set_surface_layer_for_force(&r->src, &r->gd, &r->metric);
#define ARGS cur, out, &r->wav, &r->gd, &r->metric, &r->md, &r->bdry, &r->src, \
    fd->num_of_fdx_op, fd->pair_fdx_op[ipair][istage], fd->num_of_fdy_op, fd->pair_fdy_op[ipair][istage], \
    fd->num_of_fdz_op, fd->pair_fdz_op[ipair][istage], fd->fdz_max_len, 0, 0
  switch (r->md.medium_type) {
    case CONST_MEDIUM_ELASTIC_ISO: sv_curv_col_el_iso_onestage(ARGS); break;
    case CONST_MEDIUM_ELASTIC_VTI: sv_curv_col_el_vti_onestage(ARGS); break;
    case CONST_MEDIUM_ELASTIC_ANISO: sv_curv_col_el_aniso_onestage(ARGS); break;
    case CONST_MEDIUM_VISCOELASTIC_ISO: sv_curv_col_vis_iso_onestage(ARGS); break;
    default: return 1;
  }
#undef ARGS
  memcpy(rhs, out, n * sizeof(float));
  return 0;
}

/*
 * nsteps RK4 steps with the reference's own driver, starting at it = 0 from wavefield w (level n,
 * in/out) and the PML aux vars stored by cgfd_ref_set_pml_aux (default 0). nrec record points are
 * sampled after every step through the reference's io_line_keep; rec[(it*ncmp+icmp)*nrec + ip].
 * outdir receives the PG_V_A_D file. Returns the wall seconds spent inside drv_rk_curv_col_allstep
 * through *seconds.
 */
#include <time.h>
int cgfd_ref_run(void *h, int nsteps, float *w, int nrec, const int64_t *rec_iptr, float *rec,
                 const char *outdir, double *seconds)
{
  ref_t *r = (ref_t *)h;
  size_t n = r->wav.siz_ilevel;
  int ncmp = r->wav.ncmp;
  memset(r->wav.v5d, 0, 4 * n * sizeof(float));
  memcpy(r->wav.v5d, w, n * sizeof(float));
  for (int idim = 0; idim < 3; idim++) for (int is = 0; is < 2; is++) {
    bdrypml_auxvar_t *a = &r->bdry.auxvar[idim][is];
    if (a->var) memset(a->var + a->siz_ilevel, 0, 3 * a->siz_ilevel * sizeof(float));
  }
  memset(&r->iorecv, 0, sizeof(r->iorecv));
  memset(&r->ioline, 0, sizeof(r->ioline));
  memset(&r->ioslice, 0, sizeof(r->ioslice));
  memset(&r->iosnap, 0, sizeof(r->iosnap));
  r->iorecv.max_nt = nsteps; r->iorecv.ncmp = ncmp;
  ioline_t *L = &r->ioline;
  L->max_nt = nsteps; L->ncmp = ncmp;
  int line_nr = nrec;
  int *iptr32 = NULL; float *seis = NULL;
  int *pl_nr = &line_nr; int *p_iptr[1]; float *p_seis[1];
  if (nrec > 0) {
    L->num_of_lines = 1;
    iptr32 = (int *)malloc(nrec * sizeof(int));
    for (int i = 0; i < nrec; i++) iptr32[i] = (int)rec_iptr[i];
    seis = (float *)calloc((size_t)nrec * ncmp * nsteps, sizeof(float));
    p_iptr[0] = iptr32; p_seis[0] = seis;
    L->line_nr = pl_nr; L->recv_iptr = p_iptr; L->recv_seismo = p_seis;
  }
  mkdir(outdir, 0777);
  char part[64] = "px0_py0";
  struct timespec t0, t1;
  clock_gettime(CLOCK_MONOTONIC, &t0);
  drv_rk_curv_col_allstep(&r->fd, &r->gd, &r->metric, &r->md, &r->src, &r->bdry, &r->wav, &r->mympi,
                          &r->iorecv, &r->ioline, &r->ioslice, &r->iosnap,
                          r->p.dt, nsteps, 0.0f, part, (char *)outdir, 0, 0, 0);
  clock_gettime(CLOCK_MONOTONIC, &t1);
  if (seconds) *seconds = (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec);
  /* after an odd number of steps the new level sits in buffer 3 (swap at drv_rk_curv_col.c:530) */
  int lev = (nsteps % 2) ? 3 : 0;
  memcpy(w, r->wav.v5d + (size_t)lev * n, n * sizeof(float));
  for (int idim = 0; idim < 3; idim++) for (int is = 0; is < 2; is++) {
    bdrypml_auxvar_t *a = &r->bdry.auxvar[idim][is];
    if (a->var && lev == 3) memcpy(a->var, a->var + 3 * a->siz_ilevel, a->siz_ilevel * sizeof(float));
    if (a->var) { a->pre = a->var; a->cur = a->var; a->tmp = a->var + a->siz_ilevel;
                  a->rhs = a->var + 2 * a->siz_ilevel; a->end = a->var + 3 * a->siz_ilevel; }
  }
  if (nrec > 0) {
    for (int ip = 0; ip < nrec; ip++)
      for (int ic = 0; ic < ncmp; ic++)
        for (int it = 0; it < nsteps; it++)
          rec[((size_t)it * ncmp + ic) * nrec + ip] = seis[(size_t)ip * nsteps * ncmp + (size_t)ic * nsteps + it];
    free(iptr32); free(seis);
  }
  return 0;
}
