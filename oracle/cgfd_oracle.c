/*
 * cgfd_oracle.c -- CPU restatement of the CGFD3D time-stepping hot path (isotropic elastic medium).
 *
 * TEST INFRASTRUCTURE, NOT PRODUCT. Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline
 * leg may load this. It is a plain, unfused, single-threaded C restatement of what the reference does
 * per RK stage, written against the flat problem description of include/cgfd3d_b200.h:
 *
 *   rhs_inner      forward/sv_curv_col_el_iso.c:208-441   27 one-sided 5-point derivatives, momentum, Hooke
 *   rhs_timg       forward/sv_curv_col_el.c:30-305        traction image, conservative momentum RHS (ZERO guard,
 *                                                          or MIRROR: SURVEY.md §8c hazard 1)
 *   rhs_vlow       forward/sv_curv_col_el_iso.c:451-634   stress RHS in the top 3 rows
 *   rhs_cfspml     forward/sv_curv_col_el_iso.c:644-1146  ADE CFS-PML, 6 faces
 *   rhs_src        forward/sv_curv_col_el.c:311-479       point / Gaussian force and moment sources
 *   surface force  forward/src_t.c:153-314
 *   rhs_srcdd      forward/sv_curv_col_el.c:486-632       distributed (finite-fault) sources, added at the point
 *   graves_Qs      forward/sv_curv_col_el.c:638-666       constant-Q attenuation of w_end after the last stage
 *   stage loop     forward/drv_rk_curv_col.c:167-544      RK4 axpy for wavefield and PML aux, level swap
 *
 * Parity of this restatement is PINNED against the reference itself: tests/test_cpu_oracle.py compares
 * it with oracle/_ref (the unmodified reference compiled from /root/reference) function by function and
 * over multi-step runs. Compiled with -ffp-contract=off so that, like the reference build (gcc -O3 on
 * x86-64 without -march), no FMA contraction happens.
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "../include/cgfd3d_b200.h"

enum { VX, VY, VZ, TXX, TYY, TZZ, TYZ, TXZ, TXY };
enum { JAC, XIX, XIY, XIZ, ETX, ETY, ETZ, ZTX, ZTY, ZTZ };

typedef struct {
  int on, nlay, r[6];
  size_t siz;
  float *A, *B, *D;
  float *lev[4]; /* pre, tmp, rhs, end */
} face_t;

typedef struct {
  cgfd_problem_t p;
  int nx, ny, nz;
  long L, S;
  size_t V;
  float *metric[10], *lam, *mu, *slw;
  float *lev[4]; /* pre, tmp, rhs, end */
  face_t f[3][2];
  float *mvx, *mvy, *mf;
  float *srcsl[6]; /* TxSrc TySrc TzSrc VxSrc VySrc VzSrc */
  /* Graves' attenuation: Qs array (NULL = off), forward/sv_curv_col_el.c:638-666 */
  float *Qs; float Qs_freq;
  /* distributed sources (forward/src_t.h:94-126): whole time functions in memory */
  int dd_n, dd_nt; size_t *dd_indx; float *dd_vi, *dd_mij;
} orc_t;

static float *dupf(const float *s, size_t n)
{
  float *d = (float *)calloc(n ? n : 1, sizeof(float));
  if (s && n) memcpy(d, s, n * sizeof(float));
  return d;
}

void *cgfd_oracle_create(const cgfd_problem_t *p)
{
  if (p->abi_version != CGFD_ABI_VERSION || p->medium_type != CGFD_MEDIUM_ELASTIC_ISO) return NULL;
  orc_t *o = (orc_t *)calloc(1, sizeof(orc_t));
  o->p = *p;
  o->nx = p->grid.nx; o->ny = p->grid.ny; o->nz = p->grid.nz;
  o->L = o->nx; o->S = (long)o->nx * o->ny; o->V = (size_t)o->S * o->nz;
  for (int m = 0; m < 10; m++) o->metric[m] = dupf(p->metric[m], o->V);
  o->lam = dupf(p->media[0], o->V); o->mu = dupf(p->media[1], o->V); o->slw = dupf(p->media[2], o->V);
  if (p->graves_Qs) { o->Qs = dupf(p->graves_Qs, o->V); o->Qs_freq = p->graves_Qs_freq; }
  for (int l = 0; l < 4; l++) o->lev[l] = dupf(NULL, o->V * 9);
  const cgfd_grid_t *g = &p->grid;
  for (int d = 0; d < 3; d++) for (int s = 0; s < 2; s++) {
    face_t *f = &o->f[d][s];
    const cgfd_pml_face_t *pf = &p->pml[d][s];
    f->on = pf->enabled; f->nlay = pf->enabled ? pf->nlay : 0;
    int r[6] = { g->ni1, g->ni2, g->nj1, g->nj2, g->nk1, g->nk2 };
    if (s == 0) r[2 * d + 1] = r[2 * d] + f->nlay; else r[2 * d] = r[2 * d + 1] - f->nlay;
    memcpy(f->r, r, sizeof(r));
    if (!f->on) continue;
    f->siz = (size_t)(r[1] - r[0] + 1) * (r[3] - r[2] + 1) * (r[5] - r[4] + 1);
    f->A = dupf(pf->A, f->nlay + 1); f->B = dupf(pf->B, f->nlay + 1); f->D = dupf(pf->D, f->nlay + 1);
    for (int l = 0; l < 4; l++) f->lev[l] = dupf(NULL, f->siz * 9);
  }
  if (p->free_top) {
    o->mvx = dupf(p->matVx2Vz, (size_t)o->S * 9); o->mvy = dupf(p->matVy2Vz, (size_t)o->S * 9); o->mf = dupf(p->matF2Vz, (size_t)o->S * 9);
  }
  for (int n = 0; n < 6; n++) o->srcsl[n] = dupf(NULL, (size_t)o->S);
  /* deep copies of the source tables */
  cgfd_src_t *s = &o->p.src;
  const cgfd_src_t *q = &p->src;
  size_t ns = q->total_number, nt = (size_t)q->total_number * q->max_nt * q->max_stage;
  size_t nr = (size_t)q->total_number_surface_force * q->max_nt * q->max_stage;
#define DUPI(name, n) do { int32_t *d_ = (int32_t *)calloc((n) ? (n) : 1, 4); if (q->name) memcpy(d_, q->name, (n) * 4); s->name = d_; } while (0)
  DUPI(si, ns); DUPI(sj, ns); DUPI(sk, ns); DUPI(it_begin, ns); DUPI(it_end, ns); DUPI(force_rate_indx, (size_t)q->total_number_surface_force);
  s->si_inc = dupf(q->si_inc, ns); s->sj_inc = dupf(q->sj_inc, ns); s->sk_inc = dupf(q->sk_inc, ns);
  s->Fx = dupf(q->Fx, nt); s->Fy = dupf(q->Fy, nt); s->Fz = dupf(q->Fz, nt);
  s->Mxx = dupf(q->Mxx, nt); s->Myy = dupf(q->Myy, nt); s->Mzz = dupf(q->Mzz, nt);
  s->Mxz = dupf(q->Mxz, nt); s->Myz = dupf(q->Myz, nt); s->Mxy = dupf(q->Mxy, nt);
  s->Fx_rate = dupf(q->Fx_rate, nr); s->Fy_rate = dupf(q->Fy_rate, nr); s->Fz_rate = dupf(q->Fz_rate, nr);
  if (p->ablexp_enabled) {
    o->p.ablexp_Ex = dupf(p->ablexp_Ex, o->nx); o->p.ablexp_Ey = dupf(p->ablexp_Ey, o->ny); o->p.ablexp_Ez = dupf(p->ablexp_Ez, o->nz);
  }
  return o;
}

/* distributed sources: vi [nt][4][n][3] and / or mij [nt][4][n][6] (NULL = not active); sv_curv_col_el_rhs_srcdd,
 * forward/sv_curv_col_el.c:486-632 (steps >= nt get no dd source, like dd_is_valid = 0 past dd_max_nt) */
int cgfd_oracle_set_dd(void *h, int n, const int64_t *indx, int nt, const float *vi, const float *mij)
{
  orc_t *o = h;
  o->dd_n = n; o->dd_nt = nt;
  o->dd_indx = malloc(sizeof(size_t) * n);
  for (int q = 0; q < n; q++) o->dd_indx[q] = (size_t)indx[q];
  o->dd_vi = vi ? dupf(vi, (size_t)nt * 4 * n * 3) : NULL;
  o->dd_mij = mij ? dupf(mij, (size_t)nt * 4 * n * 6) : NULL;
  return 0;
}

size_t cgfd_oracle_pml_aux_size(void *h, int d, int s) { orc_t *o = h; return o->f[d][s].on ? o->f[d][s].siz * 9 : 0; }
int cgfd_oracle_set_pml_aux(void *h, int d, int s, const float *a)
{
  orc_t *o = h; if (!o->f[d][s].on) return 1;
  memcpy(o->f[d][s].lev[0], a, o->f[d][s].siz * 9 * sizeof(float)); return 0;
}
int cgfd_oracle_get_pml_aux(void *h, int d, int s, int level, float *a)
{
  orc_t *o = h; if (!o->f[d][s].on) return 1;
  memcpy(a, o->f[d][s].lev[level], o->f[d][s].siz * 9 * sizeof(float)); return 0;
}

/* five-term one-sided difference, summed left to right (M_FD_SHIFT_PTR_MACDRP, forward/fd_t.h:26-31) */
static inline float d5(const float *p, long stride, int first, const float *c)
{
  const float *q = p + first * stride;
  return c[0] * q[0] + c[1] * q[stride] + c[2] * q[2 * stride] + c[3] * q[3 * stride] + c[4] * q[4 * stride];
}

/* Hooke's law in curvilinear coordinates from the 9 velocity derivatives dv[axis][cmp]
 * (forward/sv_curv_col_el_iso.c:409-435) */
static inline void hooke(const float e[3][3], float dv[3][3], float lam, float mu, float lam2mu, float *hT /* [9] */)
{
  float gxx = e[0][0] * dv[0][0] + e[1][0] * dv[1][0] + e[2][0] * dv[2][0];
  float gyy = e[0][1] * dv[0][1] + e[1][1] * dv[1][1] + e[2][1] * dv[2][1];
  float gzz = e[0][2] * dv[0][2] + e[1][2] * dv[1][2] + e[2][2] * dv[2][2];
  hT[TXX] = lam2mu * gxx + lam * (e[0][1] * dv[0][1] + e[1][1] * dv[1][1] + e[2][1] * dv[2][1]
                                + e[0][2] * dv[0][2] + e[1][2] * dv[1][2] + e[2][2] * dv[2][2]);
  hT[TYY] = lam2mu * gyy + lam * (e[0][0] * dv[0][0] + e[1][0] * dv[1][0] + e[2][0] * dv[2][0]
                                + e[0][2] * dv[0][2] + e[1][2] * dv[1][2] + e[2][2] * dv[2][2]);
  hT[TZZ] = lam2mu * gzz + lam * (e[0][0] * dv[0][0] + e[1][0] * dv[1][0] + e[2][0] * dv[2][0]
                                + e[0][1] * dv[0][1] + e[1][1] * dv[1][1] + e[2][1] * dv[2][1]);
  hT[TXY] = mu * (e[0][1] * dv[0][0] + e[0][0] * dv[0][1] + e[1][1] * dv[1][0] + e[1][0] * dv[1][1] + e[2][1] * dv[2][0] + e[2][0] * dv[2][1]);
  hT[TXZ] = mu * (e[0][2] * dv[0][0] + e[0][0] * dv[0][2] + e[1][2] * dv[1][0] + e[1][0] * dv[1][2] + e[2][2] * dv[2][0] + e[2][0] * dv[2][2]);
  hT[TYZ] = mu * (e[0][2] * dv[0][1] + e[0][1] * dv[0][2] + e[1][2] * dv[1][1] + e[1][1] * dv[1][2] + e[2][2] * dv[2][1] + e[2][1] * dv[2][2]);
}

static inline void load_metric(const orc_t *o, size_t p, float e[3][3])
{
  for (int a = 0; a < 3; a++) for (int c = 0; c < 3; c++) e[a][c] = o->metric[1 + 3 * a + c][p];
}

typedef struct { int first[3]; const float *c[3]; int dir[3]; } ops_t;

static void rhs_inner(const orc_t *o, const float *w, float *h, const ops_t *op)
{
  const cgfd_grid_t *g = &o->p.grid;
  const long str[3] = { 1, o->L, o->S };
  /* stress pairs entering each momentum equation: row v, column = direction of the metric vector */
  static const int T[3][3] = { { TXX, TXY, TXZ }, { TXY, TYY, TYZ }, { TXZ, TYZ, TZZ } };
  for (int k = g->nk1; k <= g->nk2; k++) for (int j = g->nj1; j <= g->nj2; j++) for (int i = g->ni1; i <= g->ni2; i++) {
    size_t p = (size_t)k * o->S + (size_t)j * o->L + i;
    float e[3][3]; load_metric(o, p, e);
    float lam = o->lam[p], mu = o->mu[p], slw = o->slw[p], lam2mu = lam + 2.0 * mu;
    float dF[3][9];
    for (int a = 0; a < 3; a++) for (int c = 0; c < 9; c++) dF[a][c] = d5(w + c * o->V + p, str[a], op->first[a], op->c[a]);
    for (int v = 0; v < 3; v++) {
      h[v * o->V + p] = slw * (e[0][0] * dF[0][T[v][0]] + e[0][1] * dF[0][T[v][1]] + e[0][2] * dF[0][T[v][2]]
                             + e[1][0] * dF[1][T[v][0]] + e[1][1] * dF[1][T[v][1]] + e[1][2] * dF[1][T[v][2]]
                             + e[2][0] * dF[2][T[v][0]] + e[2][1] * dF[2][T[v][1]] + e[2][2] * dF[2][T[v][2]]);
    }
    float dv[3][3], hT[9];
    for (int a = 0; a < 3; a++) for (int c = 0; c < 3; c++) dv[a][c] = dF[a][c];
    hooke(e, dv, lam, mu, lam2mu, hT);
    for (int c = 3; c < 9; c++) h[c * o->V + p] = hT[c];
  }
}

/* traction image: momentum RHS in conservative form in the rows whose zeta stencil crosses the surface */
static void rhs_timg(const orc_t *o, const float *w, float *h, const ops_t *op)
{
  const cgfd_grid_t *g = &o->p.grid;
  const long str[3] = { 1, o->L, o->S };
  static const int T[3][3] = { { TXX, TXY, TXZ }, { TXY, TYY, TYZ }, { TXZ, TYZ, TZZ } };
  const int kmin = g->nk2 - (op->first[2] + 4);
  for (int k = kmin; k <= g->nk2; k++) {
    const int n_free = g->nk2 - k - op->first[2];
    for (int j = g->nj1; j <= g->nj2; j++) for (int i = g->ni1; i <= g->ni2; i++) {
      size_t p = (size_t)k * o->S + (size_t)j * o->L + i, p2 = (size_t)j * o->L + i;
      float slwjac = o->slw[p] / o->metric[JAC][p];
      for (int v = 0; v < 3; v++) {
        float D[3];
        for (int a = 0; a < 3; a++) {
          float vec[5];
          for (int n = 0; n < 5; n++) {
            if (a == 2 && n >= n_free) break;
            size_t q = p + (long)(op->first[a] + n) * str[a];
            vec[n] = o->metric[JAC][q] * (o->metric[1 + 3 * a][q] * w[T[v][0] * o->V + q]
                                        + o->metric[2 + 3 * a][q] * w[T[v][1] * o->V + q]
                                        + o->metric[3 + 3 * a][q] * w[T[v][2] * o->V + q]);
          }
          if (a == 2) {
            float ts = o->srcsl[v][p2];
            vec[n_free] = ts;
            for (int n = n_free + 1; n < 5; n++) {
              int im = n_free - (n - n_free);
              float below;
              if (im >= 0) below = vec[im];
              else if (o->p.timg_mode == CGFD_TIMG_ZERO) below = 0.0f;
              else {
                size_t q = p + (long)(op->first[2] + n - 2 * (n - n_free)) * o->S;
                below = o->metric[JAC][q] * (o->metric[ZTX][q] * w[T[v][0] * o->V + q] + o->metric[ZTY][q] * w[T[v][1] * o->V + q]
                                           + o->metric[ZTZ][q] * w[T[v][2] * o->V + q]);
              }
              vec[n] = 2.0 * ts - below;
            }
          }
          float acc = op->c[a][0] * vec[0];
          for (int n = 1; n < 5; n++) acc += op->c[a][n] * vec[n];
          D[a] = acc;
        }
        h[v * o->V + p] = (D[0] + D[1] + D[2]) * slwjac;
      }
    }
  }
}

/* stress RHS of the top three rows: Dz of the velocity by matrix / 2-point / 3-point operator */
static void rhs_vlow(const orc_t *o, const float *w, float *h, const ops_t *op)
{
  const cgfd_grid_t *g = &o->p.grid;
  const cgfd_fd_t *fd = &o->p.fd;
  const int dz = op->dir[2];
  for (int n = 0; n < 3; n++) {
    int k = g->nk2 - n;
    for (int j = g->nj1; j <= g->nj2; j++) for (int i = g->ni1; i <= g->ni2; i++) {
      size_t p = (size_t)k * o->S + (size_t)j * o->L + i, p2 = (size_t)j * o->L + i;
      float e[3][3]; load_metric(o, p, e);
      float lam = o->lam[p], mu = o->mu[p], lam2mu = lam + 2.0 * mu;
      float dv[3][3];
      for (int c = 0; c < 3; c++) {
        dv[0][c] = d5(w + c * o->V + p, 1, op->first[0], op->c[0]);
        dv[1][c] = d5(w + c * o->V + p, o->L, op->first[1], op->c[1]);
      }
      if (n == 0) {
        const float *A = o->mvx + p2 * 9, *B = o->mvy + p2 * 9, *F = o->mf + p2 * 9;
        for (int r = 0; r < 3; r++) {
          float v = A[3 * r] * dv[0][0] + A[3 * r + 1] * dv[0][1] + A[3 * r + 2] * dv[0][2]
                  + B[3 * r] * dv[1][0] + B[3 * r + 1] * dv[1][1] + B[3 * r + 2] * dv[1][2];
          v += F[3 * r] * o->srcsl[3][p2] + F[3 * r + 1] * o->srcsl[4][p2] + F[3 * r + 2] * o->srcsl[5][p2];
          dv[2][r] = v;
        }
      } else {
        for (int c = 0; c < 3; c++) {
          float acc = fd->lay_coef[n][dz][0] * w[c * o->V + p + fd->lay_indx[n][dz][0] * o->S];
          for (int m = 1; m < fd->lay_len[n][dz]; m++) acc += fd->lay_coef[n][dz][m] * w[c * o->V + p + fd->lay_indx[n][dz][m] * o->S];
          dv[2][c] = acc;
        }
      }
      float hT[9];
      hooke(e, dv, lam, mu, lam2mu, hT);
      for (int c = 3; c < 9; c++) h[c * o->V + p] = hT[c];
    }
  }
}

/* ADE CFS-PML of every enabled face: h += (B-1)*rhs_n - B*aux ; aux_rhs = D*rhs_n - A*aux */
static void rhs_cfspml(orc_t *o, const float *w, float *h, const ops_t *op, int cur_level)
{
  const cgfd_grid_t *g = &o->p.grid;
  const long str[3] = { 1, o->L, o->S };
  for (int d = 0; d < 3; d++) for (int s = 0; s < 2; s++) {
    face_t *f = &o->f[d][s];
    if (!f->on) continue;
    const float *aux = f->lev[cur_level];
    float *arhs = f->lev[2];
    size_t pa = 0;
    for (int k = f->r[4]; k <= f->r[5]; k++) for (int j = f->r[2]; j <= f->r[3]; j++) for (int i = f->r[0]; i <= f->r[1]; i++, pa++) {
      int ia = (d == 0) ? i - f->r[0] : (d == 1) ? j - f->r[2] : k - f->r[4];
      float cA = f->A[ia], cB = f->B[ia], cD = f->D[ia], cB1 = cB - 1.0;
      size_t p = (size_t)k * o->S + (size_t)j * o->L + i;
      float e1 = o->metric[1 + 3 * d][p], e2 = o->metric[2 + 3 * d][p], e3 = o->metric[3 + 3 * d][p];
      float lam = o->lam[p], mu = o->mu[p], slw = o->slw[p], lam2mu = lam + 2.0 * mu;
      float D[9], r[9];
      for (int c = 0; c < 9; c++) D[c] = d5(w + c * o->V + p, str[d], op->first[d], op->c[d]);
      r[VX] = slw * (e1 * D[TXX] + e2 * D[TXY] + e3 * D[TXZ]);
      r[VY] = slw * (e1 * D[TXY] + e2 * D[TYY] + e3 * D[TYZ]);
      r[VZ] = slw * (e1 * D[TXZ] + e2 * D[TYZ] + e3 * D[TZZ]);
      r[TXX] = lam2mu * e1 * D[VX] + lam * e2 * D[VY] + lam * e3 * D[VZ];
      r[TYY] = lam * e1 * D[VX] + lam2mu * e2 * D[VY] + lam * e3 * D[VZ];
      r[TZZ] = lam * e1 * D[VX] + lam * e2 * D[VY] + lam2mu * e3 * D[VZ];
      r[TXY] = mu * (e2 * D[VX] + e1 * D[VY]);
      r[TXZ] = mu * (e3 * D[VX] + e1 * D[VZ]);
      r[TYZ] = mu * (e3 * D[VY] + e2 * D[VZ]);
      for (int c = 0; c < 9; c++) {
        float a = aux[c * f->siz + pa];
        h[c * o->V + p] += cB1 * r[c] - cB * a;
        arhs[c * f->siz + pa] = cD * r[c] - cA * a;
      }
      if (d < 2 && o->p.free_top && k == g->nk2) {
        /* terms that reach the stress RHS through the free-surface Dz conversion (iso.c:841-901, 989-1048) */
        const float *M = (d == 0 ? o->mvx : o->mvy) + ((size_t)j * o->L + i) * 9;
        float z[3];
        for (int q = 0; q < 3; q++) z[q] = M[3 * q] * D[VX] + M[3 * q + 1] * D[VY] + M[3 * q + 2] * D[VZ];
        float ztx = o->metric[ZTX][p], zty = o->metric[ZTY][p], ztz = o->metric[ZTZ][p];
        float t[9];
        t[TXX] = lam2mu * (ztx * z[0]) + lam * (zty * z[1] + ztz * z[2]);
        t[TYY] = lam2mu * (zty * z[1]) + lam * (ztx * z[0] + ztz * z[2]);
        t[TZZ] = lam2mu * (ztz * z[2]) + lam * (ztx * z[0] + zty * z[1]);
        t[TXY] = mu * (zty * z[0] + ztx * z[1]);
        t[TXZ] = mu * (ztz * z[0] + ztx * z[2]);
        t[TYZ] = mu * (ztz * z[1] + zty * z[2]);
        for (int c = 3; c < 9; c++) {
          h[c * o->V + p] += (cB - 1.0) * t[c];
          arhs[c * f->siz + pa] += cD * t[c];
        }
      }
    }
  }
}

static float fgauss(float t, float a, float t0) { float f = exp(-(t - t0) * (t - t0) / (a * a)) / (sqrtf(M_PI) * a); return f; }

/* normalised Gaussian footprint, normalisation over the rows at or below the surface (src_t.c:2110-2151) */
static void footprint(float *delt, float x0, float y0, float z0, float r, int H, int Hz2)
{
  int n1 = 2 * H + 1, ip = 0;
  for (int k = -H; k <= H; k++) for (int j = -H; j <= H; j++) for (int i = -H; i <= H; i++)
    delt[ip++] = fgauss(i - x0, r, 0.0) * fgauss(j - y0, r, 0.0) * fgauss(k - z0, r, 0.0);
  float sum = 0.0f; ip = 0;
  for (int k = -H; k <= Hz2; k++) for (int j = -H; j <= H; j++) for (int i = -H; i <= H; i++) sum += delt[ip++];
  for (ip = 0; ip < n1 * n1 * n1; ip++) delt[ip] /= sum;
}

static void set_surface_force(orc_t *o, int it, int istage)
{
  const cgfd_src_t *s = &o->p.src;
  const cgfd_grid_t *g = &o->p.grid;
  if (s->total_number_surface_force <= 0) return;
  for (int n = 0; n < 6; n++) memset(o->srcsl[n], 0, (size_t)o->S * sizeof(float));
  int H = s->ext_half_npoint, n1 = 2 * H + 1;
  float *ext = (float *)malloc(sizeof(float) * n1 * n1 * n1);
  for (int n = 0; n < s->total_number_surface_force; n++) {
    int is = s->force_rate_indx[n];
    if (it < s->it_begin[is] || it > s->it_end[is]) continue;
    size_t tab = ((size_t)is * s->max_nt + (it - s->it_begin[is])) * s->max_stage + istage;
    size_t tabr = ((size_t)n * s->max_nt + (it - s->it_begin[is])) * s->max_stage + istage;
    float F[3] = { s->Fx[tab], s->Fy[tab], s->Fz[tab] }, Fr[3] = { s->Fx_rate[tabr], s->Fy_rate[tabr], s->Fz_rate[tabr] };
    int si = s->si[is], sj = s->sj[is], sk = s->sk[is];
    if (s->itype_spatial_ext == CGFD_SRC_SPATIAL_POINT) {
      size_t p = si + sj * o->L + sk * o->S, p2 = si + sj * o->L;
      for (int c = 0; c < 3; c++) { o->srcsl[c][p2] += F[c]; o->srcsl[3 + c][p2] += Fr[c] / o->metric[JAC][p]; }
    } else {
      int kext = g->nk2 - sk;
      footprint(ext, s->si_inc[is], s->sj_inc[is], s->sk_inc[is], s->ext_func_coef, H, kext);
      int ie = (kext + H) * n1 * n1, k = sk + kext;
      for (int je = -H; je <= H; je++) for (int ix = -H; ix <= H; ix++, ie++) {
        int i = si + ix, j = sj + je;
        size_t p = i + j * o->L + k * o->S, p2 = i + j * o->L;
        for (int c = 0; c < 3; c++) { o->srcsl[c][p2] += F[c] * ext[ie]; o->srcsl[3 + c][p2] += Fr[c] * ext[ie] / o->metric[JAC][p]; }
      }
    }
  }
  free(ext);
}

static void rhs_src(const orc_t *o, float *h, int it, int istage)
{
  const cgfd_src_t *s = &o->p.src;
  const cgfd_grid_t *g = &o->p.grid;
  int H = s->ext_half_npoint, n1 = 2 * H + 1;
  float *ext = (float *)malloc(sizeof(float) * (n1 * n1 * n1 + 1));
  static const int MC[6] = { TXX, TYY, TZZ, TXZ, TYZ, TXY };
  for (int is = 0; is < s->total_number; is++) {
    if (it < s->it_begin[is] || it > s->it_end[is]) continue;
    size_t tab = ((size_t)is * s->max_nt + (it - s->it_begin[is])) * s->max_stage + istage;
    float F[3] = { 0, 0, 0 }, M[6] = { 0, 0, 0, 0, 0, 0 };
    if (s->force_actived) { F[0] = s->Fx[tab]; F[1] = s->Fy[tab]; F[2] = s->Fz[tab]; }
    if (s->moment_actived) { M[0] = s->Mxx[tab]; M[1] = s->Myy[tab]; M[2] = s->Mzz[tab]; M[3] = s->Mxz[tab]; M[4] = s->Myz[tab]; M[5] = s->Mxy[tab]; }
    int si = s->si[is], sj = s->sj[is], sk = s->sk[is];
    if (s->itype_spatial_ext == CGFD_SRC_SPATIAL_POINT) {
      size_t p = si + sj * o->L + sk * o->S;
      if (s->force_actived && (s->is_surface_force_strict == 0 || sk < g->nk2)) {
        float Vw = o->slw[p] / o->metric[JAC][p];
        for (int c = 0; c < 3; c++) h[c * o->V + p] += F[c] * Vw;
      }
      if (s->moment_actived) {
        float rj = 1.0 / o->metric[JAC][p];
        for (int c = 0; c < 6; c++) h[MC[c] * o->V + p] -= M[c] * rj;
      }
    } else {
      int k2 = (sk + H < g->nk2) ? H : g->nk2 - sk;
      footprint(ext, s->si_inc[is], s->sj_inc[is], s->sk_inc[is], s->ext_func_coef, H, k2);
      int ie = 0;
      for (int ke = -H; ke <= k2; ke++) for (int je = -H; je <= H; je++) for (int ix = -H; ix <= H; ix++, ie++) {
        int i = si + ix, j = sj + je, k = sk + ke;
        if (i < g->ni1 || i > g->ni2 || j < g->nj1 || j > g->nj2 || k < g->nk1 || k > g->nk2) continue;
        size_t p = i + j * o->L + k * o->S;
        float coef = ext[ie];
        if (s->force_actived && (s->is_surface_force_strict == 0 || k < g->nk2)) {
          float Vw = coef * o->slw[p] / o->metric[JAC][p];
          for (int c = 0; c < 3; c++) h[c * o->V + p] += F[c] * Vw;
        }
        if (s->moment_actived) {
          float rj = coef / o->metric[JAC][p];
          for (int c = 0; c < 6; c++) h[MC[c] * o->V + p] -= M[c] * rj;
        }
      }
    }
  }
  free(ext);
}

static void onestage(orc_t *o, const float *w, float *h, int it, int ipair, int istage, int aux_cur_level)
{
  const cgfd_fd_t *fd = &o->p.fd;
  ops_t op;
  for (int a = 0; a < 3; a++) {
    op.dir[a] = fd->dir[ipair][istage][a];
    op.first[a] = fd->indx[op.dir[a]][0];
    op.c[a] = fd->coef[op.dir[a]];
  }
  set_surface_force(o, it, istage);
  rhs_inner(o, w, h, &op);
  if (o->p.free_top) { rhs_timg(o, w, h, &op); rhs_vlow(o, w, h, &op); }
  rhs_cfspml(o, w, h, &op, aux_cur_level);
  if (o->p.src.total_number > 0) rhs_src(o, h, it, istage);
  if (o->dd_n > 0 && it < o->dd_nt) {
    /* sv_curv_col_el_rhs_srcdd: hV += vi * slw / J ; hT -= mij * (1 / J), table order xx yy zz yz xz xy */
    static const int TC[6] = {3, 4, 5, 6, 7, 8};   /* Txx Tyy Tzz Tyz Txz Txy in wavefield order */
    const size_t row = ((size_t)it * 4 + istage) * o->dd_n;
    for (int q = 0; q < o->dd_n; q++) {
      size_t p = o->dd_indx[q];
      if (o->dd_vi) {
        float Vw = o->slw[p] / o->metric[JAC][p];
        for (int c = 0; c < 3; c++) h[c * o->V + p] += o->dd_vi[(row + q) * 3 + c] * Vw;
      }
      if (o->dd_mij) {
        float rj = 1.0 / o->metric[JAC][p];
        for (int c = 0; c < 6; c++) h[TC[c] * o->V + p] -= o->dd_mij[(row + q) * 6 + c] * rj;
      }
    }
  }
}

int cgfd_oracle_onestage(void *hd, int it, int ipair, int istage, const float *w_cur, float *rhs)
{
  orc_t *o = hd;
  memcpy(o->lev[1], w_cur, o->V * 9 * sizeof(float));
  memset(o->lev[2], 0, o->V * 9 * sizeof(float));
  onestage(o, o->lev[1], o->lev[2], it, ipair, istage, 0);
  memcpy(rhs, o->lev[2], o->V * 9 * sizeof(float));
  return 0;
}

static void axpy_set(float *y, const float *x, float a, const float *r, size_t n) { for (size_t i = 0; i < n; i++) y[i] = x[i] + a * r[i]; }
static void axpy_add(float *y, float a, const float *r, size_t n) { for (size_t i = 0; i < n; i++) y[i] += a * r[i]; }

/* nsteps RK4 steps from step 0; w in/out (level n); rec[(it*9+c)*nrec+ip] sampled after every step */
int cgfd_oracle_run(void *hd, int nsteps, float *w, int nrec, const int64_t *rec_iptr, float *rec, double *seconds)
{
  orc_t *o = hd;
  const cgfd_fd_t *fd = &o->p.fd;
  const size_t n = o->V * 9;
  float *pre = o->lev[0], *tmp = o->lev[1], *rhs = o->lev[2], *end = o->lev[3];
  memcpy(pre, w, n * sizeof(float));
  memset(tmp, 0, n * sizeof(float)); memset(rhs, 0, n * sizeof(float)); memset(end, 0, n * sizeof(float));
  int apre = 0, aend = 3; /* aux level roles; tmp = 1, rhs = 2 */
  struct timespec t0, t1;
  clock_gettime(CLOCK_MONOTONIC, &t0);
  for (int it = 0; it < nsteps; it++) {
    int ipair = it % CGFD_NUM_PAIRS;
    for (int s = 0; s < CGFD_NUM_STAGES; s++) {
      const float *cur = (s == 0) ? pre : tmp;
      onestage(o, cur, rhs, it, ipair, s, (s == 0) ? apre : 1);
      float a = fd->rk_a[s] * o->p.dt, b = fd->rk_b[s] * o->p.dt;
      if (s < CGFD_NUM_STAGES - 1) axpy_set(tmp, pre, a, rhs, n);
      if (s == 0) axpy_set(end, pre, b, rhs, n); else axpy_add(end, b, rhs, n);
      if (s == CGFD_NUM_STAGES - 1 && o->Qs) {
        /* sv_curv_col_el_graves_Qs (forward/sv_curv_col_el.c:638-666, called at forward/drv_rk_curv_col.c:413-416) */
        const cgfd_grid_t *g = &o->p.grid;
        float coef = -3.14159265358979323846264338327950288419716939937510 * o->Qs_freq * o->p.dt;
        for (int c = 0; c < 9; c++)
          for (int k = g->nk1; k <= g->nk2; k++) for (int j = g->nj1; j <= g->nj2; j++) for (int i = g->ni1; i <= g->ni2; i++) {
            size_t p = i + j * o->L + k * o->S;
            end[c * o->V + p] *= expf(coef / o->Qs[p]);
          }
      }
      for (int d = 0; d < 3; d++) for (int q = 0; q < 2; q++) {
        face_t *f = &o->f[d][q];
        if (!f->on) continue;
        size_t m = f->siz * 9;
        if (s < CGFD_NUM_STAGES - 1) axpy_set(f->lev[1], f->lev[apre], a, f->lev[2], m);
        if (s == 0) axpy_set(f->lev[aend], f->lev[apre], b, f->lev[2], m); else axpy_add(f->lev[aend], b, f->lev[2], m);
      }
    }
    if (o->p.ablexp_enabled) {
      /* bdry_ablexp_apply (forward/bdry_t.c:840-890, called at forward/drv_rk_curv_col.c:483-485): W *= min(Ex[i], Ey[j], Ez[k]) in the shell blocks */
      for (int c = 0; c < 9; c++) for (int nb = 0; nb < 6; nb++) {
        const int32_t *B = o->p.ablexp_blk[nb];
        if (B[0] != 1) continue;
        for (int k = B[5]; k <= B[6]; k++) for (int j = B[3]; j <= B[4]; j++) for (int i = B[1]; i <= B[2]; i++) {
          float m = (o->p.ablexp_Ex[i] < o->p.ablexp_Ey[j]) ? o->p.ablexp_Ex[i] : o->p.ablexp_Ey[j];
          if (m > o->p.ablexp_Ez[k]) m = o->p.ablexp_Ez[k];
          end[c * o->V + (size_t)k * o->S + (size_t)j * o->L + i] *= m;
        }
      }
    }
    for (int ip = 0; ip < nrec; ip++) for (int c = 0; c < 9; c++) rec[((size_t)it * 9 + c) * nrec + ip] = end[c * o->V + rec_iptr[ip]];
    /* the reference re-zeroes the ghosts of rhs because it used it as output scratch (wav_zero_edge,
     * drv_rk_curv_col.c:512); here rhs ghosts are never written, so nothing to do */
    float *t = pre; pre = end; end = t;
    int ti = apre; apre = aend; aend = ti;
  }
  clock_gettime(CLOCK_MONOTONIC, &t1);
  if (seconds) *seconds = (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec);
  memcpy(w, pre, n * sizeof(float));
  /* keep level roles canonical for the next call */
  if (pre != o->lev[0]) { memcpy(o->lev[0], pre, n * sizeof(float)); }
  for (int d = 0; d < 3; d++) for (int q = 0; q < 2; q++) {
    face_t *f = &o->f[d][q];
    if (f->on && apre != 0) memcpy(f->lev[0], f->lev[apre], f->siz * 9 * sizeof(float));
  }
  return 0;
}
