/* In-memory NetCDF stand-in (see netcdf.h). A file is held as dims + dense
 * variables; the record (unlimited) dimension grows on demand. nc_close of a
 * created file serialises it as:
 *   "CGNC1\n"
 *   i32 ndims  { i32 namelen, name, i64 len, i32 is_unlimited } ...
 *   i32 nvars  { i32 namelen, name, i32 type, i32 ndims, i32 dimids[ndims],
 *                i32 natts { i32 namelen, name, i32 n, i32 v[n] }..., i64 nelem, data[nelem*4] } ...
 *   i32 nglobal_atts { ... }
 * nc_open parses the same container. All data are 4-byte (float or int). */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#include "netcdf.h"

#define MAXF 256
#define MAXD 16
#define MAXV 64
#define MAXA 32
#define MAXN 256

typedef struct { char name[MAXN]; int n; int *v; } att_t;
typedef struct { char name[MAXN]; size_t len; int unlimited; } dim_t;
typedef struct {
  char name[MAXN]; int type; int ndims; int dimids[8];
  int natts; att_t atts[MAXA];
  size_t nelem; size_t cap; char *data; /* 4-byte elements */
} var_t;
typedef struct {
  int used; int writable; char path[1024];
  int ndims; dim_t dims[MAXD];
  int nvars; var_t vars[MAXV];
  int ngatts; att_t gatts[MAXA];
} file_t;

static file_t *files[MAXF];

static file_t *getf(int id) {
  if (id < 0 || id >= MAXF || !files[id] || !files[id]->used) {
    fprintf(stderr, "netcdf shim: bad ncid %d\n", id); exit(4);
  }
  return files[id];
}
static int newf(void) {
  for (int i = 0; i < MAXF; i++) if (!files[i]) {
    files[i] = (file_t *)calloc(1, sizeof(file_t)); files[i]->used = 1; return i;
  }
  fprintf(stderr, "netcdf shim: too many open files\n"); exit(4);
}

const char *nc_strerror(int e) { return e == 0 ? "no error" : "netcdf shim error"; }

int nc_create(const char *path, int cmode, int *ncidp) {
  (void)cmode; int id = newf(); file_t *f = files[id];
  f->writable = 1; snprintf(f->path, sizeof(f->path), "%s", path);
  FILE *fp = fopen(path, "wb"); if (!fp) { free(f); files[id] = NULL; return -31; }
  fclose(fp); *ncidp = id; return 0;
}
int nc_enddef(int ncid) { (void)ncid; return 0; }

int nc_def_dim(int ncid, const char *name, size_t len, int *idp) {
  file_t *f = getf(ncid); if (f->ndims >= MAXD) return -57;
  dim_t *d = &f->dims[f->ndims]; snprintf(d->name, MAXN, "%s", name);
  d->len = len; d->unlimited = (len == 0); *idp = f->ndims++; return 0;
}
int nc_def_var(int ncid, const char *name, nc_type t, int nd, const int *dimids, int *varidp) {
  file_t *f = getf(ncid); if (f->nvars >= MAXV || nd > 8) return -48;
  var_t *v = &f->vars[f->nvars]; memset(v, 0, sizeof(*v));
  snprintf(v->name, MAXN, "%s", name); v->type = t; v->ndims = nd;
  for (int i = 0; i < nd; i++) v->dimids[i] = dimids[i];
  *varidp = f->nvars++; return 0;
}
static int put_att(att_t *arr, int *n, const char *name, size_t len, const int *op) {
  if (*n >= MAXA) return -44;
  att_t *a = &arr[*n]; snprintf(a->name, MAXN, "%s", name); a->n = (int)len;
  a->v = (int *)malloc(sizeof(int) * (len ? len : 1)); memcpy(a->v, op, sizeof(int) * len);
  (*n)++; return 0;
}
int nc_put_att_int(int ncid, int varid, const char *name, nc_type t, size_t len, const int *op) {
  (void)t; file_t *f = getf(ncid);
  if (varid == NC_GLOBAL) return put_att(f->gatts, &f->ngatts, name, len, op);
  if (varid < 0 || varid >= f->nvars) return -49;
  return put_att(f->vars[varid].atts, &f->vars[varid].natts, name, len, op);
}
int nc_get_att_int(int ncid, int varid, const char *name, int *ip) {
  file_t *f = getf(ncid); att_t *arr; int n;
  if (varid == NC_GLOBAL) { arr = f->gatts; n = f->ngatts; }
  else { if (varid < 0 || varid >= f->nvars) return -49; arr = f->vars[varid].atts; n = f->vars[varid].natts; }
  for (int i = 0; i < n; i++) if (!strcmp(arr[i].name, name)) { memcpy(ip, arr[i].v, sizeof(int) * arr[i].n); return 0; }
  return -43;
}
int nc_inq_varid(int ncid, const char *name, int *varidp) {
  file_t *f = getf(ncid);
  for (int i = 0; i < f->nvars; i++) if (!strcmp(f->vars[i].name, name)) { *varidp = i; return 0; }
  return -49;
}
int nc_inq_dimid(int ncid, const char *name, int *idp) {
  file_t *f = getf(ncid);
  for (int i = 0; i < f->ndims; i++) if (!strcmp(f->dims[i].name, name)) { *idp = i; return 0; }
  return -46;
}
int nc_inq_dimlen(int ncid, int dimid, size_t *lenp) {
  file_t *f = getf(ncid); if (dimid < 0 || dimid >= f->ndims) return -46;
  *lenp = f->dims[dimid].len; return 0;
}

/* make sure the dense buffer of v covers nrec records (record var) or the full var */
static void ensure(file_t *f, var_t *v, size_t nrec_needed) {
  size_t inner = 1; int rec = 0;
  for (int i = 0; i < v->ndims; i++) {
    dim_t *d = &f->dims[v->dimids[i]];
    if (i == 0 && d->unlimited) { rec = 1; if (d->len < nrec_needed) d->len = nrec_needed; }
    else inner *= d->len;
  }
  size_t need = rec ? inner * f->dims[v->dimids[0]].len : inner;
  if (need > v->cap) {
    size_t ncap = v->cap ? v->cap : 1024; while (ncap < need) ncap *= 2;
    v->data = (char *)realloc(v->data, ncap * 4);
    memset(v->data + v->cap * 4, 0, (ncap - v->cap) * 4); v->cap = ncap;
  }
  if (need > v->nelem) v->nelem = need;
}
static void hyperslab(file_t *f, var_t *v, const size_t *start, const size_t *count, char *buf, int put) {
  size_t dl[8], idx[8]; int nd = v->ndims;
  if (nd == 0) { ensure(f, v, 0); if (put) memcpy(v->data, buf, 4); else memcpy(buf, v->data, 4); return; }
  ensure(f, v, start[0] + count[0]);
  for (int i = 0; i < nd; i++) { dl[i] = f->dims[v->dimids[i]].len; idx[i] = 0; }
  size_t total = 1; for (int i = 0; i < nd; i++) total *= count[i];
  if (total == 0) return;
  size_t run = count[nd - 1], nrun = total / run, b = 0;
  for (size_t r = 0; r < nrun; r++) {
    size_t off = 0;
    for (int i = 0; i < nd; i++) off = off * dl[i] + (start[i] + idx[i]);
    if (put) memcpy(v->data + off * 4, buf + b * 4, run * 4);
    else memcpy(buf + b * 4, v->data + off * 4, run * 4);
    b += run;
    for (int i = nd - 2; i >= 0; i--) { if (++idx[i] < count[i]) break; idx[i] = 0; }
  }
}
static var_t *getv(file_t *f, int varid) {
  if (varid < 0 || varid >= f->nvars) { fprintf(stderr, "netcdf shim: bad varid %d\n", varid); exit(4); }
  return &f->vars[varid];
}
int nc_put_vara_float(int ncid, int varid, const size_t *s, const size_t *c, const float *op)
{ file_t *f = getf(ncid); hyperslab(f, getv(f, varid), s, c, (char *)op, 1); return 0; }
int nc_get_vara_float(int ncid, int varid, const size_t *s, const size_t *c, float *ip)
{ file_t *f = getf(ncid); hyperslab(f, getv(f, varid), s, c, (char *)ip, 0); return 0; }
static void whole(file_t *f, var_t *v, size_t *s, size_t *c) {
  for (int i = 0; i < v->ndims; i++) { s[i] = 0; c[i] = f->dims[v->dimids[i]].len; }
}
int nc_put_var_float(int ncid, int varid, const float *op)
{ file_t *f = getf(ncid); var_t *v = getv(f, varid); size_t s[8], c[8]; whole(f, v, s, c); hyperslab(f, v, s, c, (char *)op, 1); return 0; }
int nc_get_var_float(int ncid, int varid, float *ip)
{ file_t *f = getf(ncid); var_t *v = getv(f, varid); size_t s[8], c[8]; whole(f, v, s, c); hyperslab(f, v, s, c, (char *)ip, 0); return 0; }
int nc_get_var(int ncid, int varid, void *ip) { return nc_get_var_float(ncid, varid, (float *)ip); }
int nc_put_var1_float(int ncid, int varid, const size_t *idx, const float *op)
{ file_t *f = getf(ncid); var_t *v = getv(f, varid); size_t c[8]; for (int i = 0; i < 8; i++) c[i] = 1; hyperslab(f, v, idx, c, (char *)op, 1); return 0; }

static void wi32(FILE *fp, int32_t x) { fwrite(&x, 4, 1, fp); }
static void wi64(FILE *fp, int64_t x) { fwrite(&x, 8, 1, fp); }
static void wstr(FILE *fp, const char *s) { int32_t n = (int32_t)strlen(s); wi32(fp, n); fwrite(s, 1, n, fp); }
static void watts(FILE *fp, att_t *a, int n) {
  wi32(fp, n); for (int i = 0; i < n; i++) { wstr(fp, a[i].name); wi32(fp, a[i].n); fwrite(a[i].v, 4, a[i].n, fp); }
}
static int32_t ri32(FILE *fp) { int32_t x = 0; if (fread(&x, 4, 1, fp) != 1) { fprintf(stderr, "netcdf shim: short read\n"); exit(4); } return x; }
static int64_t ri64(FILE *fp) { int64_t x = 0; if (fread(&x, 8, 1, fp) != 1) { fprintf(stderr, "netcdf shim: short read\n"); exit(4); } return x; }
static void rstr(FILE *fp, char *s) { int32_t n = ri32(fp); if (n >= MAXN) n = MAXN - 1; if (fread(s, 1, n, fp) != (size_t)n) exit(4); s[n] = 0; }
static void ratts(FILE *fp, att_t *a, int *n) {
  *n = ri32(fp);
  for (int i = 0; i < *n; i++) { rstr(fp, a[i].name); a[i].n = ri32(fp); a[i].v = (int *)malloc(4 * (a[i].n ? a[i].n : 1)); if (fread(a[i].v, 4, a[i].n, fp) != (size_t)a[i].n) exit(4); }
}

int nc_close(int ncid) {
  file_t *f = getf(ncid);
  if (f->writable) {
    FILE *fp = fopen(f->path, "wb"); if (!fp) return -31;
    fwrite("CGNC1\n", 1, 6, fp);
    wi32(fp, f->ndims);
    for (int i = 0; i < f->ndims; i++) { wstr(fp, f->dims[i].name); wi64(fp, (int64_t)f->dims[i].len); wi32(fp, f->dims[i].unlimited); }
    wi32(fp, f->nvars);
    for (int i = 0; i < f->nvars; i++) {
      var_t *v = &f->vars[i]; ensure(f, v, 0);
      wstr(fp, v->name); wi32(fp, v->type); wi32(fp, v->ndims);
      for (int d = 0; d < v->ndims; d++) wi32(fp, v->dimids[d]);
      watts(fp, v->atts, v->natts);
      /* record vars: clip to the final record count */
      size_t n = 1; for (int d = 0; d < v->ndims; d++) n *= f->dims[v->dimids[d]].len;
      if (n > v->nelem) { ensure(f, v, f->dims[v->dimids[0]].len); }
      wi64(fp, (int64_t)n); fwrite(v->data, 4, n, fp);
    }
    watts(fp, f->gatts, f->ngatts);
    fclose(fp);
  }
  for (int i = 0; i < f->nvars; i++) { free(f->vars[i].data); for (int a = 0; a < f->vars[i].natts; a++) free(f->vars[i].atts[a].v); }
  for (int a = 0; a < f->ngatts; a++) free(f->gatts[a].v);
  free(f); files[ncid] = NULL; return 0;
}

int nc_open(const char *path, int mode, int *ncidp) {
  (void)mode; FILE *fp = fopen(path, "rb"); if (!fp) return -31;
  char magic[6]; if (fread(magic, 1, 6, fp) != 6 || memcmp(magic, "CGNC1\n", 6)) { fclose(fp); fprintf(stderr, "netcdf shim: %s is not a CGNC1 container\n", path); return -51; }
  int id = newf(); file_t *f = files[id]; snprintf(f->path, sizeof(f->path), "%s", path);
  f->ndims = ri32(fp);
  for (int i = 0; i < f->ndims; i++) { rstr(fp, f->dims[i].name); f->dims[i].len = (size_t)ri64(fp); f->dims[i].unlimited = ri32(fp); }
  f->nvars = ri32(fp);
  for (int i = 0; i < f->nvars; i++) {
    var_t *v = &f->vars[i]; rstr(fp, v->name); v->type = ri32(fp); v->ndims = ri32(fp);
    for (int d = 0; d < v->ndims; d++) v->dimids[d] = ri32(fp);
    ratts(fp, v->atts, &v->natts);
    v->nelem = (size_t)ri64(fp); v->cap = v->nelem ? v->nelem : 1; v->data = (char *)malloc(v->cap * 4);
    if (fread(v->data, 4, v->nelem, fp) != v->nelem) { fprintf(stderr, "netcdf shim: short data\n"); exit(4); }
  }
  ratts(fp, f->gatts, &f->ngatts);
  fclose(fp); *ncidp = id; return 0;
}
