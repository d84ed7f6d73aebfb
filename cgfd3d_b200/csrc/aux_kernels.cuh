#pragma once
#include "cgfd_dev.cuh"

namespace cgfd {

// device copy of the source tables (forward/src_t.h:21-128) plus the flattened footprints
struct SrcDev {
  int nsrc, max_nt, max_stage;
  int force_actived, moment_actived;
  const int *it_begin, *it_end;
  const float *Fx, *Fy, *Fz, *Mxx, *Myy, *Mzz, *Mxz, *Myz, *Mxy;
  const float *Fx_rate, *Fy_rate, *Fz_rate;
  // body-source footprint points: grid index, owning source, weights w*slw/J (force) and w/J (moment)
  int npts;
  const int64_t *pt_iptr;
  const int *pt_src;
  const float *pt_wV, *pt_wM;
  // surface-force footprint points on the k = nk2 slice
  int nsurf_pts;
  const int *sf_src, *sf_rate_slot, *sf_iptr2d;
  const float *sf_coef, *sf_coef_over_jac;
};

// footprint points [first, first+count) of the list (the list is sorted: points of the boundary phase first)
__global__ void k_src_inject(SrcDev S, int first, int count, int it, int istage, float *tmp, float *end, float a, float b, size_t V,
                             int kind, const float *qatt);
__global__ void k_graves_factor(float *q, size_t n, float coef);
// points sel[0 .. count) (point numbers = columns of the time-function tables vi [n][3], mij [n][6] of this step and stage)
__global__ void k_srcdd_inject(int count, const int *sel, const int64_t *iptr, const float *wV, const float *rjac, const float *vi,
                               const float *mij, float *tmp, float *end, float a, float b, size_t V, int kind, const float *qatt);
__global__ void k_srcdd_weights(int n, const int64_t *iptr, const float *slw, const float *jac, float *wV, float *rjac);
__global__ void k_src_surface(SrcDev S, int it, int istage, float *Tx, float *Ty, float *Tz, float *Vx, float *Vy, float *Vz);
__global__ void k_record(const float *w, size_t V, int ncmp, int npts, const int64_t *iptr, float *rec_it);
__global__ void k_pack_box(const float *w, int nx, int ny, int i1, int ni, int di, int j1, int nj, int dj, int k1, int nk,
                           int dk, float *out);
__global__ void k_pg(const float *w_new, const float *w_old, size_t V, int pitch, int nx, int ny, int ni1, int ni2, int nj1,
                     int nj2, int nk2, float dt, float *PG, float *Dis);
__global__ void k_vis_free(float *w, size_t V, int pitch, int nx, int ny, int ni1, int ni2, int nj1, int nj2, int nk2, const float *matD);
__global__ void k_ablexp(float *w, size_t V, int ncmp, int nx, int ny, int i1, int i2, int j1, int j2, int k1, int k2,
                         const float *Ex, const float *Ey, const float *Ez);
struct MetricOut { float *a[10]; };
// inputs / outputs of k_dvh2dvz: the k = nk2 planes [ny][nx] of the metric (reference order jac, xi_x .. zeta_z), of the media
// arrays (order of include/cgfd3d_b200.h) and, for the visco-elastic medium, of the coordinates; matrices [ny][nx][9]
struct DvhArgs {
  int med;                      // MED_* (0 iso, 1 vti, 2 aniso, 3 visco)
  int nx, ni1, ni2, nj1, nj2;
  const float *metric[10];
  const float *media[24];
  const float *x, *y, *z;
  int fd_len; const int *fd_indx; const float *fd_coef;
  float *matVx2Vz, *matVy2Vz, *matF2Vz, *matD;
};
__global__ void k_dvh2dvz(DvhArgs a);
__global__ void k_metric_cal(const float *x, const float *y, const float *z, int nx, int ny, int ni1, int ni2, int nj1, int nk1,
                             int fd_len, const int *fd_indx, const float *fd_coef, MetricOut out);
__global__ void k_metric_mirror(MetricOut out, int axis, int nx, int ny, int nz, int n1, int n2);
__global__ void k_repitch(float *pad, float *flat, int nx, int pitch, size_t rows, int to_padded);
__global__ void k_any_nonzero(const float *a, int pitch, int ny, int i1, int i2, int j1, int k1, int *flag);
__global__ void k_halo_copy(float *w, float *buf, size_t V, int ncmp, int nx, int ny, int i1, int ni, int j1, int nj, int k1,
                            int nk, int unpack);

}  // namespace cgfd
