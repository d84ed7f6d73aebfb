// Small per-stage / per-step kernels around the fused RHS pass: source injection, surface-force
// slices, output taps (record points, sub-box pack, peak ground motion), sponge, halo pack/unpack.
#include "aux_kernels.cuh"

namespace cgfd {

// Point / Gaussian body sources added after the fused stage update: the reference adds
// F*w*slw/J to hV and -M*w/J to hT before the RK axpy (forward/sv_curv_col_el.c:350-476);
// here the same term is pushed through the axpy: tmp += a*s, end += b*s.
__global__ void k_src_inject(SrcDev S, int first, int count, int it, int istage, float *tmp, float *end, float a, float b, size_t V,
                             int kind, const float *qatt)
{
  int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= count) return;
  n += first;
  const int is = S.pt_src[n];
  const int itb = S.it_begin[is], ite = S.it_end[is];
  if (it < itb || it > ite) return;
  const size_t tab = ((size_t)is * S.max_nt + (size_t)(it - itb)) * S.max_stage + istage;
  const size_t p = S.pt_iptr[n];
  float add[9];
#pragma unroll
  for (int c = 0; c < 9; c++) add[c] = 0.0f;
  if (S.force_actived) {
    const float w = S.pt_wV[n];
    add[VX] = S.Fx[tab] * w; add[VY] = S.Fy[tab] * w; add[VZ] = S.Fz[tab] * w;
  }
  if (S.moment_actived) {
    const float w = S.pt_wM[n];
    add[TXX] = -(S.Mxx[tab] * w); add[TYY] = -(S.Myy[tab] * w); add[TZZ] = -(S.Mzz[tab] * w);
    add[TXZ] = -(S.Mxz[tab] * w); add[TYZ] = -(S.Myz[tab] * w); add[TXY] = -(S.Mxy[tab] * w);
  }
#pragma unroll
  for (int c = 0; c < 9; c++) {
    if (add[c] != 0.0f) {
      // stages 0 and 2 reach w_end through w_tmp - w_pre one stage later (rk_wave, physics.cuh)
      if (kind != KIND_LAST) atomicAdd(tmp + c * V + p, a * add[c]);
      // the last stage's w_end has already been multiplied by Graves' factor: so is the source's share of it
      if (kind == KIND_MID || kind == KIND_LAST) atomicAdd(end + c * V + p, b * add[c] * ((kind == KIND_LAST && qatt) ? qatt[p] : 1.0f));
    }
  }
}

// Distributed (finite-fault) sources, sv_curv_col_el_rhs_srcdd (forward/sv_curv_col_el.c:486-632, add-at-point branch):
// hV += vi * slw/J, hT -= mij / J at every dd point, pushed through the RK axpy like k_src_inject. vi / mij point at the rows of
// this step and stage inside the resident time block: [n][3] and [n][6] (component order xx yy zz yz xz xy = TXX..TXY).
__global__ void k_srcdd_inject(int count, const int *sel, const int64_t *iptr, const float *wV, const float *rjac, const float *vi,
                               const float *mij, float *tmp, float *end, float a, float b, size_t V, int kind, const float *qatt)
{
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= count) return;
  const int is = sel[t];   // point number: column of the time-function tables
  const size_t p = iptr[is];
  float add[9];
#pragma unroll
  for (int c = 0; c < 9; c++) add[c] = 0.0f;
  if (vi) {
    const float w = wV[is];
#pragma unroll
    for (int c = 0; c < 3; c++) add[c] = vi[is * 3 + c] * w;
  }
  if (mij) {
    const float w = rjac[is];
#pragma unroll
    for (int c = 0; c < 6; c++) add[3 + c] = -(mij[is * 6 + c] * w);
  }
  const float q = (kind == KIND_LAST && qatt) ? qatt[p] : 1.0f;
#pragma unroll
  for (int c = 0; c < 9; c++) {
    if (add[c] != 0.0f) {
      if (kind != KIND_LAST) atomicAdd(tmp + c * V + p, a * add[c]);
      if (kind == KIND_MID || kind == KIND_LAST) atomicAdd(end + c * V + p, b * add[c] * q);
    }
  }
}
// weights of the dd points: slw / J (float division, like the reference) and (float)(1.0 / J)
__global__ void k_srcdd_weights(int n, const int64_t *iptr, const float *slw, const float *jac, float *wV, float *rjac)
{
  const int is = blockIdx.x * blockDim.x + threadIdx.x;
  if (is >= n) return;
  const size_t p = iptr[is];
  wV[is] = slw[p] / jac[p];
  rjac[is] = (float)(1.0 / (double)jac[p]);
}

// surface force slices TxSrc.. / VxSrc.. of one stage (forward/src_t.c:153-314); slices are
// zeroed by a memset before this kernel.
__global__ void k_src_surface(SrcDev S, int it, int istage, float *Tx, float *Ty, float *Tz, float *Vx, float *Vy, float *Vz)
{
  int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= S.nsurf_pts) return;
  const int ns = S.sf_rate_slot[n];          // index into the *_rate tables
  const int is = S.sf_src[n];
  const int itb = S.it_begin[is], ite = S.it_end[is];
  if (it < itb || it > ite) return;
  const size_t tab = ((size_t)is * S.max_nt + (size_t)(it - itb)) * S.max_stage + istage;
  const size_t tabr = ((size_t)ns * S.max_nt + (size_t)(it - itb)) * S.max_stage + istage;
  const int p2 = S.sf_iptr2d[n];
  const float c = S.sf_coef[n], cj = S.sf_coef_over_jac[n];
  atomicAdd(Tx + p2, S.Fx[tab] * c); atomicAdd(Ty + p2, S.Fy[tab] * c); atomicAdd(Tz + p2, S.Fz[tab] * c);
  atomicAdd(Vx + p2, S.Fx_rate[tabr] * cj); atomicAdd(Vy + p2, S.Fy_rate[tabr] * cj); atomicAdd(Vz + p2, S.Fz_rate[tabr] * cj);
}

// io_line_keep / integer-point io_recv_keep (forward/io_funcs.c:1584-1645): one sample per
// component per registered point per step.
__global__ void k_record(const float *w, size_t V, int ncmp, int npts, const int64_t *iptr, float *rec_it)
{
  int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= npts * ncmp) return;
  int c = n / npts, ip = n % npts;
  rec_it[(size_t)c * npts + ip] = w[c * V + iptr[ip]];
}

// strided sub-box of one component (io_snap_nc_put / io_slice_nc_put packing,
// forward/io_funcs.c:1032-1104, 1161-1262)
__global__ void k_pack_box(const float *w, int nx, int ny, int i1, int ni, int di, int j1, int nj, int dj, int k1, int nk,
                           int dk, float *out)
{
  size_t n = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t tot = (size_t)ni * nj * nk;
  if (n >= tot) return;
  int ii = n % ni, jj = (n / ni) % nj, kk = n / ((size_t)ni * nj);
  out[n] = w[((size_t)(k1 + kk * dk) * ny + (j1 + jj * dj)) * nx + (i1 + ii * di)];
}

// Qs -> exp(coef / Qs) in place (sv_curv_col_el_graves_Qs, forward/sv_curv_col_el.c:638-666: coef = -pi f0 dt, float division,
// expf). exp in double rounded to float: the correctly rounded value, which glibc's expf returns too.
__global__ void k_graves_factor(float *q, size_t n, float coef)
{
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float x = coef / q[i];
  q[i] = (q[i] != 0.0f) ? (float)exp((double)x) : 1.0f;
}

// gd_curv_metric_cal (forward/gd_t.c:190-286): centred differences of the coordinates (M_FD_SHIFT, forward/fd_t.h:11-15: first term
// assigned, the others added left to right), Jacobian and the nine metric derivatives at one physical point. Every product and
// sum is rounded separately (__fmul_rn / __fadd_rn: no FMA contraction), like the reference built by gcc for x86-64, so the
// arrays come out bit-identical. out = 10 unpadded arrays [nz][ny][nx] in the reference's order jac, xi_x .. zeta_z.
__device__ __forceinline__ float fd_shift(const float *v, size_t p, long stride, int len, const int *indx, const float *coef)
{
  float d = __fmul_rn(coef[0], v[p + indx[0] * stride]);
  for (int n = 1; n < len; n++) d = __fadd_rn(d, __fmul_rn(coef[n], v[p + indx[n] * stride]));
  return d;
}
__device__ __forceinline__ void cross_rn(const float *A, const float *B, float *C)
{
  C[0] = __fsub_rn(__fmul_rn(A[1], B[2]), __fmul_rn(A[2], B[1]));
  C[1] = __fsub_rn(__fmul_rn(A[2], B[0]), __fmul_rn(A[0], B[2]));
  C[2] = __fsub_rn(__fmul_rn(A[0], B[1]), __fmul_rn(A[1], B[0]));
}
__global__ void k_metric_cal(const float *x, const float *y, const float *z, int nx, int ny, int ni1, int ni2, int nj1, int nk1,
                             int fd_len, const int *fd_indx, const float *fd_coef, MetricOut out)
{
  const int i = ni1 + blockIdx.x * blockDim.x + threadIdx.x;
  if (i > ni2) return;
  const long L = nx, S = (long)nx * ny;
  const size_t p = (size_t)(nk1 + blockIdx.z) * S + (size_t)(nj1 + blockIdx.y) * L + i;
  float v1[3], v2[3], v3[3], g[3];
  v1[0] = fd_shift(x, p, 1, fd_len, fd_indx, fd_coef); v1[1] = fd_shift(y, p, 1, fd_len, fd_indx, fd_coef); v1[2] = fd_shift(z, p, 1, fd_len, fd_indx, fd_coef);
  v2[0] = fd_shift(x, p, L, fd_len, fd_indx, fd_coef); v2[1] = fd_shift(y, p, L, fd_len, fd_indx, fd_coef); v2[2] = fd_shift(z, p, L, fd_len, fd_indx, fd_coef);
  v3[0] = fd_shift(x, p, S, fd_len, fd_indx, fd_coef); v3[1] = fd_shift(y, p, S, fd_len, fd_indx, fd_coef); v3[2] = fd_shift(z, p, S, fd_len, fd_indx, fd_coef);
  cross_rn(v1, v2, g);
  float jac = 0.0f;   // fdlib_math_dot_product (lib/fdlib_math.c:60-69): result = 0; result += A[i] * B[i]
  for (int n = 0; n < 3; n++) jac = __fadd_rn(jac, __fmul_rn(g[n], v3[n]));
  out.a[0][p] = jac;
  cross_rn(v2, v3, g);
  out.a[1][p] = __fdiv_rn(g[0], jac); out.a[2][p] = __fdiv_rn(g[1], jac); out.a[3][p] = __fdiv_rn(g[2], jac);
  cross_rn(v3, v1, g);
  out.a[4][p] = __fdiv_rn(g[0], jac); out.a[5][p] = __fdiv_rn(g[1], jac); out.a[6][p] = __fdiv_rn(g[2], jac);
  cross_rn(v1, v2, g);
  out.a[7][p] = __fdiv_rn(g[0], jac); out.a[8][p] = __fdiv_rn(g[1], jac); out.a[9][p] = __fdiv_rn(g[2], jac);
}
// ghosts of one axis mirrored about the mid-point between the last physical and the first ghost point, over the full extent of
// the other two axes (forward/gd_t.c:288-401, applied in the order x, y, z); blockIdx.z = array
__global__ void k_metric_mirror(MetricOut out, int axis, int nx, int ny, int nz, int n1, int n2)
{
  const size_t tot = (size_t)nx * ny * nz;
  const size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= tot) return;
  const int i = q % nx, j = (q / nx) % ny, k = q / ((size_t)nx * ny);
  const int c = (axis == 0) ? i : (axis == 1) ? j : k;
  if (c >= n1 && c <= n2) return;
  const long stride = (axis == 0) ? 1 : (axis == 1) ? nx : (long)nx * ny;
  const long off = (c < n1) ? ((long)(n1 - c) * 2 - 1) : -((long)(c - n2) * 2 - 1);
  float *a = out.a[blockIdx.z];
  a[q] = a[q + off * stride];
}

// The free-surface conversion matrices of the four constitutive laws (one-shot set-up):
//   sv_curv_col_el_iso_dvh2dvz   forward/sv_curv_col_el_iso.c:1258-1375    A^-1 B, A^-1 C, A^-1
//   sv_curv_col_el_vti_dvh2dvz   forward/sv_curv_col_el_vti.c:1085-1205    -(A^-1 B), -(A^-1 C)
//   sv_curv_col_el_aniso_dvh2dvz forward/sv_curv_col_el_aniso.c:1267-1422  -(A^-1 B), -(A^-1 C)
//   sv_curv_col_vis_iso_dvh2dvz  forward/sv_curv_col_vis_iso.c:353-507     A^-1 B, A^-1 C and the rotation matD into the local
//                                                                          frame of the surface (tangent along eta, normal)
// One thread per surface point; every input is the k = nk2 plane [ny][nx] of its array. Products and sums are rounded one by
// one in the reference's order (no FMA contraction: gcc does not contract on x86-64), so the matrices come out bit-identical.
// The VTI expressions of the reference are the general anisotropic ones with the vanishing Cij left out (adding an exact zero
// changes nothing), so both go through one routine over the 21 Cij.
__device__ __forceinline__ float mul_(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float add_(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float sub_(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ void invert3x3_rn(float m[3][3])   // fdlib_math_invert3x3, lib/fdlib_math.c:7-35
{
  float inv[3][3];
  inv[0][0] = sub_(mul_(m[1][1], m[2][2]), mul_(m[2][1], m[1][2]));
  inv[0][1] = sub_(mul_(m[2][1], m[0][2]), mul_(m[0][1], m[2][2]));
  inv[0][2] = sub_(mul_(m[0][1], m[1][2]), mul_(m[0][2], m[1][1]));
  inv[1][0] = sub_(mul_(m[1][2], m[2][0]), mul_(m[1][0], m[2][2]));
  inv[1][1] = sub_(mul_(m[0][0], m[2][2]), mul_(m[2][0], m[0][2]));
  inv[1][2] = sub_(mul_(m[1][0], m[0][2]), mul_(m[0][0], m[1][2]));
  inv[2][0] = sub_(mul_(m[1][0], m[2][1]), mul_(m[1][1], m[2][0]));
  inv[2][1] = sub_(mul_(m[2][0], m[0][1]), mul_(m[0][0], m[2][1]));
  inv[2][2] = sub_(mul_(m[0][0], m[1][1]), mul_(m[0][1], m[1][0]));
  float det = add_(add_(mul_(inv[0][0], m[0][0]), mul_(inv[0][1], m[1][0])), mul_(inv[0][2], m[2][0]));
  det = __fdiv_rn(1.0f, det);
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) m[i][j] = mul_(inv[i][j], det);
}
__device__ __forceinline__ void matmul3x3_rn(const float A[3][3], const float B[3][3], float C[3][3])   // lib/fdlib_math.c:37-49
{
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) {
    float c = 0.0f;
    for (int k = 0; k < 3; k++) c = add_(c, mul_(A[i][k], B[k][j]));
    C[i][j] = c;
  }
}
// isotropic: M = sgn * [ l2m a_i b_i + mu (a_p b_p + a_q b_q) on the diagonal ; lam a_i b_j + mu a_j b_i off it ], a = zeta row
__device__ __forceinline__ void iso_mat(const float *a, const float *b, float lam, float mu, float l2m, bool neg, float M[3][3])
{
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) {
    if (i == j) {
      const int p = (i + 1) % 3 < (i + 2) % 3 ? (i + 1) % 3 : (i + 2) % 3, q = (i + 1) % 3 < (i + 2) % 3 ? (i + 2) % 3 : (i + 1) % 3;
      const float t1 = mul_(mul_(neg ? -l2m : l2m, a[i]), b[i]);
      const float t2 = mul_(mu, add_(mul_(a[p], b[p]), mul_(a[q], b[q])));
      M[i][j] = neg ? sub_(t1, t2) : add_(t1, t2);
    } else {
      const float t1 = mul_(mul_(neg ? -lam : lam, a[i]), b[j]);
      const float t2 = mul_(mul_(mu, a[j]), b[i]);
      M[i][j] = neg ? sub_(t1, t2) : add_(t1, t2);
    }
  }
}
// general anisotropic: M[i][j] = sum_p ( sum_q C[V(i,p)][V(j,q)] in_q ) out_p, V = Voigt index of a tensor index pair
__device__ __forceinline__ void aniso_mat(const float (&C)[6][6], const float *in, const float *out, float M[3][3])
{
  const int V[3][3] = {{0, 5, 4}, {5, 1, 3}, {4, 3, 2}};
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) {
    float acc = 0.0f;
    for (int p = 0; p < 3; p++) {
      const float *c = C[V[i][p]];
      const float s = add_(add_(mul_(c[V[j][0]], in[0]), mul_(c[V[j][1]], in[1])), mul_(c[V[j][2]], in[2]));
      const float t = mul_(s, out[p]);
      acc = (p == 0) ? t : add_(acc, t);
    }
    M[i][j] = acc;
  }
}
__global__ void k_dvh2dvz(DvhArgs a)
{
  const int i = a.ni1 + blockIdx.x * blockDim.x + threadIdx.x, j = a.nj1 + blockIdx.y;
  if (i > a.ni2 || j > a.nj2) return;
  const size_t p = (size_t)j * a.nx + i;
  const float xi[3] = {a.metric[1][p], a.metric[2][p], a.metric[3][p]}, et[3] = {a.metric[4][p], a.metric[5][p], a.metric[6][p]},
              zt[3] = {a.metric[7][p], a.metric[8][p], a.metric[9][p]};
  float A[3][3], B[3][3], Cm[3][3], AB[3][3], AC[3][3];
  const bool iso = (a.med == 0 || a.med == 3);
  if (iso) {
    const float lam = a.media[0][p], mu = a.media[1][p], l2m = add_(lam, mul_(2.0f, mu));
    iso_mat(zt, zt, lam, mu, l2m, false, A);
    iso_mat(zt, xi, lam, mu, l2m, true, B);
    iso_mat(zt, et, lam, mu, l2m, true, Cm);
  } else {
    float C[6][6];
    for (int r = 0; r < 6; r++) for (int c = 0; c < 6; c++) C[r][c] = 0.0f;
    if (a.med == 1) {   // VTI: c11 c13 c33 c55 c66 (forward/sv_curv_col_el_vti.c:1147-1152), c12 = c11 - 2.0 * c66 in double
      const float c11 = a.media[0][p], c13 = a.media[1][p], c33 = a.media[2][p], c55 = a.media[3][p], c66 = a.media[4][p];
      const float c12 = (float)((double)c11 - 2.0 * (double)c66);
      C[0][0] = c11; C[1][1] = c11; C[2][2] = c33; C[3][3] = c55; C[4][4] = c55; C[5][5] = c66;
      C[0][1] = C[1][0] = c12; C[0][2] = C[2][0] = c13; C[1][2] = C[2][1] = c13;
    } else {            // the 21 Cij, upper triangle row by row
      int n = 0;
      for (int r = 0; r < 6; r++) for (int c = r; c < 6; c++) { C[r][c] = C[c][r] = a.media[n][p]; n++; }
    }
    aniso_mat(C, zt, zt, A);
    aniso_mat(C, xi, zt, B);
    aniso_mat(C, et, zt, Cm);
  }
  invert3x3_rn(A);
  matmul3x3_rn(A, B, AB);
  matmul3x3_rn(A, Cm, AC);
  for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) {
    a.matVx2Vz[p * 9 + r * 3 + c] = iso ? AB[r][c] : mul_(-1.0f, AB[r][c]);
    a.matVy2Vz[p * 9 + r * 3 + c] = iso ? AC[r][c] : mul_(-1.0f, AC[r][c]);
    if (a.med == 0) a.matF2Vz[p * 9 + r * 3 + c] = A[r][c];
  }
  if (a.med == 3) {
    // rotation into the local frame of the surface: tangent along eta (centred difference of the coordinates), normal = zeta row
    const float xet = fd_shift(a.x, p, a.nx, a.fd_len, a.fd_indx, a.fd_coef), yet = fd_shift(a.y, p, a.nx, a.fd_len, a.fd_indx, a.fd_coef),
                zet = fd_shift(a.z, p, a.nx, a.fd_len, a.fd_indx, a.fd_coef);
    const float e_n = (float)(1.0 / sqrt((double)add_(add_(mul_(xet, xet), mul_(yet, yet)), mul_(zet, zet))));
    const float e_m = (float)(1.0 / sqrt((double)add_(add_(mul_(zt[0], zt[0]), mul_(zt[1], zt[1])), mul_(zt[2], zt[2]))));
    const float e_nm = mul_(e_n, e_m);
    float *D = a.matD + p * 9;
    D[0] = mul_(sub_(mul_(yet, zt[2]), mul_(zet, zt[1])), e_nm);
    D[1] = mul_(sub_(mul_(zet, zt[0]), mul_(xet, zt[2])), e_nm);
    D[2] = mul_(sub_(mul_(xet, zt[1]), mul_(yet, zt[0])), e_nm);
    D[3] = mul_(xet, e_n); D[4] = mul_(yet, e_n); D[5] = mul_(zet, e_n);
    D[6] = mul_(zt[0], e_m); D[7] = mul_(zt[1], e_m); D[8] = mul_(zt[2], e_m);
  }
}

// rows of nx floats between the unpadded host order and the padded device rows (pitch PX; `pad` is already shifted)
__global__ void k_repitch(float *pad, float *flat, int nx, int pitch, size_t rows, int to_padded)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nx) return;
  for (size_t r = blockIdx.y; r < rows; r += gridDim.y) {
    if (to_padded) pad[r * pitch + i] = flat[r * nx + i];
    else flat[r * nx + i] = pad[r * pitch + i];
  }
}

// PGV / PGA / PGD maps on the free surface (PG_calcu, forward/wav_t.c:379-455)
__global__ void k_pg(const float *w_new, const float *w_old, size_t V, int pitch, int nx, int ny, int ni1, int ni2, int nj1,
                     int nj2, int nk2, float dt, float *PG, float *Dis)
{
  int i = ni1 + blockIdx.x * blockDim.x + threadIdx.x;
  int j = nj1 + blockIdx.y;
  if (i > ni2 || j > nj2) return;
  const size_t sl = (size_t)nx * ny;   // unpadded 2-D maps
  const size_t p = ((size_t)nk2 * ny + j) * pitch + i, p1 = (size_t)j * nx + i;
  const float vx1 = w_new[p], vy1 = w_new[V + p], vz1 = w_new[2 * V + p];
  const float vx0 = w_old[p], vy0 = w_old[V + p], vz0 = w_old[2 * V + p];
  const float Ax = fabsf((vx1 - vx0) / dt), Ay = fabsf((vy1 - vy0) / dt), Az = fabsf((vz1 - vz0) / dt);
  float dx = Dis[p1] + 0.5f * (vx1 + vx0) * dt, dy = Dis[sl + p1] + 0.5f * (vy1 + vy0) * dt,
        dz = Dis[2 * sl + p1] + 0.5f * (vz1 + vz0) * dt;
  Dis[p1] = dx; Dis[sl + p1] = dy; Dis[2 * sl + p1] = dz;
  const float Vv = sqrtf(vx1 * vx1 + vy1 * vy1 + vz1 * vz1), Vh = sqrtf(vx1 * vx1 + vy1 * vy1);
  const float A = sqrtf(Ax * Ax + Ay * Ay + Az * Az), Ah = sqrtf(Ax * Ax + Ay * Ay);
  const float D = sqrtf(dx * dx + dy * dy + dz * dz), Dh = sqrtf(dx * dx + dy * dy);
  const float vals[15] = {Vv, Vh, fabsf(vx1), fabsf(vy1), fabsf(vz1), A, Ah, Ax, Ay, Az, D, Dh, fabsf(dx), fabsf(dy), fabsf(dz)};
#pragma unroll
  for (int n = 0; n < 15; n++) {
    float *q = PG + n * sl + p1;
    if (*q < vals[n]) *q = vals[n];
  }
}

// visco-elastic free surface: after the last RK stage the surface stress is rotated into the local frame of the surface,
// its normal components are dropped and it is rotated back (sv_curv_col_vis_iso_free, forward/sv_curv_col_vis_iso.c:509-623):
// Tl = (D T D^T) restricted to the two tangential directions, T <- D^T Tl D, with D = matD[(j*nx+i)*9 ..].
__global__ void k_vis_free(float *w, size_t V, int pitch, int nx, int ny, int ni1, int ni2, int nj1, int nj2, int nk2, const float *matD)
{
  int i = ni1 + blockIdx.x * blockDim.x + threadIdx.x;
  int j = nj1 + blockIdx.y;
  if (i > ni2 || j > nj2) return;
  const size_t p = ((size_t)nk2 * ny + j) * pitch + i;
  const float *D = matD + ((size_t)j * nx + i) * 9;
  const float d11 = D[0], d12 = D[1], d13 = D[2], d21 = D[3], d22 = D[4], d23 = D[5], d31 = D[6], d32 = D[7], d33 = D[8];
  const float Txx = w[TXX * V + p], Tyy = w[TYY * V + p], Tzz = w[TZZ * V + p], Tyz = w[TYZ * V + p], Txz = w[TXZ * V + p],
              Txy = w[TXY * V + p];
  float Tl[3][3] = {{0.0f, 0.0f, 0.0f}, {0.0f, 0.0f, 0.0f}, {0.0f, 0.0f, 0.0f}};
  Tl[0][0] = d11 * d11 * Txx + d12 * d12 * Tyy + d13 * d13 * Tzz + 2 * (d11 * d12 * Txy + d11 * d13 * Txz + d12 * d13 * Tyz);
  Tl[0][1] = d11 * d21 * Txx + d12 * d22 * Tyy + d13 * d23 * Tzz + (d11 * d22 + d12 * d21) * Txy + (d11 * d23 + d21 * d13) * Txz
           + (d12 * d23 + d22 * d13) * Tyz;
  Tl[1][1] = d21 * d21 * Txx + d22 * d22 * Tyy + d23 * d23 * Tzz + 2 * (d21 * d22 * Txy + d21 * d23 * Txz + d22 * d23 * Tyz);
  Tl[1][0] = Tl[0][1];
  const float Dm[3][3] = {{d11, d12, d13}, {d21, d22, d23}, {d31, d32, d33}};
  float DTl[3][3], Tg[3][3];   // DTl = D^T Tl ; Tg = DTl D (fdlib_math_matmul3x3, lib/fdlib_math.c:37-49)
#pragma unroll
  for (int r = 0; r < 3; r++)
#pragma unroll
    for (int c = 0; c < 3; c++) DTl[r][c] = Dm[0][r] * Tl[0][c] + Dm[1][r] * Tl[1][c] + Dm[2][r] * Tl[2][c];
#pragma unroll
  for (int r = 0; r < 3; r++)
#pragma unroll
    for (int c = 0; c < 3; c++) Tg[r][c] = DTl[r][0] * Dm[0][c] + DTl[r][1] * Dm[1][c] + DTl[r][2] * Dm[2][c];
  w[TXX * V + p] = Tg[0][0]; w[TYY * V + p] = Tg[1][1]; w[TZZ * V + p] = Tg[2][2];
  w[TXY * V + p] = Tg[0][1]; w[TXZ * V + p] = Tg[0][2]; w[TYZ * V + p] = Tg[1][2];
}

// exponential sponge: W *= min(Ex[i],Ey[j],Ez[k]) inside one shell block (forward/bdry_t.c:840-890)
__global__ void k_ablexp(float *w, size_t V, int ncmp, int nx, int ny, int i1, int i2, int j1, int j2, int k1, int k2,
                         const float *Ex, const float *Ey, const float *Ez)
{
  int i = i1 + blockIdx.x * blockDim.x + threadIdx.x;
  int j = j1 + blockIdx.y, k = k1 + blockIdx.z;
  if (i > i2 || j > j2 || k > k2) return;
  float d = fminf(fminf(Ex[i], Ey[j]), Ez[k]);
  size_t p = ((size_t)k * ny + j) * nx + i;
  for (int c = 0; c < ncmp; c++) w[c * V + p] *= d;
}

// halo strips <-> contiguous message buffers, layout [ivar][k][j][i] per side as
// blk_macdrp_pack_mesg / unpack_mesg (forward/blk_t.c:576-808): only j in [nj1,nj2], k in [nk1,nk2]
// for x messages and i in [ni1,ni2] for y messages (no edges / corners).
// *flag = 1 when a[k][j][i] != 0 anywhere in i1..i2 x (j1 + blockIdx.y) x (k1 + blockIdx.z); a is a shifted, padded array
__global__ void k_any_nonzero(const float *a, int pitch, int ny, int i1, int i2, int j1, int k1, int *flag)
{
  const int i = i1 + blockIdx.x * blockDim.x + threadIdx.x;
  if (i > i2) return;
  const size_t p = ((size_t)(k1 + blockIdx.z) * ny + (j1 + blockIdx.y)) * pitch + i;
  if (a[p] != 0.0f) *flag = 1;
}

__global__ void k_halo_copy(float *w, float *buf, size_t V, int ncmp, int nx, int ny, int i1, int ni, int j1, int nj, int k1,
                            int nk, int unpack)
{
  size_t n = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t per = (size_t)ni * nj * nk;
  if (n >= per * ncmp) return;
  int c = n / per;
  size_t r = n % per;
  int ii = r % ni, jj = (r / ni) % nj, kk = r / ((size_t)ni * nj);
  size_t p = c * V + ((size_t)(k1 + kk) * ny + (j1 + jj)) * nx + (i1 + ii);
  if (unpack) w[p] = buf[n]; else buf[n] = w[p];
}

}  // namespace cgfd
