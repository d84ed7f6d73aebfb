// Per-point physics of one RHS evaluation, shared by the interior and the free-surface kernels.
// Expression order follows the reference so that results differ only by FMA contraction.
#pragma once
#include "cgfd_dev.cuh"

namespace cgfd {

struct Met {
  float xix, xiy, xiz, etx, ety, etz, ztx, zty, ztz;
};

// index-space derivatives of the 9 wavefield components along xi, eta, zeta
struct Deriv {
  float x[9], y[9], z[9];
};

__device__ __forceinline__ Met load_metric(const StageArgs &P, size_t p)
{
  Met m;
  m.xix = __ldg(P.metric[M_XIX] + p); m.xiy = __ldg(P.metric[M_XIY] + p); m.xiz = __ldg(P.metric[M_XIZ] + p);
  m.etx = __ldg(P.metric[M_ETX] + p); m.ety = __ldg(P.metric[M_ETY] + p); m.etz = __ldg(P.metric[M_ETZ] + p);
  m.ztx = __ldg(P.metric[M_ZTX] + p); m.zty = __ldg(P.metric[M_ZTY] + p); m.ztz = __ldg(P.metric[M_ZTZ] + p);
  return m;
}

// momentum equation, forward/sv_curv_col_el_iso.c:398-406 (same for every medium); uses the
// stress entries of d only
__device__ __forceinline__ void momentum(const Deriv &d, const Met &m, float slw, float *h)
{
  h[VX] = slw * (m.xix * d.x[TXX] + m.xiy * d.x[TXY] + m.xiz * d.x[TXZ]
               + m.etx * d.y[TXX] + m.ety * d.y[TXY] + m.etz * d.y[TXZ]
               + m.ztx * d.z[TXX] + m.zty * d.z[TXY] + m.ztz * d.z[TXZ]);
  h[VY] = slw * (m.xix * d.x[TXY] + m.xiy * d.x[TYY] + m.xiz * d.x[TYZ]
               + m.etx * d.y[TXY] + m.ety * d.y[TYY] + m.etz * d.y[TYZ]
               + m.ztx * d.z[TXY] + m.zty * d.z[TYY] + m.ztz * d.z[TYZ]);
  h[VZ] = slw * (m.xix * d.x[TXZ] + m.xiy * d.x[TYZ] + m.xiz * d.x[TZZ]
               + m.etx * d.y[TXZ] + m.ety * d.y[TYZ] + m.etz * d.y[TZZ]
               + m.ztx * d.z[TXZ] + m.zty * d.z[TYZ] + m.ztz * d.z[TZZ]);
}

// Hooke's law, isotropic, forward/sv_curv_col_el_iso.c:409-435; uses the velocity entries of d only
__device__ __forceinline__ void hooke_iso(const Deriv &d, const Met &m, float lam, float mu, float lam2mu, float *h)
{
  h[TXX] = lam2mu * (m.xix * d.x[VX] + m.etx * d.y[VX] + m.ztx * d.z[VX])
         + lam * (m.xiy * d.x[VY] + m.ety * d.y[VY] + m.zty * d.z[VY]
                + m.xiz * d.x[VZ] + m.etz * d.y[VZ] + m.ztz * d.z[VZ]);
  h[TYY] = lam2mu * (m.xiy * d.x[VY] + m.ety * d.y[VY] + m.zty * d.z[VY])
         + lam * (m.xix * d.x[VX] + m.etx * d.y[VX] + m.ztx * d.z[VX]
                + m.xiz * d.x[VZ] + m.etz * d.y[VZ] + m.ztz * d.z[VZ]);
  h[TZZ] = lam2mu * (m.xiz * d.x[VZ] + m.etz * d.y[VZ] + m.ztz * d.z[VZ])
         + lam * (m.xix * d.x[VX] + m.etx * d.y[VX] + m.ztx * d.z[VX]
                + m.xiy * d.x[VY] + m.ety * d.y[VY] + m.zty * d.z[VY]);
  h[TXY] = mu * (m.xiy * d.x[VX] + m.xix * d.x[VY] + m.ety * d.y[VX] + m.etx * d.y[VY] + m.zty * d.z[VX] + m.ztx * d.z[VY]);
  h[TXZ] = mu * (m.xiz * d.x[VX] + m.xix * d.x[VZ] + m.etz * d.y[VX] + m.etx * d.y[VZ] + m.ztz * d.z[VX] + m.ztx * d.z[VZ]);
  h[TYZ] = mu * (m.xiz * d.x[VY] + m.xiy * d.x[VZ] + m.etz * d.y[VY] + m.ety * d.y[VZ] + m.ztz * d.z[VY] + m.zty * d.z[VZ]);
}

// Runge-Kutta stage update fused into the RHS epilogue (forward/drv_rk_curv_col.c:292-446):
//   first : tmp = pre + a*rhs ; end  = pre + b*rhs      (w_cur == w_pre, its value is `cur_c`)
//   mid   : tmp = pre + a*rhs ; end += b*rhs
//   last  :                     end += b*rhs
// pre_v / end_v are the values of w_pre / w_end at this point, fetched by the caller ahead of time
// (first: both unused; last: pre_v unused).
template <int KIND>
__device__ __forceinline__ void rk_store(float *__restrict__ tmp, float *__restrict__ end, size_t off, float cur_c,
                                         float pre_v, float end_v, float rhs, float a, float b)
{
  if (KIND == KIND_FIRST) {
    tmp[off] = cur_c + a * rhs;
    end[off] = cur_c + b * rhs;
  } else if (KIND == KIND_MID) {
    tmp[off] = pre_v + a * rhs;
    end[off] = end_v + b * rhs;
  } else {
    end[off] = end_v + b * rhs;
  }
}

// the same update done in place in shared memory: sp holds w_pre (mid) and receives w_tmp, se holds w_end (mid, last)
// and receives the new w_end; the tiles are then written out by one TMA store each
template <int KIND>
__device__ __forceinline__ void rk_smem(float *sp, float *se, float cur_c, float rhs, float a, float b)
{
  if (KIND == KIND_FIRST) {
    *sp = cur_c + a * rhs;
    *se = cur_c + b * rhs;
  } else if (KIND == KIND_MID) {
    *sp = *sp + a * rhs;
    *se = *se + b * rhs;
  } else {
    *se = *se + b * rhs;
  }
}

// ADE CFS-PML correction of one face at one slab point, isotropic
// (forward/sv_curv_col_el_iso.c:763-905 for x, 915-1054 for y, 1058-1137 for z):
//   rhs_n  = RHS terms holding the face-normal derivative only
//   h     += (B-1)*rhs_n - B*aux ;  aux_rhs = D*rhs_n - A*aux
// with the free-surface terms at k == nk2 for x/y faces (:841-901, :989-1048), followed by the RK
// update of the auxiliary variables (forward/drv_rk_curv_col.c:315-346, 371-402, 426-438).
// PART 0 handles the 6 stress components (needs the velocity derivatives along the normal),
// PART 1 the 3 velocity components (needs the stress derivatives), so that a caller can finish one
// half of the RHS before it forms the other.
template <int AXIS, int KIND, int PART>
__device__ __forceinline__ void pml_face_iso(const StageArgs &P, const PmlFaceDev &F, int i, int j, int k,
                                             const Deriv &d, const Met &m, float lam, float mu, float lam2mu, float slw,
                                             float *h)
{
  constexpr int C0 = PART ? 0 : 3, C1 = PART ? 3 : 9;
  const int ia = (AXIS == 0) ? (i - F.i1) : (AXIS == 1) ? (j - F.j1) : (k - F.k1);
  const float cA = __ldg(F.A + ia), cB = __ldg(F.B + ia), cD = __ldg(F.D + ia);
  const float cB1 = cB - 1.0f;
  const float *D_ = (AXIS == 0) ? d.x : (AXIS == 1) ? d.y : d.z;
  const float e1 = (AXIS == 0) ? m.xix : (AXIS == 1) ? m.etx : m.ztx;
  const float e2 = (AXIS == 0) ? m.xiy : (AXIS == 1) ? m.ety : m.zty;
  const float e3 = (AXIS == 0) ? m.xiz : (AXIS == 1) ? m.etz : m.ztz;
  const size_t pa = ((size_t)(k - F.k1) * F.snj + (size_t)(j - F.j1)) * F.sni + (size_t)(i - F.i1);
  float au[9], pv[9], ev[9];
#pragma unroll
  for (int c = C0; c < C1; c++) {
    const size_t o = c * F.siz + pa;
    au[c] = __ldg(F.aux_cur + o);
    if (KIND == KIND_MID) pv[c] = __ldg(F.aux_pre + o);
    if (KIND != KIND_FIRST) ev[c] = F.aux_end[o];
  }
  float r[9];
  if (PART) {
    r[VX] = slw * (e1 * D_[TXX] + e2 * D_[TXY] + e3 * D_[TXZ]);
    r[VY] = slw * (e1 * D_[TXY] + e2 * D_[TYY] + e3 * D_[TYZ]);
    r[VZ] = slw * (e1 * D_[TXZ] + e2 * D_[TYZ] + e3 * D_[TZZ]);
  } else {
    r[TXX] = lam2mu * e1 * D_[VX] + lam * e2 * D_[VY] + lam * e3 * D_[VZ];
    r[TYY] = lam * e1 * D_[VX] + lam2mu * e2 * D_[VY] + lam * e3 * D_[VZ];
    r[TZZ] = lam * e1 * D_[VX] + lam * e2 * D_[VY] + lam2mu * e3 * D_[VZ];
    r[TXY] = mu * (e2 * D_[VX] + e1 * D_[VY]);
    r[TXZ] = mu * (e3 * D_[VX] + e1 * D_[VZ]);
    r[TYZ] = mu * (e3 * D_[VY] + e2 * D_[VZ]);
  }
  float ar[9];
#pragma unroll
  for (int c = C0; c < C1; c++) {
    h[c] += cB1 * r[c] - cB * au[c];
    ar[c] = cD * r[c] - cA * au[c];
  }
  if (PART == 0 && AXIS < 2 && P.free_top && k == P.nk2) {
    const float *M = ((AXIS == 0) ? P.matVx2Vz : P.matVy2Vz) + ((size_t)j * P.nx + i) * 9;
    float z0 = __ldg(M + 0) * D_[VX] + __ldg(M + 1) * D_[VY] + __ldg(M + 2) * D_[VZ];
    float z1 = __ldg(M + 3) * D_[VX] + __ldg(M + 4) * D_[VY] + __ldg(M + 5) * D_[VZ];
    float z2 = __ldg(M + 6) * D_[VX] + __ldg(M + 7) * D_[VY] + __ldg(M + 8) * D_[VZ];
    float q[9];
    q[TXX] = lam2mu * (m.ztx * z0) + lam * (m.zty * z1 + m.ztz * z2);
    q[TYY] = lam2mu * (m.zty * z1) + lam * (m.ztx * z0 + m.ztz * z2);
    q[TZZ] = lam2mu * (m.ztz * z2) + lam * (m.ztx * z0 + m.zty * z1);
    q[TXY] = mu * (m.zty * z0 + m.ztx * z1);
    q[TXZ] = mu * (m.ztz * z0 + m.ztx * z2);
    q[TYZ] = mu * (m.ztz * z1 + m.zty * z2);
#pragma unroll
    for (int c = 3; c < 9; c++) {
      h[c] += cB1 * q[c];
      ar[c] += cD * q[c];
    }
  }
#pragma unroll
  for (int c = C0; c < C1; c++) {
    const size_t o = c * F.siz + pa;
    if (KIND == KIND_FIRST) {
      F.aux_tmp[o] = au[c] + P.a * ar[c];
      F.aux_end[o] = au[c] + P.b * ar[c];
    } else if (KIND == KIND_MID) {
      F.aux_tmp[o] = pv[c] + P.a * ar[c];
      F.aux_end[o] = ev[c] + P.b * ar[c];
    } else {
      F.aux_end[o] = ev[c] + P.b * ar[c];
    }
  }
}

// all PML faces a point belongs to, in the reference's face order x1,x2,y1,y2,z1,z2
template <int KIND, int PART>
__device__ __forceinline__ void pml_all_iso(const StageArgs &P, int i, int j, int k, const Deriv &d, const Met &m,
                                            float lam, float mu, float lam2mu, float slw, float *h)
{
#pragma unroll
  for (int s = 0; s < 2; s++) {
    const PmlFaceDev &F = P.pml[0][s];
    if (F.on && i >= F.i1 && i <= F.i2) pml_face_iso<0, KIND, PART>(P, F, i, j, k, d, m, lam, mu, lam2mu, slw, h);
  }
#pragma unroll
  for (int s = 0; s < 2; s++) {
    const PmlFaceDev &F = P.pml[1][s];
    if (F.on && j >= F.j1 && j <= F.j2) pml_face_iso<1, KIND, PART>(P, F, i, j, k, d, m, lam, mu, lam2mu, slw, h);
  }
#pragma unroll
  for (int s = 0; s < 2; s++) {
    const PmlFaceDev &F = P.pml[2][s];
    if (F.on && k >= F.k1 && k <= F.k2) pml_face_iso<2, KIND, PART>(P, F, i, j, k, d, m, lam, mu, lam2mu, slw, h);
  }
}

}  // namespace cgfd
