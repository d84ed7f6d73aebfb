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

// The same two laws on a grid with xi_y = xi_z = eta_x = eta_z == 0 (GZ kernels): the terms that multiply those metrics are
// left out. Term order is kept; results agree with the general expressions fed with zeros up to the compiler's choice of FMA
// contraction (measured: rel L2 3e-7 after 24 steps, the same distance both have from the reference).
__device__ __forceinline__ void momentum_gz(const Deriv &d, const Met &m, float slw, float *h)
{
  h[VX] = slw * (m.xix * d.x[TXX] + m.ety * d.y[TXY] + m.ztx * d.z[TXX] + m.zty * d.z[TXY] + m.ztz * d.z[TXZ]);
  h[VY] = slw * (m.xix * d.x[TXY] + m.ety * d.y[TYY] + m.ztx * d.z[TXY] + m.zty * d.z[TYY] + m.ztz * d.z[TYZ]);
  h[VZ] = slw * (m.xix * d.x[TXZ] + m.ety * d.y[TYZ] + m.ztx * d.z[TXZ] + m.zty * d.z[TYZ] + m.ztz * d.z[TZZ]);
}
__device__ __forceinline__ void hooke_iso_gz(const Deriv &d, const Met &m, float lam, float mu, float lam2mu, float *h)
{
  h[TXX] = lam2mu * (m.xix * d.x[VX] + m.ztx * d.z[VX]) + lam * (m.ety * d.y[VY] + m.zty * d.z[VY] + m.ztz * d.z[VZ]);
  h[TYY] = lam2mu * (m.ety * d.y[VY] + m.zty * d.z[VY]) + lam * (m.xix * d.x[VX] + m.ztx * d.z[VX] + m.ztz * d.z[VZ]);
  h[TZZ] = lam2mu * (m.ztz * d.z[VZ]) + lam * (m.xix * d.x[VX] + m.ztx * d.z[VX] + m.ety * d.y[VY] + m.zty * d.z[VY]);
  h[TXY] = mu * (m.xix * d.x[VY] + m.ety * d.y[VX] + m.zty * d.z[VX] + m.ztx * d.z[VY]);
  h[TXZ] = mu * (m.xix * d.x[VZ] + m.ztz * d.z[VX] + m.ztx * d.z[VZ]);
  h[TYZ] = mu * (m.ety * d.y[VZ] + m.ztz * d.z[VY] + m.zty * d.z[VZ]);
}

// Wavefield update of the four RK stages with the reconstructed w_end (kinds in cgfd_dev.cuh). The reference accumulates
// w_end += b_s dt h_s in every stage (forward/drv_rk_curv_col.c:298-300, 330-332, 354-356, 386-388, 409-411); here the
// contributions of stages 0 and 2 are recovered one stage later from w_tmp - w_pre = a dt h, which removes one w_end write
// and two w_end reads per point and step (27 of 192 floats). Same value up to float32 round-off of the field itself.
template <int KIND>
// q = Graves' attenuation factor of the point, applied to the finished w_end (forward/drv_rk_curv_col.c:409-416); 1 otherwise
__device__ __forceinline__ void rk_wave(float *__restrict__ tmp, float *__restrict__ end, size_t off, float cur_c, float pre_v,
                                        float end_v, float rhs, float a, float b, float c, float q = 1.0f)
{
  if (KIND == KIND_FIRST) {
    tmp[off] = cur_c + a * rhs;
  } else if (KIND == KIND_MID) {
    tmp[off] = pre_v + a * rhs;
    end[off] = (pre_v + c * (cur_c - pre_v)) + b * rhs;
  } else if (KIND == KIND_THIRD) {
    tmp[off] = pre_v + a * rhs;
  } else {
    end[off] = ((end_v + c * (cur_c - pre_v)) + b * rhs) * q;
  }
}

// the same update done in place in shared memory: sp holds w_pre (all kinds but FIRST) and receives w_tmp, se holds w_end
// (LAST) and receives the new w_end (MID, LAST); the tiles are then written out by one TMA store each
template <int KIND>
__device__ __forceinline__ void rk_smem(float *sp, float *se, float cur_c, float rhs, float a, float b, float c, float q = 1.0f)
{
  if (KIND == KIND_FIRST) {
    *sp = cur_c + a * rhs;
  } else if (KIND == KIND_MID) {
    const float pv = *sp;
    *sp = pv + a * rhs;
    *se = (pv + c * (cur_c - pv)) + b * rhs;
  } else if (KIND == KIND_THIRD) {
    *sp = *sp + a * rhs;
  } else {
    *se = ((*se + c * (cur_c - *sp)) + b * rhs) * q;
  }
}

// ---------------------------------------------------------------------------------------------------------
// media. MED selects the constitutive law; the media arrays arrive in the order of include/cgfd3d_b200.h:
//   iso / visco : lambda, mu, 1/rho (+ Ylam[N], Ymu[N] for visco, read separately by the attenuation step)
//   vti         : c11, c13, c33, c55, c66, 1/rho          (c12 = c11 - 2 c66, forward/sv_curv_col_el_vti.c:344-352)
//   aniso       : the 21 Cij of the upper triangle row by row, 1/rho (forward/sv_curv_col_el_aniso.c:426-453)
// NTILE = arrays staged per plane by the interior kernel.
enum { MED_ISO = 0, MED_VTI = 1, MED_ANISO = 2, MED_VIS = 3 };

template <int MED> struct Med;
template <> struct Med<MED_ISO> {
  static constexpr int NTILE = 3;
  float lam, mu, lam2mu, slw;
  template <class F> __device__ __forceinline__ void load(F g) { lam = g(0); mu = g(1); slw = g(2); lam2mu = lam + 2.0f * mu; }
  template <class F> __device__ __forceinline__ void load_slw(F g) { lam = mu = lam2mu = 0.0f; slw = g(2); }
};
template <> struct Med<MED_VIS> : Med<MED_ISO> {};
template <> struct Med<MED_VTI> {
  static constexpr int NTILE = 6;
  float c11, c13, c33, c55, c66, c12, slw;
  template <class F> __device__ __forceinline__ void load(F g)
  {
    c11 = g(0); c13 = g(1); c33 = g(2); c55 = g(3); c66 = g(4); slw = g(5); c12 = c11 - 2.0f * c66;
  }
  template <class F> __device__ __forceinline__ void load_slw(F g) { c11 = c13 = c33 = c55 = c66 = c12 = 0.0f; slw = g(5); }
};
template <> struct Med<MED_ANISO> {
  static constexpr int NTILE = 22;
  float c[21], slw;
  template <class F> __device__ __forceinline__ void load(F g)
  {
#pragma unroll
    for (int n = 0; n < 21; n++) c[n] = g(n);
    slw = g(21);
  }
  template <class F> __device__ __forceinline__ void load_slw(F g)
  {
#pragma unroll
    for (int n = 0; n < 21; n++) c[n] = 0.0f;
    slw = g(21);
  }
};

// stress rate from a velocity gradient g[j][l] = dV_j/dx_l (or one axis' share of it): hT_I = sum_J C_IJ E_J with
// E = (g00, g11, g22, g12+g21, g02+g20, g01+g10). Same products as the reference's Hooke blocks
// (vti.c:366-389, aniso.c:466-491), which expand C_IJ e_l D V_j term by term; association differs, values agree to
// float32 round-off.
__device__ __forceinline__ void stress_from_grad(const float (&g)[3][3], const Med<MED_ISO> &M, float *h)
{
  h[TXX] = M.lam2mu * g[0][0] + M.lam * (g[1][1] + g[2][2]);
  h[TYY] = M.lam2mu * g[1][1] + M.lam * (g[0][0] + g[2][2]);
  h[TZZ] = M.lam2mu * g[2][2] + M.lam * (g[0][0] + g[1][1]);
  h[TXY] = M.mu * (g[0][1] + g[1][0]);
  h[TXZ] = M.mu * (g[0][2] + g[2][0]);
  h[TYZ] = M.mu * (g[1][2] + g[2][1]);
}
__device__ __forceinline__ void stress_from_grad(const float (&g)[3][3], const Med<MED_VTI> &M, float *h)
{
  h[TXX] = M.c11 * g[0][0] + M.c12 * g[1][1] + M.c13 * g[2][2];
  h[TYY] = M.c12 * g[0][0] + M.c11 * g[1][1] + M.c13 * g[2][2];
  h[TZZ] = M.c13 * g[0][0] + M.c13 * g[1][1] + M.c33 * g[2][2];
  h[TYZ] = M.c55 * (g[1][2] + g[2][1]);
  h[TXZ] = M.c55 * (g[0][2] + g[2][0]);
  h[TXY] = M.c66 * (g[0][1] + g[1][0]);
}
__device__ __forceinline__ void stress_from_grad(const float (&g)[3][3], const Med<MED_ANISO> &M, float *h)
{
  const float E[6] = {g[0][0], g[1][1], g[2][2], g[1][2] + g[2][1], g[0][2] + g[2][0], g[0][1] + g[1][0]};
  // upper triangle, row by row: row I starts at I*6 - I*(I-1)/2 and holds columns I..5
  const float *c = M.c;
#define CIJ(I, J) ((I) <= (J) ? c[(I) * 6 - (I) * ((I) - 1) / 2 + (J) - (I)] : c[(J) * 6 - (J) * ((J) - 1) / 2 + (I) - (J)])
#pragma unroll
  for (int I = 0; I < 6; I++)
    h[TXX + I] = CIJ(I, 0) * E[0] + CIJ(I, 1) * E[1] + CIJ(I, 2) * E[2] + CIJ(I, 3) * E[3] + CIJ(I, 4) * E[4] + CIJ(I, 5) * E[5];
#undef CIJ
}
static_assert(TXX == 3 && TYY == 4 && TZZ == 5 && TYZ == 6 && TXZ == 7 && TXY == 8, "stress components follow the Voigt order");

// Hooke's law on the velocity entries of d. The isotropic media keep the reference's own expression
// (forward/sv_curv_col_el_iso.c:409-435); vti / aniso go through the physical velocity gradient.
template <int MED, bool GZ = false> __device__ __forceinline__ void hooke(const Deriv &d, const Met &m, const Med<MED> &M, float *h)
{
  if constexpr (MED == MED_ISO || MED == MED_VIS) {
    if (GZ) hooke_iso_gz(d, m, M.lam, M.mu, M.lam2mu, h);
    else hooke_iso(d, m, M.lam, M.mu, M.lam2mu, h);
  } else {
    float g[3][3];
#pragma unroll
    for (int j = 0; j < 3; j++) {
      if (GZ) {
        g[j][0] = m.xix * d.x[j] + m.ztx * d.z[j];
        g[j][1] = m.ety * d.y[j] + m.zty * d.z[j];
        g[j][2] = m.ztz * d.z[j];
      } else {
        g[j][0] = m.xix * d.x[j] + m.etx * d.y[j] + m.ztx * d.z[j];
        g[j][1] = m.xiy * d.x[j] + m.ety * d.y[j] + m.zty * d.z[j];
        g[j][2] = m.xiz * d.x[j] + m.etz * d.y[j] + m.ztz * d.z[j];
      }
    }
    stress_from_grad(g, M, h);
  }
}

// stress rate caused by the derivative along ONE grid axis with direction cosines (e1,e2,e3): D_[VX..VZ] are the three
// velocity derivatives along that axis (PML normal terms, free-surface terms)
template <int MED>
__device__ __forceinline__ void stress_one_axis(const float *D_, float e1, float e2, float e3, const Med<MED> &M, float *r)
{
  if constexpr (MED == MED_ISO || MED == MED_VIS) {
    r[TXX] = M.lam2mu * e1 * D_[VX] + M.lam * e2 * D_[VY] + M.lam * e3 * D_[VZ];
    r[TYY] = M.lam * e1 * D_[VX] + M.lam2mu * e2 * D_[VY] + M.lam * e3 * D_[VZ];
    r[TZZ] = M.lam * e1 * D_[VX] + M.lam * e2 * D_[VY] + M.lam2mu * e3 * D_[VZ];
    r[TXY] = M.mu * (e2 * D_[VX] + e1 * D_[VY]);
    r[TXZ] = M.mu * (e3 * D_[VX] + e1 * D_[VZ]);
    r[TYZ] = M.mu * (e3 * D_[VY] + e2 * D_[VZ]);
  } else {
    float g[3][3];
#pragma unroll
    for (int j = 0; j < 3; j++) { g[j][0] = e1 * D_[j]; g[j][1] = e2 * D_[j]; g[j][2] = e3 * D_[j]; }
    stress_from_grad(g, M, r);
  }
}

// ADE CFS-PML correction of one face at one slab point
// (forward/sv_curv_col_el_iso.c:763-905 for x, 915-1054 for y, 1058-1137 for z; vti.c:592-1077, aniso.c:717-1259):
//   rhs_n  = RHS terms holding the face-normal derivative only
//   h     += (B-1)*rhs_n - B*aux ;  aux_rhs = D*rhs_n - A*aux
// with the free-surface terms at k == nk2 for x/y faces (iso.c:841-901, 989-1048), followed by the RK
// update of the auxiliary variables (forward/drv_rk_curv_col.c:315-346, 371-402, 426-438).
// PART 0 handles the 6 stress components (needs the velocity derivatives along the normal),
// PART 1 the 3 velocity components (needs the stress derivatives), so that a caller can finish one
// half of the RHS before it forms the other.
// n consecutive floats of an aux record (8-byte aligned: n = 6 -> three float2, n = 3 -> float2 + float)
template <int N> __device__ __forceinline__ void aux_load(const float *p, float *v)
{
  if (N == 6) {
    const float2 a = __ldg((const float2 *)p), b = __ldg((const float2 *)p + 1), c = __ldg((const float2 *)p + 2);
    v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y; v[4] = c.x; v[5] = c.y;
  } else {
    const float2 a = __ldg((const float2 *)p);
    v[0] = a.x; v[1] = a.y; v[2] = __ldg(p + 2);
  }
}
template <int N> __device__ __forceinline__ void aux_store(float *p, const float *v)
{
  if (N == 6) {
    ((float2 *)p)[0] = make_float2(v[0], v[1]); ((float2 *)p)[1] = make_float2(v[2], v[3]); ((float2 *)p)[2] = make_float2(v[4], v[5]);
  } else {
    // the pad float of the record is written along: every 32-byte sector of a level is then written in full, and the L2 never
    // has to fetch a sector from DRAM just to merge 4 unwritten bytes into it (ncu r1x: lts__t_sectors_data_ecc = 4.9 M sectors
    // per written level and launch, 0.16 GB of DRAM reads)
    ((float2 *)p)[0] = make_float2(v[0], v[1]); ((float2 *)p)[1] = make_float2(v[2], 0.0f);
  }
}

template <int AXIS, int KIND, int PART, int MED>
__device__ __forceinline__ void pml_face(const StageArgs &P, const PmlFaceDev &F, int i, int j, int k,
                                         const Deriv &d, const Met &m, const Med<MED> &M, float *h)
{
  constexpr int C0 = PART ? 0 : 3, N = PART ? 3 : 6;   // wavefield components C0 .. C0+N-1 = record slots aux_slot(C0) ..
  const int ia = (AXIS == 0) ? (i - F.i1) : (AXIS == 1) ? (j - F.j1) : (k - F.k1);
  const float cA = __ldg(F.A + ia), cB = __ldg(F.B + ia), cD = __ldg(F.D + ia);
  const float cB1 = cB - 1.0f;
  const float *D_ = (AXIS == 0) ? d.x : (AXIS == 1) ? d.y : d.z;
  const float e1 = (AXIS == 0) ? m.xix : (AXIS == 1) ? m.etx : m.ztx;
  const float e2 = (AXIS == 0) ? m.xiy : (AXIS == 1) ? m.ety : m.zty;
  const float e3 = (AXIS == 0) ? m.xiz : (AXIS == 1) ? m.etz : m.ztz;
  const size_t pa = (((size_t)(k - F.k1) * F.snj + (size_t)(j - F.j1)) * F.sni + (size_t)(i - F.i1)) * AUX_REC + aux_slot(C0);
  float au[N], pv[N], ev[N];
  aux_load<N>(F.aux_cur + pa, au);
  if (KIND != KIND_FIRST) aux_load<N>(F.aux_pre + pa, pv);
  if (KIND == KIND_LAST) aux_load<N>(F.aux_end + pa, ev);
  float r[9];
  if (PART) {
    r[VX] = M.slw * (e1 * D_[TXX] + e2 * D_[TXY] + e3 * D_[TXZ]);
    r[VY] = M.slw * (e1 * D_[TXY] + e2 * D_[TYY] + e3 * D_[TYZ]);
    r[VZ] = M.slw * (e1 * D_[TXZ] + e2 * D_[TYZ] + e3 * D_[TZZ]);
  } else {
    stress_one_axis<MED>(D_, e1, e2, e3, M, r);
  }
  float ar[N];
#pragma unroll
  for (int n = 0; n < N; n++) {
    h[C0 + n] += cB1 * r[C0 + n] - cB * au[n];
    ar[n] = cD * r[C0 + n] - cA * au[n];
  }
  if (PART == 0 && AXIS < 2 && P.free_top && k == P.nk2) {
    const float *Mt = ((AXIS == 0) ? P.matVx2Vz : P.matVy2Vz) + ((size_t)j * P.nx + i) * 9;
    float z[3];
    z[0] = __ldg(Mt + 0) * D_[VX] + __ldg(Mt + 1) * D_[VY] + __ldg(Mt + 2) * D_[VZ];
    z[1] = __ldg(Mt + 3) * D_[VX] + __ldg(Mt + 4) * D_[VY] + __ldg(Mt + 5) * D_[VZ];
    z[2] = __ldg(Mt + 6) * D_[VX] + __ldg(Mt + 7) * D_[VY] + __ldg(Mt + 8) * D_[VZ];
    float q[9];
    if constexpr (MED == MED_ISO || MED == MED_VIS) {
      q[TXX] = M.lam2mu * (m.ztx * z[0]) + M.lam * (m.zty * z[1] + m.ztz * z[2]);
      q[TYY] = M.lam2mu * (m.zty * z[1]) + M.lam * (m.ztx * z[0] + m.ztz * z[2]);
      q[TZZ] = M.lam2mu * (m.ztz * z[2]) + M.lam * (m.ztx * z[0] + m.zty * z[1]);
      q[TXY] = M.mu * (m.zty * z[0] + m.ztx * z[1]);
      q[TXZ] = M.mu * (m.ztz * z[0] + m.ztx * z[2]);
      q[TYZ] = M.mu * (m.ztz * z[1] + m.zty * z[2]);
    } else {
      stress_one_axis<MED>(z, m.ztx, m.zty, m.ztz, M, q);
    }
    if (PART == 0) {
#pragma unroll
      for (int n = 0; n < N; n++) {
        h[C0 + n] += cB1 * q[C0 + n];
        ar[n] += cD * q[C0 + n];
      }
    }
  }
  // the auxiliary variables advance with the same four-kind update as the wavefield (rk_wave)
  float vt[N], ve[N];
#pragma unroll
  for (int n = 0; n < N; n++) {
    if (KIND == KIND_FIRST) vt[n] = au[n] + P.a * ar[n];
    else if (KIND == KIND_MID) { vt[n] = pv[n] + P.a * ar[n]; ve[n] = (pv[n] + P.c * (au[n] - pv[n])) + P.b * ar[n]; }
    else if (KIND == KIND_THIRD) vt[n] = pv[n] + P.a * ar[n];
    else ve[n] = (ev[n] + P.c * (au[n] - pv[n])) + P.b * ar[n];
  }
  if (KIND != KIND_LAST) aux_store<N>(F.aux_tmp + pa, vt);
  if (KIND == KIND_MID || KIND == KIND_LAST) aux_store<N>(F.aux_end + pa, ve);
}

// the aux records plane k of the march will read, requested into L2 one plane ahead (no registers held; the loads of
// pml_face then see L2 latency instead of DRAM latency). mask = faces the point of plane k belongs to (pml_mask_xy | pml_mask_z):
// one bit test per face instead of six range tests -- the PML planes run at the same ~1.6 instructions per clock and SM as the
// plain ones, so what they cost is their instruction count (r2j source view: 1000 against 584 per warp-plane)
template <int KIND> __device__ __forceinline__ void pml_prefetch(const StageArgs &P, int mask, int i, int j, int k)
{
#pragma unroll
  for (int ax = 0; ax < 3; ax++) {
#pragma unroll
    for (int s = 0; s < 2; s++) {
      if (!(mask & (1 << (2 * ax + s)))) continue;
      const PmlFaceDev &F = P.pml[ax][s];
      const size_t pa = (((size_t)(k - F.k1) * F.snj + (size_t)(j - F.j1)) * F.sni + (size_t)(i - F.i1)) * AUX_REC;
      asm volatile("prefetch.global.L2 [%0];" ::"l"(F.aux_cur + pa));
      if (KIND != KIND_FIRST) asm volatile("prefetch.global.L2 [%0];" ::"l"(F.aux_pre + pa));
      if (KIND == KIND_LAST) asm volatile("prefetch.global.L2 [%0];" ::"l"(F.aux_end + pa));
    }
  }
}

// all PML faces a point belongs to, in the reference's face order x1,x2,y1,y2,z1,z2.
// mask = membership bits of the point (bit 2*axis + side), see pml_mask: the interior kernel forms the x / y bits once per thread
// (they do not change along its z march) and the z bits once per plane (uniform over the block), instead of twelve range tests per plane.
__device__ __forceinline__ int pml_mask_xy(const StageArgs &P, int i, int j)
{
  int m = 0;
#pragma unroll
  for (int s = 0; s < 2; s++) {
    const PmlFaceDev &Fx = P.pml[0][s], &Fy = P.pml[1][s];
    if (Fx.on && i >= Fx.i1 && i <= Fx.i2) m |= 1 << s;
    if (Fy.on && j >= Fy.j1 && j <= Fy.j2) m |= 4 << s;
  }
  return m;
}
__device__ __forceinline__ int pml_mask_z(const StageArgs &P, int k)
{
  int m = 0;
#pragma unroll
  for (int s = 0; s < 2; s++) {
    const PmlFaceDev &Fz = P.pml[2][s];
    if (Fz.on && k >= Fz.k1 && k <= Fz.k2) m |= 16 << s;
  }
  return m;
}
template <int KIND, int PART, int MED>
__device__ __forceinline__ void pml_masked(const StageArgs &P, int mask, int i, int j, int k, const Deriv &d, const Met &m, const Med<MED> &M,
                                           float *h)
{
  if (mask & 1) pml_face<0, KIND, PART, MED>(P, P.pml[0][0], i, j, k, d, m, M, h);
  if (mask & 2) pml_face<0, KIND, PART, MED>(P, P.pml[0][1], i, j, k, d, m, M, h);
  if (mask & 4) pml_face<1, KIND, PART, MED>(P, P.pml[1][0], i, j, k, d, m, M, h);
  if (mask & 8) pml_face<1, KIND, PART, MED>(P, P.pml[1][1], i, j, k, d, m, M, h);
  if (mask & 16) pml_face<2, KIND, PART, MED>(P, P.pml[2][0], i, j, k, d, m, M, h);
  if (mask & 32) pml_face<2, KIND, PART, MED>(P, P.pml[2][1], i, j, k, d, m, M, h);
}
template <int KIND, int PART, int MED>
__device__ __forceinline__ void pml_all(const StageArgs &P, int i, int j, int k, const Deriv &d, const Met &m, const Med<MED> &M,
                                        float *h)
{
  pml_masked<KIND, PART, MED>(P, pml_mask_xy(P, i, j) | pml_mask_z(P, k), i, j, k, d, m, M, h);
}

// Generalised-Maxwell-body attenuation (sv_curv_col_vis_iso_atten, forward/sv_curv_col_vis_iso.c:250-347) fused with the RK
// update of the 6*N memory variables (components 9 + 6n + {Jxx,Jyy,Jzz,Jyz,Jxz,Jxy}, forward/wav_t.c:138-172):
// strain rate from the stress RHS, hJ_n = wl_n (EV - J_n), hT -= lam sum Ylam_n tr(J_n) + 2 mu sum Ymu_n J_n.
// h[TXX..TXY] must hold the complete stress RHS (after PML and free-surface corrections); p = point offset.
template <int KIND>
__device__ __forceinline__ void atten_update(const StageArgs &P, size_t p, float lam, float mu, float *h)
{
  const float sum_hxyz = (h[TXX] + h[TYY] + h[TZZ]) / (3.0f * lam + 2.0f * mu);
  float EV[6];   // order of the J components: xx, yy, zz, yz, xz, xy
  EV[0] = ((2.0f * h[TXX] - h[TYY] - h[TZZ]) / (2.0f * mu) + sum_hxyz) / 3.0f;
  EV[1] = ((2.0f * h[TYY] - h[TXX] - h[TZZ]) / (2.0f * mu) + sum_hxyz) / 3.0f;
  EV[2] = ((2.0f * h[TZZ] - h[TXX] - h[TYY]) / (2.0f * mu) + sum_hxyz) / 3.0f;
  EV[3] = h[TYZ] / mu * 0.5f;
  EV[4] = h[TXZ] / mu * 0.5f;
  EV[5] = h[TXY] / mu * 0.5f;
  float sum_tr = 0.0f, sum[6] = {0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f};
  const int N = P.nmaxwell;
  for (int n = 0; n < N; n++) {
    const float ylam = __ldg(P.media[3 + n] + p), ymu = __ldg(P.media[3 + N + n] + p), wl = P.wl[n];
    float J[6];
#pragma unroll
    for (int q = 0; q < 6; q++) J[q] = __ldg(P.cur + (size_t)(9 + 6 * n + q) * P.siz_vol + p);
    sum_tr += ylam * (J[0] + J[1] + J[2]);
#pragma unroll
    for (int q = 0; q < 6; q++) {
      sum[q] += ymu * J[q];
      const float hJ = wl * (EV[q] - J[q]);
      const size_t o = (size_t)(9 + 6 * n + q) * P.siz_vol + p;
      const float pv = (KIND != KIND_FIRST) ? __ldg(P.pre + o) : 0.0f;
      const float ev = (KIND == KIND_LAST) ? P.end[o] : 0.0f;
      rk_wave<KIND>(P.tmp, P.end, o, J[q], pv, ev, hJ, P.a, P.b, P.c);
    }
  }
  h[TXX] -= lam * sum_tr + 2.0f * mu * sum[0];
  h[TYY] -= lam * sum_tr + 2.0f * mu * sum[1];
  h[TZZ] -= lam * sum_tr + 2.0f * mu * sum[2];
  h[TYZ] -= 2.0f * mu * sum[3];
  h[TXZ] -= 2.0f * mu * sum[4];
  h[TXY] -= 2.0f * mu * sum[5];
}

// The same attenuation step on tiles staged in shared memory by TMA (interior kernel, nmaxwell <= VIS_MAX_STAGED): jc = memory
// variables of the level this stage reads, jp = level n (in) -> w_tmp (out), je = accumulating level (in / out), y = Ylam[0..N-1],
// Ymu[0..N-1]; every pointer is this thread's entry of tile 0, tiles are nt floats apart; component 6 n + q of body n.
template <int KIND>
__device__ __forceinline__ void atten_smem(const float *jc, float *jp, float *je, const float *y, int nt, int N, const float *wl,
                                           float lam, float mu, float *h, float a, float b, float c)
{
  // strain rate from the stress RHS. The reference divides seven times per point (forward/sv_curv_col_vis_iso.c:299-309); here
  // the two reciprocals are formed once and multiplied (each quotient differs by <= 1 ulp: far inside the tolerance, and a
  // float division costs ~10 issue slots in a kernel that is issue bound)
  const float r2mu = 1.0f / (2.0f * mu), third = 1.0f / 3.0f;
  const float sum_hxyz = (h[TXX] + h[TYY] + h[TZZ]) / (3.0f * lam + 2.0f * mu);
  float EV[6];
  EV[0] = ((2.0f * h[TXX] - h[TYY] - h[TZZ]) * r2mu + sum_hxyz) * third;
  EV[1] = ((2.0f * h[TYY] - h[TXX] - h[TZZ]) * r2mu + sum_hxyz) * third;
  EV[2] = ((2.0f * h[TZZ] - h[TXX] - h[TYY]) * r2mu + sum_hxyz) * third;
  EV[3] = h[TYZ] * r2mu;
  EV[4] = h[TXZ] * r2mu;
  EV[5] = h[TXY] * r2mu;
  float sum_tr = 0.0f, sum[6] = {0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
  for (int n = 0; n < VIS_MAX_STAGED; n++) {
    if (n < N) {
      const float ylam = y[n * nt], ymu = y[(N + n) * nt], w = wl[n];
      float J[6];
#pragma unroll
      for (int q = 0; q < 6; q++) J[q] = jc[(6 * n + q) * nt];
      sum_tr += ylam * (J[0] + J[1] + J[2]);
#pragma unroll
      for (int q = 0; q < 6; q++) {
        sum[q] += ymu * J[q];
        rk_smem<KIND>(jp + (6 * n + q) * nt, je + (6 * n + q) * nt, J[q], w * (EV[q] - J[q]), a, b, c);
      }
    }
  }
  h[TXX] -= lam * sum_tr + 2.0f * mu * sum[0];
  h[TYY] -= lam * sum_tr + 2.0f * mu * sum[1];
  h[TZZ] -= lam * sum_tr + 2.0f * mu * sum[2];
  h[TYZ] -= 2.0f * mu * sum[3];
  h[TXZ] -= 2.0f * mu * sum[4];
  h[TXY] -= 2.0f * mu * sum[5];
}

}  // namespace cgfd
