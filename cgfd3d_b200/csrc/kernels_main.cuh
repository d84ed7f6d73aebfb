// RHS + CFS-PML + free surface + RK stage update, one fused pass per stage, for one medium (template parameter MED).
// Included by kernels_{iso,vti,aniso,vis}.cu, each of which instantiates one medium.
//
//   k_main_tma : all rows below the free-surface rows. One thread per (i,j) column of a 32x8 tile marching along z.
//                Every operand of a plane -- the 9 wavefield components with their x-y halo, 9 metric and the media
//                arrays, w_pre and w_end -- is brought into a 2-slot shared-memory ring by TMA (cp.async.bulk.tensor,
//                one elected thread, mbarrier completion), two planes ahead of the arithmetic; the zeta stencil lives
//                in a 5-deep register queue per component; the RK update is done in place in the ring slot and
//                written back by TMA stores.
//                Restates *_rhs_inner (forward/sv_curv_col_el_iso.c:208-441, vti.c:200-397, aniso.c:233-497),
//                *_rhs_cfspml (iso.c:644-1146, vti.c:592-1077, aniso.c:717-1259), sv_curv_col_vis_iso_atten
//                (forward/sv_curv_col_vis_iso.c:250-347) and the RK axpy loops of forward/drv_rk_curv_col.c:292-446
//                without ever writing the RHS to memory.
//   k_top      : the top four rows when the top is a free surface: traction-image momentum RHS
//                (sv_curv_col_el_rhs_timg_z2, forward/sv_curv_col_el.c:30-305), reduced-order / matrix Dz for the
//                stress RHS (*_rhs_vlow_z2, iso.c:451-634, vti.c:407-582, aniso.c:507-707), PML with its
//                free-surface terms, attenuation, RK update.
#pragma once
#include "physics.cuh"
#include "tma.cuh"

namespace cgfd {

extern __constant__ FdConst c_fd;

template <int D> struct Ofs {            // stencil offsets of direction index D
  static constexpr int first = D ? -3 : -1;
  static constexpr int left = D ? 3 : 1;   // points on the negative side
  static constexpr int right = D ? 1 : 3;
};

constexpr int TX = TILE_X, TY = TILE_Y;
constexpr int SX = TX + 4, SY = TY + 4;

// =============================================================================================
// interior kernel: TMA-fed shared-memory ring
// =============================================================================================
constexpr int NST = 2;                               // ring depth (planes in flight per block)
// TMA needs a 16-byte aligned start along x, so the halo tile carries 4 columns on either side of
// the 32 centre columns (the operators reach at most 3): pitch 40 floats, centre at column 4.
constexpr int HX = HALO_X;
constexpr int SXT = TX + 2 * HX;
constexpr int CEN_BYTES = TX * TY * 4;               // one centre tile of one array
constexpr int CUR_BYTES = 9 * SY * SXT * 4;          // 9 components with halo
constexpr int OFF_CUR = 0;
constexpr int OFF_MET = ((CUR_BYTES + 127) / 128) * 128;
constexpr int OFF_PRE = OFF_MET + 9 * CEN_BYTES;
constexpr int OFF_END = OFF_PRE + 9 * CEN_BYTES;
constexpr int OFF_MED = OFF_END + 9 * CEN_BYTES;
constexpr int blocks_by_smem(int stage_bytes) { return 233472 / (NST * stage_bytes + 128 /*alignment slack*/ + 64 /*barriers*/ + 1024 /*reserved*/); }
template <int MED> struct Lay {   // the media tiles come last: their number depends on the medium
  static constexpr int NMT = Med<MED>::NTILE;
  // visco-elastic medium: tiles of the memory variables (cur | pre -> tmp | end, 6 per Maxwell body each) and of Ylam, Ymu
  static constexpr int NJT = (MED == MED_VIS) ? 6 * VIS_MAX_STAGED : 0, NYT = (MED == MED_VIS) ? 2 * VIS_MAX_STAGED : 0;
  static constexpr int BY_REGS = (MED == MED_VIS) ? 1 : 65536 / (TILE_X * TILE_Y * 128);   // 128 registers per thread (visco: one block)
  static constexpr int clampb(int by_smem) { return by_smem < BY_REGS ? (by_smem < 1 ? 1 : by_smem) : BY_REGS; }
  // ZA tiles: the centre box of the NEXT plane of the march (9 components), i.e. the one value per component the one-sided zeta
  // operator needs from ahead. Staged by TMA with the plane's other operands where that keeps the number of resident blocks
  // (isotropic: 2 x 112 KB; general anisotropic: 1 block anyway); otherwise (VTI, visco-elastic) each thread fetches its 9 values
  // with plain loads one plane ahead. With ZA the zeta queue in registers is three planes deep instead of six.
  static constexpr int BASE_BYTES = OFF_MED + (NMT + 3 * NJT + NYT) * CEN_BYTES;
#ifndef CGFD_NO_ZA
  static constexpr bool ZA = blocks_by_smem(BASE_BYTES + 9 * CEN_BYTES) >= 1 && clampb(blocks_by_smem(BASE_BYTES + 9 * CEN_BYTES)) >= clampb(blocks_by_smem(BASE_BYTES));
#else
  static constexpr bool ZA = false;
#endif
  static constexpr int NZT = ZA ? 9 : 0;
  static constexpr int OFF_ZA = OFF_MED + NMT * CEN_BYTES;
  static constexpr int OFF_JC = OFF_ZA + NZT * CEN_BYTES, OFF_JP = OFF_JC + NJT * CEN_BYTES, OFF_JE = OFF_JP + NJT * CEN_BYTES,
                       OFF_Y = OFF_JE + NJT * CEN_BYTES;
  static constexpr int STAGE_BYTES = OFF_Y + NYT * CEN_BYTES;
  static constexpr int SMEM_BYTES = NST * STAGE_BYTES + 128 /*alignment slack*/ + 64 /*barriers*/;
  // blocks per SM the shared memory allows (227 KB usable, 1 KB reserved per block)
  static constexpr int BY_SMEM = 233472 / (SMEM_BYTES + 1024);
  static constexpr int BLOCKS = clampb(BY_SMEM);
  // one resident block only: two thread groups per tile (stress half / velocity half of the RHS), 512 threads
#ifndef CGFD_NO_SPLIT
  static constexpr bool SPLIT = (BLOCKS == 1);
#else
  static constexpr bool SPLIT = false;
#endif
};

template <int KIND, int MED, bool GZ> __device__ __forceinline__ constexpr uint32_t stage_tx_bytes()
{
  return CUR_BYTES + ((GZ ? 5 : 9) + Lay<MED>::NMT + Lay<MED>::NZT) * CEN_BYTES + (KIND != KIND_FIRST ? 9 * CEN_BYTES : 0) +
         (KIND == KIND_LAST ? 9 * CEN_BYTES : 0);
}
// bytes of the staged memory-variable / Y tiles of one plane (visco-elastic medium, nj = 6 N tiles per level, 2 N Y tiles)
template <int KIND> __device__ __forceinline__ uint32_t vis_tx_bytes(int nmx)
{
  return (uint32_t)(nmx * CEN_BYTES) * (6 + 2 + (KIND != KIND_FIRST ? 6 : 0) + (KIND == KIND_LAST ? 6 : 0));
}

struct TmaCtx {
  unsigned char *ring;
  uint64_t *full;
  int tx, ty, t, i, j, i0, j0, k1;
  int pf;              // L2 prefetch distance in planes beyond the ring (0 = off)
  int nmx;             // visco-elastic medium: Maxwell bodies staged by TMA (0 = none: atten_update reads them with plain loads)
  int fmask;           // x / y PML faces this thread's column lies in (pml_mask_xy)
  bool active, inarr;
  size_t pij;
  const float *qptr;   // w_cur + pij: this thread's column of component 0
  uint64_t pol_keep, pol_stream;   // L2 eviction policies (thread 0 only)
};

// Loads of one plane into ring slot s, in two parts: A = the tiles nothing is stored from (wavefield with halo, metric,
// media; also arms the barrier with the byte count of the whole plane), B = the w_pre / w_end tiles, which double as the
// sources of the TMA stores of the plane that used the slot before and can only be refilled once those have been read.
template <int DX, int DY, int DZ, int KIND, int MED, bool GZ>
__device__ __forceinline__ void tma_issue_a(const StageArgs &P, const TmaMaps &M, const TmaCtx &C, int kk, int s)
{
  constexpr int YL = Ofs<DY>::left;
  unsigned char *b = C.ring + s * Lay<MED>::STAGE_BYTES;
  uint64_t *bar = C.full + s;
  uint32_t txb = stage_tx_bytes<KIND, MED, GZ>();
  if constexpr (MED == MED_VIS) txb += vis_tx_bytes<KIND>(C.nmx);
  mbar_expect_tx(bar, txb);
  // the wavefield tiles overlap their neighbours' (x-y halo): keep them in L2; everything else is touched once per stage
  tma_load_4d_hint(b + OFF_CUR, &M.cur, bar, C.i0 - HX + P.shift, C.j0 - YL, kk, 0, C.pol_keep);
  tma_load_4d_hint(b + OFF_MET, GZ ? &M.met5 : &M.met, bar, C.i0 + P.shift, C.j0, kk, 0, C.pol_stream);
  tma_load_4d_hint(b + OFF_MED, &M.med, bar, C.i0 + P.shift, C.j0, kk, 0, C.pol_stream);
  // the centre box of the plane ahead (an L2 hit now or for the halo box of that plane one iteration on)
  if constexpr (Lay<MED>::ZA) tma_load_4d_hint(b + Lay<MED>::OFF_ZA, &M.za, bar, C.i0 + P.shift, C.j0, kk + (DZ ? 1 : -1), 0, C.pol_keep);
  if constexpr (MED == MED_VIS) {
    if (C.nmx > 0) {
      tma_load_4d_hint(b + Lay<MED>::OFF_JC, &M.jcur, bar, C.i0 + P.shift, C.j0, kk, 0, C.pol_stream);
      tma_load_4d_hint(b + Lay<MED>::OFF_Y, &M.ymed, bar, C.i0 + P.shift, C.j0, kk, 0, C.pol_stream);
    }
  }
}
template <int DX, int DY, int KIND, int MED>
__device__ __forceinline__ void tma_issue_b(const StageArgs &P, const TmaMaps &M, const TmaCtx &C, int kk, int s)
{
  unsigned char *b = C.ring + s * Lay<MED>::STAGE_BYTES;
  uint64_t *bar = C.full + s;
  if (KIND != KIND_FIRST) tma_load_4d_hint(b + OFF_PRE, &M.pre, bar, C.i0 + P.shift, C.j0, kk, 0, C.pol_stream);
  if (KIND == KIND_LAST) tma_load_4d_hint(b + OFF_END, &M.end, bar, C.i0 + P.shift, C.j0, kk, 0, C.pol_stream);
  if constexpr (MED == MED_VIS) {
    if (C.nmx > 0) {
      if (KIND != KIND_FIRST) tma_load_4d_hint(b + Lay<MED>::OFF_JP, &M.jpre, bar, C.i0 + P.shift, C.j0, kk, 0, C.pol_stream);
      if (KIND == KIND_LAST) tma_load_4d_hint(b + Lay<MED>::OFF_JE, &M.jend, bar, C.i0 + P.shift, C.j0, kk, 0, C.pol_stream);
    }
  }
}
// every operand of plane kk into L2 (thread 0, PF planes ahead of the ring refill): the ring holds only NST planes per block,
// too few bytes in flight to cover DRAM latency; the L2 has room for several more
template <int DX, int DY, int KIND, int MED, bool GZ>
__device__ __forceinline__ void tma_prefetch_plane(const StageArgs &P, const TmaMaps &M, const TmaCtx &C, int kk)
{
  constexpr int YL = Ofs<DY>::left;
  tma_prefetch_4d(&M.cur, C.i0 - HX + P.shift, C.j0 - YL, kk, 0);
  tma_prefetch_4d(GZ ? &M.met5 : &M.met, C.i0 + P.shift, C.j0, kk, 0);
  tma_prefetch_4d(&M.med, C.i0 + P.shift, C.j0, kk, 0);
  if (KIND != KIND_FIRST) tma_prefetch_4d(&M.pre, C.i0 + P.shift, C.j0, kk, 0);
  if (KIND == KIND_LAST) tma_prefetch_4d(&M.end, C.i0 + P.shift, C.j0, kk, 0);
}
template <int DX, int DY, int DZ, int KIND, int MED, bool GZ>
__device__ __forceinline__ void tma_issue(const StageArgs &P, const TmaMaps &M, const TmaCtx &C, int kk, int s)
{
  tma_issue_a<DX, DY, DZ, KIND, MED, GZ>(P, M, C, kk, s);
  tma_issue_b<DX, DY, KIND, MED>(P, M, C, kk, s);
}

// ---------------------------------------------------------------------------------------------------------
// Free-surface rows inside the interior kernel (StageArgs::fuse_top): the planes k >= nk2 - 3 of the top z chunk run the plane
// function with TOP = true. Everything a plane needs arrives through the same TMA ring; what differs from a row below:
//   * stress half: the zeta derivative of the velocity comes from the surface matrices (k = nk2), the 2-point (nk2 - 1) or the
//     3-point operator (nk2 - 2) -- *_rhs_vlow_z2, forward/sv_curv_col_el_iso.c:503-593 -- all of whose rows sit in the zeta queue;
//   * velocity half, rows whose zeta stencil reaches the surface: the momentum RHS in conservative form with the anti-symmetric
//     traction image (sv_curv_col_el_rhs_timg_z2, forward/sv_curv_col_el.c:84-304). The stress values of the fluxes come from
//     the halo tile (xi, eta) and the zeta queue, the metric of the neighbour points through L1 (40 - 52 loads per point on 2 - 4
//     of the ~200 planes of a column).
template <int DZ, int MED, int QN>
__device__ __forceinline__ void top_vlow(const StageArgs &P, int i, int j, int k, Deriv &d, const float (&q1)[QN], const float (&q2)[QN],
                                         const float (&q3)[QN])
{
  const int nsurf = P.nk2 - k;   // 0 at the surface
  if (nsurf == 0) {
    // the surface point-force term exists in the isotropic operator only (iso.c:583-592)
    constexpr bool FSRC = (MED == MED_ISO || MED == MED_VIS);
    const size_t p2 = (size_t)j * P.nx + i;
    const float *A = P.matVx2Vz + p2 * 9, *B = P.matVy2Vz + p2 * 9, *F = P.matF2Vz + p2 * 9;
    float sx = (FSRC && P.VxSrc) ? __ldg(P.VxSrc + p2) : 0.0f, sy = (FSRC && P.VySrc) ? __ldg(P.VySrc + p2) : 0.0f,
          sz = (FSRC && P.VzSrc) ? __ldg(P.VzSrc + p2) : 0.0f;
#pragma unroll
    for (int r = 0; r < 3; r++) {
      float v = __ldg(A + 3 * r + 0) * d.x[VX] + __ldg(A + 3 * r + 1) * d.x[VY] + __ldg(A + 3 * r + 2) * d.x[VZ]
              + __ldg(B + 3 * r + 0) * d.y[VX] + __ldg(B + 3 * r + 1) * d.y[VY] + __ldg(B + 3 * r + 2) * d.y[VZ];
      if (FSRC) v += __ldg(F + 3 * r + 0) * sx + __ldg(F + 3 * r + 1) * sy + __ldg(F + 3 * r + 2) * sz;
      d.z[r] = v;
    }
  } else if (nsurf == 1) {
    // rows (k-1, k) for the backward operator, (k, k+1) for the forward one: q2 is the plane just visited, q3 this plane
#pragma unroll
    for (int c = 0; c < 3; c++)
      d.z[c] = DZ ? c_fd.lay2[1][0] * q2[c] + c_fd.lay2[1][1] * q3[c] : c_fd.lay2[0][0] * q3[c] + c_fd.lay2[0][1] * q2[c];
  } else if (nsurf == 2) {
#pragma unroll
    for (int c = 0; c < 3; c++)
      d.z[c] = DZ ? c_fd.lay3[1][0] * q1[c] + c_fd.lay3[1][1] * q2[c] + c_fd.lay3[1][2] * q3[c]
                  : c_fd.lay3[0][0] * q3[c] + c_fd.lay3[0][1] * q2[c] + c_fd.lay3[0][2] * q1[c];
  }
}

// sc = this thread's centre entry of component 0 in the halo tile of plane k; the queue entries hold components QB .. QB+QN-1
// (QB <= TXX); p = offset of the point in a 3-D array
template <int DX, int DY, int DZ, int QB, int QN>
__device__ __forceinline__ void timg_smem(const StageArgs &P, size_t p, int i, int j, int k, const float *sc, float slw,
                                          const float (&q0)[QN], const float (&q1)[QN], const float (&q2)[QN], const float (&q3)[QN],
                                          const float (&q4)[QN], float *h)
{
  constexpr int FX = Ofs<DX>::first, FY = Ofs<DY>::first, FZ = Ofs<DZ>::first;
  const float *cx = c_fd.coef[DX], *cy = c_fd.coef[DY], *cz = c_fd.coef[DZ];
  const long L = (long)P.siz_line, S = (long)P.siz_slice;
  const size_t V = P.siz_vol;
  const int n_free = P.nk2 - k - FZ;
  const float jac = __ldg(P.metric[M_JAC] + p);
  const float slwjac = slw / jac;
  // component triplets (T1,T2,T3) with flux_n = J*(e_x T1 + e_y T2 + e_z T3)
  constexpr int T1[3] = {TXX, TXY, TXZ}, T2[3] = {TXY, TYY, TYZ}, T3[3] = {TXZ, TYZ, TZZ};
  const float *Ts[3] = {P.TxSrc, P.TySrc, P.TzSrc};
  const size_t p2 = (size_t)j * P.nx + i;
  // stencil row n of the zeta operator is the plane k + FZ + n: queue entry n when the march runs upwards, 4 - n downwards
  auto zq = [&](int n, int c) -> float {
    const int m = DZ ? n : 4 - n;
    return m == 0 ? q0[c - QB] : m == 1 ? q1[c - QB] : m == 2 ? q2[c - QB] : m == 3 ? q3[c - QB] : q4[c - QB];
  };
  // the xi / eta flux derivatives are accumulated term by term (left to right, like M_FD_NOINDX, forward/fd_t.h:33-38);
  // only the zeta fluxes are kept, because the image terms refer back to them
  float Dxf[3], Dyf[3], fz[3][5];
#pragma unroll
  for (int n = 0; n < 5; n++) {
    {
      const size_t pp = p + (FX + n);
      const float jn = __ldg(P.metric[M_JAC] + pp), ex = __ldg(P.metric[M_XIX] + pp), ey = __ldg(P.metric[M_XIY] + pp),
                  ez = __ldg(P.metric[M_XIZ] + pp);
#pragma unroll
      for (int v = 0; v < 3; v++) {
        const float f = jn * (ex * sc[T1[v] * SY * SXT + FX + n] + ey * sc[T2[v] * SY * SXT + FX + n] + ez * sc[T3[v] * SY * SXT + FX + n]);
        if (n == 0) Dxf[v] = cx[0] * f; else Dxf[v] += cx[n] * f;
      }
    }
    {
      const size_t pp = p + (long)(FY + n) * L;
      const float jn = __ldg(P.metric[M_JAC] + pp), ex = __ldg(P.metric[M_ETX] + pp), ey = __ldg(P.metric[M_ETY] + pp),
                  ez = __ldg(P.metric[M_ETZ] + pp);
#pragma unroll
      for (int v = 0; v < 3; v++) {
        const float f = jn * (ex * sc[T1[v] * SY * SXT + (FY + n) * SXT] + ey * sc[T2[v] * SY * SXT + (FY + n) * SXT]
                            + ez * sc[T3[v] * SY * SXT + (FY + n) * SXT]);
        if (n == 0) Dyf[v] = cy[0] * f; else Dyf[v] += cy[n] * f;
      }
    }
    if (n < n_free) {
      const size_t pp = p + (long)(FZ + n) * S;
      const float jn = __ldg(P.metric[M_JAC] + pp), ex = __ldg(P.metric[M_ZTX] + pp), ey = __ldg(P.metric[M_ZTY] + pp),
                  ez = __ldg(P.metric[M_ZTZ] + pp);
#pragma unroll
      for (int v = 0; v < 3; v++) fz[v][n] = jn * (ex * zq(n, T1[v]) + ey * zq(n, T2[v]) + ez * zq(n, T3[v]));
    } else {
#pragma unroll
      for (int v = 0; v < 3; v++) fz[v][n] = 0.0f;
    }
  }
#pragma unroll
  for (int v = 0; v < 3; v++) {
    const float ts = Ts[v] ? __ldg(Ts[v] + p2) : 0.0f;
#pragma unroll
    for (int n = 0; n < 5; n++) {
      if (n == n_free) fz[v][n] = ts;
      else if (n > n_free) {
        const int im = 2 * n_free - n;    // mirror index inside the window
        float below;
        if (im >= 0) below = (im == 0) ? fz[v][0] : (im == 1) ? fz[v][1] : fz[v][2];   // im <= 2 (n_free = 3, n = 4); no dynamic indexing
        else if (P.timg_mode == 0) below = 0.0f;
        else {
          // image row k + indx[n] - 2(n - n_free) (sv_curv_col_el.c:154-159): below the window, hence outside the queue
          const size_t pp = p + (long)(FZ + n - 2 * (n - n_free)) * S;
          const float jn = __ldg(P.metric[M_JAC] + pp), ex = __ldg(P.metric[M_ZTX] + pp), ey = __ldg(P.metric[M_ZTY] + pp),
                      ez = __ldg(P.metric[M_ZTZ] + pp);
          below = jn * (ex * __ldg(P.cur + T1[v] * V + pp) + ey * __ldg(P.cur + T2[v] * V + pp) + ez * __ldg(P.cur + T3[v] * V + pp));
        }
        fz[v][n] = 2.0f * ts - below;
      }
    }
    float Dz = cz[0] * fz[v][0]; Dz += cz[1] * fz[v][1]; Dz += cz[2] * fz[v][2]; Dz += cz[3] * fz[v][3]; Dz += cz[4] * fz[v][4];
    h[v] = (Dxf[v] + Dyf[v] + Dz) * slwjac;
  }
}

// One plane. The march along z runs TOWARDS the short side of the one-sided zeta operator (upwards for
// offsets {-3..1}, downwards for {-1..3}), so that three of its neighbours are planes already visited and only one
// lies ahead: q0,q1,q2 = planes k-3d, k-2d, k-d (d = march direction), q3 = plane k, q4 receives plane k+d, which was
// requested one iteration earlier (qn) -- at the same time as the TMA of that plane, so it is one DRAM read.
//
// PART: which half of the RHS this thread forms. 0 = both (one thread per column: media with two resident blocks per SM);
// 1 = the stress half (velocity derivatives -> Hooke -> PML part 0 -> attenuation -> RK update of the six stress components),
// 2 = the velocity half (stress derivatives -> momentum -> PML part 1 -> RK update of the three velocity components).
// The halves share nothing but the read-only tiles of the ring slot and write disjoint components, so media whose shared-memory
// footprint leaves room for ONE block per SM (general anisotropic: 22 media tiles; visco-elastic: the staged memory variables) run
// blocks of two thread groups on the same tile -- 16 warps per SM instead of 8, each with a zeta queue of its own 3 / 6 components.
template <int PART> struct Part {
  static constexpr int QB = (PART == 2) ? 3 : 0;                          // first component of the zeta queue
  static constexpr int QN = (PART == 0) ? 9 : (PART == 1) ? 3 : 6;        // components in the queue = components differentiated
};
template <int DX, int DY, int DZ, int KIND, int MED, bool GZ, bool PML, int PART, bool TOP = false>
__device__ __forceinline__ void tma_plane(const StageArgs &P, const TmaMaps &M, const TmaCtx &C, int k, int it, int nplanes,
                                          float (&q0)[Part<PART>::QN], float (&q1)[Part<PART>::QN], float (&q2)[Part<PART>::QN],
                                          float (&q3)[Part<PART>::QN], float (&q4)[Part<PART>::QN], float (&qn)[Part<PART>::QN])
{
  constexpr int YL = Ofs<DY>::left;
  constexpr int FX = Ofs<DX>::first, FY = Ofs<DY>::first;
  constexpr int DIR = DZ ? 1 : -1;
  constexpr int NT = TX * TY;
  constexpr int QB = Part<PART>::QB, QN = Part<PART>::QN;
  const int s = it % NST;
  const uint32_t parity = (it / NST) & 1;
  const float *cx = c_fd.coef[DX], *cy = c_fd.coef[DY], *cz = c_fd.coef[DZ];
  constexpr bool ZA = Lay<MED>::ZA;   // q3 / q4 of this plane come from shared memory (below, after the wait)
  if constexpr (!ZA) {
#pragma unroll
    for (int c = 0; c < QN; c++) q4[c] = qn[c];
  }
  if (!ZA && C.inarr && it + 1 < nplanes && !(P.l2mode & 128)) {   // bit 7: DIAGNOSTIC (wrong results): no z-ahead loads
    const float *w = C.qptr + (long)(k + 2 * DIR) * (long)P.siz_slice;
#pragma unroll
    for (int c = 0; c < QN; c++) qn[c] = __ldg(w + (QB + c) * P.siz_vol);
  }
  // l2mode bit 9: no prefetch of the aux records (the loads are not what the PML planes wait for, r2i; the prefetch costs instructions)
  if (PML && PART != 2 && C.active && it + 1 < nplanes && !(P.l2mode & 512)) pml_prefetch<KIND>(P, C.fmask | pml_mask_z(P, k + DIR), C.i, C.j, k + DIR);
  const int pmask = PML ? (C.fmask | pml_mask_z(P, k)) : 0;
  // Graves' attenuation factor of this point (last stage only; requested before the wait below so that it is there in time)
  float qatt = 1.0f;
  if (KIND == KIND_LAST && P.qatt && C.active) qatt = __ldg(P.qatt + (size_t)k * P.siz_slice + C.pij);
  unsigned char *b = C.ring + s * Lay<MED>::STAGE_BYTES;
  mbar_wait(C.full + s, parity);
  if (C.active) {
    const float *sc = (const float *)(b + OFF_CUR) + (C.ty + YL) * SXT + C.tx + HX;
    const float *sm = (const float *)(b + OFF_MET) + C.t;
    const float *sd = (const float *)(b + OFF_MED) + C.t;
    float *sp = (float *)(b + OFF_PRE) + C.t;   // w_pre in, w_tmp out
    float *se = (float *)(b + OFF_END) + C.t;   // w_end in, w_end out
    Met m;   // tiles in device order (metric_dev_slot)
    m.xix = sm[0 * NT]; m.ety = sm[1 * NT]; m.ztx = sm[2 * NT]; m.zty = sm[3 * NT]; m.ztz = sm[4 * NT];
    if (GZ) { m.xiy = 0.0f; m.xiz = 0.0f; m.etx = 0.0f; m.etz = 0.0f; }
    else { m.xiy = sm[5 * NT]; m.xiz = sm[6 * NT]; m.etx = sm[7 * NT]; m.etz = sm[8 * NT]; }
    Med<MED> md;
    if (PART == 2) md.load_slw([&](int n) { return sd[n * NT]; });   // the velocity half needs 1/rho only
    else md.load([&](int n) { return sd[n * NT]; });
    const float slw = md.slw;
    Deriv d;
    float h[9];
    if constexpr (ZA) {
      // this plane's own values and the plane ahead: the centre of the halo tile and the ZA tile (q3 then moves into the register
      // queue, which holds the three planes behind)
      const float *za = (const float *)(b + Lay<MED>::OFF_ZA) + C.t;
#pragma unroll
      for (int c = 0; c < QN; c++) { q3[c] = sc[(QB + c) * SY * SXT]; q4[c] = za[(QB + c) * NT]; }
    }
    // derivatives of the components this thread differentiates (queue entry c - QB holds component c)
#define CGFD_DERIV(c)                                                                                                             \
    {                                                                                                                             \
      const float *r = sc + (c) * SY * SXT;                                                                                       \
      d.x[c] = cx[0] * r[FX] + cx[1] * r[FX + 1] + cx[2] * r[FX + 2] + cx[3] * r[FX + 3] + cx[4] * r[FX + 4];                     \
      d.y[c] = cy[0] * r[FY * SXT] + cy[1] * r[(FY + 1) * SXT] + cy[2] * r[(FY + 2) * SXT] + cy[3] * r[(FY + 3) * SXT] + cy[4] * r[(FY + 4) * SXT]; \
      d.z[c] = DZ ? cz[0] * q0[(c) - QB] + cz[1] * q1[(c) - QB] + cz[2] * q2[(c) - QB] + cz[3] * q3[(c) - QB] + cz[4] * q4[(c) - QB]  \
                  : cz[0] * q4[(c) - QB] + cz[1] * q3[(c) - QB] + cz[2] * q2[(c) - QB] + cz[3] * q1[(c) - QB] + cz[4] * q0[(c) - QB]; \
    }
    if (PART != 2) {
      // ---- stress half: needs the velocity derivatives only
#pragma unroll
      for (int c = 0; c < 3; c++) CGFD_DERIV(c)
      if (TOP) top_vlow<DZ, MED>(P, C.i, C.j, k, d, q1, q2, q3);   // velocity components are queue entries 0..2 of both PART 0 and 1
      hooke<MED, GZ>(d, m, md, h);
      if (PML) pml_masked<KIND, 0, MED>(P, pmask, C.i, C.j, k, d, m, md, h);
      if constexpr (MED == MED_VIS) {
        if (C.nmx > 0)
          atten_smem<KIND>((const float *)(b + Lay<MED>::OFF_JC) + C.t, (float *)(b + Lay<MED>::OFF_JP) + C.t, (float *)(b + Lay<MED>::OFF_JE) + C.t,
                           (const float *)(b + Lay<MED>::OFF_Y) + C.t, NT, C.nmx, P.wl, md.lam, md.mu, h, P.a, P.b, P.c);
        else atten_update<KIND>(P, (size_t)k * P.siz_slice + C.pij, md.lam, md.mu, h);
      }
      // the centre value of a stress component: in the queue of a thread that differentiates it, else in the halo tile
#pragma unroll
      for (int c = 3; c < 9; c++) rk_smem<KIND>(sp + c * NT, se + c * NT, PART == 0 ? q3[c - QB] : sc[c * SY * SXT], h[c], P.a, P.b, P.c, qatt);
    }
    if (PART != 1) {
      // ---- velocity half: needs the stress derivatives only
#pragma unroll
      for (int c = 3; c < 9; c++) CGFD_DERIV(c)
      // rows whose zeta stencil reaches the free surface (block-uniform): traction image instead of the plain momentum RHS
      if (TOP && k >= P.nk2 - (Ofs<DZ>::first + 4))
        timg_smem<DX, DY, DZ, QB, QN>(P, (size_t)k * P.siz_slice + C.pij, C.i, C.j, k, sc, slw, q0, q1, q2, q3, q4, h);
      else if (GZ) momentum_gz(d, m, slw, h);
      else momentum(d, m, slw, h);
      if (PML) pml_masked<KIND, 1, MED>(P, pmask, C.i, C.j, k, d, m, md, h);
#pragma unroll
      for (int c = 0; c < 3; c++) rk_smem<KIND>(sp + c * NT, se + c * NT, PART == 0 ? q3[c - QB] : sc[c * SY * SXT], h[c], P.a, P.b, P.c, qatt);
    }
#undef CGFD_DERIV
    if constexpr (ZA) {
      // the register queue moves on here, inside the branch that filled q3: q3 / q4 are dead between planes
#pragma unroll
      for (int c = 0; c < QN; c++) { q0[c] = q1[c]; q1[c] = q2[c]; q2[c] = q3[c]; }
    }
    fence_proxy_async_smem();   // the results written above are read by the TMA store below
  } else {
    // Columns / rows of the tile beyond the physical range. The TMA store clips at the tensor extent, but in units of 16 bytes:
    // when ni is not a multiple of 4 the last granule carries up to 3 ghost columns along (measured: ni = 70 -> columns 70, 71
    // written). Ghost values of the result must be zero here (physical face) or are overwritten by the halo exchange that
    // follows (inter-rank face), so these threads clear their tile entries instead of leaving stale shared memory in them.
    float *sp = (float *)(b + OFF_PRE) + C.t, *se = (float *)(b + OFF_END) + C.t;
    constexpr int CB = (PART == 1) ? 3 : 0, CE = (PART == 2) ? 3 : 9;   // the components this thread's half writes
#pragma unroll
    for (int c = CB; c < CE; c++) {
      if (KIND != KIND_LAST) sp[c * NT] = 0.0f;
      if (KIND == KIND_MID || KIND == KIND_LAST) se[c * NT] = 0.0f;
    }
    if constexpr (MED == MED_VIS) {
      if (PART != 2) {
        float *jp = (float *)(b + Lay<MED>::OFF_JP) + C.t, *je = (float *)(b + Lay<MED>::OFF_JE) + C.t;
        for (int c = 0; c < 6 * C.nmx; c++) {
          if (KIND != KIND_LAST) jp[c * NT] = 0.0f;
          if (KIND == KIND_MID || KIND == KIND_LAST) je[c * NT] = 0.0f;
        }
      }
    }
    fence_proxy_async_smem();
  }
  // every thread is done with ring slot s; its PRE / END tiles now hold w_tmp / w_end of this plane. The two thread groups of a
  // split block reach this point through different instantiations of this function: they meet at a NAMED barrier with an explicit
  // thread count (the producer / consumer idiom: warps may arrive from different program counters), not at __syncthreads()
  if (PART == 0) __syncthreads();
  else asm volatile("barrier.sync 1, %0;" ::"r"(2 * TX * TY) : "memory");
  if (threadIdx.x == 0 && threadIdx.y == 0) {
    const int tx0 = C.i0 - P.ni1, ty0 = C.j0 - P.nj1;
    const bool refill = it + NST < nplanes;
    if (refill) tma_issue_a<DX, DY, DZ, KIND, MED, GZ>(P, M, C, k + NST * DIR, s);   // most of the next plane's bytes: requested at once
    if (C.pf > 0 && it + NST + C.pf < nplanes) tma_prefetch_plane<DX, DY, KIND, MED, GZ>(P, M, C, k + (NST + C.pf) * DIR);
    if (KIND != KIND_LAST) tma_store_4d_hint(&M.out_tmp, b + OFF_PRE, tx0, ty0, k, 0, C.pol_stream);
    if (KIND == KIND_MID || KIND == KIND_LAST) tma_store_4d_hint(&M.out_end, b + OFF_END, tx0, ty0, k, 0, C.pol_stream);
    if constexpr (MED == MED_VIS) {
      if (C.nmx > 0) {
        if (KIND != KIND_LAST) tma_store_4d_hint(&M.jout_tmp, b + Lay<MED>::OFF_JP, tx0, ty0, k, 0, C.pol_stream);
        if (KIND == KIND_MID || KIND == KIND_LAST) tma_store_4d_hint(&M.jout_end, b + Lay<MED>::OFF_JE, tx0, ty0, k, 0, C.pol_stream);
      }
    }
    tma_store_commit();
    if (P.l2mode & 8) tma_store_wait_all();   // debugging switch: synchronous stores
    if (refill) {
      // the w_pre / w_end tiles may be refilled -- or, two planes on, overwritten by the threads -- once the stores have read
      // them: thread 0 passes this wait before it joins the next __syncthreads
      tma_store_wait_read();
      tma_issue_b<DX, DY, KIND, MED>(P, M, C, k + NST * DIR, s);
    }
  }
}

// the march of one thread (group) through the planes of its z chunk
template <int DX, int DY, int DZ, int KIND, int MED, bool GZ, int PART, bool TOPK>
__device__ __forceinline__ void march(const StageArgs &P, const TmaMaps &M, const TmaCtx &C, int kf, int nplanes, int ntop, bool pml_xy, int zk1,
                                      int zk2)
{
  constexpr int DIR = DZ ? 1 : -1;
  constexpr int QB = Part<PART>::QB, QN = Part<PART>::QN;
  float q0[QN], q1[QN], q2[QN], q3[QN], q4[QN], qn[QN];
#pragma unroll
  for (int c = 0; c < QN; c++) { q0[c] = q1[c] = q2[c] = q3[c] = q4[c] = qn[c] = 0.0f; }
  if (C.inarr && !(P.l2mode & 256)) {   // bit 8: DIAGNOSTIC (wrong results): no queue priming
    const long sd = (long)DIR * (long)P.siz_slice;
#pragma unroll
    for (int c = 0; c < QN; c++) {
      const float *w = P.cur + (QB + c) * P.siz_vol + (size_t)kf * P.siz_slice + C.pij;
      q0[c] = __ldg(w - 3 * sd); q1[c] = __ldg(w - 2 * sd); q2[c] = __ldg(w - sd);
      if constexpr (!Lay<MED>::ZA) { q3[c] = __ldg(w); qn[c] = __ldg(w + sd); }   // ZA: this plane and the one ahead arrive in shared memory
    }
  }
  // TOPK kernels (the top z chunk of a free-surface problem): the ntop free-surface planes (k >= nk2 - 3) are the chunk's LAST planes
  // when the march runs upwards and its FIRST when it runs downwards; DZ is a template parameter, so a kernel holds one copy of
  // their loop. The kernels of every other chunk (TOPK = false) hold none of this code: the free-surface plane function needs
  // more registers than the plain one, and inside one kernel its spills reached the hot loop (profiles/r2_experiments.txt, r2j).
  int it = 0;
#define CGFD_ROTATE                                                                             \
  if constexpr (!Lay<MED>::ZA) {   /* with ZA tiles the plane function rotates its three-deep queue itself */ \
    _Pragma("unroll") for (int c = 0; c < QN; c++) { q0[c] = q1[c]; q1[c] = q2[c]; q2[c] = q3[c]; q3[c] = q4[c]; } \
  }
  if (TOPK && !DZ) {
#pragma unroll 1
    for (; it < ntop; it++) {
      tma_plane<DX, DY, DZ, KIND, MED, GZ, true, PART, true>(P, M, C, kf + it * DIR, it, nplanes, q0, q1, q2, q3, q4, qn);
      CGFD_ROTATE
    }
  }
  const int nlow = (TOPK && DZ) ? nplanes - ntop : nplanes;
  for (; it < nlow; it++) {
    const int k = kf + it * DIR;
    if (pml_xy || k <= zk1 || k >= zk2 || (P.l2mode & 4)) tma_plane<DX, DY, DZ, KIND, MED, GZ, true, PART>(P, M, C, k, it, nplanes, q0, q1, q2, q3, q4, qn);
    else tma_plane<DX, DY, DZ, KIND, MED, GZ, false, PART>(P, M, C, k, it, nplanes, q0, q1, q2, q3, q4, qn);
    // rotate the zeta queue (register moves; unrolling by 5 instead makes the loop body outgrow the
    // instruction cache and the kernel instruction-fetch bound -- profiles/r1d_summary.txt)
    CGFD_ROTATE
  }
  if (TOPK && DZ) {
#pragma unroll 1
    for (; it < nplanes; it++) {
      tma_plane<DX, DY, DZ, KIND, MED, GZ, true, PART, true>(P, M, C, kf + it * DIR, it, nplanes, q0, q1, q2, q3, q4, qn);
      CGFD_ROTATE
    }
  }
#undef CGFD_ROTATE
}

template <int DX, int DY, int DZ, int KIND, int MED, bool GZ, bool TOPK>
__global__ void __launch_bounds__(TX *TY *(Lay<MED>::SPLIT ? 2 : 1), Lay<MED>::BLOCKS) k_main_tma(const StageArgs P, const __grid_constant__ TmaMaps M)
{
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  constexpr bool SPLIT = Lay<MED>::SPLIT;   // two thread groups per tile: threadIdx.y < TY forms the stress half, the rest the velocity half
  TmaCtx C;
  C.ring = smem_raw;
  C.full = (uint64_t *)(C.ring + NST * Lay<MED>::STAGE_BYTES);
  C.tx = threadIdx.x; C.ty = SPLIT ? threadIdx.y % TY : threadIdx.y; C.t = C.ty * TX + C.tx;
  // 1-D grid; the optional order table puts the tiles that meet an x / y PML slab first: they take ~3x as long per plane,
  // and started last they would be the tail of the launch
  const int bid = P.order ? __ldg(P.order + blockIdx.x) : (int)blockIdx.x;
  const int bxi = bid % P.nbx, byi = (bid / P.nbx) % P.nby, bzi = bid / (P.nbx * P.nby);
  C.i0 = P.ni1 + (P.bx0 + bxi) * TX; C.j0 = P.nj1 + (P.by0 + byi) * TY;
  C.i = C.i0 + C.tx; C.j = C.j0 + C.ty;
  const int k0 = P.kbeg + bzi * P.zchunk;
  C.k1 = min(k0 + P.zchunk - 1, P.kend);
  C.inarr = (C.i < P.nx) && (C.j < P.ny);
  C.active = (C.i <= P.ni2) && (C.j <= P.nj2);
  C.pij = (size_t)C.j * P.siz_line + C.i;
  C.qptr = P.cur + C.pij;
  C.nmx = (MED == MED_VIS && P.vis_staged) ? P.nmaxwell : 0;
  C.fmask = pml_mask_xy(P, C.i, C.j);
  C.pol_keep = (P.l2mode & 1) ? l2_policy_evict_last() : l2_policy_evict_normal();
  C.pol_stream = (P.l2mode & 2) ? l2_policy_evict_first() : l2_policy_evict_normal();
  const bool t0 = threadIdx.x == 0 && threadIdx.y == 0;

  if (t0) {
#pragma unroll
    for (int s = 0; s < NST; s++) mbar_init(C.full + s, 1);
    mbar_fence_init();
  }
  __syncthreads();
  constexpr int DIR = DZ ? 1 : -1;
  const int nplanes = C.k1 - k0 + 1;
  const int kf = DZ ? k0 : C.k1;   // first plane of the march
  C.pf = (P.l2mode >> 5) & 3;
  if (t0) {
#pragma unroll
    for (int s = 0; s < NST; s++)
      if (s < nplanes) tma_issue<DX, DY, DZ, KIND, MED, GZ>(P, M, C, kf + s * DIR, s);
    for (int s = NST; s < NST + C.pf; s++)
      if (s < nplanes) tma_prefetch_plane<DX, DY, KIND, MED, GZ>(P, M, C, kf + s * DIR);
  }
  // does this tile meet the slab of an x or y PML face? (block-uniform; the z faces are tested per plane)
  bool pml_xy = false;
#pragma unroll
  for (int sd = 0; sd < 2; sd++) {
    const PmlFaceDev &Fx = P.pml[0][sd], &Fy = P.pml[1][sd];
    pml_xy |= Fx.on && C.i0 <= Fx.i2 && C.i0 + TX - 1 >= Fx.i1;
    pml_xy |= Fy.on && C.j0 <= Fy.j2 && C.j0 + TY - 1 >= Fy.j1;
  }
  const int zk1 = P.pml[2][0].on ? P.pml[2][0].k2 : -1;            // planes k <= zk1 lie in the bottom slab
  const int zk2 = P.pml[2][1].on ? P.pml[2][1].k1 : (1 << 30);     // planes k >= zk2 lie in the top slab

  // planes of this chunk that are free-surface rows
  int ntop = 0;
  if (TOPK) { const int klo = max(k0, P.nk2 - 3); ntop = max(0, C.k1 - klo + 1); }

  if constexpr (SPLIT) {
    // warp-uniform: rows 0 .. TY-1 of the block are the stress group, rows TY .. 2 TY-1 the velocity group; they meet once per
    // plane at named barrier 1 (see tma_plane)
    if (threadIdx.y < TY) march<DX, DY, DZ, KIND, MED, GZ, 1, TOPK>(P, M, C, kf, nplanes, ntop, pml_xy, zk1, zk2);
    else march<DX, DY, DZ, KIND, MED, GZ, 2, TOPK>(P, M, C, kf, nplanes, ntop, pml_xy, zk1, zk2);
  } else {
    march<DX, DY, DZ, KIND, MED, GZ, 0, TOPK>(P, M, C, kf, nplanes, ntop, pml_xy, zk1, zk2);
  }
  if (t0) tma_store_wait_all();
  if (P.l2mode & 16) __syncthreads();   // debugging switch: no thread leaves before the stores are complete
}

// =============================================================================================
// free-surface rows: k in [nk2-3, nk2], one thread per point and HALF of the RHS, neighbours straight from L1/L2.
// HALF 0 = the six stress components (needs the velocity derivatives only: reduced-order / matrix Dz, Hooke, PML part 0,
// attenuation), HALF 1 = the three velocity components (needs the stress derivatives only: momentum, traction image, PML part 1).
// The two halves share nothing but read-only inputs and write disjoint components, so they run as independent blocks of one
// launch (blockIdx.z = row * 2 + half): twice the parallelism at ~half the registers of a thread that does both (168 registers,
// 12 warps per SM: latency bound, ncu r1x: long-scoreboard 4.5 per issue, 30 % of the DRAM rate, 6 % of a step for 2 % of the points).
// =============================================================================================
template <int DX, int DY, int DZ, int KIND, int MED, int HALF>
__device__ __forceinline__ void top_half(const StageArgs &P, int k)
{
  // 32 x 4 points per block: the eta neighbours of a row are mostly rows of the same block (L1 hits)
  const int i = P.ni1 + blockIdx.x * 32 + threadIdx.x;
  const int j = P.nj1 + blockIdx.y * 4 + threadIdx.y;
  if (i > P.ni2 || j > P.nj2 || k > P.kend) return;
  const size_t L = P.siz_line, S = P.siz_slice, V = P.siz_vol;
  const size_t p = (size_t)k * S + (size_t)j * L + i;
  const size_t p2 = (size_t)j * P.nx + i;
  const float *cx = c_fd.coef[DX], *cy = c_fd.coef[DY], *cz = c_fd.coef[DZ];
  constexpr int FX = Ofs<DX>::first, FY = Ofs<DY>::first, FZ = Ofs<DZ>::first;
  constexpr int C0 = HALF ? 0 : 3, NC = HALF ? 3 : 6;   // components this half updates
  constexpr int D0 = HALF ? 3 : 0, ND = HALF ? 6 : 3;   // components whose derivatives it needs
  const int nsurf = P.nk2 - k;   // 0 at the surface
  const int kmin = P.nk2 - (FZ + 4);   // first row whose momentum RHS is the traction-image one

  float cur[NC], pv[NC], ev[NC];
#pragma unroll
  for (int c = 0; c < NC; c++) {
    cur[c] = __ldg(P.cur + (C0 + c) * V + p);
    if (KIND != KIND_FIRST) pv[c] = __ldg(P.pre + (C0 + c) * V + p);
    if (KIND == KIND_LAST) ev[c] = P.end[(C0 + c) * V + p];
  }
  // In the rows where the traction image replaces the momentum RHS, the plain derivatives of the stress components are only
  // needed by the PML terms: outside the slabs their 90 loads are skipped.
  bool need_plain = true;
  if (HALF == 1 && k >= kmin) {
    need_plain = false;
#pragma unroll
    for (int ax = 0; ax < 3; ax++)
#pragma unroll
      for (int sd = 0; sd < 2; sd++) {
        const PmlFaceDev &F = P.pml[ax][sd];
        if (F.on && i >= F.i1 && i <= F.i2 && j >= F.j1 && j <= F.j2 && k >= F.k1 && k <= F.k2) need_plain = true;
      }
  }
  Deriv d;
#pragma unroll
  for (int c = D0; c < D0 + ND; c++) {
    if (!need_plain) { d.x[c] = d.y[c] = d.z[c] = 0.0f; continue; }
    const float *w = P.cur + c * V + p;
    d.x[c] = cx[0] * __ldg(w + FX) + cx[1] * __ldg(w + FX + 1) + cx[2] * __ldg(w + FX + 2) + cx[3] * __ldg(w + FX + 3)
           + cx[4] * __ldg(w + FX + 4);
    d.y[c] = cy[0] * __ldg(w + (FY + 0) * (long)L) + cy[1] * __ldg(w + (FY + 1) * (long)L)
           + cy[2] * __ldg(w + (FY + 2) * (long)L) + cy[3] * __ldg(w + (FY + 3) * (long)L)
           + cy[4] * __ldg(w + (FY + 4) * (long)L);
    // interior zeta operator; rows whose stencil leaves the grid get replaced below
    d.z[c] = cz[0] * __ldg(w + (FZ + 0) * (long)S) + cz[1] * __ldg(w + (FZ + 1) * (long)S)
           + cz[2] * __ldg(w + (FZ + 2) * (long)S) + cz[3] * __ldg(w + (FZ + 3) * (long)S)
           + cz[4] * __ldg(w + (FZ + 4) * (long)S);
  }
  const Met m = load_metric(P, p);
  Med<MED> md;
  md.load([&](int n) { return __ldg(P.media[n] + p); });
  const float slw = md.slw;
  float h[9];

  if (HALF == 0) {
    // --- velocity gradient along zeta in the top three rows (vlow, iso.c:503-593)
    if (nsurf == 0) {
      // the surface point-force term exists in the isotropic operator only (iso.c:583-592)
      constexpr bool FSRC = (MED == MED_ISO || MED == MED_VIS);
      const float *A = P.matVx2Vz + p2 * 9, *B = P.matVy2Vz + p2 * 9, *F = P.matF2Vz + p2 * 9;
      float sx = (FSRC && P.VxSrc) ? __ldg(P.VxSrc + p2) : 0.0f, sy = (FSRC && P.VySrc) ? __ldg(P.VySrc + p2) : 0.0f,
            sz = (FSRC && P.VzSrc) ? __ldg(P.VzSrc + p2) : 0.0f;
#pragma unroll
      for (int r = 0; r < 3; r++) {
        float v = __ldg(A + 3 * r + 0) * d.x[VX] + __ldg(A + 3 * r + 1) * d.x[VY] + __ldg(A + 3 * r + 2) * d.x[VZ]
                + __ldg(B + 3 * r + 0) * d.y[VX] + __ldg(B + 3 * r + 1) * d.y[VY] + __ldg(B + 3 * r + 2) * d.y[VZ];
        if (FSRC) v += __ldg(F + 3 * r + 0) * sx + __ldg(F + 3 * r + 1) * sy + __ldg(F + 3 * r + 2) * sz;
        d.z[r] = v;
      }
    } else if (nsurf == 1) {
      const long o0 = DZ ? -(long)S : 0, o1 = DZ ? 0 : (long)S;
#pragma unroll
      for (int c = 0; c < 3; c++) {
        const float *w = P.cur + c * V + p;
        d.z[c] = c_fd.lay2[DZ][0] * __ldg(w + o0) + c_fd.lay2[DZ][1] * __ldg(w + o1);
      }
    } else if (nsurf == 2) {
      const long o0 = DZ ? -2 * (long)S : 0;
#pragma unroll
      for (int c = 0; c < 3; c++) {
        const float *w = P.cur + c * V + p + o0;
        d.z[c] = c_fd.lay3[DZ][0] * __ldg(w) + c_fd.lay3[DZ][1] * __ldg(w + S) + c_fd.lay3[DZ][2] * __ldg(w + 2 * S);
      }
    }
    hooke<MED>(d, m, md, h);
    pml_all<KIND, 0, MED>(P, i, j, k, d, m, md, h);
    if constexpr (MED == MED_VIS) atten_update<KIND>(P, p, md.lam, md.mu, h);
  } else {
    if (k < kmin) momentum(d, m, slw, h);
    else {
      // --- traction image: momentum RHS in conservative form (sv_curv_col_el.c:84-304)
      const int n_free = P.nk2 - k - FZ;
      const float jac = __ldg(P.metric[M_JAC] + p);
      const float slwjac = slw / jac;
      // component triplets (T1,T2,T3) with flux_n = J*(e_x T1 + e_y T2 + e_z T3)
      const int T1[3] = {TXX, TXY, TXZ}, T2[3] = {TXY, TYY, TYZ}, T3[3] = {TXZ, TYZ, TZZ};
      const float *Ts[3] = {P.TxSrc, P.TySrc, P.TzSrc};
      // the xi / eta flux derivatives are accumulated term by term (left to right, like M_FD_NOINDX, forward/fd_t.h:33-38);
      // only the zeta fluxes are kept, because the image terms refer back to them
      float Dxf[3], Dyf[3], fz[3][5];
#pragma unroll
      for (int n = 0; n < 5; n++) {
        {
          const size_t pp = p + (FX + n);
          const float jn = __ldg(P.metric[M_JAC] + pp), ex = __ldg(P.metric[M_XIX] + pp), ey = __ldg(P.metric[M_XIY] + pp),
                      ez = __ldg(P.metric[M_XIZ] + pp);
#pragma unroll
          for (int v = 0; v < 3; v++) {
            const float f = jn * (ex * __ldg(P.cur + T1[v] * V + pp) + ey * __ldg(P.cur + T2[v] * V + pp) + ez * __ldg(P.cur + T3[v] * V + pp));
            if (n == 0) Dxf[v] = cx[0] * f; else Dxf[v] += cx[n] * f;
          }
        }
        {
          const size_t pp = p + (long)(FY + n) * (long)L;
          const float jn = __ldg(P.metric[M_JAC] + pp), ex = __ldg(P.metric[M_ETX] + pp), ey = __ldg(P.metric[M_ETY] + pp),
                      ez = __ldg(P.metric[M_ETZ] + pp);
#pragma unroll
          for (int v = 0; v < 3; v++) {
            const float f = jn * (ex * __ldg(P.cur + T1[v] * V + pp) + ey * __ldg(P.cur + T2[v] * V + pp) + ez * __ldg(P.cur + T3[v] * V + pp));
            if (n == 0) Dyf[v] = cy[0] * f; else Dyf[v] += cy[n] * f;
          }
        }
        if (n < n_free) {
          const size_t pp = p + (long)(FZ + n) * (long)S;
          const float jn = __ldg(P.metric[M_JAC] + pp), ex = __ldg(P.metric[M_ZTX] + pp), ey = __ldg(P.metric[M_ZTY] + pp),
                      ez = __ldg(P.metric[M_ZTZ] + pp);
#pragma unroll
          for (int v = 0; v < 3; v++)
            fz[v][n] = jn * (ex * __ldg(P.cur + T1[v] * V + pp) + ey * __ldg(P.cur + T2[v] * V + pp) + ez * __ldg(P.cur + T3[v] * V + pp));
        }
      }
#pragma unroll
      for (int v = 0; v < 3; v++) {
        const float ts = Ts[v] ? __ldg(Ts[v] + p2) : 0.0f;
#pragma unroll
        for (int n = 0; n < 5; n++) {
          if (n == n_free) fz[v][n] = ts;
          else if (n > n_free) {
            const int im = 2 * n_free - n;    // mirror index inside the window
            float below;
            if (im >= 0) below = fz[v][im < 0 ? 0 : im];
            else if (P.timg_mode == 0) below = 0.0f;
            else {
              // image row k + indx[n] - 2(n - n_free) (sv_curv_col_el.c:154-159)
              const size_t pp = p + (long)(FZ + n - 2 * (n - n_free)) * (long)S;
              const float jn = __ldg(P.metric[M_JAC] + pp), ex = __ldg(P.metric[M_ZTX] + pp), ey = __ldg(P.metric[M_ZTY] + pp),
                          ez = __ldg(P.metric[M_ZTZ] + pp);
              below = jn * (ex * __ldg(P.cur + T1[v] * V + pp) + ey * __ldg(P.cur + T2[v] * V + pp) + ez * __ldg(P.cur + T3[v] * V + pp));
            }
            fz[v][n] = 2.0f * ts - below;
          }
        }
        float Dz = cz[0] * fz[v][0]; Dz += cz[1] * fz[v][1]; Dz += cz[2] * fz[v][2]; Dz += cz[3] * fz[v][3]; Dz += cz[4] * fz[v][4];
        h[v] = (Dxf[v] + Dyf[v] + Dz) * slwjac;
      }
    }
    pml_all<KIND, 1, MED>(P, i, j, k, d, m, md, h);
  }
  const float qatt = (KIND == KIND_LAST && P.qatt) ? __ldg(P.qatt + p) : 1.0f;
#pragma unroll
  for (int c = 0; c < NC; c++) rk_wave<KIND>(P.tmp, P.end, (C0 + c) * V + p, cur[c], pv[c], ev[c], h[C0 + c], P.a, P.b, P.c, qatt);
}
#ifndef CGFD_TOP_BLOCKS
#define CGFD_TOP_BLOCKS 4   // resident blocks of 128 threads per SM the free-surface kernel is compiled for (<= 128 registers)
#endif
template <int DX, int DY, int DZ, int KIND, int MED>
__global__ void __launch_bounds__(128, CGFD_TOP_BLOCKS) k_top(const StageArgs P)
{
  const int k = P.kbeg + (blockIdx.z >> 1);
  if (blockIdx.z & 1) top_half<DX, DY, DZ, KIND, MED, 0>(P, k);
  else top_half<DX, DY, DZ, KIND, MED, 1>(P, k);   // the velocity half of a row (traction image: ~2x the loads) starts first
}

// =============================================================================================
// rows [P.kbeg, P.kend] of the tile rectangle [bx0,bx1) x [by0,by1) (tiles of TX x TY points counted from (ni1,nj1)) in z chunks of
// `zchunk` rows. TOPK: the launch that holds the free-surface rows (its kernels carry the free-surface plane function).
template <int DX, int DY, int DZ, int KIND, int MED, bool GZ, bool TOPK>
static void launch_main_t(const StageArgs &P0, const TmaMaps *maps, int zchunk, const int rect[4], cudaStream_t st,
                          cudaEvent_t ev0, cudaEvent_t ev1, int *nlaunch)
{
  StageArgs P = P0;
  const int bx = rect[1] - rect[0], by = rect[3] - rect[2];
  if (P.kend < P.kbeg || bx <= 0 || by <= 0 || zchunk <= 0) return;
  P.bx0 = rect[0]; P.by0 = rect[2];
  const int nk = P.kend - P.kbeg + 1;
  P.zchunk = zchunk;
  const int nzc = (nk + P.zchunk - 1) / P.zchunk;
  P.nbx = bx; P.nby = by;
  dim3 grid(bx * by * nzc), block(TX, Lay<MED>::SPLIT ? 2 * TY : TY);
  if (ev0) cudaEventRecord(ev0, st);
  k_main_tma<DX, DY, DZ, KIND, MED, GZ, TOPK><<<grid, block, Lay<MED>::SMEM_BYTES, st>>>(P, *maps);
  if (ev1) cudaEventRecord(ev1, st);
  (*nlaunch)++;
}

// the free-surface rows (whole x-y range): blockIdx.z = row * 2 + half
template <int DX, int DY, int DZ, int KIND, int MED>
static void launch_top_t(const StageArgs &P0, cudaStream_t st, int *nlaunch)
{
  if (!P0.free_top || P0.fuse_top) return;   // fused: the interior kernel's top chunk handles these rows
  StageArgs P = P0;
  const int ni = P.ni2 - P.ni1 + 1, nj = P.nj2 - P.nj1 + 1;
  const int ktop = P.nk2 - 3;
  P.kbeg = (ktop < P.nk1) ? P.nk1 : ktop; P.kend = P.nk2;
  dim3 block(32, 4), grid((ni + 31) / 32, (nj + 3) / 4, (P.kend - P.kbeg + 1) * 2);
  k_top<DX, DY, DZ, KIND, MED><<<grid, block, 0, st>>>(P);
  (*nlaunch)++;
}

template <int DX, int DY, int DZ, int KIND, int MED, bool GZ, bool TOPK> static int set_attr_g()
{
  cudaError_t e = cudaFuncSetAttribute(k_main_tma<DX, DY, DZ, KIND, MED, GZ, TOPK>, cudaFuncAttributeMaxDynamicSharedMemorySize, Lay<MED>::SMEM_BYTES);
  if (e != cudaSuccess) return 1;
  e = cudaFuncSetAttribute(k_main_tma<DX, DY, DZ, KIND, MED, GZ, TOPK>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  return e != cudaSuccess;
}
template <int DX, int DY, int DZ, int KIND, int MED> static int set_attr_t()
{
  return set_attr_g<DX, DY, DZ, KIND, MED, false, false>() | set_attr_g<DX, DY, DZ, KIND, MED, true, false>() |
         set_attr_g<DX, DY, DZ, KIND, MED, false, true>() | set_attr_g<DX, DY, DZ, KIND, MED, true, true>();
}
template <int KIND, int MED> int set_attr_k()
{
  return set_attr_t<0, 0, 0, KIND, MED>() | set_attr_t<0, 0, 1, KIND, MED>() | set_attr_t<0, 1, 0, KIND, MED>() | set_attr_t<0, 1, 1, KIND, MED>() |
         set_attr_t<1, 0, 0, KIND, MED>() | set_attr_t<1, 0, 1, KIND, MED>() | set_attr_t<1, 1, 0, KIND, MED>() | set_attr_t<1, 1, 1, KIND, MED>();
}
template <int MED> int med_blocks_per_sm() { return Lay<MED>::BLOCKS; }
template <int MED> int med_kernels_init()
{
  return set_attr_k<KIND_FIRST, MED>() | set_attr_k<KIND_MID, MED>() | set_attr_k<KIND_THIRD, MED>() | set_attr_k<KIND_LAST, MED>();
}

#define CGFD_DISPATCH_DIR(CALL)                                                                   \
  switch (dx * 4 + dy * 2 + dz) {                                                                  \
    case 0: CALL(0, 0, 0); break; case 1: CALL(0, 0, 1); break; case 2: CALL(0, 1, 0); break;      \
    case 3: CALL(0, 1, 1); break; case 4: CALL(1, 0, 0); break; case 5: CALL(1, 0, 1); break;      \
    case 6: CALL(1, 1, 0); break; default: CALL(1, 1, 1); break;                                   \
  }

template <int KIND, int MED>
void launch_main_k(const StageArgs &P, const TmaMaps *maps, int dx, int dy, int dz, int gz, int topk, int zchunk,
                   const int rect[4], cudaStream_t st, cudaEvent_t e0, cudaEvent_t e1, int *n)
{
#define CALL(a, b, c)                                                                                \
  do {                                                                                               \
    if (topk) {                                                                                      \
      if (gz) launch_main_t<a, b, c, KIND, MED, true, true>(P, maps, zchunk, rect, st, e0, e1, n);   \
      else launch_main_t<a, b, c, KIND, MED, false, true>(P, maps, zchunk, rect, st, e0, e1, n);     \
    } else {                                                                                         \
      if (gz) launch_main_t<a, b, c, KIND, MED, true, false>(P, maps, zchunk, rect, st, e0, e1, n);  \
      else launch_main_t<a, b, c, KIND, MED, false, false>(P, maps, zchunk, rect, st, e0, e1, n);    \
    }                                                                                                \
  } while (0)
  CGFD_DISPATCH_DIR(CALL)
#undef CALL
}
template <int KIND, int MED> void launch_top_k(const StageArgs &P, int dx, int dy, int dz, cudaStream_t st, int *n)
{
#define CALL(a, b, c) launch_top_t<a, b, c, KIND, MED>(P, st, n)
  CGFD_DISPATCH_DIR(CALL)
#undef CALL
}

template <int MED>
void med_launch_main(const StageArgs &P, const TmaMaps *maps, int dx, int dy, int dz, int kind, int gz, int topk, int zchunk, const int rect[4],
                     cudaStream_t st, cudaEvent_t ev0, cudaEvent_t ev1, int *nlaunch)
{
  if (kind == KIND_FIRST) launch_main_k<KIND_FIRST, MED>(P, maps, dx, dy, dz, gz, topk, zchunk, rect, st, ev0, ev1, nlaunch);
  else if (kind == KIND_MID) launch_main_k<KIND_MID, MED>(P, maps, dx, dy, dz, gz, topk, zchunk, rect, st, ev0, ev1, nlaunch);
  else if (kind == KIND_THIRD) launch_main_k<KIND_THIRD, MED>(P, maps, dx, dy, dz, gz, topk, zchunk, rect, st, ev0, ev1, nlaunch);
  else launch_main_k<KIND_LAST, MED>(P, maps, dx, dy, dz, gz, topk, zchunk, rect, st, ev0, ev1, nlaunch);
}

template <int MED> void med_launch_top(const StageArgs &P, int dx, int dy, int dz, int kind, cudaStream_t st, int *nlaunch)
{
  if (kind == KIND_FIRST) launch_top_k<KIND_FIRST, MED>(P, dx, dy, dz, st, nlaunch);
  else if (kind == KIND_MID) launch_top_k<KIND_MID, MED>(P, dx, dy, dz, st, nlaunch);
  else if (kind == KIND_THIRD) launch_top_k<KIND_THIRD, MED>(P, dx, dy, dz, st, nlaunch);
  else launch_top_k<KIND_LAST, MED>(P, dx, dy, dz, st, nlaunch);
}

// Translation units: the 64 instantiations of one medium and stage kind (8 operator pairs x general / GZ grid x plane-function
// copies) are one unit (kernels_kind.cu compiled with -DCGFD_MED=.. -DCGFD_KIND=..); the per-medium units kernels_{iso,..}.cu hold
// the dispatch over the kinds only. 16 units of similar size build in parallel.
#define CGFD_INSTANTIATE_KIND(KIND, MED)                                                                               \
  template int set_attr_k<KIND, MED>();                                                                                 \
  template void launch_main_k<KIND, MED>(const StageArgs &, const TmaMaps *, int, int, int, int, int, int, const int[4], cudaStream_t, \
                                         cudaEvent_t, cudaEvent_t, int *);                                              \
  template void launch_top_k<KIND, MED>(const StageArgs &, int, int, int, cudaStream_t, int *);
#define CGFD_EXTERN_KIND(KIND, MED)                                                                                    \
  extern template int set_attr_k<KIND, MED>();                                                                          \
  extern template void launch_main_k<KIND, MED>(const StageArgs &, const TmaMaps *, int, int, int, int, int, int, const int[4], cudaStream_t, \
                                                cudaEvent_t, cudaEvent_t, int *);                                       \
  extern template void launch_top_k<KIND, MED>(const StageArgs &, int, int, int, cudaStream_t, int *);
#define CGFD_INSTANTIATE_MEDIUM(MED)                                                                                  \
  CGFD_EXTERN_KIND(KIND_FIRST, MED) CGFD_EXTERN_KIND(KIND_MID, MED) CGFD_EXTERN_KIND(KIND_THIRD, MED) CGFD_EXTERN_KIND(KIND_LAST, MED) \
  template int med_kernels_init<MED>();                                                                                \
  template int med_blocks_per_sm<MED>();                                                                               \
  template void med_launch_main<MED>(const StageArgs &, const TmaMaps *, int, int, int, int, int, int, int, const int[4], cudaStream_t, \
                                     cudaEvent_t, cudaEvent_t, int *);                                                 \
  template void med_launch_top<MED>(const StageArgs &, int, int, int, int, cudaStream_t, int *);

}  // namespace cgfd
