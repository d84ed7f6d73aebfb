// Host side of libcgfd3d_b200.so: the C ABI of include/cgfd3d_b200.h.
// Owns the device mirrors of the reference's wav_t / bdry_t / md_t / gdcurv_metric_t / src_t data
// and runs the RK4 stage loop of forward/drv_rk_curv_col.c:167-544 on the GPU.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <array>
#include <map>
#include <string>
#include <vector>

#include "../../include/cgfd3d_b200.h"
#include "aux_kernels.cuh"
#include "cgfd_dev.cuh"
#include "halo.h"
#include "physics.cuh"

using namespace cgfd;

static thread_local std::string g_err;

// medium dispatch (ABI medium type -> kernel family)
static int med_of(int medium_type)
{
  switch (medium_type) {
    case CGFD_MEDIUM_ELASTIC_ISO: return MED_ISO;
    case CGFD_MEDIUM_ELASTIC_VTI: return MED_VTI;
    case CGFD_MEDIUM_ELASTIC_ANISO: return MED_ANISO;
    case CGFD_MEDIUM_VISCOELASTIC_ISO: return MED_VIS;
    default: return -1;
  }
}
#define MED_SWITCH(med, CALL)                                                              \
  switch (med) {                                                                           \
    case MED_ISO: CALL(MED_ISO); break; case MED_VTI: CALL(MED_VTI); break;                \
    case MED_ANISO: CALL(MED_ANISO); break; default: CALL(MED_VIS); break;                 \
  }
static int kernels_init(int med)
{
  int rc = 1;
#define CALL(M) rc = med_kernels_init<M>()
  MED_SWITCH(med, CALL)
#undef CALL
  return rc;
}
static void launch_main(int med, const StageArgs &P, const TmaMaps *maps, const int *dir, int kind, int gz, int topk, int zchunk, const int rect[4],
                        cudaStream_t st, cudaEvent_t e0, cudaEvent_t e1, int *nl)
{
#define CALL(M) med_launch_main<M>(P, maps, dir[0], dir[1], dir[2], kind, gz, topk, zchunk, rect, st, e0, e1, nl)
  MED_SWITCH(med, CALL)
#undef CALL
}
static int blocks_per_sm(int med)
{
  int n = 1;
#define CALL(M) n = med_blocks_per_sm<M>()
  MED_SWITCH(med, CALL)
#undef CALL
  return n;
}
static void launch_top(int med, const StageArgs &P, const int *dir, int kind, cudaStream_t st, int *nl)
{
#define CALL(M) med_launch_top<M>(P, dir[0], dir[1], dir[2], kind, st, nl)
  MED_SWITCH(med, CALL)
#undef CALL
}

static int fail(const std::string &m) { g_err = m; return 1; }
#define CK(call)                                                                                   \
  do {                                                                                             \
    cudaError_t e_ = (call);                                                                       \
    if (e_ != cudaSuccess)                                                                         \
      return fail(std::string(#call) + ": " + cudaGetErrorString(e_) + " (" __FILE__ ":" + std::to_string(__LINE__) + ")"); \
  } while (0)

struct PmlFaceHost {
  int on = 0, nlay = 0;
  int r[6] = {0, 0, 0, 0, 0, 0};   // i1,i2,j1,j2,k1,k2
  size_t siz = 0;
  float *A = nullptr, *B = nullptr, *D = nullptr;
  float *aux[4] = {nullptr, nullptr, nullptr, nullptr};
  float *zero = nullptr;
};

// one snapshot / slice output (iosnap_t / ioslice_t, forward/io_funcs.c:640-760): a strided sub-box of a few components,
// packed on the device every `tinv` steps from step `it1` on and copied to the caller's host buffer on the I/O stream
constexpr int SNAP_RING = 4;
struct SnapTap {
  int ncmp = 0; int cmps[32];
  int box[9];                    // i1, ni, di, j1, nj, dj, k1, nk, dk (local indices incl. ghosts)
  int it1 = 0, tinv = 1, max_frames = 0, nframes = 0;
  long nissued = 0;              // frames ever issued (ring slot = nissued % SNAP_RING; nframes restarts with every output buffer)
  size_t cmp_elems = 0;          // ni*nj*nk
  float *host = nullptr;         // [max_frames][ncmp][nk][nj][ni]
  float *ring[SNAP_RING];
  cudaEvent_t packed[SNAP_RING], copied[SNAP_RING];
};

// how one tile rectangle of the interior kernel is launched: rows per z chunk and the block order (device table)
struct LaunchPlan {
  int zchunk = 0;
  int *order = nullptr;
  // fused free surface: rows [ktop0, nk2] are ONE chunk per tile, launched with the kernels that carry the free-surface plane
  // function (TOPK); the launch above covers [nk1, ktop0 - 1]
  int ktop0 = 0, zchunk_top = 0;
  int *order_top = nullptr;
};

struct cgfd_b200_ctx {
  int device = 0;
  cudaStream_t st = nullptr;        // compute stream
  cudaStream_t st2 = nullptr;       // boundary phase: free-surface rows, tiles next to inter-rank faces, halo exchange
  cudaStream_t st3 = nullptr;       // fused free surface: the top-chunk launch of the interior tiles, beside the launch of the chunks below
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr, ev_fork3 = nullptr, ev_join3 = nullptr;
  int nsm = 148;                    // multiprocessors of the device (launch plan: resident blocks per wave)
  int l2mode = 3;                   // L2 eviction hints of the interior kernel (CGFD_L2MODE)
  int overlap = 1;                  // run the boundary phase concurrently with the interior kernel
  int top_rows = 24;                // fused free surface: rows of the TOPK launch (CGFD_TOP_ROWS)
  int top_stream = 1;               // fused free surface: top-chunk launch on its own stream (CGFD_TOP_STREAM=0: same stream, before the rest)
  int fuse_top = -1;                // free-surface rows as planes of the top z chunk (TOPK launch of k_main_tma) instead of k_top:
                                    // the default where it is the faster route (isotropic medium, measured r2l), CGFD_FUSE_TOP=0 / 1
  int toppar = 0;                   // single rank: free-surface kernel on the second stream beside the interior kernel (CGFD_TOPPAR)
  int ntx = 0, nty = 0;             // tiles of the interior kernel along x / y
  int src_nb = 0;                   // source footprint points that belong to the boundary phase (first in the list)
  cgfd_grid_t g;
  cgfd_fd_t fd;
  float dt = 0;
  int medium = 0, med = 0, nmaxwell = 0, ncmp = 9, nmedia = 0;
  float wl[CGFD_MAX_MAXWELL];
  // padded device layout (see cgfd_dev.cuh): pitch PX, index 0 of a row sits `shift` floats in
  int PX = 0, shift = 0;
  size_t V = 0, slice = 0;          // padded volume / slice (floats)
  size_t hV = 0, hslice = 0;        // host (unpadded) volume / slice
  float *lev[4] = {nullptr, nullptr, nullptr, nullptr};   // bases (unshifted)
  float *metric_blk = nullptr, *media_blk = nullptr;
  float *qatt = nullptr;            // Graves' attenuation factor per point (padded layout, unshifted base) or nullptr
  CUtensorMap map_halo[4], map_cen[4], map_out[4], map_met, map_met5, map_med;
  CUtensorMap map_jcen[4], map_jout[4], map_y;   // visco-elastic medium: memory variables of each level, Ylam / Ymu
  int vis_staged = 0;               // memory variables staged by TMA in the interior kernel (nmaxwell <= VIS_MAX_STAGED)
  int gz = 0;                       // xi_y = xi_z = eta_x = eta_z == 0 at every physical point: GZ kernels (cgfd_dev.cuh)
  bool have_maps = false;
  int zchunk = 0;                   // explicit rows per z chunk (CGFD_ZCHUNK), 0 = chosen by plan_for()
  int plan_waves = 16, plan_minchunk = 16, plan_lpt = 2;   // measured at 400x400x200: chunks of 14..28 rows within 0.5 %, 49 rows 2 % slower
  std::map<std::array<int, 5>, struct LaunchPlan> plans;
  int ipre = 0, ia = 1, ib = 2, iend = 3;   // roles of the four level buffers
  float *metric[NMETRIC];           // shifted pointers into metric_blk
  float *media[MAX_MEDIA];
  int free_top = 0, timg_mode = 0;
  PmlFaceHost pml[3][2];
  float *mats[4] = {nullptr, nullptr, nullptr, nullptr};
  float *srcslice = nullptr;     // 6 x [ny][nx]
  SrcDev src;
  bool has_src = false, has_surf = false;
  std::vector<void *> owned;     // misc device allocations
  // sponge
  int ablexp = 0; int ablexp_blk[6][7]; float *Ex = nullptr, *Ey = nullptr, *Ez = nullptr;
  // taps
  int nrec = 0, rec_max_nt = 0, rec_count = 0; int64_t *rec_iptr = nullptr; float *rec = nullptr;
  float *PG = nullptr, *Dis = nullptr;
  float *boxbuf = nullptr; size_t boxcap = 0;
  // host <-> device traffic that must not stall the stage pipeline runs on its own stream
  cudaStream_t st_io = nullptr;
  float *stage[2] = {nullptr, nullptr};            // one unpadded component each (set / get_wavefield)
  cudaEvent_t stage_full[2] = {nullptr, nullptr}, stage_free[2] = {nullptr, nullptr};
  std::vector<struct SnapTap *> snaps;
  // distributed (finite-fault) sources: points once, time-function blocks double-buffered
  struct {
    int n = 0, vi_on = 0, mij_on = 0, max_stage = 0, nt_block = 0;
    int64_t *iptr = nullptr; float *wV = nullptr, *rjac = nullptr;
    int nb = 0; int *sel = nullptr;                // point numbers, the nb boundary-phase ones first (see run_stage)
    float *vi[2] = {nullptr, nullptr}, *mij[2] = {nullptr, nullptr};
    int it_first[2] = {-1, -1}, nt[2] = {0, 0};
    int next = 0;                                  // buffer the next block goes to
    cudaEvent_t loaded[2] = {nullptr, nullptr}, used[2] = {nullptr, nullptr};
  } dd;
  // measurement
  int profiling = 0;
  std::vector<cudaEvent_t> ev;   // pairs around the main kernel
  size_t ev_used = 0;
  FILE *prof_dump = nullptr;
  double main_ms = 0; int64_t main_launches = 0, total_launches = 0;
  cudaEvent_t run0 = nullptr, run1 = nullptr, ev_rec = nullptr; double last_run_ms = 0;
  struct { int it_last = -1; cudaEvent_t done = nullptr; } blk[4];   // completion of the last asynchronous blocks (run_async / wait_block)
  int blk_next = 0;
  int variant = 0;
  int neigh[4] = {-1, -1, -1, -1};
  HaloComm *halo = nullptr;
};

extern "C" const char *cgfd_b200_last_error(void) { return g_err.c_str(); }
extern "C" int cgfd_b200_abi_version(void) { return CGFD_ABI_VERSION; }
extern "C" size_t cgfd_b200_sizeof_problem(void) { return sizeof(cgfd_problem_t); }
extern "C" int cgfd_b200_device_count(void)
{
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

template <typename T> static int upload(cgfd_b200_ctx *c, T **dst, const T *src, size_t n)
{
  *dst = nullptr;
  if (n == 0) return 0;
  CK(cudaMalloc((void **)dst, n * sizeof(T)));
  c->owned.push_back(*dst);
  // everything goes through the context's own stream: it is a non-blocking stream, so work issued on the legacy default
  // stream (a plain cudaMemset is asynchronous for device memory) would NOT be ordered before the copies and kernels
  // that follow on c->st -- with device-to-device set-up copies that race was real (zeroed metrics, NaN at the surface)
  if (src) CK(cudaMemcpyAsync(*dst, src, n * sizeof(T), cudaMemcpyDefault, c->st));
  else CK(cudaMemsetAsync(*dst, 0, n * sizeof(T), c->st));
  CK(cudaStreamSynchronize(c->st));
  return 0;
}

// host [rows][nx] <-> device padded rows (pitch PX, shifted)
static int copy_in3d(cgfd_b200_ctx *c, float *dev_base, const float *host, size_t ncomp, cudaStream_t st)
{
  CK(cudaMemcpy2DAsync(dev_base + c->shift, (size_t)c->PX * sizeof(float), host, (size_t)c->g.nx * sizeof(float),
                       (size_t)c->g.nx * sizeof(float), (size_t)c->g.ny * c->g.nz * ncomp, cudaMemcpyDefault, st));
  return 0;
}
static int copy_out3d(cgfd_b200_ctx *c, float *host, const float *dev_base, size_t ncomp, cudaStream_t st)
{
  CK(cudaMemcpy2DAsync(host, (size_t)c->g.nx * sizeof(float), dev_base + c->shift, (size_t)c->PX * sizeof(float),
                       (size_t)c->g.nx * sizeof(float), (size_t)c->g.ny * c->g.nz * ncomp, cudaMemcpyDefault, st));
  return 0;
}
// host flat index (i + j*nx + k*nx*ny) -> index relative to a shifted device pointer
static inline int64_t dev_index(const cgfd_b200_ctx *c, int64_t hp)
{
  const int64_t nx = c->g.nx, ny = c->g.ny;
  const int64_t i = hp % nx, j = (hp / nx) % ny, k = hp / (nx * ny);
  return i + j * (int64_t)c->PX + k * (int64_t)c->slice;
}

typedef CUresult (*encode_tiled_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                    const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static encode_tiled_fn get_encode()
{
  static encode_tiled_fn fn = nullptr;
  if (!fn) {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = (encode_tiled_fn)p;
  }
  return fn;
}
// 4-D map over [ncomp][nz][ny][PX] float32 with box (bx, by, 1, bc)
// phys = true: origin at the first physical point of a row / column, extents ni x nj (store maps: nothing outside the
// physical x-y range is ever written)
static int make_map(cgfd_b200_ctx *c, CUtensorMap *m, float *base, int ncomp, int bx, int by, int bc, bool phys = false,
                    const char *promo_env = "CGFD_L2PROMO")
{
  encode_tiled_fn enc = get_encode();
  if (!enc) return fail("cuTensorMapEncodeTiled is not available from this driver");
  // x extent = the logical row (shift + nx), not the padded pitch: the part of the last tile's box that hangs over the row is
  // zero-filled by the TMA unit instead of being fetched from the pad columns (4 % of every operand at 400 points per row)
  cuuint64_t dim[4] = {(cuuint64_t)(c->shift + c->g.nx), (cuuint64_t)c->g.ny, (cuuint64_t)c->g.nz, (cuuint64_t)ncomp};
  if (phys) {
    dim[0] = (cuuint64_t)(c->g.ni2 - c->g.ni1 + 1); dim[1] = (cuuint64_t)(c->g.nj2 - c->g.nj1 + 1);
    base += c->shift + c->g.ni1 + (size_t)c->g.nj1 * c->PX;
  }
  cuuint64_t str[3] = {(cuuint64_t)c->PX * 4, (cuuint64_t)c->slice * 4, (cuuint64_t)c->V * 4};
  cuuint32_t box[4] = {(cuuint32_t)bx, (cuuint32_t)by, 1, (cuuint32_t)bc};
  cuuint32_t es[4] = {1, 1, 1, 1};
  // L2 promotion = granularity of the DRAM fetch behind a box row (CGFD_L2PROMO: 0 none, 1 64 B, 2 128 B, 3 256 B)
  static const CUtensorMapL2promotion promo_tab[4] = {CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_64B,
                                                      CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B};
  int promo = 2;
  if (const char *e = getenv("CGFD_L2PROMO")) promo = atoi(e) & 3;
  if (const char *e = getenv(promo_env)) promo = atoi(e) & 3;
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, base, dim, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_NONE, promo_tab[promo], CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail("cuTensorMapEncodeTiled failed with code " + std::to_string((int)r));
  return 0;
}

// metric / media arrays may be handed over as host or as device pointers (unified addressing tells which)
static bool is_device_ptr(const void *p)
{
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
  return a.type == cudaMemoryTypeDevice;
}
struct Peek {   // read single elements of such an array on the host
  const float *p; bool dev;
  std::map<size_t, float> cache;   // device arrays: elements fetched ahead by preload() with one gather
  explicit Peek(const float *q) : p(q), dev(is_device_ptr(q)) {}
  // device arrays: fetch all of idx with one gather kernel + one copy instead of one blocking cudaMemcpy per element
  // (a Gaussian source footprint reads ~700 elements)
  void preload(const std::vector<size_t> &idx)
  {
    if (!dev || idx.empty()) return;
    std::vector<int64_t> h(idx.begin(), idx.end());
    int64_t *di = nullptr; float *dv = nullptr;
    std::vector<float> v(idx.size());
    if (cudaMalloc((void **)&di, h.size() * sizeof(int64_t)) == cudaSuccess && cudaMalloc((void **)&dv, h.size() * sizeof(float)) == cudaSuccess &&
        cudaMemcpy(di, h.data(), h.size() * sizeof(int64_t), cudaMemcpyHostToDevice) == cudaSuccess) {
      k_record<<<(unsigned)((h.size() + 127) / 128), 128>>>(p, 0, 1, (int)h.size(), di, dv);   // one component, npts = n: out[ip] = p[idx[ip]]
      if (cudaMemcpy(v.data(), dv, v.size() * sizeof(float), cudaMemcpyDeviceToHost) == cudaSuccess)
        for (size_t n = 0; n < idx.size(); n++) cache[idx[n]] = v[n];
    }
    cudaGetLastError();
    if (di) cudaFree(di);
    if (dv) cudaFree(dv);
  }
  float operator[](size_t i) const
  {
    if (!dev) return p[i];
    auto it = cache.find(i);
    if (it != cache.end()) return it->second;
    float v = 0.0f;
    cudaMemcpy(&v, p + i, sizeof(float), cudaMemcpyDeviceToHost);
    return v;
  }
};

static float fun_gauss(float t, float a, float t0)
{
  // forward/src_t.c:2178-2184
  float f = exp(-(t - t0) * (t - t0) / (a * a)) / (sqrtf(M_PI) * a);
  return f;
}

// normalised 3-D Gaussian footprint truncated at the free surface (src_cal_norm_delt3d_z2fre,
// forward/src_t.c:2110-2151)
static void norm_delt3d_z2fre(std::vector<float> &delt, float x0, float y0, float z0, float r, int L, int Lz2)
{
  int n1 = 2 * L + 1;
  delt.assign((size_t)n1 * n1 * n1, 0.0f);
  size_t ip = 0;
  for (int k = -L; k <= L; k++)
    for (int j = -L; j <= L; j++)
      for (int i = -L; i <= L; i++) {
        float D1 = fun_gauss(i - x0, r, 0.0f), D2 = fun_gauss(j - y0, r, 0.0f), D3 = fun_gauss(k - z0, r, 0.0f);
        delt[ip++] = D1 * D2 * D3;
      }
  float sum = 0.0f;
  ip = 0;
  for (int k = -L; k <= Lz2; k++)
    for (int j = -L; j <= L; j++)
      for (int i = -L; i <= L; i++) sum += delt[ip++];
  for (auto &v : delt) v /= sum;
}

static int setup_sources(cgfd_b200_ctx *c, const cgfd_problem_t *p)
{
  const cgfd_src_t &s = p->src;
  memset(&c->src, 0, sizeof(c->src));
  if (s.total_number <= 0) return 0;
  const cgfd_grid_t &g = c->g;
  const size_t L = g.nx, S = (size_t)g.nx * g.ny;
  Peek jac(p->metric[CGFD_JAC]);
  Peek slw(p->media[(p->medium_type == CGFD_MEDIUM_ELASTIC_VTI) ? 5 : (p->medium_type == CGFD_MEDIUM_ELASTIC_ANISO) ? 21 : 2]);
  if (jac.dev || slw.dev) {
    // every element the loops below read: the source points and, for Gaussian sources, their footprints up to the surface row
    std::vector<size_t> need;
    const int Hh = (s.itype_spatial_ext == CGFD_SRC_SPATIAL_POINT) ? 0 : s.ext_half_npoint;
    for (int is = 0; is < s.total_number; is++)
      for (int ke = -Hh; ke <= Hh; ke++) for (int je = -Hh; je <= Hh; je++) for (int ix = -Hh; ix <= Hh; ix++) {
        const int i = s.si[is] + ix, j = s.sj[is] + je, k = s.sk[is] + ke;
        if (i < 0 || i >= g.nx || j < 0 || j >= g.ny || k < 0 || k >= g.nz) continue;
        need.push_back(i + j * L + k * S);
      }
    jac.preload(need); slw.preload(need);
  }
  std::vector<int64_t> pt_iptr; std::vector<int> pt_src, pt_bnd; std::vector<float> pt_wV, pt_wM;
  // does point (i,j,k) belong to the boundary phase of a stage (free-surface rows, tiles next to an inter-rank face)?
  auto bnd = [&](int i, int j, int k) -> int {
    if (c->free_top && !c->fuse_top && k >= g.nk2 - 3) return 1;   // k_top runs in the boundary phase; fused rows belong to their tile
    const int tx = (i - g.ni1) / TILE_X, ty = (j - g.nj1) / TILE_Y;
    if (tx < 0 || ty < 0 || tx >= c->ntx || ty >= c->nty) return 1;
    return (c->neigh[0] >= 0 && tx == 0) || (c->neigh[1] >= 0 && tx == c->ntx - 1) || (c->neigh[2] >= 0 && ty == 0) ||
           (c->neigh[3] >= 0 && ty == c->nty - 1);
  };
  const int H = s.ext_half_npoint;
  std::vector<float> ext;
  for (int is = 0; is < s.total_number; is++) {
    int si = s.si[is], sj = s.sj[is], sk = s.sk[is];
    if (s.itype_spatial_ext == CGFD_SRC_SPATIAL_POINT) {
      size_t ip = si + sj * L + sk * S;
      float wV = 0.0f, wM = 0.0f;
      if (s.force_actived && (s.is_surface_force_strict == 0 || sk < g.nk2)) wV = slw[ip] / jac[ip];
      if (s.moment_actived) wM = (float)(1.0 / jac[ip]);
      pt_iptr.push_back(dev_index(c, (int64_t)ip)); pt_src.push_back(is); pt_wV.push_back(wV); pt_wM.push_back(wM); pt_bnd.push_back(bnd(si, sj, sk));
    } else {
      int k2 = (sk + H < g.nk2) ? H : g.nk2 - sk;
      norm_delt3d_z2fre(ext, s.si_inc[is], s.sj_inc[is], s.sk_inc[is], s.ext_func_coef, H, k2);
      size_t ie = 0;
      for (int ke = -H; ke <= k2; ke++)
        for (int je = -H; je <= H; je++)
          for (int ix = -H; ix <= H; ix++, ie++) {
            int i = si + ix, j = sj + je, k = sk + ke;
            if (i < g.ni1 || i > g.ni2 || j < g.nj1 || j > g.nj2 || k < g.nk1 || k > g.nk2) continue;
            size_t ip = i + j * L + k * S;
            float coef = ext[ie], wV = 0.0f, wM = 0.0f;
            if (s.force_actived && (s.is_surface_force_strict == 0 || k < g.nk2)) wV = coef * slw[ip] / jac[ip];
            if (s.moment_actived) wM = coef / jac[ip];
            pt_iptr.push_back(dev_index(c, (int64_t)ip)); pt_src.push_back(is); pt_wV.push_back(wV); pt_wM.push_back(wM); pt_bnd.push_back(bnd(i, j, k));
          }
    }
  }
  {  // boundary-phase points first (stable, so the order of additions into one grid point is kept)
    std::vector<size_t> ord;
    for (size_t n = 0; n < pt_bnd.size(); n++) if (pt_bnd[n]) ord.push_back(n);
    c->src_nb = (int)ord.size();
    for (size_t n = 0; n < pt_bnd.size(); n++) if (!pt_bnd[n]) ord.push_back(n);
    std::vector<int64_t> a(ord.size()); std::vector<int> b(ord.size()); std::vector<float> v(ord.size()), m(ord.size());
    for (size_t n = 0; n < ord.size(); n++) { a[n] = pt_iptr[ord[n]]; b[n] = pt_src[ord[n]]; v[n] = pt_wV[ord[n]]; m[n] = pt_wM[ord[n]]; }
    pt_iptr.swap(a); pt_src.swap(b); pt_wV.swap(v); pt_wM.swap(m);
  }
  SrcDev &d = c->src;
  d.nsrc = s.total_number; d.max_nt = s.max_nt; d.max_stage = s.max_stage;
  d.force_actived = s.force_actived; d.moment_actived = s.moment_actived;
  const size_t ntab = (size_t)s.total_number * s.max_nt * s.max_stage;
  int rc = 0;
  rc |= upload(c, (int **)&d.it_begin, s.it_begin, s.total_number);
  rc |= upload(c, (int **)&d.it_end, s.it_end, s.total_number);
  if (s.force_actived) {
    rc |= upload(c, (float **)&d.Fx, s.Fx, ntab); rc |= upload(c, (float **)&d.Fy, s.Fy, ntab); rc |= upload(c, (float **)&d.Fz, s.Fz, ntab);
  }
  if (s.moment_actived) {
    rc |= upload(c, (float **)&d.Mxx, s.Mxx, ntab); rc |= upload(c, (float **)&d.Myy, s.Myy, ntab);
    rc |= upload(c, (float **)&d.Mzz, s.Mzz, ntab); rc |= upload(c, (float **)&d.Mxz, s.Mxz, ntab);
    rc |= upload(c, (float **)&d.Myz, s.Myz, ntab); rc |= upload(c, (float **)&d.Mxy, s.Mxy, ntab);
  }
  d.npts = (int)pt_iptr.size();
  rc |= upload(c, (int64_t **)&d.pt_iptr, pt_iptr.data(), pt_iptr.size());
  rc |= upload(c, (int **)&d.pt_src, pt_src.data(), pt_src.size());
  rc |= upload(c, (float **)&d.pt_wV, pt_wV.data(), pt_wV.size());
  rc |= upload(c, (float **)&d.pt_wM, pt_wM.data(), pt_wM.size());
  c->has_src = d.npts > 0;

  // surface forces (forward/src_t.c:153-314)
  if (s.total_number_surface_force > 0 && s.force_actived && p->free_top) {
    std::vector<int> sf_src, sf_slot, sf_p2; std::vector<float> sf_c, sf_cj;
    for (int n = 0; n < s.total_number_surface_force; n++) {
      int is = s.force_rate_indx[n];
      int si = s.si[is], sj = s.sj[is], sk = s.sk[is];
      if (s.itype_spatial_ext == CGFD_SRC_SPATIAL_POINT) {
        size_t ip = si + sj * L + sk * S;
        sf_src.push_back(is); sf_slot.push_back(n); sf_p2.push_back(si + sj * (int)L);
        sf_c.push_back(1.0f); sf_cj.push_back(1.0f / jac[ip]);
      } else {
        int kext = g.nk2 - sk;
        norm_delt3d_z2fre(ext, s.si_inc[is], s.sj_inc[is], s.sk_inc[is], s.ext_func_coef, H, kext);
        int n1 = 2 * H + 1;
        size_t ie = (size_t)(kext + H) * n1 * n1;
        int k = sk + kext;
        for (int je = -H; je <= H; je++)
          for (int ix = -H; ix <= H; ix++, ie++) {
            int i = si + ix, j = sj + je;
            if (i < 0 || i >= g.nx || j < 0 || j >= g.ny) continue;
            size_t ip = i + j * L + k * S;
            sf_src.push_back(is); sf_slot.push_back(n); sf_p2.push_back(i + j * (int)L);
            sf_c.push_back(ext[ie]); sf_cj.push_back(ext[ie] / jac[ip]);
          }
      }
    }
    const size_t nrate = (size_t)s.total_number_surface_force * s.max_nt * s.max_stage;
    rc |= upload(c, (float **)&d.Fx_rate, s.Fx_rate, nrate);
    rc |= upload(c, (float **)&d.Fy_rate, s.Fy_rate, nrate);
    rc |= upload(c, (float **)&d.Fz_rate, s.Fz_rate, nrate);
    d.nsurf_pts = (int)sf_src.size();
    rc |= upload(c, (int **)&d.sf_src, sf_src.data(), sf_src.size());
    rc |= upload(c, (int **)&d.sf_rate_slot, sf_slot.data(), sf_slot.size());
    rc |= upload(c, (int **)&d.sf_iptr2d, sf_p2.data(), sf_p2.size());
    rc |= upload(c, (float **)&d.sf_coef, sf_c.data(), sf_c.size());
    rc |= upload(c, (float **)&d.sf_coef_over_jac, sf_cj.data(), sf_cj.size());
    c->has_surf = d.nsurf_pts > 0;
  }
  return rc;
}

static int check_fd(const cgfd_fd_t &fd)
{
  const int i0[5] = {-1, 0, 1, 2, 3}, i1[5] = {-3, -2, -1, 0, 1};
  for (int n = 0; n < 5; n++)
    if (fd.indx[0][n] != i0[n] || fd.indx[1][n] != i1[n])
      return fail("fd tables: the interior operators must span offsets {-1..3} and {-3..1} (forward/fd_t.c:89-113)");
  if (fd.lay_len[1][0] != 2 || fd.lay_len[1][1] != 2 || fd.lay_len[2][0] != 3 || fd.lay_len[2][1] != 3)
    return fail("fd tables: near-surface operators must have 2 and 3 points (forward/fd_t.c:75-88)");
  const int l1[2][2] = {{0, 1}, {-1, 0}}, l2[2][3] = {{0, 1, 2}, {-2, -1, 0}};
  for (int d = 0; d < 2; d++) {
    for (int n = 0; n < 2; n++) if (fd.lay_indx[1][d][n] != l1[d][n]) return fail("fd tables: unexpected 2-point offsets");
    for (int n = 0; n < 3; n++) if (fd.lay_indx[2][d][n] != l2[d][n]) return fail("fd tables: unexpected 3-point offsets");
  }
  for (int p = 0; p < 8; p++) for (int s = 0; s < 4; s++) for (int a = 0; a < 3; a++)
    if (fd.dir[p][s][a] != 0 && fd.dir[p][s][a] != 1) return fail("fd tables: direction index must be 0 or 1");
  return 0;
}

static int stage_alloc(cgfd_b200_ctx *c);
extern "C" int cgfd_b200_create(const cgfd_problem_t *p, int device, cgfd_b200_ctx **out)
{
  *out = nullptr;
  if (!p || p->abi_version != CGFD_ABI_VERSION) return fail("cgfd_b200_create: ABI version mismatch");
  const int med = med_of(p->medium_type);
  if (med < 0) return fail("cgfd_b200_create: unknown medium type " + std::to_string(p->medium_type));
  {
    const int N = (med == MED_VIS) ? p->nmaxwell : 0;
    const int want_media = (med == MED_ISO) ? 3 : (med == MED_VTI) ? 6 : (med == MED_ANISO) ? 22 : 3 + 2 * N;
    if (med == MED_VIS && (N < 1 || N > CGFD_MAX_MAXWELL)) return fail("cgfd_b200_create: visco-elastic medium needs 1..8 Maxwell bodies");
    if (p->nmedia != want_media) return fail("cgfd_b200_create: medium type " + std::to_string(p->medium_type) + " expects " + std::to_string(want_media) + " media arrays");
    if (p->ncmp != 9 + 6 * N) return fail("cgfd_b200_create: ncmp must be 9 + 6 * nmaxwell");
    for (int m = 0; m < p->nmedia; m++) if (!p->media[m]) return fail("cgfd_b200_create: media array missing");
  }
  int ndev = cgfd_b200_device_count();
  if (ndev <= 0) return fail("cgfd_b200_create: no CUDA device visible; this library has no CPU fallback");
  if (device < 0 || device >= ndev) return fail("cgfd_b200_create: bad device ordinal");
  if (check_fd(p->fd)) return 1;
  CK(cudaSetDevice(device));
#define CKD(call)                                                                                  \
  do {                                                                                             \
    cudaError_t e_ = (call);                                                                       \
    if (e_ != cudaSuccess) {                                                                       \
      cgfd_b200_destroy(c);                                                                        \
      return fail(std::string(#call) + ": " + cudaGetErrorString(e_) + " (" __FILE__ ":" + std::to_string(__LINE__) + ")"); \
    }                                                                                              \
  } while (0)
  cgfd_b200_ctx *c = new cgfd_b200_ctx();
  c->device = device;
  c->g = p->grid; c->fd = p->fd; c->dt = p->dt;
  c->medium = p->medium_type; c->med = med; c->nmaxwell = (med == MED_VIS) ? p->nmaxwell : 0; c->ncmp = p->ncmp; c->nmedia = p->nmedia;
  for (int n = 0; n < CGFD_MAX_MAXWELL; n++) c->wl[n] = p->visco_wl[n];
  c->free_top = p->free_top; c->timg_mode = p->timg_mode;
  for (int n = 0; n < 4; n++) c->neigh[n] = p->neigh[n];
  const cgfd_grid_t &g = c->g;
  if (g.ni1 != 3 || g.nj1 != 3 || g.nk1 != 3 || g.ni2 != g.nx - 4 || g.nj2 != g.ny - 4 || g.nk2 != g.nz - 4) {
    delete c; return fail("cgfd_b200_create: expected 3 ghost layers on every side");
  }
  c->shift = (32 - g.ni1 % 32) % 32;
  c->PX = ((g.nx + c->shift + 31) / 32) * 32;
  c->hslice = (size_t)g.nx * g.ny; c->hV = c->hslice * g.nz;
  c->slice = (size_t)c->PX * g.ny; c->V = c->slice * g.nz;
  if (const char *e = getenv("CGFD_ZCHUNK")) c->zchunk = atoi(e);
  if (const char *e = getenv("CGFD_WAVES")) c->plan_waves = atoi(e);
  if (const char *e = getenv("CGFD_MINCHUNK")) c->plan_minchunk = atoi(e);
  if (const char *e = getenv("CGFD_LPT")) c->plan_lpt = atoi(e);
  if (const char *e = getenv("CGFD_VARIANT")) c->variant = atoi(e);
  if (const char *e = getenv("CGFD_OVERLAP")) c->overlap = atoi(e);
  if (const char *e = getenv("CGFD_L2MODE")) c->l2mode = atoi(e);
  if (const char *e = getenv("CGFD_TOPPAR")) c->toppar = atoi(e);
  if (const char *e = getenv("CGFD_FUSE_TOP")) c->fuse_top = atoi(e) != 0;
  if (c->fuse_top < 0) c->fuse_top = (c->med == MED_ISO);
  if (c->fuse_top) c->toppar = 0;
  if (const char *e = getenv("CGFD_TOP_STREAM")) c->top_stream = atoi(e) != 0;
  if (const char *e = getenv("CGFD_TOP_ROWS")) { c->top_rows = atoi(e); if (c->top_rows < 4) c->top_rows = 4; }
  if (const char *e = getenv("CGFD_PROFILE_DUMP")) c->prof_dump = fopen(e, "a");
  {
    int lo = 0, hi = 0;
    CKD(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    CKD(cudaStreamCreateWithPriority(&c->st, cudaStreamNonBlocking, lo));
    CKD(cudaStreamCreateWithPriority(&c->st2, cudaStreamNonBlocking, hi));
    CKD(cudaStreamCreateWithPriority(&c->st3, cudaStreamNonBlocking, hi));
    CKD(cudaStreamCreateWithFlags(&c->st_io, cudaStreamNonBlocking));
    for (int b = 0; b < 2; b++) {
      CKD(cudaEventCreateWithFlags(&c->stage_full[b], cudaEventDisableTiming));
      CKD(cudaEventCreateWithFlags(&c->stage_free[b], cudaEventDisableTiming));
    }
    CKD(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
    CKD(cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming));
    CKD(cudaEventCreateWithFlags(&c->ev_join3, cudaEventDisableTiming));
    CKD(cudaEventCreateWithFlags(&c->ev_fork3, cudaEventDisableTiming));
  }
  {
    cudaDeviceProp prop;
    CKD(cudaGetDeviceProperties(&prop, device));
    c->nsm = prop.multiProcessorCount;
  }
  c->ntx = (g.ni2 - g.ni1 + 1 + TILE_X - 1) / TILE_X; c->nty = (g.nj2 - g.nj1 + 1 + TILE_Y - 1) / TILE_Y;
  if (kernels_init(c->med)) { cgfd_b200_destroy(c); return fail("cgfd_b200_create: cudaFuncSetAttribute failed (needs sm_100 shared memory sizes)"); }
  CKD(cudaEventCreate(&c->run0)); CKD(cudaEventCreate(&c->run1));

  FdConst fc;
  for (int d = 0; d < 2; d++) {
    for (int n = 0; n < 5; n++) fc.coef[d][n] = p->fd.coef[d][n];
    for (int n = 0; n < 2; n++) fc.lay2[d][n] = p->fd.lay_coef[1][d][n];
    for (int n = 0; n < 3; n++) fc.lay3[d][n] = p->fd.lay_coef[2][d][n];
  }
  CKD(cudaMemcpyToSymbol(c_fd, &fc, sizeof(fc)));

  int rc = 0;
  for (int l = 0; l < 4; l++) rc |= upload(c, &c->lev[l], (const float *)nullptr, c->V * c->ncmp);
  rc |= upload(c, &c->metric_blk, (const float *)nullptr, c->V * NMETRIC);
  rc |= upload(c, &c->media_blk, (const float *)nullptr, c->V * p->nmedia);
  if (rc) { cgfd_b200_destroy(c); return 1; }
  for (int m = 0; m < NMETRIC; m++) {   // device order: metric_dev_slot (cgfd_dev.cuh)
    float *base = c->metric_blk + (size_t)metric_dev_slot(m) * c->V;
    c->metric[m] = base + c->shift;
    rc |= copy_in3d(c, base, p->metric[m], 1, c->st);
  }
  for (int m = 0; m < p->nmedia; m++) {
    c->media[m] = c->media_blk + (size_t)m * c->V + c->shift;
    rc |= copy_in3d(c, c->media_blk + (size_t)m * c->V, p->media[m], 1, c->st);
  }
  CKD(cudaStreamSynchronize(c->st));
  if (rc) { cgfd_b200_destroy(c); return 1; }
  if (p->graves_Qs) {
    if (med == MED_VIS) { cgfd_b200_destroy(c); return fail("cgfd_b200_create: Graves Qs attenuation applies to the elastic media, not to the GMB visco-elastic one"); }
    if (upload(c, &c->qatt, (const float *)nullptr, c->V)) { cgfd_b200_destroy(c); return 1; }
    if (copy_in3d(c, c->qatt, p->graves_Qs, 1, c->st)) { cgfd_b200_destroy(c); return 1; }
    const float coef = (float)(-M_PI * (double)p->graves_Qs_freq * (double)p->dt);   // float coef = - PI * md->visco_Qs_freq * dt, PI a double literal
    k_graves_factor<<<(unsigned)((c->V + 255) / 256), 256, 0, c->st>>>(c->qatt, c->V, coef);
    CKD(cudaStreamSynchronize(c->st));
  }
  // grid class: do the four metric arrays that vanish on a vertically deformed grid vanish here? (physical points only:
  // the interior kernel uses the metric point-wise; the free-surface kernel stays general)
  {
    int *flag = nullptr;
    if (upload(c, &flag, (const int *)nullptr, 1)) { cgfd_b200_destroy(c); return 1; }
    const int zero_arrays[4] = {M_XIY, M_XIZ, M_ETX, M_ETZ};
    for (int n = 0; n < 4; n++) {
      dim3 blk(128), grd((g.ni2 - g.ni1 + 128) / 128, g.nj2 - g.nj1 + 1, g.nk2 - g.nk1 + 1);
      k_any_nonzero<<<grd, blk, 0, c->st>>>(c->metric[zero_arrays[n]], c->PX, g.ny, g.ni1, g.ni2, g.nj1, g.nk1, flag);
    }
    int hflag = 1;
    CKD(cudaMemcpyAsync(&hflag, flag, sizeof(int), cudaMemcpyDeviceToHost, c->st));
    CKD(cudaStreamSynchronize(c->st));
    c->gz = (hflag == 0);
    if (const char *e = getenv("CGFD_GZ")) { if (atoi(e) == 0) c->gz = 0; }
  }
  // tensor maps of the TMA kernel
  {
    int mrc = 0;
    for (int l = 0; l < 4 && !mrc; l++) {
      // CGFD_L2PROMO_CUR: promotion of the halo boxes alone (their 16-byte x halo drags the neighbour tile's whole 128-byte line in)
      mrc |= make_map(c, &c->map_halo[l], c->lev[l], c->ncmp, TILE_X + 2 * HALO_X, TILE_Y + 4, 9, false, "CGFD_L2PROMO_CUR");
      mrc |= make_map(c, &c->map_cen[l], c->lev[l], c->ncmp, TILE_X, TILE_Y, 9);
      mrc |= make_map(c, &c->map_out[l], c->lev[l], c->ncmp, TILE_X, TILE_Y, 9, true);
    }
    if (!mrc) mrc |= make_map(c, &c->map_met, c->metric_blk + c->V /* skip jac */, 9, TILE_X, TILE_Y, 9);
    if (!mrc) mrc |= make_map(c, &c->map_met5, c->metric_blk + c->V, 9, TILE_X, TILE_Y, 5);
    const int ntile = (med == MED_ISO || med == MED_VIS) ? 3 : p->nmedia;   // media arrays staged per plane (Med<MED>::NTILE)
    if (!mrc) mrc |= make_map(c, &c->map_med, c->media_blk, p->nmedia, TILE_X, TILE_Y, ntile);
    if (med == MED_VIS && c->nmaxwell <= VIS_MAX_STAGED && !mrc) {
      // staged attenuation: tensors that start at the first memory variable of a level / at Ylam[0]
      const int nj = 6 * c->nmaxwell;
      for (int l = 0; l < 4 && !mrc; l++) {
        mrc |= make_map(c, &c->map_jcen[l], c->lev[l] + 9 * c->V, nj, TILE_X, TILE_Y, nj);
        mrc |= make_map(c, &c->map_jout[l], c->lev[l] + 9 * c->V, nj, TILE_X, TILE_Y, nj, true);
      }
      if (!mrc) mrc |= make_map(c, &c->map_y, c->media_blk + 3 * c->V, 2 * c->nmaxwell, TILE_X, TILE_Y, 2 * c->nmaxwell);
      c->vis_staged = 1;
      if (const char *e = getenv("CGFD_VIS_STAGED")) c->vis_staged = atoi(e) != 0;
    }
    c->have_maps = (mrc == 0);
    if (mrc) { cgfd_b200_destroy(c); return 1; }
  }

  for (int idim = 0; idim < 3; idim++) for (int is = 0; is < 2; is++) {
    const cgfd_pml_face_t &f = p->pml[idim][is];
    PmlFaceHost &h = c->pml[idim][is];
    if (!f.enabled) continue;
    h.on = 1; h.nlay = f.nlay;
    int r[6] = {g.ni1, g.ni2, g.nj1, g.nj2, g.nk1, g.nk2};
    if (is == 0) r[2 * idim + 1] = r[2 * idim] + f.nlay; else r[2 * idim] = r[2 * idim + 1] - f.nlay;
    memcpy(h.r, r, sizeof(r));
    int ext = (idim == 0 ? g.ni2 - g.ni1 : idim == 1 ? g.nj2 - g.nj1 : g.nk2 - g.nk1) + 1;
    if (2 * (f.nlay + 1) > ext) { cgfd_b200_destroy(c); return fail("cgfd_b200_create: PML slabs of opposite faces overlap"); }
    h.siz = (size_t)(r[1] - r[0] + 1) * (r[3] - r[2] + 1) * (r[5] - r[4] + 1);
    rc |= upload(c, &h.A, f.A, f.nlay + 1); rc |= upload(c, &h.B, f.B, f.nlay + 1); rc |= upload(c, &h.D, f.D, f.nlay + 1);
    for (int l = 0; l < 4; l++) rc |= upload(c, &h.aux[l], (const float *)nullptr, h.siz * AUX_REC);
  }
  if (c->free_top) {
    if (!p->matVx2Vz || !p->matVy2Vz || !p->matF2Vz) { cgfd_b200_destroy(c); return fail("cgfd_b200_create: free surface needs matVx2Vz/matVy2Vz/matF2Vz"); }
    rc |= upload(c, &c->mats[0], p->matVx2Vz, c->hslice * 9);
    rc |= upload(c, &c->mats[1], p->matVy2Vz, c->hslice * 9);
    rc |= upload(c, &c->mats[2], p->matF2Vz, c->hslice * 9);
    if (p->matD) rc |= upload(c, &c->mats[3], p->matD, c->hslice * 9);
    rc |= upload(c, &c->PG, (const float *)nullptr, c->hslice * 15);
    rc |= upload(c, &c->Dis, (const float *)nullptr, c->hslice * 3);
  }
  if (p->ablexp_enabled) {
    c->ablexp = 1; memcpy(c->ablexp_blk, p->ablexp_blk, sizeof(c->ablexp_blk));
    rc |= upload(c, &c->Ex, p->ablexp_Ex, g.nx); rc |= upload(c, &c->Ey, p->ablexp_Ey, g.ny); rc |= upload(c, &c->Ez, p->ablexp_Ez, g.nz);
  }
  rc |= setup_sources(c, p);
  if (c->has_surf) rc |= upload(c, &c->srcslice, (const float *)nullptr, c->hslice * 6);
  // the two one-component staging buffers of set_wavefield / get_wavefield: every device allocation of a context happens here, so
  // that the first whole-wavefield transfer of a run costs what the later ones cost (bench.py e2e: 34 ms against 22 ms)
  rc |= stage_alloc(c);
  if (rc) { cgfd_b200_destroy(c); return 1; }
  CKD(cudaDeviceSynchronize());
  *out = c;
  return 0;
}

#undef CKD
extern "C" void cgfd_b200_destroy(cgfd_b200_ctx *c)
{
  if (!c) return;
  cudaSetDevice(c->device);
  cudaDeviceSynchronize();
  if (c->halo) halo_destroy(c->halo);
  for (void *q : c->owned) cudaFree(q);
  if (c->boxbuf) cudaFree(c->boxbuf);
  for (auto e : c->ev) cudaEventDestroy(e);
  if (c->run0) cudaEventDestroy(c->run0);
  if (c->run1) cudaEventDestroy(c->run1);
  if (c->ev_rec) cudaEventDestroy(c->ev_rec);
  for (auto &b : c->blk) if (b.done) cudaEventDestroy(b.done);
  if (c->ev_fork) cudaEventDestroy(c->ev_fork);
  if (c->ev_join) cudaEventDestroy(c->ev_join);
  if (c->ev_join3) cudaEventDestroy(c->ev_join3);
  if (c->ev_fork3) cudaEventDestroy(c->ev_fork3);
  for (SnapTap *t : c->snaps) {
    for (int r = 0; r < SNAP_RING; r++) { cudaFree(t->ring[r]); cudaEventDestroy(t->packed[r]); cudaEventDestroy(t->copied[r]); }
    delete t;
  }
  for (int b = 0; b < 2; b++) {
    if (c->dd.loaded[b]) cudaEventDestroy(c->dd.loaded[b]);
    if (c->dd.used[b]) cudaEventDestroy(c->dd.used[b]);
    if (c->stage[b]) cudaFree(c->stage[b]);
    if (c->stage_full[b]) cudaEventDestroy(c->stage_full[b]);
    if (c->stage_free[b]) cudaEventDestroy(c->stage_free[b]);
  }
  if (c->st_io) cudaStreamDestroy(c->st_io);
  if (c->prof_dump) fclose(c->prof_dump);
  if (c->st2) cudaStreamDestroy(c->st2);
  if (c->st3) cudaStreamDestroy(c->st3);
  if (c->st) cudaStreamDestroy(c->st);
  delete c;
}

// Whole-wavefield transfers go component by component through two contiguous staging buffers: one 1-D copy per component
// on the I/O stream (full PCIe rate from pinned memory; a pitched 2-D copy of 1.6 KB rows is several times slower) and a
// re-pitch kernel on the compute stream, double-buffered so that the copy of component c+1 overlaps the kernel of c.
static int stage_alloc(cgfd_b200_ctx *c)
{
  for (int b = 0; b < 2; b++)
    if (!c->stage[b]) CK(cudaMalloc((void **)&c->stage[b], c->hV * sizeof(float)));
  return 0;
}
static void repitch(cgfd_b200_ctx *c, float *dev_base, float *flat, int to_padded)
{
  const size_t rows = (size_t)c->g.ny * c->g.nz;
  dim3 blk(128), grd((c->g.nx + 127) / 128, (unsigned)(rows < 16384 ? rows : 16384));
  k_repitch<<<grd, blk, 0, c->st>>>(dev_base + c->shift, flat, c->g.nx, c->PX, rows, to_padded);
}
extern "C" int cgfd_b200_set_wavefield(cgfd_b200_ctx *c, const float *w)
{
  CK(cudaSetDevice(c->device));
  if (stage_alloc(c)) return 1;
  CK(cudaStreamSynchronize(c->st));
  for (int m = 0; m < c->ncmp; m++) {
    const int b = m & 1;
    if (m >= 2) CK(cudaStreamWaitEvent(c->st_io, c->stage_free[b], 0));
    CK(cudaMemcpyAsync(c->stage[b], w + (size_t)m * c->hV, c->hV * sizeof(float), cudaMemcpyHostToDevice, c->st_io));
    CK(cudaEventRecord(c->stage_full[b], c->st_io));
    CK(cudaStreamWaitEvent(c->st, c->stage_full[b], 0));
    repitch(c, c->lev[c->ipre] + (size_t)m * c->V, c->stage[b], 1);
    CK(cudaEventRecord(c->stage_free[b], c->st));
  }
  CK(cudaStreamSynchronize(c->st));
  CK(cudaGetLastError());
  return 0;
}
extern "C" int cgfd_b200_get_wavefield(cgfd_b200_ctx *c, float *w)
{
  CK(cudaSetDevice(c->device));
  if (stage_alloc(c)) return 1;
  CK(cudaStreamSynchronize(c->st_io));
  for (int m = 0; m < c->ncmp; m++) {
    const int b = m & 1;
    if (m >= 2) CK(cudaStreamWaitEvent(c->st, c->stage_free[b], 0));
    repitch(c, c->lev[c->ipre] + (size_t)m * c->V, c->stage[b], 0);
    CK(cudaEventRecord(c->stage_full[b], c->st));
    CK(cudaStreamWaitEvent(c->st_io, c->stage_full[b], 0));
    CK(cudaMemcpyAsync(w + (size_t)m * c->hV, c->stage[b], c->hV * sizeof(float), cudaMemcpyDeviceToHost, c->st_io));
    CK(cudaEventRecord(c->stage_free[b], c->st_io));
  }
  CK(cudaStreamSynchronize(c->st_io));
  CK(cudaStreamSynchronize(c->st));
  CK(cudaGetLastError());
  return 0;
}
extern "C" size_t cgfd_b200_pml_aux_size(cgfd_b200_ctx *c, int idim, int is) { return c->pml[idim][is].on ? c->pml[idim][is].siz * 9 : 0; }
// host [9][slab] (the reference's bdrypml_auxvar_t order, forward/bdry_t.c:300-306) <-> device [slab][AUX_REC] records
static int aux_to_device(cgfd_b200_ctx *c, PmlFaceHost &h, float *dev, const float *host)
{
  std::vector<float> rec(h.siz * AUX_REC, 0.0f);
  for (int cmp = 0; cmp < 9; cmp++)
    for (size_t n = 0; n < h.siz; n++) rec[n * AUX_REC + aux_slot(cmp)] = host[(size_t)cmp * h.siz + n];
  CK(cudaMemcpyAsync(dev, rec.data(), rec.size() * sizeof(float), cudaMemcpyHostToDevice, c->st));
  CK(cudaStreamSynchronize(c->st));
  return 0;
}
static int aux_to_host(cgfd_b200_ctx *c, PmlFaceHost &h, float *host, const float *dev)
{
  std::vector<float> rec(h.siz * AUX_REC);
  CK(cudaStreamSynchronize(c->st));
  CK(cudaMemcpyAsync(rec.data(), dev, rec.size() * sizeof(float), cudaMemcpyDeviceToHost, c->st));
  CK(cudaStreamSynchronize(c->st));
  for (int cmp = 0; cmp < 9; cmp++)
    for (size_t n = 0; n < h.siz; n++) host[(size_t)cmp * h.siz + n] = rec[n * AUX_REC + aux_slot(cmp)];
  return 0;
}
extern "C" int cgfd_b200_set_pml_aux(cgfd_b200_ctx *c, int idim, int is, const float *aux)
{
  PmlFaceHost &h = c->pml[idim][is];
  if (!h.on) return fail("set_pml_aux: face has no PML");
  CK(cudaSetDevice(c->device));
  return aux_to_device(c, h, h.aux[c->ipre], aux);
}
extern "C" int cgfd_b200_get_pml_aux(cgfd_b200_ctx *c, int idim, int is, float *aux)
{
  PmlFaceHost &h = c->pml[idim][is];
  if (!h.on) return fail("get_pml_aux: face has no PML");
  CK(cudaSetDevice(c->device));
  return aux_to_host(c, h, aux, h.aux[c->ipre]);
}
extern "C" int cgfd_b200_get_pml_aux_rhs(cgfd_b200_ctx *c, int idim, int is, float *aux)
{
  PmlFaceHost &h = c->pml[idim][is];
  if (!h.on) return fail("get_pml_aux_rhs: face has no PML");
  CK(cudaSetDevice(c->device));
  return aux_to_host(c, h, aux, h.aux[c->ib]);
}

// ---- one RK stage -----------------------------------------------------------------------------
static void fill_args(cgfd_b200_ctx *c, StageArgs &P)
{
  memset(&P, 0, sizeof(P));
  const cgfd_grid_t &g = c->g;
  P.nx = g.nx; P.ny = g.ny; P.nz = g.nz; P.shift = c->shift;
  P.ni1 = g.ni1; P.ni2 = g.ni2; P.nj1 = g.nj1; P.nj2 = g.nj2; P.nk1 = g.nk1; P.nk2 = g.nk2;
  P.siz_line = c->PX; P.siz_slice = c->slice; P.siz_vol = c->V;
  for (int m = 0; m < NMETRIC; m++) P.metric[m] = c->metric[m];
  for (int m = 0; m < c->nmedia; m++) P.media[m] = c->media[m];
  P.nmaxwell = c->nmaxwell; P.vis_staged = c->vis_staged;
  P.qatt = c->qatt ? c->qatt + c->shift : nullptr;
  for (int n = 0; n < MAX_MAXWELL; n++) P.wl[n] = c->wl[n];
  P.free_top = c->free_top; P.fuse_top = c->free_top && c->fuse_top; P.timg_mode = c->timg_mode; P.l2mode = c->l2mode;
  P.matVx2Vz = c->mats[0]; P.matVy2Vz = c->mats[1]; P.matF2Vz = c->mats[2]; P.matD = c->mats[3];
  if (c->has_surf) {
    P.TxSrc = c->srcslice; P.TySrc = c->srcslice + c->hslice; P.TzSrc = c->srcslice + 2 * c->hslice;
    P.VxSrc = c->srcslice + 3 * c->hslice; P.VySrc = c->srcslice + 4 * c->hslice; P.VzSrc = c->srcslice + 5 * c->hslice;
  }
  for (int idim = 0; idim < 3; idim++) for (int is = 0; is < 2; is++) {
    PmlFaceHost &h = c->pml[idim][is];
    PmlFaceDev &d = P.pml[idim][is];
    d.on = h.on;
    if (!h.on) continue;
    d.i1 = h.r[0]; d.i2 = h.r[1]; d.j1 = h.r[2]; d.j2 = h.r[3]; d.k1 = h.r[4]; d.k2 = h.r[5];
    d.sni = d.i2 - d.i1 + 1; d.snj = d.j2 - d.j1 + 1; d.siz = h.siz;
    d.A = h.A; d.B = h.B; d.D = h.D;
  }
}

// tile rectangles of the interior kernel: the tile columns / rows that touch an inter-rank face (boundary phase, at most
// four rectangles) and the rest. rect = {bx0, bx1, by0, by1}.
static int split_tiles(const cgfd_b200_ctx *c, bool split, int bnd[4][4], int inner[4])
{
  int x0 = 0, x1 = c->ntx, y0 = 0, y1 = c->nty, n = 0;
  if (split) {
    if (c->neigh[0] >= 0 && x1 - x0 > 0) { int r[4] = {x0, x0 + 1, 0, c->nty}; memcpy(bnd[n++], r, sizeof(r)); x0++; }
    if (c->neigh[1] >= 0 && x1 - x0 > 0) { int r[4] = {x1 - 1, x1, 0, c->nty}; memcpy(bnd[n++], r, sizeof(r)); x1--; }
    if (c->neigh[2] >= 0 && y1 - y0 > 0) { int r[4] = {x0, x1, y0, y0 + 1}; memcpy(bnd[n++], r, sizeof(r)); y0++; }
    if (c->neigh[3] >= 0 && y1 - y0 > 0) { int r[4] = {x0, x1, y1 - 1, y1}; memcpy(bnd[n++], r, sizeof(r)); y1--; }
  }
  inner[0] = x0; inner[1] = x1; inner[2] = y0; inner[3] = y1;
  return n;
}

// Launch plan of a tile rectangle: z chunks so that the launch has at least `waves` waves of resident blocks with chunks no
// shorter than `minchunk` rows, and the longest-job-first block order: blocks whose tile meets an x / y PML slab run the PML copy
// of the loop body on every plane (~3x the time of a plain block-plane), so they go first. Pure host logic (also exported as
// cgfd_b200_launch_plan for the CPU tests). pml_r[idim][iside] = {on, first index, last index} of the slab along its axis.
// lpt = 2 (default): within each class the tiles are taken in bands of one wave (148 x blocks_per_sm tiles); a band runs through
// its z chunks in the direction the kernel marches (dz = 1 upwards, 0 downwards) before the next band starts, so that the
// 4 planes a chunk re-reads for its zeta queue are the ones the previous chunk of the same tile has just fetched (L2 hits
// instead of a second DRAM read). lpt = 1: chunk-major order.
constexpr int PLAN_CHUNK_ROWS = 25, PLAN_CHUNK_ROWS_2 = 18;
static void compute_plan(const cgfd_grid_t &g, const int pml_r[2][2][3], int nk /* rows of the launch */, int nsm, int blocks_per_sm, int waves,
                         int minchunk, int zchunk_explicit, int lpt, int dz, const int rect[4], int *zchunk, std::vector<int> *order)
{
  *zchunk = 0; order->clear();
  const int bx = rect[1] - rect[0], by = rect[3] - rect[2];
  if (bx <= 0 || by <= 0 || nk <= 0) return;
  // Chunks of ~25 rows whatever the grid: a band of tiles restarts together at every chunk boundary, which keeps neighbouring tiles
  // within a few planes of each other, so that the halo lines they share are still in L2 (800x800x400: chunks of 198 rows, which
  // the wave count alone would allow, run 12 % slower than chunks of 25 -- profiles/r2_experiments.txt r2p). Small rectangles get
  // more, shorter chunks (never under `minchunk` rows) until they make `waves` waves of resident blocks.
  // measured r2r: two resident blocks per SM (iso, VTI) 16 / 18 / 20 / 22 / 25 rows: 4.867 / 4.870 / 4.876 / 4.893 / 4.899 ms per step
  // at 400x400x200, 7.82 / 7.81 / 7.82 / - / 7.76 Gpt/s at 800x800x400; one resident block (aniso, visco): 25 rows over 20 by 0.6 %
  const int rows = blocks_per_sm >= 2 ? PLAN_CHUNK_ROWS_2 : PLAN_CHUNK_ROWS;
  int nzc = (nk + rows / 2) / rows;
  if (nzc < 1) nzc = 1;
  if (zchunk_explicit > 0) nzc = (nk + zchunk_explicit - 1) / zchunk_explicit;
  else
    while (nzc < nk && (long)bx * by * nzc < (long)nsm * blocks_per_sm * waves && nk / (nzc + 1) >= minchunk) nzc++;
  *zchunk = (nk + nzc - 1) / nzc;
  nzc = (nk + *zchunk - 1) / *zchunk;
  if (!lpt) return;
  std::vector<int> cls[2];   // tiles (y * bx + x): [0] meet an x / y PML slab, [1] do not
  for (int y = 0; y < by; y++)
    for (int x = 0; x < bx; x++) {
      const int i0 = g.ni1 + (rect[0] + x) * TILE_X, j0 = g.nj1 + (rect[2] + y) * TILE_Y;
      bool pml = false;
      for (int sd = 0; sd < 2; sd++) {
        pml |= pml_r[0][sd][0] && i0 <= pml_r[0][sd][2] && i0 + TILE_X - 1 >= pml_r[0][sd][1];
        pml |= pml_r[1][sd][0] && j0 <= pml_r[1][sd][2] && j0 + TILE_Y - 1 >= pml_r[1][sd][1];
      }
      cls[pml ? 0 : 1].push_back(y * bx + x);
    }
  for (int cl = 0; cl < 2; cl++) {
    const size_t nt = cls[cl].size();
    const size_t band = (lpt >= 2) ? (size_t)nsm * blocks_per_sm : (nt ? nt : 1);
    for (size_t b0 = 0; b0 < nt; b0 += band)
      for (int zi = 0; zi < nzc; zi++) {
        const int z = (lpt >= 2 && dz == 0) ? nzc - 1 - zi : zi;
        for (size_t t = b0; t < nt && t < b0 + band; t++) order->push_back(z * by * bx + cls[cl][t]);
      }
  }
}
static const LaunchPlan *plan_for(cgfd_b200_ctx *c, const int rect[4], int dz)
{
  std::array<int, 5> key = {rect[0], rect[1], rect[2], rect[3], dz};
  auto it = c->plans.find(key);
  if (it != c->plans.end()) return &it->second;
  LaunchPlan pl;
  int pml_r[2][2][3];
  for (int d = 0; d < 2; d++) for (int sd = 0; sd < 2; sd++) {
    const PmlFaceHost &f = c->pml[d][sd];
    pml_r[d][sd][0] = f.on; pml_r[d][sd][1] = f.r[2 * d]; pml_r[d][sd][2] = f.r[2 * d + 1];
  }
  std::vector<int> order;
  const cgfd_grid_t &g = c->g;
  const bool fused = c->free_top && c->fuse_top;
  const int nk_all = (c->free_top && !fused ? g.nk2 - 4 : g.nk2) - g.nk1 + 1;
  const int bps = blocks_per_sm(c->med);
  if (!fused) {
    compute_plan(g, pml_r, nk_all, c->nsm, bps, c->plan_waves, c->plan_minchunk, c->zchunk, c->plan_lpt, dz, rect, &pl.zchunk, &order);
  } else {
    // The top `top_rows` rows (the four free-surface rows and a few below them, so that a block's start-up is spread over more than
    // four planes) are the launch of the TOPK kernels: one chunk per tile, the tiles that meet an x / y PML slab first. Kept
    // short whatever the grid: the TOPK kernels are not the hot loop's best build (section 3.1 of DESIGN.md). The rows below
    // are chunked as a launch of their own.
    const int ntop = nk_all < c->top_rows ? nk_all : c->top_rows;
    const int kt0 = g.nk2 - ntop + 1;
    pl.ktop0 = kt0; pl.zchunk_top = ntop;
    std::vector<int> otop;
    int zt = 0;
    compute_plan(g, pml_r, ntop, c->nsm, bps, c->plan_waves, c->plan_minchunk, ntop, c->plan_lpt, dz, rect, &zt, &otop);
    if (!otop.empty() && upload(c, &pl.order_top, otop.data(), otop.size())) return nullptr;
    if (kt0 > g.nk1) compute_plan(g, pml_r, kt0 - g.nk1, c->nsm, bps, c->plan_waves, c->plan_minchunk, c->zchunk, c->plan_lpt, dz, rect, &pl.zchunk, &order);
    else pl.zchunk = ntop;
  }
  if (!order.empty() && upload(c, &pl.order, order.data(), order.size())) return nullptr;
  return &(c->plans[key] = pl);
}
extern "C" int cgfd_b200_launch_plan(const cgfd_grid_t *g, const int pml_nlay[3][2], int free_top, int blocks_per_sm, int dz, const int rect[4],
                                     int *zchunk, int *order, int capacity)
{
  if (!g || !pml_nlay || !rect || !zchunk) return -1;
  int pml_r[2][2][3];
  for (int d = 0; d < 2; d++) for (int sd = 0; sd < 2; sd++) {
    const int a1 = d == 0 ? g->ni1 : g->nj1, a2 = d == 0 ? g->ni2 : g->nj2, nl = pml_nlay[d][sd];
    pml_r[d][sd][0] = nl > 0; pml_r[d][sd][1] = sd == 0 ? a1 : a2 - nl; pml_r[d][sd][2] = sd == 0 ? a1 + nl : a2;
  }
  cgfd_b200_ctx defaults;
  std::vector<int> ord;
  compute_plan(*g, pml_r, (free_top ? g->nk2 - 4 : g->nk2) - g->nk1 + 1, 148 /* B200; the context itself asks the device */, blocks_per_sm, defaults.plan_waves, defaults.plan_minchunk, 0, defaults.plan_lpt, dz, rect, zchunk, &ord);
  if (order) for (size_t n = 0; n < ord.size() && (int)n < capacity; n++) order[n] = ord[n];
  return (int)ord.size();
}

// distributed sources: the resident time block that holds step `it` (srcdd is skipped outside every loaded block, like
// dd_is_valid = 0 past dd_max_nt), and the launch of points sel[first .. first+count) of it on stream s
static int dd_block_of(const cgfd_b200_ctx *c, int it)
{
  if (c->dd.n <= 0) return -1;
  for (int bf = 0; bf < 2; bf++)
    if (c->dd.it_first[bf] >= 0 && it >= c->dd.it_first[bf] && it < c->dd.it_first[bf] + c->dd.nt[bf]) return bf;
  return -1;
}
static int launch_dd(cgfd_b200_ctx *c, int bf, int it, int istage, int first, int count, int itmp, int iend, float a, float b, int kind,
                     const float *qatt, cudaStream_t s)
{
  const size_t row = ((size_t)(it - c->dd.it_first[bf]) * c->dd.max_stage + istage) * c->dd.n;
  CK(cudaStreamWaitEvent(s, c->dd.loaded[bf], 0));
  k_srcdd_inject<<<(count + 127) / 128, 128, 0, s>>>(count, c->dd.sel + first, c->dd.iptr, c->dd.wV, c->dd.rjac,
      c->dd.vi_on ? c->dd.vi[bf] + row * 3 : nullptr, c->dd.mij_on ? c->dd.mij[bf] + row * 6 : nullptr,
      c->lev[itmp] + c->shift, c->lev[iend] + c->shift, a, b, c->V, kind, qatt);
  return 0;
}

// The interior kernel on one tile rectangle. With the fused free surface these are two launches: the top chunk of every tile (rows
// [ktop0, nk2], TOPK kernels, on s_top -- first: its blocks are the long ones) and the chunks below it (on s_lean); they touch
// disjoint rows and may run side by side. Otherwise one launch of the rows below the free-surface kernel's.
static int launch_rect(cgfd_b200_ctx *c, StageArgs &P, const TmaMaps *mp, const int *dir, int kind, const int rect[4], cudaStream_t s_lean,
                       cudaStream_t s_top, int *nl)
{
  const LaunchPlan *pl = plan_for(c, rect, dir[2]);
  if (!pl) return 1;
  const int nk1 = c->g.nk1, nk2 = c->g.nk2;
  if (P.fuse_top) {
    P.kbeg = pl->ktop0; P.kend = nk2; P.order = pl->order_top;
    launch_main(c->med, P, mp, dir, kind, c->gz, 1, pl->zchunk_top, rect, s_top, nullptr, nullptr, nl);
    P.kbeg = nk1; P.kend = pl->ktop0 - 1; P.order = pl->order;
    launch_main(c->med, P, mp, dir, kind, c->gz, 0, pl->zchunk, rect, s_lean, nullptr, nullptr, nl);
  } else {
    P.kbeg = nk1; P.kend = P.free_top ? nk2 - 4 : nk2; P.order = pl->order;
    launch_main(c->med, P, mp, dir, kind, c->gz, 0, pl->zchunk, rect, s_lean, nullptr, nullptr, nl);
  }
  return 0;
}

// Launch everything of stage `istage` of step `it`; level roles: icur -> (itmp, iend), ipre.
// Two phases on two streams:
//   boundary phase (st2, high priority): the free-surface rows, the tiles next to inter-rank faces, the source points
//        inside them, then -- when `halo_w` is given -- pack, NCCL send/recv and unpack of the ghosts of halo_w;
//   interior phase (st): every other tile, the remaining source points.
// The interior kernel never touches what the boundary phase writes (disjoint points; ghosts are written by the unpack
// only), so the halo exchange is hidden behind it. The stage ends with st waiting for st2.
static int run_stage(cgfd_b200_ctx *c, StageArgs &P, int it, int ipair, int istage, int kind, int icur, int ipre, int itmp,
                     int iend, float a, float b, float cc, float *halo_w, int halo_dx, int halo_dy)
{
  const int sh = c->shift;
  P.cur = c->lev[icur] + sh; P.pre = c->lev[ipre] + sh; P.tmp = c->lev[itmp] + sh; P.end = c->lev[iend] + sh;
  P.a = a; P.b = b; P.c = cc;
  TmaMaps maps;
  if (c->have_maps) {
    maps.cur = c->map_halo[icur]; maps.za = c->map_cen[icur]; maps.pre = c->map_cen[ipre]; maps.end = c->map_cen[iend];
    maps.met = c->map_met; maps.met5 = c->map_met5; maps.med = c->map_med;
    maps.out_tmp = c->map_out[itmp]; maps.out_end = c->map_out[iend];
    if (c->vis_staged) {
      maps.jcur = c->map_jcen[icur]; maps.jpre = c->map_jcen[ipre]; maps.jend = c->map_jcen[iend]; maps.ymed = c->map_y;
      maps.jout_tmp = c->map_jout[itmp]; maps.jout_end = c->map_jout[iend];
    }
  }
  const TmaMaps *mp = c->have_maps ? &maps : nullptr;
  for (int idim = 0; idim < 3; idim++) for (int is = 0; is < 2; is++) {
    PmlFaceHost &h = c->pml[idim][is];
    if (!h.on) continue;
    PmlFaceDev &d = P.pml[idim][is];
    d.aux_cur = h.aux[icur]; d.aux_pre = h.aux[ipre]; d.aux_tmp = h.aux[itmp]; d.aux_end = h.aux[iend];
  }
  int nl = 0;
  if (c->has_surf) {
    CK(cudaMemsetAsync(c->srcslice, 0, c->hslice * 6 * sizeof(float), c->st));
    k_src_surface<<<(c->src.nsurf_pts + 127) / 128, 128, 0, c->st>>>(c->src, it, istage, c->srcslice, c->srcslice + c->hslice,
        c->srcslice + 2 * c->hslice, c->srcslice + 3 * c->hslice, c->srcslice + 4 * c->hslice, c->srcslice + 5 * c->hslice);
    nl++;
  }
  const int *dir = c->fd.dir[ipair][istage];
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  if (c->profiling) {
    if (c->ev_used + 2 > c->ev.size()) {
      for (int n = 0; n < 2; n++) { cudaEvent_t e; CK(cudaEventCreate(&e)); c->ev.push_back(e); }
    }
    e0 = c->ev[c->ev_used]; e1 = c->ev[c->ev_used + 1]; c->ev_used += 2;
  }
  // two streams: with a halo exchange to hide, or (toppar) just to run the latency-bound free-surface kernel beside the interior one
  const bool two = (c->overlap && halo_w) || (c->toppar && c->free_top);
  // toppar = 2 (single rank): the free-surface kernel is issued AFTER the interior kernel on the low-priority stream, so that its
  // blocks fill the SMs the interior kernel's last wave leaves idle instead of running ahead of it
  const bool top_late = !halo_w && c->toppar == 2 && c->free_top;
  cudaStream_t sb = two ? c->st2 : c->st;   // stream of the boundary phase
  int bnd[4][4], inner[4];
  const int nb = split_tiles(c, halo_w != nullptr, bnd, inner);
  if (two) { CK(cudaEventRecord(c->ev_fork, c->st)); CK(cudaStreamWaitEvent(c->st2, c->ev_fork, 0)); }
  // ---- boundary phase
  if (!top_late) launch_top(c->med, P, dir, kind, sb, &nl);
  for (int n = 0; n < nb; n++)
    if (launch_rect(c, P, mp, dir, kind, bnd[n], sb, sb, &nl)) return 1;
  // Sources are pushed through the RK axpy AFTER the stage kernel of their point has written it (k_src_inject). The points were
  // classified at create time by the declared neighbours (setup_sources); the tiles are only split when an exchange follows.
  // With a split, the boundary-class points go in here, before the ghosts leave; without one, every point waits for the
  // whole-rectangle interior kernel (which would otherwise overwrite the boundary-class ones).
  const bool split = halo_w != nullptr;
  const int ddbf = dd_block_of(c, it);   // resident time block of the distributed sources that holds this step, or -1
  if (split && c->has_src && c->src_nb > 0) {
    k_src_inject<<<(c->src_nb + 127) / 128, 128, 0, sb>>>(c->src, 0, c->src_nb, it, istage, c->lev[itmp] + sh, c->lev[iend] + sh, a, b, c->V, kind, P.qatt);
    nl++;
  }
  if (split && ddbf >= 0 && c->dd.nb > 0) {
    // the reference adds srcdd to the RHS before the RK update and the pack (forward/sv_curv_col_el_iso.c:190-199): dd points in
    // the boundary-phase tiles must be in before the ghosts leave
    if (launch_dd(c, ddbf, it, istage, 0, c->dd.nb, itmp, iend, a, b, kind, P.qatt, sb)) return 1;
    nl++;
  }
  if (halo_w) {
    if (halo_exchange(c->halo, halo_w, halo_dx, halo_dy, sb)) return fail(halo_error());
    nl += halo_launches_per_exchange(c->halo);
  }
  if (two && !top_late) CK(cudaEventRecord(c->ev_join, c->st2));
  // ---- interior phase
  {
    // top_late: the interior kernel takes the high-priority stream, the free-surface kernel follows on the low-priority one
    cudaStream_t sm = top_late ? c->st2 : c->st;
    // fused free surface: the top-chunk launch of the interior tiles runs on a stream of its own beside the chunks below it
    const bool top3 = P.fuse_top && c->top_stream;
    if (e0) CK(cudaEventRecord(e0, sm));   // the timed region of the dominant kernel: both of its launches
    if (top3) { CK(cudaEventRecord(c->ev_fork3, sm)); CK(cudaStreamWaitEvent(c->st3, c->ev_fork3, 0)); }
    if (launch_rect(c, P, mp, dir, kind, inner, sm, top3 ? c->st3 : sm, &nl)) return 1;
    if (top3) { CK(cudaEventRecord(c->ev_join3, c->st3)); CK(cudaStreamWaitEvent(sm, c->ev_join3, 0)); }
    if (e1) CK(cudaEventRecord(e1, sm));
    if (top_late) {
      CK(cudaEventRecord(c->ev_join, c->st2));
      launch_top(c->med, P, dir, kind, c->st, &nl);
    }
  }
  // without a split the boundary-class points (free-surface rows, which another stream may own) are injected here as well
  if (!split && two) CK(cudaStreamWaitEvent(c->st, c->ev_join, 0));
  {
    const int first = split ? c->src_nb : 0, cnt = c->has_src ? c->src.npts - first : 0;
    if (cnt > 0) {
      k_src_inject<<<(cnt + 127) / 128, 128, 0, c->st>>>(c->src, first, cnt, it, istage, c->lev[itmp] + sh, c->lev[iend] + sh, a, b, c->V, kind, P.qatt);
      nl++;
    }
    const int dfirst = split ? c->dd.nb : 0, dcnt = c->dd.n - dfirst;
    if (ddbf >= 0 && dcnt > 0) {
      if (launch_dd(c, ddbf, it, istage, dfirst, dcnt, itmp, iend, a, b, kind, P.qatt, c->st)) return 1;
      nl++;
    }
  }
  if (ddbf >= 0) {
    // both phases are done with the block's tables once st has joined st2
    if (two) CK(cudaStreamWaitEvent(c->st, c->ev_join, 0));
    CK(cudaEventRecord(c->dd.used[ddbf], c->st));
  }
  if (two) CK(cudaStreamWaitEvent(c->st, c->ev_join, 0));
  c->total_launches += nl;
  CK(cudaGetLastError());
  return 0;
}

static int drain_profile(cgfd_b200_ctx *c)
{
  if (!c->ev_used) return 0;
  CK(cudaStreamSynchronize(c->st));
  for (size_t n = 0; n + 1 < c->ev_used; n += 2) {
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, c->ev[n], c->ev[n + 1]));
    c->main_ms += ms; c->main_launches++;
    if (c->prof_dump) fprintf(c->prof_dump, "%.4f\n", ms);   // CGFD_PROFILE_DUMP: one line per timed launch, in launch order
  }
  if (c->prof_dump) fflush(c->prof_dump);
  c->ev_used = 0;
  return 0;
}

// enqueue nsteps RK4 steps (no host synchronisation except when a snapshot ring is full)
static int run_enqueue(cgfd_b200_ctx *c, int it0, int nsteps)
{
  CK(cudaSetDevice(c->device));
  StageArgs P;
  fill_args(c, P);
  const float dt = c->dt;
  CK(cudaEventRecord(c->run0, c->st));
  for (int it = it0; it < it0 + nsteps; it++) {
    const int ipair = it % CGFD_NUM_PAIRS;
    for (int s = 0; s < CGFD_NUM_STAGES; s++) {
      const int kind = (s == 0) ? KIND_FIRST : (s == 1) ? KIND_MID : (s == 2) ? KIND_THIRD : KIND_LAST;
      // weight of (w_cur - w_pre) = a_{s-1} dt h_{s-1} in the w_end update of stages 1 and 3
      const float cc = (s == 1 || s == 3) ? (float)((double)c->fd.rk_b[s - 1] / (double)c->fd.rk_a[s - 1]) : 0.0f;
      const int icur = (s == 0) ? c->ipre : (s & 1) ? c->ia : c->ib;
      const int itmp = (s & 1) ? c->ib : c->ia;
      const float a = c->fd.rk_a[s] * dt, b = c->fd.rk_b[s] * dt;
      // ghosts of the level the NEXT rhs evaluation reads, widths of that evaluation's operator
      // (forward/drv_rk_curv_col.c:193-199, 308-312, 448-469)
      const int np = (s != CGFD_NUM_STAGES - 1) ? ipair : (it + 1) % CGFD_NUM_PAIRS;
      const int ns = (s != CGFD_NUM_STAGES - 1) ? s + 1 : 0;
      float *hw = c->halo ? ((s != CGFD_NUM_STAGES - 1) ? c->lev[itmp] : c->lev[c->iend]) + c->shift : nullptr;
      if (run_stage(c, P, it, ipair, s, kind, icur, c->ipre, itmp, c->iend, a, b, cc, hw, c->fd.dir[np][ns][0], c->fd.dir[np][ns][1])) return 1;
    }
    float *wnew = c->lev[c->iend] + c->shift, *wold = c->lev[c->ipre] + c->shift;
    const cgfd_grid_t &g = c->g;
    if (c->med == MED_VIS && c->free_top) {
      // after the halo exchange of w_end, like the reference (forward/drv_rk_curv_col.c:419-445)
      if (!c->mats[3]) return fail("visco-elastic free surface needs matD");
      dim3 blk(128), grd((g.ni2 - g.ni1 + 128) / 128, g.nj2 - g.nj1 + 1);
      k_vis_free<<<grd, blk, 0, c->st>>>(wnew, c->V, c->PX, g.nx, g.ny, g.ni1, g.ni2, g.nj1, g.nj2, g.nk2, c->mats[3]);
      c->total_launches++;
    }
    if (c->ablexp) {
      for (int n = 0; n < 6; n++) {
        const int *B = c->ablexp_blk[n];
        if (!B[0]) continue;
        dim3 blk(64), grd((B[2] - B[1] + 64) / 64, B[4] - B[3] + 1, B[6] - B[5] + 1);
        k_ablexp<<<grd, blk, 0, c->st>>>(wnew, c->V, c->ncmp, c->PX, g.ny, B[1], B[2], B[3], B[4], B[5], B[6], c->Ex, c->Ey, c->Ez);
        c->total_launches++;
      }
    }
    if (c->free_top) {
      dim3 blk(128), grd((g.ni2 - g.ni1 + 128) / 128, g.nj2 - g.nj1 + 1);
      k_pg<<<grd, blk, 0, c->st>>>(wnew, wold, c->V, c->PX, g.nx, g.ny, g.ni1, g.ni2, g.nj1, g.nj2, g.nk2, dt, c->PG, c->Dis);
      c->total_launches++;
    }
    if (c->nrec > 0 && c->rec_count < c->rec_max_nt) {
      int n = c->nrec * c->ncmp;
      k_record<<<(n + 127) / 128, 128, 0, c->st>>>(wnew, c->V, c->ncmp, c->nrec, c->rec_iptr,
                                                   c->rec + (size_t)c->rec_count * c->ncmp * c->nrec);
      c->rec_count++; c->total_launches++;
    }
    for (SnapTap *t : c->snaps) {
      // io_snap_nc_put / io_slice_nc_put (forward/io_funcs.c:991-1268): frame of w_end every tinv steps from it1 on
      if (it < t->it1 || (it - t->it1) % t->tinv != 0 || t->nframes >= t->max_frames) continue;
      const int r = (int)(t->nissued % SNAP_RING);
      if (t->nissued >= SNAP_RING) CK(cudaEventSynchronize(t->copied[r]));   // the slot's previous frame has left the device
      const size_t tot = t->cmp_elems;
      for (int m = 0; m < t->ncmp; m++)
        k_pack_box<<<(unsigned)((tot + 255) / 256), 256, 0, c->st>>>(wnew + (size_t)t->cmps[m] * c->V, c->PX, g.ny, t->box[0], t->box[1], t->box[2],
                                                                    t->box[3], t->box[4], t->box[5], t->box[6], t->box[7], t->box[8],
                                                                    t->ring[r] + (size_t)m * tot);
      c->total_launches += t->ncmp;
      CK(cudaEventRecord(t->packed[r], c->st));
      CK(cudaStreamWaitEvent(c->st_io, t->packed[r], 0));
      CK(cudaMemcpyAsync(t->host + (size_t)t->nframes * t->ncmp * tot, t->ring[r], (size_t)t->ncmp * tot * sizeof(float),
                         cudaMemcpyDeviceToHost, c->st_io));
      CK(cudaEventRecord(t->copied[r], c->st_io));
      t->nframes++; t->nissued++;
    }
    // swap levels n <-> n+1 (forward/drv_rk_curv_col.c:530-542)
    int t = c->ipre; c->ipre = c->iend; c->iend = t;
    if (c->profiling && c->ev_used > 4096) { if (drain_profile(c)) return 1; }
  }
  CK(cudaEventRecord(c->run1, c->st));
  CK(cudaGetLastError());
  return 0;
}
extern "C" int cgfd_b200_sync(cgfd_b200_ctx *c)
{
  CK(cudaSetDevice(c->device));
  CK(cudaStreamSynchronize(c->st));
  CK(cudaStreamSynchronize(c->st_io));   // every snapshot frame / record copy of the enqueued steps is in host memory on return
  CK(cudaGetLastError());
  float ms = 0;
  if (cudaEventElapsedTime(&ms, c->run0, c->run1) == cudaSuccess) c->last_run_ms = ms; else cudaGetLastError();
  if (drain_profile(c)) return 1;
  return 0;
}
extern "C" int cgfd_b200_run(cgfd_b200_ctx *c, int it0, int nsteps)
{
  if (run_enqueue(c, it0, nsteps)) return 1;
  return cgfd_b200_sync(c);
}
extern "C" int cgfd_b200_run_async(cgfd_b200_ctx *c, int it0, int nsteps, float *rec_out)
{
  const int first = c->rec_count;
  if (run_enqueue(c, it0, nsteps)) return 1;
  // everything these steps produce for the host (snapshot frames: already on the I/O stream; record samples: below) is complete
  // when the block's event on the I/O stream is
  if (!c->ev_rec) CK(cudaEventCreateWithFlags(&c->ev_rec, cudaEventDisableTiming));
  CK(cudaEventRecord(c->ev_rec, c->st));
  CK(cudaStreamWaitEvent(c->st_io, c->ev_rec, 0));
  if (rec_out && c->nrec > 0 && c->rec_count > first) {
    const size_t row = (size_t)c->ncmp * c->nrec;
    CK(cudaMemcpyAsync(rec_out, c->rec + (size_t)first * row, (size_t)(c->rec_count - first) * row * sizeof(float), cudaMemcpyDeviceToHost, c->st_io));
  }
  auto &b = c->blk[c->blk_next];
  c->blk_next = (c->blk_next + 1) % 4;
  if (!b.done) CK(cudaEventCreateWithFlags(&b.done, cudaEventDisableTiming | cudaEventBlockingSync));
  CK(cudaEventRecord(b.done, c->st_io));
  b.it_last = it0 + nsteps - 1;
  return 0;
}
extern "C" int cgfd_b200_wait_block(cgfd_b200_ctx *c, int it_last)
{
  CK(cudaSetDevice(c->device));
  for (auto &b : c->blk)
    if (b.done && b.it_last == it_last) { CK(cudaEventSynchronize(b.done)); return 0; }
  return fail("wait_block: no asynchronous block in flight ends at step " + std::to_string(it_last));
}
extern "C" int cgfd_b200_snapshot_set_output(cgfd_b200_ctx *c, int id, float *host_out, int max_frames)
{
  if (id < 0 || id >= (int)c->snaps.size() || !host_out || max_frames <= 0) return fail("snapshot_set_output: bad arguments");
  SnapTap *t = c->snaps[id];
  // frames already enqueued keep the destination they were enqueued with
  t->host = host_out; t->max_frames = max_frames; t->nframes = 0;
  return 0;
}
extern "C" int cgfd_b200_host_alloc(size_t bytes, void **out)
{
  *out = nullptr;
  CK(cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocDefault));
  return 0;
}
extern "C" void cgfd_b200_host_free(void *p) { if (p) cudaFreeHost(p); }

extern "C" int cgfd_b200_onestage(cgfd_b200_ctx *c, int it, int ipair, int istage, const float *w_cur, float *rhs)
{
  CK(cudaSetDevice(c->device));
  const size_t nb = c->V * c->ncmp * sizeof(float);
  const int sh = c->shift;
  // rhs = 0 + 1*L(w_cur): the stage update with w_pre := 0, a := 1, b := 0. Uses the level buffers as
  // scratch (the context's wavefield is overwritten; PML aux level n is preserved).
  const int icur = c->ia, iout = c->ib, izero = c->iend, iz2 = c->ipre;
  if (copy_in3d(c, c->lev[icur], w_cur, c->ncmp, c->st)) return 1;
  CK(cudaMemsetAsync(c->lev[iout], 0, nb, c->st));
  CK(cudaMemsetAsync(c->lev[izero], 0, nb, c->st));
  CK(cudaMemsetAsync(c->lev[iz2], 0, nb, c->st));
  for (int idim = 0; idim < 3; idim++) for (int is = 0; is < 2; is++) {
    PmlFaceHost &h = c->pml[idim][is];
    if (!h.on) continue;
    const size_t ab = h.siz * AUX_REC * sizeof(float);
    if (!h.zero) { if (upload(c, &h.zero, (const float *)nullptr, h.siz * AUX_REC)) return 1; }
    CK(cudaMemcpyAsync(h.aux[icur], h.aux[c->ipre], ab, cudaMemcpyDeviceToDevice, c->st));
    CK(cudaMemsetAsync(h.aux[iout], 0, ab, c->st));
    CK(cudaMemsetAsync(h.aux[izero], 0, ab, c->st));
  }
  StageArgs P;
  fill_args(c, P);
  P.cur = c->lev[icur] + sh; P.pre = c->lev[iz2] + sh; P.tmp = c->lev[iout] + sh; P.end = c->lev[izero] + sh;
  TmaMaps maps;
  if (c->have_maps) {
    maps.cur = c->map_halo[icur]; maps.za = c->map_cen[icur]; maps.pre = c->map_cen[iz2]; maps.end = c->map_cen[izero];
    maps.met = c->map_met; maps.met5 = c->map_met5; maps.med = c->map_med;
    maps.out_tmp = c->map_out[iout]; maps.out_end = c->map_out[izero];
    if (c->vis_staged) {
      maps.jcur = c->map_jcen[icur]; maps.jpre = c->map_jcen[iz2]; maps.jend = c->map_jcen[izero]; maps.ymed = c->map_y;
      maps.jout_tmp = c->map_jout[iout]; maps.jout_end = c->map_jout[izero];
    }
  }
  // aux: cur = copy of level n, pre = zeros, tmp = out, end = scratch
  // run_stage sets aux pointers from level indices; patch aux_pre to the zero buffer afterwards is not
  // possible through indices, so do it by hand here.
  P.a = 1.0f; P.b = 0.0f; P.c = 0.0f;
  for (int idim = 0; idim < 3; idim++) for (int is = 0; is < 2; is++) {
    PmlFaceHost &h = c->pml[idim][is];
    if (!h.on) continue;
    PmlFaceDev &d = P.pml[idim][is];
    d.aux_cur = h.aux[icur]; d.aux_pre = h.zero; d.aux_tmp = h.aux[iout]; d.aux_end = h.aux[izero];
  }
  int nl = 0;
  if (c->has_surf) {
    CK(cudaMemsetAsync(c->srcslice, 0, c->hslice * 6 * sizeof(float), c->st));
    k_src_surface<<<(c->src.nsurf_pts + 127) / 128, 128, 0, c->st>>>(c->src, it, istage, c->srcslice, c->srcslice + c->hslice,
        c->srcslice + 2 * c->hslice, c->srcslice + 3 * c->hslice, c->srcslice + 4 * c->hslice, c->srcslice + 5 * c->hslice);
  }
  const int *dir = c->fd.dir[ipair][istage];
  const int whole[4] = {0, c->ntx, 0, c->nty};
  launch_top(c->med, P, dir, KIND_THIRD, c->st, &nl);
  if (launch_rect(c, P, &maps, dir, KIND_THIRD, whole, c->st, c->st, &nl)) return 1;
  if (c->has_src)
    k_src_inject<<<(c->src.npts + 127) / 128, 128, 0, c->st>>>(c->src, 0, c->src.npts, it, istage, c->lev[iout] + sh, c->lev[izero] + sh, 1.0f, 0.0f, c->V, KIND_THIRD, nullptr);
  CK(cudaGetLastError());
  if (copy_out3d(c, rhs, c->lev[iout], c->ncmp, c->st)) return 1;
  CK(cudaStreamSynchronize(c->st));
  // leave a clean state: wavefield levels zero
  CK(cudaMemsetAsync(c->lev[icur], 0, nb, c->st));
  CK(cudaMemsetAsync(c->lev[iout], 0, nb, c->st));
  CK(cudaStreamSynchronize(c->st));
  return 0;
}

// ---- taps -------------------------------------------------------------------------------------
extern "C" int cgfd_b200_set_record_points(cgfd_b200_ctx *c, int n, const int64_t *iptr, int max_nt)
{
  CK(cudaSetDevice(c->device));
  c->nrec = 0; c->rec_count = 0;
  for (void **q : {(void **)&c->rec_iptr, (void **)&c->rec}) {   // a second call replaces the first one's buffers
    if (!*q) continue;
    CK(cudaStreamSynchronize(c->st));
    for (auto it = c->owned.begin(); it != c->owned.end(); ++it) if (*it == *q) { c->owned.erase(it); break; }
    cudaFree(*q); *q = nullptr;
  }
  if (n <= 0) return 0;
  for (int i = 0; i < n; i++) if (iptr[i] < 0 || (size_t)iptr[i] >= c->hV) return fail("set_record_points: index out of range");
  std::vector<int64_t> dv(n);
  for (int i = 0; i < n; i++) dv[i] = dev_index(c, iptr[i]);
  if (upload(c, &c->rec_iptr, dv.data(), n)) return 1;
  if (upload(c, &c->rec, (const float *)nullptr, (size_t)n * c->ncmp * max_nt)) return 1;
  c->nrec = n; c->rec_max_nt = max_nt;
  return 0;
}
extern "C" int cgfd_b200_get_record(cgfd_b200_ctx *c, int it_first, int nt, float *out)
{
  CK(cudaSetDevice(c->device));
  if (it_first < 0 || it_first + nt > c->rec_count) return fail("get_record: step range not recorded");
  CK(cudaStreamSynchronize(c->st));
  CK(cudaMemcpyAsync(out, c->rec + (size_t)it_first * c->ncmp * c->nrec, (size_t)nt * c->ncmp * c->nrec * sizeof(float), cudaMemcpyDeviceToHost, c->st));
  CK(cudaStreamSynchronize(c->st));
  return 0;
}
extern "C" int cgfd_b200_get_box(cgfd_b200_ctx *c, int icmp, int i1, int ni, int di, int j1, int nj, int dj, int k1, int nk, int dk,
                                 float *out)
{
  CK(cudaSetDevice(c->device));
  const cgfd_grid_t &g = c->g;
  if (icmp < 0 || icmp >= c->ncmp || ni <= 0 || nj <= 0 || nk <= 0 || i1 < 0 || j1 < 0 || k1 < 0 ||
      i1 + (ni - 1) * di >= g.nx || j1 + (nj - 1) * dj >= g.ny || k1 + (nk - 1) * dk >= g.nz)
    return fail("get_box: box outside the array");
  size_t tot = (size_t)ni * nj * nk;
  if (tot > c->boxcap) {
    if (c->boxbuf) cudaFree(c->boxbuf);
    CK(cudaMalloc((void **)&c->boxbuf, tot * sizeof(float)));
    c->boxcap = tot;
  }
  k_pack_box<<<(unsigned)((tot + 255) / 256), 256, 0, c->st>>>(c->lev[c->ipre] + c->shift + (size_t)icmp * c->V, c->PX, g.ny, i1, ni, di, j1, nj,
                                                              dj, k1, nk, dk, c->boxbuf);
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(out, c->boxbuf, tot * sizeof(float), cudaMemcpyDeviceToHost, c->st));
  CK(cudaStreamSynchronize(c->st));
  return 0;
}
// ---- set-up on the device ---------------------------------------------------
extern "C" int cgfd_b200_metric_from_coords(int device, const cgfd_grid_t *g, const float *x, const float *y, const float *z, int fd_len,
                                            const int *fd_indx, const float *fd_coef, float *const metric_out[10])
{
  if (!g || !x || !y || !z || !fd_indx || !fd_coef || !metric_out || fd_len <= 0 || fd_len > 16) return fail("metric_from_coords: bad arguments");
  for (int n = 0; n < fd_len; n++)
    if (abs(fd_indx[n]) > g->ni1 || abs(fd_indx[n]) > g->nj1 || abs(fd_indx[n]) > g->nk1) return fail("metric_from_coords: operator wider than the ghost layers");
  CK(cudaSetDevice(device));
  const size_t V = (size_t)g->nx * g->ny * g->nz, nb = V * sizeof(float);
  float *buf = nullptr; int *dindx = nullptr; float *dcoef = nullptr;
  CK(cudaMalloc((void **)&buf, 13 * nb));
  cudaError_t e = cudaMalloc((void **)&dindx, fd_len * sizeof(int));
  if (e == cudaSuccess) e = cudaMalloc((void **)&dcoef, fd_len * sizeof(float));
  auto done = [&](int rc) { cudaFree(buf); if (dindx) cudaFree(dindx); if (dcoef) cudaFree(dcoef); return rc; };
  if (e != cudaSuccess) return done(fail("metric_from_coords: out of device memory"));
  const float *src[3] = {x, y, z};
  for (int n = 0; n < 3; n++)
    if (cudaMemcpy(buf + n * V, src[n], nb, cudaMemcpyDefault) != cudaSuccess) return done(fail("metric_from_coords: copy of the coordinates failed"));
  cudaMemcpy(dindx, fd_indx, fd_len * sizeof(int), cudaMemcpyHostToDevice);
  cudaMemcpy(dcoef, fd_coef, fd_len * sizeof(float), cudaMemcpyHostToDevice);
  cudaMemset(buf + 3 * V, 0, 10 * nb);
  MetricOut out;
  for (int m = 0; m < 10; m++) out.a[m] = buf + (3 + m) * V;
  dim3 blk(128), grd((g->ni2 - g->ni1 + 128) / 128, g->nj2 - g->nj1 + 1, g->nk2 - g->nk1 + 1);
  k_metric_cal<<<grd, blk>>>(buf, buf + V, buf + 2 * V, g->nx, g->ny, g->ni1, g->ni2, g->nj1, g->nk1, fd_len, dindx, dcoef, out);
  dim3 grm((unsigned)((V + 255) / 256), 1, 10);
  k_metric_mirror<<<grm, 256>>>(out, 0, g->nx, g->ny, g->nz, g->ni1, g->ni2);
  k_metric_mirror<<<grm, 256>>>(out, 1, g->nx, g->ny, g->nz, g->nj1, g->nj2);
  k_metric_mirror<<<grm, 256>>>(out, 2, g->nx, g->ny, g->nz, g->nk1, g->nk2);
  if (cudaDeviceSynchronize() != cudaSuccess || cudaGetLastError() != cudaSuccess) return done(fail("metric_from_coords: kernel failed"));
  for (int m = 0; m < 10; m++)
    if (!metric_out[m] || cudaMemcpy(metric_out[m], out.a[m], nb, cudaMemcpyDefault) != cudaSuccess) return done(fail("metric_from_coords: copy of the result failed"));
  return done(0);
}

extern "C" int cgfd_b200_dvh2dvz(int device, const cgfd_problem_t *p, const float *x, const float *y, const float *z, int fd_len,
                                 const int *fd_indx, const float *fd_coef, float *matVx2Vz, float *matVy2Vz, float *matF2Vz, float *matD)
{
  if (!p || !matVx2Vz || !matVy2Vz) return fail("dvh2dvz: bad arguments");
  const int med = med_of(p->medium_type);
  if (med < 0) return fail("dvh2dvz: unknown medium type");
  if (med == MED_VIS && (!x || !y || !z || !fd_indx || !fd_coef || !matD || fd_len <= 0 || fd_len > 16))
    return fail("dvh2dvz: the visco-elastic medium needs the coordinates, the centred operator and matD");
  const cgfd_grid_t &g = p->grid;
  CK(cudaSetDevice(device));
  const size_t S = (size_t)g.nx * g.ny, off = (size_t)g.nk2 * S;   // the k = nk2 plane of every array
  const int nmed = (med == MED_ISO || med == MED_VIS) ? 2 : (med == MED_VTI) ? 5 : 21;
  const int nin = 9 + nmed + (med == MED_VIS ? 3 : 0);
  float *buf = nullptr; int *dindx = nullptr; float *dcoef = nullptr;
  CK(cudaMalloc((void **)&buf, (size_t)(nin + 4 * 9) * S * sizeof(float)));
  auto done = [&](int rc) { cudaFree(buf); if (dindx) cudaFree(dindx); if (dcoef) cudaFree(dcoef); return rc; };
  DvhArgs a;
  memset(&a, 0, sizeof(a));
  a.med = med; a.nx = g.nx; a.ni1 = g.ni1; a.ni2 = g.ni2; a.nj1 = g.nj1; a.nj2 = g.nj2;
  int slot = 0;
  auto plane = [&](const float *src) -> const float * {
    float *d = buf + (size_t)slot++ * S;
    if (!src || cudaMemcpy(d, src + off, S * sizeof(float), cudaMemcpyDefault) != cudaSuccess) return nullptr;
    return d;
  };
  for (int m = 1; m < 10; m++) if (!(a.metric[m] = plane(p->metric[m]))) return done(fail("dvh2dvz: copy of a metric plane failed"));
  for (int m = 0; m < nmed; m++) if (!(a.media[m] = plane(p->media[m]))) return done(fail("dvh2dvz: copy of a media plane failed"));
  if (med == MED_VIS) {
    if (!(a.x = plane(x)) || !(a.y = plane(y)) || !(a.z = plane(z))) return done(fail("dvh2dvz: copy of a coordinate plane failed"));
    if (cudaMalloc((void **)&dindx, fd_len * sizeof(int)) != cudaSuccess || cudaMalloc((void **)&dcoef, fd_len * sizeof(float)) != cudaSuccess)
      return done(fail("dvh2dvz: out of device memory"));
    cudaMemcpy(dindx, fd_indx, fd_len * sizeof(int), cudaMemcpyHostToDevice);
    cudaMemcpy(dcoef, fd_coef, fd_len * sizeof(float), cudaMemcpyHostToDevice);
    a.fd_len = fd_len; a.fd_indx = dindx; a.fd_coef = dcoef;
  }
  float *out = buf + (size_t)nin * S;
  cudaMemset(out, 0, 4 * 9 * S * sizeof(float));
  a.matVx2Vz = out; a.matVy2Vz = out + 9 * S; a.matF2Vz = out + 18 * S; a.matD = out + 27 * S;
  dim3 blk(128), grd((g.ni2 - g.ni1 + 128) / 128, g.nj2 - g.nj1 + 1);
  k_dvh2dvz<<<grd, blk>>>(a);
  if (cudaDeviceSynchronize() != cudaSuccess || cudaGetLastError() != cudaSuccess) return done(fail("dvh2dvz: kernel failed"));
  float *dst[4] = {matVx2Vz, matVy2Vz, matF2Vz, matD};
  for (int n = 0; n < 4; n++)
    if (dst[n] && cudaMemcpy(dst[n], out + (size_t)n * 9 * S, 9 * S * sizeof(float), cudaMemcpyDefault) != cudaSuccess)
      return done(fail("dvh2dvz: copy of the result failed"));
  return done(0);
}

extern "C" int cgfd_b200_dd_set_points(cgfd_b200_ctx *c, int n, const int64_t *indx, int vi_actived, int mij_actived, int max_stage,
                                       int nt_per_block)
{
  CK(cudaSetDevice(c->device));
  if (c->dd.n > 0) return fail("dd_set_points: already set");
  if (n <= 0 || !indx || max_stage <= 0 || nt_per_block <= 0 || (!vi_actived && !mij_actived)) return fail("dd_set_points: bad arguments");
  std::vector<int64_t> dv(n);
  for (int q = 0; q < n; q++) {
    if (indx[q] < 0 || (size_t)indx[q] >= c->hV) return fail("dd_set_points: index out of range");
    dv[q] = dev_index(c, indx[q]);
  }
  {
    // boundary-phase points first (free-surface rows and the tiles next to a declared inter-rank face: same rule as setup_sources)
    const cgfd_grid_t &g = c->g;
    std::vector<int> sel, rest;
    for (int q = 0; q < n; q++) {
      const int i = (int)(indx[q] % g.nx), j = (int)((indx[q] / g.nx) % g.ny), k = (int)(indx[q] / ((int64_t)g.nx * g.ny));
      const int tx = (i - g.ni1) / TILE_X, ty = (j - g.nj1) / TILE_Y;
      const bool bnd = (c->free_top && !c->fuse_top && k >= g.nk2 - 3) || i < g.ni1 || j < g.nj1 || tx >= c->ntx || ty >= c->nty ||
                       (c->neigh[0] >= 0 && tx == 0) || (c->neigh[1] >= 0 && tx == c->ntx - 1) ||
                       (c->neigh[2] >= 0 && ty == 0) || (c->neigh[3] >= 0 && ty == c->nty - 1);
      (bnd ? sel : rest).push_back(q);
    }
    c->dd.nb = (int)sel.size();
    sel.insert(sel.end(), rest.begin(), rest.end());
    if (upload(c, &c->dd.sel, sel.data(), sel.size())) return 1;
  }
  if (upload(c, &c->dd.iptr, dv.data(), n)) return 1;
  if (upload(c, &c->dd.wV, (const float *)nullptr, n) || upload(c, &c->dd.rjac, (const float *)nullptr, n)) return 1;
  const size_t rows = (size_t)nt_per_block * max_stage * n;
  for (int b = 0; b < 2; b++) {
    if (vi_actived && upload(c, &c->dd.vi[b], (const float *)nullptr, rows * 3)) return 1;
    if (mij_actived && upload(c, &c->dd.mij[b], (const float *)nullptr, rows * 6)) return 1;
    CK(cudaEventCreateWithFlags(&c->dd.loaded[b], cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&c->dd.used[b], cudaEventDisableTiming));
  }
  // 1/rho is the last media array of every medium but the visco-elastic one, where it is the third (cgfd3d_b200.h)
  const float *slw = (c->med == MED_VIS || c->med == MED_ISO) ? c->media[2] : c->media[c->nmedia - 1];
  k_srcdd_weights<<<(n + 127) / 128, 128, 0, c->st>>>(n, c->dd.iptr, slw, c->metric[M_JAC], c->dd.wV, c->dd.rjac);
  CK(cudaStreamSynchronize(c->st));
  c->dd.n = n; c->dd.vi_on = vi_actived; c->dd.mij_on = mij_actived; c->dd.max_stage = max_stage; c->dd.nt_block = nt_per_block;
  return 0;
}
extern "C" int cgfd_b200_dd_load_block(cgfd_b200_ctx *c, int it_first, int nt, const float *vi, const float *mij)
{
  CK(cudaSetDevice(c->device));
  if (c->dd.n <= 0) return fail("dd_load_block: no dd points set");
  if (nt <= 0 || nt > c->dd.nt_block || it_first < 0) return fail("dd_load_block: bad block");
  if ((c->dd.vi_on && !vi) || (c->dd.mij_on && !mij)) return fail("dd_load_block: missing table");
  const int b = c->dd.next;
  const size_t rows = (size_t)nt * c->dd.max_stage * c->dd.n;
  // the buffer may still be read by launches of the block it held before
  if (c->dd.it_first[b] >= 0) CK(cudaStreamWaitEvent(c->st_io, c->dd.used[b], 0));
  if (c->dd.vi_on) CK(cudaMemcpyAsync(c->dd.vi[b], vi, rows * 3 * sizeof(float), cudaMemcpyHostToDevice, c->st_io));
  if (c->dd.mij_on) CK(cudaMemcpyAsync(c->dd.mij[b], mij, rows * 6 * sizeof(float), cudaMemcpyHostToDevice, c->st_io));
  CK(cudaEventRecord(c->dd.loaded[b], c->st_io));
  CK(cudaEventRecord(c->dd.used[b], c->st));   // defined even if no step of this block ever runs
  c->dd.it_first[b] = it_first; c->dd.nt[b] = nt;
  c->dd.next = 1 - b;
  return 0;
}
extern "C" int cgfd_b200_add_snapshot(cgfd_b200_ctx *c, int ncmps, const int *cmps, const int box[9], int it1, int tinv, int max_frames,
                                      float *host_out)
{
  CK(cudaSetDevice(c->device));
  const cgfd_grid_t &g = c->g;
  if (ncmps <= 0 || ncmps > 32 || !cmps || !host_out || tinv <= 0 || max_frames <= 0) { fail("add_snapshot: bad arguments"); return -1; }
  for (int m = 0; m < ncmps; m++) if (cmps[m] < 0 || cmps[m] >= c->ncmp) { fail("add_snapshot: component out of range"); return -1; }
  const int i1 = box[0], ni = box[1], di = box[2], j1 = box[3], nj = box[4], dj = box[5], k1 = box[6], nk = box[7], dk = box[8];
  if (ni <= 0 || nj <= 0 || nk <= 0 || di <= 0 || dj <= 0 || dk <= 0 || i1 < 0 || j1 < 0 || k1 < 0 || i1 + (ni - 1) * di >= g.nx ||
      j1 + (nj - 1) * dj >= g.ny || k1 + (nk - 1) * dk >= g.nz) { fail("add_snapshot: box outside the array"); return -1; }
  SnapTap *t = new SnapTap();
  t->ncmp = ncmps; memcpy(t->cmps, cmps, ncmps * sizeof(int)); memcpy(t->box, box, 9 * sizeof(int));
  t->it1 = it1; t->tinv = tinv; t->max_frames = max_frames; t->host = host_out;
  t->cmp_elems = (size_t)ni * nj * nk;
  for (int r = 0; r < SNAP_RING; r++) {
    t->ring[r] = nullptr;
    if (cudaMalloc((void **)&t->ring[r], t->cmp_elems * ncmps * sizeof(float)) != cudaSuccess ||
        cudaEventCreateWithFlags(&t->packed[r], cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&t->copied[r], cudaEventDisableTiming) != cudaSuccess) { fail("add_snapshot: out of device memory"); return -1; }
  }
  c->snaps.push_back(t);
  return (int)c->snaps.size() - 1;
}
extern "C" int cgfd_b200_snapshot_frames(cgfd_b200_ctx *c, int id)
{
  if (id < 0 || id >= (int)c->snaps.size()) { fail("snapshot_frames: unknown snapshot"); return -1; }
  return c->snaps[id]->nframes;
}
extern "C" int cgfd_b200_get_pg(cgfd_b200_ctx *c, float *pg)
{
  if (!c->PG) return fail("get_pg: no free surface");
  CK(cudaSetDevice(c->device));
  CK(cudaStreamSynchronize(c->st));
  CK(cudaMemcpyAsync(pg, c->PG, c->hslice * 15 * sizeof(float), cudaMemcpyDeviceToHost, c->st));
  CK(cudaStreamSynchronize(c->st));
  return 0;
}

// ---- multi-GPU ----------------------------------------------------------------------------------
extern "C" int cgfd_b200_comm_unique_id(char id[128])
{
  if (halo_unique_id(id)) return fail(halo_error());
  return 0;
}
extern "C" int cgfd_b200_comm_init(cgfd_b200_ctx *c, const char id[128], int rank, int nranks)
{
  CK(cudaSetDevice(c->device));
  if (c->halo) return fail("comm_init: already initialised");
  // the 6*N memory variables of the visco-elastic medium have no stencil; only the 9 wavefield components travel
  // (the reference sends all ncmp, forward/drv_rk_curv_col.c:308)
  c->halo = halo_create(id, rank, nranks, c->neigh, c->g, 9, c->V, c->PX, c->st);
  if (!c->halo) return fail(halo_error());
  return 0;
}

extern "C" int cgfd_b200_halo_plan(const cgfd_grid_t *g, int dirx, int diry, int side, int send_box[6], int recv_box[6])
{
  if (!g || side < 0 || side > 3 || (dirx | diry) & ~1) return fail("halo_plan: bad arguments");
  halo_plan(*g, dirx, diry, side, send_box, recv_box);
  return 0;
}

// ---- measurement --------------------------------------------------------------------------------
extern "C" int cgfd_b200_set_profiling(cgfd_b200_ctx *c, int on)
{
  c->profiling = on; c->main_ms = 0; c->main_launches = 0; c->total_launches = 0;
  return 0;
}
extern "C" int cgfd_b200_get_profile(cgfd_b200_ctx *c, double *ms, int64_t *nmain, int64_t *ntot)
{
  if (ms) *ms = c->main_ms;
  if (nmain) *nmain = c->main_launches;
  if (ntot) *ntot = c->total_launches;
  return 0;
}
extern "C" int cgfd_b200_grid_class(cgfd_b200_ctx *c) { return c->gz; }
extern "C" int cgfd_b200_top_fused(cgfd_b200_ctx *c) { return c->free_top && c->fuse_top; }
extern "C" int cgfd_b200_last_run_ms(cgfd_b200_ctx *c, double *ms) { *ms = c->last_run_ms; return 0; }
extern "C" int cgfd_b200_set_variant(cgfd_b200_ctx *c, const char *name)
{
  if (!name || !*name || !strcmp(name, "default")) { c->variant = 0; return 0; }
  char *endp = nullptr;
  long v = strtol(name, &endp, 10);
  if (endp && *endp == 0) { c->variant = (int)v; return 0; }
  return fail(std::string("set_variant: unknown variant ") + name);
}
