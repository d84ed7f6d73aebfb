// TMA probe: which descriptor / box shapes work on this box. Build: nvcc -gencode arch=compute_100a,code=sm_100a -o tma_probe tma_probe.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda.h>
#include <cuda_runtime.h>
#include "../../cgfd3d_b200/csrc/tma.cuh"
using namespace cgfd;

typedef CUresult (*enc_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                           const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                           CUtensorMapFloatOOBfill);

struct Maps { CUtensorMap a; CUtensorMap b; };

template <int RANK>
__global__ void k_probe(const __grid_constant__ Maps M, const CUtensorMap *gmap, int use_global, int c0, int c1, int c2, int c3,
                        int nfloats, float *out)
{
  extern __shared__ __align__(1024) unsigned char smem[];
  uint64_t *bar = (uint64_t *)(smem + 65536);
  if (threadIdx.x == 0) { mbar_init(bar, 1); mbar_fence_init(); }
  __syncthreads();
  if (threadIdx.x == 0) {
    mbar_expect_tx(bar, nfloats * 4);
    const CUtensorMap *m = use_global ? gmap : &M.a;
    if (RANK == 4) tma_load_4d(smem, m, bar, c0, c1, c2, c3);
    else {
      asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                   ::"r"(smem_u32(smem)), "l"(m), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
    }
  }
  mbar_wait(bar, 0);
  for (int n = threadIdx.x; n < nfloats; n += blockDim.x) out[n] = ((float *)smem)[n];
}

int main(int argc, char **argv)
{
  int only = argc > 1 ? atoi(argv[1]) : -1;
  void *p = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
  enc_fn enc = (enc_fn)p;
  printf("entry point %p q=%d\n", p, (int)q);
  const int PX = 64, NY = 28, NZ = 26, NC = 9;
  size_t V = (size_t)PX * NY * NZ;
  std::vector<float> h(V * NC);
  for (size_t n = 0; n < h.size(); n++) h[n] = (float)(n % 100003);
  float *d, *out; cudaMalloc(&d, h.size() * 4); cudaMalloc(&out, 65536);
  cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
  CUtensorMap *gmap; cudaMalloc(&gmap, sizeof(CUtensorMap));
  cudaFuncSetAttribute(k_probe<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536 + 64);
  cudaFuncSetAttribute(k_probe<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536 + 64);
  struct Case { const char *name; int rank; cuuint32_t box[4]; int c[4]; int use_global; };
  Case cases[] = {
      {"2d box 32x8 c0=32", 2, {32, 8, 1, 1}, {32, 3, 0, 0}, 0},
      {"2d box 32x8 c0=31", 2, {32, 8, 1, 1}, {31, 3, 0, 0}, 0},
      {"2d box 32x8 c0=28", 2, {32, 8, 1, 1}, {28, 3, 0, 0}, 0},
      {"2d box 36x12 c0=32", 2, {36, 12, 1, 1}, {32, 2, 0, 0}, 0},
      {"2d box 36x8 c0=32", 2, {36, 8, 1, 1}, {32, 2, 0, 0}, 0},
      {"2d box 40x8 c0=32", 2, {40, 8, 1, 1}, {32, 2, 0, 0}, 0},
      {"2d box 48x8 c0=32", 2, {48, 8, 1, 1}, {32, 2, 0, 0}, 0},
      {"2d box 64x8 c0=0", 2, {64, 8, 1, 1}, {0, 2, 0, 0}, 0},
      {"2d box 32x12 c0=32", 2, {32, 12, 1, 1}, {32, 2, 0, 0}, 0},
      {"2d box 36x12 c0=31", 2, {36, 12, 1, 1}, {31, 2, 0, 0}, 0},
      {"4d box 32x8x1x9 grid_constant", 4, {32, 8, 1, 9}, {32, 3, 5, 0}, 0},
      {"4d box 36x12x1x9 grid_constant", 4, {36, 12, 1, 9}, {31, 2, 5, 0}, 0},
      {"4d box 36x12x1x9 global desc", 4, {36, 12, 1, 9}, {31, 2, 5, 0}, 1},
      {"4d box 36x12x1x1 grid_constant", 4, {36, 12, 1, 1}, {31, 2, 5, 0}, 0},
      {"4d box 32x12x1x9 grid_constant", 4, {32, 12, 1, 9}, {31, 2, 5, 0}, 0},
  };
  int idx = -1;
  for (auto &cs : cases) {
    idx++;
    if (only >= 0 && idx != only) continue;
    Maps M;
    cuuint64_t dim[4] = {PX, NY, NZ, NC};
    cuuint64_t str[3] = {PX * 4, (cuuint64_t)PX * NY * 4, V * 4};
    cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult r = enc(&M.a, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, cs.rank, d, dim, str, cs.box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    M.b = M.a;
    cudaMemcpy(gmap, &M.a, sizeof(CUtensorMap), cudaMemcpyHostToDevice);
    int nf = cs.box[0] * cs.box[1] * cs.box[2] * cs.box[3];
    if (cs.rank == 4) k_probe<4><<<1, 128, 65536 + 64>>>(M, gmap, cs.use_global, cs.c[0], cs.c[1], cs.c[2], cs.c[3], nf, out);
    else k_probe<2><<<1, 128, 65536 + 64>>>(M, gmap, cs.use_global, cs.c[0], cs.c[1], 0, 0, nf, out);
    cudaError_t e = cudaDeviceSynchronize();
    std::vector<float> o(nf);
    int bad = -1;
    if (e == cudaSuccess) {
      cudaMemcpy(o.data(), out, nf * 4, cudaMemcpyDeviceToHost);
      bad = 0;
      for (int c = 0; c < (int)cs.box[3]; c++) for (int y = 0; y < (int)cs.box[1]; y++) for (int x = 0; x < (int)cs.box[0]; x++) {
        int gx = cs.c[0] + x, gy = cs.c[1] + y, gz = cs.rank == 4 ? cs.c[2] : 0, gc = cs.c[3] + c;
        float ref = (gx < PX && gy < NY) ? h[(size_t)gc * V + ((size_t)gz * NY + gy) * PX + gx] : 0.0f;
        if (o[((size_t)c * cs.box[1] + y) * cs.box[0] + x] != ref) bad++;
      }
    }
    printf("%-36s encode=%d run=%s mismatches=%d\n", cs.name, (int)r, cudaGetErrorString(e), bad);
    if (e != cudaSuccess) { printf("context lost, stopping\n"); return 1; }
  }
  return 0;
}
