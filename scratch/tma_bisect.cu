// Standalone bisect of the TMA/mbarrier sequence used by corr1d_fwd.cu.  variant bits:
//  1: prefetch.tensormap   2: 4-D map (else 2-D)   4: negative start coord   8: box wider than tile
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>
#include "../pmt_learning_for_semantic_segmentation_and_disparity_b200/csrc/common.cuh"
using namespace pmt;
namespace pmt { void set_error(const char*, ...) {} const char* get_error() { return ""; } int sm_count() { return 148; } }

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__global__ void k2d(const __grid_constant__ CUtensorMap tm, float* out, int bw, int rows, int x0, int variant) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  float* s = (float*)smem_raw;
  uint64_t* bar = (uint64_t*)(smem_raw + 65536);
  if (threadIdx.x == 0) { mbar_init(bar, 1); fence_mbar_init(); }
  __syncthreads();
  if (threadIdx.x == 0) {
    if (variant & 1) tma_prefetch_desc(&tm);
    mbar_arrive_expect_tx(bar, bw * rows * 4);
    if (variant & 2) tma_load_4d(s, &tm, x0, 1, 0, 0, bar);
    else asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                      ::"r"(smem_u32(s)), "l"((uint64_t)&tm), "r"(x0), "r"(0), "r"(smem_u32(bar)) : "memory");
  }
  mbar_wait(bar, 0);
  for (int i = threadIdx.x; i < bw * rows; i += blockDim.x) out[i] = s[i];
}

int main(int argc, char** argv) {
  int variant = argc > 1 ? atoi(argv[1]) : 0;
  void* p = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
  EncodeTiledFn enc = (EncodeTiledFn)p;
  const int B = 2, C = 64, H = 8, W = 128;
  std::vector<float> h((size_t)B * C * H * W);
  for (size_t i = 0; i < h.size(); ++i) h[i] = (float)(i % 1000);
  float *d, *o; cudaMalloc(&d, h.size() * 4); cudaMalloc(&o, 65536);
  cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
  CUtensorMap tm;
  int bw = (variant & 8) ? 104 : 64, rows = 16;
  CUresult r;
  if (variant & 2) {
    cuuint64_t dims[4] = {W, H, C, B}; cuuint64_t str[3] = {W * 4, W * H * 4, (cuuint64_t)W * H * C * 4};
    cuuint32_t box[4] = {(cuuint32_t)bw, 1, (cuuint32_t)rows, 1}; cuuint32_t es[4] = {1, 1, 1, 1};
    r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, d, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  } else {
    cuuint64_t dims[2] = {W, (cuuint64_t)B * C * H}; cuuint64_t str[1] = {W * 4};
    cuuint32_t box[2] = {(cuuint32_t)bw, (cuuint32_t)rows}; cuuint32_t es[2] = {1, 1};
    r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, d, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  }
  printf("variant %d encode=%d\n", variant, (int)r);
  cudaFuncSetAttribute(k2d, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536 + 64);
  k2d<<<1, 128, 65536 + 64>>>(tm, o, bw, rows, argc > 2 ? atoi(argv[2]) : ((variant & 4) ? -19 : 0), variant);
  cudaError_t e = cudaDeviceSynchronize();
  std::vector<float> ho(bw * rows);
  cudaMemcpy(ho.data(), o, ho.size() * 4, cudaMemcpyDeviceToHost);
  printf("variant %d sync=%s first=%g %g %g last=%g\n", variant, cudaGetErrorString(e), ho[0], ho[1], ho[bw], ho[bw * rows - 1]);
  return e != cudaSuccess;
}
