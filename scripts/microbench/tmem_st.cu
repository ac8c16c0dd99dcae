// tcgen05.st throughput/latency microbenchmark
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include "../../pmt_learning_for_semantic_segmentation_and_disparity_b200/csrc/tc_common.cuh"
using namespace pmt;
namespace pmt { void set_error(const char*, ...) {} const char* get_error() { return ""; } int sm_count() { return 148; } }

// MODE 0: st16 hi + st16 lo + wait::st + fence   1: same without wait   2: wait only every 4th   3: st16 x1 + wait
template <int MODE>
__global__ void __launch_bounds__(512, 1) k(int iters, int nwarps, long long* out) {
  __shared__ uint32_t tslot;
  const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (wid == 0) { tc::tmem_alloc(&tslot, 512); tc::tmem_relinquish(); }
  tc::fence_before_sync(); __syncthreads(); tc::fence_after_sync();
  const uint32_t tb = tslot;
  long long t0 = 0, t1 = 0;
  if (wid < nwarps) {
    const int q = wid & 3, sub = (wid >> 2) & 1;
    float w[16];
#pragma unroll
    for (int t = 0; t < 16; ++t) w[t] = (float)(lane + t);
    const uint32_t ta = tb + ((uint32_t)(32 * q) << 16) + 256 + sub * 16 + ((wid >> 3) & 1) * 64;
    t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      tc::tmem_st16(ta, w);
      if (MODE != 3) tc::tmem_st16(ta + 32, w);
      if (MODE == 0 || MODE == 3 || (MODE == 2 && (it & 3) == 3)) tc::tmem_st_wait();
      tc::fence_before_sync();
#pragma unroll
      for (int t = 0; t < 16; ++t) w[t] += 1.f;
    }
    tc::tmem_st_wait();
    t1 = clock64();
  }
  if (threadIdx.x == 0) out[0] = t1 - t0;
  tc::fence_before_sync(); __syncthreads();
  if (wid == 0) { tc::fence_after_sync(); tc::tmem_dealloc(tb, 512); }
}
template <int MODE>
void run(long long* d, int nwarps) {
  const int iters = 2000;
  k<MODE><<<148, 512, 0>>>(iters, nwarps, d);
  cudaError_t e = cudaDeviceSynchronize();
  long long h; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
  printf("mode=%d warps=%d: %s %.1f cyc/iter\n", MODE, nwarps, cudaGetErrorString(e), (double)h / iters);
}
int main() {
  long long* d; cudaMalloc(&d, 64);
  for (int nw : {1, 4, 8, 16}) { run<0>(d, nw); run<1>(d, nw); run<2>(d, nw); run<3>(d, nw); }
  return 0;
}
