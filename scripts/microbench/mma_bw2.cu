// tcgen05.mma kind::tf32 microbenchmark 2: straight-line groups, rotating accumulators.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include "../../pmt_learning_for_semantic_segmentation_and_disparity_b200/csrc/tc_common.cuh"
using namespace pmt;
namespace pmt { void set_error(const char*, ...) {} const char* get_error() { return ""; } int sm_count() { return 148; } }

// group = 8 MMAs of size N; MMA i accumulates into accumulator (i % NACC) at column (i % NACC) * N (must fit 448 cols);
// A from smem (TS=0) or TMEM (TS=1); SAMEK: all MMAs read the same k-slice (else rotate over 4 slices)
template <int N, int NACC, int TS, int SAMEK>
__global__ void __launch_bounds__(128, 1) k(int iters, long long* out) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t bar[2];
  __shared__ uint32_t tslot;
  const int wid = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 48 * 1024; i += 128) reinterpret_cast<float*>(smem)[i] = 1.0f;
  if (threadIdx.x == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); fence_mbar_init(); }
  if (wid == 0) { tc::tmem_alloc(&tslot, 512); tc::tmem_relinquish(); }
  fence_proxy_async();
  tc::fence_before_sync(); __syncthreads(); tc::fence_after_sync();
  const uint32_t tb = tslot;
  if (threadIdx.x == 0) {
    const uint32_t id = tc::make_idesc(2, 0, 0, 128, N);
    const uint64_t dA = tc::smem_desc(smem_u32(smem), 16, 1024, 2);
    const uint64_t dB = tc::smem_desc(smem_u32(smem + 65536), 16, 1024, 2);
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const uint32_t d = tb + (i % NACC) * N;
        const int ks = SAMEK ? 0 : (i & 3);
        if (TS) tc::mma_tf32_ts(d, tb + 448 + 8 * ks, dB + 2 * ks, id, 1u);
        else tc::mma_tf32(d, dA + 2 * ks, dB + 2 * ks, id, 1u);
      }
    }
    const long long t1 = clock64();
    tc::mma_commit(&bar[0]);
    mbar_wait(&bar[0], 0);
    const long long t2 = clock64();
    if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  tc::fence_before_sync(); __syncthreads();
  if (wid == 0) { tc::fence_after_sync(); tc::tmem_dealloc(tb, 512); }
}
template <int N, int NACC, int TS, int SAMEK>
void run(long long* d) {
  const int iters = 1000;
  cudaFuncSetAttribute(k<N, NACC, TS, SAMEK>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  k<N, NACC, TS, SAMEK><<<148, 128, 200 * 1024>>>(iters, d);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
  printf("N=%3d nacc=%d TS=%d samek=%d: %s issue %.1f total %.1f cyc/MMA (ideal %.0f)\n", N, NACC, TS, SAMEK, cudaGetErrorString(e),
         (double)h[0] / iters / 8, (double)h[1] / iters / 8, N / 2.0);
}
int main() {
  long long* d; cudaMalloc(&d, 64);
  run<64, 1, 0, 0>(d); run<64, 2, 0, 0>(d); run<64, 4, 0, 0>(d); run<64, 1, 0, 1>(d);
  run<128, 1, 0, 0>(d); run<128, 2, 0, 0>(d); run<128, 1, 0, 1>(d);
  run<256, 1, 0, 0>(d); run<192, 1, 0, 0>(d); run<192, 2, 0, 0>(d);
  run<64, 1, 1, 0>(d); run<64, 2, 1, 0>(d); run<64, 4, 1, 0>(d); run<64, 1, 1, 1>(d);
  run<128, 1, 1, 0>(d); run<128, 2, 1, 0>(d); run<128, 1, 1, 1>(d);
  run<256, 1, 1, 0>(d); run<192, 1, 1, 0>(d); run<192, 2, 1, 0>(d);
  return 0;
}
