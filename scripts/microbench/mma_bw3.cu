// microbenchmark 3: the bwd kernel's exact per-chunk MMA sequences, straight-line.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include "../../pmt_learning_for_semantic_segmentation_and_disparity_b200/csrc/tc_common.cuh"
using namespace pmt;
namespace pmt { void set_error(const char*, ...) {} const char* get_error() { return ""; } int sm_count() { return 148; } }

// SEQ 0: [N128, N64] x4   1: [N128, N128] x4   2: [N64 x3] x4   3: [N192] x4 (reference)   NCOMMIT commits per chunk
template <int SEQ, int TS, int NCOMMIT>
__global__ void __launch_bounds__(128, 1) k(int iters, long long* out) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t bar[4];
  __shared__ uint32_t tslot;
  const int wid = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 48 * 1024; i += 128) reinterpret_cast<float*>(smem)[i] = 1.0f;
  if (threadIdx.x == 0) { for (int i = 0; i < 4; ++i) mbar_init(&bar[i], 1); fence_mbar_init(); }
  if (wid == 0) { tc::tmem_alloc(&tslot, 512); tc::tmem_relinquish(); }
  fence_proxy_async();
  tc::fence_before_sync(); __syncthreads(); tc::fence_after_sync();
  const uint32_t tb = tslot;
  if (threadIdx.x == 0) {
    const uint32_t id64 = tc::make_idesc(2, 0, 0, 128, 64), id128 = tc::make_idesc(2, 0, 0, 128, 128), id192 = tc::make_idesc(2, 0, 0, 128, 192);
    const uint64_t dA = tc::smem_desc(smem_u32(smem), 16, 1024, 2);
    const uint64_t dB = tc::smem_desc(smem_u32(smem + 65536), 16, 1024, 2);
    const uint32_t ta = tb + 384;
    auto mma = [&](uint32_t d, int a_lo, int kk, uint32_t id, int b_off) {
      if (TS) tc::mma_tf32_ts(d, ta + 32 * a_lo + 8 * kk, dB + b_off + 2 * kk, id, 1u);
      else tc::mma_tf32(d, dA + 1024 * a_lo + 2 * kk, dB + b_off + 2 * kk, id, 1u);
    };
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        if (SEQ == 0) { mma(tb, 0, kk, id128, 0); mma(tb, 1, kk, id64, 0); }
        if (SEQ == 1) { mma(tb, 0, kk, id128, 0); mma(tb, 1, kk, id128, 0); }
        if (SEQ == 2) { mma(tb, 0, kk, id64, 0); mma(tb, 1, kk, id64, 0); mma(tb, 0, kk, id64, 512); }
        if (SEQ == 3) { mma(tb, 0, kk, id192, 0); }
        if (SEQ == 4) { mma(tb, 0, kk, id128, 0); mma(tb + 128, 1, kk, id64, 0); }
      }
      if (NCOMMIT >= 1) tc::mma_commit(&bar[1]);
      if (NCOMMIT >= 2) tc::mma_commit(&bar[2]);
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
template <int SEQ, int TS, int NCOMMIT>
void run(long long* d) {
  const int iters = 1000;
  cudaFuncSetAttribute(k<SEQ, TS, NCOMMIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  k<SEQ, TS, NCOMMIT><<<148, 128, 200 * 1024>>>(iters, d);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
  printf("seq=%d TS=%d commits=%d: %s issue %.1f total %.1f cyc/chunk\n", SEQ, TS, NCOMMIT, cudaGetErrorString(e),
         (double)h[0] / iters, (double)h[1] / iters);
}
int main() {
  long long* d; cudaMalloc(&d, 64);
  run<0, 1, 0>(d); run<0, 1, 1>(d); run<0, 1, 2>(d); run<1, 1, 0>(d); run<2, 1, 0>(d); run<3, 1, 0>(d); run<4, 1, 0>(d); run<4, 1, 2>(d);
  run<0, 0, 0>(d); run<0, 0, 2>(d); run<1, 0, 0>(d); run<2, 0, 0>(d); run<3, 0, 0>(d); run<4, 0, 0>(d);
  return 0;
}
