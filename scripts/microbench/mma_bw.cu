// tcgen05.mma kind::tf32 issue/throughput microbenchmark: one CTA per SM, one thread issues `iters` groups of MMAs.
// usage: mma_bw <pattern>   patterns: see table in main
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include "../../pmt_learning_for_semantic_segmentation_and_disparity_b200/csrc/tc_common.cuh"
using namespace pmt;
namespace pmt { void set_error(const char*, ...) {} const char* get_error() { return ""; } int sm_count() { return 148; } }

struct Pat { int n1, n2; int a_tmem; int ksteps; int commit_every; int same_acc; };
// group = ksteps x { MMA(N=n1) [, MMA(N=n2)] } then (optionally) a commit
__global__ void __launch_bounds__(128, 1) k(Pat p, int iters, long long* out) {
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
    const uint32_t id1 = tc::make_idesc(2, 0, 0, 128, p.n1), id2 = tc::make_idesc(2, 0, 0, 128, p.n2 ? p.n2 : 64);
    const uint64_t dA = tc::smem_desc(smem_u32(smem), 16, 1024, 2);
    const uint64_t dB = tc::smem_desc(smem_u32(smem + 32768), 16, 1024, 2);
    uint32_t ph = 0;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      for (int kk = 0; kk < p.ksteps; ++kk) {
        const uint32_t d1 = tb, d2 = p.same_acc ? tb : tb + 256;
        if (p.a_tmem) {
          tc::mma_tf32_ts(d1, tb + 448 + 8 * (kk & 3), dB + 2 * (kk & 3), id1, 1u);
          if (p.n2) tc::mma_tf32_ts(d2, tb + 480 + 8 * (kk & 3), dB + 2 * (kk & 3), id2, 1u);
        } else {
          tc::mma_tf32(d1, dA + 2 * (kk & 3), dB + 2 * (kk & 3), id1, 1u);
          if (p.n2) tc::mma_tf32(d2, dA + 1024 + 2 * (kk & 3), dB + 2 * (kk & 3), id2, 1u);
        }
      }
      if (p.commit_every) tc::mma_commit(&bar[1]);
    }
    const long long t1 = clock64();
    tc::mma_commit(&bar[0]);
    mbar_wait(&bar[0], ph);
    const long long t2 = clock64();
    if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  tc::fence_before_sync(); __syncthreads();
  if (wid == 0) { tc::fence_after_sync(); tc::tmem_dealloc(tb, 512); }
}

int main() {
  long long* d; cudaMalloc(&d, 64);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  Pat pats[] = {
      {64, 0, 0, 4, 0, 1},  {128, 0, 0, 4, 0, 1}, {192, 0, 0, 4, 0, 1}, {256, 0, 0, 4, 0, 1},
      {128, 64, 0, 4, 0, 1}, {128, 64, 0, 4, 1, 1}, {128, 64, 0, 4, 0, 0}, {128, 64, 1, 4, 0, 1}, {128, 64, 1, 4, 1, 1},
      {128, 64, 1, 4, 0, 0}, {64, 0, 1, 4, 0, 1}, {128, 0, 1, 4, 0, 1}, {256, 0, 1, 4, 0, 1}, {160, 160, 0, 4, 0, 0}, {160, 160, 0, 4, 1, 0},
  };
  const int iters = 2000;
  for (auto& p : pats) {
    k<<<148, 128, 200 * 1024>>>(p, iters, d);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    const int nm = p.ksteps * (p.n2 ? 2 : 1);
    const double ideal = p.ksteps * (p.n1 / 2.0 + p.n2 / 2.0);
    printf("N1=%3d N2=%3d a_tmem=%d commit=%d same_acc=%d: %s issue %.1f cyc/group, total %.1f cyc/group (%.1f per MMA), ideal %.0f\n", p.n1, p.n2,
           p.a_tmem, p.commit_every, p.same_acc, cudaGetErrorString(e), (double)h[0] / iters, (double)h[1] / iters, (double)h[1] / iters / nm, ideal);
  }
  return 0;
}
