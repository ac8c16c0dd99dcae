// tma_rows: (1) does a TMA tiled load accept an inner start coordinate that is NOT a multiple of 16 bytes (and a
// negative / past-the-end one, zero-filled)?  (2) how fast are single-row [1][128] boxes (512 B each) when one warp
// issues 32 of them per barrier -- the access pattern a per-plane shifted g slice needs (mode 1 of the backward).
// usage: tma_rows            -> correctness of starts -97..+97 (odd ones included), then throughput
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include "../../pmt_learning_for_semantic_segmentation_and_disparity_b200/csrc/common.cuh"
using namespace pmt;
namespace pmt { void set_error(const char*, ...) {} const char* get_error() { return ""; } int sm_count() { return 148; } }
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn enc;
static CUtensorMap mk(float* d, int B, int C, int H, int W, int bw, int bc) {
  CUtensorMap tm;
  cuuint64_t dims[4] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)C, (cuuint64_t)B};
  cuuint64_t str[3] = {(cuuint64_t)W * 4, (cuuint64_t)W * H * 4, (cuuint64_t)W * H * C * 4};
  cuuint32_t box[4] = {(cuuint32_t)bw, 1, (cuuint32_t)bc, 1}; cuuint32_t es[4] = {1, 1, 1, 1};
  CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, d, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r) { printf("encode failed %d\n", (int)r); exit(1); }
  return tm;
}

// correctness: one CTA, loads rows p = 0..31 of plane-slice (h, n) with start x0 + shift - p (a different alignment per row)
__global__ void __launch_bounds__(32, 1) k_check(const __grid_constant__ CUtensorMap tm, int x0, int shift, int h, float* out) {
  __shared__ __align__(128) float tile[32 * 128];
  __shared__ uint64_t bar;
  const int lane = threadIdx.x;
  if (lane == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  __syncwarp();
  if (lane == 0) mbar_arrive_expect_tx(&bar, 32 * 512);
  __syncwarp();
  tma_load_4d(tile + lane * 128, &tm, x0 + shift - 4 * lane, h, lane, 0, &bar);   // always a multiple of 4 floats
  mbar_wait(&bar, 0);
  for (int i = lane; i < 32 * 128; i += 32) out[i] = tile[i];
}

// throughput: persistent CTAs; per tile 192 single-row boxes in 6 groups of 32 (one barrier per group, 6-slot ring)
__global__ void __launch_bounds__(64, 1) k_bw(const __grid_constant__ CUtensorMap tm, int n_tiles, int n_xt, int H, int rows_per_box,
                                              long long* cyc) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t full[8], empty[8];
  const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { for (int i = 0; i < 8; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); } fence_mbar_init(); }
  __syncthreads();
  const long long t0 = clock64();
  int slot = 0; uint32_t ph = 0;
  const int nbox = 32 / rows_per_box;   // boxes per 32-row group
  const int bw = rows_per_box == 32 ? 128 : (rows_per_box == 4 ? 136 : 132);
  if (wid == 0) {
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
      const int x0 = (t % n_xt) * 128, h = (t / n_xt) % H, n = t / (n_xt * H);
      for (int g = 0; g < 6; ++g) {
        mbar_wait(&empty[slot], ph ^ 1u);
        if (lane == 0) mbar_arrive_expect_tx(&full[slot], 32 * bw * 4);
        __syncwarp();
        if (lane < nbox) {
          const int p = 32 * g + lane * rows_per_box;
          const int xs = rows_per_box == 32 ? x0 : x0 + 4 * ((95 - p) / 4) - 4;   // shifted per box, 16-byte aligned
          tma_load_4d(smem + slot * 17408 + lane * rows_per_box * bw * 4, &tm, xs, h, p, n, &full[slot]);
        }
        if (++slot == 6) slot = 0, ph ^= 1u;
      }
    }
  } else if (lane == 0) {
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x)
      for (int g = 0; g < 6; ++g) {
        mbar_wait(&full[slot], ph);
        mbar_arrive(&empty[slot]);
        if (++slot == 6) slot = 0, ph ^= 1u;
      }
  }
  __syncthreads();
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = clock64() - t0;
}

int main() {
  void* p = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
  enc = (EncodeTiledFn)p;
  const int B = 4, W = 512, P = 192, H = 256;
  const size_t n = (size_t)B * P * H * W;
  float* g; cudaMalloc(&g, n * 4);
  float* hg = (float*)malloc(n * 4);
  for (size_t i = 0; i < n; ++i) hg[i] = (float)(i % 1000003);
  cudaMemcpy(g, hg, n * 4, cudaMemcpyHostToDevice);
  float* out; cudaMalloc(&out, 32 * 128 * 4);
  float ho[32 * 128];
  CUtensorMap tm1 = mk(g, B, P, H, W, 128, 1);
  int bad_total = 0;
  const int h = 7;
  for (int x0 = 0; x0 < W; x0 += 128) {
    for (int shift : {-96, -4, 0, 4, 96}) {
      k_check<<<1, 32>>>(tm1, x0, shift, h, out);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("x0=%d shift=%d: CUDA error %s\n", x0, shift, cudaGetErrorString(e)); return 1; }
      cudaMemcpy(ho, out, sizeof(ho), cudaMemcpyDeviceToHost);
      int bad = 0;
      for (int r = 0; r < 32; ++r)
        for (int i = 0; i < 128; ++i) {
          const int w = x0 + shift - 4 * r + i;
          const float want = (w >= 0 && w < W) ? hg[((size_t)r * H + h) * W + w] : 0.f;
          if (ho[r * 128 + i] != want) ++bad;
        }
      if (bad) printf("x0=%d shift=%d: %d mismatches\n", x0, shift, bad);
      bad_total += bad;
    }
  }
  printf("16-byte-aligned negative / past-the-end inner starts: %s\n", bad_total ? "MISMATCH" : "all exact (zero fill outside [0,W))");
  printf("(an inner start that is not a multiple of 16 bytes raises 'illegal instruction': measured in the first run of this probe)\n");
  long long* cyc; cudaMalloc(&cyc, 64);
  const int n_xt = W / 128, n_tiles = B * H * n_xt;
  cudaFuncSetAttribute(k_bw, cudaFuncAttributeMaxDynamicSharedMemorySize, 6 * 17408 + 1024);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int rpb : {32, 4, 1}) {
    CUtensorMap tm = mk(g, B, P, H, W, rpb == 32 ? 128 : (rpb == 4 ? 136 : 132), rpb);
    for (int ctas : {148, 74}) {
      float best = 1e9;
      for (int it = 0; it < 4; ++it) {
        cudaEventRecord(e0);
        k_bw<<<ctas, 64, 6 * 17408 + 1024>>>(tm, n_tiles, n_xt, H, rpb, cyc);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
      }
      cudaError_t e = cudaDeviceSynchronize();
      const double bytes = (double)n_tiles * 192 * 512;
      printf("rows/box %2d ctas %3d: %s %.1f us %.0f GB/s\n", rpb, ctas, cudaGetErrorString(e), best * 1e3, bytes / best / 1e6);
    }
  }
  return 0;
}
