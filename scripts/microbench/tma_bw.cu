// TMA load-throughput microbenchmark: persistent CTAs, one producer thread + one consumer thread, S-slot ring.
// usage: tma_bw <pattern> <slots> <H> [ctas]
//  pattern 0: raw0  6 boxes [32 planes][128 w] per tile of a (4,192,H,512) tensor
//  pattern 1: raw1 10 boxes [160 planes][32 w] per tile
//  pattern 2: band 10 boxes [64 planes][32 w] per tile of a (4,64,H,512) tensor (128B swizzle)
//  pattern 3: raw0 + band (two producers), pattern 4: raw1 + band
//  pattern 5: band as 5 boxes [64][64 w] no swizzle;  pattern 6: raw as 3 boxes [64 planes][128 w]
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
static CUtensorMap mk(float* d, int B, int C, int H, int W, int bw, int bc, int sw) {
  CUtensorMap tm;
  cuuint64_t dims[4] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)C, (cuuint64_t)B};
  cuuint64_t str[3] = {(cuuint64_t)W * 4, (cuuint64_t)W * H * 4, (cuuint64_t)W * H * C * 4};
  cuuint32_t box[4] = {(cuuint32_t)bw, 1, (cuuint32_t)bc, 1}; cuuint32_t es[4] = {1, 1, 1, 1};
  CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, d, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
            sw ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r) { printf("encode failed %d\n", (int)r); exit(1); }
  return tm;
}
struct Stream { int nbox, bw, bc, slots, dx, dp, x_off, p_off; };   // box k of a tile at (x0 + x_off + dx*k, h, dp*k, n)

__global__ void __launch_bounds__(128, 1) k(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                                            Stream sa, Stream sb, int n_tiles, int n_xt, int H, long long* cyc) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t full[2][16], empty[2][16];
  if (threadIdx.x == 0) {
    for (int i = 0; i < 16; ++i) { mbar_init(&full[0][i], 1); mbar_init(&empty[0][i], 1); mbar_init(&full[1][i], 1); mbar_init(&empty[1][i], 1); }
    fence_mbar_init();
  }
  __syncthreads();
  const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long t0 = clock64();
  if (lane == 0) {
    const int which = wid & 1;
    const Stream s = which ? sb : sa;
    const CUtensorMap* tm = which ? &tmB : &tmA;
    unsigned char* base = smem + (which ? sa.slots * sa.bw * sa.bc * 4 : 0);
    const int bytes = s.bw * s.bc * 4;
    if (s.nbox > 0) {
      int slot = 0; uint32_t ph = 0;
      if (wid < 2) {  // producer
        for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
          const int x0 = (t % n_xt) * 128, h = (t / n_xt) % H, n = t / (n_xt * H);
          for (int kk = 0; kk < s.nbox; ++kk) {
            mbar_wait(&empty[which][slot], ph ^ 1u);
            mbar_arrive_expect_tx(&full[which][slot], bytes);
            tma_load_4d(base + slot * bytes, tm, x0 + s.x_off + s.dx * kk, h, s.p_off + s.dp * kk, n, &full[which][slot]);
            if (++slot == s.slots) slot = 0, ph ^= 1u;
          }
        }
      } else {        // consumer
        for (int t = blockIdx.x; t < n_tiles; t += gridDim.x)
          for (int kk = 0; kk < s.nbox; ++kk) {
            mbar_wait(&full[which][slot], ph);
            mbar_arrive(&empty[which][slot]);
            if (++slot == s.slots) slot = 0, ph ^= 1u;
          }
      }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = clock64() - t0;
}

int main(int argc, char** argv) {
  int pattern = argc > 1 ? atoi(argv[1]) : 0, slots = argc > 2 ? atoi(argv[2]) : 4, H = argc > 3 ? atoi(argv[3]) : 256;
  int ctas = argc > 4 ? atoi(argv[4]) : 148;
  void* p = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
  enc = (EncodeTiledFn)p;
  const int B = 4, W = 512, P = 192, C = 64;
  float *g, *f; long long* cyc;
  cudaMalloc(&g, (size_t)B * P * H * W * 4); cudaMalloc(&f, (size_t)B * C * H * W * 4); cudaMalloc(&cyc, 64);
  cudaMemset(g, 0, (size_t)B * P * H * W * 4); cudaMemset(f, 0, (size_t)B * C * H * W * 4);
  Stream raw0{6, 128, 32, slots, 0, 32, 0, 0}, raw1{10, 32, 160, slots, 32, -32, -96, 160}, band{10, 32, 64, slots, 32, 0, -96, 0};
  Stream band64{5, 64, 64, slots, 64, 0, -96, 0}, raw64{3, 128, 64, slots, 0, 64, 0, 0}, none{0, 32, 32, 1, 0, 0, 0, 0};
  raw1.dp = -32;  // p0 = P-1+delta-32k-31 : approximate with 160 - 32k start (may go negative -> zero fill)
  Stream sa = none, sb = none; CUtensorMap tA, tB;
  tB = mk(f, B, C, H, W, 32, 64, 1);
  switch (pattern) {
    case 0: sa = raw0; tA = mk(g, B, P, H, W, 128, 32, 0); break;
    case 1: sa = raw1; tA = mk(g, B, P, H, W, 32, 160, 0); break;
    case 2: sa = band; tA = mk(f, B, C, H, W, 32, 64, 1); break;
    case 3: sa = raw0; tA = mk(g, B, P, H, W, 128, 32, 0); sb = band; break;
    case 4: sa = raw1; tA = mk(g, B, P, H, W, 32, 160, 0); sb = band; break;
    case 5: sa = band64; tA = mk(f, B, C, H, W, 64, 64, 0); break;
    case 6: sa = raw64; tA = mk(g, B, P, H, W, 128, 64, 0); break;
  }
  const int n_xt = W / 128, n_tiles = B * H * n_xt;
  const size_t smem = (size_t)sa.slots * sa.bw * sa.bc * 4 + (size_t)(sb.nbox ? sb.slots * sb.bw * sb.bc * 4 : 0) + 1024;
  if (smem > 227 * 1024) { printf("pattern %d slots %d: smem %zu too large\n", pattern, slots, smem); return 0; }
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e9;
  for (int it = 0; it < 5; ++it) {
    cudaEventRecord(e0);
    k<<<ctas, 128, smem>>>(tA, tB, sa, sb, n_tiles, n_xt, H, cyc);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
  }
  cudaError_t e = cudaDeviceSynchronize();
  long long hc; cudaMemcpy(&hc, cyc, 8, cudaMemcpyDeviceToHost);
  const double bytes = (double)n_tiles * ((double)sa.nbox * sa.bw * sa.bc * 4 + (double)sb.nbox * sb.bw * sb.bc * 4);
  printf("pattern %d slots %d H %d ctas %d: %s  %.1f us  %.2f GB  %.0f GB/s  %.1f B/cyc/SM (cta0 %lld cyc)\n", pattern, slots, H, ctas,
         cudaGetErrorString(e), best * 1e3, bytes / 1e9, bytes / best / 1e6, bytes / ctas / (double)hc, hc);
  return 0;
}
