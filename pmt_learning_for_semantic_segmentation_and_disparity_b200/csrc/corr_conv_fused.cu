// corr_conv_fused.cu -- f2: the 1 x P correlation fused with its epilogue
//   y = squeeze(correlation_sampler(a, b), 1);  y = corrConv2d(y)      (models/dsnet_t2.py:1187-1197, :879-888;
//   corrConv2d = conv2dSame(P, 128, 1, padding='same') [no bias] + ReLU   models/dsnet_t2_warp.py:664-671; dsnet_t2.py:852)
// i.e.  z[n,o,h,w] = relu( sum_p Wt[o,p] * sum_c a[n,c,h,w] * b[n,c,h,w+s_p] ),  s_p = p - (P-1)/2.
//
// This is the one correlation call the training configurations make (C=352 features at 1/8 resolution, 32 x 64, P=17):
// 17.6 MB of traffic, 0.07 GFLOP -- the launch/latency regime, where the tensor-core kernels pay their fixed pipeline
// cost (forward 19 us, backward 52 us) and the (B,17,H,W) slab then makes a round trip through HBM for a 17 -> 128
// matrix-vector product per pixel.  Here ONE CTA owns one image row (n,h): the feature rows stream through shared
// memory in 32-channel chunks (cp.async, double-buffered, zero halo = the sampler's border rule), the 17 x W slab lives
// in shared memory, the 1x1 convolution + ReLU run on it in place, and only z (and the small slab, saved for the
// backward) are written.  fp32 FFMA throughout (fixed summation order: deterministic).
//
// Backward (one CTA per row as well):  gzm = gz * [z > 0];  gcorr = Wt^T gzm;  per-row partial of gW = gzm corr^T
// (summed over rows in fixed order by a second small kernel);  then both input gradients are 17-tap filters along w
// with per-column coefficients, evaluated channel chunk by channel chunk:
//   ga[c,w]  = sum_p gcorr[p,w]      * b[c,w+s_p]          gb[c,w'] = sum_p gcorr[p,w'-s_p] * a[c,w'-s_p]
#include "common.cuh"

namespace pmt {
namespace {

constexpr int kFThreads = 256;
constexpr int kFChunk = 32;     // channels per shared-memory stage
constexpr int kFStages = 4;     // max cp.async ring depth: three chunks in flight while one is consumed (with two stages the row
                                // loop ran at one global-memory latency per chunk: 20 us for 11 chunks)
constexpr int kFMaxW = 128;

__device__ __forceinline__ void cp_async_wait_dyn(int pending) {   // pending = groups allowed to stay in flight
  if (pending >= 3) cp_async_wait<3>();
  else if (pending == 2) cp_async_wait<2>();
  else if (pending == 1) cp_async_wait<1>();
  else cp_async_wait<0>();
}

struct FusedArgs {
  const float* a;      // in1 (B,C,H,W)
  const float* b;      // in2 (B,C,H,W)
  const float* wt;     // (O,P) 1x1 convolution weight
  float* z;            // (B,O,H,W)
  float* corr;         // (B,P,H,W) saved slab
  const float* gz;     // backward
  float* ga;
  float* gb;
  float* gw_part;      // (B*H, O*P) per-row partials
  int B, C, H, W, O;
  int pad;             // zero halo on each side of a staged row (>= max |s_p|, multiple of 4)
  int Wp;              // W + 2*pad
  int G;               // channel groups per chunk = kFThreads / (W/2), rounded down (threads beyond G*W/2 idle in the filters)
  float scale;         // 1 (1-D patch: not divided by C)
  int stages;          // cp.async ring depth (2..kFStages), as many as shared memory allows
};

// stage one 32-channel chunk of rows a[n,c0..,h,:] and b[...] into [kFChunk][Wp] tiles (interior at column pad)
__device__ __forceinline__ void stage_chunk(const FusedArgs& f, float* As, float* Bs, int n, int h, int c0, int tid) {
  const int W4 = f.W / 4;
  const int total = kFChunk * W4;
  for (int i = tid; i < 2 * total; i += kFThreads) {
    const bool second = i >= total;
    const int j = second ? i - total : i;
    const int c = j / W4, q = j % W4;
    const bool valid = c0 + c < f.C;
    const float* src = (second ? f.b : f.a) + (((int64_t)n * f.C + (valid ? c0 + c : 0)) * f.H + h) * (int64_t)f.W + 4 * q;
    float* dst = (second ? Bs : As) + c * f.Wp + f.pad + 4 * q;
    cp_async16(dst, src, valid);
  }
}

template <int kP>
__global__ void __launch_bounds__(kFThreads) corr_conv_relu_fwd_kernel(const FusedArgs f) {
  extern __shared__ __align__(16) float sm[];
  constexpr int r = (kP - 1) / 2;
  const int W = f.W, Wp = f.Wp, pairs = W / 2, G = f.G, S = f.stages;
  float* As = sm;                                  // [S][kFChunk][Wp]
  float* Bs = As + S * kFChunk * Wp;               // [S][kFChunk][Wp]
  float* part = Bs + S * kFChunk * Wp;             // [G][kP][W]
  float* cs = part + G * kP * W;                   // [kP][W]
  float* ws = cs + kP * W;                         // [O][kP]
  const int tid = threadIdx.x;
  const int pair = tid % pairs, grp = tid / pairs;
  const bool worker = grp < G;
  const int n_chunks = (f.C + kFChunk - 1) / kFChunk;
  for (int i = tid; i < 2 * S * kFChunk * Wp; i += kFThreads) As[i] = 0.f;   // zero halos (never overwritten)
  for (int i = tid; i < f.O * kP; i += kFThreads) ws[i] = __ldg(f.wt + i);
  __syncthreads();
  for (int row = blockIdx.x; row < f.B * f.H; row += gridDim.x) {
    const int n = row / f.H, h = row % f.H;
    float acc0[kP], acc1[kP];
#pragma unroll
    for (int p = 0; p < kP; ++p) acc0[p] = 0.f, acc1[p] = 0.f;
    for (int k = 0; k < S - 1; ++k) {               // prologue: S-1 chunks in flight (one commit group each)
      if (k < n_chunks) stage_chunk(f, As + k * kFChunk * Wp, Bs + k * kFChunk * Wp, n, h, k * kFChunk, tid);
      cp_async_commit();
    }
    for (int k = 0; k < n_chunks; ++k) {
      const int s = k % S;
      {
        const int kn = k + S - 1, sn = kn % S;      // refill the stage consumed in the previous iteration
        if (kn < n_chunks) stage_chunk(f, As + sn * kFChunk * Wp, Bs + sn * kFChunk * Wp, n, h, kn * kFChunk, tid);
        cp_async_commit();
        cp_async_wait_dyn(S - 1);                   // chunk k has landed
      }
      __syncthreads();
      if (worker) {
        const float* Ac = As + s * kFChunk * Wp + f.pad + 2 * pair;
        const float* Bc = Bs + s * kFChunk * Wp + f.pad - r + 2 * pair;
        for (int c = grp; c < kFChunk; c += G) {
          const float2 l = *reinterpret_cast<const float2*>(Ac + c * Wp);
          float win[kP + 1];
#pragma unroll
          for (int j = 0; j < kP + 1; j += 2) {
            const float2 v = *reinterpret_cast<const float2*>(Bc + c * Wp + j);
            win[j] = v.x;
            if (j + 1 < kP + 1) win[j + 1] = v.y;
          }
#pragma unroll
          for (int p = 0; p < kP; ++p) {
            acc0[p] = fmaf(l.x, win[p], acc0[p]);
            acc1[p] = fmaf(l.y, win[p + 1], acc1[p]);
          }
        }
      }
      __syncthreads();   // the other stage is refilled next iteration
    }
    if (worker) {
#pragma unroll
      for (int p = 0; p < kP; ++p)
        *reinterpret_cast<float2*>(part + (grp * kP + p) * W + 2 * pair) = make_float2(acc0[p], acc1[p]);
    }
    __syncthreads();
    for (int i = tid; i < kP * W; i += kFThreads) {
      float sacc = 0.f;
      for (int g = 0; g < G; ++g) sacc += part[g * kP * W + i];   // fixed order
      sacc *= f.scale;
      cs[i] = sacc;
      const int p = i / W, w = i % W;
      f.corr[(((int64_t)n * kP + p) * f.H + h) * W + w] = sacc;
    }
    __syncthreads();
    // 1x1 convolution + ReLU on the slab: thread = (output group, column), weights broadcast from shared memory
    {
      const int n_og = kFThreads / W > 0 ? kFThreads / W : 1;
      const int w = tid % W, og = tid / W;
      if (og < n_og) {
        float cr[kP];
#pragma unroll
        for (int p = 0; p < kP; ++p) cr[p] = cs[p * W + w];
        for (int o = og; o < f.O; o += n_og) {
          float sacc = 0.f;
#pragma unroll
          for (int p = 0; p < kP; ++p) sacc = fmaf(ws[o * kP + p], cr[p], sacc);
          st_cs(f.z + (((int64_t)n * f.O + o) * f.H + h) * W + w, fmaxf(sacc, 0.f));
        }
      }
    }
    __syncthreads();
  }
}

template <int kP>
__global__ void __launch_bounds__(kFThreads) corr_conv_relu_bwd_kernel(const FusedArgs f) {
  extern __shared__ __align__(16) float sm[];
  constexpr int r = (kP - 1) / 2;
  constexpr int kPq = (kP + 3) & ~3;               // padded P for float4 rows
  const int W = f.W, Wp = f.Wp, pairs = W / 2, G = f.G, O = f.O, S = f.stages;
  float* As = sm;                                  // [S][kFChunk][Wp]
  float* Bs = As + S * kFChunk * Wp;               // [S][kFChunk][Wp]
  float* gzm = Bs + S * kFChunk * Wp;              // [O][W+1]   masked upstream gradient
  float* ct = gzm + O * (W + 1);                   // [W][kPq]   saved slab, transposed
  float* ws = ct + W * kPq;                        // [O][kPq]
  float* gc = ws + O * kPq;                        // [kP][W]    gradient of the slab
  float* part = gc + kP * W;                       // [n_og][kP][W]
  const int tid = threadIdx.x;
  const int pair = tid % pairs, grp = tid / pairs;
  const bool worker = grp < G;
  const int n_chunks = (f.C + kFChunk - 1) / kFChunk;
  const int n_og = kFThreads / W > 0 ? kFThreads / W : 1;
  for (int i = tid; i < 2 * S * kFChunk * Wp; i += kFThreads) As[i] = 0.f;   // zero halos
  for (int i = tid; i < O * kPq; i += kFThreads) ws[i] = (i % kPq) < kP ? __ldg(f.wt + (i / kPq) * kP + i % kPq) : 0.f;
  __syncthreads();
  for (int row = blockIdx.x; row < f.B * f.H; row += gridDim.x) {
    const int n = row / f.H, h = row % f.H;
    for (int k = 0; k < S - 1; ++k) {               // S-1 chunks in flight: they overlap with the slab work below
      if (k < n_chunks) stage_chunk(f, As + k * kFChunk * Wp, Bs + k * kFChunk * Wp, n, h, k * kFChunk, tid);
      cp_async_commit();
    }
    // ---- a. masked upstream gradient and the saved slab ----
    for (int i = tid; i < O * W; i += kFThreads) {
      const int o = i / W, w = i % W;
      const int64_t gi = (((int64_t)n * O + o) * f.H + h) * W + w;
      gzm[o * (W + 1) + w] = __ldg(f.z + gi) > 0.f ? __ldg(f.gz + gi) : 0.f;
    }
    for (int i = tid; i < kP * W; i += kFThreads) {
      const int p = i / W, w = i % W;
      ct[w * kPq + p] = __ldg(f.corr + (((int64_t)n * kP + p) * f.H + h) * W + w);
    }
    __syncthreads();
    // ---- b. gcorr[p][w] = sum_o Wt[o][p] * gzm[o][w]: thread = (output group, column), partials over groups ----
    {
      const int w = tid % W, og = tid / W;
      if (og < n_og) {
        float acc[kP];
#pragma unroll
        for (int p = 0; p < kP; ++p) acc[p] = 0.f;
        for (int o = og; o < O; o += n_og) {
          const float g = gzm[o * (W + 1) + w];
#pragma unroll
          for (int p = 0; p < kP; ++p) acc[p] = fmaf(ws[o * kPq + p], g, acc[p]);
        }
#pragma unroll
        for (int p = 0; p < kP; ++p) part[(og * kP + p) * W + w] = acc[p];
      }
    }
    // ---- c. per-row partial of gW[o][p] = sum_w gzm[o][w] * corr[p][w]  (thread = output channel) ----
    for (int o = tid; o < O; o += kFThreads) {
      float acc[kP];
#pragma unroll
      for (int p = 0; p < kP; ++p) acc[p] = 0.f;
      for (int w = 0; w < W; ++w) {
        const float g = gzm[o * (W + 1) + w];      // pitch W+1: lanes (different o) hit different banks
#pragma unroll
        for (int p = 0; p < kP; ++p) acc[p] = fmaf(g, ct[w * kPq + p], acc[p]);
      }
      float* dst = f.gw_part + ((int64_t)row * O + o) * kP;
#pragma unroll
      for (int p = 0; p < kP; ++p) dst[p] = acc[p];
    }
    __syncthreads();
    for (int i = tid; i < kP * W; i += kFThreads) {
      float sacc = 0.f;
      for (int g = 0; g < n_og; ++g) sacc += part[g * kP * W + i];
      gc[i] = sacc * f.scale;
    }
    __syncthreads();
    // ---- d. both input gradients: 17-tap filters with per-column coefficients, two columns per thread ----
    float ca0[kP], ca1[kP], cb0[kP], cb1[kP];
    if (worker) {
      const int w0 = 2 * pair;
#pragma unroll
      for (int p = 0; p < kP; ++p) {
        ca0[p] = gc[p * W + w0];
        ca1[p] = gc[p * W + w0 + 1];
        const int u0 = w0 - (p - r), u1 = w0 + 1 - (p - r);       // source column w' - s_p
        cb0[p] = (u0 >= 0 && u0 < W) ? gc[p * W + u0] : 0.f;
        cb1[p] = (u1 >= 0 && u1 < W) ? gc[p * W + u1] : 0.f;
      }
    }
    for (int k = 0; k < n_chunks; ++k) {
      const int s = k % S;
      {
        const int kn = k + S - 1, sn = kn % S;
        if (kn < n_chunks) stage_chunk(f, As + sn * kFChunk * Wp, Bs + sn * kFChunk * Wp, n, h, kn * kFChunk, tid);
        cp_async_commit();
        cp_async_wait_dyn(S - 1);
      }
      __syncthreads();
      if (worker) {
        const int w0 = 2 * pair;
        for (int c = grp; c < kFChunk; c += G) {
          const int cg = k * kFChunk + c;
          if (cg >= f.C) break;
          // ga[c][w0 + j] = sum_p ca_j[p] * b[c][w0 + j + p - r]: window b[c][w0 - r .. w0 + 1 + (kP-1-r)]
          const float* Bw = Bs + s * kFChunk * Wp + c * Wp + f.pad - r + w0;
          // gb[c][w0 + j] = sum_p cb_j[p] * a[c][w0 + j - p + r]: window a[c][w0 + r - (kP-1) .. w0 + 1 + r]
          const float* Aw = As + s * kFChunk * Wp + c * Wp + f.pad + r - (kP - 1) + w0;
          float wb[kP + 1], wa[kP + 1];
#pragma unroll
          for (int j = 0; j < kP + 1; ++j) wb[j] = Bw[j], wa[j] = Aw[j];
          float ga0 = 0.f, ga1 = 0.f, gb0 = 0.f, gb1 = 0.f;
#pragma unroll
          for (int p = 0; p < kP; ++p) {
            ga0 = fmaf(ca0[p], wb[p], ga0);
            ga1 = fmaf(ca1[p], wb[p + 1], ga1);
            gb0 = fmaf(cb0[p], wa[kP - 1 - p], gb0);      // a[c][w0 - p + r]
            gb1 = fmaf(cb1[p], wa[kP - p], gb1);          // a[c][w0 + 1 - p + r]
          }
          const int64_t oi = (((int64_t)n * f.C + cg) * f.H + h) * W + w0;
          *reinterpret_cast<float2*>(f.ga + oi) = make_float2(ga0, ga1);
          *reinterpret_cast<float2*>(f.gb + oi) = make_float2(gb0, gb1);
        }
      }
      __syncthreads();
    }
  }
}

// gW[o][p] = sum over rows of the per-row partials, in row order (deterministic)
__global__ void __launch_bounds__(256) corr_conv_gw_reduce_kernel(const float* __restrict__ part, float* __restrict__ gw, int rows,
                                                                  int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float s = 0.f;
  for (int r = 0; r < rows; ++r) s += part[(int64_t)r * n + i];
  gw[i] = s;
}

size_t fused_bwd_smem(int W, int O, int stages) {
  const int Wp = W + 16, n_og = kFThreads / W > 0 ? kFThreads / W : 1;
  return sizeof(float) * ((size_t)2 * stages * kFChunk * Wp + (size_t)O * (W + 1) + (size_t)W * 20 + (size_t)O * 20 + 17 * W +
                          (size_t)n_og * 17 * W);
}
size_t fused_fwd_smem(int W, int O, int stages) {
  const int Wp = W + 16, G = kFThreads / (W / 2);
  return sizeof(float) * ((size_t)2 * stages * kFChunk * Wp + (size_t)G * 17 * W + 17 * W + (size_t)O * 17);
}
int pick_stages(size_t (*need)(int, int, int), int W, int O) {
  for (int s = kFStages; s >= 2; --s)
    if (need(W, O, s) <= 220 * 1024) return s;
  return 0;
}

bool fused_shape_ok(int C, int H, int W, int P, int O) {
  // the backward keeps the masked (O, W) gradient of the row in shared memory next to at least two staging buffers
  return P == 17 && C >= 1 && H >= 1 && W >= 16 && W <= kFMaxW && W % 4 == 0 && O >= 1 && O <= 256 &&
         pick_stages(fused_bwd_smem, W, O) >= 2 && pick_stages(fused_fwd_smem, W, O) >= 2;
}

void fill(FusedArgs* f, int B, int C, int H, int W, int O) {
  f->B = B, f->C = C, f->H = H, f->W = W, f->O = O;
  f->pad = 8;                         // >= max |s_p| = 8 for P = 17, multiple of 4 (16-byte aligned interior)
  f->Wp = W + 2 * f->pad;
  f->G = kFThreads / (W / 2);
  f->scale = 1.f;
}

}  // namespace

int corr_conv_relu_supported(int C, int H, int W, int P, int O) { return fused_shape_ok(C, H, W, P, O) ? 1 : 0; }

int launch_corr_conv_relu_fwd(const float* a, const float* b, const float* wt, float* z, float* corr, int B, int C, int H,
                              int W, int P, int O, cudaStream_t st) {
  PMT_CHECK_ARG(fused_shape_ok(C, H, W, P, O), "corr+conv+relu: unsupported shape (needs P=17, W%%4==0, W<=128, O<=256)");
  if ((int64_t)B * H == 0) return PMT_OK;
  FusedArgs f{};
  f.a = a, f.b = b, f.wt = wt, f.z = z, f.corr = corr;
  fill(&f, B, C, H, W, O);
  f.stages = pick_stages(fused_fwd_smem, W, O);
  const size_t smem = fused_fwd_smem(W, O, f.stages);
  PMT_CUDA_OK(cudaFuncSetAttribute(corr_conv_relu_fwd_kernel<17>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int rows = B * H;
  const int grid = rows < 2 * sm_count() ? rows : 2 * sm_count();
  corr_conv_relu_fwd_kernel<17><<<grid, kFThreads, smem, st>>>(f);
  PMT_LAUNCH_OK("corr_conv_relu_fwd_kernel");
  return PMT_OK;
}

int launch_corr_conv_relu_bwd(const float* a, const float* b, const float* wt, const float* z, const float* corr,
                              const float* gz, float* ga, float* gb, float* gw, float* gw_part, int B, int C, int H, int W,
                              int P, int O, cudaStream_t st) {
  PMT_CHECK_ARG(fused_shape_ok(C, H, W, P, O), "corr+conv+relu: unsupported shape (needs P=17, W%%4==0, W<=128, O<=256)");
  if ((int64_t)B * H == 0) return PMT_OK;
  FusedArgs f{};
  f.a = a, f.b = b, f.wt = wt, f.z = const_cast<float*>(z), f.corr = const_cast<float*>(corr), f.gz = gz;
  f.ga = ga, f.gb = gb, f.gw_part = gw_part;
  fill(&f, B, C, H, W, O);
  f.stages = pick_stages(fused_bwd_smem, W, O);
  const size_t smem = fused_bwd_smem(W, O, f.stages);
  PMT_CUDA_OK(cudaFuncSetAttribute(corr_conv_relu_bwd_kernel<17>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int rows = B * H;
  const int grid = rows < 2 * sm_count() ? rows : 2 * sm_count();
  corr_conv_relu_bwd_kernel<17><<<grid, kFThreads, smem, st>>>(f);
  PMT_LAUNCH_OK("corr_conv_relu_bwd_kernel");
  const int n = O * 17;
  corr_conv_gw_reduce_kernel<<<ceil_div(n, 256), 256, 0, st>>>(gw_part, gw, rows, n);
  PMT_LAUNCH_OK("corr_conv_gw_reduce_kernel");
  return PMT_OK;
}

}  // namespace pmt
