// corr2d_rows.cu -- f3: the 2-D (pH x pW) patch of `-corrType 2dcorr`
// (models/dsnet_t2.py:129-133, :221-223, :845: patch (17,17), 289 planes, kernel_size 1, stride 1, padding 0):
//   out[n,ph,pw,h,w] = sum_c a[n,c,h,w] * b[n,c,h+sh,w+sw],   sh = ph-(pH-1)/2, sw = pw-(pW-1)/2   (OOB terms skipped)
// composed, as the survey proposed, of pH row-shifted passes of the 1-D correlation: for a fixed ph the planes
// out[n,ph,:,h,:] are the 1 x pW correlation of row h of `a` with row h+sh of `b`.
//
// The generic kernel (corr_generic.cu) gives every output element its own thread, which streams the C values of both
// inputs from global memory with no reuse (289x re-read of in2 for a 17 x 17 patch).  Here:
//   forward : one CTA per (image row, ph).  Both rows stream through shared memory in 32-channel chunks (cp.async, double
//             buffered, zero halo = border rule); a thread owns 2 columns x all pW shifts (34 accumulators) for a quarter
//             / eighth of the chunk's channels: per channel 1 + 9 LDS.64 feed 34 FFMA; channel groups are then summed in
//             fixed order.  A row of b outside the image writes zeros.
//   backward: one CTA per (image row, gradient).  All pH x pW coefficient rows of that image row (the slice of gout it
//             needs, 74 KB for 17 x 17 x 64) stay resident in shared memory; for every 32-channel chunk the pH source
//             rows stream through a double buffer and each is filtered with its own pW taps (per-column coefficients),
//             accumulating in registers over ph:
//               ga[c,h,w]  = sum_ph sum_pw g[ph,pw,h,w]        * b[c,h+sh,w+sw]
//               gb[c,h,w'] = sum_ph sum_pw g[ph,pw,h-sh,w'-sw] * a[c,h-sh,w'-sw]
// fp32 FFMA, fixed summation order (deterministic, no atomics).
#include "common.cuh"

namespace pmt {
namespace {

constexpr int kCThreads = 256;
constexpr int kCChunk = 32;     // channels per shared-memory stage
constexpr int kCMaxW = 128;
constexpr int kCPad = 8;        // zero halo (>= max |sw| = 8 for pW = 17), multiple of 4

struct C2Args {
  const float* a;
  const float* b;
  const float* g;      // backward: gout (B,pH,pW,H,W)
  float* out;          // forward
  float* ga;
  float* gb;
  int B, C, H, W, pH;
  int Wp;              // W + 2*kCPad
  int G;               // channel groups per chunk = kCThreads / (W/2)
};

// stage rows src[n, c0.., hs, :] of one tensor into a [kCChunk][Wp] tile (interior at column kCPad); a row outside the
// image or a channel >= C is zero-filled
__device__ __forceinline__ void stage_rows(const C2Args& f, const float* src, float* dst, int n, int hs, int c0, int tid) {
  const int W4 = f.W / 4;
  const bool row_ok = hs >= 0 && hs < f.H;
  for (int i = tid; i < kCChunk * W4; i += kCThreads) {
    const int c = i / W4, q = i - c * W4;
    const bool valid = row_ok && c0 + c < f.C;
    const float* s = src + (((int64_t)n * f.C + (valid ? c0 + c : 0)) * f.H + (valid ? hs : 0)) * (int64_t)f.W + 4 * q;
    cp_async16(dst + c * f.Wp + kCPad + 4 * q, s, valid);
  }
}

template <int kPW>
__global__ void __launch_bounds__(kCThreads) corr2d_fwd_rows_kernel(const C2Args f) {
  extern __shared__ __align__(16) float sm[];
  constexpr int r = (kPW - 1) / 2;
  const int W = f.W, Wp = f.Wp, pairs = W / 2, G = f.G;
  float* As = sm;                                  // [2][kCChunk][Wp]
  float* Bs = As + 2 * kCChunk * Wp;               // [2][kCChunk][Wp]
  float* part = Bs + 2 * kCChunk * Wp;             // [G][kPW][W]
  const int tid = threadIdx.x;
  const int pair = tid % pairs, grp = tid / pairs;
  const bool worker = grp < G;
  const int row = blockIdx.x, ph = blockIdx.y;
  const int n = row / f.H, h = row % f.H;
  const int hb = h + ph - (f.pH - 1) / 2;
  float* o = f.out + ((((int64_t)n * f.pH + ph) * kPW) * f.H + h) * (int64_t)W;   // plane pw at o + pw*H*W
  const int64_t pstride = (int64_t)f.H * W;
  if (hb < 0 || hb >= f.H) {                        // the whole b row is outside the image: zero planes
    for (int i = tid; i < kPW * W; i += kCThreads) o[(int64_t)(i / W) * pstride + i % W] = 0.f;
    return;
  }
  const int n_chunks = (f.C + kCChunk - 1) / kCChunk;
  for (int i = tid; i < 4 * kCChunk * Wp; i += kCThreads) As[i] = 0.f;   // zero halos (never overwritten)
  __syncthreads();
  float acc0[kPW], acc1[kPW];
#pragma unroll
  for (int p = 0; p < kPW; ++p) acc0[p] = 0.f, acc1[p] = 0.f;
  stage_rows(f, f.a, As, n, h, 0, tid);
  stage_rows(f, f.b, Bs, n, hb, 0, tid);
  cp_async_commit();
  for (int k = 0; k < n_chunks; ++k) {
    const int s = k & 1;
    if (k + 1 < n_chunks) {
      stage_rows(f, f.a, As + (s ^ 1) * kCChunk * Wp, n, h, (k + 1) * kCChunk, tid);
      stage_rows(f, f.b, Bs + (s ^ 1) * kCChunk * Wp, n, hb, (k + 1) * kCChunk, tid);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    if (worker) {
      const float* Ac = As + s * kCChunk * Wp + kCPad + 2 * pair;
      const float* Bc = Bs + s * kCChunk * Wp + kCPad - r + 2 * pair;
      for (int c = grp; c < kCChunk; c += G) {
        const float2 l = *reinterpret_cast<const float2*>(Ac + c * Wp);
        float win[kPW + 1];
#pragma unroll
        for (int j = 0; j < kPW + 1; j += 2) {
          const float2 v = *reinterpret_cast<const float2*>(Bc + c * Wp + j);
          win[j] = v.x;
          if (j + 1 < kPW + 1) win[j + 1] = v.y;
        }
#pragma unroll
        for (int p = 0; p < kPW; ++p) {
          acc0[p] = fmaf(l.x, win[p], acc0[p]);
          acc1[p] = fmaf(l.y, win[p + 1], acc1[p]);
        }
      }
    }
    __syncthreads();
  }
  if (worker) {
#pragma unroll
    for (int p = 0; p < kPW; ++p)
      *reinterpret_cast<float2*>(part + (grp * kPW + p) * W + 2 * pair) = make_float2(acc0[p], acc1[p]);
  }
  __syncthreads();
  for (int i = tid; i < kPW * W; i += kCThreads) {
    float sacc = 0.f;
    for (int g = 0; g < G; ++g) sacc += part[g * kPW * W + i];   // fixed order
    st_cs(o + (int64_t)(i / W) * pstride + i % W, sacc);
  }
}

template <int kPW>
__global__ void __launch_bounds__(kCThreads) corr2d_bwd_rows_kernel(const C2Args f) {
  extern __shared__ __align__(16) float sm[];
  constexpr int r = (kPW - 1) / 2;
  const int W = f.W, Wp = f.Wp, pairs = W / 2, G = f.G, pH = f.pH;
  float* Ss = sm;                                  // [2][kCChunk][Wp]   source rows (b for ga, a for gb)
  float* Cs = Ss + 2 * kCChunk * Wp;               // [pH][kPW][W]       coefficient rows of this image row
  const int tid = threadIdx.x;
  const int pair = tid % pairs, grp = tid / pairs;
  const bool worker = grp < G;
  const int row = blockIdx.x, mode = blockIdx.y;   // mode 0: ga, mode 1: gb
  const int n = row / f.H, h = row % f.H;
  const int rH = (pH - 1) / 2;
  const float* src = mode == 0 ? f.b : f.a;
  float* dst = mode == 0 ? f.ga : f.gb;
  const int n_chunks = (f.C + kCChunk - 1) / kCChunk;
  const int64_t pstride = (int64_t)f.H * W;
  for (int i = tid; i < 2 * kCChunk * Wp; i += kCThreads) Ss[i] = 0.f;   // zero halos
  // coefficient rows: mode 0 g[n,ph,pw,h,:], mode 1 g[n,ph,pw,h-sh,:] (zero when that row is outside the image)
  for (int i = tid; i < pH * kPW * (W / 4); i += kCThreads) {
    const int q = i % (W / 4), pp = i / (W / 4), ph = pp / kPW;
    const int hg = mode == 0 ? h : h - (ph - rH);
    const bool valid = hg >= 0 && hg < f.H;
    cp_async16(Cs + pp * W + 4 * q, f.g + ((int64_t)n * pH * kPW + pp) * pstride + (int64_t)(valid ? hg : 0) * W + 4 * q, valid);
  }
  cp_async_commit();
  cp_async_wait<0>();
  __syncthreads();
  // source row of ph: hs = h + (ph - rH) for ga, h - (ph - rH) for gb; the ph whose row lies inside the image form one
  // contiguous range [ph_lo, ph_hi)
  int ph_lo, ph_hi;
  if (mode == 0) ph_lo = rH - h, ph_hi = rH + f.H - h;
  else ph_lo = rH + h - (f.H - 1), ph_hi = rH + h + 1;
  if (ph_lo < 0) ph_lo = 0;
  if (ph_hi > pH) ph_hi = pH;
  const int n_ph = ph_hi - ph_lo;
  if (n_ph <= 0) {                                 // no source row inside the image: zero gradient row
    for (int i = tid; i < f.C * W; i += kCThreads) dst[(((int64_t)n * f.C + i / W) * f.H + h) * W + i % W] = 0.f;
    return;
  }
  const int n_steps = n_chunks * n_ph;
  auto step_src_row = [&](int step) {
    const int ph = ph_lo + step % n_ph;
    return mode == 0 ? h + (ph - rH) : h - (ph - rH);
  };
  stage_rows(f, src, Ss, n, step_src_row(0), 0, tid);
  cp_async_commit();
  const int w0 = 2 * pair;
  constexpr int kPerGrp = 8;                       // channels of a chunk per thread (kCChunk / G with G >= 4)
  float acc0[kPerGrp], acc1[kPerGrp];
  for (int step = 0; step < n_steps; ++step) {
    const int s = step & 1;
    const int kc = step / n_ph, ph = ph_lo + step % n_ph;
    if (step % n_ph == 0) {
#pragma unroll
      for (int i = 0; i < kPerGrp; ++i) acc0[i] = 0.f, acc1[i] = 0.f;
    }
    if (step + 1 < n_steps) {
      stage_rows(f, src, Ss + (s ^ 1) * kCChunk * Wp, n, step_src_row(step + 1), ((step + 1) / n_ph) * kCChunk, tid);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    if (worker) {
      // per-column coefficients of this ph
      float c0_[kPW], c1_[kPW];
      const float* S = Cs + ph * kPW * W;
#pragma unroll
      for (int p = 0; p < kPW; ++p) {
        if (mode == 0) {
          c0_[p] = S[p * W + w0];
          c1_[p] = S[p * W + w0 + 1];
        } else {
          const int u0 = w0 - (p - r), u1 = w0 + 1 - (p - r);
          c0_[p] = (u0 >= 0 && u0 < W) ? S[p * W + u0] : 0.f;
          c1_[p] = (u1 >= 0 && u1 < W) ? S[p * W + u1] : 0.f;
        }
      }
#pragma unroll
      for (int ii = 0; ii < kPerGrp; ++ii) {
        const int c = grp + ii * G;
        if (c < kCChunk) {
          // mode 0: window src[c][w0 - r .. w0 + 1 + (kPW-1-r)];  mode 1: src[c][w0 + r - (kPW-1) .. w0 + 1 + r]
          const float* Wn = Ss + s * kCChunk * Wp + c * Wp + kCPad + w0 + (mode == 0 ? -r : r - (kPW - 1));
          float win[kPW + 1];
#pragma unroll
          for (int j = 0; j < kPW + 1; ++j) win[j] = Wn[j];
          float t0 = acc0[ii], t1 = acc1[ii];
#pragma unroll
          for (int p = 0; p < kPW; ++p) {
            if (mode == 0) {
              t0 = fmaf(c0_[p], win[p], t0);
              t1 = fmaf(c1_[p], win[p + 1], t1);
            } else {
              t0 = fmaf(c0_[p], win[kPW - 1 - p], t0);
              t1 = fmaf(c1_[p], win[kPW - p], t1);
            }
          }
          acc0[ii] = t0, acc1[ii] = t1;
        }
      }
      if (step % n_ph == n_ph - 1) {               // last source row of this chunk: write the chunk's gradient rows
#pragma unroll
        for (int ii = 0; ii < kPerGrp; ++ii) {
          const int c = grp + ii * G, cg = kc * kCChunk + c;
          if (c < kCChunk && cg < f.C)
            *reinterpret_cast<float2*>(dst + (((int64_t)n * f.C + cg) * f.H + h) * W + w0) = make_float2(acc0[ii], acc1[ii]);
        }
      }
    }
    __syncthreads();
  }
}

size_t c2_fwd_smem(int W) {
  const int Wp = W + 2 * kCPad, G = kCThreads / (W / 2);
  return sizeof(float) * ((size_t)4 * kCChunk * Wp + (size_t)G * 17 * W);
}
size_t c2_bwd_smem(int W, int pH) {
  const int Wp = W + 2 * kCPad;
  return sizeof(float) * ((size_t)2 * kCChunk * Wp + (size_t)pH * 17 * W);
}

}  // namespace

// tiled path: pW == 17 (the reference's 2dcorr patch), 16 <= W <= 128, W % 4 == 0, aligned pointers, dilation_patch 1
bool corr2d_rows_ok(const void* a, const void* b, const void* third, int C, int H, int W, int pH, int pW, int dpH, int dpW) {
  if (pW != 17 || pH < 1 || pH > 33 || dpH != 1 || dpW != 1 || C < 1 || H < 1) return false;
  if (W < 16 || W > kCMaxW || W % 4 != 0 || kCThreads / (W / 2) < 4) return false;
  if (!aligned16(a) || !aligned16(b) || !aligned16(third)) return false;
  return c2_fwd_smem(W) <= 220 * 1024 && c2_bwd_smem(W, pH) <= 220 * 1024;
}

static void c2_fill(C2Args* f, int B, int C, int H, int W, int pH) {
  f->B = B, f->C = C, f->H = H, f->W = W, f->pH = pH;
  f->Wp = W + 2 * kCPad;
  f->G = kCThreads / (W / 2);
  if (f->G > kCChunk) f->G = kCChunk;
}

int launch_corr2d_fwd_rows(const float* a, const float* b, float* out, int B, int C, int H, int W, int pH, cudaStream_t st) {
  if ((int64_t)B * H == 0) return PMT_OK;
  C2Args f{};
  f.a = a, f.b = b, f.out = out;
  c2_fill(&f, B, C, H, W, pH);
  const size_t smem = sizeof(float) * ((size_t)4 * kCChunk * f.Wp + (size_t)f.G * 17 * W);
  PMT_CUDA_OK(cudaFuncSetAttribute(corr2d_fwd_rows_kernel<17>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  corr2d_fwd_rows_kernel<17><<<dim3((unsigned)(B * H), (unsigned)pH), kCThreads, smem, st>>>(f);
  PMT_LAUNCH_OK("corr2d_fwd_rows_kernel");
  return PMT_OK;
}

int launch_corr2d_bwd_rows(const float* a, const float* b, const float* g, float* ga, float* gb, int B, int C, int H, int W,
                           int pH, cudaStream_t st) {
  if ((int64_t)B * H * C == 0) return PMT_OK;
  C2Args f{};
  f.a = a, f.b = b, f.g = g, f.ga = ga, f.gb = gb;
  c2_fill(&f, B, C, H, W, pH);
  const size_t smem = c2_bwd_smem(W, pH);
  PMT_CUDA_OK(cudaFuncSetAttribute(corr2d_bwd_rows_kernel<17>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  corr2d_bwd_rows_kernel<17><<<dim3((unsigned)(B * H), 2u), kCThreads, smem, st>>>(f);
  PMT_LAUNCH_OK("corr2d_bwd_rows_kernel");
  return PMT_OK;
}

}  // namespace pmt
