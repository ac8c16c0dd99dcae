// corr_generic.cu -- shape-generic correlation kernels (any patchH x patchW, any dilation_patch,
// any W / alignment).  Used for the 2-D `-corrType 2dcorr` patches (models/dsnet_t2.py:129-133), the
// (1,21) dilation_patch=4 sampler (models/torch_dsnet.py:133-138) and as the fall-back of the 1-D
// entry points when the tiled fast path's alignment requirements do not hold.  Still CUDA, still
// deterministic (both gradients are gathers); one thread per output element, w fastest so every
// global access is coalesced along the row.
#include "common.cuh"

namespace pmt {
namespace {

struct GArgs {
  int B, C, H, W, pH, pW, dpH, dpW, rH, rW;
};

__global__ void __launch_bounds__(256) corr_generic_fwd_kernel(const float* __restrict__ in1,
                                                               const float* __restrict__ in2,
                                                               float* __restrict__ out, GArgs a,
                                                               int64_t total) {
  const int64_t plane = (int64_t)a.H * a.W;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int w = (int)(idx % a.W);
    int64_t t = idx / a.W;
    const int h = (int)(t % a.H);
    t /= a.H;
    const int pw = (int)(t % a.pW);
    t /= a.pW;
    const int ph = (int)(t % a.pH);
    const int n = (int)(t / a.pH);
    const int h2 = h + (ph - a.rH) * a.dpH, w2 = w + (pw - a.rW) * a.dpW;
    float acc = 0.f;
    if (h2 >= 0 && h2 < a.H && w2 >= 0 && w2 < a.W) {
      const float* p1 = in1 + (int64_t)n * a.C * plane + (int64_t)h * a.W + w;
      const float* p2 = in2 + (int64_t)n * a.C * plane + (int64_t)h2 * a.W + w2;
#pragma unroll 4
      for (int c = 0; c < a.C; ++c) acc = fmaf(__ldg(p1 + c * plane), __ldg(p2 + c * plane), acc);
    }
    out[idx] = acc;
  }
}

// One thread per (n, c, h, w): gin1 at (h,w) gathers over in2 shifted forward, gin2 at (h,w) gathers
// over in1 shifted backward.  Sum order: (ph, pw) ascending, as the upstream CPU loop nest.
__global__ void __launch_bounds__(256) corr_generic_bwd_kernel(const float* __restrict__ in1,
                                                               const float* __restrict__ in2,
                                                               const float* __restrict__ gout,
                                                               float* __restrict__ gin1,
                                                               float* __restrict__ gin2, GArgs a,
                                                               int64_t total) {
  const int64_t plane = (int64_t)a.H * a.W;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int w = (int)(idx % a.W);
    int64_t t = idx / a.W;
    const int h = (int)(t % a.H);
    t /= a.H;
    const int c = (int)(t % a.C);
    const int n = (int)(t / a.C);
    const float* x1 = in1 + ((int64_t)n * a.C + c) * plane;
    const float* x2 = in2 + ((int64_t)n * a.C + c) * plane;
    const float* g = gout + (int64_t)n * a.pH * a.pW * plane;
    float acc1 = 0.f, acc2 = 0.f;
    for (int ph = 0; ph < a.pH; ++ph) {
      const int sh = (ph - a.rH) * a.dpH;
      const int hf = h + sh, hb = h - sh;
      for (int pw = 0; pw < a.pW; ++pw) {
        const int sw = (pw - a.rW) * a.dpW;
        const int wf = w + sw, wb = w - sw;
        const float* gp = g + (int64_t)(ph * a.pW + pw) * plane;
        if (hf >= 0 && hf < a.H && wf >= 0 && wf < a.W)
          acc1 = fmaf(__ldg(gp + (int64_t)h * a.W + w), __ldg(x2 + (int64_t)hf * a.W + wf), acc1);
        if (hb >= 0 && hb < a.H && wb >= 0 && wb < a.W)
          acc2 = fmaf(__ldg(gp + (int64_t)hb * a.W + wb), __ldg(x1 + (int64_t)hb * a.W + wb), acc2);
      }
    }
    gin1[idx] = acc1;
    gin2[idx] = acc2;
  }
}

int grid_for(int64_t total) {
  int64_t blocks = ceil_div64(total, 256);
  const int64_t cap = (int64_t)sm_count() * 32;
  return (int)(blocks < cap ? blocks : cap);
}

}  // namespace

int launch_corr_generic_fwd(const float* in1, const float* in2, float* out, int B, int C, int H, int W,
                            int pH, int pW, int dpH, int dpW, cudaStream_t st) {
  GArgs a{B, C, H, W, pH, pW, dpH, dpW, (pH - 1) / 2, (pW - 1) / 2};
  const int64_t total = (int64_t)B * pH * pW * H * W;
  if (total == 0) return PMT_OK;
  corr_generic_fwd_kernel<<<grid_for(total), 256, 0, st>>>(in1, in2, out, a, total);
  PMT_LAUNCH_OK("corr_generic_fwd_kernel");
  return PMT_OK;
}

int launch_corr_generic_bwd(const float* in1, const float* in2, const float* gout, float* gin1,
                            float* gin2, int B, int C, int H, int W, int pH, int pW, int dpH, int dpW,
                            cudaStream_t st) {
  GArgs a{B, C, H, W, pH, pW, dpH, dpW, (pH - 1) / 2, (pW - 1) / 2};
  const int64_t total = (int64_t)B * C * H * W;
  if (total == 0) return PMT_OK;
  corr_generic_bwd_kernel<<<grid_for(total), 256, 0, st>>>(in1, in2, gout, gin1, gin2, a, total);
  PMT_LAUNCH_OK("corr_generic_bwd_kernel");
  return PMT_OK;
}

}  // namespace pmt
