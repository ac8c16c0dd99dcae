// warp1d.cu -- apply_disparity(img, x_offset, wrap_mode='edge'), models/torch_dsnet.py:10-86: a fused
// horizontal linear-interpolation gather (call sites models/dsnet_t2_warp.py:294,572,697,811,946).
// The reference runs ~25 ATen kernels and materialises C x (N*H*W) int64 index tensors twice; here
// one thread owns one (n,h,w), computes the two taps once and loops over channels.  Every fp32 step
// of the reference (including the float32 flat gather index of torch_dsnet.py:59-70 and the
// un-fused weight*pixel products) is reproduced, so the forward is bit-identical to the reference.
#include "common.cuh"

namespace pmt {
namespace {

struct Taps {
  int64_t il, ir;   // flat indices into the (N*H*W) pixel grid, as the reference computes them
  float wl, wr;     // x1 - x, x - x0
  bool pass;        // clamp passes the gradient (0 <= w+off <= W-1)
};

__device__ __forceinline__ Taps make_taps(int n, int h, int w, float off, int H, int W, int64_t total) {
  Taps t;
  const float xr = __fadd_rn((float)w, off);
  const float wm1 = (float)(W - 1);
  const float x = fminf(fmaxf(xr, 0.f), wm1);
  const float x0 = floorf(x);
  const float x1 = fminf(__fadd_rn(x0, 1.f), wm1);
  // base = dim1*arange(N) (fp32); base_y0 = base + y0*dim2; idx = base_y0 + x{0,1}  -- all fp32
  const float base = __fmul_rn((float)((int64_t)W * H), (float)n);
  const float by = __fadd_rn(base, __fmul_rn((float)h, (float)W));
  int64_t il = (int64_t)__fadd_rn(by, x0), ir = (int64_t)__fadd_rn(by, x1);
  t.il = il < total ? il : total - 1;  // the reference would raise; stay inside the buffer
  t.ir = ir < total ? ir : total - 1;
  t.wl = __fsub_rn(x1, x);
  t.wr = __fsub_rn(x, x0);
  t.pass = (xr >= 0.f) && (xr <= wm1);
  return t;
}

__global__ void __launch_bounds__(256) warp_fwd_kernel(const float* __restrict__ img,
                                                       const float* __restrict__ off,
                                                       float* __restrict__ out, int N, int C, int H,
                                                       int W, int out_cnhw) {
  const int64_t plane = (int64_t)H * W, total = (int64_t)N * plane;
  for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < total;
       q += (int64_t)gridDim.x * blockDim.x) {
    const int n = (int)(q / plane);
    const int64_t k = q % plane;
    const int h = (int)(k / W), w = (int)(k % W);
    const Taps t = make_taps(n, h, w, __ldg(off + q), H, W, total);
    const int64_t nl = t.il / plane, kl = t.il % plane, nr = t.ir / plane, kr = t.ir % plane;
    const float* pl = img + nl * C * plane + kl;
    const float* pr = img + nr * C * plane + kr;
    float* o = out_cnhw ? out + q : out + (int64_t)n * C * plane + k;
    const int64_t ostride = out_cnhw ? total : plane;
#pragma unroll 4
    for (int c = 0; c < C; ++c) {
      const float a = __fmul_rn(t.wl, __ldg(pl + c * plane));
      const float b = __fmul_rn(t.wr, __ldg(pr + c * plane));
      st_cs(o + c * ostride, __fadd_rn(a, b));
    }
  }
}

// Backward.  gimg is a data-dependent scatter (the reference's gather backward is scatter_add with
// atomics as well): fp32 RED atomics into the caller-zeroed gimg; taps with zero weight are skipped.
// goff is a per-pixel reduction over channels and is deterministic.
__global__ void __launch_bounds__(256) warp_bwd_kernel(const float* __restrict__ img,
                                                       const float* __restrict__ off,
                                                       const float* __restrict__ gout,
                                                       float* __restrict__ gimg,
                                                       float* __restrict__ goff, int N, int C, int H,
                                                       int W, int gout_cnhw) {
  const int64_t plane = (int64_t)H * W, total = (int64_t)N * plane;
  for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < total;
       q += (int64_t)gridDim.x * blockDim.x) {
    const int n = (int)(q / plane);
    const int64_t k = q % plane;
    const int h = (int)(k / W), w = (int)(k % W);
    const Taps t = make_taps(n, h, w, __ldg(off + q), H, W, total);
    const int64_t al = (t.il / plane) * C * plane + t.il % plane;
    const int64_t ar = (t.ir / plane) * C * plane + t.ir % plane;
    const float* g = gout_cnhw ? gout + q : gout + (int64_t)n * C * plane + k;
    const int64_t gstride = gout_cnhw ? total : plane;
    float sl = 0.f, sr = 0.f;
#pragma unroll 4
    for (int c = 0; c < C; ++c) {
      const float gv = __ldg(g + c * gstride);
      if (gimg != nullptr) {
        if (t.wl != 0.f) atomicAdd(gimg + al + c * plane, t.wl * gv);
        if (t.wr != 0.f) atomicAdd(gimg + ar + c * plane, t.wr * gv);
      }
      sl = fmaf(gv, __ldg(img + al + c * plane), sl);
      sr = fmaf(gv, __ldg(img + ar + c * plane), sr);
    }
    if (goff != nullptr) goff[q] = t.pass ? (sr - sl) : 0.f;
  }
}

int warp_grid(int64_t total) {
  const int64_t blocks = ceil_div64(total, 256);
  const int64_t cap = (int64_t)sm_count() * 8;
  return (int)(blocks < cap ? blocks : cap);
}

}  // namespace

int launch_warp_fwd(const float* img, const float* off, float* out, int N, int C, int H, int W,
                    int out_cnhw, cudaStream_t st) {
  const int64_t total = (int64_t)N * H * W;
  if (total == 0 || C == 0) return PMT_OK;
  warp_fwd_kernel<<<warp_grid(total), 256, 0, st>>>(img, off, out, N, C, H, W, out_cnhw);
  PMT_LAUNCH_OK("warp_fwd_kernel");
  return PMT_OK;
}

int launch_warp_bwd(const float* img, const float* off, const float* gout, float* gimg, float* goff,
                    int N, int C, int H, int W, int gout_cnhw, cudaStream_t st) {
  const int64_t total = (int64_t)N * H * W;
  if (total == 0 || C == 0) return PMT_OK;
  warp_bwd_kernel<<<warp_grid(total), 256, 0, st>>>(img, off, gout, gimg, goff, N, C, H, W, gout_cnhw);
  PMT_LAUNCH_OK("warp_bwd_kernel");
  return PMT_OK;
}

}  // namespace pmt
