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

// Fused consumers of the warp, forward.
//   kMode 1: blend  out = (1-a)*seg + a*warp  -- models/dsnet_t2_warp.py:697-698, the same three rounded fp32 steps as
//            the reference's expression, so `out` is bit-identical to it; `warped` (NCHW, may be null) receives the
//            warped tensor the model also returns.
//   kMode 2: photo-consistency  sum_c (warp*mask - left)^2 per pixel, reduced per block into partials[blockIdx.x]
//            (double), summed in index order by warp_mse_finish_kernel -> deterministic mean (torch_implementation.py:314-317).
template <int kMode>
__global__ void __launch_bounds__(256) warp_fused_fwd_kernel(const float* __restrict__ img, const float* __restrict__ off,
                                                             const float* __restrict__ aux, const float* __restrict__ att,
                                                             float* __restrict__ out, float* __restrict__ warped,
                                                             double* __restrict__ partials, int mask_pos, int N, int C,
                                                             int H, int W) {
  const int64_t plane = (int64_t)H * W, total = (int64_t)N * plane;
  float acc = 0.f;
  for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < total;
       q += (int64_t)gridDim.x * blockDim.x) {
    const int n = (int)(q / plane);
    const int64_t k = q % plane;
    const int h = (int)(k / W), w = (int)(k % W);
    const float o = __ldg(off + q);
    const Taps t = make_taps(n, h, w, o, H, W, total);
    const int64_t nl = t.il / plane, kl = t.il % plane, nr = t.ir / plane, kr = t.ir % plane;
    const float* pl = img + nl * C * plane + kl;
    const float* pr = img + nr * C * plane + kr;
    const int64_t base = (int64_t)n * C * plane + k;
    const float at = kMode == 1 ? __ldg(att + q) : 0.f;
    const float om = __fsub_rn(1.f, at);
    const float mk = (kMode == 2 && mask_pos && !(o < 0.f)) ? 0.f : 1.f;
#pragma unroll 4
    for (int c = 0; c < C; ++c) {
      const float wv = __fadd_rn(__fmul_rn(t.wl, __ldg(pl + c * plane)), __fmul_rn(t.wr, __ldg(pr + c * plane)));
      if (kMode == 1) {
        const float sv = __ldg(aux + base + c * plane);
        st_cs(out + base + c * plane, __fadd_rn(__fmul_rn(om, sv), __fmul_rn(at, wv)));
        if (warped != nullptr) st_cs(warped + base + c * plane, wv);
      } else {
        const float d = __fsub_rn(__fmul_rn(wv, mk), __ldg(aux + base + c * plane));
        acc = fmaf(d, d, acc);
      }
    }
  }
  if (kMode == 2) {
    __shared__ double red[256];
    red[threadIdx.x] = (double)acc;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
      if ((int)threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
      __syncthreads();
    }
    if (threadIdx.x == 0) partials[blockIdx.x] = red[0];
  }
}

__global__ void __launch_bounds__(32) warp_mse_finish_kernel(const double* __restrict__ partials, int n, double inv_numel,
                                                             float* __restrict__ loss) {
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int i = 0; i < n; ++i) s += partials[i];   // index order: bit-reproducible
    loss[0] = (float)(s * inv_numel);
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

// ---------------------------------------------------------------------------------------------------------------------
// Deterministic backward (and the fused consumers of SURVEY.md section 8 f4).
//
// While N*H*W < 2^24 the reference's float32 flat index is exact, so both taps of a pixel stay inside its own image row:
// the scatter into gimg never leaves the row (n, h).  One CTA owns one row.  The tap structure -- which source pixels w
// feed which destination x', with which weight -- does not depend on the channel, so it is built ONCE per row as a small
// CSR in shared memory (count with integer atomics, exclusive scan, stable fill in increasing source order by one warp
// with __match_any_sync), and every channel then GATHERS: gimg[c][x'] = sum over the bucket of x' of weight * g[c][w], in
// bucket order.  No float atomics, every element of gimg written exactly once (no zero-fill launch), bit-reproducible run
// to run.  The 8 warps of the CTA split the channels; each stages its g row (and image row, for goff) in shared memory
// with coalesced loads.  goff = sum_c g * (img[x1] - img[x0]) is accumulated per warp in registers and combined across
// warps in fixed order.
//
// kMode selects where the upstream gradient of the warped tensor comes from (the consumers of the warp fused in):
//   0 plain      ge[c][w] = gout[c][w]
//   1 blend      out = (1-a)*seg + a*warp (models/dsnet_t2_warp.py:697-698):  ge = a*gout (+ gwarp when the caller also
//                used the warped tensor); also writes gseg = (1-a)*gout and gatt[w] = sum_c gout*(warp - seg)
//   2 photo MSE  loss = mean((warp*mask - left)^2) (torch_implementation.py:314-317, mask = (disp > 0) of
//                dsnet_t2_warp.py:811):  ge = mask * scale * (warp*mask - left), scale = 2*gloss/numel
// ---------------------------------------------------------------------------------------------------------------------
constexpr int kRowWarps = 8, kRowThreads = 32 * kRowWarps;
constexpr int kMaxWPerLane = 32;   // W <= 1024: per-lane goff partials live in registers

struct RowBwdArgs {
  const float* img;      // (N,C,H,W) source image of the warp
  const float* off;      // (N,1,H,W)
  const float* gout;     // mode 0/1: upstream gradient; layout per gout_cnhw.  mode 2: unused
  const float* aux;      // mode 1: seg (N,C,H,W) the warp is blended with;  mode 2: left (N,C,H,W)
  const float* att;      // mode 1: (N,1,H,W) blend weight
  const float* gwarp;    // mode 1: optional gradient w.r.t. the warped tensor itself (same layout as gout), may be null
  float* gimg;           // (N,C,H,W) or null
  float* goff;           // (N,1,H,W) or null
  float* gaux;           // mode 1: gseg (N,C,H,W);  mode 2: gleft (N,C,H,W) or null
  float* gatt;           // mode 1: (N,1,H,W)
  float scale;           // mode 2: 2 / numel
  const float* gloss;    // mode 2: device scalar, upstream gradient of the loss (null = 1)
  int mask_pos;          // mode 2: multiply the warp by (off < 0), i.e. (disp > 0) for off = -disp
  int N, C, H, W, gout_cnhw;
};

__device__ __forceinline__ int block_exclusive_scan(int v, int* warp_sums, int tid) {
  // exclusive scan of one int per thread over the 256-thread CTA; returns the exclusive prefix, total in warp_sums[8]
  const int lane = tid & 31, wid = tid >> 5;
  int x = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const int y = __shfl_up_sync(0xffffffffu, x, d);
    if (lane >= d) x += y;
  }
  if (lane == 31) warp_sums[wid] = x;
  __syncthreads();
  if (wid == 0) {
    int s = lane < kRowWarps ? warp_sums[lane] : 0;
#pragma unroll
    for (int d = 1; d < kRowWarps; d <<= 1) {
      const int y = __shfl_up_sync(0xffffffffu, s, d);
      if (lane >= d) s += y;
    }
    if (lane < kRowWarps) warp_sums[lane] = s;   // inclusive
  }
  __syncthreads();
  const int base = wid == 0 ? 0 : warp_sums[wid - 1];
  return base + x - v;
}

template <int kMode>
__global__ void __launch_bounds__(kRowThreads) warp_bwd_rows_kernel(const RowBwdArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int W = a.W, C = a.C;
  // shared layout (all sizes multiples of 4 bytes): weights, taps, CSR, per-warp rows
  float* twl = reinterpret_cast<float*>(smem_raw);
  float* twr = twl + W;
  float* ev = twr + W;                                   // [2W] bucket weights
  int* offs = reinterpret_cast<int*>(ev + 2 * W);        // [W+1] bucket starts
  int* cur = offs + (W + 1);                             // [W] counts, then fill cursors
  int* wsum = cur + W;                                   // [8] scan scratch
  unsigned short* tx0 = reinterpret_cast<unsigned short*>(wsum + 8);
  unsigned short* tx1 = tx0 + W;
  unsigned short* ew = tx1 + W;                          // [2W] bucket sources
  unsigned char* tpass = reinterpret_cast<unsigned char*>(ew + 2 * W);   // [W]
  const int W4 = (W + 3) & ~3;
  float* rows = reinterpret_cast<float*>(smem_raw + (((size_t)(4 * W + (2 * W + 9)) * 4 + (size_t)4 * W * 2 + W + 15) & ~(size_t)15));
  const int tid = threadIdx.x, wid = tid >> 5, lane = tid & 31;
  float* gs = rows + (size_t)wid * 3 * W4;               // staged effective gradient row of this warp's channel
  float* is = gs + W4;                                   // staged image row
  float* xs = is + W4;                                   // mode 1: att row / mode 2: unused (per warp copy keeps it simple)
  const int64_t plane = (int64_t)a.H * W, total = (int64_t)a.N * plane;
  const bool need_img = a.gimg != nullptr;
  const bool need_is = a.goff != nullptr || kMode != 0;   // the image row is needed for goff and to recompute the warp

  for (int row = blockIdx.x; row < a.N * a.H; row += gridDim.x) {
    const int n = row / a.H, h = row % a.H;
    const int64_t roff = (int64_t)n * plane + (int64_t)h * W;   // offset of the row inside an (N,1,H,W) tensor
    // ---- A. taps of every pixel of the row ----
    for (int w = tid; w < W; w += kRowThreads) {
      const Taps t = make_taps(n, h, w, __ldg(a.off + roff + w), a.H, W, total);
      const int x0 = (int)(t.il - roff), x1 = (int)(t.ir - roff);     // exact below 2^24: 0 <= x0 <= x1 <= W-1
      tx0[w] = (unsigned short)x0, tx1[w] = (unsigned short)x1;
      twl[w] = t.wl, twr[w] = t.wr;
      tpass[w] = t.pass ? 1 : 0;
      cur[w] = 0;
    }
    __syncthreads();
    if (need_img) {
      // ---- B. bucket sizes (integer atomics: deterministic) ----
      for (int w = tid; w < W; w += kRowThreads) {
        if (twl[w] != 0.f) atomicAdd(&cur[tx0[w]], 1);
        if (twr[w] != 0.f) atomicAdd(&cur[tx1[w]], 1);
      }
      __syncthreads();
      // ---- C. exclusive scan -> bucket starts ----
      {
        const int per = (W + kRowThreads - 1) / kRowThreads;
        const int b0 = tid * per;
        int local = 0;
        for (int i = 0; i < per; ++i)
          if (b0 + i < W) local += cur[b0 + i];
        int run = block_exclusive_scan(local, wsum, tid);
        for (int i = 0; i < per; ++i)
          if (b0 + i < W) {
            const int c = cur[b0 + i];
            offs[b0 + i] = run;
            cur[b0 + i] = run;
            run += c;
          }
        if (tid == kRowThreads - 1) offs[W] = wsum[kRowWarps - 1];
      }
      __syncthreads();
      // ---- D. stable fill by warp 0: bucket entries in increasing (32-pixel chunk, tap kind, lane) order ----
      if (wid == 0) {
        for (int base = 0; base < W; base += 32) {
          const int w = base + lane;
#pragma unroll
          for (int kind = 0; kind < 2; ++kind) {
            const float wt = w < W ? (kind ? twr[w] : twl[w]) : 0.f;
            const int key = w < W ? (int)(kind ? tx1[w] : tx0[w]) : 0;
            const bool valid = wt != 0.f;
            const unsigned vm = __ballot_sync(0xffffffffu, valid);
            int pos = 0;
            unsigned m = 0;
            if (valid) {
              m = __match_any_sync(vm, key);
              pos = cur[key] + __popc(m & ((1u << lane) - 1u));
            }
            __syncwarp();
            if (valid && lane == __ffs(m) - 1) cur[key] += __popc(m);
            if (valid) {
              ew[pos] = (unsigned short)w;
              ev[pos] = wt;
            }
            __syncwarp();
          }
        }
      }
      __syncthreads();
    }
    // ---- E. channels: warp `wid` takes c = wid, wid+8, ... ----
    const float sc = kMode == 2 ? a.scale * (a.gloss != nullptr ? __ldg(a.gloss) : 1.f) : 0.f;
    float dgo[kMaxWPerLane];    // goff partial of pixel w = lane + 32 i
    float dga[kMode == 1 ? kMaxWPerLane : 1];   // gatt partial
#pragma unroll
    for (int i = 0; i < kMaxWPerLane; ++i) dgo[i] = 0.f;
    if (kMode == 1) {
#pragma unroll
      for (int i = 0; i < (kMode == 1 ? kMaxWPerLane : 1); ++i) dga[i] = 0.f;
      for (int w = lane; w < W; w += 32) xs[w] = __ldg(a.att + roff + w);
    }
    for (int c = wid; c < C; c += kRowWarps) {
      const int64_t coff = ((int64_t)n * C + c) * plane + (int64_t)h * W;     // row inside an (N,C,H,W) tensor
      const int64_t goff_ = a.gout_cnhw ? (int64_t)c * total + roff : coff;   // row inside gout / gwarp
      if (need_is)
        for (int w = lane; w < W; w += 32) is[w] = __ldg(a.img + coff + w);
      if (kMode == 0)
        for (int w = lane; w < W; w += 32) gs[w] = __ldg(a.gout + goff_ + w);
      __syncwarp();
      if (kMode == 1) {
        // blend: recompute the warp, split the upstream gradient
#pragma unroll
        for (int i = 0; i < kMaxWPerLane; ++i) {
          const int w = lane + 32 * i;
          if (w < W) {
            const float g = __ldg(a.gout + goff_ + w), at = xs[w];
            const float wv = __fadd_rn(__fmul_rn(twl[w], is[tx0[w]]), __fmul_rn(twr[w], is[tx1[w]]));
            const float sv = __ldg(a.aux + coff + w);
            dga[i] = fmaf(g, wv - sv, dga[i]);
            a.gaux[coff + w] = (1.f - at) * g;
            float ge = at * g;
            if (a.gwarp != nullptr) ge += __ldg(a.gwarp + goff_ + w);
            gs[w] = ge;
          }
        }
        __syncwarp();
      } else if (kMode == 2) {
#pragma unroll
        for (int i = 0; i < kMaxWPerLane; ++i) {
          const int w = lane + 32 * i;
          if (w < W) {
            float wv = __fadd_rn(__fmul_rn(twl[w], is[tx0[w]]), __fmul_rn(twr[w], is[tx1[w]]));
            const float mk = (a.mask_pos && !(__ldg(a.off + roff + w) < 0.f)) ? 0.f : 1.f;
            wv *= mk;
            const float d = sc * (wv - __ldg(a.aux + coff + w));
            if (a.gaux != nullptr) a.gaux[coff + w] = -d;
            gs[w] = mk * d;
          }
        }
        __syncwarp();
      }
      if (a.goff != nullptr) {
#pragma unroll
        for (int i = 0; i < kMaxWPerLane; ++i) {
          const int w = lane + 32 * i;
          if (w < W) dgo[i] = fmaf(gs[w], is[tx1[w]] - is[tx0[w]], dgo[i]);
        }
      }
      if (need_img) {
        float* o = a.gimg + coff;
        for (int x = lane; x < W; x += 32) {
          float acc = 0.f;
          const int e1 = offs[x + 1];
          for (int e = offs[x]; e < e1; ++e) acc = __fadd_rn(acc, __fmul_rn(ev[e], gs[ew[e]]));
          st_cs(o + x, acc);
        }
      }
      __syncwarp();   // gs / is are overwritten by the next channel
    }
    // ---- F. combine the per-warp goff / gatt partials in fixed order ----
    __syncthreads();
    if (a.goff != nullptr || kMode == 1) {
#pragma unroll
      for (int i = 0; i < kMaxWPerLane; ++i) {
        const int w = lane + 32 * i;
        if (w < W) {
          gs[w] = dgo[i];
          if (kMode == 1) is[w] = dga[i];
        }
      }
      __syncthreads();
      for (int w = tid; w < W; w += kRowThreads) {
        float s = 0.f, s2 = 0.f;
#pragma unroll
        for (int k = 0; k < kRowWarps; ++k) {
          s += rows[(size_t)k * 3 * W4 + w];
          if (kMode == 1) s2 += rows[(size_t)k * 3 * W4 + W4 + w];
        }
        if (a.goff != nullptr) a.goff[roff + w] = tpass[w] ? s : 0.f;
        if (kMode == 1) a.gatt[roff + w] = s2;
      }
    }
    __syncthreads();   // shared memory is reused by the next row
  }
}

size_t rows_smem_bytes(int W) {
  const int W4 = (W + 3) & ~3;
  const size_t head = (((size_t)(4 * W + (2 * W + 9)) * 4 + (size_t)4 * W * 2 + W + 15) & ~(size_t)15);
  return head + (size_t)kRowWarps * 3 * W4 * 4;
}

// The row kernel needs exact float32 flat indices (taps inside the row) and W <= 1024 (register partials).
bool rows_ok(int N, int H, int W) {
  return (int64_t)N * H * W < (1ll << 24) && W <= 32 * kMaxWPerLane && W < 65536 && rows_smem_bytes(W) <= 200 * 1024;
}

template <int kMode>
int launch_rows(const RowBwdArgs& a, cudaStream_t st) {
  const size_t smem = rows_smem_bytes(a.W);
  PMT_CUDA_OK(cudaFuncSetAttribute(warp_bwd_rows_kernel<kMode>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int64_t rows = (int64_t)a.N * a.H;
  const int per_sm = (int)((220 * 1024) / (smem + 1024));
  int64_t grid = (int64_t)sm_count() * (per_sm < 1 ? 1 : (per_sm > 8 ? 8 : per_sm));
  if (grid > rows) grid = rows;
  warp_bwd_rows_kernel<kMode><<<(unsigned)grid, kRowThreads, smem, st>>>(a);
  PMT_LAUNCH_OK("warp_bwd_rows_kernel");
  return PMT_OK;
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

// gimg is fully written by this call on every path (the caller does not zero it).
int launch_warp_bwd(const float* img, const float* off, const float* gout, float* gimg, float* goff,
                    int N, int C, int H, int W, int gout_cnhw, cudaStream_t st) {
  const int64_t total = (int64_t)N * H * W;
  if (total == 0 || C == 0) return PMT_OK;
  if (rows_ok(N, H, W)) {
    RowBwdArgs a{};
    a.img = img, a.off = off, a.gout = gout, a.gimg = gimg, a.goff = goff;
    a.N = N, a.C = C, a.H = H, a.W = W, a.gout_cnhw = gout_cnhw;
    return launch_rows<0>(a, st);
  }
  // N*H*W >= 2^24 (the reference's float32 indices are inexact there and may leave the row) or W > 1024: scatter with
  // fp32 atomics into gimg, zeroed here
  if (gimg != nullptr) PMT_CUDA_OK(cudaMemsetAsync(gimg, 0, sizeof(float) * (size_t)total * C, st));
  warp_bwd_kernel<<<warp_grid(total), 256, 0, st>>>(img, off, gout, gimg, goff, N, C, H, W, gout_cnhw);
  PMT_LAUNCH_OK("warp_bwd_kernel");
  return PMT_OK;
}

bool warp_rows_supported(int N, int H, int W) { return rows_ok(N, H, W); }

int launch_warp_blend_fwd(const float* img, const float* off, const float* att, const float* seg, float* out, float* warped,
                          int N, int C, int H, int W, cudaStream_t st) {
  const int64_t total = (int64_t)N * H * W;
  if (total == 0 || C == 0) return PMT_OK;
  warp_fused_fwd_kernel<1><<<warp_grid(total), 256, 0, st>>>(img, off, seg, att, out, warped, nullptr, 0, N, C, H, W);
  PMT_LAUNCH_OK("warp_fused_fwd_kernel<blend>");
  return PMT_OK;
}

int warp_mse_workspace() { return 148 * 8 * 2; }   // doubles; >= any grid warp_grid() can return on a B200-class part

int launch_warp_mse_fwd(const float* img, const float* off, const float* left, int mask_pos, double* partials, float* loss,
                        int N, int C, int H, int W, cudaStream_t st) {
  const int64_t total = (int64_t)N * H * W;
  PMT_CHECK_ARG(total > 0 && C > 0, "photo-consistency MSE of an empty tensor is undefined");
  const int grid = warp_grid(total);
  PMT_CHECK_ARG(grid <= warp_mse_workspace(), "photo-consistency MSE: workspace too small for this device");
  warp_fused_fwd_kernel<2><<<grid, 256, 0, st>>>(img, off, left, nullptr, nullptr, nullptr, partials, mask_pos, N, C, H, W);
  PMT_LAUNCH_OK("warp_fused_fwd_kernel<mse>");
  warp_mse_finish_kernel<<<1, 32, 0, st>>>(partials, grid, 1.0 / ((double)total * C), loss);
  PMT_LAUNCH_OK("warp_mse_finish_kernel");
  return PMT_OK;
}

int launch_warp_blend_bwd(const float* img, const float* off, const float* att, const float* seg, const float* gout,
                          const float* gwarp, float* gimg, float* goff, float* gatt, float* gseg, int N, int C, int H,
                          int W, cudaStream_t st) {
  if ((int64_t)N * H * W == 0 || C == 0) return PMT_OK;
  PMT_CHECK_ARG(rows_ok(N, H, W), "warp_blend backward: needs N*H*W < 2^24 and W <= 1024");
  RowBwdArgs a{};
  a.img = img, a.off = off, a.att = att, a.aux = seg, a.gout = gout, a.gwarp = gwarp;
  a.gimg = gimg, a.goff = goff, a.gatt = gatt, a.gaux = gseg;
  a.N = N, a.C = C, a.H = H, a.W = W, a.gout_cnhw = 0;
  return launch_rows<1>(a, st);
}

int launch_warp_mse_bwd(const float* img, const float* off, const float* left, int mask_pos, const float* gloss, float* gimg,
                        float* goff, float* gleft, int N, int C, int H, int W, cudaStream_t st) {
  if ((int64_t)N * H * W == 0 || C == 0) return PMT_OK;
  PMT_CHECK_ARG(rows_ok(N, H, W), "warp photo-consistency backward: needs N*H*W < 2^24 and W <= 1024");
  RowBwdArgs a{};
  a.img = img, a.off = off, a.aux = left, a.mask_pos = mask_pos, a.gloss = gloss;
  a.scale = (float)(2.0 / ((double)N * C * H * W));
  a.gimg = gimg, a.goff = goff, a.gaux = gleft;
  a.N = N, a.C = C, a.H = H, a.W = W, a.gout_cnhw = 0;
  return launch_rows<2>(a, st);
}

}  // namespace pmt
