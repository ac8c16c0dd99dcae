// warp1d.cu -- apply_disparity(img, x_offset, wrap_mode='edge'), models/torch_dsnet.py:10-86: a fused
// horizontal linear-interpolation gather (call sites models/dsnet_t2_warp.py:294,572,697,811,946).
// The reference runs ~25 ATen kernels and materialises C x (N*H*W) int64 index tensors twice; here
// one thread owns one (n,h,w), computes the two taps once and loops over channels.  Every fp32 step
// of the reference (including the float32 flat gather index of torch_dsnet.py:59-70 and the
// un-fused weight*pixel products) is reproduced, so the forward is bit-identical to the reference.
#include <cub/block/block_radix_sort.cuh>

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
// CSR in shared memory (bucket sizes with integer atomics, exclusive scan, and a STABLE block radix sort of the
// (destination, source) pairs -- cub::BlockRadixSort over the 2W taps -- so that every bucket lists its sources in
// increasing order), and every channel then GATHERS: gimg[c][x'] = sum over the bucket of x' of weight * g[c][w], in
// bucket order.  Buckets with more than kBigBucket entries (the clamped image borders collect dozens of taps) are summed
// by a whole warp in a fixed lane-strided order + shuffle tree instead of by one thread.  No float atomics, every element of gimg written exactly once (no zero-fill launch), bit-reproducible run
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
constexpr int kRowMaxW = 1024;                       // a thread owns pixels w = tid + 256 i, i < 4
constexpr int kPerThread = kRowMaxW / kRowThreads;   // 4
constexpr int kRowChunk = 8;                         // channels staged per pass
constexpr int kBigBucket = 8;                        // buckets above this size are reduced by a warp
constexpr int kMaxBig = 64;                          // big buckets per row handled by warps (more: the owning thread loops)
using RowSort = cub::BlockRadixSort<unsigned int, kRowThreads, 2 * kPerThread, unsigned int>;

struct RowBwdArgs {
  const float* img;      // (N,C,H,W) source image of the warp
  const float* off;      // (N,1,H,W)
  const float* gout;     // mode 0/1: upstream gradient; layout per gout_cnhw.  mode 2: unused
  const float* aux;      // mode 1: seg (N,C,H,W) the warp is blended with;  mode 2: left (N,C,H,W)
  const float* att;      // mode 1: (N,1,H,W) blend weight
  const float* gwarp;    // mode 1: optional gradient w.r.t. the warped tensor itself (N,C,H,W), may be null
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
  // exclusive scan of one int per thread over the 256-thread CTA; returns the exclusive prefix, total in warp_sums[7]
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

// Shared-memory layout of one row's tap structure + the staged channel chunk (floats first, then ints, shorts, bytes).
struct RowSmem {
  float *twl, *twr, *ev, *gs, *is;
  int *offs, *cur, *wsum, *big;
  unsigned short *tx0, *tx1, *ew;
  unsigned char* tpass;
};
__host__ __device__ inline size_t row_smem_layout(int W, unsigned char* base, RowSmem* r) {
  const size_t W4 = (size_t)((W + 3) & ~3);
  size_t o = 0;
  auto take = [&](size_t bytes) { const size_t at = o; o += (bytes + 15) & ~(size_t)15; return at; };
  const size_t o_twl = take(4 * W4), o_twr = take(4 * W4), o_ev = take(8 * W4),
               // gs and is are contiguous; the region also hosts the radix sort's temporary storage (used before them)
               o_gs = take(8 * kRowChunk * W4 > sizeof(RowSort::TempStorage) ? 8 * kRowChunk * W4 : sizeof(RowSort::TempStorage)),
               o_is = o_gs + 4 * kRowChunk * W4, o_offs = take(4 * (W4 + 4)), o_cur = take(4 * W4), o_ws = take(64), o_big = take(4 * (kMaxBig + 4)),
               o_x0 = take(2 * W4), o_x1 = take(2 * W4), o_ew = take(4 * W4), o_ps = take(W4);
  if (r != nullptr) {
    r->twl = reinterpret_cast<float*>(base + o_twl), r->twr = reinterpret_cast<float*>(base + o_twr);
    r->ev = reinterpret_cast<float*>(base + o_ev), r->gs = reinterpret_cast<float*>(base + o_gs);
    r->is = reinterpret_cast<float*>(base + o_is), r->offs = reinterpret_cast<int*>(base + o_offs);
    r->cur = reinterpret_cast<int*>(base + o_cur), r->wsum = reinterpret_cast<int*>(base + o_ws);
    r->big = reinterpret_cast<int*>(base + o_big);
    r->tx0 = reinterpret_cast<unsigned short*>(base + o_x0), r->tx1 = reinterpret_cast<unsigned short*>(base + o_x1);
    r->ew = reinterpret_cast<unsigned short*>(base + o_ew), r->tpass = base + o_ps;
  }
  return o;
}

template <int kMode>
__global__ void __launch_bounds__(kRowThreads) warp_bwd_rows_kernel(const RowBwdArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  RowSmem sm;
  row_smem_layout(a.W, smem_raw, &sm);
  const int W = a.W, C = a.C;
  const int W4 = (W + 3) & ~3;
  const int tid = threadIdx.x, wid = tid >> 5, lane = tid & 31;
  const int64_t plane = (int64_t)a.H * W, total = (int64_t)a.N * plane;
  const bool need_img = a.gimg != nullptr;
  const bool need_is = a.goff != nullptr || kMode != 0;   // the image rows are needed for goff and to recompute the warp
  const float sc = kMode == 2 ? a.scale * (a.gloss != nullptr ? __ldg(a.gloss) : 1.f) : 0.f;
  // 16-byte copies need 16-byte aligned row starts in global memory
  const bool vec16 = (W & 3) == 0 && ((reinterpret_cast<uintptr_t>(a.img) | reinterpret_cast<uintptr_t>(a.gout)) & 15) == 0;

  for (int row = blockIdx.x; row < a.N * a.H; row += gridDim.x) {
    const int n = row / a.H, h = row % a.H;
    const int64_t roff = (int64_t)n * plane + (int64_t)h * W;   // offset of the row inside an (N,1,H,W) tensor
    // ---- A. taps of every pixel of the row ----
    for (int w = tid; w < W; w += kRowThreads) {
      const Taps t = make_taps(n, h, w, __ldg(a.off + roff + w), a.H, W, total);
      sm.tx0[w] = (unsigned short)(t.il - roff), sm.tx1[w] = (unsigned short)(t.ir - roff);   // exact below 2^24
      sm.twl[w] = t.wl, sm.twr[w] = t.wr;
      sm.tpass[w] = t.pass ? 1 : 0;
      sm.cur[w] = 0;
    }
    __syncthreads();
    if (need_img) {
      // ---- B. bucket sizes (integer atomics: deterministic) ----
      for (int w = tid; w < W; w += kRowThreads) {
        if (sm.twl[w] != 0.f) atomicAdd(&sm.cur[sm.tx0[w]], 1);
        if (sm.twr[w] != 0.f) atomicAdd(&sm.cur[sm.tx1[w]], 1);
      }
      __syncthreads();
      // ---- C. exclusive scan -> bucket starts ----
      {
        const int per = (W + kRowThreads - 1) / kRowThreads;
        const int b0 = tid * per;
        int local = 0;
        for (int i = 0; i < per; ++i)
          if (b0 + i < W) local += sm.cur[b0 + i];
        int run = block_exclusive_scan(local, sm.wsum, tid);
        for (int i = 0; i < per; ++i)
          if (b0 + i < W) {
            const int c = sm.cur[b0 + i];
            sm.offs[b0 + i] = run;
            sm.cur[b0 + i] = run;
            run += c;
          }
        if (tid == kRowThreads - 1) sm.offs[W] = sm.wsum[kRowWarps - 1];
      }
      __syncthreads();
      // ---- D. stable sort of the taps by destination: entries of bucket k land in [offs[k], offs[k+1]) in increasing
      // (source pixel, tap kind) order.  Thread t owns pixels 4t..4t+3 (blocked arrangement = source order). ----
      {
        typename RowSort::TempStorage& tmp = *reinterpret_cast<typename RowSort::TempStorage*>(sm.gs);   // gs/is are free here
        unsigned int keys[2 * kPerThread], vals[2 * kPerThread];
#pragma unroll
        for (int i = 0; i < kPerThread; ++i) {
          const int w = kPerThread * tid + i;
          const bool in = w < W;
          const float wl = in ? sm.twl[w] : 0.f, wr = in ? sm.twr[w] : 0.f;
          keys[2 * i] = wl != 0.f ? (unsigned)sm.tx0[w] : 0x7ffu;        // taps without weight sort behind every bucket
          keys[2 * i + 1] = wr != 0.f ? (unsigned)sm.tx1[w] : 0x7ffu;
          vals[2 * i] = (unsigned)(2 * w), vals[2 * i + 1] = (unsigned)(2 * w + 1);
        }
        RowSort(tmp).Sort(keys, vals, 0, 11);
        const int n_entries = sm.offs[W];
#pragma unroll
        for (int i = 0; i < 2 * kPerThread; ++i) {
          const int pos = 2 * kPerThread * tid + i;
          if (pos < n_entries) {
            const int w = (int)(vals[i] >> 1);
            sm.ew[pos] = (unsigned short)w;
            sm.ev[pos] = (vals[i] & 1u) ? sm.twr[w] : sm.twl[w];
          }
        }
        // big buckets -> list for the warp-cooperative pass (list order is irrelevant: every bucket is summed on its own)
        if (tid == 0) sm.big[kMaxBig] = 0;
        __syncthreads();
        for (int x = tid; x < W; x += kRowThreads)
          if (sm.offs[x + 1] - sm.offs[x] > kBigBucket) {
            const int slot = atomicAdd(&sm.big[kMaxBig], 1);
            if (slot < kMaxBig) sm.big[slot] = x;
          }
      }
      __syncthreads();
    }
    // ---- E. channel chunks: all 256 threads stage kRowChunk rows, then thread t owns pixels / destinations t + 256 i ----
    float dgo[kPerThread], dga[kPerThread];
    int b_[kPerThread], n_[kPerThread];
#pragma unroll
    for (int i = 0; i < kPerThread; ++i) {
      dgo[i] = 0.f, dga[i] = 0.f;
      const int x = tid + kRowThreads * i;
      b_[i] = (need_img && x < W) ? sm.offs[x] : 0;
      n_[i] = (need_img && x < W) ? sm.offs[x + 1] - b_[i] : 0;
    }
    const int n_big_all = need_img ? sm.big[kMaxBig] : 0;
    const int n_big = n_big_all < kMaxBig ? n_big_all : kMaxBig;
    const bool warps_take_big = n_big_all <= kMaxBig;     // otherwise (pathological) the owning threads loop themselves
    for (int c0 = 0; c0 < C; c0 += kRowChunk) {
      const int nc = C - c0 < kRowChunk ? C - c0 : kRowChunk;
      // stage the chunk's rows with cp.async: every copy of the chunk is in flight before the first one is waited for
      // (a load -> store loop left one or two loads per thread in flight and made the kernel latency-bound)
      {
        const int64_t crow = ((int64_t)n * C + c0) * plane + (int64_t)h * W;    // row of channel c0 in an (N,C,H,W) tensor
        const int64_t grow = a.gout_cnhw ? (int64_t)c0 * total + roff : crow;    // ... in gout
        const int64_t gstep = a.gout_cnhw ? total : plane;
        if (vec16) {
          const int Wq = W >> 2;
          for (int i = tid; i < nc * Wq; i += kRowThreads) {
            const int cc = i / Wq, q = i - cc * Wq;
            if (need_is) cp_async16(sm.is + cc * W4 + 4 * q, a.img + crow + (int64_t)cc * plane + 4 * q, true);
            if (kMode == 0) cp_async16(sm.gs + cc * W4 + 4 * q, a.gout + grow + (int64_t)cc * gstep + 4 * q, true);
          }
        } else {
          for (int i = tid; i < nc * W; i += kRowThreads) {
            const int cc = i / W, w = i - cc * W;
            if (need_is) cp_async4(sm.is + cc * W4 + w, a.img + crow + (int64_t)cc * plane + w);
            if (kMode == 0) cp_async4(sm.gs + cc * W4 + w, a.gout + grow + (int64_t)cc * gstep + w);
          }
        }
        cp_async_commit();
        cp_async_wait<0>();
      }
      __syncthreads();
      if (kMode != 0) {
        // recompute the warp, derive the effective upstream gradient of the warped tensor and the consumer's own
        // gradients.  The global operands of the thread's pixels x channels are fetched first (independent loads).
#pragma unroll
        for (int i = 0; i < kPerThread; ++i) {
          const int w = tid + kRowThreads * i;
          if (w < W) {
            const int x0 = sm.tx0[w], x1 = sm.tx1[w];
            const float wl = sm.twl[w], wr = sm.twr[w];
            const float at = kMode == 1 ? __ldg(a.att + roff + w) : 0.f;
            const float mk = (kMode == 2 && a.mask_pos && !(__ldg(a.off + roff + w) < 0.f)) ? 0.f : 1.f;
            const int64_t c00 = ((int64_t)n * C + c0) * plane + (int64_t)h * W + w;
            float gq[kRowChunk], aq[kRowChunk], wq[kRowChunk];
#pragma unroll
            for (int cc = 0; cc < kRowChunk; ++cc) {
              const bool ok = cc < nc;
              aq[cc] = ok ? __ldg(a.aux + c00 + (int64_t)cc * plane) : 0.f;
              gq[cc] = (ok && kMode == 1) ? __ldg(a.gout + c00 + (int64_t)cc * plane) : 0.f;
              wq[cc] = (ok && kMode == 1 && a.gwarp != nullptr) ? __ldg(a.gwarp + c00 + (int64_t)cc * plane) : 0.f;
            }
#pragma unroll
            for (int cc = 0; cc < kRowChunk; ++cc) {
              if (cc < nc) {
                const int64_t coff = c00 + (int64_t)cc * plane;
                const float wv = __fadd_rn(__fmul_rn(wl, sm.is[cc * W4 + x0]), __fmul_rn(wr, sm.is[cc * W4 + x1]));
                float ge;
                if (kMode == 1) {
                  dga[i] = fmaf(gq[cc], wv - aq[cc], dga[i]);
                  a.gaux[coff] = (1.f - at) * gq[cc];
                  ge = at * gq[cc] + wq[cc];
                } else {
                  const float d = sc * (wv * mk - aq[cc]);
                  if (a.gaux != nullptr) a.gaux[coff] = -d;
                  ge = mk * d;
                }
                sm.gs[cc * W4 + w] = ge;
              }
            }
          }
        }
        __syncthreads();
      }
#pragma unroll
      for (int i = 0; i < kPerThread; ++i) {
        const int x = tid + kRowThreads * i;
        if (x < W) {
          if (a.goff != nullptr) {
            const int x0 = sm.tx0[x], x1 = sm.tx1[x];
            for (int cc = 0; cc < nc; ++cc) dgo[i] = fmaf(sm.gs[cc * W4 + x], sm.is[cc * W4 + x1] - sm.is[cc * W4 + x0], dgo[i]);
          }
          if (need_img && !(warps_take_big && n_[i] > kBigBucket)) {
            // bucket of destination x: entries b .. b+n-1; the first four inline and branch-free (smooth disparities give
            // two per bucket, random ones rarely more than four), the rest in a loop
            const int b = b_[i], nb = n_[i];
            float v_[4];
            int w_[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              v_[j] = nb > j ? sm.ev[b + j] : 0.f;
              w_[j] = nb > j ? (int)sm.ew[b + j] : 0;
            }
            float* o = a.gimg + ((int64_t)n * C + c0) * plane + (int64_t)h * W + x;
            for (int cc = 0; cc < nc; ++cc, o += plane) {
              const float* g = sm.gs + cc * W4;
              // selects, not multiplications by a zero weight: an absent entry must not turn an inf/nan of g[0] into a nan
              float acc = nb > 0 ? __fmul_rn(v_[0], g[w_[0]]) : 0.f;
              acc = nb > 1 ? __fadd_rn(acc, __fmul_rn(v_[1], g[w_[1]])) : acc;
              acc = nb > 2 ? __fadd_rn(acc, __fmul_rn(v_[2], g[w_[2]])) : acc;
              acc = nb > 3 ? __fadd_rn(acc, __fmul_rn(v_[3], g[w_[3]])) : acc;
              for (int e = b + 4; e < b + nb; ++e) acc = __fadd_rn(acc, __fmul_rn(sm.ev[e], g[sm.ew[e]]));
              st_cs(o, acc);
            }
          }
        }
      }
      if (need_img && warps_take_big) {
        // big buckets: warp `wid` takes buckets wid, wid+8, ...; lane l adds entries b+l, b+l+32, ... in order, then a
        // fixed shuffle tree -> the same order every run
        for (int j = wid; j < n_big; j += kRowWarps) {
          const int x = sm.big[j], b = sm.offs[x], nb = sm.offs[x + 1] - b;
          for (int cc = 0; cc < nc; ++cc) {
            const float* g = sm.gs + cc * W4;
            float acc = 0.f;
            for (int e = b + lane; e < b + nb; e += 32) acc = __fadd_rn(acc, __fmul_rn(sm.ev[e], g[sm.ew[e]]));
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
            if (lane == 0) a.gimg[((int64_t)n * C + c0 + cc) * plane + (int64_t)h * W + x] = acc;
          }
        }
      }
      __syncthreads();   // gs / is are overwritten by the next chunk
    }
    // ---- F. per-pixel results ----
#pragma unroll
    for (int i = 0; i < kPerThread; ++i) {
      const int w = tid + kRowThreads * i;
      if (w < W) {
        if (a.goff != nullptr) a.goff[roff + w] = sm.tpass[w] ? dgo[i] : 0.f;
        if (kMode == 1) a.gatt[roff + w] = dga[i];
      }
    }
    __syncthreads();   // shared memory is reused by the next row
  }
}

size_t rows_smem_bytes(int W) { return row_smem_layout(W, nullptr, nullptr); }

// The row kernel needs exact float32 flat indices (taps inside the row) and W <= 1024 (register partials).
bool rows_ok(int N, int H, int W) {
  return (int64_t)N * H * W < (1ll << 24) && W <= kRowMaxW && rows_smem_bytes(W) <= 200 * 1024;
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
