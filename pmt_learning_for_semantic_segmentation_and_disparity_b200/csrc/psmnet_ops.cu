// psmnet_ops.cu -- PSMNet concat cost volume, disparityregression and fused soft-argmin.
//   concat volume : models_psmnet/stackhourglass.py:110-119, matchshifted submodule.py:45-54
//   dispreg       : models_psmnet/submodule.py:56-64
//   soft-argmin   : F.softmax(dim=1) + disparityregression, stackhourglass.py:142-155
// All four are HBM-bound streaming kernels: every input element is read once and every output
// element written once with 128-bit accesses; grids are sized in multiples of the SM count.
#include <stdlib.h>

#include "common.cuh"

namespace pmt {
namespace {

// ---------------------------------------------------------------------------------------------
// concat volume forward: one CTA walks (b, c2, h) source rows; the source row is staged once in
// shared memory and re-emitted D times (shifted for the target half), so global reads are 1/D of
// the writes and every write is a full 16-byte streaming store.
// ---------------------------------------------------------------------------------------------
template <bool kVec>
__global__ void __launch_bounds__(128) concat_fwd_kernel(const float* __restrict__ ref,
                                                         const float* __restrict__ tgt,
                                                         float* __restrict__ cost, int B, int C, int D,
                                                         int H, int W, int d0) {
  extern __shared__ float srow[];
  const int64_t rows = (int64_t)B * 2 * C * H;
  for (int64_t r = blockIdx.x; r < rows; r += gridDim.x) {
    const int h = (int)(r % H);
    const int c2 = (int)((r / H) % (2 * C));
    const int b = (int)(r / ((int64_t)H * 2 * C));
    const bool is_tgt = c2 >= C;
    const float* src = (is_tgt ? tgt : ref) + (((int64_t)b * C + (is_tgt ? c2 - C : c2)) * H + h) * (int64_t)W;
    __syncthreads();  // previous row fully consumed
    if (kVec) {
      for (int w = threadIdx.x * 4; w < W; w += blockDim.x * 4)
        *reinterpret_cast<float4*>(srow + w) = __ldg(reinterpret_cast<const float4*>(src + w));
    } else {
      for (int w = threadIdx.x; w < W; w += blockDim.x) srow[w] = __ldg(src + w);
    }
    __syncthreads();
    float* dst0 = cost + ((((int64_t)b * 2 * C + c2) * D) * H + h) * (int64_t)W;
    const int64_t dstride = (int64_t)H * W;
    if (kVec) {
      const int nv = W >> 2;
      for (int e = threadIdx.x; e < D * nv; e += blockDim.x) {
        const int i = e / nv, w = (e - i * nv) * 4;
        const int d = d0 + i;
        const int sh = is_tgt ? d : 0;
        float4 v;
        v.x = (w + 0 >= d) ? srow[w + 0 - sh] : 0.f;
        v.y = (w + 1 >= d) ? srow[w + 1 - sh] : 0.f;
        v.z = (w + 2 >= d) ? srow[w + 2 - sh] : 0.f;
        v.w = (w + 3 >= d) ? srow[w + 3 - sh] : 0.f;
        st_cs4(dst0 + i * dstride + w, v);
      }
    } else {
      for (int e = threadIdx.x; e < D * W; e += blockDim.x) {
        const int i = e / W, w = e - i * W;
        const int d = d0 + i;
        dst0[i * dstride + w] = (w >= d) ? srow[w - (is_tgt ? d : 0)] : 0.f;
      }
    }
  }
}

// concat volume backward: deterministic gather; one thread per (b,c2,h,w) sums its D planes in
// descending plane order (the order autograd accumulates the reference's slice assignments).
__global__ void __launch_bounds__(256) concat_bwd_kernel(const float* __restrict__ gcost,
                                                         float* __restrict__ gref,
                                                         float* __restrict__ gtgt, int B, int C, int D,
                                                         int H, int W, int d0, int64_t total) {
  const int64_t dstride = (int64_t)H * W;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int w = (int)(idx % W);
    int64_t t = idx / W;
    const int h = (int)(t % H);
    t /= H;
    const int c2 = (int)(t % (2 * C));
    const int b = (int)(t / (2 * C));
    const bool is_tgt = c2 >= C;
    const float* g = gcost + ((((int64_t)b * 2 * C + c2) * D) * H + h) * (int64_t)W;
    float acc = 0.f;
    if (!is_tgt) {
#pragma unroll 8
      for (int i = D - 1; i >= 0; --i)
        if (w >= d0 + i) acc += __ldg(g + i * dstride + w);
      gref[(((int64_t)b * C + c2) * H + h) * (int64_t)W + w] = acc;
    } else {
#pragma unroll 8
      for (int i = D - 1; i >= 0; --i)
        if (w + d0 + i < W) acc += __ldg(g + i * dstride + w + d0 + i);
      gtgt[(((int64_t)b * C + (c2 - C)) * H + h) * (int64_t)W + w] = acc;
    }
  }
}

// Vectorised variant (W % 4 == 0, d0 % 4 == 0, 16-byte aligned pointers): one thread per 4 consecutive columns.
// Reference half: aligned 128-bit loads, per-element mask.  Target half: the 4 needed elements g[i][w+s..w+s+3]
// (s = d0+i) straddle two aligned vectors unless s % 4 == 0; planes are walked 4 at a time so the in-vector offset
// is a compile-time constant (7 vector loads per 4 planes; the second vector of a thread is the first vector of its
// neighbour, so it hits L1).  Planes are still added in descending order, like autograd does.
__device__ __forceinline__ float4 ldg4_guard(const float* row, int col, int W) {
  return col < W ? __ldg(reinterpret_cast<const float4*>(row + col)) : make_float4(0.f, 0.f, 0.f, 0.f);
}
template <int R>
__device__ __forceinline__ float4 shift_pick(const float4& a, const float4& b) {   // elements R..R+3 of (a,b)
  if (R == 0) return a;
  if (R == 1) return make_float4(a.y, a.z, a.w, b.x);
  if (R == 2) return make_float4(a.z, a.w, b.x, b.y);
  return make_float4(a.w, b.x, b.y, b.z);
}

__global__ void __launch_bounds__(256) concat_bwd_vec_kernel(const float* __restrict__ gcost,
                                                             float* __restrict__ gref, float* __restrict__ gtgt,
                                                             int B, int C, int D, int H, int W, int d0,
                                                             int64_t total4) {
  const int64_t dstride = (int64_t)H * W;
  const int W4 = W >> 2;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total4;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int w = (int)(idx % W4) * 4;
    int64_t t = idx / W4;
    const int h = (int)(t % H);
    t /= H;
    const int c2 = (int)(t % (2 * C));
    const int b = (int)(t / (2 * C));
    const bool is_tgt = c2 >= C;
    const float* g = gcost + ((((int64_t)b * 2 * C + c2) * D) * H + h) * (int64_t)W;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    auto add_masked = [&](const float4& v, int s) {   // plane with shift s contributes where the index is in range
      const float e[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const bool ok = is_tgt ? (w + u + s < W) : (w + u >= s);
        if (ok) acc[u] += e[u];
      }
    };
    int i = D - 1;
    for (; (i & 3) != 3; --i) {   // tail planes (D % 4 != 0): generic offsets
      const int s = d0 + i;
      const float* row = g + i * dstride;
      float4 v;
      if (!is_tgt) {
        v = ldg4_guard(row, w, W);
      } else {
        const int base = (w + s) & ~3, r = (w + s) & 3;
        const float4 a0 = ldg4_guard(row, base, W), a1 = ldg4_guard(row, base + 4, W);
        v = r == 0 ? a0 : r == 1 ? shift_pick<1>(a0, a1) : r == 2 ? shift_pick<2>(a0, a1) : shift_pick<3>(a0, a1);
      }
      add_masked(v, s);
    }
    for (; i >= 3; i -= 4) {      // planes i, i-1, i-2, i-3 with i % 4 == 3
      const int s0 = d0 + i - 3;  // multiple of 4
      const float* row = g + (int64_t)(i - 3) * dstride;
      if (!is_tgt) {
        float4 v[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) v[r] = ldg4_guard(row + r * dstride, w, W);
#pragma unroll
        for (int r = 3; r >= 0; --r) add_masked(v[r], s0 + r);
      } else {
        const int base = w + s0;
        float4 a[4], bb[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          a[r] = ldg4_guard(row + r * dstride, base, W);
          if (r > 0) bb[r] = ldg4_guard(row + r * dstride, base + 4, W);
        }
        add_masked(shift_pick<3>(a[3], bb[3]), s0 + 3);
        add_masked(shift_pick<2>(a[2], bb[2]), s0 + 2);
        add_masked(shift_pick<1>(a[1], bb[1]), s0 + 1);
        add_masked(a[0], s0);
      }
    }
    float* o = (is_tgt ? gtgt + (((int64_t)b * C + (c2 - C)) * H + h) * (int64_t)W
                       : gref + (((int64_t)b * C + c2) * H + h) * (int64_t)W) + w;
    *reinterpret_cast<float4*>(o) = make_float4(acc[0], acc[1], acc[2], acc[3]);
  }
}

// ---------------------------------------------------------------------------------------------
// soft-argmin / disparityregression.  kV pixels per thread (4 with 128-bit accesses when H*W%4==0),
// planes walked with stride H*W; 8 planes are loaded before use so each thread keeps 8 independent
// 16-byte loads in flight.
// ---------------------------------------------------------------------------------------------
template <int kV>
struct Vec;
template <>
struct Vec<4> {
  using T = float4;
};
template <>
struct Vec<1> {
  using T = float;
};

template <int kV>
__device__ __forceinline__ void load_vec(const float* p, float (&v)[kV]) {
  if constexpr (kV == 4) {
    const float4 t = __ldcs(reinterpret_cast<const float4*>(p));
    v[0] = t.x, v[1] = t.y, v[2] = t.z, v[3] = t.w;
  } else {
    v[0] = __ldcs(p);
  }
}
template <int kV>
__device__ __forceinline__ void store_vec(float* p, const float (&v)[kV]) {
  if constexpr (kV == 4) {
    st_cs4(p, make_float4(v[0], v[1], v[2], v[3]));
  } else {
    st_cs(p, v[0]);
  }
}

constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;
constexpr int kDU = 8;  // planes per unrolled batch

template <int kV>
__global__ void __launch_bounds__(256) softargmin_fwd_kernel(const float* __restrict__ cost,
                                                             float* __restrict__ out,
                                                             float* __restrict__ lse, int D,
                                                             int64_t plane, int64_t ngroups) {
  const int64_t gpp = plane / kV;  // pixel groups per batch item
  for (int64_t gidx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; gidx < ngroups;
       gidx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = gidx / gpp, k = (gidx % gpp) * kV;
    const float* base = cost + b * D * plane + k;
    float m[kV], s[kV], t[kV];
#pragma unroll
    for (int v = 0; v < kV; ++v) m[v] = -INFINITY, s[v] = 0.f, t[v] = 0.f;
    // blocks start at different planes and wrap (see dispreg_fwd_kernel): no lockstep walk over the planes
    const int nb_ = (D + kDU - 1) / kDU;
    const int boff_ = (int)(blockIdx.x % 4) * (nb_ / 4);
    for (int t_ = 0; t_ < nb_; ++t_) {
      const int d0 = ((t_ + boff_) % nb_) * kDU;
      float x[kDU][kV];
#pragma unroll
      for (int j = 0; j < kDU; ++j) {
        if (d0 + j < D) {
          load_vec<kV>(base + (int64_t)(d0 + j) * plane, x[j]);
        } else {
#pragma unroll
          for (int v = 0; v < kV; ++v) x[j][v] = -INFINITY;
        }
      }
#pragma unroll
      for (int v = 0; v < kV; ++v) {
        float bm = x[0][v];
#pragma unroll
        for (int j = 1; j < kDU; ++j) bm = fmaxf(bm, x[j][v]);
        const float mn = fmaxf(m[v], bm * kLog2e);
        const float sc = exp2f(m[v] - mn);  // 0 on the first batch (m = -inf)
        float ss = s[v] * sc, tt = t[v] * sc;
#pragma unroll
        for (int j = 0; j < kDU; ++j) {
          const float e = exp2f(fmaf(x[j][v], kLog2e, -mn));  // exp(x - max); 0 for padded planes
          ss += e;
          tt = fmaf(e, (float)(d0 + j), tt);
        }
        m[v] = mn, s[v] = ss, t[v] = tt;
      }
    }
    float o[kV], l[kV];
#pragma unroll
    for (int v = 0; v < kV; ++v) {
      o[v] = t[v] / s[v];
      l[v] = (m[v] + log2f(s[v])) * kLn2;
    }
    store_vec<kV>(out + b * plane + k, o);
    if (lse != nullptr) store_vec<kV>(lse + b * plane + k, l);
  }
}

template <int kV>
__global__ void __launch_bounds__(256) softargmin_bwd_kernel(const float* __restrict__ cost,
                                                             const float* __restrict__ out,
                                                             const float* __restrict__ lse,
                                                             const float* __restrict__ gout,
                                                             float* __restrict__ gcost, int D,
                                                             int64_t plane, int64_t ngroups) {
  const int64_t gpp = plane / kV;
  for (int64_t gidx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; gidx < ngroups;
       gidx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = gidx / gpp, k = (gidx % gpp) * kV;
    float o[kV], l[kV], g[kV];
    load_vec<kV>(out + b * plane + k, o);
    load_vec<kV>(lse + b * plane + k, l);
    load_vec<kV>(gout + b * plane + k, g);
#pragma unroll
    for (int v = 0; v < kV; ++v) l[v] *= kLog2e;
    const float* cbase = cost + b * D * plane + k;
    float* gbase = gcost + b * D * plane + k;
    // (walking the planes in a per-block staggered order, which helps the forward kernels, was measured SLOWER here:
    // 147 vs 139 us -- this kernel also writes a plane per plane read)
    for (int d0 = 0; d0 < D; d0 += kDU) {
      float x[kDU][kV];
#pragma unroll
      for (int j = 0; j < kDU; ++j)
        if (d0 + j < D) load_vec<kV>(cbase + (int64_t)(d0 + j) * plane, x[j]);
#pragma unroll
      for (int j = 0; j < kDU; ++j) {
        if (d0 + j < D) {
          float r[kV];
#pragma unroll
          for (int v = 0; v < kV; ++v) {
            const float p = exp2f(fmaf(x[j][v], kLog2e, -l[v]));
            r[v] = g[v] * p * ((float)(d0 + j) - o[v]);
          }
          store_vec<kV>(gbase + (int64_t)(d0 + j) * plane, r);
        }
      }
    }
  }
}

template <int kV>
__global__ void __launch_bounds__(256) dispreg_fwd_kernel(const float* __restrict__ x,
                                                          float* __restrict__ out, int D,
                                                          int64_t plane, int64_t ngroups) {
  constexpr int kU = kDU;   // 16 planes in flight was measured slower (180 vs 100 us: register pressure)
  const int64_t gpp = plane / kV;
  for (int64_t gidx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; gidx < ngroups;
       gidx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = gidx / gpp, k = (gidx % gpp) * kV;
    const float* base = x + b * D * plane + k;
    float acc[kV];
#pragma unroll
    for (int v = 0; v < kV; ++v) acc[v] = 0.f;
    // blocks start at different planes (then wrap): threads of different blocks do not walk the planes in lockstep,
    // which spreads the concurrent requests over more DRAM pages / L2 slices
    const int nb = (D + kU - 1) / kU;
    const int boff = (int)(blockIdx.x % 4) * (nb / 4);
    for (int t = 0; t < nb; ++t) {
      const int d0 = ((t + boff) % nb) * kU;
      float xv[kU][kV];
#pragma unroll
      for (int j = 0; j < kU; ++j) {
        if (d0 + j < D) {
          load_vec<kV>(base + (int64_t)(d0 + j) * plane, xv[j]);
        } else {
#pragma unroll
          for (int v = 0; v < kV; ++v) xv[j][v] = 0.f;
        }
      }
      // the reference multiplies then sums (x*disp).sum(1): keep the products un-fused, in d order
#pragma unroll
      for (int j = 0; j < kU; ++j)
#pragma unroll
        for (int v = 0; v < kV; ++v) acc[v] = __fadd_rn(acc[v], __fmul_rn(xv[j][v], (float)(d0 + j)));
    }
    store_vec<kV>(out + b * plane + k, acc);
  }
}

template <int kV>
__global__ void __launch_bounds__(256) dispreg_bwd_kernel(const float* __restrict__ gout,
                                                          float* __restrict__ gx, int D,
                                                          int64_t plane, int64_t ngroups) {
  const int64_t gpp = plane / kV;
  for (int64_t gidx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; gidx < ngroups;
       gidx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = gidx / gpp, k = (gidx % gpp) * kV;
    float g[kV];
    load_vec<kV>(gout + b * plane + k, g);
    float* base = gx + b * D * plane + k;
    for (int d = 0; d < D; ++d) {
      float r[kV];
#pragma unroll
      for (int v = 0; v < kV; ++v) r[v] = (float)d * g[v];
      store_vec<kV>(base + (int64_t)d * plane, r);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// f1: trilinear upsample fused into the soft-argmin.  PSMNet does
//   cost = F.upsample(cost3, [maxdisp, H, W], mode='trilinear'); pred = disparityregression(softmax(cost, 1))
// (models_psmnet/stackhourglass.py:149-155): the (B,maxdisp,H,W) tensor (100.7 MB per pair) exists only to be
// reduced again.  Here one thread owns an output pixel, interpolates the 4 spatial taps of every low-res plane on
// the fly (source indices and weights exactly as ATen's align_corners=False rule: src = scale*(dst+0.5)-0.5
// clamped at 0, scale = in/out), blends consecutive planes along d and feeds the online softmax -- the upsampled
// volume never touches HBM (reads 1.57 MB instead of 100.7 MB per pair).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float fast_exp2(float x) {   // MUFU.EX2; exp2(-inf) = 0
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

struct UpArgs {
  int B, Dq, Hq, Wq, D, H, W;
  float sd, sh, sw;  // in/out scale per axis
};

__device__ __forceinline__ void src_index(float scale, int dst, int in_size, int& i0, int& i1, float& l0, float& l1) {
  float src = scale * ((float)dst + 0.5f) - 0.5f;
  src = src < 0.f ? 0.f : src;
  i0 = (int)src;
  if (i0 > in_size - 1) i0 = in_size - 1;
  i1 = i0 + (i0 < in_size - 1 ? 1 : 0);
  l1 = src - (float)i0;
  l0 = 1.f - l1;
}

__global__ void __launch_bounds__(256) upsample_softargmin_fwd_kernel(const float* __restrict__ lowres,
                                                                      float* __restrict__ out,
                                                                      float* __restrict__ lse, UpArgs a,
                                                                      int64_t total) {
  // per-CTA table of the disparity-axis source planes and weights (identical for every pixel)
  extern __shared__ float up_tab[];
  int* tab_i0 = reinterpret_cast<int*>(up_tab);
  float* tab_l1 = up_tab + a.D;
  for (int d = threadIdx.x; d < a.D; d += blockDim.x) {
    int i0, i1;
    float l0, l1;
    src_index(a.sd, d, a.Dq, i0, i1, l0, l1);
    tab_i0[d] = i0;
    tab_l1[d] = (i1 == i0) ? 0.f : l1;  // at the last plane both taps coincide
  }
  __syncthreads();
  const int64_t qplane = (int64_t)a.Hq * a.Wq;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int w = (int)(idx % a.W);
    const int h = (int)((idx / a.W) % a.H);
    const int b = (int)(idx / ((int64_t)a.W * a.H));
    int h0, h1, w0, w1;
    float hl0, hl1, wl0, wl1;
    src_index(a.sh, h, a.Hq, h0, h1, hl0, hl1);
    src_index(a.sw, w, a.Wq, w0, w1, wl0, wl1);
    const float* base = lowres + (int64_t)b * a.Dq * qplane;
    const int o00 = h0 * a.Wq + w0, o01 = h0 * a.Wq + w1, o10 = h1 * a.Wq + w0, o11 = h1 * a.Wq + w1;
    auto plane = [&](int dq) {
      const float* p = base + (int64_t)dq * qplane;
      return hl0 * (wl0 * __ldg(p + o00) + wl1 * __ldg(p + o01)) + hl1 * (wl0 * __ldg(p + o10) + wl1 * __ldg(p + o11));
    };
    int cur = 0;
    float s_cur = plane(0), s_nxt = plane(a.Dq > 1 ? 1 : 0);
    float m = -INFINITY, s = 0.f, t = 0.f;
    for (int d0 = 0; d0 < a.D; d0 += kDU) {
      float x[kDU];
#pragma unroll
      for (int j = 0; j < kDU; ++j) {
        const int d = d0 + j;
        if (d < a.D) {
          const int i0 = tab_i0[d];
          const float l1 = tab_l1[d];
          while (cur < i0) {  // advance the two-plane window
            ++cur;
            s_cur = s_nxt;
            s_nxt = plane(cur + 1 < a.Dq ? cur + 1 : cur);
          }
          x[j] = (1.f - l1) * s_cur + l1 * s_nxt;
        } else {
          x[j] = -INFINITY;
        }
      }
      float bm = x[0];
#pragma unroll
      for (int j = 1; j < kDU; ++j) bm = fmaxf(bm, x[j]);
      const float mn = fmaxf(m, bm * kLog2e);
      const float sc = exp2f(m - mn);
      float ss = s * sc, tt = t * sc;
#pragma unroll
      for (int j = 0; j < kDU; ++j) {
        const float e = exp2f(fmaf(x[j], kLog2e, -mn));
        ss += e;
        tt = fmaf(e, (float)(d0 + j), tt);
      }
      m = mn, s = ss, t = tt;
    }
    out[idx] = t / s;
    if (lse != nullptr) lse[idx] = (m + log2f(s)) * kLn2;
  }
}

// Tiled variant: a CTA owns a kUpTH x kUpTW block of output pixels and stages the low-res footprint of that block
// (every plane; a few rows x a few columns) in shared memory, so the 4 spatial taps per plane are LDS (mostly
// broadcast: 4 neighbouring pixels share their taps) instead of scattered global loads.  Same arithmetic, same order.
constexpr int kUpTH = 4, kUpTW = 64;

// kBwd = false: forward (writes out, lse).  kBwd = true: first half of the backward -- out / lse are INPUTS, and with
// g_x[d] = gout * p_d * (d - out) the kernel writes U[b,q,h,w] = sum_d g_x[d] * Wd(d -> q), the gradient w.r.t. the
// bilinearly sampled source planes (Wd = the disparity-axis interpolation weights); upsample_adjoint2d_kernel then
// applies the transposed spatial interpolation.
template <bool kBwd>
__global__ void __launch_bounds__(kUpTH * kUpTW) upsample_softargmin_tiled_kernel(const float* __restrict__ lowres,
                                                                                   float* __restrict__ out,
                                                                                   float* __restrict__ lse, UpArgs a,
                                                                                   int fr_max, int fc_max,
                                                                                   const float* __restrict__ gout,
                                                                                   float* __restrict__ U) {
  extern __shared__ float up_smem[];
  int* tab_beg = reinterpret_cast<int*>(up_smem);   // [Dq+1]: first output plane whose lower source plane is >= q
  float* tab_l1 = up_smem + a.Dq + 1;                // [D]: weight of the upper source plane
  float* tile = up_smem + a.Dq + 1 + a.D;            // [Dq][fr][fc]
  const int tid = threadIdx.x;
  const int h_first = blockIdx.y * kUpTH, w_first = blockIdx.x * kUpTW, b = blockIdx.z;
  const int h_last = min(h_first + kUpTH, a.H) - 1, w_last = min(w_first + kUpTW, a.W) - 1;
  int r0, r1, c0, c1, tmp;
  float f0, f1;
  src_index(a.sh, h_first, a.Hq, r0, tmp, f0, f1);
  src_index(a.sh, h_last, a.Hq, tmp, r1, f0, f1);
  src_index(a.sw, w_first, a.Wq, c0, tmp, f0, f1);
  src_index(a.sw, w_last, a.Wq, tmp, c1, f0, f1);
  const int fr = r1 - r0 + 1, fc = c1 - c0 + 1;   // <= fr_max, fc_max (host-side bound)
  for (int d = tid; d <= a.D; d += blockDim.x) {
    int i0 = a.Dq, i1 = a.Dq, ip = -1;
    float l0, l1 = 0.f;
    if (d < a.D) src_index(a.sd, d, a.Dq, i0, i1, l0, l1);
    if (d > 0) {
      int j1;
      float m0, m1;
      src_index(a.sd, d - 1, a.Dq, ip, j1, m0, m1);
    }
    if (d < a.D) tab_l1[d] = (i1 == i0) ? 0.f : l1;   // at the last plane both taps coincide
    for (int q = ip + 1; q <= i0 && q <= a.Dq; ++q) tab_beg[q] = d;   // i0 is non-decreasing in d
  }
  const int64_t qplane = (int64_t)a.Hq * a.Wq;
  const float* base = lowres + (int64_t)b * a.Dq * qplane;
  const int per_plane = fr * fc;
  for (int i = tid; i < a.Dq * per_plane; i += blockDim.x) {
    const int q = i / per_plane, rem = i - q * per_plane;
    const int r = rem / fc, c = rem - r * fc;
    tile[i] = __ldg(base + q * qplane + (int64_t)(r0 + r) * a.Wq + (c0 + c));
  }
  __syncthreads();
  const int w = w_first + (tid % kUpTW), h = h_first + (tid / kUpTW);
  if (w >= a.W || h >= a.H) return;
  int h0, h1, w0, w1;
  float hl0, hl1, wl0, wl1;
  src_index(a.sh, h, a.Hq, h0, h1, hl0, hl1);
  src_index(a.sw, w, a.Wq, w0, w1, wl0, wl1);
  const float* p00 = tile + (h0 - r0) * fc + (w0 - c0);
  const int d01 = w1 - w0, d10 = (h1 - h0) * fc;
  // bilinear sample of low-res plane q at this pixel, pre-multiplied by log2(e) (the soft-max runs in base 2)
  auto plane = [&](int q) {
    const float* p = p00 + q * per_plane;
    return kLog2e * (hl0 * (wl0 * p[0] + wl1 * p[d01]) + hl1 * (wl0 * p[d10] + wl1 * p[d10 + d01]));
  };
  const int64_t idx = ((int64_t)b * a.H + h) * a.W + w;
  if (kBwd) {
    const float g = __ldg(gout + idx), o = out[idx], l2 = lse[idx] * kLog2e;
    float* up = U + ((int64_t)b * a.Dq * a.H + h) * (int64_t)a.W + w;
    const int64_t ustride = (int64_t)a.H * a.W;
    float s_cur = plane(0), carry = 0.f;
    float df = (float)tab_beg[0];
    for (int q = 0; q < a.Dq; ++q) {
      const float s_nxt = plane(q + 1 < a.Dq ? q + 1 : q);
      const int dend = tab_beg[q + 1];
      const float c = s_cur - l2, diff = s_nxt - s_cur;
      float a0 = 0.f, a1 = 0.f;
      for (int d = tab_beg[q]; d < dend; ++d) {
        const float l1 = tab_l1[d];
        const float gx = g * fast_exp2(fmaf(l1, diff, c)) * (df - o);   // gout * p_d * (d - pred)
        a1 = fmaf(gx, l1, a1);
        a0 += gx - gx * l1;
        df += 1.f;
      }
      up[q * ustride] = carry + a0;   // plane q: lower tap of its own group + upper tap of the previous group
      carry = a1;
      s_cur = s_nxt;
    }
    return;
  }
  // Output planes d in [beg[q], beg[q+1]) blend source planes q and q+1: x_d = s_q + l1_d * (s_{q+1} - s_q), a convex
  // combination, so max(s_q, s_{q+1}) bounds every x_d of the group: the running maximum is updated once per source
  // plane (one rescale) instead of once per output plane.
  float s_cur = plane(0);
  float m = -INFINITY, s = 0.f, t = 0.f;
  float df = (float)tab_beg[0];
  for (int q = 0; q < a.Dq; ++q) {
    const float s_nxt = plane(q + 1 < a.Dq ? q + 1 : q);
    const int dend = tab_beg[q + 1];
    int d = tab_beg[q];
    if (d < dend) {
      const float mn = fmaxf(m, fmaxf(s_cur, s_nxt));
      const float sc = fast_exp2(m - mn);
      s *= sc, t *= sc, m = mn;
      const float c = s_cur - mn, diff = s_nxt - s_cur;
      for (; d < dend; ++d) {
        const float e = fast_exp2(fmaf(tab_l1[d], diff, c));
        s += e;
        t = fmaf(e, df, t);
        df += 1.f;
      }
    }
    s_cur = s_nxt;
  }
  out[idx] = t / s;
  if (lse != nullptr) lse[idx] = (m + log2f(s)) * kLn2;
}

// Second half of the backward: glow[b,q,hq,wq] = sum_{h,w} U[b,q,h,w] * Wh(h -> hq) * Ww(w -> wq), the adjoint of the
// bilinear (align_corners=False) up-sampling of one plane.  One thread per low-res element gathers from the window of
// output rows / columns whose two taps include it (deterministic, no atomics).
constexpr int kAdjMaxWin = 40;
__device__ __forceinline__ int adjoint_window(float scale, int q, int in_size, int out_size, float* wgt, int& lo) {
  // outputs o with src(o) in (q-1, q+1): o in ((q-0.5)/scale - 0.5, (q+1.5)/scale - 0.5); one more on each side for rounding
  int o_lo = (int)floorf(((float)q - 0.5f) / scale - 0.5f) - 1;
  int o_hi = (int)ceilf(((float)q + 1.5f) / scale - 0.5f) + 1;
  o_lo = o_lo < 0 ? 0 : o_lo;
  o_hi = o_hi > out_size - 1 ? out_size - 1 : o_hi;
  lo = o_lo;
  int n = o_hi - o_lo + 1;
  n = n > kAdjMaxWin ? kAdjMaxWin : n;   // the launcher guarantees the window fits
  for (int i = 0; i < n; ++i) {
    int i0, i1;
    float l0, l1;
    src_index(scale, o_lo + i, in_size, i0, i1, l0, l1);
    wgt[i] = (i0 == q ? l0 : 0.f) + (i1 == q ? l1 : 0.f);
  }
  return n;
}

__global__ void __launch_bounds__(256) upsample_adjoint2d_kernel(const float* __restrict__ U, float* __restrict__ glow,
                                                                 UpArgs a, int64_t total) {
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int wq = (int)(idx % a.Wq);
    const int hq = (int)((idx / a.Wq) % a.Hq);
    const int64_t bq = idx / ((int64_t)a.Wq * a.Hq);   // b * Dq + q
    float wh[kAdjMaxWin], ww[kAdjMaxWin];
    int h_lo, w_lo;
    const int nh = adjoint_window(a.sh, hq, a.Hq, a.H, wh, h_lo);
    const int nw = adjoint_window(a.sw, wq, a.Wq, a.W, ww, w_lo);
    const float* up = U + bq * a.H * (int64_t)a.W;
    float acc = 0.f;
    for (int i = 0; i < nh; ++i) {
      if (wh[i] == 0.f) continue;
      const float* row = up + (int64_t)(h_lo + i) * a.W + w_lo;
      float r = 0.f;
      for (int j = 0; j < nw; ++j) r = fmaf(ww[j], __ldg(row + j), r);
      acc = fmaf(wh[i], r, acc);
    }
    glow[idx] = acc;
  }
}

int stream_grid(int64_t nthreads, int block = 256) {
  const int64_t blocks = ceil_div64(nthreads, block);
  const int64_t cap = (int64_t)sm_count() * (2048 / block);  // full residency per SM
  return (int)(blocks < cap ? blocks : cap);
}
// Block size for the plane-walking kernels: with few, long-running threads (config 3: 131 072 threads of 192 planes
// each = 3.46 blocks of 256 per SM) the last wave leaves a quarter of the SMs idle; 128-thread blocks quantise finer.
int stream_block(int64_t nthreads) { return nthreads < (int64_t)sm_count() * 2048 ? 128 : 256; }

}  // namespace

int launch_concat_fwd(const float* ref, const float* tgt, float* cost, int B, int C, int D, int H,
                      int W, int d0, cudaStream_t st) {
  const int64_t rows = (int64_t)B * 2 * C * H;
  if (rows == 0 || D == 0 || W == 0) return PMT_OK;
  const int64_t cap = (int64_t)sm_count() * 16;
  const int grid = (int)(rows < cap ? rows : cap);
  const size_t smem = (size_t)round_up(W, 4) * sizeof(float);
  PMT_CHECK_ARG(smem <= 48 * 1024, "concat volume: W=%d too wide for the row stage", W);
  const bool vec = (W % 4 == 0) && aligned16(ref) && aligned16(tgt) && aligned16(cost);
  if (vec)
    concat_fwd_kernel<true><<<grid, 128, smem, st>>>(ref, tgt, cost, B, C, D, H, W, d0);
  else
    concat_fwd_kernel<false><<<grid, 128, smem, st>>>(ref, tgt, cost, B, C, D, H, W, d0);
  PMT_LAUNCH_OK("concat_fwd_kernel");
  return PMT_OK;
}

int launch_concat_bwd(const float* gcost, float* gref, float* gtgt, int B, int C, int D, int H, int W,
                      int d0, cudaStream_t st) {
  const int64_t total = (int64_t)B * 2 * C * H * W;
  if (total == 0) return PMT_OK;
  if (W % 4 == 0 && d0 % 4 == 0 && d0 >= 0 && aligned16(gcost) && aligned16(gref) && aligned16(gtgt)) {
    concat_bwd_vec_kernel<<<stream_grid(total / 4), 256, 0, st>>>(gcost, gref, gtgt, B, C, D, H, W, d0, total / 4);
    PMT_LAUNCH_OK("concat_bwd_vec_kernel");
    return PMT_OK;
  }
  concat_bwd_kernel<<<stream_grid(total), 256, 0, st>>>(gcost, gref, gtgt, B, C, D, H, W, d0, total);
  PMT_LAUNCH_OK("concat_bwd_kernel");
  return PMT_OK;
}

#define PMT_DISPATCH_VEC(kernel, vec_ok, total_pixels, ...)                                   \
  do {                                                                                        \
    if (vec_ok) {                                                                             \
      const int64_t ng = (total_pixels) / 4;                                                  \
      const int blk = stream_block(ng);                                                       \
      kernel<4><<<stream_grid(ng, blk), blk, 0, st>>>(__VA_ARGS__, ng);                       \
    } else {                                                                                  \
      const int64_t ng = (total_pixels);                                                      \
      const int blk = stream_block(ng);                                                       \
      kernel<1><<<stream_grid(ng, blk), blk, 0, st>>>(__VA_ARGS__, ng);                       \
    }                                                                                         \
    PMT_LAUNCH_OK(#kernel);                                                                   \
  } while (0)

int launch_softargmin_fwd(const float* cost, float* out, float* lse, int B, int D, int H, int W,
                          cudaStream_t st) {
  const int64_t plane = (int64_t)H * W;
  if (B * plane == 0) return PMT_OK;
  const bool vec = plane % 4 == 0 && aligned16(cost) && aligned16(out) && (lse == nullptr || aligned16(lse));
  PMT_DISPATCH_VEC(softargmin_fwd_kernel, vec, B * plane, cost, out, lse, D, plane);
  return PMT_OK;
}

int launch_softargmin_bwd(const float* cost, const float* out, const float* lse, const float* gout,
                          float* gcost, int B, int D, int H, int W, cudaStream_t st) {
  const int64_t plane = (int64_t)H * W;
  if (B * plane == 0) return PMT_OK;
  const bool vec = plane % 4 == 0 && aligned16(cost) && aligned16(out) && aligned16(lse) &&
                   aligned16(gout) && aligned16(gcost);
  PMT_DISPATCH_VEC(softargmin_bwd_kernel, vec, B * plane, cost, out, lse, gout, gcost, D, plane);
  return PMT_OK;
}

// footprint / shared-memory plan of the tiled kernels; false if the shape needs the gather kernel instead
static bool upsample_tiled_plan(const UpArgs& a, int* fr_max, int* fc_max, size_t* bytes) {
  // footprint bound of a kUpTH x kUpTW output block: (block-1)*scale + 2 source rows/columns, +1 for rounding
  *fr_max = (int)((kUpTH - 1) * a.sh) + 3 < a.Hq ? (int)((kUpTH - 1) * a.sh) + 3 : a.Hq;
  *fc_max = (int)((kUpTW - 1) * a.sw) + 3 < a.Wq ? (int)((kUpTW - 1) * a.sw) + 3 : a.Wq;
  *bytes = (size_t)(a.Dq + 1 + a.D) * 4 + (size_t)a.Dq * *fr_max * *fc_max * 4;
  return *bytes <= 64 * 1024 && ceil_div64(a.H, kUpTH) <= 65535 && a.B <= 65535;
}

int launch_upsample_softargmin_fwd(const float* lowres, float* out, float* lse, int B, int Dq, int Hq, int Wq, int D,
                                   int H, int W, cudaStream_t st) {
  const int64_t total = (int64_t)B * H * W;
  if (total == 0) return PMT_OK;
  UpArgs a{B, Dq, Hq, Wq, D, H, W, (float)Dq / (float)D, (float)Hq / (float)H, (float)Wq / (float)W};
  PMT_CHECK_ARG(D <= 4096, "upsample_softargmin: maxdisp %d too large for the weight table", D);
  int fr_max, fc_max;
  size_t tile_bytes;
  const int64_t gy = ceil_div64(H, kUpTH), gx = ceil_div64(W, kUpTW);
  static const bool force_gather = getenv("PMT_UPSOFT_GATHER") != nullptr;
  if (upsample_tiled_plan(a, &fr_max, &fc_max, &tile_bytes) && !force_gather) {
    PMT_CUDA_OK(cudaFuncSetAttribute(upsample_softargmin_tiled_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     64 * 1024));
    upsample_softargmin_tiled_kernel<false><<<dim3((unsigned)gx, (unsigned)gy, (unsigned)B), kUpTH * kUpTW, tile_bytes, st>>>(
        lowres, out, lse, a, fr_max, fc_max, nullptr, nullptr);
    PMT_LAUNCH_OK("upsample_softargmin_tiled_kernel");
    return PMT_OK;
  }
  upsample_softargmin_fwd_kernel<<<stream_grid(total), 256, (size_t)D * 8, st>>>(lowres, out, lse, a, total);
  PMT_LAUNCH_OK("upsample_softargmin_fwd_kernel");
  return PMT_OK;
}

// 1 if the fused backward supports this shape (tiled plan fits, adjoint windows fit), else the caller re-materialises
int upsample_softargmin_bwd_supported(int B, int Dq, int Hq, int Wq, int D, int H, int W) {
  if (B <= 0 || D <= 0 || H <= 0 || W <= 0 || D > 4096) return 0;
  UpArgs a{B, Dq, Hq, Wq, D, H, W, (float)Dq / (float)D, (float)Hq / (float)H, (float)Wq / (float)W};
  int fr_max, fc_max;
  size_t bytes;
  if (!upsample_tiled_plan(a, &fr_max, &fc_max, &bytes)) return 0;
  const float win_h = 2.f / a.sh + 5.f, win_w = 2.f / a.sw + 5.f;
  return win_h <= (float)kAdjMaxWin && win_w <= (float)kAdjMaxWin ? 1 : 0;
}

int launch_upsample_softargmin_bwd(const float* lowres, const float* out, const float* lse, const float* gout, float* U,
                                   float* glow, int B, int Dq, int Hq, int Wq, int D, int H, int W, cudaStream_t st) {
  if ((int64_t)B * Dq * Hq * Wq == 0) return PMT_OK;
  PMT_CHECK_ARG(upsample_softargmin_bwd_supported(B, Dq, Hq, Wq, D, H, W) == 1,
                "upsample_softargmin backward: shape not supported by the fused kernels");
  UpArgs a{B, Dq, Hq, Wq, D, H, W, (float)Dq / (float)D, (float)Hq / (float)H, (float)Wq / (float)W};
  int fr_max, fc_max;
  size_t tile_bytes;
  upsample_tiled_plan(a, &fr_max, &fc_max, &tile_bytes);
  const int64_t gy = ceil_div64(H, kUpTH), gx = ceil_div64(W, kUpTW);
  PMT_CUDA_OK(cudaFuncSetAttribute(upsample_softargmin_tiled_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   64 * 1024));
  upsample_softargmin_tiled_kernel<true><<<dim3((unsigned)gx, (unsigned)gy, (unsigned)B), kUpTH * kUpTW, tile_bytes, st>>>(
      lowres, const_cast<float*>(out), const_cast<float*>(lse), a, fr_max, fc_max, gout, U);
  PMT_LAUNCH_OK("upsample_softargmin_tiled_kernel<bwd>");
  const int64_t total = (int64_t)B * Dq * Hq * Wq;
  upsample_adjoint2d_kernel<<<stream_grid(total), 256, 0, st>>>(U, glow, a, total);
  PMT_LAUNCH_OK("upsample_adjoint2d_kernel");
  return PMT_OK;
}

int launch_dispreg_fwd(const float* x, float* out, int B, int D, int H, int W, cudaStream_t st) {
  const int64_t plane = (int64_t)H * W;
  if (B * plane == 0) return PMT_OK;
  const bool vec = plane % 4 == 0 && aligned16(x) && aligned16(out);
  PMT_DISPATCH_VEC(dispreg_fwd_kernel, vec, B * plane, x, out, D, plane);
  return PMT_OK;
}

int launch_dispreg_bwd(const float* gout, float* gx, int B, int D, int H, int W, cudaStream_t st) {
  const int64_t plane = (int64_t)H * W;
  if (B * plane == 0) return PMT_OK;
  const bool vec = plane % 4 == 0 && aligned16(gout) && aligned16(gx);
  PMT_DISPATCH_VEC(dispreg_bwd_kernel, vec, B * plane, gout, gx, D, plane);
  return PMT_OK;
}

}  // namespace pmt
