// corr1d_bwd.cu -- backward of the 1 x P horizontal correlation, both gradients as DETERMINISTIC
// GATHERS (the upstream CUDA extension launches B x 2 grids of (C,H,W) 25-thread blocks; no atomics
// here either, and a fixed summation order so results are bit-reproducible run to run).
//
//   gin1[n,c,h,w]  = sum_p g[n,0,p,h,w]      * in2[n,c,h,w+p-rW]
//   gin2[n,c,h,w'] = sum_p g[n,0,p,h,w'-s_p] * in1[n,c,h,w'-s_p],   s_p = p-rW     (in-bounds terms)
//
// Per image row both are a product of a feature band A[c][j] with the banded matrix Gg[x][j] built
// from g.  A CTA owns 128 output columns x of one row, for one of the two gradients (blockIdx.y):
//   * it un-skews its slice of g ONCE into shared memory as T[strip][k][16]: for the 16 columns of a
//     strip, row k holds the g values that multiply band column (16*strip + k).  Both operands of
//     the contraction are then k-contiguous, so the inner loop is a plain register-tiled GEMM:
//     per 4 k-steps a thread issues 4 LDS.128 of A and 16 LDS.128 of T for 256 FFMA;
//   * feature bands arrive in 32-channel passes through cp.async (16-byte copies with zero fill
//     implement the image border), double buffered; rows are padded to an odd number of 16-byte
//     chunks so the 8 channel-lanes of a warp hit distinct banks;
//   * a warp = one strip x 8 channel groups x 4 k-splits; the 4 partial sums are combined with two
//     xor-shuffles (fixed order), then each lane stores one 64-byte row segment.
// The T build costs one pass over g per gradient; it is amortised over all C channels.
#include "common.cuh"

namespace pmt {
namespace {

constexpr int kXT = 128;           // output columns per CTA
constexpr int kXS = 16;            // columns per strip (thread tile width)
constexpr int kNSb = kXT / kXS;    // strips per CTA == warps per CTA
constexpr int kCP = 32;            // channels per pass
constexpr int kNCG = kCP / 4;      // channel groups (lanes); thread rows = cg + 8*cc
constexpr int kKS = 4;             // k-splits (lanes)
constexpr int kThreads = 32 * kNSb;

struct BwdMode {
  int oo;      // band column j (aligned) <-> image column x0 + oo + j
  int delta;   // T row index = k' + delta
  int KP;      // contraction length, = kKS * KSL
  int KSL;     // per-split length (multiple of 4)
  int BWA;     // band width in floats (multiple of 4)
  int AST;     // padded band row stride in floats ((AST/4) odd)
  int TS;      // floats per strip of T
};

struct BwdArgs {
  int C, H, W, P, rW;
  int n_xtiles, n_passes;
  BwdMode m[2];
};

__device__ __forceinline__ void load_band(float* Abuf, const float* __restrict__ src, const BwdArgs& a,
                                          const BwdMode& m, int n, int h, int x0, int pass, int tid) {
  const int nch = m.BWA >> 2;
  const int col0 = x0 + m.oo;
  for (int e = tid; e < kCP * nch; e += kThreads) {
    const int r = e / nch, t = e - r * nch;
    const int c = pass * kCP + r;
    const int col = col0 + 4 * t;
    const bool valid = (c < a.C) && (col >= 0) && (col < a.W);
    const float* g = valid ? src + (((int64_t)n * a.C + c) * a.H + h) * (int64_t)a.W + col : src;
    cp_async16(Abuf + r * m.AST + 4 * t, g, valid);
  }
}

__global__ void __launch_bounds__(kThreads, 1)
corr1d_bwd_kernel(const float* __restrict__ in1, const float* __restrict__ in2,
                  const float* __restrict__ gout, float* __restrict__ gin1, float* __restrict__ gin2,
                  const BwdArgs a) {
  extern __shared__ __align__(16) float smem[];
  const int mode = blockIdx.y;  // 0: gin1 from in2, 1: gin2 from in1
  const BwdMode m = a.m[mode];
  const float* __restrict__ src = mode == 0 ? in2 : in1;
  float* __restrict__ dst = mode == 0 ? gin1 : gin2;

  const int tid = threadIdx.x, wid = tid >> 5, lane = tid & 31;
  int bid = blockIdx.x;
  const int xt = bid % a.n_xtiles;
  bid /= a.n_xtiles;
  const int h = bid % a.H;
  const int n = bid / a.H;
  const int x0 = xt * kXT;

  float* T = smem;
  float* A0 = smem + kNSb * m.TS;
  const int abuf = kCP * m.AST;

  // pass 0 of the feature band goes in flight before anything else
  load_band(A0, src, a, m, n, h, x0, 0, tid);
  cp_async_commit();

  // ---- build T: zero the ragged head/tail rows, then scatter the g slice ----
  {
    const int nz1 = kXS - 1 + m.delta;           // rows [0, nz1) are partly empty
    const int tail0 = a.P + m.delta;             // rows [tail0, KP) are partly empty
    const int nzr = nz1 + (m.KP - tail0);
    for (int e = tid; e < kNSb * nzr * 4; e += kThreads) {
      const int part = e & 3;
      int rr = (e >> 2) % nzr;
      const int s = (e >> 2) / nzr;
      const int k = rr < nz1 ? rr : tail0 + (rr - nz1);
      *reinterpret_cast<float4*>(T + s * m.TS + k * 16 + 4 * (k / m.KSL) + 4 * part) =
          make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  __syncthreads();
  {
    const int xl = tid & (kXT - 1);
    const int s = xl >> 4, i = xl & 15;
    const float* gplane = gout + ((int64_t)n * a.P * a.H + h) * (int64_t)a.W;
    const int64_t pstride = (int64_t)a.H * a.W;
    float* Ts = T + s * m.TS + i;
    if (mode == 0) {
      const int x = x0 + xl;
      const bool ok = x < a.W;
#pragma unroll 8
      for (int p = tid >> 7; p < a.P; p += kThreads / kXT) {
        const float v = ok ? __ldg(gplane + p * pstride + x) : 0.f;
        const int k = p + i + m.delta;
        Ts[k * 16 + 4 * (k / m.KSL)] = v;
      }
    } else {
#pragma unroll 8
      for (int p = tid >> 7; p < a.P; p += kThreads / kXT) {
        const int w = x0 + xl + a.rW - p;
        const float v = (w >= 0 && w < a.W) ? __ldg(gplane + p * pstride + w) : 0.f;
        const int k = (a.P - 1 - p) + i + m.delta;
        Ts[k * 16 + 4 * (k / m.KSL)] = v;
      }
    }
  }

  // ---- main loop over channel passes ----
  const int s = wid;
  const int cg = lane & 7, ks = lane >> 3;
  const float* Tw = T + s * m.TS + 4 * ks;  // this split's rows live at +4*ks
  const int kbeg = ks * m.KSL;

  for (int pass = 0; pass < a.n_passes; ++pass) {
    float* Acur = A0 + (pass & 1) * abuf;
    if (pass + 1 < a.n_passes) load_band(A0 + ((pass + 1) & 1) * abuf, src, a, m, n, h, x0, pass + 1, tid);
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();  // this pass's band (and, first time round, T) is visible to every warp

    float acc[4][16];
#pragma unroll
    for (int cc = 0; cc < 4; ++cc)
#pragma unroll
      for (int i = 0; i < 16; ++i) acc[cc][i] = 0.f;

    const float* Ar = Acur + cg * m.AST + kXS * s;
#pragma unroll 1
    for (int kk = kbeg; kk < kbeg + m.KSL; kk += 4) {
      float av[4][4];
#pragma unroll
      for (int cc = 0; cc < 4; ++cc) {
        const float4 t = *reinterpret_cast<const float4*>(Ar + cc * kNCG * m.AST + kk);
        av[cc][0] = t.x, av[cc][1] = t.y, av[cc][2] = t.z, av[cc][3] = t.w;
      }
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const float* Tr = Tw + (kk + t) * 16;
        const float4 b0 = *reinterpret_cast<const float4*>(Tr);
        const float4 b1 = *reinterpret_cast<const float4*>(Tr + 4);
        const float4 b2 = *reinterpret_cast<const float4*>(Tr + 8);
        const float4 b3 = *reinterpret_cast<const float4*>(Tr + 12);
        const float b[16] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w,
                             b2.x, b2.y, b2.z, b2.w, b3.x, b3.y, b3.z, b3.w};
#pragma unroll
        for (int cc = 0; cc < 4; ++cc)
#pragma unroll
          for (int i = 0; i < 16; ++i) acc[cc][i] = fmaf(av[cc][t], b[i], acc[cc][i]);
      }
    }

    // combine the 4 k-splits (lanes differing in bits 3,4), fixed order
#pragma unroll
    for (int cc = 0; cc < 4; ++cc)
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        float v = acc[cc][i];
        v += __shfl_xor_sync(0xffffffffu, v, 8);
        v += __shfl_xor_sync(0xffffffffu, v, 16);
        acc[cc][i] = v;
      }
    // lane (cg, ks) stores channel row cg + 8*ks of this pass: 16 consecutive columns
#pragma unroll
    for (int cc = 0; cc < 4; ++cc) {
      if (ks == cc) {
        const int c = pass * kCP + cg + kNCG * cc;
        const int x = x0 + kXS * s;
        if (c < a.C) {
          float* o = dst + (((int64_t)n * a.C + c) * a.H + h) * (int64_t)a.W + x;
#pragma unroll
          for (int v4 = 0; v4 < 4; ++v4)
            if (x + 4 * v4 < a.W)
              *reinterpret_cast<float4*>(o + 4 * v4) =
                  make_float4(acc[cc][4 * v4], acc[cc][4 * v4 + 1], acc[cc][4 * v4 + 2], acc[cc][4 * v4 + 3]);
        }
      }
    }
    __syncthreads();  // everyone is done with Acur before pass+2 overwrites it
  }
}

void fill_mode(BwdMode* m, int P, int oo_unaligned) {
  const int delta = ((oo_unaligned % 4) + 4) % 4;
  m->delta = delta;
  m->oo = oo_unaligned - delta;
  const int need = P + kXS - 1 + delta;
  m->KSL = round_up(ceil_div(need, kKS), 4);
  m->KP = m->KSL * kKS;
  m->BWA = kXS * (kNSb - 1) + m->KP;
  m->AST = ((m->BWA / 4) % 2 == 0) ? m->BWA + 4 : m->BWA;
  m->TS = m->KP * 16 + 16;
}

size_t bwd_smem_bytes(const BwdMode& m) {
  return ((size_t)kNSb * m.TS + 2u * kCP * m.AST) * sizeof(float);
}

}  // namespace

bool corr1d_bwd_fast_ok(const void* in1, const void* in2, const void* gout, int C, int W, int P, int dilp) {
  (void)gout;
  if (dilp != 1 || P < 1 || C < 1 || W % 4 != 0 || !aligned16(in1) || !aligned16(in2)) return false;
  BwdMode m0, m1;
  const int rW = (P - 1) / 2;
  fill_mode(&m0, P, -rW);
  fill_mode(&m1, P, -(P - 1 - rW));
  const size_t lim = 227 * 1024;
  return bwd_smem_bytes(m0) <= lim && bwd_smem_bytes(m1) <= lim;
}

int launch_corr1d_bwd_tiled(const float* in1, const float* in2, const float* gout, float* gin1,
                            float* gin2, int B, int C, int H, int W, int P, cudaStream_t st) {
  BwdArgs a;
  a.C = C, a.H = H, a.W = W, a.P = P, a.rW = (P - 1) / 2;
  a.n_xtiles = ceil_div(W, kXT);
  a.n_passes = ceil_div(C, kCP);
  fill_mode(&a.m[0], P, -a.rW);
  fill_mode(&a.m[1], P, -(P - 1 - a.rW));
  const size_t s0 = bwd_smem_bytes(a.m[0]), s1 = bwd_smem_bytes(a.m[1]);
  const size_t smem_bytes = s0 > s1 ? s0 : s1;
  PMT_CUDA_OK(cudaFuncSetAttribute(corr1d_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)smem_bytes));
  const int64_t gx = (int64_t)B * H * a.n_xtiles;
  PMT_CHECK_ARG(gx < (1ll << 31), "corr1d bwd: grid too large");
  corr1d_bwd_kernel<<<dim3((unsigned)gx, 2), kThreads, smem_bytes, st>>>(in1, in2, gout, gin1, gin2, a);
  PMT_LAUNCH_OK("corr1d_bwd_kernel");
  return PMT_OK;
}

}  // namespace pmt
