// corr1d_bwd_tc.cu -- tensor-core (tcgen05 + TMEM) backward of the 1 x P horizontal correlation.
// Both gradients are deterministic gathers (no atomics), one launch, blockIdx.y selects the gradient:
//   mode 0: gin1[c,x] = sum_j Gd[x][j] * in2[c][x0+oo+j]        mode 1: gin2[c,x] = sum_j Gd[x][j] * in1[c][x0+oo+j]
// where Gd[x][j] is the band matrix made of g (mode 0: g[j-delta-x][x]; mode 1: g[x+P-1+delta-j][x0+oo+j]).
// Per CTA: 128 output columns x of one image row, all C channels (C <= 128):
//   D[x (M=128 TMEM lanes)][c (N=C columns)] = sum over the band (K = 32*NKC columns, 320 for P=192).
//   * B operand = the feature band: rows c, K contiguous along w -> K-major, fetched by TMA in
//     32-column x C-channel boxes with the 128-byte swizzle, straight from NCHW;
//   * A operand = Gd, K-major too, built on the fly from raw g staged by TMA (no global-load latency
//     in the builders): mode 0 keeps the tile's [P][128] slice of g resident (32-row boxes, consumed as
//     they land), mode 1 streams [160][32] blocks (one per K chunk, OOB rows/columns zero-filled by
//     TMA); 8 builder warps re-lay them out shared->shared into the swizzled K-major stage with
//     bank-conflict-free LDS/STS; for 3xTF32 they write hi and lo copies and also split the band;
//   * warp 1 issues tcgen05.mma kind::tf32 (M=128, N=C, K=8), 4 k-steps per 32-column chunk; ring of
//     stages with mbarriers (TMA -> builders -> MMA -> free);
//   * epilogue: builder warps 0-3 read TMEM (tcgen05.ld) and store gin[c][x] rows directly -- a warp
//     writes 32 consecutive columns of one channel per instruction (coalesced 128 bytes).
#include <stdlib.h>

#include "tc_common.cuh"

namespace pmt {
namespace {

constexpr int kTM = 128;         // output columns per CTA (UMMA M)
constexpr int kKC = 32;          // band columns per ring stage (4 k-steps of 8)
constexpr int kGdBytes = kTM * kKC * 4;  // 16 KB: A operand chunk
constexpr int kBuilders = 8;     // builder warps
constexpr int kThreads = 32 * (2 + kBuilders);

struct TcBwdMode {
  int oo;      // band column j <-> image column x0 + oo + j  (multiple of 4)
  int delta;
};

constexpr int kRawRows1 = 160;                   // mode-1 raw block rows (>= 128+32-1)
constexpr int kRawSlot1 = kRawRows1 * kKC * 4;   // 20 KB

struct TcBwdArgs {
  int C, H, W, P, rW;
  int raw_bytes;       // raw g staging region at the start of shared memory
  int n_gboxes;        // mode 0: 32-row boxes of the resident g slice
  int Cbox;            // channels rounded up to 16 (UMMA N)
  int NKC;             // K chunks of 32 band columns
  int n_xtiles;
  int stages;
  int stage_bytes, lo_off;
  int bar_off;
  int tmem_cols;
  TcBwdMode m[2];
};

// K-major, 128-byte swizzle: row r of a [rows][32 floats] chunk
__device__ __forceinline__ uint32_t kmajor_off(int row, int col) {
  return (uint32_t)((row >> 3) * 1024 + (row & 7) * 128 + ((((col >> 2) ^ (row & 7)) << 4) | ((col & 3) << 2)));
}

__device__ __forceinline__ void split_tf32(float x, float& hi, float& lo) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
  hi = __uint_as_float(u);
  lo = x - hi;
}

template <int kPasses>
__global__ void __launch_bounds__(kThreads, 1)
corr1d_bwd_tc_kernel(const __grid_constant__ CUtensorMap tmIn1, const __grid_constant__ CUtensorMap tmIn2,
                     const __grid_constant__ CUtensorMap tmG0, const __grid_constant__ CUtensorMap tmG1,
                     float* __restrict__ gin1, float* __restrict__ gin2, const TcBwdArgs a) {
  extern __shared__ __align__(1024) unsigned char smem[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + a.bar_off);   // TMA band chunk landed
  uint64_t* built = full + 8;                                        // builders finished the stage
  uint64_t* empty = built + 8;                                       // MMAs finished reading the stage
  uint64_t* raw_full = empty + 8;                                    // raw g box / block landed
  uint64_t* raw_empty = raw_full + 8;                                // mode 1: builders done with a raw slot
  uint64_t* tmem_full = raw_empty + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);
  unsigned char* stage0 = smem + a.raw_bytes;

  const int mode = blockIdx.y;
  const TcBwdMode m = a.m[mode];
  const CUtensorMap* tmBand = mode == 0 ? &tmIn2 : &tmIn1;
  float* __restrict__ dst = mode == 0 ? gin1 : gin2;

  const int tid = threadIdx.x, wid = tid >> 5, lane = tid & 31;
  int bid = blockIdx.x;
  const int xt = bid % a.n_xtiles;
  bid /= a.n_xtiles;
  const int h = bid % a.H;
  const int n = bid / a.H;
  const int x0 = xt * kTM;
  const int band_bytes = a.Cbox * 128;

  if (tid == 0) {
    for (int s = 0; s < a.stages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&built[s], kBuilders);
      mbar_init(&empty[s], 1);
    }
    for (int s = 0; s < 8; ++s) mbar_init(&raw_full[s], 1);
    for (int s = 0; s < 2; ++s) mbar_init(&raw_empty[s], kBuilders);
    mbar_init(tmem_full, 1);
    fence_mbar_init();
  }
  if (wid == 1) {
    tc::tmem_alloc(tmem_slot, (uint32_t)a.tmem_cols);
    tc::tmem_relinquish();
  }
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  if (wid == 0) {
    // ===== TMA producer: raw g (resident boxes / streamed blocks) + one swizzled band box per stage =====
    if (lane == 0) {
      tma_prefetch_desc(tmBand);
      if (mode == 0) {
        for (int b = 0; b < a.n_gboxes; ++b) {
          mbar_arrive_expect_tx(&raw_full[b], 32u * kTM * 4u);
          tma_load_4d(smem + b * (32 * kTM * 4), &tmG0, x0, h, 32 * b, n, &raw_full[b]);
        }
      }
      for (int k = 0; k < a.NKC; ++k) {
        if (mode == 1) {
          const int slot = k & 1;
          mbar_wait(&raw_empty[slot], ((uint32_t)(k >> 1) & 1u) ^ 1u);
          mbar_arrive_expect_tx(&raw_full[slot], (uint32_t)kRawSlot1);
          tma_load_4d(smem + slot * kRawSlot1, &tmG1, x0 + m.oo + kKC * k, h,
                      a.P - 1 + m.delta - kKC * k - (kKC - 1), n, &raw_full[slot]);
        }
        const int st = k % a.stages;
        const uint32_t ph = (uint32_t)(k / a.stages) & 1u;
        mbar_wait(&empty[st], ph ^ 1u);
        unsigned char* sb = stage0 + (size_t)st * a.stage_bytes + kGdBytes;
        mbar_arrive_expect_tx(&full[st], (uint32_t)band_bytes);
        tma_load_4d(sb, tmBand, x0 + m.oo + kKC * k, h, 0, n, &full[st]);
      }
    }
  } else if (wid == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      const uint32_t idesc = tc::make_idesc(2, 0, 0, kTM, a.Cbox);
      for (int k = 0; k < a.NKC; ++k) {
        const int st = k % a.stages;
        const uint32_t ph = (uint32_t)(k / a.stages) & 1u;
        mbar_wait(&built[st], ph);
        if (kPasses == 1) mbar_wait(&full[st], ph);  // 3x: the builders already waited for (and rewrote) the band
        tc::fence_after_sync();
        const uint32_t sa = smem_u32(stage0 + (size_t)st * a.stage_bytes);
        const uint32_t sb = sa + kGdBytes;
#pragma unroll
        for (int kk = 0; kk < kKC / 8; ++kk) {
          const uint32_t acc = (k > 0 || kk > 0) ? 1u : 0u;
          const uint64_t dA = tc::smem_desc(sa + kk * 32, 16, 1024, 2);
          const uint64_t dB = tc::smem_desc(sb + kk * 32, 16, 1024, 2);
          if (kPasses == 3) {
            const uint64_t dAl = tc::smem_desc(sa + a.lo_off + kk * 32, 16, 1024, 2);
            const uint64_t dBl = tc::smem_desc(sb + a.lo_off + kk * 32, 16, 1024, 2);
            tc::mma_tf32(tmem_base, dAl, dB, idesc, acc);
            tc::mma_tf32(tmem_base, dA, dBl, idesc, 1u);
            tc::mma_tf32(tmem_base, dA, dB, idesc, 1u);
          } else {
            tc::mma_tf32(tmem_base, dA, dB, idesc, acc);
          }
        }
        tc::mma_commit(&empty[st]);
      }
      tc::mma_commit(tmem_full);
    }
  } else {
    // ===== builder warps =====
    const int bw = wid - 2;  // 0..7
    const int64_t pstride = (int64_t)a.H * a.W;
    int boxes_ready = 0;
    for (int k = 0; k < a.NKC; ++k) {
      const int st = k % a.stages;
      const uint32_t ph = (uint32_t)(k / a.stages) & 1u;
      unsigned char* sa = stage0 + (size_t)st * a.stage_bytes;
      if (mode == 0) {
        // rows p <= 32k+31-delta are needed: boxes 0..k of the resident [P][128] slice
        const int need = (k + 1 < a.n_gboxes) ? k + 1 : a.n_gboxes;
        while (boxes_ready < need) mbar_wait(&raw_full[boxes_ready++], 0);
        mbar_wait(&empty[st], ph ^ 1u);
        const float* Gt = reinterpret_cast<const float*>(smem);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int it = bw + kBuilders * i;      // 32 warp tasks per chunk: (16-byte column c4, 32-row block xb)
          const int c4 = it & 7, xb = it >> 3;
          const int xl = 32 * xb + lane;
          const int pb = kKC * k + 4 * c4 - m.delta - xl;  // p of column jj = 4*c4 + t is pb + t
          float v[4];
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            const int p = pb + t;
            v[t] = (p >= 0 && p < a.P) ? Gt[p * kTM + xl] : 0.f;
          }
          const uint32_t off = kmajor_off(xl, 4 * c4);
          if (kPasses == 3) {
            float4 hi, lo;
            split_tf32(v[0], hi.x, lo.x);
            split_tf32(v[1], hi.y, lo.y);
            split_tf32(v[2], hi.z, lo.z);
            split_tf32(v[3], hi.w, lo.w);
            *reinterpret_cast<float4*>(sa + off) = hi;
            *reinterpret_cast<float4*>(sa + a.lo_off + off) = lo;
          } else {
            *reinterpret_cast<float4*>(sa + off) = make_float4(v[0], v[1], v[2], v[3]);
          }
        }
      } else {
        const int slot = k & 1;
        mbar_wait(&raw_full[slot], (uint32_t)(k >> 1) & 1u);
        mbar_wait(&empty[st], ph ^ 1u);
        const float* raw = reinterpret_cast<const float*>(smem + slot * kRawSlot1);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int r = bw + kBuilders * i;       // 32 warp tasks per chunk: 4 rows x 32 columns (lane = column)
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            const int xl = 4 * r + t;
            const float v = raw[(xl + 31 - lane) * kKC + lane];   // g[p_first + xl + 31 - jj][column jj]
            const uint32_t off = kmajor_off(xl, lane);
            if (kPasses == 3) {
              float hi, lo;
              split_tf32(v, hi, lo);
              *reinterpret_cast<float*>(sa + off) = hi;
              *reinterpret_cast<float*>(sa + a.lo_off + off) = lo;
            } else {
              *reinterpret_cast<float*>(sa + off) = v;
            }
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&raw_empty[slot]);
      }
      if (kPasses == 3) {
        // split the landed feature band chunk into hi / lo (position-wise, layout agnostic)
        mbar_wait(&full[st], ph);
        unsigned char* sb = sa + kGdBytes;
        const int nch = band_bytes / 16;
        for (int c = (bw * 32 + lane); c < nch; c += kBuilders * 32) {
          float4* q = reinterpret_cast<float4*>(sb + 16 * c);
          const float4 x = *q;
          float4 hi, lo;
          split_tf32(x.x, hi.x, lo.x);
          split_tf32(x.y, hi.y, lo.y);
          split_tf32(x.z, hi.z, lo.z);
          split_tf32(x.w, hi.w, lo.w);
          *q = hi;
          *reinterpret_cast<float4*>(sb + a.lo_off + 16 * c) = lo;
        }
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(&built[st]);
    }
    // ===== epilogue (builder warps 0..3): TMEM -> coalesced global stores =====
    if (bw < 4) {
      const int q = wid & 3;
      const int xl = 32 * q + lane;
      mbar_wait(tmem_full, 0);
      tc::fence_after_sync();
      const bool ok = x0 + xl < a.W;
      float* o = dst + ((int64_t)n * a.C * a.H + h) * (int64_t)a.W + x0 + xl;
      for (int cb = 0; cb < a.Cbox; cb += 32) {
        float v[32];
        tc::tmem_ld32(tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)cb, v);
#pragma unroll
        for (int cc = 0; cc < 32; ++cc)
          if (ok && cb + cc < a.C) o[(int64_t)(cb + cc) * pstride] = v[cc];
      }
    }
  }

  tc::fence_before_sync();
  __syncthreads();
  if (wid == 1) {
    tc::fence_after_sync();
    tc::tmem_dealloc(tmem_base, (uint32_t)a.tmem_cols);
  }
}

int fill_args(TcBwdArgs* a, int C, int H, int W, int P, int passes) {
  a->C = C, a->H = H, a->W = W, a->P = P, a->rW = (P - 1) / 2;
  if (C > 128) return 1;
  a->Cbox = round_up(C, 32);  // the epilogue reads TMEM in 32-column groups; UMMA N % 16 == 0
  const int oo0 = -a->rW, oo1 = -(P - 1 - a->rW);
  a->m[0].delta = ((oo0 % 4) + 4) % 4;
  a->m[0].oo = oo0 - a->m[0].delta;
  a->m[1].delta = ((oo1 % 4) + 4) % 4;
  a->m[1].oo = oo1 - a->m[1].delta;
  const int dmax = a->m[0].delta > a->m[1].delta ? a->m[0].delta : a->m[1].delta;
  a->NKC = ceil_div(kTM + P - 1 + dmax, kKC);
  a->n_xtiles = ceil_div(W, kTM);
  const int hi_bytes = kGdBytes + a->Cbox * 128;
  a->lo_off = hi_bytes;
  a->stage_bytes = hi_bytes * (passes == 3 ? 2 : 1);
  a->n_gboxes = ceil_div(P, 32);
  if (a->n_gboxes > 8) return 1;
  const int raw0 = a->n_gboxes * 32 * kTM * 4, raw1 = 2 * kRawSlot1;
  a->raw_bytes = round_up(raw0 > raw1 ? raw0 : raw1, 1024);
  int stages = (227 * 1024 - 512 - a->raw_bytes) / a->stage_bytes;
  if (stages > 6) stages = 6;
  if (stages > a->NKC) stages = a->NKC;
  if (stages < 2) return 1;
  a->stages = stages;
  a->bar_off = a->raw_bytes + stages * a->stage_bytes;
  int cols = 32;
  while (cols < a->Cbox) cols *= 2;
  a->tmem_cols = cols;
  return 0;
}

}  // namespace

bool corr1d_bwd_tc_ok(const void* in1, const void* in2, int C, int H, int W, int P, int dilp, int passes) {
  if (dilp != 1 || P < 1 || C < 1 || W % 4 != 0 || !aligned16(in1) || !aligned16(in2)) return false;
  TcBwdArgs a;
  return fill_args(&a, C, H, W, P, passes) == 0;
}

int launch_corr1d_bwd_tc(const float* in1, const float* in2, const float* gout, float* gin1, float* gin2, int B,
                         int C, int H, int W, int P, int passes, cudaStream_t st) {
  TcBwdArgs a;
  PMT_CHECK_ARG(passes == 1 || passes == 3, "corr1d tc: passes must be 1 (tf32) or 3 (3xtf32)");
  PMT_CHECK_ARG(fill_args(&a, C, H, W, P, passes) == 0, "corr1d tc bwd: unsupported shape C=%d P=%d", C, P);
  CUtensorMap tm1, tm2, tmG0, tmG1;
  if (int e = make_tmap_nchw_ex(&tm1, in1, B, C, H, W, kKC, a.Cbox, 1)) return e;
  if (int e = make_tmap_nchw_ex(&tm2, in2, B, C, H, W, kKC, a.Cbox, 1)) return e;
  if (int e = make_tmap_nchw_ex(&tmG0, gout, B, P, H, W, kTM, 32, 0)) return e;
  if (int e = make_tmap_nchw_ex(&tmG1, gout, B, P, H, W, kKC, kRawRows1, 0)) return e;
  const int smem_bytes = a.bar_off + 512;
  const int64_t gx = (int64_t)B * H * a.n_xtiles;
  PMT_CHECK_ARG(gx < (1ll << 31), "corr1d tc bwd: grid too large");
  if (passes == 3) {
    PMT_CUDA_OK(cudaFuncSetAttribute(corr1d_bwd_tc_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    corr1d_bwd_tc_kernel<3><<<dim3((unsigned)gx, 2), kThreads, smem_bytes, st>>>(tm1, tm2, tmG0, tmG1, gin1, gin2, a);
  } else {
    PMT_CUDA_OK(cudaFuncSetAttribute(corr1d_bwd_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    corr1d_bwd_tc_kernel<1><<<dim3((unsigned)gx, 2), kThreads, smem_bytes, st>>>(tm1, tm2, tmG0, tmG1, gin1, gin2, a);
  }
  PMT_LAUNCH_OK("corr1d_bwd_tc_kernel");
  return PMT_OK;
}

}  // namespace pmt
