// corr1d_fwd.cu -- forward of the 1 x P horizontal correlation (SpatialCorrelationSampler with
// kernel_size=1, patch_size=(1,P), stride=1, padding=0, dilation_patch=1; reference call sites
// models/dsnet_t2.py:847-851 -> :879, :1078-1082 -> :1188, models/dsnet_t2_warp.py:615-619 -> :664).
//
//   out[n,0,p,h,w] = sum_c L[n,c,h,w] * R[n,c,h,w+p-rW],  rW=(P-1)/2,  0 outside the image.
//
// B200 design.  For one image row (n,h) the op is the band |w'-w+rW| < P of the Gram matrix
// G[w,w'] = sum_c L[c,w] R[c,w'] -- a GEMM with K=C whose operands are both "K-major rows", which is
// exactly how NCHW rows sit in memory.  One CTA owns 64 output columns and all P shifts:
//   * a producer warp streams 16-channel slabs of the L tile [16][64] and of the R band [16][BW]
//     (BW = 56+8*NQ floats, = 256 for P=192; its first column is rounded down to a multiple of 4
//     because TMA wants a 16-byte aligned inner start) into a 3-stage shared-memory ring with TMA
//     (cp.async.bulk.tensor, 4-D map over (W,H,C,B)); TMA's zero OOB fill implements the sampler's
//     "skip terms outside the image" (left/right halo, ragged last tile, C % 16) for free;
//   * compute threads hold 8x8 register tiles of G: thread (s,q) owns rows w = 8s..8s+7 and the two
//     4-wide column chunks u=q and u=q+NQ of that strip's band, so per channel it issues
//     4 LDS.128 for 64 FFMA (the SGEMM ratio) and the two R loads of a warp each touch 14
//     consecutive 16-byte chunks (2 shared-memory wavefronts, the minimum);
//   * the epilogue un-skews G into out[p][w]: register tiles are scattered into a [P][64] staging
//     tile (aliasing the drained ring) whose rows are rotated by p>>2 words, which makes both the
//     scattered writes and the row reads bank-conflict free; rows leave as coalesced 128-byte
//     streaming stores.
// Roofline: 2*C*Sum_p max(0,W-|s_p|) in-bounds FLOPs per row against 4*(2*C*W + P*W) bytes; at
// C=64, P=192 the op is FP32-pipe bound (AI 17.4 FLOP/B fwd) -- see DESIGN.md.
#include "common.cuh"

namespace pmt {
namespace {

constexpr int kWT = 64;            // output columns per CTA
constexpr int kNS = kWT / 8;       // 8-wide strips per CTA
constexpr int kCK = 16;            // channels per pipeline stage
constexpr int kStages = 3;
constexpr int kMaxThreads = 256;   // 7 compute warps (P<=193) + 1 producer warp

struct FwdArgs {
  int C, H, W, P, rW;
  int delta;      // band origin is rounded down to a multiple of 4 columns (TMA needs a 16-byte
                  // aligned start in the innermost dimension): smem column j <-> w' = w0 - rW - delta + j
  int NQ;         // column-chunk pairs per strip
  int BW;         // R band width in floats (inner box dim of the R tensor map)
  int n_wtiles, n_cchunks;
  int FG, REM;    // NQ = 8*FG + REM
  int ncw;        // compute warps
  int bar_off;    // byte offset of the mbarriers inside dynamic shared memory
};

__global__ void __launch_bounds__(kMaxThreads, 2)
corr1d_fwd_kernel(const __grid_constant__ CUtensorMap tmL, const __grid_constant__ CUtensorMap tmR,
                  float* __restrict__ out, const FwdArgs a) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  float* smem = reinterpret_cast<float*>(smem_raw);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + a.bar_off);
  uint64_t* empty = full + kStages;

  const int tid = threadIdx.x, wid = tid >> 5, lane = tid & 31;
  const int stage_floats = kCK * (kWT + a.BW);

  int bid = blockIdx.x;
  const int wt = bid % a.n_wtiles;
  bid /= a.n_wtiles;
  const int h = bid % a.H;
  const int n = bid / a.H;
  const int w0 = wt * kWT;

  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], a.ncw);
    }
    fence_mbar_init();
  }
  __syncthreads();

  if (wid == a.ncw) {
    // ===== TMA producer warp =====
    if (lane == 0) {
      tma_prefetch_desc(&tmL);
      tma_prefetch_desc(&tmR);
      const uint32_t bytes = (uint32_t)stage_floats * 4u;
      for (int k = 0; k < a.n_cchunks; ++k) {
        const int st = k % kStages;
        const uint32_t ph = (uint32_t)(k / kStages) & 1u;
        mbar_wait(&empty[st], ph ^ 1u);
        float* Ls = smem + st * stage_floats;
        float* Rs = Ls + kCK * kWT;
        mbar_arrive_expect_tx(&full[st], bytes);
        tma_load_4d(Ls, &tmL, w0, h, k * kCK, n, &full[st]);
        tma_load_4d(Rs, &tmR, w0 - a.rW - a.delta, h, k * kCK, n, &full[st]);
      }
    }
    return;
  }

  // ===== compute warps: lane -> (strip s, chunk pair q) =====
  int s, q;
  bool active = true;
  const int nfull = (kNS / 4) * a.FG;
  if (wid < nfull) {
    const int qg = wid % a.FG, sg = wid / a.FG;
    s = sg * 4 + (lane >> 3);
    q = qg * 8 + (lane & 7);
  } else {
    const int idx = (wid - nfull) * 32 + lane;
    s = idx / a.REM;
    q = a.FG * 8 + idx % a.REM;
    active = s < kNS;
    if (!active) s = 0, q = 0;
  }

  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  const int loff = 8 * s;
  const int roff = 8 * s + 4 * q;
  const int roff2 = roff + 4 * a.NQ;
  const int BW = a.BW;

  for (int k = 0; k < a.n_cchunks; ++k) {
    const int st = k % kStages;
    const uint32_t ph = (uint32_t)(k / kStages) & 1u;
    mbar_wait(&full[st], ph);
    const float* Ls = smem + st * stage_floats + loff;
    const float* Rs = smem + st * stage_floats + kCK * kWT;
    if (active) {
#pragma unroll 4
      for (int c = 0; c < kCK; ++c) {
        const float4 l0 = *reinterpret_cast<const float4*>(Ls + c * kWT);
        const float4 l1 = *reinterpret_cast<const float4*>(Ls + c * kWT + 4);
        const float4 r0 = *reinterpret_cast<const float4*>(Rs + c * BW + roff);
        const float4 r1 = *reinterpret_cast<const float4*>(Rs + c * BW + roff2);
        const float l[8] = {l0.x, l0.y, l0.z, l0.w, l1.x, l1.y, l1.z, l1.w};
        const float r[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(l[i], r[j], acc[i][j]);
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[st]);
  }

  // ===== epilogue: un-skew through a rotated [P][64] staging tile that aliases the ring =====
  const int nct = a.ncw * 32;
  named_bar_sync(1, nct);  // every compute warp has finished reading the ring
  float* tile = smem;
  if (active) {
#pragma unroll
    for (int hf = 0; hf < 2; ++hf) {
      const int u = q + hf * a.NQ;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
          const int p = 4 * u + jj - i - a.delta;
          if (p >= 0 && p < a.P) tile[p * kWT + ((8 * s + i + (p >> 2)) & (kWT - 1))] = acc[i][hf * 4 + jj];
        }
      }
    }
  }
  named_bar_sync(1, nct);
  for (int p = wid; p < a.P; p += a.ncw) {
    const int rot = p >> 2;
    float* row = out + (((int64_t)n * a.P + p) * a.H + h) * (int64_t)a.W + w0;
    const float* trow = tile + p * kWT;
#pragma unroll
    for (int t = 0; t < kWT; t += 32) {
      const int wl = t + lane;
      if (w0 + wl < a.W) st_cs(row + wl, trow[(wl + rot) & (kWT - 1)]);
    }
  }
}

}  // namespace

bool corr1d_fwd_fast_ok(const void* in1, const void* in2, int W, int P, int dilp) {
  if (dilp != 1 || P < 1 || W % 4 != 0 || !aligned16(in1) || !aligned16(in2)) return false;
  const int rW = (P - 1) / 2, delta = ((-rW % 4) + 4) % 4;
  const int U = (P + 6 + delta) / 4 + 1, NQ = (U + 1) / 2;
  return 8 * (kNS - 1) + 8 * NQ <= 256;  // TMA box limit
}

int launch_corr1d_fwd_tiled(const float* in1, const float* in2, float* out, int B, int C, int H, int W,
                            int P, cudaStream_t st) {
  FwdArgs a;
  a.C = C, a.H = H, a.W = W, a.P = P, a.rW = (P - 1) / 2;
  a.delta = ((-a.rW % 4) + 4) % 4;
  const int U = (P + 6 + a.delta) / 4 + 1;
  a.NQ = (U + 1) / 2;
  a.BW = 8 * (kNS - 1) + 8 * a.NQ;
  a.n_wtiles = ceil_div(W, kWT);
  a.n_cchunks = ceil_div(C, kCK);
  a.FG = a.NQ / 8, a.REM = a.NQ % 8;
  a.ncw = (kNS / 4) * a.FG + ceil_div(kNS * a.REM, 32);
  PMT_CHECK_ARG((a.ncw + 1) * 32 <= kMaxThreads, "corr1d fwd: P=%d needs %d warps", P, a.ncw + 1);
  const int ring_bytes = kStages * kCK * (kWT + a.BW) * 4;
  const int tile_bytes = P * kWT * 4;
  a.bar_off = round_up(ring_bytes > tile_bytes ? ring_bytes : tile_bytes, 128);
  const int smem_bytes = a.bar_off + 2 * kStages * 8;

  CUtensorMap tmL, tmR;
  if (int e = make_tmap_nchw(&tmL, in1, B, C, H, W, kWT, kCK)) return e;
  if (int e = make_tmap_nchw(&tmR, in2, B, C, H, W, a.BW, kCK)) return e;

  static int configured_smem = 0;
  if (smem_bytes > configured_smem) {
    PMT_CUDA_OK(cudaFuncSetAttribute(corr1d_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     smem_bytes));
    configured_smem = smem_bytes;
  }
  const int64_t grid = (int64_t)B * H * a.n_wtiles;
  PMT_CHECK_ARG(grid < (1ll << 31), "corr1d fwd: grid too large");
  corr1d_fwd_kernel<<<(unsigned)grid, (a.ncw + 1) * 32, smem_bytes, st>>>(tmL, tmR, out, a);
  PMT_LAUNCH_OK("corr1d_fwd_kernel");
  return PMT_OK;
}

}  // namespace pmt
