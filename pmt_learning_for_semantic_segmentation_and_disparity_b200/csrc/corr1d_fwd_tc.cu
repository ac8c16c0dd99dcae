// corr1d_fwd_tc.cu -- tensor-core (tcgen05 + TMEM) forward of the 1 x P horizontal correlation.
//
// Per image row the op is the band of the Gram matrix G[w,w'] = sum_c L[c,w] R[c,w'].  NCHW rows are
// "MN-major" operands (w contiguous, channel = K strided), which tcgen05.mma kind::tf32 reads straight
// from shared memory, so fp32 features go TMA -> SWIZZLE_128B smem -> tensor core with no cast pass:
//   * CTA tile = 128 output columns x the whole band (N = 32*NB columns, 320 for P=192); accumulator
//     D[128 lanes][N cols] fp32 lives in TMEM;
//   * warp 0 streams 16-channel slabs (boxes of 32 columns x 16 channels, 2 KB, 128-byte swizzle) of
//     the L tile and the R band through a ring with TMA; warp 1 issues tcgen05.mma (M=128, N<=256,
//     K=8 per instruction) and commits to mbarriers; 4 epilogue warps read TMEM with tcgen05.ld (every 32-column
//     block exactly once), un-skew D[w][j] -> out[p = j - w - delta][w] into ping-pong [32][128] staging tiles
//     (bank-conflict free: the row stride is 128 words) and issue one TMA store per 32 planes;
//   * kPasses = 1: plain TF32 (10-bit mantissa inputs, fp32 accumulate) -- the reduced-precision variant;
//     kPasses = 3: "3xTF32": 4 transform warps split every staged value into hi = tf32(x), lo = x - hi
//     (exact in fp32) and the MMA warp accumulates hi*hi + hi*lo + lo*hi, recovering fp32-class
//     accuracy (error ~2^-22 per product) at 3x the tensor work, still far below the FP32-pipe time.
#include <stdlib.h>

#include "tc_common.cuh"

namespace pmt {
namespace {

constexpr int kTM = 128;              // output columns per CTA (= UMMA M = TMEM lanes)
constexpr int kCK = 16;               // channels per ring stage (2 k-steps of 8)
constexpr int kBoxBytes = kCK * 128;  // one 32-column x 16-channel box
constexpr int kLBlocks = kTM / 32;    // 4 boxes for the L tile
constexpr int kMaxNB = 10;            // band boxes (P <= 193)
constexpr int kRowsPerStep = 32;      // output planes per staging buffer / TMA store (= one TMEM column block)
constexpr int kStepBytes = kRowsPerStep * kTM * 4;
#ifndef PMT_FWD_XF_GROUPS
#define PMT_FWD_XF_GROUPS 1
#endif
constexpr int kXfGroups = PMT_FWD_XF_GROUPS;   // groups of 4 transform warps (3xTF32); needs lo_stages >= kXfGroups
constexpr int kFwdThreads3 = 32 * (6 + 4 * kXfGroups);

struct TcFwdArgs {
  int C, H, W, P, rW, delta;
  int NB;                // band boxes of 32 columns
  int N1, N2;            // MMA N of the two band halves (N2 may be 0)
  int tmem_cols;         // power of two >= 32*NB
  int n_wtiles, n_cchunks, n_tiles;
  int stages;            // raw ring stages (TMA destination; the raw fp32 doubles as the tf32 "hi" operand)
  int lo_stages;         // lo ring stages (3xTF32 only)
  int stage_bytes;       // (4+NB)*kBoxBytes
  int lo_ring_off;       // byte offset of the lo ring
  int tile_off;          // byte offset of the [32][128] staging buffers
  int n_steps;           // epilogue steps of kRowsPerStep output planes
  int n_bufs;            // staging buffers (2 or 3)
  int bar_off;           // byte offset of the barriers
  int debug;             // PMT_TC_DEBUG: 1 = constant tile, 2 = tcgen05.st pattern instead of MMA result
  // TMEM-A variant (corr1d_fwd_tca_kernel): the L tile is the A operand in TMEM, only the R band goes through the rings
  int l_stages;          // L staging ring ([16 ch][128 w] fp32, no swizzle, 8 KB per stage)
  int l_ring_off;        // byte offset of the L staging ring
  int a_base;            // first TMEM column of the A ring (4 slots x (16 hi + 16 lo) columns)
};
constexpr int kLStageBytes = kCK * kTM * 4;   // 8 KB
constexpr int kASlots = 4;
constexpr int kASlotCols = 2 * kCK;           // 16 hi + 16 lo columns

// MN-major tf32 operand: 32 columns x 4 channel rows per swizzle atom (Swizzle<2,5,2>, 128-byte rows),
// column blocks kBoxBytes apart (LBO), 4-row k-groups 512 bytes apart (SBO).
__device__ __forceinline__ uint64_t mn_desc(uint32_t saddr, uint32_t lbo, uint32_t) {
  return tc::smem_desc(saddr, lbo, 512, 1);
}

template <int kPasses>
__global__ void __launch_bounds__(kPasses == 3 ? kFwdThreads3 : 192, 1)
corr1d_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmL, const __grid_constant__ CUtensorMap tmR,
                     const __grid_constant__ CUtensorMap tmO, const TcFwdArgs a) {
  extern __shared__ __align__(1024) unsigned char smem[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + a.bar_off);
  uint64_t* empty = full + 8;
  uint64_t* xf_done = empty + 8;
  uint64_t* lo_empty = xf_done + 8;
  uint64_t* tmem_full = lo_empty + 8;
  uint64_t* tmem_empty = tmem_full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 1);

  const int tid = threadIdx.x, wid = tid >> 5, lane = tid & 31;
  const int nboxes = kLBlocks + a.NB;

  if (tid == 0) {
    for (int s = 0; s < a.stages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int s = 0; s < 8; ++s) {
      mbar_init(&xf_done[s], 4);
      mbar_init(&lo_empty[s], 1);
    }
    mbar_init(tmem_full, 1);
    mbar_init(tmem_empty, 4);
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

  // Persistent CTA: tiles blockIdx.x, blockIdx.x + gridDim.x, ...  Every role walks the same tile list and the
  // same global chunk counter g, so the smem ring keeps streaming across tile boundaries.
  if (wid == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      tma_prefetch_desc(&tmL);
      tma_prefetch_desc(&tmR);
    }
    const uint32_t bytes = (uint32_t)nboxes * kBoxBytes;
    int st = 0;
    uint32_t eph = 1;  // parity to wait for on empty[st] (the first pass over the ring is free)
    for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
      const int wt = tile % a.n_wtiles, h = (tile / a.n_wtiles) % a.H, n = tile / (a.n_wtiles * a.H);
      const int w0 = wt * kTM;
      for (int k = 0; k < a.n_cchunks; ++k) {
        mbar_wait(&empty[st], eph);
        unsigned char* sbase = smem + (size_t)st * a.stage_bytes;
        if (lane == 0) mbar_arrive_expect_tx(&full[st], bytes);
        __syncwarp();
        if (lane < nboxes) {
          if (lane < kLBlocks)
            tma_load_4d(sbase + lane * kBoxBytes, &tmL, w0 + 32 * lane, h, k * kCK, n, &full[st]);
          else
            tma_load_4d(sbase + lane * kBoxBytes, &tmR, w0 - a.rW - a.delta + 32 * (lane - kLBlocks), h, k * kCK, n,
                        &full[st]);
        }
        if (++st == a.stages) st = 0, eph ^= 1u;
      }
    }
  } else if (wid == 1) {
    // ===== MMA issuer: the warp runs the loop converged, one elected lane issues (uniform-register code, see
    // tc::elect_one) =====
    {
      const uint32_t idesc1 = tc::make_idesc(2, 1, 1, kTM, a.N1);
      const uint32_t idesc2 = tc::make_idesc(2, 1, 1, kTM, a.N2 > 0 ? a.N2 : 16);
      const uint64_t d0 = mn_desc(smem_u32(smem), kBoxBytes, 1024);                       // raw ring (hi operand)
      const uint64_t dl0 = mn_desc(smem_u32(smem + a.lo_ring_off), kBoxBytes, 1024);      // lo ring
      const uint32_t st_step = (uint32_t)a.stage_bytes >> 4;
      const uint32_t offB1 = (uint32_t)(kLBlocks * kBoxBytes) >> 4;
      const uint32_t offB2 = offB1 + ((uint32_t)((a.N1 / 32) * kBoxBytes) >> 4);
      const bool skip = PMT_DBG(a, 16) != 0, two = a.N2 > 0;
      int st = 0, ls = 0, it = 0;
      uint32_t fph = 0, lph = 0;
      for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x, ++it) {
        mbar_wait(tmem_empty, ((uint32_t)it & 1u) ^ 1u);  // epilogue of the previous tile has drained TMEM
        tc::fence_after_sync();
        for (int k = 0; k < a.n_cchunks; ++k) {
          if (kPasses == 3) mbar_wait(&xf_done[ls], lph);
          else mbar_wait(&full[st], fph);
          tc::fence_after_sync();
          const uint64_t dh = d0 + (uint64_t)(st_step * (uint32_t)st);
          const uint64_t dl = dl0 + (uint64_t)(st_step * (uint32_t)ls);
          if (tc::elect_one()) {
#pragma unroll
            for (int kk = 0; kk < kCK / 8; ++kk) {
              if (skip) break;
              const uint32_t acc = (k > 0 || kk > 0) ? 1u : 0u;
              const uint32_t ko = (uint32_t)(kk * 1024) >> 4;
              const uint64_t dA = dh + ko, dB1 = dh + offB1 + ko, dB2 = dh + offB2 + ko;
              if (kPasses == 3) {
                const uint64_t dAl = dl + ko, dB1l = dl + offB1 + ko, dB2l = dl + offB2 + ko;
                // small cross terms first, then the dominant hi*hi term
                tc::mma_tf32(tmem_base, dAl, dB1, idesc1, acc);
                tc::mma_tf32(tmem_base, dA, dB1l, idesc1, 1u);
                tc::mma_tf32(tmem_base, dA, dB1, idesc1, 1u);
                if (two) {
                  tc::mma_tf32(tmem_base + a.N1, dAl, dB2, idesc2, acc);
                  tc::mma_tf32(tmem_base + a.N1, dA, dB2l, idesc2, 1u);
                  tc::mma_tf32(tmem_base + a.N1, dA, dB2, idesc2, 1u);
                }
              } else {
                tc::mma_tf32(tmem_base, dA, dB1, idesc1, acc);
                if (two) tc::mma_tf32(tmem_base + a.N1, dA, dB2, idesc2, acc);
              }
            }
            tc::mma_commit(&empty[st]);  // ring slot reusable once these MMAs have read it
            if (kPasses == 3) tc::mma_commit(&lo_empty[ls]);
            if (k == a.n_cchunks - 1) tc::mma_commit(tmem_full);     // accumulator of this tile complete
          }
          __syncwarp();
          if (++st == a.stages) st = 0, fph ^= 1u;
          if (++ls == a.lo_stages) ls = 0, lph ^= 1u;
        }
      }
    }
  } else if (wid < 6) {
    // ===== epilogue warps: TMEM -> registers -> un-skewed staging tile -> TMA store =====
    const int q = wid & 3;               // TMEM lane quarter this warp may access
    const int wl = 32 * q + lane;        // output column within the tile (= TMEM lane)
    int it = 0, sbuf = 0;   // staging buffer of the current step (a running index: `gstep % n_bufs` was 10% of the kernel's instructions)
    for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x, ++it) {
      const int wt = tile % a.n_wtiles, h = (tile / a.n_wtiles) % a.H, n = tile / (a.n_wtiles * a.H);
      mbar_wait(tmem_full, (uint32_t)it & 1u);
      tc::fence_after_sync();
      // 32 output planes per step through two ping-pong staging buffers, one TMA store per step.  TMEM is read in
      // 32-column blocks that start at column delta, so that block b holds, for lane w = 32q+lane, the planes
      // p = 32(b-q) + jj - lane (jj = column inside the block): the planes [32s, 32s+32) of this warp's 32 output
      // columns are the upper triangle (jj >= lane) of block s+q and the lower triangle (jj < lane) of block s+q+1.
      // Every block is therefore loaded from TMEM exactly once and kept in registers for the following step
      // (TMEM reads, 64 B/clk, were the longest part of the drain when every block was fetched for two steps).
      float vp[32], vc[32];
      tc::tmem_ld32(tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)(32 * q + a.delta), vp);
      const int r0 = (32 - lane) & 31;
      for (int s = 0; s < a.n_steps; ++s, sbuf = (sbuf + 1 == a.n_bufs) ? 0 : sbuf + 1) {
        float* tile_s = reinterpret_cast<float*>(smem + a.tile_off + sbuf * kStepBytes);
        tc::tmem_ld32(tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)(32 * (s + q + 1) + a.delta), vc);
        if (s == a.n_steps - 1) {
          tc::fence_before_sync();
          __syncwarp();
          if (lane == 0) mbar_arrive(tmem_empty);  // TMEM drained: the next tile's MMAs may start
        }
        if (wid == 2 && lane == 0) {   // the store that last used this buffer has read it
          if (a.n_bufs == 3) tc::tma_store_wait_read<2>();
          else tc::tma_store_wait_read<1>();
        }
        named_bar_sync(1, 128);
        if (!PMT_DBG(a, 8)) {
          float* col = tile_s + wl;
#pragma unroll
          for (int jj = 0; jj < 32; ++jj) {
            const float v = (jj >= lane) ? vp[jj] : vc[jj];
            col[((r0 + jj) & 31) * kTM] = v;       // staging row = plane - 32s = (jj - lane) mod 32
          }
        }
#pragma unroll
        for (int jj = 0; jj < 32; ++jj) vp[jj] = vc[jj];
        fence_proxy_async();                       // generic-proxy writes -> visible to the TMA store
        named_bar_sync(1, 128);
        if (wid == 2 && lane == 0) {
          tc::tma_store_4d(&tmO, tile_s, wt * kTM, h, kRowsPerStep * s, n);
          tc::tma_store_commit();
        }
      }
    }
    if (wid == 2 && lane == 0) tc::tma_store_wait<0>();
  } else {
    // ===== transform warps (kPasses == 3): split staged fp32 into tf32 hi + lo =====
    // kXfGroups groups of 4 warps take the K chunks round-robin.  One group is the default: a second one was measured
    // at the headline shape and changed nothing (166.6 vs 167.6 us) -- the split is not what the MMA warp waits for.
    const int t = (tid - 6 * 32) & 127;      // 0..127 inside the group
    const int xg = (tid - 6 * 32) >> 7;      // group
    const int nchunks = nboxes * (kBoxBytes / 16);
    int st = xg % a.stages, ls = xg % a.lo_stages;
    uint32_t fph = (uint32_t)(xg / a.stages) & 1u, leph = ((uint32_t)(xg / a.lo_stages) & 1u) ^ 1u;
    int g = 0;
    for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
      for (int k = 0; k < a.n_cchunks; ++k, ++g) {
        if (g % kXfGroups != xg) continue;
        mbar_wait(&full[st], fph);
        mbar_wait(&lo_empty[ls], leph);
        const unsigned char* sbase = smem + (size_t)st * a.stage_bytes;
        unsigned char* lbase = smem + a.lo_ring_off + (size_t)ls * a.stage_bytes;
        // hi operand = the raw fp32 left in place (kind::tf32 ignores the low 13 mantissa bits, verified on B200:
        // the 3-term sum stays at ~1e-6); lo = x - trunc_tf32(x) is exact in fp32 and goes to the lo ring.
        // loads are issued in batches of 8 before the first store: the LDS latency is paid once per batch
        for (int cb = t; cb < nchunks && !PMT_DBG(a, 4); cb += 8 * 128) {
          float4 x[8];
#pragma unroll
          for (int u = 0; u < 8; ++u)
            if (cb + u * 128 < nchunks) x[u] = *reinterpret_cast<const float4*>(sbase + 16 * (cb + u * 128));
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            if (cb + u * 128 < nchunks) {
              float4 lo;
              lo.x = x[u].x - __uint_as_float(__float_as_uint(x[u].x) & 0xffffe000u);
              lo.y = x[u].y - __uint_as_float(__float_as_uint(x[u].y) & 0xffffe000u);
              lo.z = x[u].z - __uint_as_float(__float_as_uint(x[u].z) & 0xffffe000u);
              lo.w = x[u].w - __uint_as_float(__float_as_uint(x[u].w) & 0xffffe000u);
              *reinterpret_cast<float4*>(lbase + 16 * (cb + u * 128)) = lo;
            }
          }
        }
        fence_proxy_async();  // make the rewritten stage visible to the tensor core's async-proxy reads
        __syncwarp();
        if (lane == 0) mbar_arrive(&xf_done[ls]);
        st += kXfGroups;
        while (st >= a.stages) st -= a.stages, fph ^= 1u;
        ls += kXfGroups;
        while (ls >= a.lo_stages) ls -= a.lo_stages, leph ^= 1u;
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

// --------------------------------------------------------------------------------------------------------------------
// TMEM-A variant.  The SS kernel above is bound by shared-memory bandwidth (section 4.3 of DESIGN.md): every MMA
// re-reads the 128 x 8 L slab from shared memory (twice per pass: once per band half), and the L tile goes through the
// lo ring as well.  Here the L tile never feeds the tensor core from shared memory: TMA drops each 16-channel slab as a
// plain [16][128] block, four builder warps (thread = output column = TMEM lane, conflict-free LDS of its column) write
// it with tcgen05.st as hi (raw fp32: kind::tf32 ignores the low 13 mantissa bits) and lo = x - trunc_tf32(x) into a
// 4-slot A ring in TMEM columns [384, 512), and the MMAs take A from TMEM (TS form).  Per 128 x 192 tile that removes
// 192 KB of operand fetches and the L share of the lo ring (32 KB read + 32 KB written) from shared memory.
//   warp 0 TMA producer | warp 1 MMA issuer | warps 2-9 epilogue (two per TMEM lane quarter, each takes one half of
//   every 32-column block: the drain, serialised with the MMA phase on the single 320-column accumulator, was the longest
//   part of a tile with four warps) | warps 10-13 A builders | warps 14-15 R lo split (3xTF32)
// --------------------------------------------------------------------------------------------------------------------
constexpr int kTcaEpiWarps = 8;
constexpr int kTcaThreads3 = 32 * 16, kTcaThreads1 = 32 * 14;   // 16 warps: 128 registers per thread

template <int kPasses>
__global__ void __launch_bounds__(kPasses == 3 ? kTcaThreads3 : kTcaThreads1, 1)
corr1d_fwd_tca_kernel(const __grid_constant__ CUtensorMap tmL, const __grid_constant__ CUtensorMap tmR,
                      const __grid_constant__ CUtensorMap tmO, const TcFwdArgs a) {
  extern __shared__ __align__(1024) unsigned char smem[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + a.bar_off);   // R raw stage landed (TMA)
  uint64_t* empty = full + 8;                                       // R raw stage consumed (MMA commit)
  uint64_t* xf_done = empty + 8;                                    // R lo stage written (4 split warps)
  uint64_t* lo_empty = xf_done + 8;                                 // R lo stage consumed (MMA commit)
  uint64_t* l_full = lo_empty + 8;                                  // L slab landed (TMA)
  uint64_t* l_empty = l_full + 8;                                   // L slab copied to TMEM (4 builder warps)
  uint64_t* a_built = l_empty + 8;                                  // A slot written (4 builder warps)
  uint64_t* a_empty = a_built + kASlots;                            // A slot consumed (MMA commit)
  uint64_t* tmem_full = a_empty + kASlots;
  uint64_t* tmem_empty = tmem_full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 1);

  const int tid = threadIdx.x, wid = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    for (int s = 0; s < 8; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
      mbar_init(&xf_done[s], 2);
      mbar_init(&lo_empty[s], 1);
      mbar_init(&l_full[s], 1);
      mbar_init(&l_empty[s], 4);
    }
    for (int s = 0; s < kASlots; ++s) {
      mbar_init(&a_built[s], 4);
      mbar_init(&a_empty[s], 1);
    }
    mbar_init(tmem_full, 1);
    mbar_init(tmem_empty, kTcaEpiWarps);
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
    // ===== TMA producer: per 16-channel chunk NB band boxes (swizzled, B operand) + one plain L slab =====
    if (lane == 0) {
      tma_prefetch_desc(&tmL);
      tma_prefetch_desc(&tmR);
    }
    const uint32_t rbytes = (uint32_t)a.NB * kBoxBytes;
    int st = 0, ls = 0;
    uint32_t eph = 1, leph = 1;  // parities to wait for on empty[] / l_empty[] (the first pass over a ring is free)
    for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
      const int wt = tile % a.n_wtiles, h = (tile / a.n_wtiles) % a.H, n = tile / (a.n_wtiles * a.H);
      const int w0 = wt * kTM;
      for (int k = 0; k < a.n_cchunks; ++k) {
        mbar_wait(&l_empty[ls], leph);
        if (lane == a.NB) {
          mbar_arrive_expect_tx(&l_full[ls], (uint32_t)kLStageBytes);
          tma_load_4d(smem + a.l_ring_off + ls * kLStageBytes, &tmL, w0, h, k * kCK, n, &l_full[ls]);
        }
        mbar_wait(&empty[st], eph);
        unsigned char* sbase = smem + (size_t)st * a.stage_bytes;
        if (lane == 0) mbar_arrive_expect_tx(&full[st], rbytes);
        __syncwarp();
        if (lane < a.NB)
          tma_load_4d(sbase + lane * kBoxBytes, &tmR, w0 - a.rW - a.delta + 32 * lane, h, k * kCK, n, &full[st]);
        if (++st == a.stages) st = 0, eph ^= 1u;
        if (++ls == a.l_stages) ls = 0, leph ^= 1u;
      }
    }
  } else if (wid == 1) {
    // ===== MMA issuer (warp converged, one elected lane issues): A from TMEM, B = R band (MN-major) from shared memory =====
    {
      const uint32_t idesc1 = tc::make_idesc(2, 0, 1, kTM, a.N1);
      const uint32_t idesc2 = tc::make_idesc(2, 0, 1, kTM, a.N2 > 0 ? a.N2 : 16);
      const uint64_t d0 = mn_desc(smem_u32(smem), kBoxBytes, 1024);                       // raw ring (hi operand)
      const uint64_t dl0 = mn_desc(smem_u32(smem + a.lo_ring_off), kBoxBytes, 1024);      // lo ring
      const uint32_t st_step = (uint32_t)a.stage_bytes >> 4;
      const uint32_t offB2 = (uint32_t)((a.N1 / 32) * kBoxBytes) >> 4;
      const bool two = a.N2 > 0;
      int st = 0, ls = 0, as = 0, it = 0;
      uint32_t fph = 0, lph = 0, aph = 0;
      for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x, ++it) {
        mbar_wait(tmem_empty, ((uint32_t)it & 1u) ^ 1u);  // epilogue of the previous tile has drained TMEM
        tc::fence_after_sync();
        for (int k = 0; k < a.n_cchunks; ++k) {
          if (kPasses == 3) mbar_wait(&xf_done[ls], lph);
          else mbar_wait(&full[st], fph);
          mbar_wait(&a_built[as], aph);
          tc::fence_after_sync();
          const uint64_t dh = d0 + (uint64_t)(st_step * (uint32_t)st);
          const uint64_t dl = dl0 + (uint64_t)(st_step * (uint32_t)ls);
          const uint32_t ta = tmem_base + (uint32_t)(a.a_base + as * kASlotCols);
          if (tc::elect_one()) {
#pragma unroll
            for (int kk = 0; kk < kCK / 8; ++kk) {
              const uint32_t acc = (k > 0 || kk > 0) ? 1u : 0u;
              const uint32_t ko = (uint32_t)(kk * 1024) >> 4;
              const uint32_t ah = ta + 8 * kk, al = ta + kCK + 8 * kk;
              if (kPasses == 3) {
                // small cross terms first, then the dominant hi*hi term
                tc::mma_tf32_ts(tmem_base, al, dh + ko, idesc1, acc);
                tc::mma_tf32_ts(tmem_base, ah, dl + ko, idesc1, 1u);
                tc::mma_tf32_ts(tmem_base, ah, dh + ko, idesc1, 1u);
                if (two) {
                  tc::mma_tf32_ts(tmem_base + a.N1, al, dh + offB2 + ko, idesc2, acc);
                  tc::mma_tf32_ts(tmem_base + a.N1, ah, dl + offB2 + ko, idesc2, 1u);
                  tc::mma_tf32_ts(tmem_base + a.N1, ah, dh + offB2 + ko, idesc2, 1u);
                }
              } else {
                tc::mma_tf32_ts(tmem_base, ah, dh + ko, idesc1, acc);
                if (two) tc::mma_tf32_ts(tmem_base + a.N1, ah, dh + offB2 + ko, idesc2, acc);
              }
            }
            tc::mma_commit(&empty[st]);  // ring slots reusable once these MMAs have read them
            if (kPasses == 3) tc::mma_commit(&lo_empty[ls]);
            tc::mma_commit(&a_empty[as]);
            if (k == a.n_cchunks - 1) tc::mma_commit(tmem_full);     // accumulator of this tile complete
          }
          __syncwarp();
          if (++st == a.stages) st = 0, fph ^= 1u;
          if (++ls == a.lo_stages) ls = 0, lph ^= 1u;
          if (++as == kASlots) as = 0, aph ^= 1u;
        }
      }
    }
  } else if (wid < 2 + kTcaEpiWarps) {
    // ===== epilogue warps: TMEM -> registers -> un-skewed staging tile -> TMA store =====
    // Block b (32 TMEM columns from column delta + 32 b) holds for lane w = 32q+lane the planes 32(b-q) + jj - lane, so
    // the planes [32s, 32s+32) of this warp's 32 output columns are the upper triangle (jj >= lane) of block s+q and the
    // lower triangle of block s+q+1: every block is read from TMEM once and kept for the next step.  Two warps share a
    // lane quarter: warp `hsel` owns columns jj in [16 hsel, 16 hsel + 16) of every block.
    const int q = wid & 3;               // TMEM lane quarter this warp may access
    const int hsel = (wid - 2) >> 2;     // which half of every block
    const int wl = 32 * q + lane;        // output column within the tile (= TMEM lane)
    int it = 0, sbuf = 0;   // staging buffer of the current step (a running index: `gstep % n_bufs` was 10% of the kernel's instructions)
    for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x, ++it) {
      const int wt = tile % a.n_wtiles, h = (tile / a.n_wtiles) % a.H, n = tile / (a.n_wtiles * a.H);
      mbar_wait(tmem_full, (uint32_t)it & 1u);
      tc::fence_after_sync();
      // software pipeline over the TMEM loads: block s+q+2 is requested before the staging stores of step s, so its
      // latency is hidden behind them (vp = block s+q, vc = block s+q+1, vn = block s+q+2 in flight)
      uint32_t vp[16], vc[16], vn[16];
      const uint32_t tcol = tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)(a.delta + 16 * hsel);
      tc::tmem_ld16_async(tcol + (uint32_t)(32 * q), vp);
      tc::tmem_ld16_async(tcol + (uint32_t)(32 * (q + 1)), vc);
      tc::tmem_ld_wait16(vp);
      tc::tmem_ld_wait16(vc);
      const int r0 = (32 - lane + 16 * hsel) & 31;
      // One step: P = block s+q, C = block s+q+1, N = block s+q+2 (requested here).  The three register sets change roles
      // from step to step by NAME (the loop below is unrolled by three), not by copying 32 registers per step.
      auto step = [&](int s, uint32_t (&P)[16], uint32_t (&C)[16], uint32_t (&N)[16]) {
        float* tile_s = reinterpret_cast<float*>(smem + a.tile_off + sbuf * kStepBytes);
        sbuf = (sbuf + 1 == a.n_bufs) ? 0 : sbuf + 1;
        if (s + 1 < a.n_steps) {
          tc::tmem_ld16_async(tcol + (uint32_t)(32 * (s + q + 2)), N);
        } else {
          tc::fence_before_sync();
          __syncwarp();
          if (lane == 0) mbar_arrive(tmem_empty);  // every block of this tile has been read: the next tile's MMAs may start
        }
        if (wid == 2 && lane == 0) {   // the store that last used this buffer has read it
          if (a.n_bufs == 3) tc::tma_store_wait_read<2>();
          else tc::tma_store_wait_read<1>();
        }
        named_bar_sync(1, 32 * kTcaEpiWarps);
        float* col = tile_s + wl;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const int jj = 16 * hsel + j;
          const uint32_t v = (jj >= lane) ? P[j] : C[j];
          col[((r0 + j) & 31) * kTM] = __uint_as_float(v);   // staging row = plane - 32s = (jj - lane) mod 32
        }
        fence_proxy_async();                       // generic-proxy writes -> visible to the TMA store
        named_bar_sync(1, 32 * kTcaEpiWarps);
        if (wid == 2 && lane == 0) {
          tc::tma_store_4d(&tmO, tile_s, wt * kTM, h, kRowsPerStep * s, n);
          tc::tma_store_commit();
        }
        if (s + 1 < a.n_steps) tc::tmem_ld_wait16(N);   // N has landed (its latency overlapped the stores above)
      };
      for (int s = 0;;) {
        step(s, vp, vc, vn);
        if (++s >= a.n_steps) break;
        step(s, vc, vn, vp);
        if (++s >= a.n_steps) break;
        step(s, vn, vp, vc);
        if (++s >= a.n_steps) break;
      }
    }
    if (wid == 2 && lane == 0) tc::tma_store_wait<0>();
  } else if (wid < 2 + kTcaEpiWarps + 4) {
    // ===== A builders: L slab (shared memory, [16 ch][128 w]) -> TMEM A slot, hi = raw fp32, lo = x - trunc_tf32(x) =====
    const int q = wid & 3;
    const int xl = 32 * q + lane;        // output column = TMEM lane = A row
    int ls = 0, as = 0;
    uint32_t lph = 0, aeph = 1;
    for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
      for (int k = 0; k < a.n_cchunks; ++k) {
        mbar_wait(&l_full[ls], lph);
        const float* Ls = reinterpret_cast<const float*>(smem + a.l_ring_off + ls * kLStageBytes) + xl;
        float v[kCK];
#pragma unroll
        for (int c = 0; c < kCK; ++c) v[c] = Ls[c * kTM];   // lanes read consecutive words: conflict-free
        mbar_wait(&a_empty[as], aeph);
        tc::fence_after_sync();
        const uint32_t ta = tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)(a.a_base + as * kASlotCols);
        tc::tmem_st16(ta, v);
        if (kPasses == 3) {
#pragma unroll
          for (int c = 0; c < kCK; ++c) v[c] = v[c] - __uint_as_float(__float_as_uint(v[c]) & 0xffffe000u);
          tc::tmem_st16(ta + kCK, v);
        }
        tc::tmem_st_wait();
        tc::fence_before_sync();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(&a_built[as]);
          mbar_arrive(&l_empty[ls]);
        }
        if (++ls == a.l_stages) ls = 0, lph ^= 1u;
        if (++as == kASlots) as = 0, aeph ^= 1u;
      }
    }
  } else {
    // ===== R lo split (kPasses == 3): lo = x - trunc_tf32(x) of the landed band stage into the lo ring =====
    constexpr int kXfThreads = 64;
    const int t = tid - 32 * (2 + kTcaEpiWarps + 4);   // 0..63
    const int nchunks = a.NB * (kBoxBytes / 16);
    int st = 0, ls = 0;
    uint32_t fph = 0, leph = 1;
    for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
      for (int k = 0; k < a.n_cchunks; ++k) {
        mbar_wait(&full[st], fph);
        mbar_wait(&lo_empty[ls], leph);
        const unsigned char* sbase = smem + (size_t)st * a.stage_bytes;
        unsigned char* lbase = smem + a.lo_ring_off + (size_t)ls * a.stage_bytes;
        for (int cb = t; cb < nchunks; cb += 8 * kXfThreads) {
          float4 x[8];
#pragma unroll
          for (int u = 0; u < 8; ++u)
            if (cb + u * kXfThreads < nchunks) x[u] = *reinterpret_cast<const float4*>(sbase + 16 * (cb + u * kXfThreads));
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            if (cb + u * kXfThreads < nchunks) {
              float4 lo;
              lo.x = x[u].x - __uint_as_float(__float_as_uint(x[u].x) & 0xffffe000u);
              lo.y = x[u].y - __uint_as_float(__float_as_uint(x[u].y) & 0xffffe000u);
              lo.z = x[u].z - __uint_as_float(__float_as_uint(x[u].z) & 0xffffe000u);
              lo.w = x[u].w - __uint_as_float(__float_as_uint(x[u].w) & 0xffffe000u);
              *reinterpret_cast<float4*>(lbase + 16 * (cb + u * kXfThreads)) = lo;
            }
          }
        }
        fence_proxy_async();  // make the lo stage visible to the tensor core's async-proxy reads
        __syncwarp();
        if (lane == 0) mbar_arrive(&xf_done[ls]);
        if (++st == a.stages) st = 0, fph ^= 1u;
        if (++ls == a.lo_stages) ls = 0, leph ^= 1u;
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

int fill_args(TcFwdArgs* a, int C, int H, int W, int P, int passes) {
  a->C = C, a->H = H, a->W = W, a->P = P, a->rW = (P - 1) / 2;
  a->delta = ((-a->rW % 4) + 4) % 4;
  a->NB = ceil_div(kTM + P - 1 + a->delta, 32);
  if (a->NB > kMaxNB) return 1;
  const int nb1 = (a->NB + 1) / 2;
  a->N1 = 32 * nb1;
  a->N2 = 32 * (a->NB - nb1);
  a->n_steps = ceil_div(P, kRowsPerStep);
  int cols = 32;
  const int last_col = 32 * (a->n_steps + 4) + a->delta + 32;   // the epilogue reads blocks 0 .. n_steps+3, shifted by delta
  while (cols < 32 * a->NB || cols < last_col) cols *= 2;
  if (cols > 512) return 1;
  a->tmem_cols = cols;
  a->n_wtiles = ceil_div(W, kTM);
  a->n_cchunks = ceil_div(C, kCK);
  a->n_tiles = 0;  // set by the launcher (needs B)
  a->stage_bytes = (kLBlocks + a->NB) * kBoxBytes;
  a->n_bufs = 2;
  if (PMT_ENV_INT("PMT_FWD_NBUF", 2) == 3) a->n_bufs = 3;
  const int budget = 227 * 1024 - 1024 - a->n_bufs * kStepBytes;
  int total = budget / a->stage_bytes;  // ring stages that fit next to the two staging buffers
  if (passes == 3) {
    a->lo_stages = total >= 6 ? 2 : 1;   // measured at the headline shape (7 stages): 5 raw + 2 lo beats 4 + 3
    if (const int e = PMT_ENV_INT("PMT_FWD_LO_STAGES", 0)) a->lo_stages = e;
    if (a->lo_stages < kXfGroups) return 1;
    a->stages = total - a->lo_stages;
  } else {
    a->lo_stages = 1;
    a->stages = total;
  }
  if (a->stages > 8) a->stages = 8;
  if (a->stages < 1) return 1;
  a->lo_ring_off = a->stages * a->stage_bytes;
  a->tile_off = a->lo_ring_off + (passes == 3 ? a->lo_stages * a->stage_bytes : 0);
  a->bar_off = a->tile_off + a->n_bufs * kStepBytes;
  a->debug = PMT_ENV_INT("PMT_TC_DEBUG", 0);
  return 0;
}

// Ring layout of the TMEM-A variant (fills the fields the kernel above uses on top of fill_args()).
int fill_args_tca(TcFwdArgs* a, int passes) {
  if (32 * a->NB > 384) return 1;            // accumulator must stay below the A ring
  a->a_base = 384;
  a->tmem_cols = 512;
  a->stage_bytes = a->NB * kBoxBytes;        // R band only
  a->n_bufs = 2;
  a->l_stages = 3;
  const int budget = 227 * 1024 - 1024 - a->n_bufs * kStepBytes - a->l_stages * kLStageBytes;
  int total = budget / a->stage_bytes;
  if (passes == 3) {
    a->lo_stages = total >= 8 ? 3 : (total >= 5 ? 2 : 1);
    if (const int e = PMT_ENV_INT("PMT_FWD_LO_STAGES", 0)) a->lo_stages = e;
    if (a->lo_stages < 1 || a->lo_stages > 8) return 1;
    a->stages = total - a->lo_stages;
  } else {
    a->lo_stages = 1;   // unused, keeps the ring arithmetic defined
    a->stages = total;
  }
  if (a->stages > 8) a->stages = 8;
  if (a->stages < 1) return 1;
  a->lo_ring_off = a->stages * a->stage_bytes;
  a->l_ring_off = a->lo_ring_off + (passes == 3 ? a->lo_stages * a->stage_bytes : 0);
  a->tile_off = a->l_ring_off + a->l_stages * kLStageBytes;
  a->bar_off = a->tile_off + a->n_bufs * kStepBytes;
  return 0;
}

}  // namespace

int make_tmap_nchw_ex(CUtensorMap* map, const float* base, int B, int C, int H, int W, int box_w, int box_c,
                      int swizzle128);

bool corr1d_fwd_tc_ok(const void* in1, const void* in2, const void* out, int C, int H, int W, int P, int dilp,
                      int passes) {
  if (dilp != 1 || P < 1 || C < 1 || W % 4 != 0 || !aligned16(in1) || !aligned16(in2) || !aligned16(out)) return false;
  TcFwdArgs a;
  (void)H;
  return fill_args(&a, C, H, W, P, passes) == 0;
}

int launch_corr1d_fwd_tc(const float* in1, const float* in2, float* out, int B, int C, int H, int W, int P,
                         int passes, cudaStream_t st) {
  TcFwdArgs a;
  PMT_CHECK_ARG(passes == 1 || passes == 3, "corr1d tc: passes must be 1 (tf32) or 3 (3xtf32)");
  PMT_CHECK_ARG(fill_args(&a, C, H, W, P, passes) == 0, "corr1d tc: unsupported shape P=%d", P);
  CUtensorMap tmL, tmR, tmO;
  const bool tmem_a = PMT_ENV_INT("PMT_FWD_SS", 0) == 0 && fill_args_tca(&a, passes) == 0;
  if (tmem_a) {
    // L tile as the A operand in TMEM (default): plain [16][128] slabs for the builders, swizzled boxes for the R band
    if (int e = make_tmap_nchw_ex(&tmL, in1, B, C, H, W, kTM, kCK, 0)) return e;
    if (int e = make_tmap_nchw_ex(&tmR, in2, B, C, H, W, 32, kCK, 2)) return e;
    if (int e = make_tmap_nchw_ex(&tmO, out, B, P, H, W, kTM, kRowsPerStep, 0)) return e;
    const int smem_bytes = a.bar_off + 512;
    const int64_t tiles = (int64_t)B * H * a.n_wtiles;
    PMT_CHECK_ARG(tiles < (1ll << 31), "corr1d tc: too many tiles");
    a.n_tiles = (int)tiles;
    const int64_t grid = tiles < sm_count() ? tiles : sm_count();  // persistent: one CTA per SM
    if (passes == 3) {
      PMT_CUDA_OK(cudaFuncSetAttribute(corr1d_fwd_tca_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
      corr1d_fwd_tca_kernel<3><<<(unsigned)grid, kTcaThreads3, smem_bytes, st>>>(tmL, tmR, tmO, a);
    } else {
      PMT_CUDA_OK(cudaFuncSetAttribute(corr1d_fwd_tca_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
      corr1d_fwd_tca_kernel<1><<<(unsigned)grid, kTcaThreads1, smem_bytes, st>>>(tmL, tmR, tmO, a);
    }
    PMT_LAUNCH_OK("corr1d_fwd_tca_kernel");
    return PMT_OK;
  }
  PMT_CHECK_ARG(fill_args(&a, C, H, W, P, passes) == 0, "corr1d tc: unsupported shape P=%d", P);   // SS layout again
  if (int e = make_tmap_nchw_ex(&tmL, in1, B, C, H, W, 32, kCK, 2)) return e;
  if (int e = make_tmap_nchw_ex(&tmR, in2, B, C, H, W, 32, kCK, 2)) return e;
  if (int e = make_tmap_nchw_ex(&tmO, out, B, P, H, W, kTM, kRowsPerStep, 0)) return e;
  const int smem_bytes = a.bar_off + 512;
  const int64_t tiles = (int64_t)B * H * a.n_wtiles;
  PMT_CHECK_ARG(tiles < (1ll << 31), "corr1d tc: too many tiles");
  a.n_tiles = (int)tiles;
  const int64_t grid = tiles < sm_count() ? tiles : sm_count();  // persistent: one CTA per SM
  if (passes == 3) {
    PMT_CUDA_OK(cudaFuncSetAttribute(corr1d_fwd_tc_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    corr1d_fwd_tc_kernel<3><<<(unsigned)grid, kFwdThreads3, smem_bytes, st>>>(tmL, tmR, tmO, a);
  } else {
    PMT_CUDA_OK(cudaFuncSetAttribute(corr1d_fwd_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    corr1d_fwd_tc_kernel<1><<<(unsigned)grid, 192, smem_bytes, st>>>(tmL, tmR, tmO, a);
  }
  PMT_LAUNCH_OK("corr1d_fwd_tc_kernel");
  return PMT_OK;
}

}  // namespace pmt
